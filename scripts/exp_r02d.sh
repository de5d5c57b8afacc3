set -x
run() { tag=$1; shift; env "$@" python bench.py --no-secondary --no-cpu-baseline --steps 96 > gpurun_out/r02d_$tag.json 2> gpurun_out/r02d_$tag.err; python - <<PY
import json
d=json.load(open("gpurun_out/r02d_$tag.json"))
print("$tag", "ms/step %.4f" % d["ms_per_step"], "frac %.3f" % d["roofline"]["frac"], "single %.3f" % d["config"]["single_stream_ms_per_step"], "errs", d["config"]["error_rows"], d["kernel_ms"])
PY
}
for v in b1 b2 b3 b4; do
  P2E_LIB=$PWD/scratch/lib_$v.so python -m pytest tests/test_gpu_f8.py tests/test_gpu_chain.py -m gpu -q -x 2>&1 | tail -n 2
  run ${v}_split1 P2E_LIB=$PWD/scratch/lib_$v.so
  run ${v}_split0 P2E_LIB=$PWD/scratch/lib_$v.so P2E_OPT_CHAIN_SPLIT=0
done

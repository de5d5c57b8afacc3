set -x
python -m pytest tests/test_c_host.py tests/test_ntt.py tests/test_gpu_chain.py -m gpu -q > gpurun_out/r02b_pytest.log 2>&1; tail -n 5 gpurun_out/r02b_pytest.log
run() { tag=$1; shift; env "$@" python bench.py --no-secondary --no-cpu-baseline --steps 96 > gpurun_out/r02b_$tag.json 2> gpurun_out/r02b_$tag.err; python - <<PY
import json
d=json.load(open("gpurun_out/r02b_$tag.json"))
print("$tag", "ms/step %.4f" % d["ms_per_step"], "frac %.3f" % d["roofline"]["frac"], "single %.3f" % d["config"]["single_stream_ms_per_step"], d["kernel_ms"])
PY
}
for s in 8 12 16 24 32; do run base_s$s P2E_BENCH_STREAMS=$s; done
for v in a1 a2 a3; do for s in 8 16; do run ${v}_s$s P2E_LIB=$PWD/scratch/lib_$v.so P2E_BENCH_STREAMS=$s; done; done

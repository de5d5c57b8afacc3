set -x
python -m pytest tests/test_genstream_ecdsa.py -m gpu -q 2>&1 | tail -n 3
run() { tag=$1; shift; env "$@" python bench.py --no-secondary --no-cpu-baseline --steps 96 > gpurun_out/r02l_$tag.json 2> gpurun_out/r02l_$tag.err; python - <<PY
import json
d=json.load(open("gpurun_out/r02l_$tag.json"))
print("$tag", "ms/step %.4f" % d["ms_per_step"], "frac %.3f" % d["roofline"]["frac"], "single %.3f" % d["config"]["single_stream_ms_per_step"], "streams", d["config"]["streams"])
PY
}
run conn32
run conn8 CUDA_DEVICE_MAX_CONNECTIONS=8
run conn32_s16 P2E_BENCH_STREAMS=16
run conn32_s4 P2E_BENCH_STREAMS=4

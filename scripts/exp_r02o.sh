set -x
python -m pytest tests/test_gpu_chain.py tests/test_builder.py tests/test_c_host.py tests/test_genstream_ecdsa.py -m gpu -q 2>&1 | tail -n 6
run() { tag=$1; shift; env "$@" python bench.py --no-secondary --no-cpu-baseline --steps 128 > gpurun_out/r02o_$tag.json 2> gpurun_out/r02o_$tag.err; python - <<PY
import json
d=json.load(open("gpurun_out/r02o_$tag.json"))
print("$tag", "ms/step %.4f" % d["ms_per_step"], "frac %.3f" % d["roofline"]["frac"], "single %.3f" % d["config"]["single_stream_ms_per_step"], "streams", d["config"]["streams"], d["kernel_ms"])
PY
}
run s16
run s24 P2E_BENCH_STREAMS=24
python scripts/chain_phases.py 16 > gpurun_out/r02o_phases.txt 2>&1; cat gpurun_out/r02o_phases.txt

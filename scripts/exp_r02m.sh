set -x
python -m pytest tests/test_genstream_ecdsa.py -m gpu -q 2>&1 | tail -n 15
run() { tag=$1; shift; env "$@" python bench.py --no-secondary --no-cpu-baseline --steps 128 > gpurun_out/r02m_$tag.json 2> gpurun_out/r02m_$tag.err; python - <<PY
import json
d=json.load(open("gpurun_out/r02m_$tag.json"))
print("$tag", "ms/step %.4f" % d["ms_per_step"], "frac %.3f" % d["roofline"]["frac"], "single %.3f" % d["config"]["single_stream_ms_per_step"], "streams", d["config"]["streams"], d["kernel_ms"]["k_prep"])
PY
}
run s16 P2E_BENCH_STREAMS=16
run s24 P2E_BENCH_STREAMS=24
run s32 P2E_BENCH_STREAMS=32
run s48 P2E_BENCH_STREAMS=48
run s16_nosplit P2E_BENCH_STREAMS=16 P2E_OPT_CHAIN_SPLIT=0
run s32_nosplit P2E_BENCH_STREAMS=32 P2E_OPT_CHAIN_SPLIT=0

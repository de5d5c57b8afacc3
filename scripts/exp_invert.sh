run() { name=$1; shift; env "$@" python bench.py --no-secondary --no-cpu-baseline --steps 40 > gpurun_out/x_$name.json 2> gpurun_out/x_$name.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/x_$name.json"))
    print("$name", round(d["ms_per_step"],3), round(d["config"]["single_stream_ms_per_step"],3), d["kernel_ms"])
except Exception as e:
    print("$name ERR", e)
PY
}
python -m pytest tests/test_gpu_chain.py tests/test_builder.py -m gpu -x -q 2>&1 | tail -3
run tree16_s4 A=1
run tree16_s8 P2E_BENCH_STREAMS=8
run tree8_s8 P2E_BENCH_STREAMS=8 P2E_CHAIN_INV_GROUPS=8
run tree32_s8 P2E_BENCH_STREAMS=8 P2E_CHAIN_INV_GROUPS=32
run groups_s8 P2E_BENCH_STREAMS=8 P2E_CHAIN_INVERT=groups
run tree16_s12 P2E_BENCH_STREAMS=12

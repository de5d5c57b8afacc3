run() { name=$1; shift; env "$@" python bench.py --no-secondary --no-cpu-baseline --steps 40 > gpurun_out/x_$name.json 2> gpurun_out/x_$name.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/x_$name.json"))
    print("$name", round(d["ms_per_step"],3), round(d["config"]["single_stream_ms_per_step"],3))
except Exception as e:
    print("$name ERR", e)
PY
}
run s8 A=1
run s8b A=1
run s4 P2E_BENCH_STREAMS=4
run s6 P2E_BENCH_STREAMS=6
run s12 P2E_BENCH_STREAMS=12
run s16 P2E_BENCH_STREAMS=16
run s8_tpb64 P2E_CHAIN_WIDE_TPB=64
run s8_serial P2E_CHAIN_REPLAY=serial
run s16_serial P2E_CHAIN_REPLAY=serial P2E_BENCH_STREAMS=16
run s8_nopri P2E_CHAIN_NO_PRIORITY=1
run s8_groups P2E_CHAIN_INVERT=groups
run s8_g8 P2E_CHAIN_INV_GROUPS=8
run s8_g32 P2E_CHAIN_INV_GROUPS=32

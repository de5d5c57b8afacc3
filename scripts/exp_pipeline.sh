run() { name=$1; shift; env "$@" python bench.py --no-secondary --no-cpu-baseline --steps 40 > gpurun_out/x_$name.json 2> gpurun_out/x_$name.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/x_$name.json"))
    print("$name", round(d["ms_per_step"],3), round(d["config"]["single_stream_ms_per_step"],3), d["kernel_ms"])
except Exception as e:
    print("$name ERR", e)
PY
}
run base A=1
run tpb64 P2E_CHAIN_WIDE_TPB=64
run tpb32 P2E_CHAIN_WIDE_TPB=32
run serial4 P2E_CHAIN_REPLAY=serial
run serial8 P2E_CHAIN_REPLAY=serial P2E_BENCH_STREAMS=8
run serial8_64 P2E_CHAIN_REPLAY=serial P2E_BENCH_STREAMS=8 P2E_CHAIN_WIDE_TPB=64
run nopri P2E_CHAIN_NO_PRIORITY=1
run s8_64 P2E_BENCH_STREAMS=8 P2E_CHAIN_WIDE_TPB=64
run s3 P2E_BENCH_STREAMS=3

set -x
run() { tag=$1; shift; env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 96 --warmup 3 --no-secondary > gpurun_out/r02i_$tag.json 2> gpurun_out/r02i_$tag.err; python - <<PY
import json
d=json.load(open("gpurun_out/r02i_$tag.json"))
print("$tag", "ms/step %.4f" % d["ms_per_step"], "value %.3fM" % (d["value"]/1e6), "streams", d["config"]["streams"], "e2e %.0f" % d["e2e"]["value"])
PY
}
run base
run conn32 CUDA_DEVICE_MAX_CONNECTIONS=32
run nosplit P2E_OPT_CHAIN_SPLIT=0
run nosplit_conn32 P2E_OPT_CHAIN_SPLIT=0 CUDA_DEVICE_MAX_CONNECTIONS=32
run s8 P2E_BENCH_STREAMS=8
run s32_conn32 P2E_BENCH_STREAMS=32 CUDA_DEVICE_MAX_CONNECTIONS=32

set -x
N=$1
run() { tag=$1; shift; env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 96 --warmup 3 --no-secondary > gpurun_out/r02j_n${N}_$tag.json 2> gpurun_out/r02j_n${N}_$tag.err; python - <<PY
import json
d=json.load(open("gpurun_out/r02j_n${N}_$tag.json"))
print("n$N $tag", "ms/step %.4f" % d["ms_per_step"], "value %.3fM" % (d["value"]/1e6), "streams", d["config"]["streams"], "e2e %.0f" % d["e2e"]["value"], "gather", d.get("gather",{}).get("ms_gather_alone"), "d2h", d.get("d2h_probe",{}).get("aggregate_GBps"))
PY
}
run base
run s32 P2E_BENCH_STREAMS=32

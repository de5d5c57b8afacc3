# one strong-scaling bench line at N GPUs (the driver's own launch): scripts/scale_run.sh N TAG STEPS [ENV=VALUE ...]  (under gpurun --gpus N)
N=$1; TAG=${2:-r02s}; STEPS=${3:-96}; shift; shift; shift
env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps $STEPS --warmup 5 --no-secondary > gpurun_out/${TAG}_n${N}.json 2> gpurun_out/${TAG}_n${N}.err
rc=$?
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_n${N}.json"))
print("n$N ${TAG} steps $STEPS", "ms/step %.4f" % d["ms_per_step"], "value %.3fM" % (d["value"]/1e6), "streams", d["config"]["streams"], "e2e %.0f" % d["e2e"]["value"], "gather", d.get("gather",{}).get("ms_gather_alone"), "d2h", d.get("d2h_probe",{}).get("aggregate_GBps"), "weak", (d.get("weak_scaling") or {}).get("value"))
PY
exit $rc

for s in 8 12 15; do
  P2E_BENCH_STREAMS=$s python bench.py --no-secondary --no-cpu-baseline --steps 60 > gpurun_out/x_streams_$s.json 2> gpurun_out/x_streams_$s.err
  python - <<P
import json
d=json.load(open("gpurun_out/x_streams_$s.json"))
print("streams $s", round(d["ms_per_step"],4), round(d["value"]))
P
done

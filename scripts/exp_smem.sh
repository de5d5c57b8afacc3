for cfg in "0 8" "81920 8" "122880 8" "81920 12" "58000 8"; do
  set -- $cfg
  P2E_CHAIN_WIDE_SMEM=$1 P2E_BENCH_STREAMS=$2 python bench.py --no-secondary --no-cpu-baseline --steps 40 > gpurun_out/x_smem_$1_$2.json 2> gpurun_out/x_smem_$1_$2.err
  python - <<P
import json
d=json.load(open("gpurun_out/x_smem_$1_$2.json"))
print("smem $1 streams $2", round(d["ms_per_step"],4), round(d["value"]), d["config"]["single_stream_ms_per_step"], d["kernel_ms"])
P
done

# round-1 (session 3) evidence: bench lines, reference arm, launch list, ncu summaries
set -x
python bench.py --steps 20 --warmup 3 > gpurun_out/r01c_bench_n1.json 2> gpurun_out/r01c_bench_n1.err
python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/r01c_bench_reference_n1.json 2> gpurun_out/r01c_bench_reference_n1.err
for w in eval evalfc evalcs evalcsfc evalrc; do
  P2E_BENCH_LOG_ROWS=23 python bench.py --workload $w --no-secondary --no-cpu-baseline --steps 10 > gpurun_out/r01c_bench_${w}_2p23.json 2> gpurun_out/r01c_bench_${w}_2p23.err
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r01c_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-secondary --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_prep|k_chain' -s 11 -c 11 -o gpurun_out/r01c_prof_chain -f \
    python scripts/profile_once.py ecdsa > gpurun_out/ncu_chain_full.log 2>&1
tail -n 3 gpurun_out/ncu_chain_full.log
head -c 400 gpurun_out/r01c_bench_n1.json

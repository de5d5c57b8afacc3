# round-1 evidence: bench lines, reference arm, ncu launch list, ncu --set full summaries
set -x
python bench.py --steps 20 --warmup 3 > gpurun_out/r01b_bench_n1.json 2> gpurun_out/r01b_bench_n1.err
python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/r01b_bench_reference_n1.json 2> gpurun_out/r01b_bench_reference_n1.err
python scripts/profile_once.py ecdsa > gpurun_out/po_ecdsa.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01b_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-secondary --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_prep|k_chain' -s 11 -c 11 -o gpurun_out/r01b_prof_chain -f \
    python scripts/profile_once.py ecdsa > gpurun_out/ncu_chain_full.log 2>&1
python scripts/profile_once.py ux > gpurun_out/po_ux.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_eval_ux|k_ux_' -c 24 -o gpurun_out/r01b_prof_ux -f \
    python scripts/profile_once.py ux > gpurun_out/ncu_ux_full.log 2>&1
tail -2 gpurun_out/po_ecdsa.log gpurun_out/po_ux.log gpurun_out/ncu_chain_full.log gpurun_out/ncu_ux_full.log
head -c 600 gpurun_out/r01b_bench_n1.json

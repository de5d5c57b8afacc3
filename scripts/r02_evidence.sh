# round-2 evidence: tests, bench lines, reference arm, launch list, ncu summaries (run under gpurun, one GPU)
set -x
TAG=${1:-r02n}
python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; tail -n 3 gpurun_out/${TAG}_pytest_gpu.log
python bench.py --steps 64 --warmup 3 > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err
python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_reference_n1.json 2> gpurun_out/${TAG}_bench_reference_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-secondary --no-cpu-baseline > gpurun_out/${TAG}_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_prep|k_chain' -s 9 -c 9 -o gpurun_out/${TAG}_prof_chain -f \
    python scripts/profile_once.py ecdsa > gpurun_out/${TAG}_ncu_chain_full.log 2>&1
tail -n 3 gpurun_out/${TAG}_ncu_chain_full.log
head -c 600 gpurun_out/${TAG}_bench_n1.json

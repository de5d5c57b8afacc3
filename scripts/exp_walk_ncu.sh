P2E_OPT_GENSTREAM_WALK_SHAPE=832 ncu --set full --clock-control none --import-source on -k regex:'k_gs_walk' -s 2 -c 1 -o gpurun_out/walk_prof -f \
  python bench.py --workload genstream_ecdsa --no-secondary --no-cpu-baseline --steps 2 --warmup 3 > gpurun_out/walk_ncu.log 2>&1
tail -n 3 gpurun_out/walk_ncu.log
ls -la gpurun_out/walk_prof.ncu-rep

set -x
python -m pytest tests/test_genstream_ecdsa.py tests/test_genstream.py tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -n 8
python bench.py --workload genstream_ecdsa --no-secondary --no-cpu-baseline --steps 3 > gpurun_out/r02k_gs_ecdsa.json 2> gpurun_out/r02k_gs_ecdsa.err; tail -3 gpurun_out/r02k_gs_ecdsa.err; head -c 900 gpurun_out/r02k_gs_ecdsa.json; echo
python bench.py --workload genstream --no-secondary --no-cpu-baseline --steps 10 > gpurun_out/r02k_gs.json 2> gpurun_out/r02k_gs.err; head -c 300 gpurun_out/r02k_gs.json; echo

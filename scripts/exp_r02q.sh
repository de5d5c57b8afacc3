python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 96 --warmup 16 > gpurun_out/r02q_bench.json 2> gpurun_out/r02q_bench.err; tail -c 1500 gpurun_out/r02q_bench.json
python scripts/small_batch.py 4096 16 256
python scripts/small_batch.py 512 32 1024 chain_split=0
python scripts/small_batch.py 512 32 1024
python scripts/chain_timeline.py 32 256 0 512 | tail -1

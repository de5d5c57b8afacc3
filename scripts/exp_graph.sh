python -m pytest tests/test_gpu_chain.py -m gpu -x -q 2>&1 | tail -3
python scripts/small_batch.py 512 32 1024
python scripts/small_batch.py 512 32 1024 chain_graph=0
python scripts/small_batch.py 512 32 20
python scripts/small_batch.py 512 32 20 chain_graph=0
python scripts/small_batch.py 4096 32 384
python scripts/small_batch.py 4096 32 20
python scripts/small_batch.py 4096 32 20 chain_graph=0

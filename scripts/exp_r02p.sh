for cfg in "512 16" "512 32" "512 64" "512 128" "1024 32" "1024 64" "2048 32" "256 64" "256 128"; do python scripts/small_batch.py $cfg 1024 chain_split=0; done
python scripts/small_batch.py 4096 16 256 chain_split=0

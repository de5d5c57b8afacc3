for cfg in "4096 12" "4096 16" "4096 24" "4096 32" "8192 8" "8192 16" "16384 8" "2048 32" "2048 64"; do python scripts/small_batch.py $cfg 256; done

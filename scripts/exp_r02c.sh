set -x
python scripts/chain_timeline.py 8 32 1 > gpurun_out/r02c_timeline_s8.txt 2>&1
python scripts/chain_timeline.py 8 32 0 > gpurun_out/r02c_timeline_s8_nosplit.txt 2>&1
python scripts/chain_timeline.py 16 48 1 > gpurun_out/r02c_timeline_s16.txt 2>&1
python scripts/chain_timeline.py 2 16 1 > gpurun_out/r02c_timeline_s2.txt 2>&1
python scripts/chain_timeline.py 4 16 1 > gpurun_out/r02c_timeline_s4.txt 2>&1
head -45 gpurun_out/r02c_timeline_s8.txt

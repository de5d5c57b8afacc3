python bench.py --no-secondary --no-cpu-baseline --steps 128 --warmup 16 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ecdsa', d['ms_per_step'], d['value'], d['roofline']['frac'], d['kernel_ms'], d['config']['single_stream_ms_per_step'], d['config']['single_stream_ms_per_step_two_lane_replay'], d['config']['streams'], d['device_formats'], d['e2e']['value'])"
python scripts/small_batch.py 2048 32 512
python scripts/small_batch.py 1024 32 512
python scripts/small_batch.py 512 32 1024

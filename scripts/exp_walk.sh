python -m pytest tests/test_genstream.py tests/test_genstream_ecdsa.py -m gpu -x -q 2>&1 | tail -2
for sh in 832 816 808 416 408 404 208; do P2E_OPT_GENSTREAM_WALK_SHAPE=$sh python bench.py --workload genstream_ecdsa --no-secondary --steps 20 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('genstream_ecdsa shape $sh', d['ms_per_step'], d['value'])"; done
for sh in 832 408 404; do P2E_OPT_GENSTREAM_WALK_SHAPE=$sh python bench.py --workload genstream --no-secondary --steps 20 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('genstream shape $sh', d['ms_per_step'], d['value'])"; done

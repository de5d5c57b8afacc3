python -m pytest tests/test_genstream.py tests/test_genstream_ecdsa.py tests/test_gpu_f8.py -m gpu -x -q 2>&1 | tail -2
for v in "" scratch/lib_mi30.so scratch/lib_mi2.so; do
  echo "lib=$v"
  P2E_LIB=$v python bench.py --workload genstream_ecdsa --no-secondary --no-cpu-baseline --steps 20 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('genstream_ecdsa', d['ms_per_step'], d['value'], d['gpu_launches'])"
  P2E_LIB=$v python bench.py --no-secondary --no-cpu-baseline --steps 96 --warmup 16 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ecdsa', d['ms_per_step'], d['kernel_ms'])"
done

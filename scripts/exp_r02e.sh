set -x
python -m pytest tests/test_gpu_f8.py tests/test_gpu_chain.py tests/test_c_host.py -m gpu -q 2>&1 | tail -n 3
python bench.py --no-secondary --no-cpu-baseline --steps 96 > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/r02e_bench.json"))
print("r02e", "ms/step %.4f" % d["ms_per_step"], "frac %.3f" % d["roofline"]["frac"], "single %.3f" % d["config"]["single_stream_ms_per_step"], "errs", d["config"]["error_rows"], d["kernel_ms"])
print(d["device_formats"]); print(d["e2e"]["value"], d["e2e"]["ms_per_step"])
PY

#!/bin/bash
# GPU parity suite + terminal timing + one bench line (no ncu).
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 4 gpurun_out/pytest_gpu.log
python tools/time_terminal.py 1000000 120 3
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print({k:(d[k]['value'] if isinstance(d[k],dict) and 'value' in d[k] else d[k]) for k in ('value','e2e','e2e_dense','cpu_baseline')})
print(d['roofline'])
PY
tail -n 3 gpurun_out/bench.err

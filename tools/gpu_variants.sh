#!/bin/bash
# Times kernel variants built with `python -m em_model_manned_bayes_b200.build --variant <name> <defines>`.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 3 gpurun_out/pytest_gpu.log
echo "== base"; python tools/quick_time.py tracks
for v in "$@"; do
  echo "== $v"; EMB200_LIB=$PWD/em_model_manned_bayes_b200/libemb200_$v.so python tools/quick_time.py tracks
done
python tools/time_terminal.py 1000000 120 3

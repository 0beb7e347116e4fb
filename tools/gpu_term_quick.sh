#!/bin/bash
# terminal-path GPU tests + timing (+ optional variants: bash tools/gpu_term_quick.sh <variant> ...)
mkdir -p gpurun_out
python -m pytest tests/test_terminal_traj.py tests/test_dist.py -m gpu -x -q > gpurun_out/pytest_term.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_term.log
tail -n 4 gpurun_out/pytest_term.log
{
python tools/time_terminal.py 1000000 120 3
for v in "$@"; do
  echo "== $v"; EMB200_LIB=$PWD/em_model_manned_bayes_b200/libemb200_$v.so python tools/time_terminal.py 1000000 120 3
done
} 2>&1 | tee gpurun_out/time_terminal.log

mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
python tools/time_terminal.py 1000000 120 3 > gpurun_out/time_terminal.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_terminal -c 1 -f -o gpurun_out/prof_terminal python tools/time_terminal.py 200000 120 1 > gpurun_out/prof_terminal.log 2>&1
tail -n 5 gpurun_out/pytest_gpu.log; cat gpurun_out/time_terminal.log; tail -3 gpurun_out/prof_terminal.log

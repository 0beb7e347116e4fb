#!/bin/bash
# Quick perf iteration on one B200: GPU parity tests, microbench, kernel timings, bench line.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
[ -x tools/ubench/ubench ] && [ -n "$UBENCH" ] && tools/ubench/ubench > gpurun_out/ubench.log 2>&1
python tools/quick_time.py > gpurun_out/quick_time.log 2>&1
python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
if [ -n "$NCU" ]; then
ncu --set full --clock-control none --import-source on -k regex:k_tracks -s 1 -c 1 -f -o gpurun_out/prof_tracks \
    python bench.py --steps 1 --warmup 1 --no-cpu --e2e-steps 1 --tracks 262144 > gpurun_out/prof_tracks.log 2>&1
fi
tail -n 3 gpurun_out/pytest_gpu.log; cat gpurun_out/quick_time.log; cat gpurun_out/ubench.log 2>/dev/null; cut -c1-400 gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err

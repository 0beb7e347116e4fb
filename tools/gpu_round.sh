#!/bin/bash
# One gpurun call: GPU parity tests, smoke, bench, ncu launch list of the bench command, one ncu --set full capture.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
python tools/quick_time.py > gpurun_out/quick_time.log 2>&1
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --e2e-steps 1 > gpurun_out/bench_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_tracks -s 1 -c 1 -f -o gpurun_out/prof_tracks \
    python bench.py --steps 1 --warmup 1 --no-cpu --e2e-steps 1 --tracks 262144 > gpurun_out/prof_tracks.log 2>&1
tail -n 3 gpurun_out/pytest_gpu.log gpurun_out/smoke.log gpurun_out/quick_time.log gpurun_out/bench.json

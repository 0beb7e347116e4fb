#!/bin/bash
# Round-2 evidence in one gpurun call: GPU tests, smoke, bench (+ reference arm), ncu launch list of the bench command,
# ncu --set full captures of the track kernel and the initial-network kernel (summarised here: gpurun_out/ is capped at 64 MiB).
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --e2e-steps 1 > gpurun_out/bench_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_tracks_fast -s 2 -c 1 -f -o /tmp/prof_tracks \
    python tools/quick_time.py tracks > gpurun_out/prof_tracks.log 2>&1
python tools/ncu_summary.py /tmp/prof_tracks.ncu-rep 750000000 > gpurun_out/ncu_tracks_summary.txt 2>&1
ncu -i /tmp/prof_tracks.ncu-rep --page source --csv > gpurun_out/ncu_tracks_source.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:k_initial_fast -s 4 -c 1 -f -o /tmp/prof_init \
    python tools/time_initial2.py 16777216 > gpurun_out/prof_init.log 2>&1
python tools/ncu_summary.py /tmp/prof_init.ncu-rep 16777216 > gpurun_out/ncu_init_summary.txt 2>&1
ncu --set full --clock-control none -k regex:k_tracks_fast -s 1 -c 1 -f -o /tmp/prof_slow python tools/time_slow.py > gpurun_out/prof_slow.log 2>&1
python tools/ncu_summary.py /tmp/prof_slow.ncu-rep 314572800 > gpurun_out/ncu_slow_summary.txt 2>&1
tail -n 3 gpurun_out/pytest_gpu.log gpurun_out/smoke.log; cut -c1-300 gpurun_out/bench.json; tail -n 2 gpurun_out/bench.err

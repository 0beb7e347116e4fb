#!/bin/bash
# Round-2 closing evidence in one gpurun call: everything tools/gpu_round2.sh records plus the terminal chain kernel's
# --set full summary and per-source-line table.
bash tools/gpu_round2.sh
python tools/time_terminal.py 1000000 120 3 > gpurun_out/time_terminal.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_terminal -c 1 -f -o /tmp/prof_terminal python tools/time_terminal.py 200000 120 1 > gpurun_out/prof_terminal.log 2>&1
python tools/ncu_summary.py /tmp/prof_terminal.ncu-rep 88000000 > gpurun_out/ncu_term_summary.txt 2>&1
ncu -i /tmp/prof_terminal.ncu-rep --page source --print-source cuda,sass --csv 2>/dev/null | gzip -9 > gpurun_out/ncu_term_source_cuda.csv.gz
ncu -i /tmp/prof_terminal.ncu-rep --page source --csv 2>/dev/null | gzip -9 > gpurun_out/ncu_term_source.csv.gz
cat gpurun_out/time_terminal.log

#!/bin/bash
# time the bench workload kernel for each tuning variant (libemb200_<v>.so); scratch tool
mkdir -p gpurun_out
: > gpurun_out/tune.log
for lib in em_model_manned_bayes_b200/libemb200.so em_model_manned_bayes_b200/libemb200_*.so; do
  echo "== $lib" >> gpurun_out/tune.log
  EMB200_LIB=$PWD/$lib python tools/quick_time.py tracks >> gpurun_out/tune.log 2>&1
done
cat gpurun_out/tune.log

#!/bin/bash
# Times kernel variants (no test suite): bash tools/gpu_variants_quick.sh <variant> ...
mkdir -p gpurun_out
{
echo "== base"; python tools/quick_time.py tracks
for v in "$@"; do
  echo "== $v"; EMB200_LIB=$PWD/em_model_manned_bayes_b200/libemb200_$v.so python tools/quick_time.py tracks
done
} 2>&1 | tee gpurun_out/variants.txt

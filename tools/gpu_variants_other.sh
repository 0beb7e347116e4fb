#!/bin/bash
# Times the non-headline kernels for library variants: bash tools/gpu_variants_other.sh <variant> ...
mkdir -p gpurun_out
{
echo "== base"; python tools/prof_other.py 3
for v in "$@"; do
  echo "== $v"; EMB200_LIB=$PWD/em_model_manned_bayes_b200/libemb200_$v.so python tools/prof_other.py 3
done
} 2>&1 | tee gpurun_out/variants_other.txt

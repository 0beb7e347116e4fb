#!/bin/bash
# bench.py at N GPUs of one box, launched exactly as the driver does.  Usage: gpurun --gpus N -- bash tools/gpu_scale.sh N
N=${1:-2}
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
fi
echo "rc=$?"; tail -c 1500 gpurun_out/scale_n$N.json; tail -n 5 gpurun_out/scale_n$N.err

#!/bin/bash
# Multi-GPU bench pass, launched exactly like the driver does.  usage: bash tools/gpu_scale.sh TAG N [N ...]
TAG=$1; shift
O=gpurun_out; mkdir -p $O
for N in "$@"; do
  if [ "$N" = 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 50 --warmup 5 --no-cpu-baseline > $O/scale_${TAG}_n1.json 2> $O/scale_${TAG}_n1.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $N --steps 50 --warmup 5 > $O/scale_${TAG}_n$N.json 2> $O/scale_${TAG}_n$N.err
  fi
  echo "N=$N rc=$?"
done

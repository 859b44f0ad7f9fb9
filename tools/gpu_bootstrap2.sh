#!/bin/bash
# The reference-compatible launcher on 2 GPUs (N = 4 workers, swaps, the reference's default batch 10), peer exchange
# vs NCCL: the final weights must be identical.
O=$PWD/gpurun_out; mkdir -p $O
for X in peer nccl; do
  D=$(mktemp -d); cd $D
  MDGAN_EXCHANGE=$X timeout 300 python $GRAFT_REPO_ROOT/distributed-gan_b200/bootstrap.py --backend nccl --world_size 5 --ranks 0..4 \
    --dataset CIFAR10 --epochs 8 --local_epochs 1 --swap_interval 3 --device cuda --batch_size 10 --iid 1 --seed 3 --beta_1 0.5 \
    --generator_lr 0.0002 --discriminator_lr 0.0002 --log_interval 4 --gpus 2 --synthetic 640 --master_port 29577 > $O/bootstrap2_$X.log 2>&1
  echo "$X rc=$?"; ls weights saved_images | tr '\n' ' '; echo
  python - <<P
import torch,hashlib
g=torch.load("weights/generator_final.pt"); d=torch.load("weights/worker_3/discriminator.pth")
h=hashlib.sha256(b"".join(v.numpy().tobytes() for v in list(g.values())+list(d.values()))).hexdigest()
print("$X sha256", h[:16], "G keys", len(g))
P
  cd - > /dev/null
done
tail -3 $O/bootstrap2_peer.log

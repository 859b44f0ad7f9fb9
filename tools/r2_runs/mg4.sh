#!/bin/bash
# BASELINE config 3: CIFAR-10 shape, K = 4 workers on 4 GPUs, discriminator swap every iteration inside the timed window.
N=4
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 250 $TR --master-port 29631 bench.py --gpus $N --dataset CIFAR10 --swap-interval 1 --steps 20 --warmup 5 --no-shapes > $O/r2mg4_cifar_swap1.json 2> $O/r2mg4_cifar_swap1.err; echo "rc=$?"; tail -2 $O/r2mg4_cifar_swap1.err | cut -c1-300
python - <<PY
import json
d = json.loads(open("gpurun_out/r2mg4_cifar_swap1.json").read().strip().splitlines()[-1])
print("ms", round(d["ms_per_step"], 4), "value", round(d["value"], 1), "e2e ms", round(d["e2e"]["ms_per_step"], 4), d["setup"]["exchange"], d["setup"].get("push"),
      "swaps", d["setup"]["swaps_in_timed_window"], "bit_identical", d.get("multi_gpu_bit_identical"), d["config"]["workload"])
print("per rank", d["setup"]["per_rank_ms"])
PY

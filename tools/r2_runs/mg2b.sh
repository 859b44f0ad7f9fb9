#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 2 --steps 5 --warmup 3 --no-shapes --no-selfcheck > gpurun_out/r2mg2b.json 2> gpurun_out/r2mg2b.err
python -c "
import json; d=json.loads(open('gpurun_out/r2mg2b.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['setup'])"
tail -3 gpurun_out/r2mg2b.err

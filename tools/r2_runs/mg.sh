#!/bin/bash
# Multi-GPU validation: usage gpu_r2_mg.sh N   (N = 2 or 8 GPUs)
N=${1:-2}
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi -L | wc -l
( timeout 900 python -m pytest tests/test_multigpu_gpu.py tests/test_actors_gpu.py -q 2>&1 | tail -15 ) > $O/r2mg${N}_pytest.log; tail -4 $O/r2mg${N}_pytest.log
timeout 400 $TR --master-port 29601 bench.py --gpus $N --steps 20 --warmup 5 > $O/r2mg${N}_bench.json 2> $O/r2mg${N}_bench.err; echo "bench rc=$?"; tail -3 $O/r2mg${N}_bench.err | cut -c1-300
MDGAN_PEER_MULTICAST=0 timeout 300 $TR --master-port 29602 bench.py --gpus $N --steps 20 --warmup 5 --no-shapes --no-selfcheck > $O/r2mg${N}_bench_nomc.json 2> $O/r2mg${N}_bench_nomc.err; echo "nomc rc=$?"
MDGAN_EXCHANGE=nccl timeout 300 $TR --master-port 29603 bench.py --gpus $N --steps 20 --warmup 5 --no-shapes --no-selfcheck > $O/r2mg${N}_bench_nccl.json 2> $O/r2mg${N}_bench_nccl.err; echo "nccl rc=$?"
python - <<PY
import json
for tag in ("bench", "bench_nomc", "bench_nccl"):
    try:
        d = json.loads(open("gpurun_out/r2mg${N}_%s.json" % tag).read().strip().splitlines()[-1])
        print(tag, "ms", round(d["ms_per_step"], 4), "value", round(d["value"], 1), "e2e ms", round(d["e2e"]["ms_per_step"], 4), d["setup"]["exchange"],
              "bit_identical", d.get("multi_gpu_bit_identical"), (d.get("multi_gpu_check") or {}).get("mismatches"),
              {k: round(v["ms_per_step"], 4) for k, v in d.get("shapes", {}).items()})
    except Exception as e:
        print(tag, "FAILED", e)
PY

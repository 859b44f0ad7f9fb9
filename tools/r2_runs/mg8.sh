#!/bin/bash
# 8-GPU validation (expensive: keep it short).
N=8
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi -L | wc -l
( timeout 400 python -m pytest tests/test_multigpu_gpu.py -q -k "peer" 2>&1 | tail -8 ) > $O/r2mg8_pytest.log; tail -3 $O/r2mg8_pytest.log
timeout 300 $TR --master-port 29611 bench.py --gpus $N --steps 20 --warmup 5 > $O/r2mg8_bench.json 2> $O/r2mg8_bench.err; echo "bench rc=$?"; tail -3 $O/r2mg8_bench.err | cut -c1-300
MDGAN_PEER_MULTICAST=0 timeout 200 $TR --master-port 29612 bench.py --gpus $N --steps 20 --warmup 5 --no-shapes --no-selfcheck > $O/r2mg8_bench_nomc.json 2> $O/r2mg8_bench_nomc.err; echo "nomc rc=$?"
python - <<PY
import json
for tag in ("bench", "bench_nomc"):
    try:
        d = json.loads(open("gpurun_out/r2mg8_%s.json" % tag).read().strip().splitlines()[-1])
        print(tag, "ms", round(d["ms_per_step"], 4), "value", round(d["value"], 1), "e2e ms", round(d["e2e"]["ms_per_step"], 4), d["setup"]["exchange"], d["setup"].get("push"),
              "bit_identical", d.get("multi_gpu_bit_identical"), (d.get("multi_gpu_check") or {}).get("mismatches"),
              {k: round(v["ms_per_step"], 4) for k, v in d.get("shapes", {}).items()})
        print("   per rank ms", d["setup"].get("per_rank_ms")); print("   exchange", d["setup"].get("exchange_us_per_rank_eager"))
    except Exception as e:
        print(tag, "FAILED", e)
PY

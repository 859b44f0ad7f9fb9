#!/bin/bash
# wgrad producer restructure (no loads in flight at the proxy fence): correctness + timing.
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
WGRAD_BENCH_ONLY=wgrad timeout 100 python tools/conv_bench.py 1 > $O/r2c8_convbench.log 2>&1; cat $O/r2c8_convbench.log
( timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 ) > $O/r2c8_pytest.log; tail -3 $O/r2c8_pytest.log
timeout 150 python bench.py --dataset CelebA --steps 30 --warmup 5 --no-cpu-baseline > $O/r2c8_bench_celeba.json 2> $O/r2c8_bench_celeba.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2c8_bench_celeba.json").read().strip().splitlines()[-1])
print("ms", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], d["roofline"]["kernel"], d["roofline"]["achieved"], d["roofline"]["frac"], {k:v["ms_per_step"] for k,v in d["shapes"].items()})
for k,v in list(d["per_op"].items())[:8]: print("  ", k, v)
PY
MDGAN_CONV_TA=0 MDGAN_WGRAD_TA=0 MDGAN_BN_FUSED_STATS=0 MDGAN_CONV_UP2=0 timeout 100 python tools/stress_conv.py --dataset CIFAR10 --n 128 --reps 300 --save /tmp/ref.pt > $O/r2c8_stress_nota.log 2>&1
MDGAN_BN_FUSED_STATS=0 MDGAN_CONV_UP2=0 timeout 200 python tools/stress_conv.py --dataset CIFAR10 --n 128 --reps 20000 --against /tmp/ref.pt > $O/r2c8_stress_ta.log 2>&1
MDGAN_CONV_TA=0 MDGAN_WGRAD_TA=0 MDGAN_BN_FUSED_STATS=0 MDGAN_CONV_UP2=0 timeout 100 python tools/stress_conv.py --dataset CelebA --n 128 --reps 200 --save /tmp/refc.pt > $O/r2c8_stress_celeba_nota.log 2>&1
MDGAN_BN_FUSED_STATS=0 MDGAN_CONV_UP2=0 timeout 200 python tools/stress_conv.py --dataset CelebA --n 128 --reps 10000 --against /tmp/refc.pt > $O/r2c8_stress_celeba_ta.log 2>&1
timeout 200 python tools/stress_conv.py --dataset CelebA --n 128 --reps 10000 > $O/r2c8_stress_celeba.log 2>&1
grep -h "stress_conv" $O/r2c8_stress_*.log | grep -v "first rep"

#!/bin/bash
# BatchNorm apply kernels with 32-bit index math: correctness + timing.
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
( timeout 400 python -m pytest tests/test_kernels_gpu.py tests/test_nets_gpu.py tests/test_mdgan_gpu.py -q -x 2>&1 | tail -4 ) > $O/r2c12_pytest.log; tail -2 $O/r2c12_pytest.log
for i in 1 2; do
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r2c12_bench_$i.json 2> $O/r2c12_bench_$i.err; echo "rc=$?"
done
python - <<'PY'
import json
for i in (1, 2):
    d=json.loads(open(f"gpurun_out/r2c12_bench_{i}.json").read().strip().splitlines()[-1])
    print("ms", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], {k:v["ms_per_step"] for k,v in d["shapes"].items()}, {k:v["us_per_iter"] for k,v in d["per_op"].items() if k.startswith("bn")})
PY

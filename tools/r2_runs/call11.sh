#!/bin/bash
# wgrad transposer groups + PDL re-test on the current build.
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
WGRAD_BENCH_ONLY=wgrad timeout 60 python tools/conv_bench.py 1 > $O/r2c11_cb_tg1.log 2>&1
MDGAN_WGRAD_TG=2 WGRAD_BENCH_ONLY=wgrad timeout 60 python tools/conv_bench.py 1 > $O/r2c11_cb_tg2.log 2>&1
paste <(cut -c1-40 $O/r2c11_cb_tg1.log) <(cut -c28-40 $O/r2c11_cb_tg2.log)
for pdl in 0 1; do
  MDGAN_PDL=$pdl timeout 100 python bench.py --dataset CelebA --steps 30 --warmup 5 --no-cpu-baseline --no-shapes > $O/r2c11_bench_pdl$pdl.json 2> $O/r2c11_bench_pdl$pdl.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2c11_bench_pdl$pdl.json").read().strip().splitlines()[-1])
    print("pdl$pdl ms", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], {k:v["us_per_iter"] for k,v in list(d["per_op"].items())[:8]})
except Exception as e: print("pdl$pdl failed", e)
PY
done

#!/bin/bash
# wgrad producer groups 2 / 4 / 8 + PDL re-test.
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
L=$PWD/distributed-gan_b200/mdgan_b200
for v in default hi2 hi8; do
  lib=$L/libmdgan_b200.so; [ $v != default ] && lib=$L/libmdgan_b200_$v.so
  [ -f $lib ] || continue
  MDGAN_B200_LIB=$lib WGRAD_BENCH_ONLY=wgrad timeout 100 python tools/conv_bench.py 1 > $O/r2c9_cb_$v.log 2>&1; echo "== $v"; cat $O/r2c9_cb_$v.log
done
( timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_nets_gpu.py -q -x 2>&1 | tail -4 ) > $O/r2c9_pytest.log; tail -2 $O/r2c9_pytest.log
for pdl in 0 1; do
  MDGAN_PDL=$pdl timeout 150 python bench.py --dataset CelebA --steps 30 --warmup 5 --no-cpu-baseline --no-shapes > $O/r2c9_bench_pdl$pdl.json 2> $O/r2c9_bench_pdl$pdl.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2c9_bench_pdl$pdl.json").read().strip().splitlines()[-1])
print("pdl$pdl ms", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], {k:v["us_per_iter"] for k,v in list(d["per_op"].items())[:6]})
PY
done

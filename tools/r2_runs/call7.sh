#!/bin/bash
# After the elect.sync fix: do transposer groups / pipeline depth matter now?
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
DEEP=$PWD/distributed-gan_b200/mdgan_b200/libmdgan_b200_deep.so
timeout 100 python tools/conv_bench.py 1 > $O/r2c7_cb_tg1.log 2>&1
MDGAN_CONV_TG=2 MDGAN_WGRAD_TG=2 timeout 100 python tools/conv_bench.py 1 > $O/r2c7_cb_tg2.log 2>&1
MDGAN_B200_LIB=$DEEP timeout 100 python tools/conv_bench.py 1 > $O/r2c7_cb_deep1.log 2>&1
MDGAN_B200_LIB=$DEEP MDGAN_CONV_TG=2 MDGAN_WGRAD_TG=2 timeout 100 python tools/conv_bench.py 1 > $O/r2c7_cb_deep2.log 2>&1
paste <(cut -c1-40 $O/r2c7_cb_tg1.log) <(cut -c28-40 $O/r2c7_cb_tg2.log) <(cut -c28-40 $O/r2c7_cb_deep1.log) <(cut -c28-40 $O/r2c7_cb_deep2.log)
timeout 150 python bench.py --dataset CelebA --steps 30 --warmup 5 --no-cpu-baseline --no-shapes > $O/r2c7_bench.json 2> $O/r2c7_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2c7_bench.json").read().strip().splitlines()[-1])
print("ms", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], d["roofline"]["kernel"], d["roofline"]["achieved"], d["roofline"]["frac"])
for k,v in d["per_op"].items(): print("  ", k, v)
PY

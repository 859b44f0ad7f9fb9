#!/bin/bash
# pytest (GPU) then A/B bench of an env toggle.  usage: bash tools/gpu_ab.sh TAG "ENV=VAL" [datasets...]
TAG=$1; TOGGLE=$2; shift; shift
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_${TAG}.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_${TAG}.log
for D in "$@"; do
  timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --dataset $D > $O/bench_${TAG}_${D}.json 2> $O/bench_${TAG}_${D}.err; echo "bench $D rc=$?"
  env $TOGGLE timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --dataset $D > $O/bench_${TAG}_${D}_off.json 2> $O/bench_${TAG}_${D}_off.err; echo "bench $D ($TOGGLE) rc=$?"
done
python - <<'P'
import json,glob,sys
for f in sorted(glob.glob('gpurun_out/bench_%s_*.json' % sys.argv[1] if len(sys.argv)>1 else 'gpurun_out/bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), 'launches', d['gpu_launches_per_step'])
    except Exception as e: print(f, 'ERR', e)
P

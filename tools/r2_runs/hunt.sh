#!/bin/bash
# Stall hunt for the shipped build: N benchmark processes back to back, each dumps its Python stack after 30 s
# (faulthandler) and is killed at 45 s.  usage: gpu_r2_hunt.sh [runs]
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
RUNS=${1:-8}
( timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_nets_gpu.py -q -x 2>&1 | tail -3 ) > $O/r2h_pytest.log; tail -2 $O/r2h_pytest.log
WGRAD_BENCH_ONLY=wgrad timeout 60 python tools/conv_bench.py 1
ok=0; bad=0
for i in $(seq 1 $RUNS); do
  timeout 45 python -c "
import faulthandler, sys, runpy
faulthandler.dump_traceback_later(30, exit=True)
sys.argv = ['bench.py'] + sys.argv[1:]
runpy.run_path('bench.py', run_name='__main__')
" --steps 10 --warmup 3 --no-cpu-baseline --no-shapes > $O/r2h_$i.json 2> $O/r2h_$i.err
  if [ $? -eq 0 ]; then ok=$((ok+1)); python -c "
import json; print('run $i ms', json.loads(open('$O/r2h_$i.json').read().strip().splitlines()[-1])['ms_per_step'])"
  else bad=$((bad+1)); echo "run $i STALLED"; grep -A8 "most recent call first" $O/r2h_$i.err | head -12; fi
done
echo "stall hunt: $ok ok, $bad stalled of $RUNS"

#!/bin/bash
# The whole GPU suite and one default bench line on the FINAL tree of round 2.
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
( timeout 105 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -12 ) > $O/r2k_pytest.log; tail -3 $O/r2k_pytest.log
timeout 45 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r2k_bench.json 2> $O/r2k_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2k_bench.json").read().strip().splitlines()[-1])
    print("ms", round(d["ms_per_step"], 4), "value", round(d["value"], 2), "e2e", round(d["e2e"]["value"], 2), round(d["e2e"]["ms_per_step"], 4),
          {k: (round(v["ms_per_step"], 4) if "ms_per_step" in v else v) for k, v in (d.get("shapes") or {}).items()})
except Exception as e:
    print("bench FAILED", e)
PY

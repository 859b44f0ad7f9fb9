#!/bin/bash
# Last round-2 call (7 GPU-minutes left): the whole GPU suite with the additions of this session (early upload, ingest
# from MNIST-format files, the MLP path), the early-upload A/B on the end-to-end step, one default bench line.
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
( timeout 250 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -40 ) > $O/r2g_pytest.log; tail -4 $O/r2g_pytest.log
timeout 80 python tools/e2e_ab.py --steps 40 --rounds 2 > $O/r2g_e2e_ab.json 2> $O/r2g_e2e_ab.err; echo "e2e_ab rc=$?"; cat $O/r2g_e2e_ab.json
timeout 140 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r2g_bench.json 2> $O/r2g_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2g_bench.json").read().strip().splitlines()[-1])
    print("ms", round(d["ms_per_step"], 4), "value", round(d["value"], 2), "e2e", round(d["e2e"]["value"], 2), round(d["e2e"]["ms_per_step"], 4),
          {k: (round(v["ms_per_step"], 4) if "ms_per_step" in v else v) for k, v in (d.get("shapes") or {}).items()})
except Exception as e:
    print("bench FAILED", e)
PY

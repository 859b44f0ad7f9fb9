#!/bin/bash
# Device-side dropout masks of the MLP path (z_source = "device"): its test, where the MLP step spends its time, and a
# complete default bench line of the final tree.
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
( timeout 100 python -m pytest tests/test_mlp_gpu.py -q -m gpu -k "device_mask or engine_matches_oracle" -p no:cacheprovider 2>&1 | tail -30 ) > $O/r2i_pytest_mlp.log; tail -3 $O/r2i_pytest_mlp.log
timeout 60 python tools/mlp_profile.py > $O/r2i_mlp_profile.json 2> $O/r2i_mlp_profile.err; echo "mlp profile rc=$?"; cat $O/r2i_mlp_profile.json
timeout 120 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r2i_bench.json 2> $O/r2i_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2i_bench.json").read().strip().splitlines()[-1])
    print("ms", round(d["ms_per_step"], 4), "value", round(d["value"], 2), "e2e", round(d["e2e"]["value"], 2), round(d["e2e"]["ms_per_step"], 4),
          {k: (round(v["ms_per_step"], 4) if "ms_per_step" in v else v) for k, v in (d.get("shapes") or {}).items()})
except Exception as e:
    print("bench FAILED", e)
PY

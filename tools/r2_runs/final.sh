#!/bin/bash
# Final round-2 call: the driver's sequence (GPU tests, smoke, reference arm, bench) + the profiles of the final build
# + the K = 1 batch sweep (BASELINE config 5).
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
( timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 ) > $O/r2f_pytest.log; tail -3 $O/r2f_pytest.log
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2f_smoke.log 2>&1; tail -1 $O/r2f_smoke.log | cut -c1-200
timeout 400 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/r2f_bench_reference.json 2> $O/r2f_bench_reference.err; echo "ref rc=$?"
for i in 1 2; do
  timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r2f_bench_$i.json 2> $O/r2f_bench_$i.err; echo "bench $i rc=$?"
done
python - <<'PY'
import json
for f in ("r2f_bench_1", "r2f_bench_2", "r2f_bench_reference"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "ms", round(d["ms_per_step"], 4), "value", round(d["value"], 2), "e2e", round(d["e2e"]["value"], 2), (d.get("cpu_baseline") or {}).get("kind"),
              {k: round(v["ms_per_step"], 4) for k, v in (d.get("shapes") or {}).items()}, (d.get("roofline") or {}).get("kernel"), (d.get("roofline") or {}).get("frac"))
    except Exception as e:
        print(f, "FAILED", e)
PY
CMD="python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-shapes"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/r02_launchlist_celeba_b64.csv $CMD > $O/r02_launchlist.log 2>&1; echo "launch list rc=$?"
timeout 500 ncu --set full --clock-control none --import-source on -k "regex:conv_gemm_ta|wgrad_gemm_ta" -s 75 -c 25 -f -o /tmp/r02_full_gemm $CMD > $O/r02_full_gemm.log 2>&1; echo "full gemm rc=$?"
ncu -i /tmp/r02_full_gemm.ncu-rep --page raw --csv > $O/r02_full_gemm_raw.csv 2>/dev/null
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -s 268 -c 95 --csv --log-file $O/r02_metrics_all.csv $CMD > $O/r02_metrics_all.log 2>&1; echo "metrics rc=$?"
for D in CelebA CIFAR10; do
  for B in 32 128 256 512 1024; do
    timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-shapes --dataset $D --batch $B > $O/r2f_sweep_${D}_b$B.json 2> $O/r2f_sweep_${D}_b$B.err; echo "$D b=$B rc=$?"
  done
done
du -sh $O

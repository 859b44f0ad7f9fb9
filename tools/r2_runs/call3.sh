#!/bin/bash
# Round-2 GPU call 3 (lean): smoke gate, suite, bench variants, parity numbers.  Every command under a short timeout.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c3_smoke.log 2>&1 || { tail -30 gpurun_out/r2c3_smoke.log; echo SMOKE FAILED; }
tail -2 gpurun_out/r2c3_smoke.log | cut -c1-400
( timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -60 ) > gpurun_out/r2c3_pytest.log
tail -5 gpurun_out/r2c3_pytest.log
timeout 300 python tools/parity_report.py all trajectory,unpatched > gpurun_out/r2c3_parity_report.jsonl 2> gpurun_out/r2c3_parity_report.err
for v in default up2off fusedoff; do
  case $v in
    default) env="";;
    up2off) env="MDGAN_CONV_UP2=0";;
    fusedoff) env="MDGAN_BN_FUSED_STATS=0";;
  esac
  env $env timeout 150 python bench.py --dataset CelebA --steps 30 --warmup 5 --no-cpu-baseline --no-shapes > gpurun_out/r2c3_bench_celeba_$v.json 2> gpurun_out/r2c3_bench_celeba_$v.err
  env $env timeout 100 python tools/conv_bench.py 1 > gpurun_out/r2c3_convbench_$v.log 2>&1
done
timeout 200 python tools/stress_conv.py --dataset CelebA --n 128 --reps 20000 > gpurun_out/r2c3_stress_celeba.log 2>&1
timeout 200 python tools/stress_conv.py --dataset CIFAR10 --n 128 --reps 40000 > gpurun_out/r2c3_stress_cifar.log 2>&1
grep -h "stress_conv" gpurun_out/r2c3_stress_*.log | grep -v "first rep"
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c3_bench_default.json 2> gpurun_out/r2c3_bench_default.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2c3_bench_reference.json 2> gpurun_out/r2c3_bench_reference.err
cut -c1-300 gpurun_out/r2c3_bench_default.json

#!/bin/bash
# Round-2 profiling call: launch list of the benchmark workload + ncu --set full of one iteration's GEMM kernels +
# DRAM metrics of the bandwidth kernels.  Reports are exported to CSV on the box (gpurun_out is capped at 64 MiB).
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
CMD="python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-shapes"
timeout 300 $CMD > $O/r02_plain.log 2>&1 || { tail -20 $O/r02_plain.log; exit 1; }
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/r02_launchlist_celeba_b64.csv $CMD > $O/r02_launchlist.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:conv_gemm_ta|wgrad_gemm_ta" -s 75 -c 25 -f -o /tmp/r02_full_gemm $CMD > $O/r02_full_gemm.log 2>&1; echo "full gemm rc=$?"
ncu -i /tmp/r02_full_gemm.ncu-rep --page raw --csv > $O/r02_full_gemm_raw.csv 2>/dev/null
ncu -i /tmp/r02_full_gemm.ncu-rep --page source --csv 2>/dev/null | gzip > $O/r02_full_gemm_src.csv.gz
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -s 268 -c 95 --csv --log-file $O/r02_metrics_all.csv $CMD > $O/r02_metrics_all.log 2>&1; echo "metrics rc=$?"
timeout 120 python -m pytest tests/test_kernels_gpu.py -q -k "fused" 2>&1 | tail -3
ls -la $O/r02_*; du -sh $O

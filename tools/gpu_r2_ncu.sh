#!/bin/bash
# Round-2 profiling call: launch list of the benchmark workload + ncu --set full of one iteration's kernels.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
CMD="python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-shapes"
timeout 300 $CMD > $O/r02_plain.log 2>&1 || { tail -20 $O/r02_plain.log; exit 1; }
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/r02_launchlist_celeba_b64.csv $CMD > $O/r02_launchlist.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:conv_gemm_ta|wgrad_gemm_ta" -s 75 -c 25 -f -o $O/r02_full_gemm $CMD > $O/r02_full_gemm.log 2>&1; echo "full gemm rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:bn_|thin_|adam|pack_weights|wgrad_unpack|head_|tanh_|reduce_slices" -s 201 -c 67 -f -o $O/r02_full_other $CMD > $O/r02_full_other.log 2>&1; echo "full other rc=$?"
ls -la $O/r02_*

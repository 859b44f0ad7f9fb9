#!/bin/bash
# SGEMM rework of the MLP path (32 x 32 tiles on the 64..128-row layers, 128-bit shared-memory loads, register prefetch):
# its tests + the step time of the MLP workload.
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
( timeout 150 python -m pytest tests/test_mlp_gpu.py -q -m gpu -p no:cacheprovider 2>&1 | tail -30 ) > $O/r2h_pytest_mlp.log; tail -3 $O/r2h_pytest_mlp.log
timeout 60 python tools/e2e_ab.py --dataset MNIST --steps 40 --rounds 1 > $O/r2h_mlp_step.json 2> $O/r2h_mlp_step.err; echo "mlp step rc=$?"; cat $O/r2h_mlp_step.json

#!/bin/bash
# Split-K SGEMM for the 64..128-row layers of the MLP path: tests + step profile with it on and off.
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
( timeout 120 python -m pytest tests/test_mlp_gpu.py -q -m gpu -p no:cacheprovider 2>&1 | tail -30 ) > $O/r2j_pytest_mlp.log; tail -3 $O/r2j_pytest_mlp.log
timeout 50 python tools/mlp_profile.py > $O/r2j_mlp_profile_splitk.json 2> $O/r2j_mlp_profile_splitk.err; echo "splitk rc=$?"; cat $O/r2j_mlp_profile_splitk.json
MDGAN_SGEMM_SPLITK=0 timeout 50 python tools/mlp_profile.py > $O/r2j_mlp_profile_nosplit.json 2> $O/r2j_mlp_profile_nosplit.err; echo "nosplit rc=$?"; cut -c1-200 $O/r2j_mlp_profile_nosplit.json

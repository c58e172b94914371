#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
: > gpurun_out/r02j.txt
export MCPC_NATIVE_LIB=$PWD/montecarlopredictivecoding_b200/libmcpc_b200_debug.so
MCPC_C5_T=100 MCPC_WIDE_TIMING=1 timeout 300 python scripts/bench_configs.py c5 2>&1 | grep -E "wide timeline|tile|SM clock|alive|ms_per_step" | tail -48 | cut -c1-200 >> gpurun_out/r02j.txt
unset MCPC_NATIVE_LIB
timeout 900 python -m pytest tests/test_gpu_wide.py -q --timeout 400 2>&1 | tail -2 >> gpurun_out/r02j.txt
cat gpurun_out/r02j.txt

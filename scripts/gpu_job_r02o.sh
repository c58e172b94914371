#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
: > gpurun_out/r02o.txt
export MCPC_C5_T=100 MCPC_C5_LIVE_PEAK=0
for rep in 1 2 3; do
for lib in libmcpc_b200_debug.so libmcpc_b200_debug_ns5.so; do
  echo -n "$lib: " >> gpurun_out/r02o.txt
  MCPC_NATIVE_LIB=$PWD/montecarlopredictivecoding_b200/$lib timeout 300 python scripts/bench_configs.py c5 2>&1 | tail -1 | grep -o '"ms_per_step": [0-9.]*' >> gpurun_out/r02o.txt
done
done
cat gpurun_out/r02o.txt

#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
: > gpurun_out/r02g.txt
for w in 0 6 12 24; do
  echo "## WPF=$w T=20" >> gpurun_out/r02g.txt
  MCPC_WIDE_WPF=$w timeout 300 python scripts/bench_configs.py c5 2>&1 | tail -1 | cut -c100-250 >> gpurun_out/r02g.txt
done
timeout 900 python -m pytest tests/test_gpu_wide.py -q --timeout 400 2>&1 | tail -3 >> gpurun_out/r02g.txt
export MCPC_NATIVE_LIB=$PWD/montecarlopredictivecoding_b200/libmcpc_b200_debug.so
echo "## timeline WPF=12 (debug lib)" >> gpurun_out/r02g.txt
MCPC_C5_T=8 MCPC_WIDE_TIMING=1 timeout 300 python scripts/bench_configs.py c5 2>&1 | grep -A 8 "wide timeline" | tail -20 >> gpurun_out/r02g.txt
unset MCPC_NATIVE_LIB
cat gpurun_out/r02g.txt

#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_wide.py -q --timeout 400 2>&1 | tail -8 > gpurun_out/r02f.txt
for cg in 2 1; do
  echo "## staged epilogue inputs CG=$cg T=20" >> gpurun_out/r02f.txt
  MCPC_WIDE_CG=$cg timeout 300 python scripts/bench_configs.py c5 2>&1 | tail -1 | cut -c100-330 >> gpurun_out/r02f.txt
done
echo "## T=100" >> gpurun_out/r02f.txt
MCPC_C5_T=100 timeout 300 python scripts/bench_configs.py c5 2>&1 | tail -1 | cut -c100-330 >> gpurun_out/r02f.txt
export MCPC_NATIVE_LIB=$PWD/montecarlopredictivecoding_b200/libmcpc_b200_debug.so
echo "## timeline (debug lib)" >> gpurun_out/r02f.txt
MCPC_C5_T=8 MCPC_WIDE_TIMING=1 timeout 300 python scripts/bench_configs.py c5 2>&1 | grep -A 12 "wide timeline" | tail -34 >> gpurun_out/r02f.txt
for mode in 1 2; do
  echo "## debug lib EPI_MODE=$mode" >> gpurun_out/r02f.txt
  MCPC_WIDE_EPI_MODE=$mode timeout 300 python scripts/bench_configs.py c5 2>&1 | tail -1 | cut -c100-330 >> gpurun_out/r02f.txt
done
unset MCPC_NATIVE_LIB
cat gpurun_out/r02f.txt

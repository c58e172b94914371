#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_n2_n3.py tests/test_gpu_wide.py -q --timeout 400 2>&1 | grep -v Warning | tail -25 > gpurun_out/r02c_tests.txt
: > gpurun_out/c5_dbg.txt
export MCPC_NATIVE_LIB=$PWD/montecarlopredictivecoding_b200/libmcpc_b200_debug.so
for mode in 0 1 2 3; do
  echo "## debug lib EPI_MODE=$mode CG=2 S=4 T=20" >> gpurun_out/c5_dbg.txt
  MCPC_WIDE_EPI_MODE=$mode timeout 300 python scripts/bench_configs.py c5 2>&1 | tail -1 >> gpurun_out/c5_dbg.txt
done
echo "## debug lib EPI_MODE=1 CG=1" >> gpurun_out/c5_dbg.txt
MCPC_WIDE_CG=1 MCPC_WIDE_EPI_MODE=1 timeout 300 python scripts/bench_configs.py c5 2>&1 | tail -1 >> gpurun_out/c5_dbg.txt
unset MCPC_NATIVE_LIB
echo "## release T=100" >> gpurun_out/c5_dbg.txt
MCPC_C5_T=100 timeout 300 python scripts/bench_configs.py c5 2>&1 | tail -1 >> gpurun_out/c5_dbg.txt
cat gpurun_out/r02c_tests.txt gpurun_out/c5_dbg.txt

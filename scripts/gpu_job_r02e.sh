#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_wide.py -q --timeout 400 2>&1 | tail -3 > gpurun_out/r02e.txt
for cs in 0 1; do
  echo "## per-layer layout CS=$cs T=20" >> gpurun_out/r02e.txt
  MCPC_WIDE_CS=$cs timeout 300 python scripts/bench_configs.py c5 2>&1 | tail -1 | cut -c100-330 >> gpurun_out/r02e.txt
done
echo "## per-layer layout T=100" >> gpurun_out/r02e.txt
MCPC_C5_T=100 timeout 300 python scripts/bench_configs.py c5 2>&1 | tail -1 | cut -c100-330 >> gpurun_out/r02e.txt
export MCPC_NATIVE_LIB=$PWD/montecarlopredictivecoding_b200/libmcpc_b200_debug.so
echo "## timeline (debug lib)" >> gpurun_out/r02e.txt
MCPC_C5_T=8 MCPC_WIDE_TIMING=1 timeout 300 python scripts/bench_configs.py c5 2>&1 | grep -A 12 "wide timeline" | tail -45 >> gpurun_out/r02e.txt
echo "## timeline, no stores (mode 2)" >> gpurun_out/r02e.txt
MCPC_C5_T=8 MCPC_WIDE_TIMING=1 MCPC_WIDE_EPI_MODE=2 timeout 300 python scripts/bench_configs.py c5 2>&1 | grep -A 12 "wide timeline" | tail -45 >> gpurun_out/r02e.txt
unset MCPC_NATIVE_LIB
timeout 1200 python -m pytest tests/test_gpu_bf16_bound.py -q -s --timeout 600 2>&1 | grep -E "C2 bf16|C3 bf16|C4 |\[fp32\]|\[bf16\]|MSE ours|passed|failed|^E  " | cut -c1-700 >> gpurun_out/r02e.txt
cat gpurun_out/r02e.txt

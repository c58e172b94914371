#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
: > gpurun_out/r02k.txt
timeout 900 python -m pytest tests/test_gpu_wide.py -q --timeout 400 2>&1 | tail -4 >> gpurun_out/r02k.txt
for pdl in 0 1; do
  echo "## PDL=$pdl T=20" >> gpurun_out/r02k.txt
  MCPC_WIDE_PDL=$pdl timeout 300 python scripts/bench_configs.py c5 2>&1 | tail -1 | cut -c100-250 >> gpurun_out/r02k.txt
  echo "## PDL=$pdl T=100" >> gpurun_out/r02k.txt
  MCPC_WIDE_PDL=$pdl MCPC_C5_T=100 timeout 300 python scripts/bench_configs.py c5 2>&1 | tail -1 | cut -c100-250 >> gpurun_out/r02k.txt
done
export MCPC_NATIVE_LIB=$PWD/montecarlopredictivecoding_b200/libmcpc_b200_debug.so
MCPC_C5_T=100 MCPC_WIDE_TIMING=1 timeout 300 python scripts/bench_configs.py c5 2>&1 | grep -E "wide timeline|tile|SM clock|alive|ms_per_step" | grep -v wgrad | head -24 | cut -c1-200 >> gpurun_out/r02k.txt
unset MCPC_NATIVE_LIB
cat gpurun_out/r02k.txt

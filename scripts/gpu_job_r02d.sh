#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_bf16_bound.py tests/test_gpu_n2_n3.py -q -s --timeout 600 2>&1 | grep -v "Warning\|warnings.warn" | tail -60 > gpurun_out/r02d_tests.txt
: > gpurun_out/c5_cs.txt
for cs in 0 1; do
  echo "## CS=$cs CG=2 S=4 T=20" >> gpurun_out/c5_cs.txt
  MCPC_WIDE_CS=$cs timeout 300 python scripts/bench_configs.py c5 2>&1 | tail -1 >> gpurun_out/c5_cs.txt
done
for cs in 0 1; do
  echo "## CS=$cs T=100" >> gpurun_out/c5_cs.txt
  MCPC_WIDE_CS=$cs MCPC_C5_T=100 timeout 300 python scripts/bench_configs.py c5 2>&1 | tail -1 >> gpurun_out/c5_cs.txt
done
timeout 900 python -m pytest tests/test_gpu_wide.py -q --timeout 400 2>&1 | tail -3 >> gpurun_out/r02d_tests.txt
cat gpurun_out/r02d_tests.txt gpurun_out/c5_cs.txt

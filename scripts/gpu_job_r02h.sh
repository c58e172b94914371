#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
: > gpurun_out/r02h.txt
for g in 1 0; do
  echo "## G32=$g T=20" >> gpurun_out/r02h.txt
  MCPC_WIDE_G32=$g timeout 300 python scripts/bench_configs.py c5 2>&1 | tail -1 | cut -c100-250 >> gpurun_out/r02h.txt
done
echo "## G32=0 T=100" >> gpurun_out/r02h.txt
MCPC_C5_T=100 timeout 300 python scripts/bench_configs.py c5 2>&1 | tail -1 | cut -c100-250 >> gpurun_out/r02h.txt
timeout 900 python -m pytest tests/test_gpu_wide.py -q -s --timeout 400 2>&1 | grep -E "C5|passed|failed|odd widths|^E " | cut -c1-400 >> gpurun_out/r02h.txt
MCPC_WIDE_G32=1 timeout 900 python -m pytest tests/test_gpu_wide.py -q --timeout 400 2>&1 | tail -2 >> gpurun_out/r02h.txt
timeout 600 python -m pytest tests/test_gpu_reference_callsites.py tests/test_gpu_n2_n3.py -q -s --timeout 400 2>&1 | grep -E "MAP rep|passed|failed|^E " | cut -c1-300 >> gpurun_out/r02h.txt
cat gpurun_out/r02h.txt

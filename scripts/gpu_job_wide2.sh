#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_wide.py -q --timeout 400 2>&1 | grep -v Warning | tail -15 > gpurun_out/wide_all.txt
: > gpurun_out/c5_sweep2.txt
for pf in 0 1; do
  echo "## EPIPF=$pf CG=2 S=4 T=20" >> gpurun_out/c5_sweep2.txt
  MCPC_WIDE_EPIPF=$pf timeout 300 python scripts/bench_configs.py c5 2>&1 | tail -1 >> gpurun_out/c5_sweep2.txt
done
echo "## EPIPF=1 CG=2 S=4 T=100" >> gpurun_out/c5_sweep2.txt
MCPC_C5_T=100 timeout 300 python scripts/bench_configs.py c5 2>&1 | tail -1 >> gpurun_out/c5_sweep2.txt
MCPC_C5_T=6 timeout 900 ncu --set full --clock-control none --import-source on -k regex:wide_kernel -s 20 -c 3 -o gpurun_out/r02b_wide \
  python scripts/bench_configs.py c5 > gpurun_out/r02b_wide_ncu.log 2>&1
cat gpurun_out/wide_all.txt gpurun_out/c5_sweep2.txt

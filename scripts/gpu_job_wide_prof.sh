#!/bin/bash
# GPU job: profile the streaming path on C5 (launch list + full ncu capture of the three kernels) and time T=100.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export MCPC_C5_T=6
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r02a_wide_launches.csv \
  python scripts/bench_configs.py c5 > gpurun_out/r02a_wide_launch_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wide_kernel -s 20 -c 5 -o gpurun_out/r02a_wide \
  python scripts/bench_configs.py c5 > gpurun_out/r02a_wide_ncu.log 2>&1
export MCPC_C5_T=100
for s in 2 4; do
  echo "## CG=2 SLOTS=$s T=100" >> gpurun_out/c5_t100.txt
  MCPC_WIDE_SLOTS=$s timeout 300 python scripts/bench_configs.py c5 2>&1 | tail -1 >> gpurun_out/c5_t100.txt
done
cat gpurun_out/c5_t100.txt

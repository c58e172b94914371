#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python scripts/host_profile.py > gpurun_out/r02_host_profile.txt 2>&1
# launch list of the bench command (cold-cache, serialised: shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 400 --csv --log-file gpurun_out/r02_bf16_launches.csv \
  python bench.py --steps 4 --warmup 3 --no-other-workloads --no-cpu-baseline > gpurun_out/r02_launch_bench.log 2>&1
# full captures: C2 resident kernel + its weight gradient, the three streaming kernels
timeout 900 ncu --set full --clock-control none --import-source on -k regex:infer_tc_kernel -s 4 -c 1 -o gpurun_out/r02_infer_tc \
  python bench.py --steps 2 --warmup 3 --no-other-workloads --no-cpu-baseline > gpurun_out/r02_ncu_tc.log 2>&1
MCPC_C5_T=8 MCPC_C5_LIVE_PEAK=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:wide_kernel -s 30 -c 5 -o gpurun_out/r02_wide \
  python scripts/bench_configs.py c5 > gpurun_out/r02_ncu_wide.log 2>&1
MCPC_C5_T=8 MCPC_C5_LIVE_PEAK=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/r02_wide_launches.csv \
  python scripts/bench_configs.py c5 > gpurun_out/r02_wide_launch_bench.log 2>&1
cat gpurun_out/r02_host_profile.txt
ls -la gpurun_out/r02_*

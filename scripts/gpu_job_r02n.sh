#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
: > gpurun_out/r02n.txt
timeout 900 python -m pytest tests/test_gpu_wide.py -q --timeout 400 2>&1 | tail -2 >> gpurun_out/r02n.txt
for rep in 1 2 3; do
  MCPC_C5_T=100 timeout 300 python scripts/bench_configs.py c5 2>&1 | tail -1 | grep -o '"ms_per_step": [0-9.]*\|"frac_of_live_cublas": [0-9.]*\|"live_cublas_bf16_tflops_sustained": [0-9.]*' | tr '\n' ' ' >> gpurun_out/r02n.txt
  echo >> gpurun_out/r02n.txt
done
MCPC_C5_T=8 MCPC_C5_LIVE_PEAK=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:wide_kernel -s 30 -c 5 -o gpurun_out/r02_wide \
  python scripts/bench_configs.py c5 > gpurun_out/r02_ncu_wide.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:infer_tc_kernel.*, 1>' -s 1 -c 1 -o gpurun_out/r02_infer_tc \
  python bench.py --steps 2 --warmup 3 --no-other-workloads --no-cpu-baseline > gpurun_out/r02_ncu_tc.log 2>&1
cat gpurun_out/r02n.txt

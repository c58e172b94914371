#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
: > gpurun_out/r02l.txt
export MCPC_C5_T=100
for rep in 1 2; do
for cfg in "MCPC_WIDE_SLOTS=4 MCPC_WIDE_PDL=0" "MCPC_WIDE_SLOTS=4 MCPC_WIDE_PDL=1" "MCPC_WIDE_SLOTS=8 MCPC_WIDE_PDL=0" "MCPC_WIDE_SLOTS=2 MCPC_WIDE_PDL=0" "MCPC_WIDE_CG=1 MCPC_WIDE_PDL=0" "MCPC_WIDE_CS=1 MCPC_WIDE_PDL=0"; do
  echo "## rep $rep: $cfg" >> gpurun_out/r02l.txt
  env $cfg timeout 300 python scripts/bench_configs.py c5 2>&1 | tail -1 | grep -o '"ms_per_step": [0-9.]*' >> gpurun_out/r02l.txt
done
done
cat gpurun_out/r02l.txt

#!/bin/bash
# GPU job: the streaming (wide) path -- parity tests per CTA-group mode, then the C5 timing sweep.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
for sel in cg1 cg2 "not cg1 and not cg2"; do
  tag=$(echo "$sel" | tr ' ' '_')
  timeout 1200 python -m pytest tests/test_gpu_wide.py -q -s -k "$sel" --timeout 400 2>&1 | grep -v Warning | tail -60 > "gpurun_out/wide_${tag}.txt"
  echo "== $sel: exit $?" >> "gpurun_out/wide_${tag}.txt"
done
: > gpurun_out/c5_sweep.txt
for cg in 1 2; do for s in 1 4 8; do
  echo "## CG=$cg SLOTS=$s" >> gpurun_out/c5_sweep.txt
  MCPC_WIDE_CG=$cg MCPC_WIDE_SLOTS=$s timeout 300 python scripts/bench_configs.py c5 2>&1 | tail -3 >> gpurun_out/c5_sweep.txt
done; done
tail -5 gpurun_out/wide_*.txt
cat gpurun_out/c5_sweep.txt

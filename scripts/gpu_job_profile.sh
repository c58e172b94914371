#!/bin/bash
# ncu evidence for the C2 bench command: launch list of the bench itself + one --set full capture of the resident kernel and
# of the tensor-core weight-gradient kernel (MIX=50 SAMP=100 scripts/tc_timing.py = the bench's MCPC call, B=1024, T=150).
# TAG names the output files (profiles are kept per round / iteration).
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
TAG=${TAG:-r02c}
if [ -z "$ONLY_TC" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_bf16_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-other-workloads > gpurun_out/prof_bench.log 2>&1
MIX=50 SAMP=100 timeout 600 ncu --set full --clock-control none --import-source on -k regex:wgrad_tc_kernel -s 1 -c 1 \
  -o gpurun_out/${TAG}_wgrad_tc python scripts/tc_timing.py >> gpurun_out/prof_tc.log 2>&1
fi
MIX=50 SAMP=100 timeout 600 ncu --set full --clock-control none --import-source on -k regex:infer_tc_kernel -s 1 -c 1 \
  -f -o gpurun_out/${TAG}_infer_tc python scripts/tc_timing.py > gpurun_out/prof_tc.log 2>&1
ls -la gpurun_out/*.ncu-rep; tail -3 gpurun_out/prof_tc.log

cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out; rm -f gpurun_out/micro.txt
MIX=50 SAMP=100 MCPC_TC_TIMING=100 timeout 120 python scripts/tc_timing.py > gpurun_out/tc_trace.txt 2>&1
timeout 120 python scripts/host_gaps.py > gpurun_out/micro.txt 2>&1
timeout 120 python scripts/host_timeline.py >> gpurun_out/micro.txt 2>&1
cat gpurun_out/micro.txt

cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out; rm -f gpurun_out/micro.txt
MIX=50 SAMP=100 MCPC_TC_TIMING=100 timeout 120 python scripts/tc_timing.py > gpurun_out/tc_trace.txt 2>&1
MODE=map MCPC_TC_TIMING=100 timeout 120 python scripts/tc_timing.py > gpurun_out/tc_trace_map.txt 2>&1
tail -2 gpurun_out/tc_trace.txt | cut -c1-800; tail -2 gpurun_out/tc_trace_map.txt | cut -c1-800

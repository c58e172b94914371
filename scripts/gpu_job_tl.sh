cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out; rm -f gpurun_out/micro.txt
MIX=50 SAMP=100 MCPC_TC_TIMING=100 timeout 120 python scripts/tc_timing.py > gpurun_out/tc_trace.txt 2>&1
tail -3 gpurun_out/tc_trace.txt | cut -c1-700 >> gpurun_out/micro.txt
timeout 120 python scripts/host_timeline.py 2>&1 | tail -16 >> gpurun_out/micro.txt
timeout 900 python -m pytest tests/test_gpu_bf16.py tests/test_gpu_parity.py tests/test_gpu_bf16_bound.py tests/test_gpu_n2_n3.py -x -q 2>&1 | grep -v "Warning\|warnings.warn" | tail -8 >> gpurun_out/micro.txt
cat gpurun_out/micro.txt

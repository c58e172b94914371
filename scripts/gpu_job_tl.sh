cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_bf16.py tests/test_gpu_parity.py -x -q 2>&1 | grep -v "Warning\|warnings.warn" | tail -30 > gpurun_out/micro.txt
cat gpurun_out/micro.txt

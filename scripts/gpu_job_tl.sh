cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_wide.py tests/test_gpu_bf16.py -x -q 2>&1 | grep -v "Warning\|warnings.warn" | tail -40 > gpurun_out/micro.txt
cat gpurun_out/micro.txt

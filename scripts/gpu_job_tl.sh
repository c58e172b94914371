cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out; rm -f gpurun_out/micro.txt
timeout 900 python -m pytest tests/test_gpu_bf16.py tests/test_gpu_n2_n3.py tests/test_gpu_parity.py tests/test_gpu_bf16_bound.py tests/test_gpu_reference_callsites.py -x -q 2>&1 | grep -v "Warning\|warnings.warn" | tail -8 >> gpurun_out/micro.txt
cat gpurun_out/micro.txt | cut -c1-300

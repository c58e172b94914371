cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
python scripts/host_gaps.py > gpurun_out/micro.txt 2>&1
PLAIN=1 python scripts/host_timeline.py >> gpurun_out/micro.txt 2>&1
MCPC_SPIN_WAIT=1 PLAIN=1 python scripts/host_timeline.py >> gpurun_out/micro.txt 2>&1
python scripts/host_timeline.py >> gpurun_out/micro.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_bf16.py tests/test_gpu_parity.py tests/test_gpu_bf16_bound.py -x -q 2>&1 | tail -3 >> gpurun_out/micro.txt
cat gpurun_out/micro.txt | tail -50

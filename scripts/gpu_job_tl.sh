cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out; rm -f gpurun_out/micro.txt
for ov in 1 0 1 0; do
echo "== MCPC_TC_DW_OVERLAP=$ov" >> gpurun_out/micro.txt
MCPC_TC_DW_OVERLAP=$ov timeout 120 python scripts/host_timeline.py 2>&1 | tail -16 >> gpurun_out/micro.txt
done
timeout 600 python -m pytest tests/test_gpu_bf16.py -x -q 2>&1 | grep -v "Warning\|warnings.warn" | tail -5 >> gpurun_out/micro.txt
cat gpurun_out/micro.txt

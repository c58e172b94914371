cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out; rm -f gpurun_out/micro.txt
for i in 1 2; do
echo "== new lib" >> gpurun_out/micro.txt
timeout -s KILL 100 python scripts/bench_configs.py c4 2>&1 | grep -o '"ms": [0-9.]*' >> gpurun_out/micro.txt
echo "== old lib (edddb91)" >> gpurun_out/micro.txt
MCPC_NATIVE_LIB=$PWD/tmp_old/libmcpc_b200_old.so timeout -s KILL 100 python scripts/bench_configs.py c4 2>&1 | grep -o '"ms": [0-9.]*' >> gpurun_out/micro.txt
done
cat gpurun_out/micro.txt

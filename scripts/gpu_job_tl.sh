cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out; rm -f gpurun_out/micro.txt
cat > /tmp/c3ab.py <<'PY'
import os, sys
sys.path.insert(0, "scripts"); sys.path.insert(0, ".")
import bench_configs as bc
for l0 in ["0", "1", "2", "0", "1", "2"]:
    os.environ["MCPC_TC_NZ_L0"] = l0
    r = bc.c3("bf16", B=65536, T=1000)
    r2 = bc.c3("bf16", B=1024, T=1000)
    print("nz_l0", l0, "B=65536:", round(r["us_per_step"], 1), "us/step;  B=1024:", round(r2["us_per_step"], 2), "us/step", flush=True)
PY
timeout -s KILL 150 python /tmp/c3ab.py >> gpurun_out/micro.txt 2>&1
cat gpurun_out/micro.txt | cut -c1-300

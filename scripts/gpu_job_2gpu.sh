#!/bin/bash
# 2-GPU check: the NCCL sharded-run test and the bench line at N=2 (as the driver launches it).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/n2.txt
timeout 900 python -m pytest tests/test_gpu_dp_nccl.py -q --timeout 800 2>&1 | grep -E "passed|failed|Error" >> gpurun_out/n2.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps ${N2_STEPS:-100} --warmup 3 --no-cpu-baseline > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
tail -3 gpurun_out/bench_n2.err >> gpurun_out/n2.txt
cat gpurun_out/n2.txt | cut -c1-300
python - <<'PY'
import json
for ln in open('gpurun_out/bench_n2.json'):
    if ln.startswith('{'):
        d=json.loads(ln)
        print({k: d[k] for k in ('value','n_gpus','ms_per_step','ms_median')}, d['e2e']['value'])
        for k,v in (d.get('other_workloads') or {}).items():
            print(k, {kk: v.get(kk) for kk in ('n_gpus','ms_per_step','ms','latent_updates_per_s','frac_of_live_cublas','error')} if isinstance(v, dict) else v)
PY

#!/bin/bash
# C2 kernel iteration: bf16 parity tests of the resident kernel, then the headline line (no CPU arm).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout -s KILL 400 python -m pytest tests/test_gpu_bf16.py tests/test_gpu_bf16_bound.py -q -m gpu --timeout 300 -x 2>&1 | grep -v "Warning\|warnings.warn" | grep -E "^E |passed|failed|Error" | head -30 > gpurun_out/c2_tests.txt
cat gpurun_out/c2_tests.txt
for v in "" $C2_VARIANTS; do
  echo "== $v" | tee -a gpurun_out/c2_tests.txt
  env $v timeout -s KILL 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline $C2_BENCH_ARGS 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print({k:d[k] for k in ('ms_per_step','ms_median')}, 'kernel_ms', d['roofline']['kernel_ms'], 'e2e', d['e2e']['ms_per_step'])
for k,v in d.get('other_workloads',{}).items(): print(k, {a:b for a,b in v.items() if a in ('us_per_step','ms','ms_per_step','frac_of_live_cublas')})" | tee -a gpurun_out/c2_tests.txt
done

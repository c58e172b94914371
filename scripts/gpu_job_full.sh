#!/bin/bash
# Full GPU check: every -m gpu test, smoke(), the bench line.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests -q -m gpu --timeout 300 -x 2>&1 | grep -v "Warning\|warnings.warn" | tail -25 > gpurun_out/full_tests.txt
timeout -s KILL 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 >> gpurun_out/full_tests.txt
timeout -s KILL 400 python bench.py --steps 50 --warmup 3 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err
timeout -s KILL 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
cat gpurun_out/full_tests.txt; tail -3 gpurun_out/bench_ours.err; cat gpurun_out/bench_ours.json | cut -c1-3000; cat gpurun_out/bench_ref.json | cut -c1-600

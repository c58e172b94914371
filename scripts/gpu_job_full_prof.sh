cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests -q -m gpu --timeout 300 -x 2>&1 | grep -v "Warning\|warnings.warn" | tail -6 > gpurun_out/full_tests.txt
timeout -s KILL 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 >> gpurun_out/full_tests.txt
timeout -s KILL 400 python bench.py --steps 200 --warmup 5 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err
TAG=r02d bash scripts/gpu_job_profile.sh > gpurun_out/prof_stdout.txt 2>&1
tail -5 gpurun_out/full_tests.txt; cut -c1-300 gpurun_out/bench_ours.json; tail -2 gpurun_out/prof_stdout.txt

cd "${GRAFT_REPO_ROOT:-/root/repo}"
bash scripts/gpu_job_full.sh > gpurun_out/full_stdout.txt 2>&1
TAG=r02c bash scripts/gpu_job_profile.sh > gpurun_out/prof_stdout.txt 2>&1
tail -5 gpurun_out/full_tests.txt; cut -c1-400 gpurun_out/bench_ours.json; tail -2 gpurun_out/prof_stdout.txt

#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
: > gpurun_out/r02i.txt
export MCPC_NATIVE_LIB=$PWD/montecarlopredictivecoding_b200/libmcpc_b200_debug.so
nvidia-smi --query-gpu=clocks.sm,clocks.mem,power.draw,clocks_event_reasons.active --format=csv -lms 100 > gpurun_out/r02i_smi.csv &
SMI=$!
for mode in 0 2 1; do
  echo "## EPI_MODE=$mode T=100 (sustained)" >> gpurun_out/r02i.txt
  MCPC_C5_T=100 MCPC_WIDE_TIMING=1 MCPC_WIDE_EPI_MODE=$mode timeout 300 python scripts/bench_configs.py c5 2>&1 | grep -E "SM clock|ms_per_step" | tail -4 | cut -c1-240 >> gpurun_out/r02i.txt
  sleep 2
done
kill $SMI
unset MCPC_NATIVE_LIB
python - <<'PY' >> gpurun_out/r02i.txt
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02i_smi.csv'))][1:]
busy=[(float(r[0].split()[0]), float(r[2].split()[0])) for r in rows if float(r[2].split()[0])>400]
import statistics
print("samples under load:", len(busy), "median sm MHz", statistics.median(b[0] for b in busy), "median W", statistics.median(b[1] for b in busy))
# print a coarse trace
for i in range(0,len(rows),5): print(rows[i])
PY
cat gpurun_out/r02i.txt | head -80

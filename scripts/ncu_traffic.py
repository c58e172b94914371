#!/usr/bin/env python
"""Write profiles/ncu_traffic.json: DRAM bytes per launch of a kernel from ONE `ncu --set full` capture, so that
bench.py's roofline.traffic is read from a committed measurement that names its capture (never a literal in bench.py).

    python scripts/ncu_traffic.py <key> <report.ncu-rep> <kernel-regex>
e.g. python scripts/ncu_traffic.py C2_bf16_B1024 gpurun_out/r02_c2_infer_tc.ncu-rep infer_tc_kernel
"""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    key, rep, pattern = sys.argv[1], sys.argv[2], sys.argv[3]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    vals = []
    for r in rows[2:]:
        if not re.search(pattern, r[col["Kernel Name"]]):
            continue
        tot = 0.0
        for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            tot += float(r[col[m]].replace(",", "")) * scale[units[col[m]]]
        dur = float(r[col["gpu__time_duration.sum"]].replace(",", ""))
        dur_unit = units[col["gpu__time_duration.sum"]]
        vals.append((tot, dur, dur_unit, r[col["Kernel Name"]]))
    if not vals:
        raise SystemExit(f"no kernel matching {pattern!r} in {rep}")
    tot = sum(v[0] for v in vals) / len(vals)
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    table = {}
    if os.path.exists(path):
        with open(path) as fh:
            table = json.load(fh)
    table[key] = {"dram_bytes_per_launch": int(round(tot)), "launches_in_capture": len(vals), "kernel": vals[0][3][:120],
                  "duration_under_ncu": f"{vals[0][1]} {vals[0][2]}", "capture": os.path.basename(rep)}
    with open(path, "w") as fh:
        json.dump(table, fh, indent=1, sort_keys=True)
    print(key, table[key])


if __name__ == "__main__":
    main()

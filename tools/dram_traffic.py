"""`python tools/dram_traffic.py profiles/rNN_ncu_full_hot_kernels.csv [first_id last_id] > profiles/dram_traffic.json`

Per-kernel DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum of a `ncu --set full` capture, one row per launch,
as written by tools/ncu_summary.py) keyed by the kernel's base name; bench.py reads this file at run time for
`roofline.traffic` instead of carrying a number in its source.  With [first_id last_id] only the launches of one step
are used and `step_total_bytes` is their sum."""
import csv
import json
import re
import sys

path = sys.argv[1]
lo = int(sys.argv[2]) if len(sys.argv) > 3 else None
hi = int(sys.argv[3]) if len(sys.argv) > 3 else None
rows = list(csv.reader(open(path)))
head, units = rows[0], rows[1]
ci = {c: i for i, c in enumerate(head)}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
t_scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
out, total = {}, 0.0
for r in rows[2:]:
    if not r or not r[0].strip().isdigit():
        continue
    i = int(r[0])
    if lo is not None and not (lo <= i <= hi):
        continue
    name = re.match(r"[\w:]+", r[ci["Kernel Name"]].replace("void ", "").replace("mdeggs::", "")).group(0)
    rd = float(r[ci["dram__bytes_read.sum"]]) * scale[units[ci["dram__bytes_read.sum"]]]
    wr = float(r[ci["dram__bytes_write.sum"]]) * scale[units[ci["dram__bytes_write.sum"]]]
    us = float(r[ci["gpu__time_duration.sum"]]) * t_scale[units[ci["gpu__time_duration.sum"]]]
    e = out.setdefault(name, {"launches": 0, "dram_read_bytes": 0.0, "dram_write_bytes": 0.0, "duration_us": 0.0})
    e["launches"] += 1
    e["dram_read_bytes"] += rd
    e["dram_write_bytes"] += wr
    e["duration_us"] += us
    total += rd + wr
for e in out.values():  # per launch
    for k in ("dram_read_bytes", "dram_write_bytes", "duration_us"):
        e[k] = e[k] / e["launches"]
json.dump({"source": path, "launch_ids": [lo, hi], "per_launch": out,
           "step_total_bytes": total if lo is not None else None}, sys.stdout, indent=1)

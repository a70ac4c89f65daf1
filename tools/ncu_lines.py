"""python tools/ncu_lines.py <report.ncu-rep> <kernel regex> <mangled-name prefix> [top]

Per-source-line instruction counts and stall samples of one kernel: joins `ncu --page source --csv` (SASS rows of the
first profiled launch that matches) with `nvdisasm --print-line-info` of the cubin inside libmdeggs.so."""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

rep, kre, mangled = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(root, "multideggs_b200/csrc/libmdeggs.so")], cwd=tmp,
               capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
lines = dis.split("\n")
start = next(i for i, l in enumerate(lines) if ".section" in l and ".text." + mangled in l)
cur, byoff = None, {}
for l in lines[start + 1:]:
    if ".section" in l:
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", l)
    if m:
        byoff[int(m.group(1), 16)] = (cur, m.group(2))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(src.split("\n")))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
line_ie, line_ss = collections.Counter(), collections.Counter()
base = None
for r in rows[2:]:
    if not r or not r[0].startswith("0x"):
        break  # next launch
    a = int(r[0], 16)
    base = a if base is None else base
    c = byoff.get(a - base, (None, ""))[0]
    line_ie[c] += int(r[ix["Instructions Executed"]])
    line_ss[c] += int(r[ix["# Samples"]])
tot, tots = sum(line_ie.values()), sum(line_ss.values())
print(f"total warp instructions {tot}, stall samples {tots}")
for c, v in line_ie.most_common(top):
    print(f"{str(c):40s} {v:10d} {100 * v / tot:5.1f} %   samples {100 * line_ss[c] / max(tots, 1):5.1f} %")

#!/usr/bin/env python
"""Per-launch timeline of one hot-path call (MDEGGS trace option): python tools/trace.py [cfg3|cfg5|...] [host]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import multideggs_b200 as md

name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
host = len(sys.argv) > 2 and sys.argv[2] == "host"
prob = bench._workload(name).problem()
dev = torch.device("cuda", 0)
eng = md.Engine([0])
if host:
    cprob, cres, keep, _, _ = bench._host_problem(torch, prob)
else:
    cprob, cres, keep = bench._device_problem(torch, prob, dev)
flush = torch.zeros(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
for _ in range(3):
    eng.run_omic_raw(cprob, cres)
eng.set_option("trace", 1)
for i in range(2):
    flush.add_(1.0)
    torch.cuda.synchronize()
    sys.stderr.write(f"--- traced call {i} ({name}{', host buffers' if host else ''}) ---\n")
    eng.run_omic_raw(cprob, cres)
eng.set_option("trace", 0)
print({k: round(v, 4) if isinstance(v, float) else v for k, v in eng.stats().items()})

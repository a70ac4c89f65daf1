#!/usr/bin/env python
"""A/B of the zero-copy input path against the H2D band copy, same process, alternating rounds."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
import multideggs_b200 as md
prob = bench._workload("cfg3").problem()
eng = md.Engine([0])
hprob, hres, hkeep, h2d, d2h = bench._host_problem(torch, prob)
res = {0: [], 1: []}
for rnd in range(6):
    for z in (0, 1):
        eng.set_option("zero_copy", z)
        for _ in range(3):
            eng.run_omic_raw(hprob, hres)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(20):
            eng.run_omic_raw(hprob, hres)
        torch.cuda.synchronize()
        res[z].append((time.perf_counter() - t0) / 20 * 1e3)
for z in (0, 1):
    print("zero_copy", z, "ms/step per round:", [round(v, 3) for v in res[z]], "median", round(float(np.median(res[z])), 3))

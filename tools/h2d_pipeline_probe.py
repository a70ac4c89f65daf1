#!/usr/bin/env python
"""A/B of the streamed input copy (option h2d_pipeline: chunk size in MB, 0 = copy first), same process, alternating."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
import multideggs_b200 as md
prob = bench._workload("cfg3").problem()
eng = md.Engine([0])
hprob, hres, hkeep, h2d, d2h = bench._host_problem(torch, prob)
modes = [int(a) for a in sys.argv[1:]] or [0, 1006, 2003, 2006, 2012]   # copy streams * 1000 + chunk MB
res = {m: [] for m in modes}
for rnd in range(5):
    for m in modes:
        eng.set_option("h2d_pipeline", m % 1000)
        eng.set_option("copy_streams", max(1, m // 1000))
        for _ in range(3):
            eng.run_omic_raw(hprob, hres)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(20):
            eng.run_omic_raw(hprob, hres)
        torch.cuda.synchronize()
        res[m].append((time.perf_counter() - t0) / 20 * 1e3)
for m in modes:
    print("h2d_pipeline", m, "ms/step per round:", [round(v, 3) for v in res[m]], "median", round(float(np.median(res[m])), 3),
          "fit_path", hex(int(eng.stats()["fit_path"])))
eng.set_option("trace", 1)
eng.set_option("h2d_pipeline", 1)
eng.run_omic_raw(hprob, hres)

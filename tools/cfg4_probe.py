#!/usr/bin/env python
"""Config 4 through the public API: 10 folds + final call of multiDEGGs_filter on 5,000 x 24; where does the time go?"""
import os, sys, time, cProfile, pstats
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, pandas as pd
import multideggs_b200 as md
from multideggs_b200 import synth
wl = synth.cfg4()
x = pd.DataFrame(wl.assay.values.T, index=wl.assay.samples, columns=wl.assay.genes)
y = wl.group.astype(float)
eng = md.Engine([0])
rng = np.random.default_rng(4)
folds = np.array_split(rng.permutation(24), 10)
def run_all():
    out = []
    for hold in [None] + folds:
        keep = np.ones(24, dtype=bool)
        if hold is not None:
            keep[hold] = False
        try:
            r = md.multiDEGGs_filter(y[keep], x.iloc[keep], engine=eng)
            out.append(r["pairs"].shape[0])
        except ValueError as e:
            out.append(str(e)[:30])
    return out
run_all()
t0 = time.perf_counter(); res = run_all(); dt = time.perf_counter() - t0
print("11 calls:", round(dt * 1e3, 1), "ms", res)
print("engine ms_total of last call:", eng.stats()["ms_total"])
pr = cProfile.Profile(); pr.enable(); run_all(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)

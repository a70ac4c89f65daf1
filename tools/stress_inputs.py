#!/usr/bin/env python
"""Stress of the host input paths: alternating problem sizes / paths (plain copy, streamed band, host packing, device
buffers) on ONE context, many repetitions; every result must equal the first result of its problem bit for bit."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import multideggs_b200 as md
from multideggs_b200 import synth, _abi
want = ("p_edge", "status", "cutoff", "median", "best", "sig_count", "tot_edges", "padj_best", "cat_best")
probs = [synth.random_problem(6000, 500, 5000, 3, 1), synth.random_problem(900, 60, 700, 2, 2),
         synth.random_problem(2600, 700, 4000, 2, 3), synth.random_problem(4000, 200, 3000, 2, 4, _abi.METHOD_RLM),
         synth.random_problem(3000, 400, 60000, 2, 5)]
eng = md.Engine([0])
ref = []
eng.set_option("host_pack", 0); eng.set_option("h2d_pipeline", 0)
for p in probs:
    ref.append(eng.run_omic(p, want))
rng = np.random.default_rng(0)
n_bad = 0
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 300):
    i = int(rng.integers(len(probs)))
    eng.set_option("host_pack", int(rng.integers(3)))
    eng.set_option("h2d_pipeline", int(rng.choice([0, 1, 2, 5])))
    eng.set_option("graph", int(rng.integers(2)))
    for rep in range(int(rng.integers(1, 4))):
        got = eng.run_omic(probs[i], want)
        for k in want:
            if not np.array_equal(got[k], ref[i][k], equal_nan=True):
                n_bad += 1
                print("MISMATCH", it, i, k, hex(int(eng.stats()["fit_path"])))
print("stress done, mismatches:", n_bad)
sys.exit(1 if n_bad else 0)

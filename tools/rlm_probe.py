#!/usr/bin/env python
"""rlm stress leg alone (for ncu): cfg3 matrix/network, 2 shuffled categories."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
import multideggs_b200 as md
from multideggs_b200 import _abi as A
prob = bench._workload("cfg3").problem()
n = prob.dims[1]
rng = np.random.default_rng(7)
grp2 = (np.arange(n) % 2 + 1).astype(np.int32)
rng.shuffle(grp2)
rprob = A.Problem(assay=prob.assay, from_idx=prob.from_idx, to_idx=prob.to_idx, group=grp2, K=2, pct=prob.pct,
                  method=A.METHOD_RLM, padj=A.PADJ_BH)
eng = md.Engine([0])
eng.set_option("graph", 0)
for _ in range(3):
    eng.run_omic(rprob, ("iters", "status", "best", "tot_edges"))
    print(eng.stats()["ms_fit"])

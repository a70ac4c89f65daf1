"""Config 1 (bundled RNAseq / Proteomics / Olink layers, lm + bonferroni): the three layers one after the other
on one context vs. submitted together on three contexts (mdeggs_run_omics). Prints microseconds per layer set."""
import ctypes as C
import time

import numpy as np

import multideggs_b200 as md
import os

import pandas as pd

from multideggs_b200 import _abi, host
from multideggs_b200.engine import EnginePool


def main():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    d = np.load(os.path.join(root, "tests", "golden", "bundled.npz"), allow_pickle=False)
    b = {}
    for key, name in (("rnaseq", "RNAseq"), ("proteomic", "Proteomics"), ("olink", "Olink")):
        b[name] = pd.DataFrame(d[f"{key}_x"], index=d[f"{key}_genes"], columns=d[f"{key}_samples"])
    b["metadata"] = pd.DataFrame({
        "response": pd.Categorical(d["meta_response"], categories=list(d["meta_response_levels"]))},
        index=d["meta_samples"])
    fac = host.tidy_metadata(metadata=b["metadata"], category_variable="response", verbose=False)
    pct = _abi.r_seq(0.35, 0.98, 0.05)
    for method, padj in (("lm", "bonferroni"), ("rlm", "BH")):
        preps = [host.prepare_omic(b[k], k, fac, method, None, pct, padj) for k in ("RNAseq", "Proteomics", "Olink")]
        probs = [p.problem for p in preps]
        wants = [host.want_fields(p) for p in preps]
        eng = md.Engine([0])
        pool = EnginePool(0)
        cps = [p.to_c() for p in probs]
        crs = [_abi.alloc_result(p.dims, tuple(w)) for p, w in zip(probs, wants)]
        n = len(probs)
        engs = pool.engines(n)
        ctxs = (C.c_void_p * n)(*[e._ctx for e in engs])
        cprobs = (_abi.CProblem * n)(*cps)
        cress = (_abi.CResult * n)(*[c[0] for c in crs])
        rcs = (C.c_int * n)()
        lib = eng._lib

        def seq():
            for cp, cr in zip(cps, crs):
                eng.run_omic_raw(cp, cr[0])

        def bat():
            lib.mdeggs_run_omics(ctxs, n, cprobs, cress, rcs)

        for name, fn in (("sequential", seq), ("batched", bat)):
            for _ in range(20):
                fn()
            t = []
            for _ in range(200):
                t0 = time.perf_counter()
                fn()
                t.append(time.perf_counter() - t0)
            print(f"{method}/{padj} {name}: median {np.median(t) * 1e6:.1f} us per 3-layer set, min {np.min(t) * 1e6:.1f}")
        pool.close()
        eng.close()


if __name__ == "__main__":
    main()

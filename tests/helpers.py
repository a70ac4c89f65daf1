"""Shared test helpers: run the host flow with the ORACLE as the compute (CPU tests), compare result dicts."""
import numpy as np

import pyoracle
from multideggs_b200 import _abi, host

ALL_FIELDS = tuple(_abi.RESULT_FIELDS)


def oracle_single_omic(assay, name, factor, method, network, pct, padj, cores=1, **kw):
    prep = host.prepare_omic(assay, name, factor, method, network, pct, padj, **kw)
    res = pyoracle.run_omic(prep.problem, host.want_fields(prep))
    return host.assemble_networks(prep, res, cores=cores)


def rel_close(a, b, rtol, atol=0.0):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    both_nan = np.isnan(a) & np.isnan(b)
    ok = both_nan | (np.abs(a - b) <= atol + rtol * np.maximum(np.abs(a), np.abs(b)))
    return ok


def assert_parity(gpu: dict, ref: dict, problem: _abi.Problem, rtol=1e-9, what=""):
    """Integer / index / selection outputs bit-exact; statistics within 1e-9 relative (north_star).
    lm p-values carry the reference's own 2.2e-16 quantisation (2*(1-pt)), hence the absolute floor."""
    exact = ("status", "cat", "cat_best", "tot_edges", "sig_count", "fail_count", "best", "iters")
    for k in exact:
        if k in gpu and k in ref:
            assert np.array_equal(gpu[k], ref[k]), f"{what}{k} differs: {np.flatnonzero(np.ravel(gpu[k] != ref[k]))[:10]}"
    for k in ("cutoff", "median", "df"):
        if k in gpu and k in ref:
            assert np.array_equal(gpu[k], ref[k], equal_nan=True), f"{what}{k} not bit-exact"
    p_floor = 4.5e-16 if (problem.K == 2 and problem.method == _abi.METHOD_LM) else 0.0
    tol = {"stat": (rtol, 0.0), "coef": (rtol, 1e-12), "p_edge": (rtol, p_floor), "padj": (rtol, p_floor * 1e7),
           "padj_best": (rtol, p_floor * 1e7)}
    for k, (rt, at) in tol.items():
        if k in gpu and k in ref:
            ok = rel_close(gpu[k], ref[k], rt, at)
            assert ok.all(), (f"{what}{k}: {int((~ok).sum())} of {ok.size} outside tolerance; first "
                              f"{np.ravel(gpu[k])[np.flatnonzero(~np.ravel(ok))[:3]]} vs "
                              f"{np.ravel(ref[k])[np.flatnonzero(~np.ravel(ok))[:3]]}")

"""ctypes front-end of the CPU restatement `libmdeggs_oracle.so`.

TEST INFRASTRUCTURE ONLY (see the header of oracle/mdeggs_oracle.c): imported by tests/,
`__graft_entry__.smoke()` and bench.py's cpu_baseline / `--impl reference` legs, never by the product.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
from multideggs_b200 import _abi  # noqa: E402  (struct layout only)

MODE_UNIQUE, MODE_REFERENCE = 0, 1
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libmdeggs_oracle.so")
    src = os.path.join(_HERE, "mdeggs_oracle.c")
    stale = (not os.path.exists(so)) or os.path.getmtime(so) < os.path.getmtime(src)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B", "libmdeggs_oracle.so"], check=True, capture_output=True)
    return so


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        d, i, vp = C.c_double, C.c_int, C.c_void_p
        L.mdeggs_oracle_run_omic.argtypes = [C.POINTER(_abi.CProblem), C.POINTER(_abi.CResult), i, i]
        L.mdeggs_oracle_run_omic.restype = i
        L.mdeggs_oracle_threads.restype = i
        L.orc_quantile7_sorted.argtypes = [vp, C.c_int64, d]
        L.orc_quantile7_sorted.restype = d
        L.orc_seq.argtypes = [d, d, d, vp, i]
        L.orc_seq.restype = i
        L.orc_median.argtypes = [vp, C.c_int64, C.c_int64, vp]
        L.orc_median.restype = d
        for nm, na in (("orc_lbeta", 2), ("orc_pt", 2), ("orc_p_two_sided_ref", 2), ("orc_p_two_sided_true", 2),
                       ("orc_pf_upper", 3)):
            getattr(L, nm).argtypes = [d] * na
            getattr(L, nm).restype = d
        L.orc_pbeta.argtypes = [d, d, d, i]
        L.orc_pbeta.restype = d
        L.orc_pbeta_xy.argtypes = [d, d, d, d, i]
        L.orc_pbeta_xy.restype = d
        L.orc_fit_lm_interaction.argtypes = [vp, vp, vp, i, vp, vp, vp, vp]
        L.orc_fit_lm_interaction.restype = i
        L.orc_fit_aov.argtypes = [vp, vp, vp, i, i, vp, vp, vp, vp]
        L.orc_fit_aov.restype = i
        L.orc_fit_rlm.argtypes = [vp, vp, vp, i, i, i, vp, vp, vp, vp, vp]
        L.orc_fit_rlm.restype = i
        L.orc_p_adjust.argtypes = [vp, C.c_int64, i, vp]
        L.orc_p_adjust.restype = None
        _LIB = L
    return _LIB


def _qvalue_smoother_last(x: np.ndarray, y: np.ndarray, df: float = 3.0) -> float:
    """Value at the last knot of the natural cubic smoothing spline (Reinsch form) whose smoother has trace df:
    restatement of what qvalue::pi0est(pi0.method = "smoother") asks of stats::smooth.spline(lambda, pi0, df = 3).
    Solves (I + a K) f = y directly for the a that gives trace df (bisection on a)."""
    n = x.shape[0]
    h = np.diff(x)
    Q = np.zeros((n, n - 2))
    R = np.zeros((n - 2, n - 2))
    for j in range(n - 2):
        Q[j:j + 3, j] = [1.0 / h[j], -1.0 / h[j] - 1.0 / h[j + 1], 1.0 / h[j + 1]]
        R[j, j] = (h[j] + h[j + 1]) / 3.0
        if j + 1 < n - 2:
            R[j, j + 1] = R[j + 1, j] = h[j + 1] / 6.0
    K = Q @ np.linalg.solve(R, Q.T)
    eye = np.eye(n)
    lo, hi = 1e-12, 1e12
    for _ in range(200):
        mid = np.sqrt(lo * hi)
        if np.trace(np.linalg.inv(eye + mid * K)) > df:
            lo = mid
        else:
            hi = mid
    return float(np.linalg.solve(eye + np.sqrt(lo * hi) * K, y)[-1])


def qvalue(p: np.ndarray) -> np.ndarray:
    """`padj_method = "q.value"` of calc_pvalues_network (core_functions.R:1012-1025): qvalue::qvalue(p)$qvalues
    (pi0 by the smoother over lambda = seq(0.05, 0.95, 0.05); qvalues = pi0 * pmin(1, cummin(p[o] * m / i)) with o the
    decreasing order); on any qvalue() error retry with pi0 = 1 when there is more than one row, else NA."""
    p = np.asarray(p, dtype=np.float64)
    ok = ~np.isnan(p)
    pv = p[ok]
    m = pv.shape[0]

    def q_of(pi0):
        o = np.argsort(-pv, kind="stable")
        cm = np.minimum.accumulate(pv[o] * m / np.arange(m, 0, -1, dtype=np.float64))
        q = np.empty(m)
        q[o] = pi0 * np.minimum(1.0, cm)
        out = np.full(p.shape[0], np.nan)
        out[ok] = q
        return out

    lam = seq(0.05, 0.95, 0.05)
    good = m >= 2 and pv.min() >= 0 and pv.max() <= 1 and pv.max() >= lam.max()
    if good:
        pi0_l = np.array([np.mean(pv >= l) / (1.0 - l) for l in lam])
        pi0 = min(_qvalue_smoother_last(lam, pi0_l), 1.0)
        good = pi0 > 0
    if good:
        return q_of(pi0)
    return q_of(1.0) if p.shape[0] > 1 else np.full(p.shape[0], np.nan)


def _run_omic_qvalue(problem: _abi.Problem, want, mode: int, nthreads: int) -> dict[str, np.ndarray]:
    """PADJ_QVALUE: the C restatement returns raw p-values and the category of every edge at every percentile; the
    q-values, the significant-pair counts and the best percentile (core_functions.R:796-827, 471-495) follow here."""
    import copy
    raw = copy.copy(problem)
    raw.padj = _abi.PADJ_HOST
    need = tuple(dict.fromkeys(tuple(w for w in want if w not in ("padj", "padj_best", "cat_best", "sig_count", "best"))
                               + ("p_edge", "cat", "tot_edges", "fail_count")))
    res = run_omic(raw, need, mode, nthreads)
    G, n, E, K, P = problem.dims
    p_edge, cat, tot, fail = res["p_edge"], res["cat"].reshape(P, E), res["tot_edges"], res["fail_count"]
    padj = np.full((P, E), np.nan)
    sig = np.full(P, -1, dtype=np.int64)
    for pi in range(P):
        if tot[pi] <= 0:
            continue
        cnt = 0
        for k in range(1, K + 1):
            msk = cat[pi] == k
            if msk.any():
                padj[pi, msk] = qvalue(p_edge[msk])
                cnt += int(np.sum(padj[pi, msk] < 0.05))
        sig[pi] = cnt
    best, bv = 0, -1
    for pi in range(P):
        if tot[pi] > 0 and fail[pi] == 0 and sig[pi] > bv:
            best, bv = pi + 1, sig[pi]
    out = dict(res)
    out["padj"] = padj
    out["sig_count"] = sig
    out["best"] = np.array([best], dtype=np.int32)
    out["padj_best"] = padj[best - 1].copy() if best > 0 else np.full(E, np.nan)
    out["cat_best"] = cat[best - 1].copy() if best > 0 else np.zeros(E, dtype=np.int8)
    return {k: v for k, v in out.items() if k in want}


def run_omic(problem: _abi.Problem, want=tuple(_abi.RESULT_FIELDS), mode: int = MODE_UNIQUE,
             nthreads: int = 0) -> dict[str, np.ndarray]:
    if problem.padj == _abi.PADJ_QVALUE:
        return _run_omic_qvalue(problem, tuple(want), mode, nthreads)
    cprob = problem.to_c()
    cres, arrays = _abi.alloc_result(problem.dims, want)
    rc = lib().mdeggs_oracle_run_omic(C.byref(cprob), C.byref(cres), mode, nthreads)
    if rc != 0:
        raise RuntimeError(f"oracle returned {rc}")
    return arrays


def _d(x) -> np.ndarray:
    return np.ascontiguousarray(x, dtype=np.float64)


def fit_lm(a, b, g01):
    a, b, g = _d(a), _d(b), _d(g01)
    n = a.shape[0]
    ws = np.empty(6 * n)
    coef = np.empty(4)
    t = C.c_double()
    p = C.c_double()
    st = lib().orc_fit_lm_interaction(a.ctypes.data, b.ctypes.data, g.ctypes.data, n, ws.ctypes.data,
                                      coef.ctypes.data, C.addressof(t), C.addressof(p))
    return st, coef, t.value, p.value


def fit_aov(a, b, grp0, K):
    a, b = _d(a), _d(b)
    g = np.ascontiguousarray(grp0, dtype=np.int32)
    n = a.shape[0]
    ws = np.empty((2 * K + 2) * n)
    coef = np.empty(2 * K)
    f = C.c_double()
    p = C.c_double()
    st = lib().orc_fit_aov(a.ctypes.data, b.ctypes.data, g.ctypes.data, n, K, ws.ctypes.data, coef.ctypes.data,
                           C.addressof(f), C.addressof(p))
    return st, coef, f.value, p.value


def fit_rlm(a, b, g01, tested_coef=3, cov_mode=_abi.RLM_COV_XTWX):
    a, b, g = _d(a), _d(b), _d(g01)
    n = a.shape[0]
    ws = np.empty(9 * n)
    coef = np.empty(4)
    f = C.c_double()
    p = C.c_double()
    it = C.c_int()
    st = lib().orc_fit_rlm(a.ctypes.data, b.ctypes.data, g.ctypes.data, n, tested_coef, cov_mode, ws.ctypes.data,
                           coef.ctypes.data, C.addressof(f), C.addressof(p), C.addressof(it))
    return st, coef, f.value, p.value, it.value


def p_adjust(p, method: int) -> np.ndarray:
    p = _d(p)
    out = np.empty_like(p)
    lib().orc_p_adjust(p.ctypes.data, p.shape[0], method, out.ctypes.data)
    return out


def quantile7(values, probs) -> np.ndarray:
    v = _d(values).ravel()
    v = np.sort(v[~np.isnan(v)])
    return np.array([lib().orc_quantile7_sorted(v.ctypes.data, v.shape[0], float(q)) for q in probs])


def median(x) -> float:
    x = _d(x)
    scr = np.empty_like(x)
    return lib().orc_median(x.ctypes.data, x.shape[0], 1, scr.ctypes.data)


def seq(frm, to, by) -> np.ndarray:
    out = np.empty(4096)
    n = lib().orc_seq(frm, to, by, out.ctypes.data, 4096)
    return out[:n].copy()

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


def run_omic(problem: _abi.Problem, want=tuple(_abi.RESULT_FIELDS), mode: int = MODE_UNIQUE,
             nthreads: int = 0) -> dict[str, np.ndarray]:
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

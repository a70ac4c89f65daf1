"""ctypes mirror of include/mdeggs.h (structs, constants) and numpy marshalling helpers.

Pure description of the C ABI: no compute, no fallback.  `Problem`/`Result` hold references to the
numpy arrays whose pointers are placed in the C structs, so nothing is freed while a call runs.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Any

import numpy as np

ABI_VERSION = 1
MAX_K = 8

METHOD_LM, METHOD_RLM = 0, 1
PADJ_NONE, PADJ_BONFERRONI, PADJ_BH, PADJ_HOST, PADJ_HOLM, PADJ_HOCHBERG, PADJ_BY, PADJ_QVALUE, PADJ_HOMMEL = range(9)
DEVICE_PADJ = (PADJ_BONFERRONI, PADJ_BH, PADJ_HOLM, PADJ_HOCHBERG, PADJ_BY, PADJ_QVALUE, PADJ_HOMMEL)
HOMMEL_MAX_E = 65536   # MDEGGS_HOMMEL_MAX_E: hommel is O(n^2) per segment, larger jobs are adjusted by the host
RLM_COV_XTWX, RLM_COV_XTX = 0, 1
LAYOUT_COLMAJOR, LAYOUT_ROWMAJOR = 0, 1
MEM_HOST, MEM_DEVICE = 0, 1
EDGE_OK, EDGE_SINGULAR, EDGE_NONFINITE, EDGE_NOT_CONVERGED, EDGE_ZERO_SCALE, EDGE_SELFLOOP, EDGE_ALIASED, EDGE_NA_OMIT = 0, 1, 2, 3, 4, 5, 6, 7

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_bp = C.POINTER(C.c_int8)
_lp = C.POINTER(C.c_int64)


class CProblem(C.Structure):
    _fields_ = [
        ("assay", C.c_void_p), ("ld", C.c_int64), ("G", C.c_int32), ("n", C.c_int32),
        ("layout", C.c_int32), ("mem", C.c_int32),
        ("from_", C.c_void_p), ("to", C.c_void_p), ("E", C.c_int64),
        ("group", C.c_void_p), ("K", C.c_int32),
        ("pct", C.c_void_p), ("P", C.c_int32),
        ("method", C.c_int32), ("padj", C.c_int32), ("tested_coef", C.c_int32), ("rlm_cov_mode", C.c_int32),
    ]


class CResult(C.Structure):
    _fields_ = [(nm, C.c_void_p) for nm in (
        "cutoff", "median", "p_edge", "stat", "coef", "df", "status", "iters", "cat", "padj",
        "tot_edges", "sig_count", "fail_count", "best", "cat_best", "padj_best")]


class CStats(C.Structure):
    _fields_ = [
        ("n_nodes", C.c_int64), ("unique_fits", C.c_int64), ("ref_equiv_tests", C.c_int64),
        ("kernel_launches", C.c_int64), ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64),
        ("fit_path", C.c_int32), ("n_ranks", C.c_int32),
        ("ms_total", C.c_float), ("ms_h2d", C.c_float), ("ms_gather", C.c_float), ("ms_stats", C.c_float),
        ("ms_quantile", C.c_float), ("ms_fit", C.c_float), ("ms_tail", C.c_float),
        ("ms_collective", C.c_float), ("ms_percolate", C.c_float), ("ms_adjust", C.c_float),
        ("ms_d2h", C.c_float), ("fit_bytes", C.c_int64),
    ]

    def asdict(self) -> dict[str, Any]:
        return {nm: getattr(self, nm) for nm, _ in self._fields_}


# name -> (dtype, shape function(G, n, E, K, P))
RESULT_FIELDS = {
    "cutoff": (np.float64, lambda G, n, E, K, P: (P,)),
    "median": (np.float64, lambda G, n, E, K, P: (K, G)),
    "p_edge": (np.float64, lambda G, n, E, K, P: (E,)),
    "stat": (np.float64, lambda G, n, E, K, P: (E,)),
    "coef": (np.float64, lambda G, n, E, K, P: (E, 2 * K)),
    "df": (np.float64, lambda G, n, E, K, P: (2,)),
    "status": (np.int32, lambda G, n, E, K, P: (E,)),
    "iters": (np.int32, lambda G, n, E, K, P: (E,)),
    "cat": (np.int8, lambda G, n, E, K, P: (P, E)),
    "padj": (np.float64, lambda G, n, E, K, P: (P, E)),
    "tot_edges": (np.int64, lambda G, n, E, K, P: (P,)),
    "sig_count": (np.int64, lambda G, n, E, K, P: (P,)),
    "fail_count": (np.int64, lambda G, n, E, K, P: (P,)),
    "best": (np.int32, lambda G, n, E, K, P: (1,)),
    "cat_best": (np.int8, lambda G, n, E, K, P: (E,)),
    "padj_best": (np.float64, lambda G, n, E, K, P: (E,)),
}
DEFAULT_WANT = ("cutoff", "median", "p_edge", "stat", "status", "df", "tot_edges", "sig_count", "fail_count", "best",
                "cat_best", "padj_best")


def _ptr(a: np.ndarray | None) -> int | None:
    return None if a is None else a.ctypes.data


@dataclass
class Problem:
    """Host-side description of one omic layer, marshalled into `CProblem`."""

    assay: np.ndarray            # 2-D float64, genes x samples; F-order => COLMAJOR, C-order => ROWMAJOR
    from_idx: np.ndarray         # int32 1-based
    to_idx: np.ndarray
    group: np.ndarray            # int32 1..K per sample
    K: int
    pct: np.ndarray              # float64
    method: int = METHOD_LM
    padj: int = PADJ_BONFERRONI
    tested_coef: int = 3
    rlm_cov_mode: int = RLM_COV_XTWX
    _keep: list = field(default_factory=list, repr=False)

    def to_c(self) -> CProblem:
        a = self.assay
        if a.dtype != np.float64 or a.ndim != 2:
            raise ValueError("assay must be a 2-D float64 array (genes x samples)")
        G, n = a.shape
        if a.flags.f_contiguous and not (a.flags.c_contiguous and G > 1 and n > 1):
            layout, ld = LAYOUT_COLMAJOR, max(G, 1)
        elif a.flags.c_contiguous:
            layout, ld = LAYOUT_ROWMAJOR, max(n, 1)
        else:
            a = np.asfortranarray(a)
            layout, ld = LAYOUT_COLMAJOR, max(G, 1)
        fr = np.ascontiguousarray(self.from_idx, dtype=np.int32)
        to = np.ascontiguousarray(self.to_idx, dtype=np.int32)
        gr = np.ascontiguousarray(self.group, dtype=np.int32)
        pc = np.ascontiguousarray(self.pct, dtype=np.float64)
        if gr.shape[0] != n:
            raise ValueError("group must have one code per sample")
        self._keep = [a, fr, to, gr, pc]
        return CProblem(_ptr(a), ld, G, n, layout, MEM_HOST, _ptr(fr), _ptr(to), fr.shape[0], _ptr(gr), int(self.K),
                        _ptr(pc), pc.shape[0], int(self.method), int(self.padj), int(self.tested_coef),
                        int(self.rlm_cov_mode))

    @property
    def dims(self) -> tuple[int, int, int, int, int]:
        G, n = self.assay.shape
        return G, n, int(np.asarray(self.from_idx).shape[0]), int(self.K), int(np.asarray(self.pct).shape[0])


def alloc_result(dims: tuple[int, int, int, int, int], want=DEFAULT_WANT) -> tuple[CResult, dict[str, np.ndarray]]:
    """Allocate caller-owned numpy output buffers for the requested fields."""
    G, n, E, K, P = dims
    arrays: dict[str, np.ndarray] = {}
    cres = CResult()
    for nm in want:
        dt, shp = RESULT_FIELDS[nm]
        shape = shp(G, n, E, K, P)
        if dt is np.float64:
            arr = np.full(shape, np.nan, dtype=dt)
        else:
            arr = np.zeros(shape, dtype=dt)
        arrays[nm] = arr
        setattr(cres, nm, _ptr(arr) if arr.size else None)
    return cres, arrays


def r_seq(frm: float, to: float, by: float) -> np.ndarray:
    """base::seq(from, to, by) in double arithmetic (core_functions.R:137; fs:86): from + k*by, pmin(., to)."""
    kmax = int(np.floor((to - frm) / by + 1e-10))
    return np.array([min(frm + k * by, to) for k in range(kmax + 1)], dtype=np.float64)

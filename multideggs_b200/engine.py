"""ctypes binding of libmdeggs.so (the CUDA engine behind include/mdeggs.h).

There is deliberately NO fallback: if the shared library is missing, cannot be loaded, or no CUDA
device is usable, `Engine()` raises.  The library is built in-tree by `__graft_entry__.build()` /
`python -m multideggs_b200.build` into multideggs_b200/csrc/libmdeggs.so.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Sequence

import numpy as np

from . import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MDEGGS_LIB") or os.path.join(_HERE, "csrc", "libmdeggs.so")  # MDEGGS_LIB: A/B builds
_LIB: C.CDLL | None = None

# every symbol include/mdeggs.h declares (tests check the library exports all of them)
ABI_SYMBOLS = (
    "mdeggs_create", "mdeggs_create_rank", "mdeggs_nccl_unique_id", "mdeggs_destroy", "mdeggs_run_omic",
    "mdeggs_pair_features", "mdeggs_pair_moments", "mdeggs_host_pack_rows", "mdeggs_measure_fp64_peak", "mdeggs_measure_l2_peak", "mdeggs_last_stats",
    "mdeggs_trace_count", "mdeggs_trace_get", "mdeggs_set_option", "mdeggs_last_error", "mdeggs_strerror", "mdeggs_abi_version",
    "mdeggs_dev_pt", "mdeggs_dev_pf_upper", "mdeggs_dev_p_two_sided_ref", "mdeggs_dev_p_two_sided_tab",
    "mdeggs_submit_omic", "mdeggs_wait", "mdeggs_run_omics",
)


class EngineError(RuntimeError):
    pass


def host_pack_rows(assay: np.ndarray, from_idx: np.ndarray, to_idx: np.ndarray, chunk_rows: int = 0):
    """The packing step of option "host_pack" on its own (host threads only, no GPU): the network-node rows of a
    column-major matrix gathered in ascending row order. Returns (packed [nodes x n, Fortran order], rows 0-based)."""
    lib = load_library()
    a = np.asfortranarray(assay, dtype=np.float64)
    G, n = a.shape
    f = np.ascontiguousarray(from_idx, dtype=np.int32)
    t = np.ascontiguousarray(to_idx, dtype=np.int32)
    nn = C.c_int32(0)
    rc = lib.mdeggs_host_pack_rows(a.ctypes.data, G, G, n, f.ctypes.data, t.ctypes.data, f.shape[0], 0, None, 0, None,
                                   C.byref(nn))
    if rc != 0:
        raise EngineError(f"mdeggs_host_pack_rows failed: {lib.mdeggs_strerror(rc).decode()}")
    out = np.full((nn.value, n), np.nan, dtype=np.float64, order="F")
    rows = np.zeros(nn.value, dtype=np.int32)
    if nn.value:
        rc = lib.mdeggs_host_pack_rows(a.ctypes.data, G, G, n, f.ctypes.data, t.ctypes.data, f.shape[0], int(chunk_rows),
                                       out.ctypes.data, nn.value, rows.ctypes.data, C.byref(nn))
        if rc != 0:
            raise EngineError(f"mdeggs_host_pack_rows failed: {lib.mdeggs_strerror(rc).decode()}")
    return out, rows


class OmicError(EngineError):
    """The engine rejected THIS omic layer (MDEGGS_EINVAL / EUNSUPPORTED / ENOMEM: n <= 2K, rlm rows too long for shared
    memory, an edge index outside 1..G ...).  The host treats it like any other per-omic error of the reference's
    tryCatch (core_functions.R:219-233): message + NULL layer.  CUDA / NCCL failures and a missing library stay plain
    EngineError and are never swallowed."""


_LAYER_CODES = (-1, -4, -5)  # MDEGGS_EINVAL, MDEGGS_ENOMEM, MDEGGS_EUNSUPPORTED


def load_library() -> C.CDLL:
    """dlopen libmdeggs.so and declare prototypes. Raises EngineError if it is not built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise EngineError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    vp, i, i64 = C.c_void_p, C.c_int, C.c_int64
    L.mdeggs_create.argtypes = [C.POINTER(vp), C.POINTER(C.c_int), i]
    L.mdeggs_create_rank.argtypes = [C.POINTER(vp), i, i, i, vp]
    L.mdeggs_nccl_unique_id.argtypes = [vp]
    L.mdeggs_destroy.argtypes = [vp]
    L.mdeggs_run_omic.argtypes = [vp, C.POINTER(_abi.CProblem), C.POINTER(_abi.CResult)]
    L.mdeggs_submit_omic.argtypes = [vp, C.POINTER(_abi.CProblem), C.POINTER(_abi.CResult)]
    L.mdeggs_wait.argtypes = [vp]
    L.mdeggs_run_omics.argtypes = [C.POINTER(vp), i, C.POINTER(_abi.CProblem), C.POINTER(_abi.CResult),
                                   C.POINTER(C.c_int)]
    L.mdeggs_pair_features.argtypes = [vp, vp, i64, C.c_int32, vp, vp, C.c_int32, C.c_int32, C.c_int32, vp]
    L.mdeggs_pair_moments.argtypes = [vp, vp, vp, C.c_int32, vp]
    L.mdeggs_host_pack_rows.argtypes = [vp, C.c_int64, C.c_int32, C.c_int32, vp, vp, C.c_int64, C.c_int32, vp, C.c_int64,
                                        vp, vp]
    L.mdeggs_last_stats.argtypes = [vp, C.POINTER(_abi.CStats)]
    L.mdeggs_set_option.argtypes = [vp, C.c_char_p, i64]
    L.mdeggs_measure_fp64_peak.argtypes = [vp, C.POINTER(C.c_double)]
    L.mdeggs_measure_l2_peak.argtypes = [vp, i64, C.POINTER(C.c_double)]
    L.mdeggs_trace_count.argtypes = [vp]
    L.mdeggs_trace_get.argtypes = [vp, i, C.c_char_p, i, C.POINTER(C.c_int), C.POINTER(C.c_float)]
    L.mdeggs_last_error.argtypes = [vp]
    L.mdeggs_last_error.restype = C.c_char_p
    L.mdeggs_strerror.argtypes = [i]
    L.mdeggs_strerror.restype = C.c_char_p
    L.mdeggs_abi_version.restype = i
    L.mdeggs_dev_pt.argtypes = [vp, vp, vp, i64, vp]
    L.mdeggs_dev_pf_upper.argtypes = [vp, vp, vp, vp, i64, vp]
    L.mdeggs_dev_p_two_sided_ref.argtypes = [vp, vp, vp, i64, vp]
    L.mdeggs_dev_p_two_sided_tab.argtypes = [vp, vp, C.c_double, C.c_int32, i64, vp]
    for nm in ABI_SYMBOLS:
        if getattr(L, nm).restype is C.c_int or getattr(L, nm).restype is None:
            getattr(L, nm).restype = i
    if L.mdeggs_abi_version() != _abi.ABI_VERSION:
        raise EngineError("libmdeggs.so ABI version mismatch; rebuild")
    _LIB = L
    return L


class Engine:
    """Owns one `mdeggs_ctx` (device memory, streams, NCCL communicators). Not re-entrant.

    Engine(devices=[0])                     single process, one or several GPUs of this host
    Engine.for_rank(device, rank, world, unique_id)   one process per GPU (torchrun)
    """

    def __init__(self, devices: Sequence[int] = (0,), _handle: C.c_void_p | None = None):
        self._lib = load_library()
        self._ctx = C.c_void_p()
        if _handle is not None:
            self._ctx = _handle
            return
        arr = (C.c_int * len(devices))(*devices)
        rc = self._lib.mdeggs_create(C.byref(self._ctx), arr, len(devices))
        if rc != 0:
            raise EngineError(f"mdeggs_create failed: {self._lib.mdeggs_strerror(rc).decode()}")

    @classmethod
    def for_rank(cls, device: int, rank: int, world: int, unique_id: bytes | None) -> "Engine":
        lib = load_library()
        h = C.c_void_p()
        buf = C.create_string_buffer(unique_id, 128) if unique_id is not None else None
        rc = lib.mdeggs_create_rank(C.byref(h), device, rank, world, buf)
        if rc != 0:
            raise EngineError(f"mdeggs_create_rank failed: {lib.mdeggs_strerror(rc).decode()}")
        return cls(_handle=h)

    @staticmethod
    def nccl_unique_id() -> bytes:
        lib = load_library()
        buf = C.create_string_buffer(128)
        rc = lib.mdeggs_nccl_unique_id(buf)
        if rc != 0:
            raise EngineError(f"mdeggs_nccl_unique_id failed: {lib.mdeggs_strerror(rc).decode()}")
        return buf.raw

    def close(self) -> None:
        if self._ctx:
            self._lib.mdeggs_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int, what: str) -> None:
        if rc != 0:
            msg = self._lib.mdeggs_last_error(self._ctx)
            cls = OmicError if rc in _LAYER_CODES and what.startswith(("mdeggs_run", "mdeggs_submit", "mdeggs_wait")) \
                else EngineError
            raise cls(f"{what}: {self._lib.mdeggs_strerror(rc).decode()}: {msg.decode() if msg else ''}")

    # -- the hot path ---------------------------------------------------------------------------
    def run_omic(self, problem: _abi.Problem, want: Sequence[str] = _abi.DEFAULT_WANT) -> dict[str, np.ndarray]:
        """Host buffers in, host buffers out (numpy). One call = one omic layer."""
        cprob = problem.to_c()
        cres, arrays = _abi.alloc_result(problem.dims, tuple(want))
        self._check(self._lib.mdeggs_run_omic(self._ctx, C.byref(cprob), C.byref(cres)), "mdeggs_run_omic")
        return arrays

    def run_omic_raw(self, cprob: _abi.CProblem, cres: _abi.CResult) -> None:
        """Pre-marshalled structs (device or pinned-host pointers); used by bench.py."""
        self._check(self._lib.mdeggs_run_omic(self._ctx, C.byref(cprob), C.byref(cres)), "mdeggs_run_omic")

    def submit_omic_raw(self, cprob: _abi.CProblem, cres: _abi.CResult) -> None:
        """First half of run_omic_raw: enqueue and return; the buffers stay borrowed until wait()."""
        self._check(self._lib.mdeggs_submit_omic(self._ctx, C.byref(cprob), C.byref(cres)), "mdeggs_submit_omic")

    def wait(self) -> None:
        self._check(self._lib.mdeggs_wait(self._ctx), "mdeggs_wait")

    def set_option(self, name: str, value: int) -> None:
        """"graph": 1/0 CUDA-graph replay of repeated identical calls; "fit_path": -1 auto, 0 streaming, 1 dense."""
        self._check(self._lib.mdeggs_set_option(self._ctx, name.encode(), int(value)), "mdeggs_set_option")

    def fp64_peak_tflops(self) -> float:
        """Measured FP64 FMA peak of the GPU (DFMA-chain microbenchmark)."""
        v = C.c_double()
        self._check(self._lib.mdeggs_measure_fp64_peak(self._ctx, C.byref(v)), "mdeggs_measure_fp64_peak")
        return v.value

    def l2_peak_gbs(self, nbytes: int = 64 << 20) -> float:
        """Measured L2 -> SM read bandwidth over an L2-resident buffer of `nbytes` (GB/s)."""
        v = C.c_double()
        self._check(self._lib.mdeggs_measure_l2_peak(self._ctx, int(nbytes), C.byref(v)), "mdeggs_measure_l2_peak")
        return v.value

    def trace(self) -> list[tuple[str, int, float]]:
        """(kernel, stream, end_us) of every launch of the last call (option "trace" must be on)."""
        out = []
        buf = C.create_string_buffer(64)
        sid, t = C.c_int(), C.c_float()
        for k in range(self._lib.mdeggs_trace_count(self._ctx)):
            self._check(self._lib.mdeggs_trace_get(self._ctx, k, buf, 64, C.byref(sid), C.byref(t)), "mdeggs_trace_get")
            out.append((buf.value.decode(), sid.value, t.value))
        return out

    def stats(self) -> dict:
        st = _abi.CStats()
        self._check(self._lib.mdeggs_last_stats(self._ctx, C.byref(st)), "mdeggs_last_stats")
        return st.asdict()

    def pair_features(self, x: np.ndarray, a: np.ndarray, b: np.ndarray, ratio: bool = True) -> np.ndarray:
        """predict.multiDEGGs_filter feature construction (feature_selection_for_ML.R:423-431)."""
        x = np.asfortranarray(x, dtype=np.float64)
        a = np.ascontiguousarray(a, dtype=np.int32)
        b = np.ascontiguousarray(b, dtype=np.int32)
        n = x.shape[0]
        out = np.empty((n, a.shape[0]), dtype=np.float64, order="F")
        self._check(self._lib.mdeggs_pair_features(self._ctx, x.ctypes.data, max(n, 1), n, a.ctypes.data,
                                                   b.ctypes.data, a.shape[0], int(ratio), _abi.MEM_HOST,
                                                   out.ctypes.data), "mdeggs_pair_features")
        return out

    def pair_moments(self, a_rows: np.ndarray, b_rows: np.ndarray, K: int) -> np.ndarray:
        """[n_pairs, K, 6] = n_k, mean_A, mean_B, Sxx_A, Syy_B, Sxy per pair and category, from what the last run_omic of
        this engine left on the device (the plot_regressions refit, plotting_functions.R:125-162). Rows are 1-based."""
        a = np.ascontiguousarray(a_rows, dtype=np.int32)
        b = np.ascontiguousarray(b_rows, dtype=np.int32)
        out = np.empty((a.shape[0], K, 6), dtype=np.float64)
        self._check(self._lib.mdeggs_pair_moments(self._ctx, a.ctypes.data, b.ctypes.data, a.shape[0], out.ctypes.data),
                    "mdeggs_pair_moments")
        return out

    # -- device math twins (tests) --------------------------------------------------------------
    def _vec(self, fn, *cols) -> np.ndarray:
        cols = [np.ascontiguousarray(np.broadcast_to(np.asarray(c, dtype=np.float64), np.broadcast(*cols).shape)).ravel()
                for c in cols]
        out = np.empty_like(cols[0])
        self._check(fn(self._ctx, *[c.ctypes.data for c in cols], cols[0].shape[0], out.ctypes.data), "dev math")
        return out

    def dev_pt(self, x, df) -> np.ndarray:
        return self._vec(self._lib.mdeggs_dev_pt, x, df)

    def dev_pf_upper(self, f, df1, df2) -> np.ndarray:
        return self._vec(self._lib.mdeggs_dev_pf_upper, f, df1, df2)

    def dev_p_two_sided_ref(self, t, df) -> np.ndarray:
        return self._vec(self._lib.mdeggs_dev_p_two_sided_ref, t, df)

    def dev_p_two_sided_tab(self, t, df: float, use_spline: bool = True) -> np.ndarray:
        """2 * (1 - pt(|t|, df)) the way a whole run evaluates it (one df): per-run tables, optionally the spline."""
        x = np.ascontiguousarray(t, dtype=np.float64)
        out = np.empty_like(x)
        self._check(self._lib.mdeggs_dev_p_two_sided_tab(self._ctx, x.ctypes.data, float(df), int(bool(use_spline)),
                                                         x.size, out.ctypes.data), "mdeggs_dev_p_two_sided_tab")
        return out


class EnginePool:
    """One engine (context: streams, scratch, staging) per omic layer of a get_diffNetworks call, all on one GPU:
    mdeggs_run_omics submits every layer before waiting for any, so the layers overlap on the device
    (the per-omic loop of core_functions.R:214-234 as one launch set)."""

    def __init__(self, device: int = 0):
        self.device = device
        self._engines: list[Engine] = []

    def engines(self, n: int) -> list[Engine]:
        while len(self._engines) < n:
            self._engines.append(Engine([self.device]))
        return self._engines[:n]

    def close(self) -> None:
        for e in self._engines:
            e.close()
        self._engines = []

    def run_omics(self, problems: Sequence[_abi.Problem], wants: Sequence[Sequence[str]]):
        """-> list of (result arrays | EngineError), one per layer; a failing layer does not stop the others."""
        n = len(problems)
        if n == 0:
            return []
        engs = self.engines(n)
        lib = engs[0]._lib
        cprobs = (_abi.CProblem * n)()
        cress = (_abi.CResult * n)()
        keep = []
        for i, (pr, want) in enumerate(zip(problems, wants)):
            cp = pr.to_c()
            cr, arrays = _abi.alloc_result(pr.dims, tuple(want))
            cprobs[i] = cp
            cress[i] = cr
            keep.append((cp, cr, arrays))
        ctxs = (C.c_void_p * n)(*[e._ctx for e in engs])
        rcs = (C.c_int * n)()
        lib.mdeggs_run_omics(ctxs, n, cprobs, cress, rcs)
        out = []
        for i in range(n):
            if rcs[i] == 0:
                out.append(keep[i][2])
            else:
                msg = lib.mdeggs_last_error(engs[i]._ctx)
                cls = OmicError if rcs[i] in _LAYER_CODES else EngineError
                out.append(cls(f"mdeggs_run_omics[{i}]: {lib.mdeggs_strerror(rcs[i]).decode()}: "
                               f"{msg.decode() if msg else ''}"))
        return out


_DEFAULT: Engine | None = None
_DEFAULT_POOL: EnginePool | None = None


def default_pool() -> EnginePool:
    """Lazily created process-wide pool on the first GPU of MDEGGS_GPUS (default 0)."""
    global _DEFAULT_POOL
    if _DEFAULT_POOL is None:
        env = os.environ.get("MDEGGS_GPUS")
        _DEFAULT_POOL = EnginePool(int(env.split(",")[0]) if env else 0)
    return _DEFAULT_POOL



def default_engine() -> Engine:
    """Lazily created process-wide engine (never before a fork: created on first use)."""
    global _DEFAULT
    if _DEFAULT is None:
        env = os.environ.get("MDEGGS_GPUS")
        devices = [int(t) for t in env.split(",")] if env else [0]
        _DEFAULT = Engine(devices)
    return _DEFAULT

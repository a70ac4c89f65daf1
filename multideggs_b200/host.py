"""Host-side logic that stays on the CPU in the reference too: names, factors, edge matching, the
`deggs` layout and host-side p-value adjustment for methods the device does not implement.

Everything here is string / index bookkeeping (no regression arithmetic).  It mirrors, step by step,
  tidy_metadata                          R/core_functions.R:564-686
  get_diffNetworks_singleOmic (prep)     R/core_functions.R:280-338
  calc_pvalues_network (layout)          R/core_functions.R:999-1043
  best-percentile selection              R/core_functions.R:471-495
and produces the `mdeggs_problem` handed to the C ABI (include/mdeggs.h).  The compute is done by
`engine.Engine.run_omic` (CUDA); this module never computes a regression and has no CPU fallback.
"""
from __future__ import annotations

import functools

import gzip
import os
import warnings
from dataclasses import dataclass
from typing import Any, Callable, Iterable, Sequence

import numpy as np
import pandas as pd

from . import _abi

NO_LINKS = "No specific links for this category."
P_ADJUST_METHODS = ("holm", "hochberg", "hommel", "bonferroni", "BH", "BY", "fdr", "none")

_message: Callable[[str], None] = lambda s: print(s)  # R's message(); tests may patch


def message(s: str) -> None:
    _message(s)


# ----------------------------------------------------------------------------------------------
# containers
# ----------------------------------------------------------------------------------------------
@dataclass
class Assay:
    """A numeric matrix with dimnames: genes (rows) x samples (columns), like an R matrix."""

    values: np.ndarray
    genes: np.ndarray
    samples: np.ndarray

    @staticmethod
    def coerce(x: Any) -> "Assay":
        if isinstance(x, Assay):
            return x
        if isinstance(x, pd.DataFrame):
            v = x.to_numpy(dtype=np.float64)
            return Assay(np.asfortranarray(v), np.asarray(x.index, dtype=object).astype(str),
                         np.asarray(x.columns, dtype=object).astype(str))
        raise TypeError("assayData must be a pandas.DataFrame or multideggs_b200.Assay (genes x samples)")

    def to_frame(self) -> pd.DataFrame:
        return pd.DataFrame(self.values, index=self.genes, columns=self.samples)


class Factor:
    """Named factor: `codes` are 1-based integer codes into `levels` (0 = NA), `names` the sample IDs."""

    def __init__(self, codes: np.ndarray, levels: Sequence[str], names: np.ndarray):
        self.codes = np.asarray(codes, dtype=np.int32)
        self.levels = [str(l) for l in levels]
        self.names = np.asarray(names, dtype=object).astype(str)

    def __len__(self) -> int:
        return self.codes.shape[0]

    @property
    def values(self) -> np.ndarray:
        lv = np.array([None] + self.levels, dtype=object)
        return lv[self.codes]

    def subset(self, keep: np.ndarray) -> "Factor":
        return Factor(self.codes[keep], self.levels, self.names[keep])

    def droplevels(self) -> "Factor":
        used = [i for i in range(1, len(self.levels) + 1) if np.any(self.codes == i)]
        remap = np.zeros(len(self.levels) + 1, dtype=np.int32)
        for new, old in enumerate(used, 1):
            remap[old] = new
        return Factor(remap[self.codes], [self.levels[i - 1] for i in used], self.names)

    def table(self) -> dict[str, int]:
        return {l: int(np.sum(self.codes == i)) for i, l in enumerate(self.levels, 1)}

    def to_series(self) -> pd.Series:
        return pd.Series(pd.Categorical(self.values, categories=self.levels), index=self.names)


def _as_factor(values: Iterable, names: np.ndarray) -> Factor:
    if isinstance(values, pd.Series) and isinstance(values.dtype, pd.CategoricalDtype):
        cat = values.cat
        return Factor(cat.codes.to_numpy().astype(np.int32) + 1, [str(c) for c in cat.categories], names)
    if isinstance(values, pd.Categorical):
        return Factor(np.asarray(values.codes, dtype=np.int32) + 1, [str(c) for c in values.categories], names)
    arr = np.asarray(values, dtype=object)
    isna = np.array([v is None or (isinstance(v, float) and np.isnan(v)) for v in arr])
    present = arr[~isna]
    numeric = all(isinstance(v, (int, float, np.integer, np.floating)) and not isinstance(v, bool) for v in present)
    uniq = sorted(set(present.tolist())) if numeric else sorted({str(v) for v in present})

    def lab(v):
        if numeric:
            return str(int(v)) if float(v).is_integer() else repr(float(v))
        return str(v)

    levels = [lab(v) for v in uniq]
    pos = {l: i + 1 for i, l in enumerate(levels)}
    codes = np.array([0 if na else pos[lab(v)] for v, na in zip(arr, isna)], dtype=np.int32)
    return Factor(codes, levels, names)


# ----------------------------------------------------------------------------------------------
# tidy_metadata (core_functions.R:564-686)
# ----------------------------------------------------------------------------------------------
def tidy_metadata(category_subset=None, metadata=None, category_variable=None, id_variable=None,
                  verbose: bool = False) -> Factor:
    if id_variable is not None:
        raise NotImplementedError("mixed models (id_variable) are outside the GPU hot path")
    if isinstance(metadata, Factor):
        fac_src, names = metadata.to_series(), metadata.names
    elif isinstance(metadata, pd.Series):
        if metadata.index is None:
            raise ValueError("Named vector required: metadata vector must have names.")
        fac_src, names = metadata, np.asarray(metadata.index, dtype=object).astype(str)
    elif isinstance(metadata, pd.DataFrame):
        if category_variable is None:
            raise ValueError("category_variable must be specified when metadata is\n           a matrix or data.frame.")
        if category_variable not in metadata.columns:
            raise ValueError(f"category_variable '{category_variable}' not found in metadata.")
        fac_src = metadata[category_variable]
        names = np.asarray(metadata.index, dtype=object).astype(str)
    elif isinstance(metadata, dict):
        names = np.array(list(metadata.keys()), dtype=object).astype(str)
        fac_src = pd.Series(list(metadata.values()), index=names)
    else:
        raise ValueError("metadata must be a named vector/factor, matrix, or data.frame.")
    was_factor = isinstance(fac_src.dtype, pd.CategoricalDtype)
    vec = fac_src

    if category_subset is not None:
        present = {str(v) for v in np.asarray(vec, dtype=object) if v is not None and v == v}
        if not all(str(c) in present for c in category_subset):
            raise ValueError("category_subset values must be contained in metadata")
        keep = np.array([str(v) in {str(c) for c in category_subset} for v in np.asarray(vec, dtype=object)])
        if not keep.any():
            raise ValueError("No samples remain after filtering for specified category_subset.")
        vec = vec[keep]
        names = names[keep]

    fac = _as_factor(vec, names)
    if not was_factor and verbose:
        message("category values were converted to factor")

    tbl = fac.table()
    small = [l for l, c in tbl.items() if c < 5]
    if small:
        for g in small:
            message(f"The {g} category did not contain enough samples (less than five \n"
                    "              observations). This category will be removed.")
        big = {l for l, c in tbl.items() if c >= 5}
        keep = np.array([v in big for v in fac.values])
        fac = fac.subset(keep).droplevels()
        if len(fac.levels) < 2:
            raise ValueError("At least two categories with at least 5 observations are required.")

    n_na = int(np.sum(fac.codes == 0))
    if n_na > 0:
        fac = fac.subset(fac.codes != 0).droplevels()
        if verbose:
            message(f"category values contained NAs. {n_na} samples were removed.")
    return fac


# ----------------------------------------------------------------------------------------------
# network
# ----------------------------------------------------------------------------------------------
_OMIC_NETWORK: pd.DataFrame | None = None


def omic_network() -> pd.DataFrame:
    """The built-in biological network (reference internal data `omic_network`, R/sysdata.rda;
    53,035 directed KEGG signalling edges, built by data-raw/omic_network.R)."""
    global _OMIC_NETWORK
    if _OMIC_NETWORK is None:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "omic_network.tsv.gz")
        with gzip.open(path, "rt") as fh:
            _OMIC_NETWORK = pd.read_csv(fh, sep="\t", dtype=str, keep_default_na=False)
    return _OMIC_NETWORK


@dataclass
class IndexedNetwork:
    """A network already expressed as 1-based row indices into one assay (skips string matching;
    used for very large user networks such as the all-pairs benchmark)."""

    from_idx: np.ndarray
    to_idx: np.ndarray


# ----------------------------------------------------------------------------------------------
# per-omic preparation (core_functions.R:280-338)
# ----------------------------------------------------------------------------------------------
@dataclass
class OmicPrep:
    name: str
    problem: _abi.Problem
    genes: np.ndarray          # rownames of the (sample-aligned) assay
    samples: np.ndarray
    categories: list[str]
    padj_method: str


def _padj_code(padj_method: str, n_edges: int = 0) -> int:
    if padj_method == "none":
        return _abi.PADJ_NONE
    if padj_method == "bonferroni":
        return _abi.PADJ_BONFERRONI
    if padj_method in ("BH", "fdr"):
        return _abi.PADJ_BH
    if padj_method == "holm":
        return _abi.PADJ_HOLM
    if padj_method == "hochberg":
        return _abi.PADJ_HOCHBERG
    if padj_method == "BY":
        return _abi.PADJ_BY
    if padj_method == "q.value":
        return _abi.PADJ_QVALUE   # pi0 smoother + q-values on the device (SURVEY section 8 f, N1)
    if padj_method == "hommel":  # on the device up to MDEGGS_HOMMEL_MAX_E edges, else raw p + masks -> p_adjust() here
        return _abi.PADJ_HOMMEL if n_edges <= _abi.HOMMEL_MAX_E else _abi.PADJ_HOST
    raise ValueError(f"padj_method must be one of {P_ADJUST_METHODS + ('q.value',)}")


_NET_MATCH_CACHE: dict = {}


def _match_network(network, genes: np.ndarray, assayDataName: str):
    """edges <- network rows whose two ends are rownames (match(): first occurrence), duplicates removed keeping the
    first (core_functions.R:319-331), as 1-based row indices.  Repeated calls with the same network object and the
    same rownames (nestedcv folds) reuse the result."""
    net = omic_network() if network is None else network
    # user networks are identified by CONTENT (a data.frame mutated in place, or a new one at a recycled id(), must
    # not hit the entry of the old network)
    ident = "builtin" if network is None else pd.util.hash_pandas_object(net[["from", "to"]], index=False).values.tobytes()
    key = (ident, len(net), hash(genes.tobytes()), genes.shape[0])
    hit = _NET_MATCH_CACHE.get(key)
    if hit is not None:
        return hit
    first: dict[str, int] = {}
    for i, g in enumerate(genes.tolist()):
        first.setdefault(g, i + 1)  # match(): first occurrence
    nf = pd.Series(net["from"].to_numpy(dtype=object)).map(first)
    nt = pd.Series(net["to"].to_numpy(dtype=object)).map(first)
    ok = (nf.notna() & nt.notna()).to_numpy()
    fr = nf.to_numpy()[ok].astype(np.int64)
    to = nt.to_numpy()[ok].astype(np.int64)
    if fr.size == 0:
        raise ValueError(f"Rownames of {assayDataName} don't match any link of the biological\n"
                         "    network. Check if names need conversion to gene symbols or entrez IDs. \n"
                         "    Alternatively, provide more data.")
    # remove duplicate edges, keep first (core_functions.R:329-331)
    k2 = fr * (len(genes) + 1) + to
    _, first_pos = np.unique(k2, return_index=True)
    sel = np.sort(first_pos)
    out = (fr[sel].astype(np.int32), to[sel].astype(np.int32))
    if len(_NET_MATCH_CACHE) >= 8:
        _NET_MATCH_CACHE.pop(next(iter(_NET_MATCH_CACHE)))
    _NET_MATCH_CACHE[key] = out
    return out


def prepare_omic(assayData, assayDataName: str, metadata: Factor, regression_method: str, network,
                 percentile_vector, padj_method: str, tested_coef: int = 3,
                 rlm_cov_mode: int = _abi.RLM_COV_XTWX) -> OmicPrep:
    if not isinstance(assayData, (Assay, pd.DataFrame)):
        raise ValueError(f"{assayDataName} is not a matrix or data.frame")
    assay = Assay.coerce(assayData)
    meta_pos = {nm: i for i, nm in enumerate(metadata.names)}
    in_meta = np.array([s in meta_pos for s in assay.samples], dtype=bool)
    if not in_meta.any():
        raise ValueError(f"Sample IDs in {assayDataName} don't match the IDs in metadata \n"
                         "          (the metadata names/rownames must match the sample IDs used as column \n"
                         f"          names of {assayDataName}.")
    values, samples = assay.values, assay.samples
    if not in_meta.all():
        values, samples = values[:, in_meta], samples[in_meta]
    # remove columns with 0 or 1 non-NA values (core_functions.R:292)
    if np.isnan(values).any():
        keep = np.isnan(values).sum(axis=0) < (values.shape[0] - 1)
        if not keep.all():
            values, samples = values[:, keep], samples[keep]
    have = set(samples.tolist())
    missing = [nm for nm in metadata.names if nm not in have]
    if missing:
        message(f"The following samples IDs are missing in {assayDataName}:\n" + ", ".join(missing))
    codes = metadata.codes[[meta_pos[s] for s in samples]]
    if len(np.unique(codes)) == 1:
        raise ValueError(f"All sample IDs in {assayDataName} belong to one category. \n"
                         "           No differential analysis is possible.")
    genes = assay.genes
    if isinstance(network, IndexedNetwork):
        fr = np.asarray(network.from_idx, dtype=np.int32)
        to = np.asarray(network.to_idx, dtype=np.int32)
    else:
        fr, to = _match_network(network, genes, assayDataName)
    K = len(metadata.levels)
    if K > _abi.MAX_K:
        raise ValueError(f"more than {_abi.MAX_K} categories are not supported by the GPU engine")
    method = {"lm": _abi.METHOD_LM, "rlm": _abi.METHOD_RLM}[regression_method]
    prob = _abi.Problem(assay=values if values.flags.f_contiguous or values.flags.c_contiguous
                        else np.asfortranarray(values),
                        from_idx=fr, to_idx=to, group=codes.astype(np.int32), K=K,
                        pct=np.asarray(percentile_vector, dtype=np.float64), method=method,
                        padj=_padj_code(padj_method, len(fr)), tested_coef=tested_coef, rlm_cov_mode=rlm_cov_mode)
    return OmicPrep(assayDataName, prob, genes, samples, list(metadata.levels), padj_method)


def want_fields(prep: OmicPrep, extra: Sequence[str] = ()) -> tuple[str, ...]:
    base = ["cutoff", "p_edge", "status", "tot_edges", "sig_count", "fail_count", "best", "cat_best"]
    code = prep.problem.padj
    if code in _abi.DEVICE_PADJ:
        base.append("padj_best")
    if code == _abi.PADJ_HOST:
        base.append("cat")
    return tuple(dict.fromkeys(base + list(extra)))


# ----------------------------------------------------------------------------------------------
# host-side multiple-testing adjustment (methods the device leaves to the host; SURVEY §8 f N1/N2)
# ----------------------------------------------------------------------------------------------
def p_adjust(p: np.ndarray, method: str) -> np.ndarray:
    """stats::p.adjust (used at core_functions.R:1029). NaN = NA."""
    p = np.asarray(p, dtype=np.float64)
    out = p.copy()
    nna = ~np.isnan(p)
    n = int(nna.sum())
    if n <= 1 or method == "none":
        return out
    pv = p[nna]
    if method == "bonferroni":
        res = np.minimum(1.0, n * pv)
    elif method in ("BH", "fdr", "BY", "hochberg"):
        o = np.argsort(-pv, kind="stable")
        ro = np.argsort(o, kind="stable")
        i = np.arange(n, 0, -1, dtype=np.float64)
        if method in ("BH", "fdr"):
            v = (n / i) * pv[o]
        elif method == "BY":
            q = np.sum(1.0 / np.arange(1, n + 1))
            v = (q * n / i) * pv[o]
        else:
            v = (n + 1.0 - i) * pv[o]
        res = np.minimum(1.0, np.minimum.accumulate(v))[ro]
    elif method == "holm":
        o = np.argsort(pv, kind="stable")
        ro = np.argsort(o, kind="stable")
        i = np.arange(1, n + 1, dtype=np.float64)
        res = np.minimum(1.0, np.maximum.accumulate((n + 1.0 - i) * pv[o]))[ro]
    elif method == "hommel":
        o = np.argsort(pv, kind="stable")
        ps = pv[o]
        ro = np.argsort(o, kind="stable")
        i = np.arange(1, n + 1, dtype=np.float64)
        q = pa = np.full(n, np.min(n * ps / i))
        for m in range(n - 1, 1, -1):
            i1 = np.arange(0, n - m + 1)
            i2 = np.arange(n - m + 1, n)
            q1 = np.min(m * ps[i2] / np.arange(2, m + 1))
            q = q.copy()
            q[i1] = np.minimum(m * ps[i1], q1)
            q[i2] = q[n - m]
            pa = np.maximum(pa, q)
        res = np.maximum(pa, ps)[ro]
    else:
        raise ValueError(f"unknown p.adjust method {method}")
    out[nna] = res
    return out


@functools.lru_cache(maxsize=8)
def _smoother_last_row(x_bytes: bytes, df: float) -> np.ndarray:
    """Last row of the smoother matrix of the natural cubic smoothing spline on knots `x` whose trace is `df`
    (Reinsch form).  For fixed knots, unit weights and a target df the penalty parameter does not depend on the
    data, so the fitted value at the last knot is one dot product with this row."""
    x = np.frombuffer(x_bytes, dtype=np.float64)
    n = x.shape[0]
    h = np.diff(x)
    Q = np.zeros((n, n - 2))
    R = np.zeros((n - 2, n - 2))
    for j in range(n - 2):
        Q[j, j] = 1.0 / h[j]
        Q[j + 1, j] = -1.0 / h[j] - 1.0 / h[j + 1]
        Q[j + 2, j] = 1.0 / h[j + 1]
        R[j, j] = (h[j] + h[j + 1]) / 3.0
        if j + 1 < n - 2:
            R[j, j + 1] = R[j + 1, j] = h[j + 1] / 6.0
    Kmat = Q @ np.linalg.solve(R, Q.T)
    ev, U = np.linalg.eigh(Kmat)
    ev = np.clip(ev, 0.0, None)
    lo, hi = 1e-12, 1e12
    for _ in range(200):
        mid = np.sqrt(lo * hi)
        if np.sum(1.0 / (1.0 + mid * ev)) > df:
            lo = mid
        else:
            hi = mid
    lam = np.sqrt(lo * hi)
    return U[-1] @ (U / (1.0 + lam * ev)).T


def _smooth_spline_df3_last(x: np.ndarray, y: np.ndarray, df: float = 3.0) -> float:
    """Fitted value at the last knot of the natural cubic smoothing spline whose smoother matrix has
    trace `df`.  Stand-in for stats::smooth.spline(lambda, pi0, df = 3) inside qvalue::pi0est — numerically
    UNPINNED against R (R matches df only to its own optimiser tolerance)."""
    row = _smoother_last_row(np.ascontiguousarray(x, dtype=np.float64).tobytes(), float(df))
    return float(row @ np.asarray(y, dtype=np.float64))


def qvalue_or_fallback(p: np.ndarray) -> np.ndarray:
    """`padj_method = "q.value"` branch of calc_pvalues_network (core_functions.R:1012-1025):
    qvalue::qvalue(p)$qvalues; on error retry with pi0 = 1 when there is more than one row, else NA."""
    p = np.asarray(p, dtype=np.float64)
    m_all = p.shape[0]
    nna = ~np.isnan(p)
    pv = p[nna]
    m = pv.shape[0]

    def q_from_pi0(pi0: float) -> np.ndarray:
        o = np.argsort(-pv, kind="stable")
        ro = np.argsort(o, kind="stable")
        i = np.arange(m, 0, -1, dtype=np.float64)
        q = pi0 * np.minimum(1.0, np.minimum.accumulate(pv[o] * m / i))[ro]
        out = np.full(m_all, np.nan)
        out[nna] = q
        return out

    try:
        if m == 0 or pv.min() < 0 or pv.max() > 1:
            raise ValueError("p-values not in valid range [0, 1]")
        lam = _abi.r_seq(0.05, 0.95, 0.05)
        if pv.max() < lam.max():
            raise ValueError("maximum p-value is smaller than lambda range")
        if m < 2:
            raise ValueError("lfdr density estimate needs at least two p-values")
        ps_sorted = np.sort(pv)
        pi0_l = (m - np.searchsorted(ps_sorted, lam, side="left")) / m / (1.0 - lam)
        pi0 = min(_smooth_spline_df3_last(lam, pi0_l, 3.0), 1.0)
        if pi0 <= 0:
            raise ValueError("estimated pi0 <= 0")
        return q_from_pi0(pi0)
    except ValueError:
        if m_all > 1:
            return q_from_pi0(1.0)
        return np.full(m_all, np.nan)


def host_adjust(p: np.ndarray, padj_method: str) -> np.ndarray:
    if padj_method == "q.value":
        return qvalue_or_fallback(p)
    return p_adjust(p, padj_method)


# ----------------------------------------------------------------------------------------------
# result assembly (core_functions.R:471-495, 999-1043)
# ----------------------------------------------------------------------------------------------
def assemble_networks(prep: OmicPrep, res: dict[str, np.ndarray], cores: int = 1, verbose: bool = False) -> dict:
    """Builds what get_diffNetworks_singleOmic returns: {category: DataFrame | NO_LINKS, ..., 'sig_pvalues_count'}."""
    prob = prep.problem
    fr, to = np.asarray(prob.from_idx), np.asarray(prob.to_idx)
    P = len(prob.pct)
    K = prob.K
    p_edge = res["p_edge"]
    tot = res["tot_edges"]
    code = prob.padj
    if code == _abi.PADJ_HOST:
        cat_all = res["cat"]
        sig = np.full(P, -1, dtype=np.int64)
        padj_rows: dict[int, np.ndarray] = {}
        for pi in range(P):
            if tot[pi] <= 0:
                continue
            row = np.full(p_edge.shape[0], np.nan)
            cnt = 0
            for k in range(1, K + 1):
                m = cat_all[pi] == k
                if m.any():
                    row[m] = host_adjust(p_edge[m], prep.padj_method)
                    cnt += int(np.sum(row[m] < 0.05))
            sig[pi] = cnt
            padj_rows[pi] = row
        valid = [pi for pi in range(P) if tot[pi] > 0]
        get_cat = lambda pi: cat_all[pi]
        get_padj = lambda pi: padj_rows[pi]
    else:
        sig = np.asarray(res["sig_count"]).copy()
        valid = [pi for pi in range(P) if tot[pi] > 0]
        best_dev = int(res["best"][0])
        get_cat = lambda pi: res["cat_best"] if pi == best_dev - 1 else None
        get_padj = lambda pi: res.get("padj_best") if pi == best_dev - 1 else None

    # error semantics of the reference (quirk q7): a failing pair (solve() / .lm.fit / rlm() error) aborts the
    # omic when cores == 1 (mclapply degenerates to lapply) and only its percentile otherwise
    # (core_functions.R:475-483).  The engine reports such pairs in fail_count[p] and skips those percentiles.
    fail = np.asarray(res["fail_count"])
    if np.any(fail[valid] > 0) if valid else False:
        if cores == 1:
            raise ArithmeticError("a tested gene pair has a singular design or non-finite values "
                                  "(R: 'system is computationally singular' / 'NA/NaN/Inf in x')")
        valid = [pi for pi in valid if fail[pi] == 0]
    if not valid:
        raise ValueError("No significant differences detected across categories.")
    best = max(valid, key=lambda pi: (sig[pi], -pi))  # which.max: first maximum
    if verbose:
        message("Percolation analysis: genes whose expression is below the "
                f"{_r_num(prob.pct[best] * 100)}th percentile are removed from networks.")
    cat = get_cat(best)
    padj = get_padj(best)
    out: dict[str, Any] = {}
    for k, level in enumerate(prep.categories, 1):
        m = np.flatnonzero(cat == k)
        if m.size == 0:
            out[level] = NO_LINKS
            continue
        f_names = prep.genes[fr[m] - 1]
        t_names = prep.genes[to[m] - 1]
        cols = {"from": f_names, "to": t_names, "p.value": p_edge[m]}
        if prep.padj_method != "none":
            cols["p.adj"] = padj[m]
        # (row names as one object array: handing pandas a Python list costs a per-element type inference, 3-7x the
        # time of everything else in this function on config 4)
        names = np.empty(m.size, dtype=object)
        names[:] = [f"{a}-{b}" for a, b in zip(f_names.tolist(), t_names.tolist())]
        out[level] = pd.DataFrame(cols, index=pd.Index(names, dtype=object))
    out["sig_pvalues_count"] = float(sig[best])
    return out


def _r_num(x: float) -> str:
    s = f"{x:.15g}"
    return s

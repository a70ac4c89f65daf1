"""Seeded synthetic workloads of BASELINE.json's configs (SURVEY §8 d): numpy `default_rng(seed)` (PCG64),
value model of the reference's data-raw/synthetic_data.R (base N(5, 1), planted pairs whose correlation
differs between categories, clamp to [0, 23]; synthetic_data.R:65-76, 210-211, 254-255).

  cfg3: 20,000 genes x 1,000 samples, 3 categories, the built-in 53,035-edge omic_network  (seed 20261019)
  cfg4: 5,000 features x 24 samples, 2 categories, multiDEGGs_filter defaults                (seed 20261020)
  cfg5: 4,000 features x 2,000 samples, 2 categories, all i<j pairs (7,998,000 edges)        (seed 20261021)
Every generator takes size overrides so tests can run scaled-down versions of the same recipe.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import _abi, host


@dataclass
class Workload:
    name: str
    assay: host.Assay          # genes x samples, Fortran order (R column-major)
    group: np.ndarray          # int32 codes 1..K per sample
    levels: list[str]
    from_idx: np.ndarray       # int32, 1-based rows of `assay`
    to_idx: np.ndarray
    percentiles: np.ndarray
    method: int = _abi.METHOD_LM
    padj: int = _abi.PADJ_BH

    def problem(self) -> _abi.Problem:
        return _abi.Problem(assay=self.assay.values, from_idx=self.from_idx, to_idx=self.to_idx, group=self.group,
                            K=len(self.levels), pct=self.percentiles, method=self.method, padj=self.padj)

    def factor(self) -> host.Factor:
        return host.Factor(self.group, self.levels, self.assay.samples)


def _plant(values: np.ndarray, group: np.ndarray, fr: np.ndarray, to: np.ndarray, frac: float, rho: float, sd: float,
           rng: np.random.Generator, K: int) -> None:
    E = fr.shape[0]
    m = int(round(frac * E))
    if m == 0:
        return
    chosen = rng.choice(E, size=m, replace=False)
    cats = rng.integers(1, K + 1, size=m)
    for e, c in zip(chosen.tolist(), cats.tolist()):
        cols = np.flatnonzero(group == c)
        z1 = rng.standard_normal(cols.shape[0])
        z2 = rng.standard_normal(cols.shape[0])
        a, b = fr[e] - 1, to[e] - 1
        if a == b:
            continue
        values[a, cols] = 5.0 + sd * z1
        values[b, cols] = 5.0 + sd * (rho * z1 + np.sqrt(1.0 - rho * rho) * z2)


def _groups(n: int, sizes: list[int], rng: np.random.Generator) -> np.ndarray:
    g = np.concatenate([np.full(s, k + 1, dtype=np.int32) for k, s in enumerate(sizes)])
    assert g.shape[0] == n
    rng.shuffle(g)
    return g


def cfg3(n_genes: int = 20000, n_samples: int = 1000, seed: int = 20261019, planted: float = 0.10,
         max_edges: int | None = None) -> Workload:
    """20,000 x 1,000 RNA-seq-like matrix, 3 categories, built-in network (aov interaction F + BH)."""
    rng = np.random.default_rng(seed)
    net = host.omic_network()
    nf, nt = net["from"].to_numpy(dtype=object), net["to"].to_numpy(dtype=object)
    if max_edges is not None:
        nf, nt = nf[:max_edges], nt[:max_edges]
    order: dict[str, int] = {}
    for a in np.concatenate([nf, nt]).tolist():  # first appearance in c(from, to)
        order.setdefault(a, len(order))
    node_names = list(order.keys())
    n_nodes = len(node_names)
    n_genes = max(n_genes, n_nodes)
    genes = np.array(node_names + [f"FILL{i + 1:05d}" for i in range(n_genes - n_nodes)], dtype=object).astype(str)
    samples = np.array([f"S{i + 1:04d}" for i in range(n_samples)], dtype=object).astype(str)
    base = n_samples // 3
    sizes = [n_samples - 2 * base, base, base]
    group = _groups(n_samples, sizes, rng)
    values = np.asfortranarray(5.0 + rng.standard_normal((n_genes, n_samples)))
    fr = np.array([order[a] + 1 for a in nf.tolist()], dtype=np.int32)
    to = np.array([order[a] + 1 for a in nt.tolist()], dtype=np.int32)
    key = fr.astype(np.int64) * (n_genes + 1) + to
    _, first = np.unique(key, return_index=True)
    sel = np.sort(first)
    fr, to = fr[sel], to[sel]
    _plant(values, group, fr, to, planted, 0.8, 2.0, rng, 3)
    np.clip(values, 0.0, 23.0, out=values)
    return Workload("cfg3", host.Assay(values, genes, samples), group, ["A", "B", "C"], fr, to,
                    _abi.r_seq(0.35, 0.98, 0.05), _abi.METHOD_LM, _abi.PADJ_BH)


def cfg4(n_features: int = 5000, n_samples: int = 24, seed: int = 20261020, planted: float = 0.10) -> Workload:
    """Small-sample percentile path used inside nestedcv folds (multiDEGGs_filter: 15 percentiles, q.value)."""
    rng = np.random.default_rng(seed)
    net = host.omic_network()
    nf, nt = net["from"].to_numpy(dtype=object), net["to"].to_numpy(dtype=object)
    order: dict[str, int] = {}
    for a in np.concatenate([nf, nt]).tolist():
        order.setdefault(a, len(order))
    names = list(order.keys())[:n_features]
    keep = {g: i for i, g in enumerate(names)}
    genes = np.array(names, dtype=object).astype(str)
    samples = np.array([f"S{i + 1:04d}" for i in range(n_samples)], dtype=object).astype(str)
    group = _groups(n_samples, [n_samples - n_samples // 2, n_samples // 2], rng)
    values = np.asfortranarray(5.0 + rng.standard_normal((len(names), n_samples)))
    ok = np.array([(a in keep) and (b in keep) for a, b in zip(nf.tolist(), nt.tolist())])
    fr = np.array([keep[a] + 1 for a in nf[ok].tolist()], dtype=np.int32)
    to = np.array([keep[b] + 1 for b in nt[ok].tolist()], dtype=np.int32)
    key = fr.astype(np.int64) * (len(names) + 1) + to
    _, first = np.unique(key, return_index=True)
    sel = np.sort(first)
    fr, to = fr[sel], to[sel]
    _plant(values, group, fr, to, planted, 0.8, 1.0, rng, 2)
    np.clip(values, 0.0, 23.0, out=values)
    return Workload("cfg4", host.Assay(values, genes, samples), group, ["1", "2"], fr, to,
                    _abi.r_seq(0.25, 0.98, 0.05), _abi.METHOD_LM, _abi.PADJ_HOST)


def cfg5(n_features: int = 4000, n_samples: int = 2000, seed: int = 20261021, planted: float = 0.01) -> Workload:
    """User-supplied all-pairs network: every i<j pair of n_features features, lm + global BH."""
    rng = np.random.default_rng(seed)
    genes = np.array([f"F{i + 1:04d}" for i in range(n_features)], dtype=object).astype(str)
    samples = np.array([f"S{i + 1:04d}" for i in range(n_samples)], dtype=object).astype(str)
    group = _groups(n_samples, [n_samples - n_samples // 2, n_samples // 2], rng)
    values = np.asfortranarray(5.0 + rng.standard_normal((n_features, n_samples)))
    iu, ju = np.triu_indices(n_features, k=1)
    fr = (iu + 1).astype(np.int32)
    to = (ju + 1).astype(np.int32)
    E = fr.shape[0]
    m = int(round(planted * E))
    if m:
        # planting pair by pair would overwrite rows many times for a dense network; plant disjoint pairs instead
        perm = rng.permutation(n_features)
        n_pairs = min(m, n_features // 2)
        cats = rng.integers(1, 3, size=n_pairs)
        for q in range(n_pairs):
            a, b = perm[2 * q], perm[2 * q + 1]
            cols = np.flatnonzero(group == cats[q])
            z1 = rng.standard_normal(cols.shape[0])
            z2 = rng.standard_normal(cols.shape[0])
            values[a, cols] = 5.0 + 2.0 * z1
            values[b, cols] = 5.0 + 2.0 * (0.8 * z1 + 0.6 * z2)
    np.clip(values, 0.0, 23.0, out=values)
    return Workload("cfg5", host.Assay(values, genes, samples), group, ["1", "2"], fr, to,
                    _abi.r_seq(0.35, 0.98, 0.05), _abi.METHOD_LM, _abi.PADJ_BH)


def random_problem(G: int, n: int, E: int, K: int, seed: int, method: int = _abi.METHOD_LM,
                   padj: int = _abi.PADJ_BH, pct: np.ndarray | None = None, min_per_group: int = 5,
                   row_major: bool = False, shift: float = 0.6) -> _abi.Problem:
    """Random edge list over random rows with category-dependent expression shifts so that the percolation
    step selects non-trivial edge sets (used by parity tests)."""
    rng = np.random.default_rng(seed)
    group = rng.integers(1, K + 1, size=n).astype(np.int32)
    for k in range(1, K + 1):  # guarantee min_per_group samples per category
        need = min_per_group - int(np.sum(group == k))
        if need > 0:
            counts = np.bincount(group, minlength=K + 1)
            donors = np.flatnonzero(group == int(np.argmax(counts)))
            group[donors[:need]] = k
    values = 5.0 + rng.standard_normal((G, n))
    values += shift * rng.standard_normal((G, K))[:, group - 1]  # per gene, per category mean shift
    n_corr = max(1, G // 4)
    for _ in range(n_corr):
        a, b = rng.integers(0, G, size=2)
        if a == b:
            continue
        c = int(rng.integers(1, K + 1))
        cols = np.flatnonzero(group == c)
        values[b, cols] = values[b, cols].mean() + 0.9 * (values[a, cols] - values[a, cols].mean()) \
            + 0.3 * rng.standard_normal(cols.shape[0])
    values = np.clip(values, 0.0, 23.0)
    pairs = set()
    fr, to = [], []
    while len(fr) < E:
        a, b = (int(v) for v in rng.integers(1, G + 1, size=2))
        if a == b or (a, b) in pairs:
            continue
        pairs.add((a, b))
        fr.append(a)
        to.append(b)
    vals = np.ascontiguousarray(values) if row_major else np.asfortranarray(values)
    return _abi.Problem(assay=vals, from_idx=np.array(fr, dtype=np.int32), to_idx=np.array(to, dtype=np.int32),
                        group=group, K=K, pct=_abi.r_seq(0.35, 0.98, 0.05) if pct is None else pct, method=method,
                        padj=padj)

"""Independent numpy / scipy / mpmath twin of the C oracle (TEST INFRASTRUCTURE ONLY).

Written separately from oracle/mdeggs_oracle.c, with different numerical routes (numpy lstsq / QR, scipy
special functions, mpmath at 50 digits, string set operations for the percolation step exactly as
R/core_functions.R:731-770 does them), so that an error of restatement in one is caught by the other.
Pure-Python loops: small cases only.
"""
from __future__ import annotations

import itertools

import mpmath as mp
import numpy as np
from scipy import special

mp.mp.dps = 50


# ---- tails ---------------------------------------------------------------------------------------------
def t_two_sided_mp(t: float, df: float) -> float:
    """P(|T_df| > t), exact to double precision via mpmath's incomplete beta."""
    t, df = mp.mpf(t), mp.mpf(df)
    if t == 0:
        return 1.0
    x = df / (df + t * t)
    return float(mp.betainc(df / 2, mp.mpf(1) / 2, 0, x, regularized=True))


def f_upper_mp(f: float, d1: float, d2: float) -> float:
    f, d1, d2 = mp.mpf(f), mp.mpf(d1), mp.mpf(d2)
    x = d2 / (d2 + d1 * f)
    return float(mp.betainc(d2 / 2, d1 / 2, 0, x, regularized=True))


def p_two_sided_ref(t: float, df: float) -> float:
    """2 * (1 - pt(|t|, df)) with R's pt() composition (core_functions.R:1072), tails from scipy."""
    val = special.betainc(df / 2.0, 0.5, df / (df + t * t))  # = P(|T| > |t|)
    u = (0.5 - val / 2.0) + 0.5
    return 2.0 * (1.0 - u)


# ---- fits ----------------------------------------------------------------------------------------------
def lm_interaction(a, b, g01):
    """fit_lm_interaction (core_functions.R:1058-1074) with numpy lstsq and an explicit inverse."""
    a, b, g = (np.asarray(v, dtype=np.float64) for v in (a, b, g01))
    X = np.column_stack([np.ones_like(a), a, g, a * g])
    coef, *_ = np.linalg.lstsq(X, b, rcond=None)
    resid = b - X @ coef
    dof = len(b) - 4
    sigma2 = float(resid @ resid) / dof
    xtx_inv = np.linalg.inv(X.T @ X)
    t = coef[3] / np.sqrt(sigma2 * xtx_inv[3, 3])
    return coef, float(t), p_two_sided_ref(float(t), dof)


def aov_interaction(a, b, grp0, K):
    """summary(aov(B ~ A * metadata))[[1]][["Pr(>F)"]][3] (core_functions.R:966-968): sequential SS by nested fits."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    grp0 = np.asarray(grp0)
    n = len(b)
    ind = [(grp0 == k).astype(np.float64) for k in range(1, K)]
    X_add = np.column_stack([np.ones(n), a] + ind)
    X_full = np.column_stack([X_add] + [a * i for i in ind])

    def rss(X):
        c, *_ = np.linalg.lstsq(X, b, rcond=None)
        r = b - X @ c
        return float(r @ r), c

    rss_add, _ = rss(X_add)
    rss_full, coef = rss(X_full)
    # degrees of freedom from the matrix ranks: with a predictor that is constant inside a category the interaction
    # loses a column (aov() drops aliased columns, it does not fail); full rank gives (K - 1, n - 2K)
    r_add, r_full = np.linalg.matrix_rank(X_add, tol=1e-9), np.linalg.matrix_rank(X_full, tol=1e-9)
    d1, d2 = r_full - r_add, n - r_full
    if d1 <= 0:
        return coef, float("nan"), float("nan")   # no A:metadata row in the table: Pr(>F)[3] is NA
    f = ((rss_add - rss_full) / d1) / (rss_full / d2)
    p = float(special.fdtrc(d1, d2, f))
    return coef, f, p


def rlm_wald(a, b, g01, tested_coef=3, method="XtWX"):
    """MASS::rlm defaults + summary.rlm + sfsmisc::f.robftest(var = tested_coef) (core_functions.R:916-925)."""
    a, b, g = (np.asarray(v, dtype=np.float64) for v in (a, b, g01))
    n = len(b)
    X = np.column_stack([np.ones(n), a, g, a * g])
    k = 1.345

    def wls(w):
        sw = np.sqrt(w)
        c, *_ = np.linalg.lstsq(X * sw[:, None], b * sw, rcond=None)
        return c, b - X @ c

    coef, resid = wls(np.ones(n))
    done, it, scale = False, 0, 0.0
    for it in range(1, 21):
        old = resid.copy()
        scale = np.median(np.abs(resid)) / 0.6745
        if scale == 0:
            done = True
            break
        w = np.minimum(1.0, k / np.abs(resid / scale))
        coef, resid = wls(w)
        conv = np.sqrt(np.sum((old - resid) ** 2) / max(1e-20, np.sum(old ** 2)))
        if conv <= 1e-4:
            done = True
            break
    p = 4
    rdf = n - p
    w = np.minimum(1.0, k / np.abs(resid / scale))
    S = np.sum((resid * w) ** 2) / rdf
    psiprime = (np.abs(resid / scale) <= k).astype(np.float64)
    mn = psiprime.mean()
    kappa = 1 + p * psiprime.var(ddof=1) / (n * mn ** 2)
    stddev = np.sqrt(S) * (kappa / mn)
    Xs = X * np.sqrt(w / w.mean())[:, None] if method == "XtWX" else X
    cov = np.linalg.inv(Xs.T @ Xs)
    v = tested_coef - 1
    f = coef[v] ** 2 / (stddev ** 2 * cov[v, v])
    return coef, float(f), float(special.fdtrc(1, rdf, f)), it


# ---- p.adjust, quantile ---------------------------------------------------------------------------------
def p_adjust_loops(p, method: str):
    """stats::p.adjust written as explicit loops over the definition (core_functions.R:1029)."""
    p = np.asarray(p, dtype=np.float64)
    out = p.copy()
    idx = [i for i in range(len(p)) if not np.isnan(p[i])]
    n = len(idx)
    if n <= 1 or method == "none":
        return out
    if method == "bonferroni":
        for i in idx:
            out[i] = min(1.0, n * p[i])
        return out
    if method in ("BH", "BY", "hochberg"):
        order = sorted(idx, key=lambda i: (-p[i], i))       # order(p, decreasing = TRUE), stable
        q = sum(1.0 / j for j in range(1, n + 1)) if method == "BY" else 1.0
        run = float("inf")
        for pos, i in enumerate(order):
            rank = float(n - pos)                               # i = n:1
            if method == "BH":
                v = (n / rank) * p[i]
            elif method == "BY":
                v = (q * n / rank) * p[i]
            else:
                v = (n + 1.0 - rank) * p[i]
            run = min(run, v)
            out[i] = min(1.0, run)
        return out
    if method == "holm":
        order = sorted(idx, key=lambda i: (p[i], i))
        run = -float("inf")
        for pos, i in enumerate(order):
            run = max(run, (n + 1.0 - (pos + 1)) * p[i])
            out[i] = min(1.0, run)
        return out
    raise ValueError(method)


def quantile7(x, probs):
    x = np.sort(np.asarray(x, dtype=np.float64).ravel())
    x = x[~np.isnan(x)]
    N = len(x)
    out = []
    for p in probs:
        idx = 1 + (N - 1) * p
        lo, hi = int(np.floor(idx)), int(np.ceil(idx))
        q = x[lo - 1]
        if idx > lo and x[hi - 1] != q:
            h = idx - lo
            q = (1 - h) * q + h * x[hi - 1]
        out.append(q)
    return np.array(out)


# ---- whole path, string set logic as in R ----------------------------------------------------------------
def run_omic(prob) -> dict:
    """get_diffNetworks_singleOmic after edge selection (core_functions.R:333-495), as the R code does it:
    named medians, `paste(from, to)` strings, intersect/unique for common links."""
    A = np.asarray(prob.assay, dtype=np.float64)
    G, n = A.shape
    fr, to = np.asarray(prob.from_idx) - 1, np.asarray(prob.to_idx) - 1
    grp = np.asarray(prob.group)
    K, E = prob.K, len(fr)
    nodes = sorted(set(fr.tolist()) | set(to.tolist()))
    names = {g: f"g{g}" for g in range(G)}
    sub = A[nodes]
    med = {k: {names[g]: (np.nanmedian(A[g, grp == k]) if np.any(grp == k) and not np.all(np.isnan(A[g, grp == k]))
                          else np.nan) for g in nodes} for k in range(1, K + 1)}
    median_out = np.full((K, G), np.nan)
    for k in range(1, K + 1):
        for g in nodes:
            median_out[k - 1, g] = med[k][names[g]]
    # per-edge p-values (independent of the percentile)
    p_edge = np.full(E, np.nan)
    g01 = (grp - 1).astype(np.float64)
    for e in range(E):
        a, b = A[fr[e]], A[to[e]]
        if fr[e] == to[e]:
            p_edge[e] = 1.0
            continue
        if K == 2 and prob.method == 0:
            p_edge[e] = lm_interaction(a, b, g01)[2]
        elif K == 2:
            p_edge[e] = rlm_wald(a, b, g01, prob.tested_coef, "XtWX" if prob.rlm_cov_mode == 0 else "XtX")[2]
        else:
            p_edge[e] = aov_interaction(a, b, grp - 1, K)[2]
    edge_str = [f"{names[fr[e]]} {names[to[e]]}" for e in range(E)]
    cutoffs = quantile7(sub, prob.pct)
    P = len(prob.pct)
    cat = np.zeros((P, E), dtype=np.int8)
    padj = np.full((P, E), np.nan)
    tot = np.zeros(P, dtype=np.int64)
    sig = np.full(P, -1, dtype=np.int64)
    method = {0: "none", 1: "bonferroni", 2: "BH", 3: "none"}[prob.padj]
    for pi, cut in enumerate(cutoffs):
        nets = {}
        for k in range(1, K + 1):
            active = {nm for nm, v in med[k].items() if v > cut}
            nets[k] = [s for e, s in enumerate(edge_str) if names[fr[e]] in active and names[to[e]] in active]
        common = set()
        for k1, k2 in itertools.combinations(range(1, K + 1), 2):
            common |= set(nets[k1]) & set(nets[k2])
        count = 0
        for k in range(1, K + 1):
            keep = [s for s in nets[k] if s not in common]
            rows = [edge_str.index(s) for s in keep]
            tot[pi] += len(rows)
            if not rows:
                continue
            cat[pi, rows] = k
            pv = p_edge[rows]
            adj = p_adjust_loops(pv, method) if method != "none" else pv
            if prob.padj in (1, 2):
                padj[pi, rows] = adj
            count += int(np.sum(adj < 0.05))
        if tot[pi] > 0:
            sig[pi] = count
    best = 0
    valid = [pi for pi in range(P) if tot[pi] > 0]
    if valid and prob.padj != 3:
        best = max(valid, key=lambda pi: (sig[pi], -pi)) + 1
    if prob.padj == 3:
        sig[:] = -1
    return {"cutoff": cutoffs, "median": median_out, "p_edge": p_edge, "cat": cat, "padj": padj, "tot_edges": tot,
            "sig_count": sig, "best": best}

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
    # an adjusted value is at most (segment size) x p (bonferroni n p, BH (n/i) p, holm (n+1-i) p; BY adds the
    # harmonic factor), so the 2.2e-16 grid of the reference's p is amplified by the segment size at most
    n_seg = int(np.max(ref["tot_edges"])) if "tot_edges" in ref and np.size(ref["tot_edges"]) else int(problem.dims[2])
    amp = max(n_seg, 1) * (np.log(max(n_seg, 1)) + 1.0 if problem.padj == _abi.PADJ_BY else 1.0)
    # stat: t = (slope_2 - slope_1) / se cancels when |t| << 1 (cfg4 has a t of -2.8e-6 whose two routes, QR in the oracle
    # and the closed form in the engine, differ by 1e-14 absolute); 1e-12 absolute = 1e-9 relative at |t| = 1e-3
    tol = {"stat": (rtol, 1e-12), "coef": (rtol, 1e-12), "p_edge": (rtol, p_floor), "padj": (rtol, p_floor * amp),
           "padj_best": (rtol, p_floor * amp)}
    for k, (rt, at) in tol.items():
        if k in gpu and k in ref:
            ok = rel_close(gpu[k], ref[k], rt, at)
            assert ok.all(), (f"{what}{k}: {int((~ok).sum())} of {ok.size} outside tolerance; first "
                              f"{np.ravel(gpu[k])[np.flatnonzero(~np.ravel(ok))[:3]]} vs "
                              f"{np.ravel(ref[k])[np.flatnonzero(~np.ravel(ok))[:3]]}")


# Numbers printed by real R that the reference itself holds (vignette screenshots of the bundled data, lm + bonferroni,
# vignettes/Differential_Network_Analysis.Rmd:75-83, 146, 157, 213-221): the only numeric anchors to R for this path.
#   vignettes/multiDEGGs_1.png       Non_responder TNF -> TNFRSF1A: p adj 0.000e+00 (Proteomics), 1.330e-03 (RNAseq)
#   vignettes/multiDEGGs_2.png       RNAseq TGFB3 -> TGFBR1: p value 5.329e-15  (the 2.2e-16 grid of 2*(1-pt), core:1072)
#   vignettes/plot_regressions.png   RNAseq AKT2 ~ MTOR: Padj = 0.0017  (format(x, digits = 2), plotting_functions.R:227)
# The tables print formatC(x, format = "e", digits = 3) (plotting_functions.R:367-371).
R_PINNED = (
    ("RNAseq", "Non_responder", "TNF-TNFRSF1A", "p.adj", "e3", "1.330e-03"),
    ("Proteomics", "Non_responder", "TNF-TNFRSF1A", "p.adj", "e3", "0.000e+00"),
    ("RNAseq", "Non_responder", "TGFB3-TGFBR1", "p.value", "e3", "5.329e-15"),
    ("RNAseq", "Non_responder", "AKT2-MTOR", "p.adj", "g2", "0.0017"),
)


def r_format(x: float, how: str) -> str:
    """formatC(x, format = "e", digits = 3) / format(x, digits = 2) for the magnitudes pinned above."""
    return f"{x:.3e}" if how == "e3" else f"{x:.2g}"


def check_r_pinned(networks: dict) -> None:
    """networks[omic][category] = data.frame-like with rownames from-to (the `deggs$diffNetworks` layout)."""
    for omic, cat, row, col, how, printed in R_PINNED:
        v = float(networks[omic][cat].loc[row, col])
        assert r_format(v, how) == printed, (omic, row, col, v, printed)
    assert float(networks["Proteomics"]["Non_responder"].loc["TNF-TNFRSF1A", "p.adj"]) == 0.0   # exact zero

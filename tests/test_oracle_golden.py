"""CPU tests: the oracle against (i) every known answer of the reference's own tests, (ii) the survey-derived
candidate goldens, (iii) an independent numpy/scipy/mpmath twin."""
import numpy as np
import pandas as pd
import pytest

import multideggs_b200 as md
from multideggs_b200 import _abi, host, synth
from helpers import check_r_pinned, oracle_single_omic
import oracle_np


@pytest.fixture(scope="module")
def factor(bundled):
    return host.tidy_metadata(metadata=bundled["metadata"], category_variable="response")


def test_reference_known_answers_lm_bonferroni(bundled, factor, oracle):
    """tests/testthat/test-core_functions.R:2-51 and 95-127"""
    out = {k: oracle_single_omic(bundled[k], k, factor, "lm", None, md.DEFAULT_PERCENTILES, "bonferroni")
           for k in ("RNAseq", "Proteomics", "Olink")}
    assert out["RNAseq"]["sig_pvalues_count"] == 7
    assert out["RNAseq"]["Responder"].shape[0] == 3
    assert out["RNAseq"]["Non_responder"].shape[0] == 4
    assert out["Proteomics"]["Responder"] == md.NO_LINKS
    assert out["Olink"]["Non_responder"].shape[1] == 4
    assert out["Olink"]["Responder"] == md.NO_LINKS
    nr = out["RNAseq"]["Non_responder"]
    assert "TNF" in set(nr["from"]) and "TNFRSF1A" in set(nr["to"])
    assert list(nr.index) == ["TNF-TNFRSF1A", "IL1B-IL1R2", "TGFB3-TGFBR1", "AKT2-MTOR"]     # network order
    assert list(out["RNAseq"]["Responder"].index) == ["FANCD2-FAN1", "FASLG-FAS", "MAP2K2-MAPK3"]
    # survey probe: Proteomics 8/0 with 5 significant, Olink 1/0 with 0 significant
    assert out["Proteomics"]["Non_responder"].shape[0] == 8 and out["Proteomics"]["sig_pvalues_count"] == 5
    assert out["Olink"]["Non_responder"].shape[0] == 1 and out["Olink"]["sig_pvalues_count"] == 0


def test_real_R_numbers_held_by_the_reference(bundled, factor, oracle):
    """lm + bonferroni is pinned to real R output at 2-4 significant digits, incl. the exact 0 and the 2.2e-16 grid."""
    out = {k: oracle_single_omic(bundled[k], k, factor, "lm", None, md.DEFAULT_PERCENTILES, "bonferroni")
           for k in ("RNAseq", "Proteomics")}
    check_r_pinned(out)
    # vignettes/Feature_Selection.Rmd:133-136: the pairs multiDEGGs_filter keeps on the full bundled data (q.value path)
    y = pd.Series(bundled["metadata"]["response"].cat.codes.to_numpy() + 1.0, index=bundled["RNAseq"].columns)
    q = oracle_single_omic(bundled["RNAseq"], "assayData1", host.tidy_metadata(metadata=y), "lm", None,
                           md.FILTER_PERCENTILES, "q.value")
    kept = {i.replace("-", ":") for f in q.values() if isinstance(f, pd.DataFrame)
            for i in f.index[(f["p.adj"] < 0.05).to_numpy()]}
    assert kept == {"TNF:TNFRSF1A", "AKT2:MTOR", "IL1B:IL1R2", "FASLG:FAS", "TGFB3:TGFBR1", "MAP2K2:MAPK3", "FANCD2:FAN1"}


def test_reference_known_answers_rlm_BH(bundled, factor, oracle):
    """tests/testthat/test-core_functions.R:55-91"""
    rna = oracle_single_omic(bundled["RNAseq"], "RNAseq", factor, "rlm", None, md.DEFAULT_PERCENTILES, "BH")
    olk = oracle_single_omic(bundled["Olink"], "Olink", factor, "rlm", None, md.DEFAULT_PERCENTILES, "BH")
    deggs = md.Deggs(diffNetworks={"RNAseq": rna, "Olink": olk},
                     assayData={"RNAseq": bundled["RNAseq"], "Olink": bundled["Olink"]}, metadata=factor,
                     regression_method="rlm", mixedModel=False, id_variable=None, lmer_ctrl=None,
                     category_subset=None, padj_method="BH")
    res = md.get_multiOmics_diffNetworks(deggs)
    assert len(res) == 2 and res["Non_responder"].shape[1] == 5
    assert list(res["Responder"].columns) == ["from", "to", "p.value", "p.adj", "layer"]
    assert res["Responder"].shape[0] == 3
    assert "AKT2" in set(res["Non_responder"]["from"]) and "MTOR" in set(res["Non_responder"]["to"])
    assert "FANCD2" in set(res["Responder"]["from"]) and "FAN1" in set(res["Responder"]["to"])
    # all four (tested coefficient x covariance) variants keep every RNAseq edge significant (survey §4)
    for coef in (3, 4):
        for cov in (_abi.RLM_COV_XTWX, _abi.RLM_COV_XTX):
            r = oracle_single_omic(bundled["RNAseq"], "RNAseq", factor, "rlm", None, md.DEFAULT_PERCENTILES, "BH",
                                   tested_coef=coef, rlm_cov_mode=cov)
            assert r["sig_pvalues_count"] == 7


def test_survey_candidate_goldens(bundled, factor, oracle):
    """beta4 / t / p_ref of the RNAseq edges (numpy Householder restatement, not R)."""
    x = bundled["RNAseq"]
    g01 = (factor.codes[[list(factor.names).index(s) for s in x.columns]] - 1).astype(np.float64)
    gold = {
        ("FANCD2", "FAN1"): (1.302296224268076, 12.21588864074, 0.0),
        ("TNF", "TNFRSF1A"): (0.6934148610492126, 3.722572919314, 3.325023365317037e-04),
        ("IL1B", "IL1R2"): (0.6813539284726465, 3.581216328857, 5.389194511264961e-04),
        ("FASLG", "FAS"): (0.9370486297350311, 5.314149168563, 6.947839152893209e-07),
        ("TGFB3", "TGFBR1"): (1.176162521803456, 9.282036526726, 5.329070518200751e-15),
        ("MAP2K2", "MAPK3"): (1.306747495135146, 8.184698024226, 1.153299677980613e-12),
        ("AKT2", "MTOR"): (0.6208898017638079, 3.644554120184, 4.347082799509572e-04),
        ("TNF", "FAS"): (-0.2025562440093338, -1.051878468603, 2.954958539142014e-01),
    }
    for (a, b), (beta4, t, p) in gold.items():
        st, coef, tt, pp = oracle.fit_lm(x.loc[a].to_numpy(), x.loc[b].to_numpy(), g01)
        assert st == 0
        assert abs(coef[3] - beta4) <= 1e-12 * abs(beta4)
        assert abs(tt - t) <= 2e-12 * abs(t)
        assert abs(pp - p) <= 1e-9 * p + 4.5e-16
    cut = oracle.quantile7(x.to_numpy(), md.DEFAULT_PERCENTILES)
    expect = [3.936, 4.395, 4.797, 5.129, 5.556, 5.961, 6.319, 6.755, 7.167, 7.602, 8.119, 8.803, 9.748]
    assert np.allclose(cut, expect, atol=5e-4)
    prep = host.prepare_omic(x, "RNAseq", factor, "lm", None, md.DEFAULT_PERCENTILES, "bonferroni")
    res = oracle.run_omic(prep.problem)
    assert list(res["tot_edges"]) == [7, 7, 7, 7, 7, 6, 6, 5, 3, 2, 0, 0, 0]
    assert int(res["best"][0]) == 1          # which.max: FIRST maximum (percentile 0.35)


def test_multiDEGGs_filter_flow_with_oracle(bundled, factor, oracle):
    """tests/testthat/test-ML_functions.R:12-17, 66-69: q.value path (bundled data => BH fallback), row order."""
    y = pd.Series(bundled["metadata"]["response"].cat.codes.to_numpy() + 1.0, index=bundled["RNAseq"].columns)
    fac = host.tidy_metadata(metadata=y)
    assert fac.levels == ["1", "2"]
    out = oracle_single_omic(bundled["RNAseq"], "assayData1", fac, "lm", None, md.FILTER_PERCENTILES, "q.value")
    frames = [v for v in out.values() if isinstance(v, pd.DataFrame)]
    pairs = pd.concat([f[(f["p.adj"] < 0.05).to_numpy()] for f in frames])
    assert pairs.shape == (7, 4)
    assert list(pairs.index[:5]) == ["TNF-TNFRSF1A", "IL1B-IL1R2", "TGFB3-TGFBR1", "AKT2-MTOR", "FANCD2-FAN1"]


# ---- oracle vs the independent twin -------------------------------------------------------------------
@pytest.mark.parametrize("df", [1.0, 2.0, 5.0, 18.0, 96.0, 996.0, 1996.0, 4e5])
def test_pt_against_mpmath(oracle, df):
    ts = [0.0, 1e-6, 0.3, 1.0, 2.5, 5.5, 9.9, 12.2, 28.8, 40.0, -3.3]
    for t in ts:
        true_two_sided = oracle_np.t_two_sided_mp(abs(t), df)
        got = oracle.lib().orc_p_two_sided_true(t, df)
        rtol = 2e-12 if df <= 2000 else 1e-10   # long continued fractions at df ~ 1e5 accumulate ~1e-11
        assert abs(got - true_two_sided) <= rtol * true_two_sided + 1e-305, (t, df)
        ref = oracle.lib().orc_p_two_sided_ref(t, df)
        assert abs(ref - true_two_sided) <= 1e-9 * true_two_sided + 4.5e-16, (t, df)


@pytest.mark.parametrize("d1,d2", [(1.0, 96.0), (2.0, 994.0), (3.0, 20.0), (7.0, 1992.0), (1.0, 1e5), (2.0, 6.0)])
def test_pf_against_mpmath(oracle, d1, d2):
    for f in [1e-8, 0.1, 0.9, 1.0, 3.7, 25.0, 300.0, 5e3, 1e5]:
        true = oracle_np.f_upper_mp(f, d1, d2)
        got = oracle.lib().orc_pf_upper(f, d1, d2)
        assert abs(got - true) <= 5e-12 * true + 1e-305, (f, d1, d2)


def test_two_sided_ref_is_quantised_like_R():
    """2*(1-pt) is exactly 0 once the true p drops below ~1.5e-16 and sits on a 2.2e-16 grid."""
    import pyoracle
    L = pyoracle.lib()
    assert L.orc_p_two_sided_ref(12.21588864074, 96.0) == 0.0
    p = L.orc_p_two_sided_ref(9.282036526726, 96.0)
    assert p == 5.329070518200751e-15 and (p / 2.220446049250313e-16) == round(p / 2.220446049250313e-16)


@pytest.mark.parametrize("seed,n,K", [(1, 12, 2), (2, 100, 2), (3, 57, 3), (4, 240, 4), (5, 33, 5)])
def test_fits_against_numpy_twin(oracle, seed, n, K):
    rng = np.random.default_rng(seed)
    grp = rng.integers(0, K, size=n)
    grp[:K * 3] = np.repeat(np.arange(K), 3)
    a = 5 + rng.standard_normal(n)
    b = 5 + 0.5 * a * (grp == 1) + rng.standard_normal(n)
    if K == 2:
        st, coef, t, p = oracle.fit_lm(a, b, grp.astype(float))
        c2, t2, p2_true = oracle_np.lm_interaction(a, b, grp.astype(float))
        assert st == 0 and np.allclose(coef, c2, rtol=1e-10, atol=1e-12)
        assert abs(t - t2) <= 1e-10 * abs(t2)
        assert abs(p - p2_true) <= 1e-9 * p2_true + 4.5e-16
        st, coef, f, p, it = oracle.fit_rlm(a, b, grp.astype(float), 3, _abi.RLM_COV_XTWX)
        c3, f3, p3, it3 = oracle_np.rlm_wald(a, b, grp.astype(float), 3, "XtWX")
        assert st == 0 and it == it3 and np.allclose(coef, c3, rtol=1e-9, atol=1e-12)
        assert abs(f - f3) <= 1e-8 * abs(f3) and abs(p - p3) <= 1e-7 * p3
        st, coef, f, p, it = oracle.fit_rlm(a, b, grp.astype(float), 4, _abi.RLM_COV_XTX)
        c3, f3, p3, it3 = oracle_np.rlm_wald(a, b, grp.astype(float), 4, "XtX")
        assert abs(f - f3) <= 1e-8 * abs(f3) and abs(p - p3) <= 1e-7 * p3
    else:
        st, coef, f, p = oracle.fit_aov(a, b, grp, K)
        c2, f2, p2 = oracle_np.aov_interaction(a, b, grp, K)
        assert st == 0 and np.allclose(coef, c2, rtol=1e-9, atol=1e-11)
        assert abs(f - f2) <= 1e-9 * abs(f2) and abs(p - p2) <= 1e-9 * p2


def test_aov_with_aliased_interaction_columns(oracle):
    """K >= 3 goes through aov() (core_functions.R:963-968), which does not fail on a predictor that is constant inside a
    category (an RNA-seq gene that is all zero in one group): lm() pivots the aliased A:metadata column out, the term
    keeps the remaining degrees of freedom, and only when no interaction column is left is Pr(>F)[3] NA."""
    rng = np.random.default_rng(8)
    n, K = 90, 3
    grp = np.repeat(np.arange(K), n // K)
    rng.shuffle(grp)
    b = 5 + rng.standard_normal(n)
    for const_in in ([1], [0], [2], [0, 2], [0, 1, 2]):
        a = 5 + rng.standard_normal(n)
        for k in const_in:
            a[grp == k] = 0.0 if k != 1 else 3.25
        bb = b + 0.8 * a * (grp == 1)
        st, coef, f, p = oracle.fit_aov(a, bb, grp, K)
        _, f2, p2 = oracle_np.aov_interaction(a, bb, grp, K)
        assert st == _abi.EDGE_ALIASED, const_in
        if len(const_in) >= 2:       # at most one category with a slope of its own: no interaction row at all
            assert np.isnan(p) and np.isnan(f) and np.isnan(p2) and np.all(np.isnan(coef))
            continue
        assert abs(f - f2) <= 1e-9 * abs(f2) and abs(p - p2) <= 1e-9 * p2, const_in
        # closed form the engine uses: the constant category contributes its Syy to RSS_full and loses one df each way
        sxx = np.array([np.sum((a[grp == k] - a[grp == k].mean()) ** 2) for k in range(K)])
        syy = np.array([np.sum((bb[grp == k] - bb[grp == k].mean()) ** 2) for k in range(K)])
        sxy = np.array([np.sum((a[grp == k] - a[grp == k].mean()) * (bb[grp == k] - bb[grp == k].mean())) for k in range(K)])
        ok = sxx > 0
        rss_full = np.sum(syy[ok] - sxy[ok] ** 2 / sxx[ok]) + np.sum(syy[~ok])
        rss_add = syy.sum() - sxy.sum() ** 2 / sxx.sum()
        f3 = ((rss_add - rss_full) / (K - 2)) / (rss_full / (n - 2 * K + 1))
        assert abs(f - f3) <= 1e-9 * abs(f3)
        # NA coefficient: the aliased category's own interaction, or (reference category constant) the last one
        na_at = K + (const_in[0] if const_in[0] > 0 else K - 1)
        assert np.isnan(coef[na_at]) and np.sum(np.isnan(coef)) == 1
    # a full-rank fit is unchanged
    a = 5 + rng.standard_normal(n)
    st, coef, f, p = oracle.fit_aov(a, b, grp, K)
    assert st == _abi.EDGE_OK and not np.any(np.isnan(coef))


@pytest.mark.parametrize("K,method", [(3, _abi.METHOD_LM), (4, _abi.METHOD_LM), (2, _abi.METHOD_RLM), (2, _abi.METHOD_LM)])
def test_pairwise_na_removal_like_model_frame(oracle, K, method):
    """rlm(B ~ A * metadata) and aov(B ~ A * metadata) (core_functions.R:916-917, 966-967) build a model frame:
    na.omit drops, per pair, the samples where either gene is NA, and the test uses the complete cases and THEIR degrees
    of freedom; fit_lm_interaction (core_functions.R:1065) hands NA to .lm.fit, which throws.  The oracle's complete-case
    route is checked against the independent twin run on the complete cases."""
    prob = synth.random_problem(40, 90, 160, K, 77 + K, method, _abi.PADJ_BH, min_per_group=12)
    rng = np.random.default_rng(5)
    v = prob.assay
    na_rows = rng.choice(40, size=12, replace=False)
    for g in na_rows:
        v[g, rng.choice(90, size=int(rng.integers(1, 9)), replace=False)] = np.nan
    inf_row = int(na_rows[0])
    v[inf_row, 0] = np.inf                                   # an Inf is not dropped: lm.fit / lm.wfit throw
    one = int(na_rows[1])                                    # a category that keeps a single complete case
    cols = np.flatnonzero(prob.group == 1)
    v[one, cols[1:]] = np.nan
    res = oracle.run_omic(prob, ("p_edge", "stat", "status", "iters", "fail_count", "tot_edges", "cat"))
    fr, to = prob.from_idx - 1, prob.to_idx - 1
    touched = np.isnan(v[fr]).any(axis=1) | np.isnan(v[to]).any(axis=1)
    assert touched.sum() > 30
    st = res["status"]
    if K == 2 and method == _abi.METHOD_LM:
        assert np.all(st[touched] == _abi.EDGE_NONFINITE) and np.all(np.isnan(res["p_edge"][touched]))
        return
    has_inf = np.isinf(v[fr]).any(axis=1) | np.isinf(v[to]).any(axis=1)
    seen = set()
    for e in np.flatnonzero(touched):
        a, b = v[fr[e]], v[to[e]]
        keep = ~(np.isnan(a) | np.isnan(b))
        a2, b2, g2 = a[keep], b[keep], prob.group[keep] - 1
        if np.isinf(a2).any() or np.isinf(b2).any():
            assert st[e] == _abi.EDGE_NONFINITE and np.isnan(res["p_edge"][e])
            seen.add("inf")
            continue
        cnt = np.bincount(g2, minlength=K)
        if K == 2 and (cnt < 2).any():
            assert st[e] == _abi.EDGE_SINGULAR, (e, cnt)     # rlm: "'x' is singular"
            seen.add("singular")
            continue
        if K >= 3:
            # lm() drops the unused levels of the complete cases (drop.unused.levels = TRUE): renumber them
            lev = np.flatnonzero(cnt > 0)
            g3 = np.searchsorted(lev, g2)
            _, f2, p2 = oracle_np.aov_interaction(a2, b2, g3, len(lev))
            assert st[e] in (_abi.EDGE_NA_OMIT, _abi.EDGE_ALIASED), (e, st[e])
            seen.add(int(st[e]))
            assert abs(res["stat"][e] - f2) <= 1e-9 * abs(f2) and abs(res["p_edge"][e] - p2) <= 1e-9 * p2, e
        else:
            _, f2, p2, it2 = oracle_np.rlm_wald(a2, b2, g2.astype(float), 3, "XtWX")
            assert st[e] in (_abi.EDGE_NA_OMIT, _abi.EDGE_NOT_CONVERGED)
            seen.add(int(st[e]))
            assert res["iters"][e] == it2
            assert abs(res["stat"][e] - f2) <= 1e-8 * abs(f2) and abs(res["p_edge"][e] - p2) <= 1e-7 * p2, e
    assert _abi.EDGE_NA_OMIT in seen and "inf" in seen
    assert has_inf.any()
    # untouched pairs keep the run's degrees of freedom and status
    assert np.all(np.isin(st[~touched], (_abi.EDGE_OK, _abi.EDGE_SELFLOOP, _abi.EDGE_NOT_CONVERGED)))
    # soft statuses never count as failures
    hard = np.isin(st, (_abi.EDGE_SINGULAR, _abi.EDGE_NONFINITE))
    P, E = len(prob.pct), len(fr)
    cat = res["cat"].reshape(P, E)
    assert np.array_equal(res["fail_count"], ((cat != 0) & hard[None, :]).sum(axis=1))


@pytest.mark.parametrize("method", ["none", "bonferroni", "BH", "holm", "hochberg", "BY"])
def test_p_adjust_three_ways(oracle, method):
    rng = np.random.default_rng(11)
    p = rng.uniform(size=200) ** 3
    p[[5, 17]] = np.nan
    p[40:44] = p[39]           # ties
    p[60] = 0.0
    code = {"none": 0, "bonferroni": 1, "BH": 2, "holm": 3, "hochberg": 4, "BY": 5}[method]
    a = oracle.p_adjust(p, code)
    b = host.p_adjust(p, method)
    c = oracle_np.p_adjust_loops(p, method)
    if method == "BY":   # q = sum(1/(1:n)): R accumulates in long double, so only ~1e-15 agreement is defined
        assert np.allclose(a, b, rtol=1e-14, equal_nan=True) and np.allclose(a, c, rtol=1e-14, equal_nan=True)
    else:
        assert np.array_equal(a, b, equal_nan=True) and np.array_equal(a, c, equal_nan=True)
    one = np.array([0.3])
    assert oracle.p_adjust(one, code)[0] == 0.3      # n <= 1: unchanged


def test_quantile_median_seq(oracle):
    rng = np.random.default_rng(5)
    x = np.round(rng.normal(5, 1, size=1001), 2)
    x[::50] = np.nan
    probs = md.DEFAULT_PERCENTILES
    assert np.array_equal(oracle.quantile7(x, probs), oracle_np.quantile7(x, probs))
    assert np.allclose(oracle.quantile7(x, probs), np.nanquantile(x, probs), rtol=1e-14)
    for m in (1, 2, 7, 8):
        assert oracle.median(x[1:1 + m]) == np.nanmedian(x[1:1 + m])
    # base::seq in double arithmetic: not the decimal literals
    s = oracle.seq(0.35, 0.98, 0.05)
    assert len(s) == 13 and s[1] == 0.39999999999999997 and s[12] == 0.9500000000000001
    assert np.array_equal(s, md.DEFAULT_PERCENTILES)
    f = oracle.seq(0.25, 0.98, 0.05)
    assert len(f) == 15 and f[7] == 0.6000000000000001 and np.array_equal(f, md.FILTER_PERCENTILES)


@pytest.mark.parametrize("K,method,padj", [(2, "lm", _abi.PADJ_BH), (3, "lm", _abi.PADJ_BONFERRONI),
                                           (2, "rlm", _abi.PADJ_BH), (2, "lm", _abi.PADJ_NONE)])
def test_whole_path_against_numpy_twin(oracle, K, method, padj):
    """The C pipeline vs a line-by-line numpy/string restatement of core_functions.R:702-832 on a small case."""
    m = _abi.METHOD_LM if method == "lm" else _abi.METHOD_RLM
    prob = synth.random_problem(24, 40, 60, K, 99 + K, m, padj)
    res = oracle.run_omic(prob)
    twin = oracle_np.run_omic(prob)
    assert np.array_equal(res["tot_edges"], twin["tot_edges"])
    assert np.array_equal(res["cat"], twin["cat"])
    assert np.array_equal(res["sig_count"], twin["sig_count"])
    assert int(res["best"][0]) == twin["best"]
    assert np.array_equal(res["cutoff"], twin["cutoff"])
    assert np.array_equal(res["median"], twin["median"], equal_nan=True)
    ok = np.isclose(res["p_edge"], twin["p_edge"], rtol=1e-7, atol=4.5e-16)
    assert ok.all()
    if padj in (_abi.PADJ_BH, _abi.PADJ_BONFERRONI):
        assert np.allclose(res["padj"], twin["padj"], rtol=1e-7, atol=1e-15, equal_nan=True)


def test_reference_mode_equals_unique_mode(oracle):
    prob = synth.random_problem(40, 30, 150, 2, 5)
    a = oracle.run_omic(prob, mode=oracle.MODE_UNIQUE)
    b = oracle.run_omic(prob, mode=oracle.MODE_REFERENCE, nthreads=3)
    for k in a:
        assert np.array_equal(a[k], b[k], equal_nan=True), k

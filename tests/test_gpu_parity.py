"""-m gpu: the CUDA path (through the C ABI) against the CPU oracle on the same inputs."""
import numpy as np
import pandas as pd
import pytest

import multideggs_b200 as md
from multideggs_b200 import _abi, host, synth
from helpers import ALL_FIELDS, assert_parity, check_r_pinned, rel_close

pytestmark = pytest.mark.gpu


def _both(engine, oracle, prob, want=ALL_FIELDS):
    gpu = engine.run_omic(prob, want)
    ref = oracle.run_omic(prob, want)
    return gpu, ref


@pytest.mark.parametrize("omic", ["RNAseq", "Proteomics", "Olink"])
@pytest.mark.parametrize("method,padj", [("lm", "bonferroni"), ("lm", "BH"), ("lm", "none"), ("lm", "q.value"),
                                         ("rlm", "BH"), ("rlm", "bonferroni")])
def test_bundled_engine_vs_oracle(engine, oracle, bundled, omic, method, padj):
    fac = host.tidy_metadata(metadata=bundled["metadata"], category_variable="response")
    pct = md.FILTER_PERCENTILES if padj == "q.value" else md.DEFAULT_PERCENTILES
    prep = host.prepare_omic(bundled[omic], omic, fac, method, None, pct, padj)
    gpu, ref = _both(engine, oracle, prep.problem)
    assert_parity(gpu, ref, prep.problem, what=f"{omic}/{method}/{padj}: ")


@pytest.mark.parametrize("G,n,E,K,seed", [
    (40, 10, 60, 2, 1), (64, 22, 200, 2, 2), (100, 100, 500, 2, 3), (300, 1000, 800, 2, 4), (50, 2000, 300, 2, 5),
    (80, 37, 300, 3, 6), (120, 333, 900, 3, 7), (60, 1000, 400, 3, 8), (90, 64, 500, 4, 9), (70, 120, 300, 5, 10),
    (33, 17, 40, 2, 11), (257, 129, 1025, 3, 12), (60, 400, 200, 8, 13), (45, 90, 150, 7, 14), (40, 66, 90, 6, 15),
])
@pytest.mark.parametrize("padj", [_abi.PADJ_BH, _abi.PADJ_BONFERRONI, _abi.PADJ_HOLM, _abi.PADJ_HOCHBERG, _abi.PADJ_BY])
def test_random_lm_aov(engine, oracle, G, n, E, K, seed, padj):
    prob = synth.random_problem(G, n, E, K, seed, _abi.METHOD_LM, padj)
    gpu, ref = _both(engine, oracle, prob)
    assert_parity(gpu, ref, prob, what=f"G{G} n{n} E{E} K{K}: ")
    assert ref["tot_edges"].sum() > 0


@pytest.mark.parametrize("G,n,E,seed", [(40, 30, 80, 21), (60, 100, 200, 22), (50, 501, 150, 23), (30, 1200, 60, 24)])
@pytest.mark.parametrize("coef,cov", [(3, _abi.RLM_COV_XTWX), (4, _abi.RLM_COV_XTWX), (3, _abi.RLM_COV_XTX)])
def test_random_rlm(engine, oracle, G, n, E, seed, coef, cov):
    prob = synth.random_problem(G, n, E, 2, seed, _abi.METHOD_RLM, _abi.PADJ_BH)
    prob.tested_coef, prob.rlm_cov_mode = coef, cov
    gpu, ref = _both(engine, oracle, prob)
    assert_parity(gpu, ref, prob, rtol=1e-8, what=f"rlm G{G} n{n}: ")


def test_plot_regressions_refit_from_cached_moments(engine, bundled, oracle):
    """plot_regressions (plotting_functions.R:125-172): interaction p of the pair, the printed Padj, and per category the
    line lm(y ~ x) with its confidence band, rebuilt from the moments the engine run left on the device
    (mdeggs_pair_moments).  Checked against numpy fits on the bundled data and against the R-printed Padj = 0.0017 of
    AKT2 ~ MTOR (vignettes/plot_regressions.png)."""
    from scipy import stats as st
    deggs = md.get_diffNetworks(assayData={k: bundled[k] for k in ("RNAseq", "Proteomics", "Olink")},
                                metadata=bundled["metadata"], category_variable="response", regression_method="lm",
                                padj_method="bonferroni", verbose=False, show_progressBar=False, cores=1)
    d = md.plot_regressions_data(deggs, "RNAseq", "AKT2", "MTOR", engine=engine)
    assert d["prefix"] == "Padj" and f"{d['sig_interaction']:.2g}" == "0.0017"
    assay = bundled["RNAseq"]
    fac = host.tidy_metadata(metadata=bundled["metadata"], category_variable="response")
    grp = pd.Series(fac.values, index=fac.names)[assay.columns]
    x, y = assay.loc["AKT2"].to_numpy(), assay.loc["MTOR"].to_numpy()
    # interaction p: coef(summary(lm(y ~ x * g)))[4, 4] = 2 pt(-|t|)
    st_, coef, t, p_ref = oracle.fit_lm(x, y, (grp.to_numpy() == d["categories"][1]).astype(float))
    assert abs(d["p_interaction"] - 2 * st.t.sf(abs(t), len(x) - 4)) <= 1e-9 * d["p_interaction"]
    for cat in d["categories"]:
        m = (grp == cat).to_numpy()
        slope, icpt = np.polyfit(x[m], y[m], 1)
        L = d["lines"][cat]
        assert abs(L["slope"] - slope) <= 1e-9 * abs(slope) and abs(L["intercept"] - icpt) <= 1e-9 * abs(icpt) + 1e-12
        resid = y[m] - (icpt + slope * x[m])
        sigma = np.sqrt(resid @ resid / (m.sum() - 2))
        se = sigma * np.sqrt(1 / m.sum() + (d["new_x"] - x[m].mean()) ** 2 / np.sum((x[m] - x[m].mean()) ** 2))
        tq = st.t.ppf(0.975, m.sum() - 2)
        assert np.allclose(L["lwr"], icpt + slope * d["new_x"] - tq * se, rtol=1e-9, atol=1e-12)
        assert np.allclose(L["upr"], icpt + slope * d["new_x"] + tq * se, rtol=1e-9, atol=1e-12)
    assert len(d["new_x"]) == 100 and d["new_x"][0] < x.min() and d["new_x"][-1] > x.max()
    # raw moments of several pairs at once, a row that is no node of the run -> NaN
    prep = host.prepare_omic(assay, "RNAseq", fac, "lm", None, md.DEFAULT_PERCENTILES, "bonferroni")
    engine.run_omic(prep.problem, ("best",))
    fr, to = prep.problem.from_idx[:5], prep.problem.to_idx[:5]
    mom = engine.pair_moments(fr, to, 2)
    for j in range(5):
        for k, cat in enumerate(d["categories"]):
            m = (grp == cat).to_numpy()
            a, b = prep.problem.assay[fr[j] - 1][m], prep.problem.assay[to[j] - 1][m]
            want = [m.sum(), a.mean(), b.mean(), np.sum((a - a.mean()) ** 2), np.sum((b - b.mean()) ** 2),
                    np.sum((a - a.mean()) * (b - b.mean()))]
            assert np.allclose(mom[j, k], want, rtol=1e-10, atol=1e-12), (j, k)
    with pytest.raises(ValueError):
        md.plot_regressions_data(deggs, "RNAseq", "NOT_A_GENE", "MTOR", engine=engine)


def _na_problem(K, method, padj, seed, G=60, n=96, E=400):
    """random problem with missing values in a quarter of the rows, one row with an Inf as well, and one category of one
    row reduced to a single complete case (rlm: singular, aov: aliased / fewer levels)"""
    prob = synth.random_problem(G, n, E, K, seed, method, padj, min_per_group=14)
    rng = np.random.default_rng(seed + 1000)
    v = prob.assay
    na_rows = rng.choice(G, size=G // 4, replace=False)
    for g in na_rows:
        v[g, rng.choice(n, size=int(rng.integers(1, 10)), replace=False)] = np.nan
    v[int(na_rows[0]), 1] = np.inf
    cols = np.flatnonzero(prob.group == 1)
    v[int(na_rows[1]), cols[1:]] = np.nan
    v[int(na_rows[2]), cols] = np.nan                       # a whole category missing for this gene
    v[int(na_rows[3]), :] = np.nan                          # an all-NA row: no complete case
    return prob


@pytest.mark.parametrize("K,method", [(3, _abi.METHOD_LM), (5, _abi.METHOD_LM), (2, _abi.METHOD_RLM), (2, _abi.METHOD_LM)])
@pytest.mark.parametrize("padj", [_abi.PADJ_BH, _abi.PADJ_BONFERRONI])
def test_pairwise_na_removal(engine, oracle, K, method, padj):
    """rlm() / aov() drop, per pair, the samples where either gene is NA (model.frame, core_functions.R:916-917,
    966-967) and test on the complete cases; fit_lm_interaction (K == 2, lm) errors instead (core_functions.R:1065)."""
    prob = _na_problem(K, method, padj, 300 + K)
    gpu, ref = _both(engine, oracle, prob)
    assert_parity(gpu, ref, prob, rtol=1e-8 if method == _abi.METHOD_RLM else 1e-9, what=f"NA K{K} m{method}: ")
    st = ref["status"]
    if K == 2 and method == _abi.METHOD_LM:
        assert (st == _abi.EDGE_NONFINITE).sum() > 50 and not (st == _abi.EDGE_NA_OMIT).any()
    else:
        assert (st == _abi.EDGE_NA_OMIT).sum() > 50 and (st == _abi.EDGE_NONFINITE).any() and (st == _abi.EDGE_SINGULAR).any()
        assert np.all(np.isfinite(ref["p_edge"][st == _abi.EDGE_NA_OMIT]))


def test_pairwise_na_removal_dense_path_and_sharded_queue(oracle):
    """same through the dense tile moments (src 1 of k_finish) and with the K8 radix-sort ordering"""
    prob = _na_problem(3, _abi.METHOD_LM, _abi.PADJ_BH, 411, G=48, n=64, E=900)
    ref = oracle.run_omic(prob, ALL_FIELDS)
    eng = md.Engine([0])
    for opt, val in (("fit_path", 1), ("sort_path", 0), ("front_path", 0), ("bh_onepass", 0)):
        eng.set_option(opt, val)
        gpu = eng.run_omic(prob, ALL_FIELDS)
        assert_parity(gpu, ref, prob, what=f"NA {opt}={val}: ")
    eng.close()


def test_row_major_layout_matches_col_major(engine):
    a = synth.random_problem(50, 40, 120, 2, 31)
    b = synth.random_problem(50, 40, 120, 2, 31, row_major=True)
    ra, rb = engine.run_omic(a, ALL_FIELDS), engine.run_omic(b, ALL_FIELDS)
    for k in ra:
        assert np.array_equal(ra[k], rb[k], equal_nan=True), k


def test_edge_cases_status(engine, oracle):
    prob = synth.random_problem(30, 40, 50, 2, 41)
    v = prob.assay.copy()
    v[3, :] = 7.0                                  # constant gene -> singular when it is the predictor
    v[5, prob.group == 1] = 2.5                    # constant inside one category
    v[8, 4] = np.nan                               # NA
    v[9, 2] = np.inf
    prob.assay = np.asfortranarray(v)
    fr = np.concatenate([prob.from_idx, [4, 6, 9, 10, 12, 12]]).astype(np.int32)
    to = np.concatenate([prob.to_idx, [1, 2, 3, 11, 12, 13]]).astype(np.int32)   # includes a self-loop 12->12
    prob.from_idx, prob.to_idx = fr, to
    gpu, ref = _both(engine, oracle, prob)
    assert set(np.unique(gpu["status"])) >= {0, 1, 2, 5}
    assert_parity(gpu, ref, prob, what="edge cases: ")


@pytest.mark.parametrize("K,fit_path", [(3, 0), (4, 0), (3, 1)])
def test_aov_aliased_columns(engine, oracle, K, fit_path):
    """K >= 3: a predictor that is constant inside a category (RNA-seq gene all zero in one group) is NOT a failure:
    aov() drops the aliased A:metadata column (core_functions.R:966-968), status = ALIASED, reduced d.f., fail_count
    untouched; with a single category left that has a slope, p is NA.  Same through the dense-tile path."""
    prob = synth.random_problem(40, 90, 160, K, 4100 + K)
    v = prob.assay.copy()
    v[2, prob.group == 2] = 0.0                    # constant in a non-reference category
    v[4, prob.group == 1] = 0.0                    # constant in the reference category
    v[6, prob.group != 3] = 1.5                    # a slope of its own only in category 3 (K = 3: no interaction left)
    v[7, :] = 2.0                                  # constant everywhere
    prob.assay = np.asfortranarray(v)
    extra = [(3, 1), (3, 9), (5, 2), (5, 11), (7, 12), (8, 13), (1, 3), (2, 5)]   # aliased genes as predictor / response
    prob.from_idx = np.concatenate([prob.from_idx, [a for a, _ in extra]]).astype(np.int32)
    prob.to_idx = np.concatenate([prob.to_idx, [b for _, b in extra]]).astype(np.int32)
    engine.set_option("fit_path", fit_path)
    try:
        gpu, ref = _both(engine, oracle, prob)
    finally:
        engine.set_option("fit_path", -1)
    assert engine.stats()["fit_path"] & 0xff == fit_path
    assert (ref["status"] == _abi.EDGE_ALIASED).sum() >= 4 and not (ref["status"] == _abi.EDGE_SINGULAR).any()
    assert np.isnan(ref["p_edge"][ref["status"] == _abi.EDGE_ALIASED]).any()
    assert np.isfinite(ref["p_edge"][ref["status"] == _abi.EDGE_ALIASED]).any()
    assert ref["fail_count"].sum() == 0
    assert_parity(gpu, ref, prob, what=f"aliased aov K{K}: ")


def test_ties_and_clamped_values(engine, oracle):
    prob = synth.random_problem(60, 50, 200, 3, 51)
    v = np.round(prob.assay, 1)                    # heavy ties: medians / quantiles hit equal values
    v[v < 4.0] = 0.0                               # clamp floor as in data-raw/synthetic_data.R:254-255
    prob.assay = np.asfortranarray(v)
    gpu, ref = _both(engine, oracle, prob)
    assert_parity(gpu, ref, prob, what="ties: ")


def test_zero_sample_category_and_empty_percentiles(engine, oracle):
    prob = synth.random_problem(40, 30, 100, 2, 61)
    prob.K = 3                                     # level 3 exists but has no sample (quirk q2)
    prob.pct = np.array([0.0, 0.5, 0.999999, 1.0])
    gpu, ref = _both(engine, oracle, prob)
    assert_parity(gpu, ref, prob, what="empty category: ")


def test_device_tails_match_oracle(engine, oracle):
    rng = np.random.default_rng(7)
    t = np.concatenate([rng.standard_normal(2000) * 3, [0.0, 1e-8, 40.0, -40.0, 9.9, 10.1, 28.8]])
    for df in (1.0, 2.0, 5.0, 18.0, 96.0, 996.0, 1996.0, 4e5):
        got = engine.dev_pt(t, df)
        ref = np.array([oracle.lib().orc_pt(float(x), df) for x in t])
        assert np.all(np.abs(got - ref) <= (1e-13 if df <= 2000 else 1e-10)), df   # long CFs at df ~ 1e5
        got2 = engine.dev_p_two_sided_ref(t, df)
        ref2 = np.array([oracle.lib().orc_p_two_sided_ref(float(x), df) for x in t])
        assert np.all(np.abs(got2 - ref2) <= 1e-9 * ref2 + 4.5e-16), df
    f = np.abs(rng.standard_normal(2000)) * 10
    f = np.concatenate([f, [0.0, 1e-12, 1e3, 1e6]])
    for d1, d2 in ((1.0, 96.0), (2.0, 994.0), (3.0, 20.0), (7.0, 1992.0), (1.0, 1e5)):
        got = engine.dev_pf_upper(f, d1, d2)
        ref = np.array([oracle.lib().orc_pf_upper(float(x), d1, d2) for x in f])
        assert np.all(np.abs(got - ref) <= 1e-10 * ref + 1e-300), (d1, d2)


def test_device_tails_against_mpmath(engine):
    """The device tail routines checked DIRECTLY against 50-digit mpmath (independent of the oracle, whose tails share
    the continued-fraction algorithm): pt lower tail at negative t = half the two-sided tail; pf upper tail incl. the
    closed form for 2 numerator d.f.; the reference's 2*(1-pt(|t|)) within its own 2.2e-16 quantisation."""
    import oracle_np
    ts = np.array([1e-6, 0.3, 1.0, 1.8, 2.5, 4.0, 5.5, 9.9, 12.2, 28.8, 40.0])
    for df in (1.0, 2.0, 5.0, 18.0, 96.0, 996.0, 1996.0, 4e5):
        true2 = np.array([oracle_np.t_two_sided_mp(float(x), df) for x in ts])
        rtol = 2e-12 if df <= 2000 else 1e-10     # long continued fractions at df ~ 1e5 accumulate ~1e-11
        lower = engine.dev_pt(-ts, df)            # pt(-|t|) = P(T < -|t|): no cancellation, meaningful to 1e-300
        assert np.all(np.abs(lower - true2 / 2) <= rtol * true2 / 2 + 1e-305), df
        ref = engine.dev_p_two_sided_ref(ts, df)
        assert np.all(np.abs(ref - true2) <= 1e-9 * true2 + 4.5e-16), df
        assert np.all((ref / 2.220446049250313e-16) == np.round(ref / 2.220446049250313e-16)), "2.2e-16 grid"
    fs = np.array([1e-8, 0.1, 0.9, 1.0, 3.7, 25.0, 300.0, 5e3, 1e5])
    for d1, d2 in ((1.0, 96.0), (2.0, 994.0), (2.0, 6.0), (3.0, 20.0), (7.0, 1992.0), (1.0, 1e5), (2.0, 1994.0)):
        true = np.array([oracle_np.f_upper_mp(float(x), d1, d2) for x in fs])
        got = engine.dev_pf_upper(fs, d1, d2)
        assert np.all(np.abs(got - true) <= 5e-12 * true + 1e-305), (d1, d2)


# ---- the reference's own test expectations, through the public API on the GPU ------------------------
def test_get_diffNetworks_multi_omics(engine, bundled):
    """tests/testthat/test-core_functions.R:2-51"""
    assays = {k: bundled[k] for k in ("RNAseq", "Proteomics", "Olink")}
    deggs = md.get_diffNetworks(assayData=assays, metadata=bundled["metadata"], category_variable="response",
                                regression_method="lm", padj_method="bonferroni", verbose=False,
                                show_progressBar=False, cores=1, engine=engine)
    assert isinstance(deggs, md.Deggs) and tuple(deggs.keys()) == md.Deggs.FIELDS
    assert deggs["diffNetworks"]["RNAseq"]["sig_pvalues_count"] == 7
    assert deggs["diffNetworks"]["RNAseq"]["Responder"].shape[0] == 3
    assert deggs["diffNetworks"]["Proteomics"]["Responder"] == md.NO_LINKS
    assert deggs["diffNetworks"]["Olink"]["Non_responder"].shape[1] == 4
    assert deggs["diffNetworks"]["Olink"]["Responder"] == md.NO_LINKS
    nr = deggs["diffNetworks"]["RNAseq"]["Non_responder"]
    assert "TNF" in set(nr["from"]) and "TNFRSF1A" in set(nr["to"])


def test_real_R_numbers_held_by_the_reference_on_gpu(engine, bundled):
    """The reference's vignette screenshots hold real R output of this exact call (lm + bonferroni on the bundled data,
    vignettes/Differential_Network_Analysis.Rmd:75-83; multiDEGGs_1.png, multiDEGGs_2.png, plot_regressions.png):
    the engine, through the public API, must print the same numbers (helpers.R_PINNED)."""
    assays = {k: bundled[k] for k in ("RNAseq", "Proteomics", "Olink")}
    deggs = md.get_diffNetworks(assayData=assays, metadata=bundled["metadata"], category_variable="response",
                                regression_method="lm", padj_method="bonferroni", verbose=False,
                                show_progressBar=False, cores=2, engine=engine)
    check_r_pinned(deggs["diffNetworks"])
    # vignettes/Feature_Selection.Rmd:133-136: pairs kept by multiDEGGs_filter on the full bundled data
    y = bundled["metadata"]["response"].cat.codes.to_numpy() + 1.0
    res = md.multiDEGGs_filter(y=y, x=bundled["RNAseq"].T, engine=engine)
    assert set(res["pairs"]["from"] + ":" + res["pairs"]["to"]) == {
        "TNF:TNFRSF1A", "AKT2:MTOR", "IL1B:IL1R2", "FASLG:FAS", "TGFB3:TGFBR1", "MAP2K2:MAPK3", "FANCD2:FAN1"}


def test_get_multiOmics_diffNetworks_rlm_BH(engine, bundled):
    """tests/testthat/test-core_functions.R:55-91"""
    assays = {k: bundled[k] for k in ("RNAseq", "Olink")}
    deggs = md.get_diffNetworks(assayData=assays, metadata=bundled["metadata"], category_variable="response",
                                regression_method="rlm", padj_method="BH", verbose=False, show_progressBar=False,
                                cores=1, engine=engine)
    res = md.get_multiOmics_diffNetworks(deggs)
    assert len(res) == 2
    assert res["Non_responder"].shape[1] == 5
    assert list(res["Responder"].columns) == ["from", "to", "p.value", "p.adj", "layer"]
    assert res["Responder"].shape[0] == 3
    assert "AKT2" in set(res["Non_responder"]["from"]) and "MTOR" in set(res["Non_responder"]["to"])
    assert "FANCD2" in set(res["Responder"]["from"]) and "FAN1" in set(res["Responder"]["to"])


def test_get_sig_deggs_single_omic(engine, bundled):
    """tests/testthat/test-core_functions.R:95-127"""
    deggs = md.get_diffNetworks(assayData=bundled["RNAseq"], metadata=bundled["metadata"],
                                category_variable="response", regression_method="lm", padj_method="bonferroni",
                                verbose=False, show_progressBar=False, cores=1, engine=engine)
    res = md.get_sig_deggs(deggs)
    assert res.shape == (7, 4)
    for a, b in (("TNF", "TNFRSF1A"), ("TGFB3", "TGFBR1"), ("IL1B", "IL1R2")):
        assert a in set(res["from"]) and b in set(res["to"])


def test_multiDEGGs_filter_and_predict(engine, bundled):
    """tests/testthat/test-ML_functions.R:2-17, 57-69 (the nestedcv-free parts)"""
    x = bundled["RNAseq"].T
    y = bundled["metadata"]["response"].cat.codes.to_numpy() + 1.0
    res = md.multiDEGGs_filter(y=y, x=x, engine=engine)
    assert res["pairs"].shape == (7, 4)
    assert "IL1B" in set(res["pairs"]["from"]) and "IL1R2" in set(res["pairs"]["to"])
    assert "FASLG" in set(res["pairs"]["from"]) and "FAS" in set(res["pairs"]["to"])
    res5 = md.multiDEGGs_filter(y=y, x=x, nfilter=5, engine=engine)
    assert res5["pairs"].shape[0] == 5          # first-N rows in network order (fs:148-150)
    assert "FANCD2" in set(res5["pairs"]["from"]) and "FAN1" in set(res5["pairs"]["to"])
    feats = md.predict(res, x, engine=engine)
    assert feats.shape == (100, 7) and "TNF:TNFRSF1A" in feats.columns
    a, b = x["TNF"].to_numpy(), x["TNFRSF1A"].to_numpy()
    expect = np.where(np.isfinite(a / b), a / b, 0.0)
    assert np.array_equal(feats["TNF:TNFRSF1A"].to_numpy(), expect)


def test_cfg3_scaled_down(engine, oracle):
    wl = synth.cfg3(n_genes=3000, n_samples=120, max_edges=4000)
    prob = wl.problem()
    gpu, ref = _both(engine, oracle, prob, ("cutoff", "p_edge", "stat", "status", "tot_edges", "sig_count",
                                           "fail_count", "best", "cat_best", "padj_best", "median"))
    assert_parity(gpu, ref, prob, what="cfg3-small: ")


def test_cfg5_scaled_down(engine, oracle):
    wl = synth.cfg5(n_features=120, n_samples=200, planted=0.01)
    prob = wl.problem()
    gpu, ref = _both(engine, oracle, prob)
    assert_parity(gpu, ref, prob, what="cfg5-small: ")


def test_stats_and_launch_count(engine):
    prob = synth.random_problem(50, 40, 120, 2, 71)
    engine.run_omic(prob)
    st = engine.stats()
    assert st["kernel_launches"] > 5 and st["unique_fits"] == 120 and st["n_nodes"] > 0
    assert st["fit_bytes"] == 120 * (16 * 40 + 16)


@pytest.mark.parametrize("nf,n,K", [(300, 64, 2), (257, 100, 3), (130, 37, 2), (385, 520, 2)])
def test_dense_tile_path_matches_oracle(oracle, nf, n, K, monkeypatch):
    """All-pairs network through the dense cross-moment kernel (forced), same parity bar as the streaming kernel."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    monkeypatch.setenv("MDEGGS_FIT_PATH", "1")
    eng = md.Engine([0])
    rng = np.random.default_rng(nf)
    prob = synth.random_problem(nf, n, 10, K, 900 + nf)
    iu, ju = np.triu_indices(nf, k=1)
    sel = rng.permutation(iu.shape[0])[: min(iu.shape[0], 20000)]
    flip = rng.random(sel.shape[0]) < 0.3                       # some edges reversed: B ~ A vs A ~ B
    fr = np.where(flip, ju[sel], iu[sel]) + 1
    to = np.where(flip, iu[sel], ju[sel]) + 1
    prob.from_idx, prob.to_idx = fr.astype(np.int32), to.astype(np.int32)
    gpu = eng.run_omic(prob, ALL_FIELDS)
    assert eng.stats()["fit_path"] & 0xff == 1
    ref = oracle.run_omic(prob, ALL_FIELDS)
    assert_parity(gpu, ref, prob, what=f"dense nf{nf} n{n} K{K}: ")
    monkeypatch.setenv("MDEGGS_FIT_PATH", "0")
    eng0 = md.Engine([0])
    gpu0 = eng0.run_omic(prob, ALL_FIELDS)
    assert eng0.stats()["fit_path"] & 0xff == 0
    assert np.array_equal(gpu0["status"], gpu["status"]) and np.array_equal(gpu0["cat_best"], gpu["cat_best"])
    assert np.allclose(gpu0["stat"], gpu["stat"], rtol=1e-10, atol=0, equal_nan=True)
    # the cp.async form of the dense kernel (kept for drivers without tensor maps) against the TMA-fed one
    eng.set_option("dense_path", 0)
    gpu1 = eng.run_omic(prob, ALL_FIELDS)
    assert eng.stats()["fit_path"] & 0xff == 1
    assert np.array_equal(gpu1["status"], gpu["status"]) and np.array_equal(gpu1["cat_best"], gpu["cat_best"])
    assert np.allclose(gpu1["stat"], gpu["stat"], rtol=1e-9, atol=1e-12, equal_nan=True)
    eng.close()
    eng0.close()


def test_single_process_multi_gpu_matches_one_gpu(engine):
    """Edges sharded over the GPUs of one process (ncclCommInitAll + all-gather): identical to one GPU."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    eng2 = md.Engine(list(range(min(4, torch.cuda.device_count()))))
    for K, method in ((2, _abi.METHOD_LM), (3, _abi.METHOD_LM), (2, _abi.METHOD_RLM)):
        prob = synth.random_problem(80, 60, 333, K, 500 + K, method, _abi.PADJ_BH)
        a = engine.run_omic(prob, ALL_FIELDS)
        b = eng2.run_omic(prob, ALL_FIELDS)
        for k in a:
            assert np.array_equal(a[k], b[k], equal_nan=True), (K, method, k)
    assert eng2.stats()["n_ranks"] >= 2
    assert int(eng2.stats()["fit_path"]) & 0x4000          # p-value blocks pushed over peer memory (mdeggs_p2p.cuh)
    eng2.set_option("p2p", 0)                               # ... and the ncclAllGather path
    prob = synth.random_problem(80, 60, 333, 2, 502, _abi.METHOD_LM, _abi.PADJ_BH)
    a, b = engine.run_omic(prob, ALL_FIELDS), eng2.run_omic(prob, ALL_FIELDS)
    assert not int(eng2.stats()["fit_path"]) & 0x4000
    for k in a:
        assert np.array_equal(a[k], b[k], equal_nan=True), k
    eng2.set_option("p2p", 1)
    # best-percentile outputs only: the BH segments are sharded over the ranks and the counts all-reduced;
    # dense-tile path: every rank computes only the tile rows of its edge block
    want = _abi.DEFAULT_WANT
    for fit_path in (0, 1):
        engine.set_option("fit_path", fit_path)
        eng2.set_option("fit_path", fit_path)
        for padj in (_abi.PADJ_BH, _abi.PADJ_HOLM, _abi.PADJ_BONFERRONI, _abi.PADJ_QVALUE, _abi.PADJ_HOMMEL):
            prob = synth.random_problem(300, 50, 2500, 2, 900 + padj, _abi.METHOD_LM, padj)
            a = engine.run_omic(prob, want)
            b = eng2.run_omic(prob, want)
            for k in a:
                assert np.array_equal(a[k], b[k], equal_nan=True), (fit_path, padj, k)
    engine.set_option("fit_path", -1)
    eng2.close()


def test_one_process_per_gpu_peer_push_matches_one_gpu():
    """One process per GPU (torchrun, 2 ranks): the peer-memory all-gather (k_push_gather over IPC-mapped buffers) and the
    ncclAllGather path both reproduce the single-GPU results bit for bit, eagerly and under graph replay, across buffer
    re-allocations (tools/p2p_check.py)."""
    import os, subprocess, sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(root, "tools", "p2p_check.py")],
                         capture_output=True, text=True, timeout=600)
    assert res.returncode == 0 and "p2p_check ok" in res.stdout, res.stdout[-2000:] + res.stderr[-4000:]


def test_node_band_copy_and_leading_dimension(engine, oracle):
    """Only the band of node rows crosses PCIe; ld > G (a sub-matrix of a bigger R matrix) is honoured."""
    import ctypes as C
    base = synth.random_problem(200, 40, 10, 2, 81)
    rng = np.random.default_rng(81)
    pairs = set()
    while len(pairs) < 150:
        a, b = (int(v) for v in rng.integers(51, 121, size=2))      # nodes only in rows 51..120
        if a != b:
            pairs.add((a, b))
    fr = np.array([p[0] for p in pairs], dtype=np.int32)
    to = np.array([p[1] for p in pairs], dtype=np.int32)
    base.from_idx, base.to_idx = fr, to
    ref = oracle.run_omic(base, ALL_FIELDS)
    gpu = engine.run_omic(base, ALL_FIELDS)
    assert engine.stats()["h2d_bytes"] < 0.5 * base.assay.nbytes
    assert_parity(gpu, ref, base, what="band: ")
    # same data embedded in a taller column-major buffer: ld = G + 9
    G, n = base.assay.shape
    big = np.asfortranarray(np.full((G + 9, n), 777.0))
    big[:G, :] = base.assay
    cprob = base.to_c()
    cprob.assay, cprob.ld = big.ctypes.data, G + 9
    cres, arrays = _abi.alloc_result(base.dims, ALL_FIELDS)
    engine.run_omic_raw(cprob, cres)
    for k in gpu:
        assert np.array_equal(gpu[k], arrays[k], equal_nan=True), k


def test_graph_replay_is_bit_identical(oracle):
    """Third identical raw call is replayed from the captured CUDA graph: same bytes out, new inputs honoured."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    eng = md.Engine([0])
    prob = synth.random_problem(90, 70, 400, 3, 601)
    cprob = prob.to_c()
    cres, arr = _abi.alloc_result(prob.dims, ALL_FIELDS)
    eng.run_omic_raw(cprob, cres)
    first = {k: v.copy() for k, v in arr.items()}
    assert not (eng.stats()["fit_path"] & 0x100)
    for _ in range(3):
        eng.run_omic_raw(cprob, cres)
    assert eng.stats()["fit_path"] & 0x100, "graph replay did not engage"
    for k in first:
        assert np.array_equal(first[k], arr[k], equal_nan=True), k
    # same buffers, new CONTENT (values and percentiles): the replay must see it
    prob.assay[...] = np.asfortranarray(synth.random_problem(90, 70, 400, 3, 602).assay)
    prob._keep[4][...] = _abi.r_seq(0.25, 0.98, 0.06)[: len(prob.pct)]
    prob2 = _abi.Problem(assay=prob.assay.copy(order="F"), from_idx=prob.from_idx, to_idx=prob.to_idx, group=prob.group,
                         K=3, pct=prob._keep[4].copy(), padj=prob.padj)
    eng.run_omic_raw(cprob, cres)
    assert eng.stats()["fit_path"] & 0x100
    ref = oracle.run_omic(prob2, ALL_FIELDS)
    assert_parity(arr, ref, prob2, what="graph replay, new content: ")
    eng.set_option("graph", 0)
    eng.run_omic_raw(cprob, cres)
    assert not (eng.stats()["fit_path"] & 0x100)
    eng.close()


def test_quantile_select_tie_groups_and_overflow(oracle):
    """Radix-select corner paths: a tie group larger than the per-CTA shared-memory sort (slow in-CTA digit path) and
    a candidate list that overflows (falls back to scanning the matrix). Cut-offs stay bit-exact."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    eng = md.Engine([0])
    prob = synth.random_problem(500, 100, 600, 2, 701)
    v = np.round(prob.assay, 1)
    v[v < 4.9] = 0.0                      # ~45% exact zeros: one 64-bit-identical group of > 8192 values
    prob.assay = np.asfortranarray(v)
    prob.pct = np.array([0.05, 0.2, 0.35, 0.45, 0.5, 0.6, 0.75, 0.9, 0.99, 1.0])
    ref = oracle.run_omic(prob, ALL_FIELDS)
    gpu = eng.run_omic(prob, ALL_FIELDS)
    assert (v == 0.0).sum() > 8192 and ref["cutoff"][0] == 0.0
    assert_parity(gpu, ref, prob, what="tie group: ")
    eng.set_option("rsel_cap", 1000)       # force the compaction to overflow
    gpu2 = eng.run_omic(prob, ALL_FIELDS)
    assert np.array_equal(gpu2["cutoff"], ref["cutoff"])
    smooth = synth.random_problem(300, 80, 400, 3, 702)
    ref3 = oracle.run_omic(smooth, ("cutoff", "tot_edges", "best"))
    gpu3 = eng.run_omic(smooth, ("cutoff", "tot_edges", "best"))
    for k in ref3:
        assert np.array_equal(gpu3[k], ref3[k]), k
    eng.close()


def _cfg4_flow(engine, oracle, wl, folds, raw_parity):
    """multiDEGGs_filter on the full data and on every training fold (fs:77-90 as nestedcv calls it): the public API on
    the GPU must select exactly the pairs the same host flow selects from the oracle's numbers; with `raw_parity` the
    engine's per-edge / per-percentile outputs of every call are also compared field by field."""
    import pandas as pd
    from helpers import oracle_single_omic
    n = wl.assay.values.shape[1]
    x = pd.DataFrame(wl.assay.values.T, index=wl.assay.samples, columns=wl.assay.genes)
    y = wl.group.astype(float)
    n_with_pairs = 0
    for hold in [None] + folds:
        keep = np.ones(n, dtype=bool)
        if hold is not None:
            keep[hold] = False
        xs, ys = x.iloc[keep], y[keep]
        try:
            got = md.multiDEGGs_filter(ys, xs, engine=engine)
            got_pairs = list(got["pairs"].index) if got["pairs"].shape[0] else []
        except ValueError as e:
            got_pairs = str(e)
        fac = host.tidy_metadata(metadata=pd.Series(ys, index=xs.index))
        counts = md.Assay(np.ascontiguousarray(xs.to_numpy().T), wl.assay.genes, np.asarray(xs.index).astype(str))
        try:
            nets = oracle_single_omic(counts, "assayData1", fac, "lm", None, md.FILTER_PERCENTILES, "q.value")
            frames = [v for v in nets.values() if isinstance(v, pd.DataFrame)]
            sig = [f[(f["p.adj"] < 0.05).to_numpy()] for f in frames]
            exp_pairs = [i for f in sig for i in f.index]
            if len(exp_pairs) <= 1:
                exp_pairs = []          # t-test fallback branch (fs:157-162): no pairs
        except ValueError as e:
            exp_pairs = str(e)
        if isinstance(exp_pairs, str):
            # get_diffNetworks swallows the per-omic error (message + NULL layer); get_sig_deggs then finds nothing
            assert got_pairs == [] or isinstance(got_pairs, str)
        else:
            assert got_pairs == exp_pairs
            n_with_pairs += len(exp_pairs) > 0
        if raw_parity:
            prep = host.prepare_omic(counts, "assayData1", fac, "lm", None, md.FILTER_PERCENTILES, "q.value")
            want = ("cutoff", "median", "p_edge", "stat", "status", "tot_edges", "sig_count", "fail_count", "best",
                    "cat", "cat_best", "padj_best")
            gpu, ref = _both(engine, oracle, prep.problem, want)
            assert_parity(gpu, ref, prep.problem, what=f"cfg4 fold (held out {hold}): ")
    return n_with_pairs


def test_cfg4_filter_flow_scaled_down(engine, oracle):
    """Config 4 (multiDEGGs_filter inside nestedcv folds), scaled down: 1,500 features, 3 folds + full data."""
    wl = synth.cfg4(n_features=1500, n_samples=24, planted=0.15)
    folds = np.array_split(np.random.default_rng(4).permutation(24), 3)
    _cfg4_flow(engine, oracle, wl, folds, raw_parity=False)


def test_cfg4_full_size_ten_folds(engine, oracle):
    """Config 4 at BASELINE size: 5,000 features x 24 samples (12 / 12), 10 stratified folds + the full data = 11 calls
    of multiDEGGs_filter (15 percentiles, q.value), every call compared with the oracle field by field."""
    wl = synth.cfg4()
    assert wl.assay.values.shape == (5000, 24)
    rng = np.random.default_rng(20261020)
    folds = [[] for _ in range(10)]
    for c in (1, 2):  # stratified: deal each category's samples round the folds
        idx = rng.permutation(np.flatnonzero(wl.group == c))
        for j, s in enumerate(idx.tolist()):
            folds[(j + (c - 1) * 5) % 10].append(s)
    folds = [np.array(sorted(f)) for f in folds]
    assert sorted(np.concatenate(folds).tolist()) == list(range(24))
    assert _cfg4_flow(engine, oracle, wl, folds, raw_parity=True) >= 1


@pytest.mark.parametrize("padj", ["none", "holm", "hochberg", "BY", "hommel", "q.value"])
def test_host_adjusted_methods_through_api(engine, oracle, bundled, padj):
    """padj methods the device leaves to the host (or 'none'): same networks from engine and oracle numbers."""
    import pandas as pd
    from helpers import oracle_single_omic
    fac = host.tidy_metadata(metadata=bundled["metadata"], category_variable="response")
    got = md.get_diffNetworks_singleOmic(bundled["RNAseq"], "RNAseq", fac, "lm", network=None,
                                         percentile_vector=md.DEFAULT_PERCENTILES, padj_method=padj, engine=engine)
    exp = oracle_single_omic(bundled["RNAseq"], "RNAseq", fac, "lm", None, md.DEFAULT_PERCENTILES, padj)
    assert got.keys() == exp.keys() and got["sig_pvalues_count"] == exp["sig_pvalues_count"]
    for k in got:
        if isinstance(got[k], pd.DataFrame):
            assert list(got[k].index) == list(exp[k].index) and list(got[k].columns) == list(exp[k].columns)
            for col in got[k].columns[2:]:
                assert np.allclose(got[k][col].to_numpy(), exp[k][col].to_numpy(), rtol=1e-9, atol=4.5e-16, equal_nan=True)
        else:
            assert got[k] == exp[k]


@pytest.mark.parametrize("n,K,method", [(2600, 2, _abi.METHOD_LM), (2600, 2, _abi.METHOD_RLM), (3300, 3, _abi.METHOD_LM),
                                        (9000, 2, _abi.METHOD_RLM), (30000, 2, _abi.METHOD_LM)])
def test_large_sample_counts(engine, oracle, n, K, method):
    """Segments beyond the register-resident K2 kernel (> 1,024 samples per category: row staged in shared memory),
    beyond shared memory (30,000 samples: global-memory path), and rlm with the residual-only staging."""
    prob = synth.random_problem(24, n, 40, K, 1000 + n, method, _abi.PADJ_BH)
    gpu, ref = _both(engine, oracle, prob)
    assert_parity(gpu, ref, prob, rtol=1e-8 if method == _abi.METHOD_RLM else 1e-9, what=f"n={n}: ")


def test_rlm_too_many_samples_is_reported(engine):
    prob = synth.random_problem(8, 40000, 6, 2, 77, _abi.METHOD_RLM, _abi.PADJ_BH)
    with pytest.raises(md.EngineError, match="unsupported"):
        engine.run_omic(prob)


def test_invalid_edge_index_is_reported(engine):
    prob = synth.random_problem(30, 20, 40, 2, 78)
    prob.to_idx = prob.to_idx.copy()
    prob.to_idx[7] = 31                       # > G
    with pytest.raises(md.EngineError, match="outside 1..G"):
        engine.run_omic(prob)
    big = synth.random_problem(30, 20, 40, 2, 79)   # the context stays usable afterwards
    assert engine.run_omic(big)["best"].shape == (1,)


@pytest.mark.parametrize("padj", [_abi.PADJ_BH, _abi.PADJ_HOLM, _abi.PADJ_HOCHBERG, _abi.PADJ_BY, _abi.PADJ_BONFERRONI])
def test_bucket_order_equals_radix_sort(padj):
    """K8 ordering through p-value buckets (default up to 2^26 edges) must give bit-identical adjusted values,
    counts and best percentile as the CUB radix-sort path, including tie groups (p == 1 self-loops, NaN)."""
    eng = md.Engine([0])
    eng.set_option("graph", 0)
    want = ("p_edge", "padj", "padj_best", "sig_count", "tot_edges", "best", "cat_best", "status")
    for K, seed, E in ((2, 11, 5000), (3, 12, 40000)):
        prob = synth.random_problem(300, 64, E, K, seed, _abi.METHOD_LM, padj)
        prob.from_idx[:50] = prob.to_idx[:50]              # self-loops: a tie group at p == 1
        prob.assay[7, 3] = np.nan                          # a non-finite gene: NaN p-values (sorted last)
        eng.set_option("sort_path", 0)
        ref = eng.run_omic(prob, want)
        assert not (eng.stats()["fit_path"] & 0x200)
        eng.set_option("sort_path", 1)
        got = eng.run_omic(prob, want)
        assert eng.stats()["fit_path"] & 0x200, "bucket ordering did not engage"
        for k in want:
            assert np.array_equal(got[k], ref[k], equal_nan=True), (K, k)
    eng.close()


def test_bucket_order_oversize_buckets(oracle):
    """Degenerate p-value distributions: thousands of identical p-values (one oversize tie bucket) and thousands of
    distinct p-values inside one bucket (global-memory sort path)."""
    eng = md.Engine([0])
    eng.set_option("graph", 0)
    rng = np.random.default_rng(5)
    G, n, K = 400, 40, 2
    # (a) all edges are the same gene pair repeated -> one huge tie group; (b) edges between near-identical genes:
    # tiny p-values of almost the same magnitude -> one binade bucket with thousands of distinct keys
    base = rng.normal(5, 1, size=n)
    assay = rng.normal(5, 1, size=(G, n))
    grp = (np.arange(n) % 2 + 1).astype(np.int32)
    sign = np.where(grp == 1, 1.0, -1.0)
    for g in range(200):
        assay[g] = base * sign + rng.normal(0, 1e-3, size=n) if g % 2 else base + rng.normal(0, 1e-3, size=n)
    assay = np.asfortranarray(assay)
    E = 6000
    fr = np.concatenate([np.full(3000, 301), 2 * rng.integers(0, 100, 3000) + 1]).astype(np.int32)
    to = np.concatenate([np.full(3000, 302), 2 * rng.integers(0, 100, 3000) + 2]).astype(np.int32)
    prob = _abi.Problem(assay=assay, from_idx=fr, to_idx=to, group=grp, K=K, pct=np.array([0.1, 0.5]),
                     method=_abi.METHOD_LM, padj=_abi.PADJ_BH)
    want = ("p_edge", "padj", "sig_count", "tot_edges", "best", "status")
    eng.set_option("sort_path", 0)
    ref = eng.run_omic(prob, want)
    eng.set_option("sort_path", 1)
    got = eng.run_omic(prob, want)
    for k in want:
        assert np.array_equal(got[k], ref[k], equal_nan=True), k
    orc = oracle.run_omic(prob, want)
    assert np.array_equal(got["sig_count"], orc["sig_count"])
    eng.close()


# ---- BASELINE.json's full sizes -------------------------------------------------------------------------------
def _bh_properties(res, padj_code=_abi.PADJ_BH):
    """Size-independent invariants of one result: counts agree with the rows, BH is monotone in p inside a category."""
    best = int(res["best"][0])
    assert best >= 1
    cat, padj, p = res["cat_best"], res["padj_best"], res["p_edge"]
    tested = cat != 0
    assert int(tested.sum()) == int(res["tot_edges"][best - 1])
    assert int(np.sum(padj[tested] < 0.05)) == int(res["sig_count"][best - 1])
    assert np.all(np.isnan(padj[~tested]))
    assert np.all((padj[tested] >= p[tested]) | np.isnan(p[tested]))  # an adjusted p never drops below the raw p
    assert np.nanmax(padj[tested]) <= 1.0
    for c in np.unique(cat[tested]):
        m = (cat == c) & ~np.isnan(p)
        o = np.argsort(p[m], kind="stable")
        assert np.all(np.diff(padj[m][o]) >= 0), "BH values must be non-decreasing in p inside a segment"


def test_cfg3_full_size_vs_oracle(engine, oracle):
    """Config 3 at BASELINE.json's size (20,000 x 1,000, K = 3, 53,035 edges): whole-result parity with the oracle
    (the oracle's unique-edge mode needs a couple of seconds for it) plus the size-independent invariants."""
    wl = synth.cfg3()
    prob = wl.problem()
    want = ("cutoff", "median", "p_edge", "stat", "status", "tot_edges", "sig_count", "fail_count", "best",
            "cat_best", "padj_best")
    gpu = engine.run_omic(prob, want)
    ref = oracle.run_omic(prob, want)
    assert_parity(gpu, ref, prob, what="cfg3-full: ")
    _bh_properties(gpu)
    # edge-permutation invariance: per-edge results follow the permutation bit for bit, the counts do not move
    rng = np.random.default_rng(3)
    perm = rng.permutation(prob.from_idx.shape[0])
    prob2 = _abi.Problem(assay=prob.assay, from_idx=prob.from_idx[perm].copy(), to_idx=prob.to_idx[perm].copy(),
                         group=prob.group, K=prob.K, pct=prob.pct, method=prob.method, padj=prob.padj)
    gpu2 = engine.run_omic(prob2, want)
    for k in ("p_edge", "stat", "status", "cat_best", "padj_best"):
        assert np.array_equal(gpu2[k], gpu[k][perm], equal_nan=True), k
    for k in ("cutoff", "tot_edges", "sig_count", "best"):
        assert np.array_equal(gpu2[k], gpu[k]), k


def test_cfg5_full_size_sampled_oracle_and_properties(engine, oracle):
    """Config 5 at BASELINE.json's size (4,000 x 2,000, all 7,998,000 pairs): the oracle re-fits a random sample of
    the edges (per-edge statistics depend on the two rows only), the rest is covered by invariants and by the
    streaming and dense-tile fit paths agreeing with each other on every edge."""
    wl = synth.cfg5()
    prob = wl.problem()
    E = prob.from_idx.shape[0]
    assert E == 7_998_000
    want = ("p_edge", "stat", "status", "tot_edges", "sig_count", "best", "cat_best", "padj_best", "cutoff")
    gpu = engine.run_omic(prob, want)
    assert engine.stats()["fit_path"] & 0xff == 1      # dense tile path chosen for an all-pairs network
    _bh_properties(gpu)
    rng = np.random.default_rng(11)
    pick = np.sort(rng.choice(E, size=4000, replace=False))
    sub = _abi.Problem(assay=prob.assay, from_idx=prob.from_idx[pick].copy(), to_idx=prob.to_idx[pick].copy(),
                       group=prob.group, K=prob.K, pct=prob.pct, method=prob.method, padj=prob.padj)
    ref = oracle.run_omic(sub, ("p_edge", "stat", "status"))
    assert np.array_equal(gpu["status"][pick], ref["status"])
    assert rel_close(gpu["stat"][pick], ref["stat"], 1e-9).all()
    assert rel_close(gpu["p_edge"][pick], ref["p_edge"], 1e-9, 4.5e-16).all()
    # the cut-offs are order statistics of the whole 4,000 x 2,000 matrix: every row is a node here
    flat = np.sort(prob.assay, axis=None)
    N = flat.shape[0]
    for pc, got in zip(prob.pct, gpu["cutoff"]):
        idx = 1.0 + (N - 1) * pc
        lo_i = int(np.floor(idx))
        q = flat[lo_i - 1]
        hi_v = flat[int(np.ceil(idx)) - 1]
        if idx > lo_i and hi_v != q:
            h = idx - lo_i
            q = (1.0 - h) * q + h * hi_v
        assert got == q
    # streaming path on the same problem: same edges tested, statistics within 1e-9 (different summation order)
    engine.set_option("fit_path", 0)
    try:
        gpu_s = engine.run_omic(prob, ("stat", "status", "tot_edges", "best"))
    finally:
        engine.set_option("fit_path", -1)
    assert np.array_equal(gpu_s["status"], gpu["status"])
    assert np.array_equal(gpu_s["tot_edges"], gpu["tot_edges"])
    assert rel_close(gpu_s["stat"], gpu["stat"], 1e-9, 1e-10).all()


def _quantile_problem(values: np.ndarray, pct, seed=0, K=2):
    """Every row is a network node (a ring of edges), so the cut-offs are order statistics of the whole matrix."""
    G, n = values.shape
    rng = np.random.default_rng(seed)
    group = (np.arange(n) % K + 1).astype(np.int32)
    rng.shuffle(group)
    fr = np.arange(1, G + 1, dtype=np.int32)
    to = np.roll(fr, -1)
    return _abi.Problem(assay=np.asfortranarray(values), from_idx=fr, to_idx=to, group=group, K=K,
                        pct=np.asarray(pct, dtype=np.float64), method=_abi.METHOD_LM, padj=_abi.PADJ_BH)


@pytest.mark.parametrize("kind", ["lognormal_outliers", "counts", "few_values", "constant", "with_nan", "negative",
                                  "two_genes", "wide_dynamic_range", "sorted_rows", "bimodal_gap"])
def test_quantile_select_distributions(engine, oracle, kind):
    """K3 is sample-adaptive: whatever the value distribution (heavy tails, integer counts with huge tie groups,
    constants, NaN, gaps wider than the sample sees), every cut-off must stay bit-identical to the oracle's type-7
    quantile of the whole node matrix, for interior and extreme percentiles alike."""
    import zlib
    rng = np.random.default_rng(zlib.crc32(kind.encode()))
    G, n = 600, 96
    if kind == "lognormal_outliers":
        v = np.exp(rng.normal(0, 2.5, size=(G, n)))
        v[rng.integers(0, G, 20), rng.integers(0, n, 20)] = 1e12
    elif kind == "counts":
        v = rng.negative_binomial(2, 0.3, size=(G, n)).astype(np.float64)      # integers, ~10 % zeros
        v[: G // 2] = np.floor(np.log2(1.0 + v[: G // 2]) * 4) / 4
    elif kind == "few_values":
        v = rng.choice([0.0, 1.5, 1.5000000001, 7.25], size=(G, n), p=[0.6, 0.2, 0.1, 0.1])
    elif kind == "constant":
        v = np.full((G, n), 3.25)
    elif kind == "with_nan":
        v = rng.normal(5, 1, size=(G, n))
        v[rng.random((G, n)) < 0.07] = np.nan
    elif kind == "negative":
        v = rng.normal(-3, 4, size=(G, n))
        v[0, 0], v[1, 1] = -0.0, 0.0
    elif kind == "two_genes":
        G = 2
        v = rng.normal(5, 1, size=(G, n))
    elif kind == "wide_dynamic_range":
        v = 10.0 ** rng.uniform(-200, 200, size=(G, n))
    elif kind == "sorted_rows":
        v = np.sort(rng.normal(5, 1, size=(G, n)), axis=0)     # the stratified sample sees a biased slice per row
    else:  # bimodal_gap: two narrow modes, the targets fall on both sides of an empty stretch
        v = np.where(rng.random((G, n)) < 0.5, rng.normal(1, 1e-6, size=(G, n)), rng.normal(1e6, 1e-3, size=(G, n)))
    pct = [0.0, 0.001, 0.05, 0.35, 0.5, 0.6, 0.65, 0.9, 0.98, 0.999, 1.0]
    prob = _quantile_problem(np.asarray(v, dtype=np.float64), pct)
    want = ("cutoff", "median", "tot_edges", "best")
    gpu = engine.run_omic(prob, want)
    ref = oracle.run_omic(prob, want)
    assert np.array_equal(gpu["cutoff"], ref["cutoff"], equal_nan=True), (kind, gpu["cutoff"], ref["cutoff"])
    assert np.array_equal(gpu["median"], ref["median"], equal_nan=True), kind
    assert np.array_equal(gpu["tot_edges"], ref["tot_edges"]), kind


def test_quantile_select_large_matrix_with_ties(engine, oracle):
    """A matrix large enough that tie groups exceed the shared-memory finish (fat slots answered by min == max) while
    other targets take the regular candidate path; also with a tiny candidate buffer (everything fat -> slow path)."""
    rng = np.random.default_rng(99)
    G, n = 3000, 200
    v = np.round(rng.normal(5, 1, size=(G, n)), 2)
    v[v < 4.2] = 0.0
    v[(v > 5.5) & (v < 5.6)] = 5.55
    pct = [0.01, 0.1, 0.2, 0.25, 0.5, 0.7, 0.72, 0.74, 0.9, 0.99]
    prob = _quantile_problem(v, pct)
    ref = oracle.run_omic(prob, ("cutoff",))
    gpu = engine.run_omic(prob, ("cutoff",))
    assert np.array_equal(gpu["cutoff"], ref["cutoff"])
    eng = md.Engine([0])
    eng.set_option("rsel_cap", 64)
    gpu2 = eng.run_omic(prob, ("cutoff",))
    eng.close()
    assert np.array_equal(gpu2["cutoff"], ref["cutoff"])


@pytest.mark.parametrize("E", [1, 2, 33, 120_000])
def test_bucket_order_edge_counts(oracle, E):
    """K8 ordering at the extremes: one and two edges, and enough edges for the device-wide bucket scan."""
    eng = md.Engine([0])
    G = 20 if E < 100 else 600
    prob = synth.random_problem(G, 40, E, 2, 500 + E)
    want = ("p_edge", "padj", "sig_count", "tot_edges", "best")
    got = eng.run_omic(prob, want)
    eng.set_option("sort_path", 0)
    ref_sort = eng.run_omic(prob, want)
    eng.close()
    for k in want:
        assert np.array_equal(got[k], ref_sort[k], equal_nan=True), k
    if E <= 33:
        ref = oracle.run_omic(prob, want)
        assert np.array_equal(got["sig_count"], ref["sig_count"]) and np.array_equal(got["best"], ref["best"])


@pytest.mark.parametrize("G,n,E,K,seed", [(60, 30, 200, 2, 41), (300, 64, 5000, 2, 42), (200, 90, 3000, 3, 43),
                                          (40, 24, 60, 2, 44), (500, 40, 20000, 2, 45)])
def test_device_qvalue_matches_oracle(engine, oracle, G, n, E, K, seed):
    """padj "q.value" on the device (pi0 smoother + q-values per (percentile, category) segment) against the oracle's
    restatement of qvalue::qvalue, including the reference's fallback rules (core_functions.R:1012-1025)."""
    prob = synth.random_problem(G, n, E, K, seed, _abi.METHOD_LM, _abi.PADJ_QVALUE,
                                pct=_abi.r_seq(0.25, 0.98, 0.05))
    want = ("p_edge", "status", "cat", "padj", "padj_best", "cat_best", "sig_count", "tot_edges", "fail_count", "best")
    gpu = engine.run_omic(prob, want)
    ref = oracle.run_omic(prob, want)
    assert_parity(gpu, ref, prob, what=f"q.value G{G} E{E} K{K}: ")


def test_device_qvalue_fallback_branches(engine, oracle):
    """Segments where qvalue() errors: a single tested edge (NA), all p-values below max(lambda) (pi0 = 1 retry)."""
    rng = np.random.default_rng(77)
    G, n = 40, 60
    grp = (np.arange(n) % 2 + 1).astype(np.int32)
    v = rng.normal(5, 1, size=(G, n))
    # strongly planted pairs: every tested p-value is tiny (< 0.95 = max lambda) -> qvalue() stops -> pi0 = 1
    for g in range(0, 20, 2):
        z = rng.normal(0, 1, n)
        v[g] = 5 + 2 * z
        v[g + 1] = 5 + 2 * np.where(grp == 1, z, -z) + rng.normal(0, 0.05, n)
    fr = np.arange(1, 20, 2, dtype=np.int32)
    to = fr + 1
    prob = _abi.Problem(assay=np.asfortranarray(v), from_idx=fr, to_idx=to.astype(np.int32), group=grp, K=2,
                        pct=np.array([0.0, 0.2, 0.5]), method=_abi.METHOD_LM, padj=_abi.PADJ_QVALUE)
    want = ("p_edge", "cat", "padj", "sig_count", "tot_edges", "best", "padj_best")
    gpu = engine.run_omic(prob, want)
    ref = oracle.run_omic(prob, want)
    assert_parity(gpu, ref, prob, what="q.value fallback: ")
    assert ref["tot_edges"].max() > 0
    one = _abi.Problem(assay=np.asfortranarray(v), from_idx=fr[:1].copy(), to_idx=to[:1].astype(np.int32), group=grp,
                       K=2, pct=np.array([0.0]), method=_abi.METHOD_LM, padj=_abi.PADJ_QVALUE)
    g1, r1 = engine.run_omic(one, want), oracle.run_omic(one, want)
    assert_parity(g1, r1, one, what="q.value single row: ")


# ---- all omic layers of one call as one launch set (mdeggs_submit_omic / mdeggs_wait / mdeggs_run_omics) ----
def _frames_equal(a, b):
    assert type(a) is type(b)
    if isinstance(a, pd.DataFrame):
        pd.testing.assert_frame_equal(a, b, check_exact=True)
    elif isinstance(a, dict):
        assert list(a) == list(b)
        for k in a:
            _frames_equal(a[k], b[k])
    else:
        assert a == b


@pytest.mark.parametrize("method,padj", [("lm", "bonferroni"), ("rlm", "BH"), ("lm", "q.value")])
def test_batched_layers_equal_sequential_layers(engine, bundled, method, padj):
    """get_diffNetworks without an explicit engine runs its layers through mdeggs_run_omics (one context per layer);
    the result must be identical to the layer-by-layer calls (core_functions.R:214-234)."""
    assays = {k: bundled[k] for k in ("RNAseq", "Proteomics", "Olink")}
    kw = dict(assayData=assays, metadata=bundled["metadata"], category_variable="response", regression_method=method,
              padj_method=padj, verbose=False, show_progressBar=False, cores=1)
    seq = md.get_diffNetworks(engine=engine, **kw)
    bat = md.get_diffNetworks(**kw)
    assert list(bat["diffNetworks"]) == list(seq["diffNetworks"]) == list(assays)
    _frames_equal(bat["diffNetworks"], seq["diffNetworks"])
    bat2 = md.get_diffNetworks(**kw)  # second batched call: every context replays its graph
    _frames_equal(bat2["diffNetworks"], seq["diffNetworks"])


def test_run_omics_layers_of_different_shapes_and_a_failing_layer(engine, oracle):
    from multideggs_b200.engine import EngineError, EnginePool
    pool = EnginePool(0)
    probs = [synth.random_problem(60, 40, 150, 2, 1201, _abi.METHOD_LM, _abi.PADJ_BH),
             synth.random_problem(300, 64, 2500, 3, 1202, _abi.METHOD_LM, _abi.PADJ_BONFERRONI),
             synth.random_problem(30, 50, 40, 2, 1203, _abi.METHOD_RLM, _abi.PADJ_BH),
             synth.random_problem(30, 20, 40, 2, 1204)]
    probs[3].from_idx = probs[3].from_idx.copy()
    probs[3].from_idx[5] = 31  # outside 1..G: this layer fails, the others must not notice
    try:
        for _ in range(2):
            res = pool.run_omics(probs, [ALL_FIELDS] * 4)
            assert isinstance(res[3], EngineError) and "outside 1..G" in str(res[3])
            for pr, got in zip(probs[:3], res[:3]):
                assert_parity(got, oracle.run_omic(pr, ALL_FIELDS), pr, rtol=1e-8 if pr.method == _abi.METHOD_RLM else 1e-9)
    finally:
        pool.close()


def test_submit_wait_protocol(engine):
    from multideggs_b200.engine import EngineError
    prob = synth.random_problem(50, 40, 120, 2, 1210)
    ref = engine.run_omic(prob, ALL_FIELDS)
    cp = prob.to_c()
    cr, arrays = _abi.alloc_result(prob.dims, tuple(ALL_FIELDS))
    with pytest.raises(EngineError, match="without a submitted call"):
        engine.wait()
    engine.submit_omic_raw(cp, cr)
    with pytest.raises(EngineError, match="still pending"):
        engine.submit_omic_raw(cp, cr)
    engine.wait()
    for k in ref:
        np.testing.assert_array_equal(arrays[k], ref[k])


# ---- device-resident callers (MDEGGS_MEM_DEVICE): what bench.py's `value` leg and a GPU-side pipeline use ----
def _device_call(torch, prob, want):
    dev = torch.device("cuda", 0)
    G, n, E, K, P = prob.dims
    t_assay = torch.from_numpy(np.ascontiguousarray(np.asfortranarray(prob.assay).T)).to(dev)  # memory: col-major G x n
    t_from = torch.from_numpy(np.ascontiguousarray(prob.from_idx, dtype=np.int32)).to(dev)
    t_to = torch.from_numpy(np.ascontiguousarray(prob.to_idx, dtype=np.int32)).to(dev)
    group = np.ascontiguousarray(prob.group, dtype=np.int32)
    pct = np.ascontiguousarray(prob.pct, dtype=np.float64)
    cprob = _abi.CProblem(t_assay.data_ptr(), G, G, n, _abi.LAYOUT_COLMAJOR, _abi.MEM_DEVICE, t_from.data_ptr(),
                          t_to.data_ptr(), E, group.ctypes.data, K, pct.ctypes.data, P, prob.method, prob.padj,
                          prob.tested_coef, prob.rlm_cov_mode)
    cres, outs = _abi.CResult(), {}
    tdt = {np.float64: torch.float64, np.int32: torch.int32, np.int8: torch.int8, np.int64: torch.int64}
    for nm in want:
        dt, shp = _abi.RESULT_FIELDS[nm]
        t = torch.full(shp(G, n, E, K, P), -7, dtype=tdt[dt], device=dev)  # poison: every field must be written
        outs[nm] = t
        setattr(cres, nm, t.data_ptr())
    return cprob, cres, outs, (t_assay, t_from, t_to, group, pct)


@pytest.mark.parametrize("K,method,padj", [(2, _abi.METHOD_LM, _abi.PADJ_BH), (3, _abi.METHOD_LM, _abi.PADJ_HOLM),
                                           (2, _abi.METHOD_RLM, _abi.PADJ_BONFERRONI), (2, _abi.METHOD_LM, _abi.PADJ_NONE),
                                           (3, _abi.METHOD_LM, _abi.PADJ_QVALUE)])
def test_device_resident_caller_equals_host_caller(K, method, padj):
    """Same pointers, new data on every call: call 1 is eager, call 2 captures the graph (output kernel included) and
    launches it, calls 3-4 replay it; every result must equal the host-buffer call on the same data, bit for bit."""
    import torch
    href = md.Engine([0])
    eng = md.Engine([0])
    prob = synth.random_problem(120, 48, 700, K, 1300 + K, method, padj)
    cprob, cres, outs, keep = _device_call(torch, prob, ALL_FIELDS)
    try:
        for it in range(4):
            fresh = synth.random_problem(120, 48, 700, K, 1310 + 7 * it + K, method, padj)
            prob.assay[...] = np.asfortranarray(fresh.assay)
            keep[0].copy_(torch.from_numpy(np.ascontiguousarray(np.asfortranarray(prob.assay).T)))
            for t in outs.values():
                t.fill_(-7)
            eng.run_omic_raw(cprob, cres)
            torch.cuda.synchronize()
            ref = href.run_omic(prob, ALL_FIELDS)
            replayed = bool(eng.stats()["fit_path"] & 0x100)
            assert replayed == (it >= 1), (it, replayed)
            for nm in ALL_FIELDS:
                got = outs[nm].cpu().numpy()
                if nm in ("padj", "padj_best") and padj == _abi.PADJ_NONE:
                    continue  # no adjusted values exist: neither caller's buffer is written
                assert np.array_equal(got, ref[nm], equal_nan=True), (it, nm)
    finally:
        eng.close()
        href.close()


# ---- K8 without keeping the adjusted values: significant counts from one rank pass, best row re-adjusted ----
@pytest.mark.parametrize("padj", [_abi.PADJ_BH, _abi.PADJ_BONFERRONI, _abi.PADJ_HOLM, _abi.PADJ_HOCHBERG, _abi.PADJ_BY,
                                  _abi.PADJ_QVALUE])
@pytest.mark.parametrize("K,E,seed", [(2, 3000, 1401), (3, 5000, 1402), (2, 40, 1403)])
def test_count_only_adjustment_equals_three_pass(padj, K, E, seed):
    """Big jobs do not keep P x E adjusted values: K8 then counts p.adj < 0.05 with the step-up rule (k_bh_sigrank) and
    re-adjusts only the chosen percentile.  Forced here on small problems ("padj_keep_cap" = 0) and compared, bit for
    bit, with the path that keeps everything; also with "count_only" = 0 (three passes + final-pass rerun)."""
    want = tuple(f for f in ALL_FIELDS if f != "padj")
    prob = synth.random_problem(200, 60, E, K, seed, _abi.METHOD_LM, padj)
    # plant some strong pairs so that several percentiles have significant counts > 0
    rng = np.random.default_rng(seed)
    a = np.asfortranarray(prob.assay.copy())
    for e in rng.choice(E, size=max(4, E // 40), replace=False):
        i, j = prob.from_idx[e] - 1, prob.to_idx[e] - 1
        cols = np.flatnonzero(np.asarray(prob.group) == 1)
        a[j, cols] = a[i, cols] * (1.5 + 0.01 * rng.standard_normal(cols.size)) + 0.05 * rng.standard_normal(cols.size)
    prob.assay[...] = a
    full = md.Engine([0])
    engs = {}
    try:
        ref = full.run_omic(prob, ALL_FIELDS)
        assert ref["sig_count"].max() > 0
        for name, opts in (("count", {"padj_keep_cap": 0}), ("three", {"padj_keep_cap": 0, "count_only": 0})):
            e = md.Engine([0])
            engs[name] = e
            for k, v in opts.items():
                e.set_option(k, v)
            for rep in range(3):  # eager, capture, replay
                got = e.run_omic(prob, want)
                for f in want:
                    assert np.array_equal(got[f], ref[f], equal_nan=True), (name, rep, f)
    finally:
        full.close()
        for e in engs.values():
            e.close()


@pytest.mark.parametrize("padj", [_abi.PADJ_BH, _abi.PADJ_BONFERRONI, _abi.PADJ_HOLM, _abi.PADJ_HOCHBERG, _abi.PADJ_BY])
def test_count_only_skips_tiles_beyond_005(padj):
    """The count-only K8 does not visit the tiles of the p order that start at p >= 0.05 (no adjusted value below 0.05 can
    come from there; segment sizes from k_percolate; holm takes the rank of the first skipped member).  Many tiles, strong
    and null pairs, NaN p-values: counts, best percentile and its rows equal the path that keeps everything."""
    want = tuple(f for f in ALL_FIELDS if f != "padj")
    for K, seed in ((2, 1451), (3, 1452)):
        prob = synth.random_problem(300, 48, 30000, K, seed, _abi.METHOD_LM, padj, shift=1.2)
        prob.assay[5, 2] = np.nan
        rng = np.random.default_rng(seed)
        a = np.asfortranarray(prob.assay.copy())
        for e in rng.choice(30000, size=1500, replace=False):
            i, j = prob.from_idx[e] - 1, prob.to_idx[e] - 1
            cols = np.flatnonzero(np.asarray(prob.group) == 1 + int(e) % K)
            a[j, cols] = a[i, cols] * 1.5 + 0.05 * rng.standard_normal(cols.size)
        prob.assay[...] = a
        full, cnt, noskip = md.Engine([0]), md.Engine([0]), md.Engine([0])
        try:
            cnt.set_option("padj_keep_cap", 0)
            noskip.set_option("padj_keep_cap", 0)
            noskip.set_option("count_skip", 0)
            ref = full.run_omic(prob, ALL_FIELDS)
            assert (ref["sig_count"] > 0).sum() >= 2
            for e in (cnt, noskip):
                got = e.run_omic(prob, want)
                for f in want:
                    assert np.array_equal(got[f], ref[f], equal_nan=True), (K, f)
        finally:
            for e in (full, cnt, noskip):
                e.close()


# ---- hommel on the device (SURVEY section 8 f, N2) ----
@pytest.mark.parametrize("G,n,E,K,seed", [(40, 30, 60, 2, 1501), (120, 64, 900, 2, 1502), (200, 60, 5000, 3, 1503),
                                          (64, 40, 300, 4, 1504), (30, 24, 3, 2, 1505)])
def test_device_hommel_matches_oracle(engine, oracle, G, n, E, K, seed):
    """stats::p.adjust(method = "hommel") per (percentile, category) segment: O(n^2) on the device (k_hm_*), compared
    with the C restatement of R's loop; segments of 0, 1, 2 and many members occur across the percentiles."""
    prob = synth.random_problem(G, n, E, K, seed, _abi.METHOD_LM, _abi.PADJ_HOMMEL)
    for rep in range(3):  # eager, capture, replay
        gpu, ref = _both(engine, oracle, prob)
        assert_parity(gpu, ref, prob, what=f"hommel G{G} E{E} K{K} rep{rep}: ")
    assert ref["tot_edges"].sum() > 0
    # the same numbers as the host-side adjustment of the raw p-values (the path bigger jobs take)
    raw = synth.random_problem(G, n, E, K, seed, _abi.METHOD_LM, _abi.PADJ_HOST)
    r = engine.run_omic(raw, ALL_FIELDS)
    P = len(prob.pct)
    cat = r["cat"].reshape(P, E)
    for pi in range(P):
        for k in range(1, K + 1):
            msk = cat[pi] == k
            if msk.any():
                exp = host.p_adjust(r["p_edge"][msk], "hommel")
                assert np.allclose(gpu["padj"].reshape(P, E)[pi, msk], exp, rtol=1e-12, atol=0, equal_nan=True)


def test_device_hommel_edge_limit(engine):
    big = synth.random_problem(400, 20, _abi.HOMMEL_MAX_E + 1, 2, 1510, _abi.METHOD_LM, _abi.PADJ_HOMMEL)
    with pytest.raises(md.EngineError, match="hommel"):
        engine.run_omic(big)
    assert host._padj_code("hommel", _abi.HOMMEL_MAX_E + 1) == _abi.PADJ_HOST
    assert host._padj_code("hommel", 1000) == _abi.PADJ_HOMMEL


@pytest.mark.parametrize("G,n,E,K,row_major,with_nan", [
    (333, 90, 500, 2, False, False),      # odd leading dimension: the tensor map is refused, plain loader
    (400, 90, 500, 3, False, True),       # even: TMA boxes; NaN-bearing rows
    (400, 150, 700, 2, True, False),      # row-major: plain loader
    (96, 1000, 300, 3, False, False),     # config-3 segment size (12 values per lane), 8-row tiles, 5 boxes
    (64, 1500, 200, 2, False, False),     # 750 per category: 32 values per lane
    (64, 2000, 200, 2, False, False),     # config-5 sample count: 4-row tiles
    (5000, 24, 3000, 2, False, False),    # config-4 shape: many tiles of few samples, scattered node rows
])
def test_front_paths_agree(oracle, G, n, E, K, row_major, with_nan):
    """K1 + K2 as one TMA-fed kernel (front_path 1), the same kernel fed by its plain cp.async loader (2) and the
    separate gather + gene-statistics kernels (0): identical medians / cut-offs / selections, all equal to the oracle."""
    prob = synth.random_problem(G, n, E, K, 4242 + G + n, row_major=row_major)
    if with_nan:
        prob.assay = prob.assay.copy(order="F")
        prob.assay[int(prob.from_idx[0]) - 1, 3] = np.nan
        prob.assay[int(prob.to_idx[5]) - 1, [0, 7]] = np.nan
    ref = oracle.run_omic(prob, ALL_FIELDS)
    eng = md.Engine([0])
    eng.set_option("graph", 0)
    outs = {}
    for path in (0, 1, 2):
        eng.set_option("front_path", path)
        outs[path] = eng.run_omic(prob, ALL_FIELDS)
        bits = int(eng.stats()["fit_path"])
        assert bool(bits & 0xC00) == (path != 0), (path, hex(bits))
        if path == 2:
            assert bits & 0x800
        assert_parity(outs[path], ref, prob, what=f"front_path {path}: ")
    for k in ("median", "cutoff", "status", "cat", "best", "sig_count", "tot_edges"):
        assert np.array_equal(outs[0][k], outs[1][k], equal_nan=True), k
        assert np.array_equal(outs[1][k], outs[2][k], equal_nan=True), k
    eng.close()


@pytest.mark.parametrize("padj", [_abi.PADJ_BH, _abi.PADJ_BONFERRONI, _abi.PADJ_HOLM, _abi.PADJ_HOCHBERG, _abi.PADJ_BY])
@pytest.mark.parametrize("K,E,seed", [(2, 3000, 1601), (3, 9000, 1602), (2, 40, 1603), (3, 100_000, 1604)])
def test_onepass_adjustment_equals_three_pass(padj, K, E, seed):
    """K8 in one launch (k_bh_onepass: tile aggregates read back through published words, up to 64 tiles) against the
    three-launch form: bit-identical adjusted values, counts and selection."""
    G = 120 if E < 10_000 else 700
    prob = synth.random_problem(G, 48, E, K, seed, _abi.METHOD_LM, padj)
    want = ("padj", "padj_best", "cat_best", "sig_count", "tot_edges", "best")
    eng = md.Engine([0])
    eng.set_option("graph", 0)
    res = {}
    for mode in (1, 0):
        eng.set_option("bh_onepass", mode)
        res[mode] = eng.run_omic(prob, want)
    for k in want:
        assert np.array_equal(res[0][k], res[1][k], equal_nan=True), k
    eng.close()


@pytest.mark.parametrize("chunk_mb", [2, 5])        # 7 and 2 chunks of rows
def test_streamed_input_copy_equals_plain_copy(oracle, chunk_mb):
    """Host matrices cross PCIe in row chunks while the compute phase already runs (option h2d_pipeline; the front
    kernel's TMA loader waits for the chunk of its tile, the quantile sample only draws from the first chunk): every
    output equals the copy-first run bit for bit, eagerly and replayed from the CUDA graph, and matches the oracle."""
    prob = synth.random_problem(2600, 700, 4000, 3, 7711)   # 14.6 MB column-major
    ref = oracle.run_omic(prob, ALL_FIELDS)
    eng = md.Engine([0])
    eng.set_option("h2d_pipeline", 0)
    plain = eng.run_omic(prob, ALL_FIELDS)
    assert not int(eng.stats()["fit_path"]) & 0x1000
    assert_parity(plain, ref, prob, what="copy first: ")
    eng.set_option("h2d_pipeline", chunk_mb)
    for rep in range(4):                                     # the third call captures the graph, the fourth replays it
        got = eng.run_omic(prob, ALL_FIELDS)
        bits = int(eng.stats()["fit_path"])
        assert bits & 0x1000, hex(bits)
        if rep == 3:
            assert bits & 0x100, hex(bits)
        for k in ALL_FIELDS:
            assert np.array_equal(got[k], plain[k], equal_nan=True), (k, rep)
    eng.close()


@pytest.mark.parametrize("G,n,E,K,method,chunk_mb", [
    (6000, 500, 5000, 3, _abi.METHOD_LM, 0),     # 24 MB, node rows scattered: packed + streamed in 4 MB chunks (auto)
    (6000, 500, 5000, 2, _abi.METHOD_LM, 1),     # ... 1 MB chunks
    (900, 60, 700, 2, _abi.METHOD_RLM, 0),       # small: packed, sent with one plain copy
    (2500, 301, 40000, 3, _abi.METHOD_LM, 0),    # every row a node: contiguous runs, odd sample count
])
def test_host_packed_rows_equal_plain_copy(oracle, G, n, E, K, method, chunk_mb):
    """Option host_pack: the host gathers the network-node rows into pinned staging with its worker threads and the call
    runs on the packed problem (node-rank edge list), chunks streamed to the device behind the already enqueued compute
    phase: every output equals the run on the caller's matrix bit for bit - eager, captured, replayed - and the oracle;
    mdeggs_pair_moments keeps taking rows of the caller's matrix; an edge index outside 1..G is still reported."""
    prob = synth.random_problem(G, n, E, K, 9100 + G + n, method)
    ref = oracle.run_omic(prob, ALL_FIELDS)
    eng = md.Engine([0])
    eng.set_option("host_pack", 0)
    plain = eng.run_omic(prob, ALL_FIELDS)
    assert not int(eng.stats()["fit_path"]) & 0x2000
    assert_parity(plain, ref, prob, what="plain copy: ")
    a_rows = np.array([prob.from_idx[0], prob.to_idx[1], 1, G], dtype=np.int32)
    b_rows = np.array([prob.to_idx[0], prob.from_idx[1], 2, G - 1], dtype=np.int32)
    mom_plain = eng.pair_moments(a_rows, b_rows, K)
    eng.set_option("host_pack", 2)
    if chunk_mb:
        eng.set_option("h2d_pipeline", chunk_mb)
    for rep in range(4):
        got = eng.run_omic(prob, ALL_FIELDS)
        bits = int(eng.stats()["fit_path"])
        assert bits & 0x2000, hex(bits)
        if G * n * 8 > (16 << 20):
            assert bits & 0x1000, hex(bits)              # streamed behind the compute phase
        for k in ALL_FIELDS:
            assert np.array_equal(got[k], plain[k], equal_nan=True), (k, rep)
    assert np.array_equal(eng.pair_moments(a_rows, b_rows, K), mom_plain, equal_nan=True)
    bad = synth.random_problem(G, n, E, K, 9100 + G + n, method)
    bad.from_idx = bad.from_idx.copy()
    bad.from_idx[3] = G + 1
    with pytest.raises(md.EngineError, match="outside 1..G"):
        eng.run_omic(bad, ALL_FIELDS)
    got = eng.run_omic(prob, ALL_FIELDS)                    # the context is usable afterwards
    for k in ALL_FIELDS:
        assert np.array_equal(got[k], plain[k], equal_nan=True), k
    eng.close()


def test_pageable_matrices_are_packed_by_default(oracle):
    """A pageable host matrix (numpy here, an R matrix in the drop-in) above 2 MB goes through the packing path without
    any option; a small one keeps the plain copy."""
    eng = md.Engine([0])
    big = synth.random_problem(4000, 200, 3000, 2, 5150)
    got = eng.run_omic(big, ALL_FIELDS)
    assert int(eng.stats()["fit_path"]) & 0x2000
    assert_parity(got, oracle.run_omic(big, ALL_FIELDS), big, what="packed by default: ")
    small = synth.random_problem(300, 40, 200, 2, 5151)
    got = eng.run_omic(small, ALL_FIELDS)
    assert not int(eng.stats()["fit_path"]) & 0x2000
    assert_parity(got, oracle.run_omic(small, ALL_FIELDS), small, what="small: ")
    eng.close()


@pytest.mark.parametrize("df", [1.0, 2.0, 6.0, 20.0, 96.0, 996.0, 1996.0, 19996.0, 4e5])
def test_tail_table_against_series_and_mpmath(engine, df):
    """Option tail_spline: the per-run interpolation table of log P(|T| > t) in s = sqrt(log1p(t^2 / df)) against the
    per-edge series / continued fractions it is built from (dense grid of t, incl. interval boundaries and the far tail,
    where the table ends and the fraction takes over) and against 50-digit mpmath (core_functions.R:1072)."""
    import oracle_np
    rng = np.random.default_rng(int(df))
    t = np.concatenate([np.linspace(0.0, 6.0, 6001), np.geomspace(1e-9, 1e3, 4000), rng.uniform(0, 60, 4000),
                        np.sqrt(df) * np.sqrt(np.expm1((np.arange(1, 2049) * (np.sqrt(min(1360.0 / df, 680.0)) / 2048)) ** 2)),
                        [0.0, 1e-300, 1e140, 1e160, np.inf, -3.3]])
    t = t[np.isfinite(t) | np.isinf(t)]
    tab = engine.dev_p_two_sided_tab(t, df, True)
    ser = engine.dev_p_two_sided_tab(t, df, False)
    # (the per-edge series gives 1 - w for |t| ~ 3.5-4.5 with up to ~1e-10 relative noise from the cancellation; the
    # nodes of the table take the fraction there; both are inside the 1e-9 parity bar) + one step of the 2.2e-16 grid
    ok = np.abs(tab - ser) <= 5e-10 * ser + 2.3e-16
    assert ok.all(), (df, t[~ok][:5], tab[~ok][:5], ser[~ok][:5])
    assert np.all((tab / 2.220446049250313e-16) == np.round(tab / 2.220446049250313e-16)), "2.2e-16 grid"
    ts = np.concatenate([[1e-6, 0.3, 1.0, 1.8, 2.5, 5.5, 9.9, 12.2, 28.8, 40.0], np.linspace(3.3, 4.7, 29)])
    true2 = np.array([oracle_np.t_two_sided_mp(float(x), df) for x in ts])
    got = engine.dev_p_two_sided_tab(ts, df, True)
    rtol = 5e-12 if df <= 2000 else 2e-10     # (long fractions at df ~ 1e5 accumulate ~1e-11)
    assert np.all(np.abs(got - true2) <= rtol * true2 + 2.3e-16), (df, np.max(np.abs(got - true2) / true2))


@pytest.mark.parametrize("E", [40_000, 300])
def test_tail_table_inside_the_pipeline(oracle, E):
    """Two-category lm with the tail table on (forced for the small job) and off: p-values within one grid step of each
    other, statuses / selections identical, both within the parity bar of the oracle."""
    prob = synth.random_problem(500, 64, E, 2, 6060 + E, _abi.METHOD_LM, _abi.PADJ_BH)
    ref = oracle.run_omic(prob, ALL_FIELDS)
    eng = md.Engine([0])
    outs = {}
    for mode in (0, 2):
        eng.set_option("tail_spline", mode)
        outs[mode] = eng.run_omic(prob, ALL_FIELDS)
        assert_parity(outs[mode], ref, prob, what=f"tail_spline {mode}: ")
    a, b = outs[0]["p_edge"], outs[2]["p_edge"]
    assert np.all((np.isnan(a) & np.isnan(b)) | (np.abs(a - b) <= 5e-10 * np.abs(a) + 2.3e-16))
    for k in ("status", "cat", "best", "sig_count", "tot_edges", "cat_best"):
        assert np.array_equal(outs[0][k], outs[2][k]), k
    eng.close()


def test_streamed_packing_survives_blocking_launches():
    """With CUDA_LAUNCH_BLOCKING=1 (or under a profiler that serialises launches) a launch only returns when the kernel
    has run, so the engine must not launch the compute phase before the chunks it waits for are enqueued: it probes the
    launch behaviour once per context (k_async_probe) and packs first.  Same results, no time-out."""
    import os, subprocess, sys, textwrap
    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = textwrap.dedent("""
        import sys, numpy as np
        sys.path.insert(0, %r); sys.path.insert(0, %r)
        import multideggs_b200 as md
        from multideggs_b200 import synth
        prob = synth.random_problem(6000, 500, 5000, 2, 424242)
        want = ("p_edge", "status", "cutoff", "median", "best", "sig_count", "padj_best")
        eng = md.Engine([0])
        eng.set_option("host_pack", 0); eng.set_option("h2d_pipeline", 0)
        plain = eng.run_omic(prob, want)
        eng.set_option("host_pack", 2); eng.set_option("h2d_pipeline", 1)
        for rep in range(4):
            got = eng.run_omic(prob, want)
            assert int(eng.stats()["fit_path"]) & 0x3000 == 0x3000
            for k in want:
                assert np.array_equal(got[k], plain[k], equal_nan=True), (k, rep)
        print("ok")
    """) % (ROOT, os.path.join(ROOT, "tests"))
    env = dict(os.environ, CUDA_LAUNCH_BLOCKING="1")
    res = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0 and "ok" in res.stdout, res.stdout + res.stderr

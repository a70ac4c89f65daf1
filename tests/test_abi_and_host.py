"""CPU tests: the C-ABI library loads and exports every declared symbol (no compute without a GPU), and the
host-side mirror of the R API behaves like the reference (names, factor handling, errors, layout)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pandas as pd
import pytest

import multideggs_b200 as md
from multideggs_b200 import _abi, engine as eng_mod, host
from helpers import oracle_single_omic

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_header_symbol():
    import __graft_entry__ as g
    g.build()
    lib = eng_mod.load_library()
    header = open(os.path.join(ROOT, "include", "mdeggs.h")).read()
    declared = set(re.findall(r"\b(mdeggs_[a-z0-9_]+)\s*\(", header))
    assert declared == set(eng_mod.ABI_SYMBOLS)
    for nm in declared:
        assert hasattr(lib, nm), nm
    assert lib.mdeggs_abi_version() == _abi.ABI_VERSION
    assert lib.mdeggs_strerror(-1) == b"invalid argument"
    out = subprocess.run(["nm", "-D", "--defined-only", eng_mod.LIB_PATH], capture_output=True, text=True).stdout
    for nm in declared:
        assert f" T {nm}" in out


@pytest.mark.parametrize("G,n,E,chunk_rows,contiguous", [
    (500, 37, 300, 0, False),       # scattered node rows, one chunk, a last column block of 5
    (3000, 64, 2500, 160, False),   # several row chunks (the unit order the streamed copy uses)
    (400, 20, 4000, 48, True),      # every row is a node: contiguous runs are copied with memcpy
    (64, 3, 1, 0, False),           # a single edge
])
def test_host_packing_of_node_rows(G, n, E, chunk_rows, contiguous):
    """The host-side packing step of option "host_pack" (worker pool, no GPU): node rows of the edge list, ascending,
    every value in place - against numpy (core_functions.R:333-336: assayData[rownames %in% nodes, ])."""
    rng = np.random.default_rng(G + n)
    a = np.asfortranarray(rng.normal(size=(G, n)))
    a[rng.integers(0, G, 5), rng.integers(0, n, 5)] = np.nan
    if contiguous:
        f = np.arange(1, G + 1, dtype=np.int32).repeat(E // G + 1)[:E]
        t = np.roll(f, 1)
    else:
        f = rng.integers(1, G + 1, E).astype(np.int32)
        t = rng.integers(1, G + 1, E).astype(np.int32)
    for rep in range(3):                                    # the pool is reused from job to job
        packed, rows = eng_mod.host_pack_rows(a, f, t, chunk_rows)
        want_rows = np.unique(np.concatenate([f, t])) - 1
        assert np.array_equal(rows, want_rows)
        assert packed.shape == (want_rows.size, n)
        assert np.array_equal(packed, a[want_rows, :], equal_nan=True)
    # indices outside 1..G are no nodes (the device check reports them)
    f2 = f.copy()
    f2[0] = G + 7
    t2 = t.copy()
    t2[-1] = 0
    packed, rows = eng_mod.host_pack_rows(a, f2, t2, chunk_rows)
    want_rows = np.unique(np.concatenate([f2[1:], t2[:-1], t2[:1] if E > 1 else t2[:0], f2[-1:] if E > 1 else f2[:0]])) - 1
    want_rows = want_rows[(want_rows >= 0) & (want_rows < G)]
    assert np.array_equal(rows, want_rows)
    assert np.array_equal(packed, a[want_rows, :], equal_nan=True)


def test_struct_layout_matches_header():
    """sizeof/offsetof cross-check of the ctypes mirrors against the C compiler."""
    src = r'''
#include <stdio.h>
#include <stddef.h>
#include "mdeggs.h"
int main(void){
  printf("%zu %zu %zu %zu %zu %zu %zu\n", sizeof(mdeggs_problem), offsetof(mdeggs_problem, from), offsetof(mdeggs_problem, group),
         offsetof(mdeggs_problem, pct), offsetof(mdeggs_problem, method), sizeof(mdeggs_result), sizeof(mdeggs_stats));
  printf("%zu %zu\n", offsetof(mdeggs_result, fail_count), offsetof(mdeggs_stats, fit_bytes));
  return 0; }'''
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        c = os.path.join(td, "l.c")
        open(c, "w").write(src)
        exe = os.path.join(td, "l")
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe], check=True)
        a, b = subprocess.run([exe], capture_output=True, text=True).stdout.strip().split("\n")
    vals = [int(v) for v in a.split()]
    P, R, S = _abi.CProblem, _abi.CResult, _abi.CStats
    assert vals == [C.sizeof(P), P.from_.offset, P.group.offset, P.pct.offset, P.method.offset, C.sizeof(R), C.sizeof(S)]
    assert [int(v) for v in b.split()] == [R.fail_count.offset, S.fit_bytes.offset]


def test_no_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(md.EngineError):
        md.Engine([0])


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "multideggs_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".c", ".R")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "pyoracle" not in txt and "oracle_np" not in txt and "mdeggs_oracle" not in txt, f


# ---- tidy_metadata (core_functions.R:564-686) ------------------------------------------------------------
def test_tidy_metadata_variants(bundled):
    meta = bundled["metadata"]
    f = host.tidy_metadata(metadata=meta, category_variable="response")
    assert f.levels == ["Non_responder", "Responder"] and f.table() == {"Non_responder": 58, "Responder": 42}
    assert list(f.names[:3]) == ["PT001", "PT002", "PT003"]
    # character vector -> factor with sorted levels
    f2 = host.tidy_metadata(metadata=pd.Series(meta["response"].astype(str).to_numpy(), index=meta.index))
    assert f2.levels == f.levels and np.array_equal(f2.codes, f.codes)
    # numeric named vector: as.factor(c(1,2)) -> "1","2"
    f3 = host.tidy_metadata(metadata=pd.Series(meta["response"].cat.codes.to_numpy() + 1.0, index=meta.index))
    assert f3.levels == ["1", "2"]
    # category_subset + small-category removal + NA removal
    g = meta["gender"].astype(object).copy()
    g.iloc[:3] = "Other"
    g.iloc[10] = None
    msgs = []
    host._message, old = msgs.append, host._message
    try:
        f4 = host.tidy_metadata(metadata=pd.Series(g.to_numpy(), index=meta.index), verbose=True)
    finally:
        host._message = old
    assert f4.levels == ["Female", "Male"] and len(f4) == 100 - 3 - 1
    assert any("Other category did not contain enough samples" in m for m in msgs)
    # (R drops the NA sample silently here: `NA %in% names(tbl)[tbl >= 5]` is FALSE, core_functions.R:651-653)
    g2 = meta["gender"].astype(object).copy()
    g2.iloc[10] = None
    msgs.clear()
    host._message, old = msgs.append, host._message
    try:
        f5 = host.tidy_metadata(metadata=pd.Series(g2.to_numpy(), index=meta.index), verbose=True)
    finally:
        host._message = old
    assert len(f5) == 99 and any("contained NAs. 1 samples were removed" in m for m in msgs)
    with pytest.raises(ValueError, match="category_subset values must be contained"):
        host.tidy_metadata(category_subset=["nope"], metadata=meta, category_variable="response")
    with pytest.raises(ValueError, match="category_variable must be specified"):
        host.tidy_metadata(metadata=meta)
    sub = host.tidy_metadata(category_subset=["Responder", "Non_responder"], metadata=meta, category_variable="response")
    assert len(sub) == 100


def test_prepare_omic_edges_nodes_and_errors(bundled):
    fac = host.tidy_metadata(metadata=bundled["metadata"], category_variable="response")
    prep = host.prepare_omic(bundled["RNAseq"], "RNAseq", fac, "lm", None, md.DEFAULT_PERCENTILES, "bonferroni")
    genes = list(bundled["RNAseq"].index)
    pairs = [(genes[a - 1], genes[b - 1]) for a, b in zip(prep.problem.from_idx, prep.problem.to_idx)]
    assert len(pairs) == 8 and ("TNF", "TNFRSF1A") in pairs and ("TNF", "FAS") in pairs
    assert prep.problem.K == 2 and prep.categories == ["Non_responder", "Responder"]
    # duplicated edges are dropped keeping the first (core_functions.R:329-331); extra columns are ignored
    net = pd.DataFrame({"from": ["TNF", "TNF", "XXX", "IL1B"], "to": ["FAS", "FAS", "TNF", "IL1R2"], "w": [1, 2, 3, 4]})
    p2 = host.prepare_omic(bundled["RNAseq"], "RNAseq", fac, "lm", net, md.DEFAULT_PERCENTILES, "BH")
    assert p2.problem.dims[2] == 2
    with pytest.raises(ValueError, match="don't match any link"):
        host.prepare_omic(bundled["RNAseq"], "RNAseq", fac, "lm", pd.DataFrame({"from": ["A"], "to": ["B"]}),
                          md.DEFAULT_PERCENTILES, "BH")
    with pytest.raises(ValueError, match="don't match the IDs in metadata"):
        bad = bundled["RNAseq"].copy()
        bad.columns = [f"X{i}" for i in range(100)]
        host.prepare_omic(bad, "RNAseq", fac, "lm", None, md.DEFAULT_PERCENTILES, "BH")
    # samples missing from the assay are reported; only one category left -> error
    msgs = []
    host._message, old = msgs.append, host._message
    try:
        part = bundled["RNAseq"].iloc[:, :60]
        p3 = host.prepare_omic(part, "RNAseq", fac, "lm", None, md.DEFAULT_PERCENTILES, "BH")
        assert p3.problem.assay.shape == (14, 60) and any("missing in RNAseq" in m for m in msgs)
        only = bundled["RNAseq"].loc[:, (bundled["metadata"]["response"] == "Responder").to_numpy()]
        with pytest.raises(ValueError, match="belong to one category"):
            host.prepare_omic(only, "RNAseq", fac, "lm", None, md.DEFAULT_PERCENTILES, "BH")
    finally:
        host._message = old
    with pytest.raises(ValueError, match="padj_method"):
        host.prepare_omic(bundled["RNAseq"], "RNAseq", fac, "lm", None, md.DEFAULT_PERCENTILES, "bogus")


def test_get_diffNetworks_argument_checks(bundled):
    with pytest.raises(ValueError, match="should be one of"):
        md.get_diffNetworks(bundled["RNAseq"], bundled["metadata"], "response", regression_method="glm")
    with pytest.raises(ValueError, match="category_variable must be %in% colnames"):
        md.get_diffNetworks(bundled["RNAseq"], bundled["metadata"], "nope")
    with pytest.raises(ValueError, match="category_variable is NULL"):
        md.get_diffNetworks(bundled["RNAseq"], bundled["metadata"])
    with pytest.raises(ValueError, match="not a list or a matrix"):
        md.get_diffNetworks(3.0, bundled["metadata"], "response")
    with pytest.raises(NotImplementedError):
        md.get_diffNetworks(bundled["RNAseq"], bundled["metadata"], "response", mixedModel=True, id_variable="patient_id")


def test_consumers_on_oracle_built_deggs(bundled):
    """get_sig_deggs / get_multiOmics_diffNetworks layout (core_functions.R:1101-1230) on a deggs object whose
    networks were computed by the oracle through the same host assembly code."""
    fac = host.tidy_metadata(metadata=bundled["metadata"], category_variable="response")
    nets = {k: oracle_single_omic(bundled[k], k, fac, "lm", None, md.DEFAULT_PERCENTILES, "bonferroni")
            for k in ("RNAseq", "Proteomics", "Olink")}
    deggs = md.Deggs(diffNetworks=nets, assayData={k: bundled[k] for k in nets}, metadata=fac, regression_method="lm",
                     mixedModel=False, id_variable=None, lmer_ctrl=None, category_subset=None, padj_method="bonferroni")
    sig = md.get_sig_deggs(deggs)                       # tests/testthat/test-core_functions.R:110-125
    assert sig.shape == (7, 4) and list(sig.columns) == ["from", "to", "p.value", "p.adj"]
    assert md.get_sig_deggs(deggs, "Proteomics").shape == (5, 4)
    assert md.get_sig_deggs(deggs, 3).shape[0] == 0     # Olink: nothing significant
    multi = md.get_multiOmics_diffNetworks(deggs)
    assert set(multi) == {"Non_responder", "Responder"}
    assert set(multi["Non_responder"]["layer"].unique()) == {"RNAseq", "Proteomics"}
    one = md.Deggs(deggs, assayData={"RNAseq": bundled["RNAseq"]})
    with pytest.raises(ValueError, match="Only one omic dataset"):
        md.get_multiOmics_diffNetworks(one)
    none_d = md.Deggs(deggs, padj_method="none")
    assert "p.value" in md.get_sig_deggs(none_d).columns


def test_failing_pair_semantics_follow_cores(bundled, oracle):
    """quirk q7: cores == 1 -> a singular tested pair aborts the omic; cores > 1 -> only that percentile is lost."""
    fac = host.tidy_metadata(metadata=bundled["metadata"], category_variable="response")
    x = bundled["RNAseq"].copy()
    x.loc["TNF", :] = 9.0          # constant predictor, expressed above every low cut-off in both categories
    x.loc["TNFRSF1A", (bundled["metadata"]["response"] == "Responder").to_numpy()] = 0.0   # specific to Non_responder
    with pytest.raises(ArithmeticError):
        oracle_single_omic(x, "RNAseq", fac, "lm", None, md.DEFAULT_PERCENTILES, "bonferroni", cores=1)
    out = oracle_single_omic(x, "RNAseq", fac, "lm", None, md.DEFAULT_PERCENTILES, "bonferroni", cores=2)
    nr = out["Non_responder"]
    assert isinstance(nr, str) or "TNF-TNFRSF1A" not in nr.index


def test_one_bad_layer_does_not_stop_the_others(bundled, oracle, monkeypatch):
    """core_functions.R:219-233: the per-omic tryCatch turns ANY error of a layer into message() + NULL.  An engine-side
    rejection of one layer (OmicError) must behave the same in the batched multi-omic path and the per-layer path."""
    import pyoracle
    from multideggs_b200 import api
    from multideggs_b200.engine import OmicError

    class FakePool:  # the oracle as the compute of the good layers; the second layer is rejected by "the engine"
        def run_omics(self, problems, wants):
            return [OmicError("mdeggs_run_omics[1]: unsupported configuration: test") if i == 1
                    else pyoracle.run_omic(p, w) for i, (p, w) in enumerate(zip(problems, wants))]

    class FakeEngine:
        calls = 0

        def run_omic(self, problem, want):
            FakeEngine.calls += 1
            if FakeEngine.calls == 2:
                raise OmicError("mdeggs_run_omic: invalid argument: test")
            return pyoracle.run_omic(problem, want)

    assays = {k: bundled[k] for k in ("RNAseq", "Proteomics", "Olink")}
    msgs = []
    monkeypatch.setattr(host, "_message", msgs.append)
    monkeypatch.setattr(api, "default_pool", lambda: FakePool())
    for kw in ({}, {"engine": FakeEngine()}):
        msgs.clear()
        deggs = md.get_diffNetworks(assayData=assays, metadata=bundled["metadata"], category_variable="response",
                                    regression_method="lm", padj_method="bonferroni", verbose=False,
                                    show_progressBar=False, cores=1, **kw)
        nets = deggs["diffNetworks"]
        assert list(nets) == ["RNAseq", "Proteomics", "Olink"]
        assert nets["Proteomics"] is None and any("test" in m for m in msgs)
        assert nets["RNAseq"]["sig_pvalues_count"] == 7 and nets["Olink"]["sig_pvalues_count"] == 0


def test_host_adjust_methods(oracle):
    rng = np.random.default_rng(3)
    p = rng.uniform(size=50) ** 2
    for m in ("holm", "hochberg", "hommel", "BY", "BH", "bonferroni"):
        q = host.p_adjust(p, m)
        assert np.all(q >= p - 1e-15) and np.all(q <= 1.0)
    # hommel is between hochberg and bonferroni-holm in power
    assert np.all(host.p_adjust(p, "hommel") <= host.p_adjust(p, "hochberg") + 1e-15)
    # the C restatement of R's hommel loop (oracle, method 6) and the numpy one agree bit for bit, with ties and NA
    for n in (1, 2, 3, 4, 9, 64, 257):
        v = rng.uniform(size=n) ** 3
        if n > 3:
            v[2] = v[0]
            v[1] = np.nan
        assert np.array_equal(oracle.p_adjust(v, 6), host.p_adjust(v, "hommel"), equal_nan=True), n
    # q.value: max(p) < 0.95 -> qvalue() errors -> pi0 = 1 retry == BH up to the multiply/divide order
    small = p[p < 0.9]
    assert np.allclose(host.qvalue_or_fallback(small), host.p_adjust(small, "BH"), rtol=1e-14)
    assert np.isnan(host.qvalue_or_fallback(np.array([0.2]))).all()       # 1-row network -> NA
    full = np.concatenate([p, [0.97, 0.99]])
    q = host.qvalue_or_fallback(full)
    assert np.all(q <= host.p_adjust(full, "BH") + 1e-12)                  # pi0 <= 1 shrinks BH


def test_rdata_reader_against_reference_if_present():
    ref = "/root/reference/data/synthetic_rnaseqData.rda"
    if not os.path.exists(ref):
        pytest.skip("reference not mounted (GPU box)")
    from multideggs_b200 import rdata
    m, rn, cn = rdata.as_matrix(rdata.read_rda(ref)["synthetic_rnaseqData"])
    d = np.load(os.path.join(ROOT, "tests", "golden", "bundled.npz"))
    assert np.array_equal(m, d["rnaseq_x"]) and rn == list(d["rnaseq_genes"]) and cn == list(d["rnaseq_samples"])
    net = rdata.as_dataframe(rdata.read_rda("/root/reference/R/sysdata.rda")["omic_network"])
    ours = md.omic_network()
    assert list(ours["from"]) == list(net["from"]) and list(ours["to"]) == list(net["to"]) and len(ours) == 53035

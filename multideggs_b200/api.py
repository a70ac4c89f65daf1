"""Python mirror of the reference's exported R API for the differential-network hot path.

Same function names, argument names/meaning, return layout and error behaviour as
  get_diffNetworks               R/core_functions.R:128-250
  get_diffNetworks_singleOmic    R/core_functions.R:258-529
  get_sig_deggs                  R/core_functions.R:1101-1149
  get_multiOmics_diffNetworks    R/core_functions.R:1175-1230
  multiDEGGs_filter              R/feature_selection_for_ML.R:77-166
  predict.multiDEGGs_filter      R/feature_selection_for_ML.R:406-463
R is not installed in this image, so this module plays the role of the R host in tests and benchmarks;
the R host itself (same flow, `.Call` into the same C ABI) is shipped under multideggs_b200/r/.
All arithmetic of the path runs in the CUDA engine (engine.Engine.run_omic); there is no CPU fallback.
"""
from __future__ import annotations

from typing import Any, Sequence

import numpy as np
import pandas as pd

from . import _abi, host
from .engine import Engine, OmicError, default_engine, default_pool
from .host import NO_LINKS, Assay, Factor, IndexedNetwork

DEFAULT_PERCENTILES = _abi.r_seq(0.35, 0.98, 0.05)   # core_functions.R:137
FILTER_PERCENTILES = _abi.r_seq(0.25, 0.98, 0.05)    # feature_selection_for_ML.R:86


class Deggs(dict):
    """The S3 class "deggs" (core_functions.R:237-248): a 9-field list."""

    FIELDS = ("diffNetworks", "assayData", "metadata", "regression_method", "mixedModel", "id_variable",
              "lmer_ctrl", "category_subset", "padj_method")

    def __getattr__(self, k: str) -> Any:
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


def get_diffNetworks_singleOmic(assayData, assayDataName, metadata: Factor, regression_method, mixedModel=False,
                                id_variable=None, lmer_ctrl=None, network=None,
                                percentile_vector=DEFAULT_PERCENTILES, padj_method="bonferroni",
                                show_progressBar=False, verbose=False, cores=1, *, engine: Engine | None = None,
                                tested_coef: int = 3, rlm_cov_mode: int = _abi.RLM_COV_XTWX) -> dict:
    """One omic layer: host prep -> ONE engine call -> data.frame assembly (core_functions.R:258-529).

    `cores` and `show_progressBar` are accepted for signature compatibility; `cores` only selects the
    reference's error semantics for failing pairs (quirk q7), the GPU path is not forked."""
    if verbose:
        host.message(f"Generating {assayDataName} network layer...")
    if mixedModel:
        raise NotImplementedError("mixed models are outside the GPU hot path (SURVEY §2 #7)")
    prep = host.prepare_omic(assayData, assayDataName, metadata, regression_method, network, percentile_vector,
                             padj_method, tested_coef, rlm_cov_mode)
    eng = engine if engine is not None else default_engine()
    res = eng.run_omic(prep.problem, host.want_fields(prep))
    return host.assemble_networks(prep, res, cores=cores, verbose=verbose)


def get_diffNetworks(assayData, metadata, category_variable=None, regression_method="lm", mixedModel=False,
                     id_variable=None, lmer_ctrl=None, category_subset=None, network=None,
                     percentile_vector=DEFAULT_PERCENTILES, padj_method="bonferroni", show_progressBar=True,
                     verbose=True, cores=1, *, engine: Engine | None = None, tested_coef: int = 3,
                     rlm_cov_mode: int = _abi.RLM_COV_XTWX) -> Deggs:
    """core_functions.R:128-250."""
    if regression_method not in ("lm", "rlm"):
        raise ValueError("'arg' should be one of 'lm', 'rlm'")
    if isinstance(assayData, (pd.DataFrame, Assay)):
        assay_list = {"assayData1": assayData}
    elif isinstance(assayData, dict):
        assay_list = dict(assayData)
    elif isinstance(assayData, (list, tuple)):
        assay_list = {f"assayData{i + 1}": a for i, a in enumerate(assayData)}
    else:
        raise ValueError("assayDataList is not a list or a matrix or data.frame")
    if category_variable is not None:
        if isinstance(metadata, pd.DataFrame) and category_variable not in metadata.columns:
            raise ValueError("category_variable must be %in% colnames(metadata). If metadata is a \n"
                             "       named vector category_variable must be NULL.")
    elif isinstance(metadata, pd.DataFrame):
        raise ValueError("metadata is a matrix or dataframe but category_variable is NULL.\n"
                         "     Please specify category_variable to select the correct column \n     from metadata.")
    if mixedModel:
        raise NotImplementedError("mixed models are outside the GPU hot path (SURVEY §2 #7)")
    meta = host.tidy_metadata(category_subset=category_subset, metadata=metadata,
                              category_variable=category_variable, id_variable=None, verbose=verbose)
    nets: dict[str, Any] = {}
    if engine is None and len(assay_list) > 1:
        # every layer prepared on the host, then ONE launch set for all of them (mdeggs_run_omics), then assembly;
        # each step keeps the reference's per-omic tryCatch (core_functions.R:219-233)
        preps: dict[str, host.OmicPrep] = {}
        for name, a in assay_list.items():
            try:
                if verbose:
                    host.message(f"Generating {name} network layer...")
                preps[name] = host.prepare_omic(a, name, meta, regression_method, network, percentile_vector,
                                                padj_method, tested_coef, rlm_cov_mode)
            except (ValueError, ArithmeticError) as e:
                host.message(str(e))
                nets[name] = None
        results = default_pool().run_omics([p.problem for p in preps.values()],
                                           [host.want_fields(p) for p in preps.values()])
        for (name, prep), res in zip(preps.items(), results):
            if isinstance(res, OmicError):  # the engine rejected this layer: same tryCatch as any other per-omic error
                host.message(str(res))
                nets[name] = None
                continue
            if isinstance(res, Exception):  # CUDA / NCCL failure: never swallowed
                raise res
            try:
                nets[name] = host.assemble_networks(prep, res, cores=cores, verbose=verbose)
            except (ValueError, ArithmeticError) as e:
                host.message(str(e))
                nets[name] = None
        nets = {name: nets[name] for name in assay_list}  # layer order of the input list
    else:
        for name, a in assay_list.items():
            try:  # tryCatch(..., error = function(e) message(conditionMessage(e)))  core_functions.R:219-233
                nets[name] = get_diffNetworks_singleOmic(a, name, meta, regression_method, mixedModel, id_variable,
                                                         lmer_ctrl, network, percentile_vector, padj_method,
                                                         show_progressBar, verbose, cores, engine=engine,
                                                         tested_coef=tested_coef, rlm_cov_mode=rlm_cov_mode)
            except (ValueError, ArithmeticError, OmicError) as e:
                host.message(str(e))
                nets[name] = None
    return Deggs(diffNetworks=nets, assayData=assay_list, metadata=meta, regression_method=regression_method,
                 mixedModel=mixedModel, id_variable=id_variable, lmer_ctrl=lmer_ctrl,
                 category_subset=category_subset, padj_method=padj_method)


def _omic_entry(deggs_object: Deggs, assayDataName):
    nets = deggs_object["diffNetworks"]
    if isinstance(assayDataName, int):
        return list(nets.values())[assayDataName - 1]  # R is 1-based
    return nets[assayDataName]


def get_sig_deggs(deggs_object: Deggs, assayDataName=1, sig_threshold=0.05) -> pd.DataFrame:
    """core_functions.R:1101-1149."""
    if not isinstance(deggs_object, Deggs):
        raise ValueError("Input must be a 'deggs' object")
    sig_var = "p.value" if deggs_object["padj_method"] == "none" else "p.adj"
    entry = _omic_entry(deggs_object, assayDataName) or {}
    frames = {k: v for k, v in entry.items() if isinstance(v, pd.DataFrame)}
    if not frames:
        host.message("Warning: No valid network found")
        return pd.DataFrame()
    parts = []
    for _, sub in frames.items():
        if sub.shape[0] == 0:
            continue
        rows = (sub[sig_var] < sig_threshold).to_numpy()
        if rows.any():
            parts.append(sub[rows])
    if not parts:
        host.message(f"Warning: No significant interactions found with {sig_var} < {sig_threshold}")
        return pd.DataFrame()
    return pd.concat(parts, axis=0)


def get_multiOmics_diffNetworks(deggs_object: Deggs, sig_threshold=0.05) -> dict[str, pd.DataFrame]:
    """core_functions.R:1175-1230."""
    if not isinstance(deggs_object, Deggs):
        raise ValueError("Input must be a 'deggs' object")
    if len(deggs_object["assayData"]) == 1:
        raise ValueError("Only one omic dataset was provided. This function is used only in \n"
                         "         multi-omic scenarios. Use get_sig_deggs() instead.")
    categories = deggs_object["metadata"].levels
    sig_var = "p.value" if deggs_object["padj_method"] == "none" else "p.adj"
    out: dict[str, pd.DataFrame] = {}
    for category in categories:
        parts = []
        for omic in deggs_object["assayData"].keys():
            entry = deggs_object["diffNetworks"].get(omic)
            if not isinstance(entry, dict):
                parts.append(pd.DataFrame())
                continue
            net = entry.get(category)
            if not isinstance(net, pd.DataFrame):
                parts.append(pd.DataFrame())
                continue
            net = net.copy()
            net["layer"] = omic
            parts.append(net[(net[sig_var] < sig_threshold).to_numpy()])
        nonempty = [p for p in parts if p.shape[1] > 0]
        if nonempty:
            merged = pd.concat(nonempty, axis=0)
            merged["layer"] = pd.Categorical(merged["layer"])
        else:
            merged = pd.DataFrame()
        out[category] = merged
    return out


# ----------------------------------------------------------------------------------------------
# nestedcv glue (feature_selection_for_ML.R)
# ----------------------------------------------------------------------------------------------
def plot_regressions_data(deggs_object: Deggs, assayDataName=1, gene_A: str | None = None, gene_B: str | None = None, *,
                          engine: Engine | None = None, n_points: int = 100, level: float = 0.95) -> dict:
    """The numbers behind plot_regressions (plotting_functions.R:60-240) without the drawing: the pair's interaction
    p-value, what the figure prints (`Padj` looked up in the object, plotting_functions.R:175-190), and for each category
    the line lm(y ~ x) with its confidence band at `n_points` abscissae (predict(interval = "confidence"), 125-172).

    The pair goes through the engine as a one-edge network; the per-category lines are then rebuilt from the moments that
    run left on the device (`mdeggs_pair_moments`: n_k, means, Sxx, Syy, Sxy), no second pass over the rows on the host.
    regression_method "rlm" draws MASS::rlm(y ~ x) per category in R, a robust fit with its own scale per category; that
    drawing glue is not mirrored (NotImplementedError), the interaction p-value of the rlm branch is."""
    from scipy import stats as _st
    if not isinstance(deggs_object, Deggs):
        raise ValueError("deggs_object must be of class deggs")
    assay_list = deggs_object["assayData"]
    name = list(assay_list.keys())[assayDataName - 1] if isinstance(assayDataName, int) else assayDataName
    assay = host.Assay.coerce(assay_list[name])
    genes = set(assay.genes.tolist())
    if gene_A not in genes:
        raise ValueError(f"gene_A is not in rownames({name})")
    if gene_B not in genes:
        raise ValueError(f"gene_B is not in rownames({name})")
    metadata: Factor = deggs_object["metadata"]
    method = deggs_object["regression_method"]
    net = pd.DataFrame({"from": [gene_A], "to": [gene_B]})
    prep = host.prepare_omic(assay, name, metadata, method, net, [0.0], "none")
    eng = engine if engine is not None else default_engine()
    res = eng.run_omic(prep.problem, ("p_edge", "stat", "status", "df"))
    K = prep.problem.K
    stat, df = float(res["stat"][0]), res["df"]
    if K == 2 and method == "lm":
        # coef(summary(lmfit))[4, 4]: the plain two-sided t tail 2 pt(-|t|), not the reference's 2 (1 - pt(|t|))
        p_interaction = float(2.0 * eng.dev_pt(-abs(stat), df[1])[0]) if np.isfinite(stat) else float("nan")
    else:
        p_interaction = float(res["p_edge"][0])
    padj_method = deggs_object["padj_method"]
    sig_interaction = p_interaction
    if padj_method != "none":
        entry = _omic_entry(deggs_object, name) or {}
        frames = [v for k, v in entry.items() if isinstance(v, pd.DataFrame) and k in prep.categories and v.shape[0]]
        sig_interaction = float("nan")
        if frames:
            allf = pd.concat(frames, axis=0)
            hit = allf[(allf["from"] == gene_A) & (allf["to"] == gene_B)]
            if hit.shape[0] == 0:
                hit = allf[(allf["from"] == gene_B) & (allf["to"] == gene_A)]
            if hit.shape[0] > 0:
                sig_interaction = float(hit["p.adj"].iloc[0])
    if method != "lm":
        raise NotImplementedError("per-category MASS::rlm(y ~ x) lines are drawing glue outside the hot path; "
                                  f"interaction p = {p_interaction!r}")
    mom = eng.pair_moments(prep.problem.from_idx, prep.problem.to_idx, K)[0]          # [K, 6]
    x = prep.problem.assay[int(prep.problem.from_idx[0]) - 1]
    lo, hi = float(np.nanmin(x)), float(np.nanmax(x))
    adj = (hi - lo) * 0.05
    new_x = np.linspace(lo - adj, hi + adj, n_points)
    lines = {}
    for k, cat in enumerate(prep.categories):
        n_k, ma, mb, sxx, syy, sxy = (float(v) for v in mom[k])
        if n_k < 1:
            lines[cat] = None                                                        # no sample of this category here
            continue
        slope = sxy / sxx if sxx > 0 else float("nan")
        icpt = mb - slope * ma
        rdf = n_k - 2
        sigma2 = (syy - slope * sxy) / rdf if rdf > 0 else float("nan")
        fit = icpt + slope * new_x
        se = np.sqrt(sigma2 * (1.0 / n_k + (new_x - ma) ** 2 / sxx)) if sxx > 0 else np.full_like(new_x, np.nan)
        tq = float(_st.t.ppf(0.5 + level / 2.0, rdf)) if rdf > 0 else float("nan")
        lines[cat] = {"intercept": icpt, "slope": slope, "sigma": float(np.sqrt(sigma2)) if sigma2 >= 0 else float("nan"),
                      "n": int(n_k), "fit": fit, "lwr": fit - tq * se, "upr": fit + tq * se}
    return {"assayDataName": name, "gene_A": gene_A, "gene_B": gene_B, "categories": list(prep.categories),
            "p_interaction": p_interaction, "sig_interaction": sig_interaction,
            "prefix": "P" if padj_method == "none" else "Padj", "new_x": new_x, "lines": lines}


class MultiDEGGsFilter(dict):
    """S3 class "multiDEGGs_filter": list(keep = <names>, pairs = <data.frame>)."""


def multiDEGGs_filter(y, x, keep_single_genes: bool = False, nfilter: int | None = None, *,
                      engine: Engine | None = None, ttest_filter=None) -> MultiDEGGsFilter:
    """feature_selection_for_ML.R:77-166.  `x` is samples x features (DataFrame); `y` the outcome vector.

    `ttest_filter(y, x, nfilter)` is nestedcv's fallback (fs:157-162); it is not part of this hot path, so the
    caller may inject one — otherwise an empty `keep` is returned in that case."""
    if not isinstance(x, pd.DataFrame):
        raise TypeError("x must be a pandas.DataFrame (samples x features)")
    # counts <- t(x): x is column-major samples x genes, i.e. each gene is contiguous => ROWMAJOR genes x samples
    vals = np.ascontiguousarray(x.to_numpy(dtype=np.float64).T)
    counts = Assay(vals, np.asarray(x.columns, dtype=object).astype(str), np.asarray(x.index, dtype=object).astype(str))
    yv = pd.Series(np.asarray(y), index=counts.samples)
    deggs_object = get_diffNetworks(assayData=counts, metadata=yv, regression_method="lm",
                                    percentile_vector=FILTER_PERCENTILES, padj_method="q.value",
                                    show_progressBar=False, verbose=False, cores=1, engine=engine)
    pairs = get_sig_deggs(deggs_object, 1, 0.05)
    keep: list[str] = []
    if keep_single_genes:
        if nfilter is not None:
            selected: list[int] = []
            total = 0
            for i in range(pairs.shape[0]):
                if total >= nfilter:
                    break
                cur = [pairs.iloc[i, 0], pairs.iloc[i, 1]]
                new = [g for g in dict.fromkeys(cur) if g not in keep]
                add = 1 + len(new)
                if total + add <= nfilter:
                    selected.append(i)
                    keep += new
                    total += add
                elif len(new) == 0 and total + 1 <= nfilter:
                    selected.append(i)
                    total += 1
                else:
                    remaining = nfilter - total
                    if remaining > 0 and new:
                        add_g = new[:min(remaining, len(new))]
                        keep += add_g
                        total += len(add_g)
                    break
            pairs = pairs.iloc[selected]
        else:
            keep = list(dict.fromkeys([g for i in range(pairs.shape[0]) for g in (pairs.iloc[i, 0], pairs.iloc[i, 1])]))
    elif nfilter is not None and pairs.shape[0] > nfilter:
        pairs = pairs.iloc[:nfilter]  # first-N rows in network order, NOT by significance (fs:148-150)
    sel = MultiDEGGsFilter(keep=keep, pairs=pairs)
    if pairs.shape[0] <= 1:
        kt = list(ttest_filter(y, x, nfilter)) if ttest_filter is not None else []
        sel = MultiDEGGsFilter(keep=kt, pairs=pd.DataFrame())
    return sel


def predict(obj: MultiDEGGsFilter, newdata: pd.DataFrame, interaction_type: str = "ratio", sep: str = ":", *,
            engine: Engine | None = None) -> pd.DataFrame:
    """predict.multiDEGGs_filter -> .predict_multiDEGGs (feature_selection_for_ML.R:406-463).
    The public S3 wrappers always pass interaction.type = "ratio" (fs:461)."""
    result = None
    keep = list(obj["keep"])
    if keep:
        result = newdata[keep]
    pairs = obj["pairs"]
    if pairs.shape[0] > 0:
        cols = {c: i for i, c in enumerate(newdata.columns)}
        a = np.array([cols[g] for g in pairs.iloc[:, 0]], dtype=np.int32)
        b = np.array([cols[g] for g in pairs.iloc[:, 1]], dtype=np.int32)
        eng = engine if engine is not None else default_engine()
        feats = eng.pair_features(newdata.to_numpy(dtype=np.float64), a, b, ratio=(interaction_type == "ratio"))
        names = [f"{x}{sep}{y}" for x, y in zip(pairs.iloc[:, 0], pairs.iloc[:, 1])]
        comb = pd.DataFrame(feats, index=newdata.index, columns=names)
        result = comb if result is None else pd.concat([result, comb], axis=1)
    if result is None:
        raise ValueError("The custom filter is empty")
    return result

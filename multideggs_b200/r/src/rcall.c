/*
 * rcall.c — the thin `.Call` shim between the R host (R/gpu_backend.R) and the C ABI of libmdeggs
 * (include/mdeggs.h).  It replaces what the reference does inside R between core_functions.R:336 and :495
 * (per-percentile mclapply loop -> calc_pvalues_percentile -> calc_pvalues_network -> fit_lm_interaction)
 * by ONE call per omic layer.
 *
 * Contract (SURVEY §8 b): inputs are borrowed R vectors (never written, never retained); every output is
 * allocated with Rf_allocVector and PROTECTed BEFORE any GPU work; the C ABI never longjmps, so Rf_error is
 * raised only after mdeggs_run_omic has returned (all device work quiesced).  The CUDA context lives in a
 * process-wide external pointer created lazily on first use (never before a fork()).
 *
 * Build: R CMD INSTALL links this file against libmdeggs.so.  In this repository it is
 * only compile-checked against the stub headers in src/stub/ (R is not installed in the build image).
 */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include <stdint.h>
#include <string.h>

#include "mdeggs.h"

static SEXP g_ctx_ptr = NULL; /* external pointer owning the mdeggs_ctx */

static void ctx_finalizer(SEXP p) {
  mdeggs_ctx* ctx = (mdeggs_ctx*)R_ExternalPtrAddr(p);
  if (ctx) mdeggs_destroy(ctx);
  R_ClearExternalPtr(p);
}

static mdeggs_ctx* get_ctx(int n_gpus) {
  if (g_ctx_ptr != NULL) {
    mdeggs_ctx* c = (mdeggs_ctx*)R_ExternalPtrAddr(g_ctx_ptr);
    if (c) return c;
  }
  int ids[16];
  if (n_gpus < 1) n_gpus = 1;
  if (n_gpus > 16) n_gpus = 16;
  for (int i = 0; i < n_gpus; ++i) ids[i] = i;
  mdeggs_ctx* ctx = NULL;
  int rc = mdeggs_create(&ctx, ids, n_gpus);
  if (rc != MDEGGS_OK) Rf_error("multiDEGGs GPU engine: cannot create context: %s", mdeggs_strerror(rc));
  g_ctx_ptr = R_MakeExternalPtr(ctx, R_NilValue, R_NilValue);
  R_PreserveObject(g_ctx_ptr);
  R_RegisterCFinalizerEx(g_ctx_ptr, ctx_finalizer, TRUE);
  return ctx;
}

static SEXP i64_to_real(const int64_t* v, int n) {
  SEXP out = PROTECT(Rf_allocVector(REALSXP, n));
  for (int i = 0; i < n; ++i) REAL(out)[i] = v[i] < 0 ? NA_REAL : (double)v[i];
  UNPROTECT(1);
  return out;
}

/* .Call("mdeggs_R_run", assay, from_idx, to_idx, group, K, percentiles, method, padj, tested_coef,
 *       rlm_cov_mode, n_gpus, want_all)  ->  named list */
SEXP mdeggs_R_run(SEXP assay, SEXP from_idx, SEXP to_idx, SEXP group, SEXP K_, SEXP percentiles, SEXP method_,
                  SEXP padj_, SEXP tested_coef_, SEXP cov_mode_, SEXP n_gpus_, SEXP want_all_) {
  if (TYPEOF(assay) != REALSXP || !Rf_isMatrix(assay)) Rf_error("assay must be a double matrix (genes x samples)");
  if (TYPEOF(from_idx) != INTSXP || TYPEOF(to_idx) != INTSXP || TYPEOF(group) != INTSXP)
    Rf_error("from_idx, to_idx and group must be integer vectors");
  if (TYPEOF(percentiles) != REALSXP) Rf_error("percentiles must be double");
  const int G = Rf_nrows(assay), n = Rf_ncols(assay);
  const R_xlen_t E = XLENGTH(from_idx);
  const int K = Rf_asInteger(K_), P = (int)XLENGTH(percentiles);
  const int padj = Rf_asInteger(padj_), want_all = Rf_asLogical(want_all_);
  if (XLENGTH(to_idx) != E) Rf_error("from_idx and to_idx differ in length");
  if (XLENGTH(group) != n) Rf_error("group must have one code per sample (column of assay)");

  mdeggs_problem in;
  memset(&in, 0, sizeof in);
  in.assay = REAL(assay);
  in.ld = G;
  in.G = G;
  in.n = n;
  in.layout = MDEGGS_LAYOUT_COLMAJOR;
  in.mem = MDEGGS_MEM_HOST;
  in.from = INTEGER(from_idx);
  in.to = INTEGER(to_idx);
  in.E = (int64_t)E;
  in.group = INTEGER(group);
  in.K = K;
  in.pct = REAL(percentiles);
  in.P = P;
  in.method = Rf_asInteger(method_);
  in.padj = padj;
  in.tested_coef = Rf_asInteger(tested_coef_);
  in.rlm_cov_mode = Rf_asInteger(cov_mode_);

  /* outputs: allocate + protect everything first */
  int np = 0;
  SEXP cutoff = PROTECT(Rf_allocVector(REALSXP, P)); ++np;
  SEXP p_edge = PROTECT(Rf_allocVector(REALSXP, E)); ++np;
  SEXP stat = PROTECT(Rf_allocVector(REALSXP, E)); ++np;
  SEXP status = PROTECT(Rf_allocVector(INTSXP, E)); ++np;
  SEXP df = PROTECT(Rf_allocVector(REALSXP, 2)); ++np;
  SEXP best = PROTECT(Rf_allocVector(INTSXP, 1)); ++np;
  SEXP cat_best = PROTECT(Rf_allocVector(RAWSXP, E)); ++np;
  SEXP padj_best = PROTECT(Rf_allocVector(REALSXP, E)); ++np;
  SEXP median = R_NilValue, coef = R_NilValue, iters = R_NilValue, cat = R_NilValue, padj_all = R_NilValue;
  const int host_adjust = (padj == MDEGGS_PADJ_HOST);
  if (want_all) {
    median = PROTECT(Rf_allocMatrix(REALSXP, G, K)); ++np;
    coef = PROTECT(Rf_allocMatrix(REALSXP, 2 * K, (int)E)); ++np;
    iters = PROTECT(Rf_allocVector(INTSXP, E)); ++np;
  }
  if (want_all || host_adjust) {
    cat = PROTECT(Rf_allocVector(RAWSXP, (R_xlen_t)P * E)); ++np;
  }
  if (want_all && !host_adjust && padj != MDEGGS_PADJ_NONE) {
    padj_all = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)P * E)); ++np;
  }
  int64_t* tot = (int64_t*)R_alloc((size_t)(P > 0 ? P : 1), sizeof(int64_t));
  int64_t* sig = (int64_t*)R_alloc((size_t)(P > 0 ? P : 1), sizeof(int64_t));
  int64_t* fail = (int64_t*)R_alloc((size_t)(P > 0 ? P : 1), sizeof(int64_t));
  for (R_xlen_t i = 0; i < E; ++i) REAL(padj_best)[i] = NA_REAL;

  mdeggs_result out;
  memset(&out, 0, sizeof out);
  out.cutoff = REAL(cutoff);
  out.p_edge = REAL(p_edge);
  out.stat = REAL(stat);
  out.status = INTEGER(status);
  out.df = REAL(df);
  out.best = INTEGER(best);
  out.cat_best = (int8_t*)RAW(cat_best);
  out.padj_best = REAL(padj_best);
  out.tot_edges = tot;
  out.sig_count = sig;
  out.fail_count = fail;
  if (median != R_NilValue) out.median = REAL(median);
  if (coef != R_NilValue) out.coef = REAL(coef);
  if (iters != R_NilValue) out.iters = INTEGER(iters);
  if (cat != R_NilValue) out.cat = (int8_t*)RAW(cat);
  if (padj_all != R_NilValue) out.padj = REAL(padj_all);

  mdeggs_ctx* ctx = get_ctx(Rf_asInteger(n_gpus_));
  int rc = mdeggs_run_omic(ctx, &in, &out); /* returns only after all device work has completed */
  if (rc != MDEGGS_OK) {
    char msg[600];
    snprintf(msg, sizeof msg, "multiDEGGs GPU engine: %s: %s", mdeggs_strerror(rc), mdeggs_last_error(ctx));
    UNPROTECT(np);
    Rf_error("%s", msg);
  }

  const char* names[] = {"cutoff", "p_edge", "stat", "status", "df", "best", "cat_best", "padj_best", "tot_edges",
                         "sig_count", "fail_count", "median", "coef", "iters", "cat", "padj"};
  const int nfield = 16;
  SEXP res = PROTECT(Rf_allocVector(VECSXP, nfield)); ++np;
  SEXP nms = PROTECT(Rf_allocVector(STRSXP, nfield)); ++np;
  for (int i = 0; i < nfield; ++i) SET_STRING_ELT(nms, i, Rf_mkChar(names[i]));
  SET_VECTOR_ELT(res, 0, cutoff);
  SET_VECTOR_ELT(res, 1, p_edge);
  SET_VECTOR_ELT(res, 2, stat);
  SET_VECTOR_ELT(res, 3, status);
  SET_VECTOR_ELT(res, 4, df);
  SET_VECTOR_ELT(res, 5, best);
  SET_VECTOR_ELT(res, 6, cat_best);
  SET_VECTOR_ELT(res, 7, padj_best);
  SET_VECTOR_ELT(res, 8, i64_to_real(tot, P));
  SET_VECTOR_ELT(res, 9, i64_to_real(sig, P));
  SET_VECTOR_ELT(res, 10, i64_to_real(fail, P));
  SET_VECTOR_ELT(res, 11, median);
  SET_VECTOR_ELT(res, 12, coef);
  SET_VECTOR_ELT(res, 13, iters);
  SET_VECTOR_ELT(res, 14, cat);
  SET_VECTOR_ELT(res, 15, padj_all);
  Rf_setAttrib(res, R_NamesSymbol, nms);
  UNPROTECT(np);
  return res;
}

/* .Call("mdeggs_R_pair_features", x, a, b, ratio): predict.multiDEGGs_filter features (fs:423-431) */
SEXP mdeggs_R_pair_features(SEXP x, SEXP a, SEXP b, SEXP ratio) {
  if (TYPEOF(x) != REALSXP || !Rf_isMatrix(x)) Rf_error("x must be a double matrix (samples x features)");
  const int n = Rf_nrows(x), np_ = (int)XLENGTH(a);
  SEXP out = PROTECT(Rf_allocMatrix(REALSXP, n, np_));
  mdeggs_ctx* ctx = get_ctx(1);
  int rc = mdeggs_pair_features(ctx, REAL(x), n, n, INTEGER(a), INTEGER(b), np_, Rf_asLogical(ratio), MDEGGS_MEM_HOST,
                                REAL(out));
  UNPROTECT(1);
  if (rc != MDEGGS_OK) Rf_error("multiDEGGs GPU engine: %s", mdeggs_strerror(rc));
  return out;
}

static const R_CallMethodDef call_methods[] = {{"mdeggs_R_run", (DL_FUNC)&mdeggs_R_run, 12},
                                               {"mdeggs_R_pair_features", (DL_FUNC)&mdeggs_R_pair_features, 4},
                                               {NULL, NULL, 0}};

void R_init_multiDEGGs(DllInfo* dll) {
  R_registerRoutines(dll, NULL, call_methods, NULL, NULL);
  R_useDynamicSymbols(dll, FALSE);
}

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

#include <unistd.h>

#include "mdeggs.h"

/* Cached contexts are keyed on (process id, number of GPUs):
 *  - a context created in a parent must never be used in a forked child (mclapply / nestedcv cv.cores > 1): the child
 *    forgets the parent's pointer WITHOUT destroying it (the CUDA state behind it is not valid there) and creates its
 *    own on first use;
 *  - options(multiDEGGs.gpus = ...) changed between calls: the old context is destroyed and a new one is created. */
static SEXP g_ctx_ptr = NULL; /* external pointer owning the mdeggs_ctx */
static int g_ctx_pid = 0, g_ctx_gpus = 0;

static void ctx_finalizer(SEXP p) {
  mdeggs_ctx* ctx = (mdeggs_ctx*)R_ExternalPtrAddr(p);
  if (ctx) mdeggs_destroy(ctx);
  R_ClearExternalPtr(p);
}

static void forget_ctx(SEXP* slot, int destroy) {
  if (*slot == NULL) return;
  if (destroy) ctx_finalizer(*slot);
  else R_ClearExternalPtr(*slot); /* another process's context: leak it here, never touch it */
  R_ReleaseObject(*slot);
  *slot = NULL;
}

static mdeggs_ctx* get_ctx(int n_gpus) {
  if (n_gpus < 1) n_gpus = 1;
  if (n_gpus > 16) n_gpus = 16;
  const int pid = (int)getpid();
  if (g_ctx_ptr != NULL && g_ctx_pid != pid) forget_ctx(&g_ctx_ptr, 0);
  if (g_ctx_ptr != NULL && g_ctx_gpus != n_gpus) forget_ctx(&g_ctx_ptr, 1);
  if (g_ctx_ptr != NULL) {
    mdeggs_ctx* c = (mdeggs_ctx*)R_ExternalPtrAddr(g_ctx_ptr);
    if (c) return c;
  }
  int ids[16];
  for (int i = 0; i < n_gpus; ++i) ids[i] = i;
  mdeggs_ctx* ctx = NULL;
  int rc = mdeggs_create(&ctx, ids, n_gpus);
  if (rc != MDEGGS_OK) Rf_error("multiDEGGs GPU engine: cannot create context: %s", mdeggs_strerror(rc));
  g_ctx_ptr = R_MakeExternalPtr(ctx, R_NilValue, R_NilValue);
  R_PreserveObject(g_ctx_ptr);
  R_RegisterCFinalizerEx(g_ctx_ptr, ctx_finalizer, TRUE);
  g_ctx_pid = pid;
  g_ctx_gpus = n_gpus;
  return ctx;
}

static SEXP i64_to_real(const int64_t* v, int n) {
  SEXP out = PROTECT(Rf_allocVector(REALSXP, n));
  for (int i = 0; i < n; ++i) REAL(out)[i] = v[i] < 0 ? NA_REAL : (double)v[i];
  UNPROTECT(1);
  return out;
}

/* One omic layer on the R side of the boundary: the borrowed inputs, the R vectors that receive the results
 * (allocated and PROTECTed before any GPU work; `np` counts them) and the two C structs handed to the engine. */
typedef struct {
  mdeggs_problem in;
  mdeggs_result out;
  SEXP cutoff, p_edge, stat, status, df, best, cat_best, padj_best, median, coef, iters, cat, padj_all;
  int64_t *tot, *sig, *fail;
  int P, np;
} layer_t;

/* argument checks + output allocation; may raise an R error (nothing of this layer is on the GPU yet) */
static void layer_prepare(layer_t* L, SEXP assay, SEXP from_idx, SEXP to_idx, SEXP group, SEXP K_, SEXP percentiles,
                          SEXP method_, SEXP padj_, SEXP tested_coef_, SEXP cov_mode_, SEXP want_all_) {
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

  memset(L, 0, sizeof *L);
  mdeggs_problem* in = &L->in;
  in->assay = REAL(assay);
  in->ld = G;
  in->G = G;
  in->n = n;
  in->layout = MDEGGS_LAYOUT_COLMAJOR;
  in->mem = MDEGGS_MEM_HOST;
  in->from = INTEGER(from_idx);
  in->to = INTEGER(to_idx);
  in->E = (int64_t)E;
  in->group = INTEGER(group);
  in->K = K;
  in->pct = REAL(percentiles);
  in->P = P;
  in->method = Rf_asInteger(method_);
  in->padj = padj;
  in->tested_coef = Rf_asInteger(tested_coef_);
  in->rlm_cov_mode = Rf_asInteger(cov_mode_);

  /* outputs: allocate + protect everything first */
  int np = 0;
  L->P = P;
  L->cutoff = PROTECT(Rf_allocVector(REALSXP, P)); ++np;
  L->p_edge = PROTECT(Rf_allocVector(REALSXP, E)); ++np;
  L->stat = PROTECT(Rf_allocVector(REALSXP, E)); ++np;
  L->status = PROTECT(Rf_allocVector(INTSXP, E)); ++np;
  L->df = PROTECT(Rf_allocVector(REALSXP, 2)); ++np;
  L->best = PROTECT(Rf_allocVector(INTSXP, 1)); ++np;
  L->cat_best = PROTECT(Rf_allocVector(RAWSXP, E)); ++np;
  L->padj_best = PROTECT(Rf_allocVector(REALSXP, E)); ++np;
  L->median = L->coef = L->iters = L->cat = L->padj_all = R_NilValue;
  const int host_adjust = (padj == MDEGGS_PADJ_HOST);
  if (want_all) {
    L->median = PROTECT(Rf_allocMatrix(REALSXP, G, K)); ++np;
    L->coef = PROTECT(Rf_allocMatrix(REALSXP, 2 * K, (int)E)); ++np;
    L->iters = PROTECT(Rf_allocVector(INTSXP, E)); ++np;
  }
  if (want_all || host_adjust) {
    L->cat = PROTECT(Rf_allocVector(RAWSXP, (R_xlen_t)P * E)); ++np;
  }
  if (want_all && !host_adjust && padj != MDEGGS_PADJ_NONE) {
    L->padj_all = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)P * E)); ++np;
  }
  L->np = np;
  L->tot = (int64_t*)R_alloc((size_t)(P > 0 ? P : 1), sizeof(int64_t));
  L->sig = (int64_t*)R_alloc((size_t)(P > 0 ? P : 1), sizeof(int64_t));
  L->fail = (int64_t*)R_alloc((size_t)(P > 0 ? P : 1), sizeof(int64_t));
  for (R_xlen_t i = 0; i < E; ++i) REAL(L->padj_best)[i] = NA_REAL;

  mdeggs_result* out = &L->out;
  out->cutoff = REAL(L->cutoff);
  out->p_edge = REAL(L->p_edge);
  out->stat = REAL(L->stat);
  out->status = INTEGER(L->status);
  out->df = REAL(L->df);
  out->best = INTEGER(L->best);
  out->cat_best = (int8_t*)RAW(L->cat_best);
  out->padj_best = REAL(L->padj_best);
  out->tot_edges = L->tot;
  out->sig_count = L->sig;
  out->fail_count = L->fail;
  if (L->median != R_NilValue) out->median = REAL(L->median);
  if (L->coef != R_NilValue) out->coef = REAL(L->coef);
  if (L->iters != R_NilValue) out->iters = INTEGER(L->iters);
  if (L->cat != R_NilValue) out->cat = (int8_t*)RAW(L->cat);
  if (L->padj_all != R_NilValue) out->padj = REAL(L->padj_all);
}

/* results of one layer -> named R list (2 more PROTECTs, counted in L->np) */
static SEXP layer_pack(layer_t* L) {
  const char* names[] = {"cutoff", "p_edge", "stat", "status", "df", "best", "cat_best", "padj_best", "tot_edges",
                         "sig_count", "fail_count", "median", "coef", "iters", "cat", "padj"};
  const int nfield = 16;
  SEXP res = PROTECT(Rf_allocVector(VECSXP, nfield)); ++L->np;
  SEXP nms = PROTECT(Rf_allocVector(STRSXP, nfield)); ++L->np;
  for (int i = 0; i < nfield; ++i) SET_STRING_ELT(nms, i, Rf_mkChar(names[i]));
  SET_VECTOR_ELT(res, 0, L->cutoff);
  SET_VECTOR_ELT(res, 1, L->p_edge);
  SET_VECTOR_ELT(res, 2, L->stat);
  SET_VECTOR_ELT(res, 3, L->status);
  SET_VECTOR_ELT(res, 4, L->df);
  SET_VECTOR_ELT(res, 5, L->best);
  SET_VECTOR_ELT(res, 6, L->cat_best);
  SET_VECTOR_ELT(res, 7, L->padj_best);
  SET_VECTOR_ELT(res, 8, i64_to_real(L->tot, L->P));
  SET_VECTOR_ELT(res, 9, i64_to_real(L->sig, L->P));
  SET_VECTOR_ELT(res, 10, i64_to_real(L->fail, L->P));
  SET_VECTOR_ELT(res, 11, L->median);
  SET_VECTOR_ELT(res, 12, L->coef);
  SET_VECTOR_ELT(res, 13, L->iters);
  SET_VECTOR_ELT(res, 14, L->cat);
  SET_VECTOR_ELT(res, 15, L->padj_all);
  Rf_setAttrib(res, R_NamesSymbol, nms);
  return res;
}

/* .Call("mdeggs_R_run", assay, from_idx, to_idx, group, K, percentiles, method, padj, tested_coef,
 *       rlm_cov_mode, n_gpus, want_all)  ->  named list */
SEXP mdeggs_R_run(SEXP assay, SEXP from_idx, SEXP to_idx, SEXP group, SEXP K_, SEXP percentiles, SEXP method_,
                  SEXP padj_, SEXP tested_coef_, SEXP cov_mode_, SEXP n_gpus_, SEXP want_all_) {
  layer_t L;
  layer_prepare(&L, assay, from_idx, to_idx, group, K_, percentiles, method_, padj_, tested_coef_, cov_mode_,
                want_all_);
  mdeggs_ctx* ctx = get_ctx(Rf_asInteger(n_gpus_));
  int rc = mdeggs_run_omic(ctx, &L.in, &L.out); /* returns only after all device work has completed */
  if (rc != MDEGGS_OK) {
    char msg[600];
    snprintf(msg, sizeof msg, "multiDEGGs GPU engine: %s: %s", mdeggs_strerror(rc), mdeggs_last_error(ctx));
    UNPROTECT(L.np);
    Rf_error("%s", msg);
  }
  SEXP res = layer_pack(&L);
  UNPROTECT(L.np);
  return res;
}

/* .Call("mdeggs_R_run_layers", layers): every omic layer of one get_diffNetworks call (the loop of
 * core_functions.R:214-234) as ONE launch set.  `layers` is a list; each element is the list of the first ten
 * arguments of mdeggs_R_run followed by want_all (n_gpus does not apply: one single-GPU context per layer).
 * Returns a list of the same length: the result list of the layer, or a character(1) error message where the
 * engine failed for that layer (the caller turns it into message() like the reference's per-omic tryCatch). */
#define MDEGGS_R_MAX_LAYERS 16
static SEXP g_layer_ctx[MDEGGS_R_MAX_LAYERS];
static int g_layer_pid[MDEGGS_R_MAX_LAYERS];

static mdeggs_ctx* get_layer_ctx(int i) {
  const int pid = (int)getpid();
  if (g_layer_ctx[i] != NULL && g_layer_pid[i] != pid) forget_ctx(&g_layer_ctx[i], 0); /* forked child: see get_ctx */
  if (g_layer_ctx[i] != NULL) {
    mdeggs_ctx* c = (mdeggs_ctx*)R_ExternalPtrAddr(g_layer_ctx[i]);
    if (c) return c;
  }
  int dev = 0;
  mdeggs_ctx* ctx = NULL;
  int rc = mdeggs_create(&ctx, &dev, 1);
  if (rc != MDEGGS_OK) Rf_error("multiDEGGs GPU engine: cannot create context: %s", mdeggs_strerror(rc));
  g_layer_ctx[i] = R_MakeExternalPtr(ctx, R_NilValue, R_NilValue);
  R_PreserveObject(g_layer_ctx[i]);
  R_RegisterCFinalizerEx(g_layer_ctx[i], ctx_finalizer, TRUE);
  g_layer_pid[i] = pid;
  return ctx;
}

SEXP mdeggs_R_run_layers(SEXP layers) {
  if (TYPEOF(layers) != VECSXP) Rf_error("layers must be a list");
  const int nl = (int)XLENGTH(layers);
  if (nl > MDEGGS_R_MAX_LAYERS) Rf_error("at most %d layers per call", MDEGGS_R_MAX_LAYERS);
  layer_t* L = (layer_t*)R_alloc((size_t)(nl > 0 ? nl : 1), sizeof(layer_t));
  mdeggs_ctx* ctxs[MDEGGS_R_MAX_LAYERS];
  mdeggs_problem in[MDEGGS_R_MAX_LAYERS];
  mdeggs_result out[MDEGGS_R_MAX_LAYERS];
  int rcs[MDEGGS_R_MAX_LAYERS];
  int np = 0;
  for (int i = 0; i < nl; ++i) { /* every R allocation and every context before any GPU work */
    SEXP a = VECTOR_ELT(layers, i);
    if (TYPEOF(a) != VECSXP || XLENGTH(a) != 11) Rf_error("layer %d: expected a list of 11 arguments", i + 1);
    layer_prepare(&L[i], VECTOR_ELT(a, 0), VECTOR_ELT(a, 1), VECTOR_ELT(a, 2), VECTOR_ELT(a, 3), VECTOR_ELT(a, 4),
                  VECTOR_ELT(a, 5), VECTOR_ELT(a, 6), VECTOR_ELT(a, 7), VECTOR_ELT(a, 8), VECTOR_ELT(a, 9),
                  VECTOR_ELT(a, 10));
    np += L[i].np;
    ctxs[i] = get_layer_ctx(i);
    in[i] = L[i].in;
    out[i] = L[i].out;
  }
  mdeggs_run_omics(ctxs, nl, in, out, rcs); /* returns only after the device work of every layer has completed */
  SEXP res = PROTECT(Rf_allocVector(VECSXP, nl)); ++np;
  for (int i = 0; i < nl; ++i) {
    if (rcs[i] == MDEGGS_OK) {
      const int before = L[i].np;
      SET_VECTOR_ELT(res, i, layer_pack(&L[i]));
      np += L[i].np - before;
    } else {
      char msg[600];
      snprintf(msg, sizeof msg, "multiDEGGs GPU engine: %s: %s", mdeggs_strerror(rcs[i]), mdeggs_last_error(ctxs[i]));
      SET_VECTOR_ELT(res, i, Rf_mkString(msg));
    }
  }
  UNPROTECT(np);
  return res;
}

/* .Call("mdeggs_R_pair_features", x, a, b, ratio): predict.multiDEGGs_filter features (fs:423-431) */
SEXP mdeggs_R_pair_features(SEXP x, SEXP a, SEXP b, SEXP ratio) {
  if (TYPEOF(x) != REALSXP || !Rf_isMatrix(x)) Rf_error("x must be a double matrix (samples x features)");
  if (TYPEOF(a) != INTSXP || TYPEOF(b) != INTSXP || XLENGTH(a) != XLENGTH(b))
    Rf_error("a and b must be integer vectors of the same length (0-based column indices)");
  const int n = Rf_nrows(x), np_ = (int)XLENGTH(a), nc = Rf_ncols(x);
  for (int j = 0; j < np_; ++j)
    if (INTEGER(a)[j] < 0 || INTEGER(a)[j] >= nc || INTEGER(b)[j] < 0 || INTEGER(b)[j] >= nc)
      Rf_error("pair %d refers to a column outside 0..%d", j + 1, nc - 1);
  SEXP out = PROTECT(Rf_allocMatrix(REALSXP, n, np_));
  mdeggs_ctx* ctx = get_ctx(1);
  int rc = mdeggs_pair_features(ctx, REAL(x), n, n, INTEGER(a), INTEGER(b), np_, Rf_asLogical(ratio), MDEGGS_MEM_HOST,
                                REAL(out));
  UNPROTECT(1);
  if (rc != MDEGGS_OK) Rf_error("multiDEGGs GPU engine: %s", mdeggs_strerror(rc));
  return out;
}

/* .Call("mdeggs_R_pair_moments", from_idx, to_idx, K): per-category n, means, Sxx, Syy, Sxy of gene pairs from what the
 * last mdeggs_R_run of this process left on the device: the lm(y ~ x) lines and confidence bands of plot_regressions
 * (plotting_functions.R:125-172).  Returns a 6 x K x n_pairs array. */
SEXP mdeggs_R_pair_moments(SEXP a, SEXP b, SEXP K) {
  if (TYPEOF(a) != INTSXP || TYPEOF(b) != INTSXP || XLENGTH(a) != XLENGTH(b))
    Rf_error("from_idx and to_idx must be integer vectors of the same length (1-based gene rows)");
  const int np_ = (int)XLENGTH(a), k = Rf_asInteger(K);
  if (k < 2 || k > MDEGGS_MAX_K) Rf_error("K outside 2..%d", MDEGGS_MAX_K);
  SEXP out = PROTECT(Rf_alloc3DArray(REALSXP, 6, k, np_));
  mdeggs_ctx* ctx = get_ctx(1);
  int rc = mdeggs_pair_moments(ctx, INTEGER(a), INTEGER(b), np_, REAL(out));
  UNPROTECT(1);
  if (rc != MDEGGS_OK) Rf_error("multiDEGGs GPU engine: %s (%s)", mdeggs_strerror(rc), mdeggs_last_error(ctx));
  return out;
}

static const R_CallMethodDef call_methods[] = {{"mdeggs_R_run", (DL_FUNC)&mdeggs_R_run, 12},
                                               {"mdeggs_R_run_layers", (DL_FUNC)&mdeggs_R_run_layers, 1},
                                               {"mdeggs_R_pair_features", (DL_FUNC)&mdeggs_R_pair_features, 4},
                                               {"mdeggs_R_pair_moments", (DL_FUNC)&mdeggs_R_pair_moments, 3},
                                               {NULL, NULL, 0}};

void R_init_multiDEGGs(DllInfo* dll) {
  R_registerRoutines(dll, NULL, call_methods, NULL, NULL);
  R_useDynamicSymbols(dll, FALSE);
}

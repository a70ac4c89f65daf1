/*
 * mdeggs.h — C ABI of the B200 engine for the multiDEGGs differential-correlation hot path.
 *
 * One call (`mdeggs_run_omic`) replaces, for one omic layer, everything the reference does between
 * "edges and node rows selected" (R/core_functions.R:319-336) and "best percentile chosen"
 * (R/core_functions.R:471-495):
 *
 *   reference (all citations relative to the reference repo)            this ABI
 *   ------------------------------------------------------------------  ------------------------------
 *   assayData[rownames %in% nodes, ]            core_functions.R:336     node-row gather (kernel K1)
 *   apply(category_vec, 1, median, na.rm=TRUE)  core_functions.R:353-365 out.median      (kernel K2)
 *   quantile(as.matrix(assayData), p)           core_functions.R:718     out.cutoff      (kernel K3)
 *   fit_lm_interaction()                        core_functions.R:1058-1074  out.p_edge/stat/coef (K4)
 *   aov(B ~ A * metadata), row 3                core_functions.R:963-968    out.p_edge/stat/coef (K5)
 *   MASS::rlm + sfsmisc::f.robftest(var = 3)    core_functions.R:914-925    out.p_edge/stat/iters (K6)
 *   calc_pvalues_percentile() set logic         core_functions.R:723-779    out.cat/tot_edges   (K7)
 *   p.adjust(bonferroni|BH) + count < 0.05      core_functions.R:1027-1035, 796-808  out.padj/sig_count (K8)
 *   which.max(sig_pvalues)                      core_functions.R:475-495    out.best
 *
 * The R host keeps strings, factors, data.frames, message()/stop() texts; it calls
 * this ABI through the `.Call` shim in multideggs_b200/r/src/rcall.c.  All functions are `extern "C"`,
 * take plain pointers and sizes, never throw, and return 0 or a negative MDEGGS_E* code.
 */
#ifndef MDEGGS_H
#define MDEGGS_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MDEGGS_ABI_VERSION 1
#define MDEGGS_MAX_K 8 /* categories per run (the reference has no limit; >8 is rejected) */

/* regression_method (core_functions.R:143) */
#define MDEGGS_METHOD_LM 0
#define MDEGGS_METHOD_RLM 1
/* padj_method (core_functions.R:58-62, 1012, 1027) */
#define MDEGGS_PADJ_NONE 0       /* "none": count raw p < 0.05 */
#define MDEGGS_PADJ_BONFERRONI 1 /* "bonferroni" */
#define MDEGGS_PADJ_BH 2         /* "BH" / "fdr" */
#define MDEGGS_PADJ_HOST 3       /* any method adjusted by the host (big hommel jobs, a host-side q.value): raw p + masks returned */
#define MDEGGS_PADJ_HOLM 4       /* "holm"      pmin(1, cummax((n+1-i) p[o])), o ascending                      */
#define MDEGGS_PADJ_HOCHBERG 5   /* "hochberg"  pmin(1, cummin((n+1-i) p[o])), o descending                     */
#define MDEGGS_PADJ_BY 6         /* "BY"        pmin(1, cummin(q n/i p[o])), q = sum(1/(1:n))                   */
#define MDEGGS_PADJ_QVALUE 7     /* "q.value"   qvalue::qvalue(p)$qvalues = pi0 pmin(1, cummin(p[o] m/i)), pi0 by the
                                    smoother over lambda = seq(.05,.95,.05); on qvalue's errors pi0 = 1 if the
                                    category network has > 1 row, else NA (core_functions.R:1012-1025)              */
#define MDEGGS_PADJ_HOMMEL 8     /* "hommel"    pmax(pa, p), pa = max over m of the Simes-type bounds of stats::p.adjust;
                                    O(n^2) per (percentile, category) segment: accepted up to MDEGGS_HOMMEL_MAX_E edges,
                                    beyond that MDEGGS_EUNSUPPORTED (use MDEGGS_PADJ_HOST and adjust on the host) */
#define MDEGGS_HOMMEL_MAX_E 65536
/* rlm covariance used by the Wald test (MASS::summary.rlm `method`) */
#define MDEGGS_RLM_COV_XTWX 0
#define MDEGGS_RLM_COV_XTX 1
/* assay memory layout */
#define MDEGGS_LAYOUT_COLMAJOR 0 /* R matrix genes x samples: element (g,s) at assay[g + s*ld]          */
#define MDEGGS_LAYOUT_ROWMAJOR 1 /* genes x samples, sample fastest: element (g,s) at assay[g*ld + s]
                                    (= the samples x genes column-major `x` of multiDEGGs_filter, fs:81) */
/* where the caller's buffers live */
#define MDEGGS_MEM_HOST 0
#define MDEGGS_MEM_DEVICE 1 /* all problem/result pointers are device pointers on the context's first GPU */

/* per-edge status (never an error of the call; the host decides what R would have done) */
#define MDEGGS_EDGE_OK 0
#define MDEGGS_EDGE_SINGULAR 1    /* K == 2: solve(XtX) would throw (core_functions.R:1068) / rlm refuses a singular design */
#define MDEGGS_EDGE_NONFINITE 2   /* K == 2 lm: NA/NaN/Inf in either gene row (.lm.fit errors, core_functions.R:1065);
                                     rlm / aov: an Inf among the pair's complete cases (lm.wfit / lm.fit error)  */
#define MDEGGS_EDGE_NOT_CONVERGED 3 /* rlm: 20 IRLS iterations without convergence (R: warning only)  */
#define MDEGGS_EDGE_ZERO_SCALE 4  /* rlm: MAD scale == 0                                              */
#define MDEGGS_EDGE_SELFLOOP 5    /* exact fit: from == to (the built-in network has one self-loop) or a response
                                     that is constant inside every category; R returns rounding noise / noise,
                                     the engine returns p = 1, stat = 0 (documented deviation)              */
#define MDEGGS_EDGE_ALIASED 6     /* K >= 3 only, soft: the predictor is constant inside a category (or the category has
                                     one sample), so aov() drops the aliased A:metadata column(s) instead of failing
                                     (core_functions.R:966-968): F and p use the reduced degrees of freedom
                                     (Ke - 1 - aliased, n - 2 Ke + aliased), not the run's `df`; when no interaction
                                     column is left the summary table has no third term row and p is NA (NaN)      */
#define MDEGGS_EDGE_NA_OMIT 7     /* rlm and aov only, soft: the samples where either gene is NA/NaN were dropped for this
                                     pair, as model.frame(na.action = na.omit) does inside rlm() / aov()
                                     (core_functions.R:916-917, 966-967): the fit, F and p use the n' complete cases and
                                     their degrees of freedom (K >= 3: Ke' - 1, n' - 2 Ke'; rlm: 1, n' - 4), not the
                                     run's `df`.  Fewer than 2 levels or no residual d.f. left: MDEGGS_EDGE_SINGULAR   */

/* error codes */
#define MDEGGS_OK 0
#define MDEGGS_EINVAL (-1)
#define MDEGGS_ECUDA (-2)
#define MDEGGS_ENCCL (-3)
#define MDEGGS_ENOMEM (-4)
#define MDEGGS_EUNSUPPORTED (-5)

typedef struct mdeggs_ctx mdeggs_ctx;

typedef struct mdeggs_problem {
  const double* assay; /* G x n doubles, borrowed, never written                                        */
  int64_t ld;          /* leading dimension (>= G col-major, >= n row-major)                            */
  int32_t G, n;        /* rows (genes/features) and columns (samples) AFTER sample alignment            */
  int32_t layout;      /* MDEGGS_LAYOUT_*                                                               */
  int32_t mem;         /* MDEGGS_MEM_*                                                                  */
  const int32_t* from; /* [E] 1-based row index of gene A (predictor), network order, already de-duplicated */
  const int32_t* to;   /* [E] 1-based row index of gene B (response)                                    */
  int64_t E;
  const int32_t* group; /* [n] category code 1..K of each sample (as.integer(factor)); always HOST memory */
  int32_t K;
  const double* pct; /* [P] percentile_vector as the doubles R computed (seq(from,to,by)); always HOST memory */
  int32_t P;
  int32_t method;       /* MDEGGS_METHOD_*; ignored when K >= 3 (core_functions.R:963)                  */
  int32_t padj;         /* MDEGGS_PADJ_*                                                                */
  int32_t tested_coef;  /* rlm only: 1-based coefficient index of the Wald test; reference uses 3        */
  int32_t rlm_cov_mode; /* MDEGGS_RLM_COV_*                                                             */
} mdeggs_problem;

/* Caller-allocated outputs; any pointer may be NULL (= not wanted, not computed/copied).
 * Arrays indexed [P x E] have E fastest: x[p*E + e]. */
typedef struct mdeggs_result {
  double* cutoff;     /* [P]     type-7 quantile of the node-filtered matrix                            */
  double* median;     /* [K x G] median[k*G + g]; NaN for rows that are not network nodes               */
  double* p_edge;     /* [E]     raw p-value of the edge (independent of percentile/category)           */
  double* stat;       /* [E]     t (K==2 lm), Wald F (rlm), interaction F (K>=3)                        */
  double* coef;       /* [2K x E] coef[e*2K + j], R order: (Intercept), A, metadata2..K, A:metadata2..K  */
  double* df;         /* [2]     numerator, denominator degrees of freedom                              */
  int32_t* status;    /* [E]     MDEGGS_EDGE_*                                                          */
  int32_t* iters;     /* [E]     rlm IRLS iterations (0 otherwise)                                      */
  int8_t* cat;        /* [P x E] 0 = not tested at this percentile, c = specific to category c (1..K)   */
  double* padj;       /* [P x E] adjusted p (NaN where cat == 0); not filled for PADJ_NONE/PADJ_HOST    */
  int64_t* tot_edges; /* [P]     edges left at this percentile (core_functions.R:773-779)               */
  int64_t* sig_count; /* [P]     sig_pvalues_count (core_functions.R:796-808); -1 where tot_edges == 0  */
  int64_t* fail_count; /* [P]    tested edges whose fit failed hard (status SINGULAR/NONFINITE): in R such a
                                 percentile is lost (cores > 1) or the whole omic errors (cores == 1), quirk q7 */
  int32_t* best;      /* [1]     1-based index of the first maximum of sig_count over percentiles with
                                 tot_edges > 0 and fail_count == 0 (core_functions.R:475-495); 0 if none   */
  int8_t* cat_best;   /* [E]     cat row of the best percentile                                         */
  double* padj_best;  /* [E]     padj row of the best percentile                                        */
} mdeggs_result;

typedef struct mdeggs_stats {
  int64_t n_nodes;        /* rows kept by the node filter                                               */
  int64_t unique_fits;    /* pair fits executed (== E)                                                  */
  int64_t ref_equiv_tests; /* sum over percentiles of tot_edges: fits the reference would have run      */
  int64_t kernel_launches; /* CUDA kernels launched by the last call (this library's + CUB's)           */
  int64_t h2d_bytes, d2h_bytes;
  int32_t fit_path;       /* bits 0-7: 0 warp-per-edge streaming, 1 dense tile cross-moments, 2 rlm IRLS;
                             bit 8: the compute phase was replayed from a CUDA graph;
                             bit 9: K8 ordered the edges through p-value buckets (no radix sort);
                             bit 10 / 11: fused front kernel fed by TMA / by its plain loader;
                             bit 12: ... while the host matrix was still streaming in (option "h2d_pipeline");
                             bit 13: the call ran on host-packed node rows (option "host_pack");
                             bit 14: several ranks exchanged their p-value blocks by a peer-memory push (option "p2p") */
  int32_t n_ranks;
  float ms_total;         /* device time of the whole pipeline (CUDA events on the engine stream)       */
  float ms_h2d, ms_gather, ms_stats, ms_quantile, ms_fit, ms_tail, ms_collective, ms_percolate, ms_adjust, ms_d2h;
  int64_t fit_bytes;      /* algorithmic bytes of the fit kernel: E * (16 n + 16)                        */
} mdeggs_stats;

/* Single process, n_dev >= 1 GPUs of this host (edges sharded, NCCL all-gather when n_dev > 1). */
int mdeggs_create(mdeggs_ctx** out, const int* device_ids, int n_dev);
/* One process per GPU (torchrun): `unique_id` = 128 bytes from mdeggs_nccl_unique_id() on rank 0,
 * distributed to all ranks by the caller (torch.distributed broadcast). world == 1 ignores it. */
int mdeggs_create_rank(mdeggs_ctx** out, int device_id, int rank, int world, const void* unique_id);
int mdeggs_nccl_unique_id(void* out128);
int mdeggs_destroy(mdeggs_ctx* ctx);

/* Runs the whole path for one omic layer and waits for completion. With MDEGGS_MEM_DEVICE nothing
 * crosses PCIe except `group`, `pct` (tiny) and one 8-byte count. */
int mdeggs_run_omic(mdeggs_ctx* ctx, const mdeggs_problem* in, mdeggs_result* out);

/* The same call in two halves: mdeggs_submit_omic enqueues the input copies, the whole path and the output copies on
 * the context's streams and returns without waiting; mdeggs_wait blocks until that work is done, delivers the small
 * host results and the statistics, and reports errors found on the device. Every buffer named by `in` / `out`
 * stays borrowed until mdeggs_wait returns. One call may be in flight per context. */
int mdeggs_submit_omic(mdeggs_ctx* ctx, const mdeggs_problem* in, mdeggs_result* out);
int mdeggs_wait(mdeggs_ctx* ctx);

/* All omic layers of one get_diffNetworks call (the per-omic loop of core_functions.R:214-234) in one launch set:
 * layer i is submitted on ctxs[i] (distinct contexts, normally all on the same GPU) before any is waited for, so
 * the layers' kernels and copies overlap on the device instead of paying one launch/sync latency chain each.
 * A failing layer does not stop the others (the reference's per-omic tryCatch, core_functions.R:219-233):
 * layer_rc[i] (optional) receives its code, the return value is the first non-zero one. */
int mdeggs_run_omics(mdeggs_ctx* const* ctxs, int n_layers, const mdeggs_problem* in, mdeggs_result* out,
                     int* layer_rc);

/* Feature construction of predict.multiDEGGs_filter (feature_selection_for_ML.R:406-445):
 * out[s + j*n] = x[s + a[j]*ldx] / x[s + b[j]*ldx]  (or product), non-finite -> 0; x is n x G col-major,
 * a/b are 0-based column indices. mem as above. */
int mdeggs_pair_features(mdeggs_ctx* ctx, const double* x, int64_t ldx, int32_t n, const int32_t* a,
                         const int32_t* b, int32_t n_pairs, int32_t ratio, int32_t mem, double* out);

/* Per-category sufficient statistics of gene pairs, from the gene-major matrix and the per-gene moments that the LAST
 * completed mdeggs_run_omic of this context left on the device: the refit behind the per-category regression lines and
 * confidence bands of plot_regressions (plotting_functions.R:125-162: lm(y ~ x) inside each category +
 * predict(interval = "confidence")) without sending the rows again.  a_rows / b_rows: 1-based gene rows as in from / to
 * (host memory); out (host): out[(j K + k) 6 + 0..5] = n_k, mean_A, mean_B, Sxx_A, Syy_B, Sxy of pair j in category k
 * (centred sums over the category's samples); NaN for a row that was no network node of that run.
 * slope = Sxy / Sxx_A, intercept = mean_B - slope mean_A, sigma^2 = (Syy_B - slope Sxy) / (n_k - 2),
 * se(fit at x0)^2 = sigma^2 (1 / n_k + (x0 - mean_A)^2 / Sxx_A). */
int mdeggs_pair_moments(mdeggs_ctx* ctx, const int32_t* a_rows, const int32_t* b_rows, int32_t n_pairs, double* out);

/* Host-only utility (no GPU needed), the packing step of option "host_pack" on its own: the rows of the column-major
 * G x n matrix `src` (leading dimension ld) that are end points of the edge list (core_functions.R:333-336) are gathered,
 * in ascending row order, into the column-major block dst (leading dimension ld_dst >= node count) by the engine's pool
 * of packing threads, in row chunks of chunk_rows (0 = one chunk).  rows_out (optional) receives the 0-based source row
 * of every packed row, *n_nodes the node count; dst == NULL only counts. */
int mdeggs_host_pack_rows(const double* src, int64_t ld, int32_t G, int32_t n, const int32_t* from, const int32_t* to,
                          int64_t E, int32_t chunk_rows, double* dst, int64_t ld_dst, int32_t* rows_out, int32_t* n_nodes);

/* Measures this GPU's FP64 FMA peak with a register-only DFMA-chain kernel (TFLOP/s): the denominator of the
 * rlm (K6) roofline, which is FP64-pipe bound (SURVEY section 8 d). */
int mdeggs_measure_fp64_peak(mdeggs_ctx* ctx, double* tflops_out);
/* Measures this GPU's L2 -> SM read bandwidth (GB/s) over an L2-resident buffer of `bytes` (>= 1 MiB; use the size of
 * the gene-major matrix, <= ~100 MB): the denominator of the pair-fit kernel's roofline, whose 16 n + 16 algorithmic
 * bytes per fit are served by L2, not by HBM (the matrix is read ~12 times per run and fits the 126 MB L2). */
int mdeggs_measure_l2_peak(mdeggs_ctx* ctx, int64_t bytes, double* gbs_out);
int mdeggs_last_stats(const mdeggs_ctx* ctx, mdeggs_stats* out);
/* Per-launch timeline of the last call when option "trace" was 1 (printed to stderr as well), 2 or 3 (silent): entry i =
 * kernel name, stream (0 main, 1 side) and the time its CUDA event fired, in microseconds after the call's first
 * event.  bench.py derives the per-kernel durations of its roofline table from it (CUDA events on the launching
 * stream; traced calls launch eagerly). */
int mdeggs_trace_count(const mdeggs_ctx* ctx);
int mdeggs_trace_get(const mdeggs_ctx* ctx, int i, char* name, int name_cap, int* stream_id, float* end_us);
/* Engine options: "graph" (1 = replay the compute phase from a CUDA graph when a call repeats with identical
 * shapes and buffers [default], 0 = always launch eagerly, which also keeps the per-stage timings);
 * "front_path" (-1 auto, 0 separate gather + gene-statistics kernels, 1 fused TMA-fed front kernel, 2 the same kernel fed by
 * its plain loader); "bh_onepass" (1 = the multiple-testing adjustment of small edge lists in one launch [default]);
 * "fit_path" (-1 = choose streaming vs dense-tile pair fits automatically [default], 0 = streaming, 1 = dense);
 * "sort_path" (-1 = K8 orders the edges through p-value buckets up to 2^26 edges and with a CUB radix sort beyond
 * [default], 0 = always the radix sort, 1 = buckets whenever possible);
 * "trace" (1 = print the end time of every launch of the next calls to stderr; implies eager launches);
 * "h2d_pipeline" (1 = the band of a column-major HOST matrix crosses PCIe in row chunks of 12 MB on two copy streams while
 * the compute phase already runs: the TMA loader of the front kernel waits for the chunk that holds its tile [default];
 * 0 = the whole band is copied before the compute phase starts; n > 1 = chunks of n MB);
 * "p2p" (several GPUs: 1 = the all-gather of the p-value / status blocks is a push over peer memory -
 * every rank writes its block straight into the buffers of its peers over NVLink (IPC-mapped with one process per GPU,
 * addressed directly with one process for all GPUs) and raises a flag word there [default; falls back to NCCL when the
 * mappings cannot be made], 0 = ncclAllGather);
 * "tail_spline" (1 = two-category lm on edge lists of 16,384 edges or more per rank reads its t tails from ONE per-run
 * interpolation table, built on the side stream from the same series / continued fractions [default]; 0 = every edge
 * runs its own series / fraction; 2 = the table on every two-category lm call);
 * "host_pack" (1 = a column-major host matrix that is PAGEABLE, or whose network-node rows cover less than 3/4 of the
 * band they span, is not copied as it is: worker threads gather exactly the node rows into pinned staging, row chunk by
 * row chunk, and every finished chunk is sent while the next one is packed and while the front kernel already works
 * on the chunks that arrived; the call then runs on the packed problem, results are identical [default]; 0 = never;
 * 2 = whenever the input is a column-major host matrix; MDEGGS_PACK_THREADS caps the threads [all the process may use]);
 * "zero_copy" (1 = a pinned host assay is read by the gather kernel straight over PCIe instead of being copied first;
 * measured 2 % slower than the DMA copy on this pool's boxes, hence off by default);
 * "rsel_cap" (tests: cap the quantile candidate list so that the overflow path is exercised; 0 = default);
 * "padj_keep_cap" (bytes; the P x E matrix of adjusted values is kept in device scratch for the best-percentile row
 * only up to this size [default 256 MiB], beyond it K8 counts first and re-adjusts the chosen percentile);
 * "count_only" (1 = when no adjusted value of the all-percentile pass is kept, the significant counts come from a
 * single rank pass (step-up rule) instead of the cumulative-min passes [default], 0 = always the three passes). */
int mdeggs_set_option(mdeggs_ctx* ctx, const char* name, int64_t value);
const char* mdeggs_last_error(const mdeggs_ctx* ctx);
const char* mdeggs_strerror(int code);
int mdeggs_abi_version(void);

/* Scalar device-math twins exported for tests (host wrappers that launch a 1-thread kernel):
 * Student-t lower tail as R's pt(), F upper tail as pf(lower.tail=FALSE), and the reference's
 * 2*(1-pt(|t|,df)) expression. n values each. */
int mdeggs_dev_pt(mdeggs_ctx* ctx, const double* x, const double* df, int64_t n, double* out);
int mdeggs_dev_pf_upper(mdeggs_ctx* ctx, const double* f, const double* df1, const double* df2, int64_t n, double* out);
int mdeggs_dev_p_two_sided_ref(mdeggs_ctx* ctx, const double* t, const double* df, int64_t n, double* out);
/* ... and the form the pipeline uses for a whole run (ONE df): the per-run series / continued-fraction tables and,
 * with use_spline, the interpolation table of log P(|T| > t) in s = sqrt(log1p(t^2 / df)) (option "tail_spline"). */
int mdeggs_dev_p_two_sided_tab(mdeggs_ctx* ctx, const double* t, double df, int32_t use_spline, int64_t n, double* out);

#ifdef __cplusplus
}
#endif
#endif /* MDEGGS_H */

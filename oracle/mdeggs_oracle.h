/*
 * mdeggs_oracle.h — CPU restatement (TEST INFRASTRUCTURE ONLY; see mdeggs_oracle.c header).
 * Same problem/result structs as the product ABI (include/mdeggs.h) so tests compare field by field.
 */
#ifndef MDEGGS_ORACLE_H
#define MDEGGS_ORACLE_H
#include <stdint.h>
#include "mdeggs.h"
#ifdef __cplusplus
extern "C" {
#endif

/* mode of mdeggs_oracle_run_omic */
#define MDEGGS_ORACLE_MODE_UNIQUE 0    /* each unique edge fitted once; OpenMP over edges                      */
#define MDEGGS_ORACLE_MODE_REFERENCE 1 /* like R: every (percentile, category, edge) refitted; OpenMP over
                                          percentiles only (what mclapply does, core_functions.R:448-467)     */

int mdeggs_oracle_run_omic(const mdeggs_problem* in, mdeggs_result* out, int mode, int nthreads);
int mdeggs_oracle_threads(void);

double orc_quantile7_sorted(const double* xs, int64_t N, double p);
int orc_seq(double from, double to, double by, double* out, int cap);
double orc_median(const double* x, int64_t n, int64_t stride, double* scratch);
double orc_lbeta(double a, double b);
double orc_pbeta_xy(double x, double y, double a, double b, int lower);
double orc_pbeta(double x, double a, double b, int lower);
double orc_pt(double x, double n);
double orc_p_two_sided_ref(double t, double df);
double orc_p_two_sided_true(double t, double df);
double orc_pf_upper(double f, double d1, double d2);
int orc_fit_lm_interaction(const double* a, const double* b, const double* g01, int n, double* ws, double* coef4,
                           double* t_out, double* p_out);
int orc_fit_aov(const double* a, const double* b, const int32_t* grp, int n, int K, double* ws, double* coef,
                double* f_out, double* p_out);
int orc_fit_rlm(const double* a, const double* b, const double* g01, int n, int tested_coef, int cov_mode,
                double* ws, double* coef4, double* f_out, double* p_out, int* iters_out);
void orc_p_adjust(const double* p, int64_t len, int method, double* out);

#ifdef __cplusplus
}
#endif
#endif

// mdeggs_math.cuh — device-side distribution tails for the pair tests.
//
// Replaces, on the GPU, the R numerics the reference reaches through
//   stats::pt  in  2 * (1 - pt(abs(t), df))                      R/core_functions.R:1072
//   stats::pf(lower.tail = FALSE) inside summary(aov)[["Pr(>F)"]]  R/core_functions.R:966-968
//   stats::pf(lower.tail = FALSE) inside sfsmisc::f.robftest      R/core_functions.R:919-922
// R implements both on top of pbeta (TOMS 708 bratio, not vendored in the reference).  Here the
// regularised incomplete beta is evaluated by its continued fraction, always on the side that yields the
// wanted tail without cancellation (generic modified-Lentz version for the scalar twins; table-driven
// forward recurrence for the per-run batched kernel, see TailConsts below); log Beta uses Stirling-corrected differences so that df up to ~1e6 keeps
// ~1e-13 relative accuracy.  pt()/pf() reproduce R's argument switching (nmath pt.c / pf.c), and the
// two-sided lm p-value reproduces the reference's operation order bit for bit given `val`.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace mdeggs {

__device__ __forceinline__ double stirling_corr(double x) {
  // lgamma(x) - ((x - .5) log x - x + .5 log(2 pi)), x >= 10
  double r = 1.0 / (x * x);
  double s = 1.0 / 156.0;
  s = fma(s, r, -691.0 / 360360.0);
  s = fma(s, r, 1.0 / 1188.0);
  s = fma(s, r, -1.0 / 1680.0);
  s = fma(s, r, 1.0 / 1260.0);
  s = fma(s, r, -1.0 / 360.0);
  s = fma(s, r, 1.0 / 12.0);
  return s / x;
}

__device__ inline double lbeta_dev(double a, double b) {
  double p = fmin(a, b), q = fmax(a, b);
  const double half_log_2pi = 0.918938533204672741780329736406;
  if (p >= 10.0) {
    double corr = stirling_corr(p) + stirling_corr(q) - stirling_corr(p + q);
    return -0.5 * log(q) + half_log_2pi + corr + (p - 0.5) * log(p / (p + q)) + q * log1p(-p / (p + q));
  }
  if (q >= 10.0) {
    double corr = stirling_corr(q) - stirling_corr(p + q);
    return lgamma(p) + corr + p - p * log(p + q) + (q - 0.5) * log1p(-p / (p + q));
  }
  return lgamma(p) + lgamma(q) - lgamma(p + q);
}

// continued fraction of I_x(a,b) (modified Lentz); fast for x < (a+1)/(a+b+2)
__device__ inline double beta_cf_dev(double a, double b, double x) {
  const double tiny = 1e-300, eps = 1e-16;
  double qab = a + b, qap = a + 1.0, qam = a - 1.0;
  double c = 1.0, d = 1.0 - qab * x / qap;
  if (fabs(d) < tiny) d = tiny;
  d = 1.0 / d;
  double h = d;
  for (int m = 1; m <= 100000; ++m) {
    double dm = (double)m, m2 = 2.0 * dm;
    double aa = dm * (b - dm) * x / ((qam + m2) * (a + m2));
    d = 1.0 + aa * d;
    if (fabs(d) < tiny) d = tiny;
    c = 1.0 + aa / c;
    if (fabs(c) < tiny) c = tiny;
    d = 1.0 / d;
    h *= d * c;
    aa = -(a + dm) * (qab + dm) * x / ((a + m2) * (qap + m2));
    d = 1.0 + aa * d;
    if (fabs(d) < tiny) d = tiny;
    c = 1.0 + aa / c;
    if (fabs(c) < tiny) c = tiny;
    d = 1.0 / d;
    double del = d * c;
    h *= del;
    if (fabs(del - 1.0) < eps) break;
  }
  return h;
}

// lower ? I_x(a,b) : 1 - I_x(a,b), with y = 1 - x supplied by the caller (no cancellation in either tail)
__device__ inline double pbeta_xy_dev(double x, double y, double a, double b, bool lower) {
  if (isnan(x) || isnan(y) || isnan(a) || isnan(b)) return nan("");
  if (x <= 0.0) return lower ? 0.0 : 1.0;
  if (y <= 0.0) return lower ? 1.0 : 0.0;
  double lx = (x < 0.5) ? log(x) : log1p(-y);
  double ly = (y < 0.5) ? log(y) : log1p(-x);
  double front = exp(a * lx + b * ly - lbeta_dev(a, b));
  if (x < (a + 1.0) / (a + b + 2.0)) {
    double w = front * beta_cf_dev(a, b, x) / a;
    return lower ? w : 1.0 - w;
  }
  double w1 = front * beta_cf_dev(b, a, y) / b;
  return lower ? 1.0 - w1 : w1;
}

// `val` of nmath/pt.c: the two-sided tail probability P(|T| > |x|)
__device__ inline double pt_val_dev(double x, double n) {
  double nx = 1.0 + (x / n) * x;
  if (n > x * x) {
    double den = n + x * x;
    return pbeta_xy_dev(x * x / den, n / den, 0.5, n / 2.0, false);
  }
  return pbeta_xy_dev(1.0 / nx, (x / n) * x / nx, n / 2.0, 0.5, true);
}

// stats::pt(x, n), lower tail
__device__ inline double pt_dev(double x, double n) {
  if (isnan(x) || isnan(n) || n <= 0.0) return nan("");
  if (isinf(x)) return x < 0 ? 0.0 : 1.0;
  double val = pt_val_dev(x, n) * 0.5;
  if (x <= 0.0) return val;
  return __dadd_rn(__dsub_rn(0.5, val), 0.5);  // R_D_Cval: (0.5 - val + 0.5)
}

// the reference's  2 * (1 - pt(abs(t), df))  (core_functions.R:1072), same operation order, no contraction
__device__ inline double p_two_sided_ref_dev(double t, double df) {
  if (isnan(t)) return nan("");
  double u = pt_dev(fabs(t), df);
  return __dmul_rn(2.0, __dsub_rn(1.0, u));
}

// stats::pf(f, d1, d2, lower.tail = FALSE)
__device__ inline double pf_upper_dev(double f, double d1, double d2) {
  if (isnan(f) || !(d1 > 0.0) || !(d2 > 0.0)) return nan("");
  if (f <= 0.0) return 1.0;
  if (isinf(f)) return 0.0;
  double den = d2 + d1 * f;
  if (d1 * f > d2) return pbeta_xy_dev(d2 / den, d1 * f / den, d2 / 2.0, d1 / 2.0, true);
  return pbeta_xy_dev(d1 * f / den, d2 / den, d1 / 2.0, d2 / 2.0, false);
}

// ---------------------------------------------------------------------------------------------------
// Table-driven tails for a whole run.  Within one omic every edge uses the SAME degrees of freedom, so the
// continued-fraction coefficients of I_x(a,b) depend on the edge only through x:  d_j = x * e_j(a,b).
// k_cf_tables precomputes e_j for both argument orders and log Beta once; per edge the fraction is then
// evaluated by the forward three-term recurrence (2 FMAs per term, one division + convergence test every
// 4 terms) instead of modified Lentz (6 divisions per iteration): ~7x fewer dependent FP64 divisions.
// ---------------------------------------------------------------------------------------------------
constexpr int CF_TABLE_LEN = 1024;

struct TailConsts {
  double a, b;        // (0.5, df/2) for pt, (df1/2, df2/2) for pf
  double lbeta_ab;    // log Beta(a, b)
  const double* e_ab;  // e_j for I_x(a, b),  j = 1..CF_TABLE_LEN
  const double* e_ba;  // e_j for I_x(b, a)
  const double* c_ab;  // c_k = (a + b + k - 1) / (a + k), k = 1..CF_TABLE_LEN: term ratios of the power series of I_x(a, b)
  const double* c_ba;  // ... of I_x(b, a)
  // piecewise Chebyshev table of the log of the two-sided t tail for the run's degrees of freedom (k_tail_spline);
  // null: not built for this run
  const double* sp;
  double sp_inv_h, sp_inv_n;
};
constexpr int CF_TABLE_DOUBLES = 4 * CF_TABLE_LEN + 2;  // e_ab, e_ba, lbeta (+ pad), c_ab, c_ba

// e_1 = -(a+b)/(a+1); e_{2m} = m(b-m)/((a+2m-1)(a+2m)); e_{2m+1} = -(a+m)(a+b+m)/((a+2m)(a+2m+1))
__global__ void k_cf_tables(double a, double b, double* __restrict__ e_ab, double* __restrict__ e_ba,
                            double* __restrict__ lbeta_out, double* __restrict__ c_ab, double* __restrict__ c_ba) {
  for (int j = threadIdx.x + 1; j <= CF_TABLE_LEN; j += blockDim.x) {
    for (int o = 0; o < 2; ++o) {
      double p = o ? b : a, q = o ? a : b;
      double m = (double)(j >> 1), e;
      if (j & 1) e = -(p + m) * (p + q + m) / ((p + 2.0 * m) * (p + 2.0 * m + 1.0));
      else e = m * (q - m) / ((p + 2.0 * m - 1.0) * (p + 2.0 * m));
      (o ? e_ba : e_ab)[j - 1] = e;
      (o ? c_ba : c_ab)[j - 1] = (p + q + (double)(j - 1)) / (p + (double)j);
    }
  }
  if (threadIdx.x == 0) *lbeta_out = lbeta_dev(a, b);
}

// Power series of the incomplete beta on the side below the mode:
//   I_x(a, b) = x^a (1 - x)^b / (a B(a, b)) * sum_{k >= 0} t_k,   t_0 = 1,  t_k = t_{k-1} c_k x,
// all terms positive (no cancellation), one multiply and one add per term on the dependent chain and NO division,
// against 2 FMAs per term + a division every 4 terms for the continued fraction.  The ratio c_k x falls with k, so the
// terms rise to a maximum near k ~ x (a + b) and then decay faster than geometrically: ~x (a + b) + 30 terms.
// `budget` / `ok` as for beta_cf_tab.  Returns the sum.
__device__ inline double beta_series_tab(const double* __restrict__ c, double x, int budget = CF_TABLE_LEN,
                                         bool* ok = nullptr) {
  double term = 1.0, sum = 1.0;
  const int jmax = budget < CF_TABLE_LEN ? budget : CF_TABLE_LEN;
  for (int j = 0; j < jmax; j += 4) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      term *= x * __ldg(c + j + u);
      sum += term;
    }
    if (term <= 1e-18 * sum) return sum;  // (only reached on the decaying side: a rising term is >= t_0 = 1)
  }
  if (ok) *ok = false;
  return -1.0;  // (callers check `ok` or the sign)
}
constexpr double BETA_SERIES_MAX_XAB = 14.0;  // x (a + b) up to which the series is used beyond the mode switch (~70 terms)

// continued fraction value h(x) = 1/(1 + d_1/(1 + d_2/(1 + ...))), d_j = x e_j
// `budget` (a multiple of 4) bounds the number of terms; when it is exhausted and `ok` is given, *ok = false and the
// caller retries later with the full table (k_finish: the slow lanes of a CTA are compacted first, so that a few lanes
// near the mode of the distribution - O(sqrt(max(a, b))) terms - do not hold their whole warps for the duration)
__device__ inline double beta_cf_tab(const double* __restrict__ e, double a, double b, double x,
                                     int budget = CF_TABLE_LEN, bool* ok = nullptr) {
  double Am = 0.0, A = 1.0, Bm = 1.0, B = 1.0;  // (A_0, A_1), (B_0, B_1)
  double f_old = 1.0;
  const int jmax = budget < CF_TABLE_LEN ? budget : CF_TABLE_LEN;
  for (int j = 0; j < jmax; j += 4) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      double d = x * __ldg(e + j + u);
      double An = fma(d, Am, A), Bn = fma(d, Bm, B);
      Am = A;
      Bm = B;
      A = An;
      B = Bn;
    }
    double r = 1.0 / B;
    A *= r;
    Am *= r;
    Bm *= r;
    B = 1.0;
    if (fabs(A - f_old) <= 4e-16 * fabs(A)) return A;
    f_old = A;
  }
  if (ok) {
    *ok = false;
    return 0.0;
  }
  return beta_cf_dev(a, b, x);  // table exhausted (never seen for |t| < 40, df < 1e6): generic Lentz
}

// pbeta_xy for the run's (a, b) pair; `swapped` = the caller's first parameter is tc.b.
// Where x (a + b) <= 14 the power series replaces the continued fractions: below the mode switch of pbeta
// (x < (a + 1) / (a + b + 2)) and on the stretch beyond it where the complementary fraction needs O(sqrt(max(a, b)))
// terms (|t| up to ~5 at 2,000 degrees of freedom), as long as the wanted tail 1 - w keeps its relative accuracy.
// `strict`: the series is only taken where it yields the wanted tail directly (k_tail_spline: the nodes of the table
// are worth the long fraction; 1 - w from the series carries ~1e-16 / (1 - w) of relative noise).
__device__ inline double pbeta_xy_tab(const TailConsts& tc, bool swapped, double x, double y, bool lower,
                                      int budget = CF_TABLE_LEN, bool* ok = nullptr, bool strict = false) {
  const double a = swapped ? tc.b : tc.a, b = swapped ? tc.a : tc.b;
  if (isnan(x) || isnan(y)) return nan("");
  if (x <= 0.0) return lower ? 0.0 : 1.0;
  if (y <= 0.0) return lower ? 1.0 : 0.0;
  double lx = (x < 0.5) ? log(x) : log1p(-y);
  double ly = (y < 0.5) ? log(y) : log1p(-x);
  double front = exp(a * lx + b * ly - tc.lbeta_ab);
  const bool below = x < (a + 1.0) / (a + b + 2.0);
  if (x * (a + b) <= BETA_SERIES_MAX_XAB && (!strict || below || lower)) {  // (beyond: more terms than the fractions)
    bool conv = true;
    const double sum = beta_series_tab(swapped ? tc.c_ba : tc.c_ab, x, budget, &conv);
    if (conv) {
      const double w = front * sum / a;
      if (below || lower || 1.0 - w >= 1e-5) return lower ? w : 1.0 - w;
    } else if (ok) {  // out of budget: the caller retries with the full table
      *ok = false;
      return 0.0;
    }
  }
  if (below) {
    double w = front * beta_cf_tab(swapped ? tc.e_ba : tc.e_ab, a, b, x, budget, ok) / a;
    return lower ? w : 1.0 - w;
  }
  double w1 = front * beta_cf_tab(swapped ? tc.e_ab : tc.e_ba, b, a, y, budget, ok) / b;
  return lower ? 1.0 - w1 : w1;
}

// `val` of nmath/pt.c for the run's degrees of freedom (tc = {0.5, n/2}): P(|T| > x), x >= 0 finite
__device__ inline double pt_val_tab(const TailConsts& tc, double x, double n, int budget = CF_TABLE_LEN,
                                    bool* ok = nullptr, bool strict = false) {
  if (n > x * x) {
    double den = n + x * x;
    return pbeta_xy_tab(tc, false, x * x / den, n / den, false, budget, ok, strict);
  }
  double nx = 1.0 + (x / n) * x;
  return pbeta_xy_tab(tc, true, 1.0 / nx, (x / n) * x / nx, true, budget, ok, strict);
}

// ---------------------------------------------------------------------------------------------------
// One table per run instead of one series / continued fraction per edge.  Every edge of a run has the same degrees of
// freedom n, so the two-sided tail val(t) = P(|T| > t) = I_x(n/2, 1/2), x = n / (n + t^2), is ONE smooth function of t.
// In the variable s = sqrt(log1p(t^2 / n)) (x = exp(-s^2)) its logarithm g(s) = log val is analytic at s = 0
// (val = 1 - c t + ..., t = sqrt(n expm1(s^2))) and close to the parabola -(n/2) s^2 in the far tail, so a piecewise
// Chebyshev interpolant of degree 7 on TS_M uniform intervals of [0, s_max], (n/2) s_max^2 = 680 (val ~ 1e-295),
// reproduces g to ~1e-13 for every n (the interval is ~1/50 of the scale 1/sqrt(n) on which g bends).  k_tail_spline
// evaluates the 8 Chebyshev nodes of every interval with the run's series / continued fractions (16,384 evaluations,
// once per run, on the side stream) and stores the Chebyshev coefficients; per edge the tail is then log1p, sqrt, a
// Clenshaw recurrence of 8 terms and exp - no loop, no division, no divergent lanes.  Arguments beyond s_max (val
// below ~1e-295) keep the fraction.  The result feeds the same  2 * (1 - (0.5 - val / 2 + 0.5))  composition.
// ---------------------------------------------------------------------------------------------------
constexpr int TS_M = 2048;
constexpr int TS_C = 8;  // coefficients per interval
inline double tail_spline_smax(double n) {
  const double s2 = 1360.0 / n;
  return sqrt(s2 < 680.0 ? s2 : 680.0);  // (t^2 = n expm1(s^2) stays finite)
}
__global__ void __launch_bounds__(256) k_tail_spline(TailConsts tc, const double* __restrict__ lbeta_p, double n,
                                                     double h, double* __restrict__ sp) {
  tc.lbeta_ab = *lbeta_p;
  tc.sp = nullptr;
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;  // (TS_M * TS_C threads: whole groups of 8 lanes)
  const int i = gid / TS_C, k = gid % TS_C;
  const double theta = cospi(((double)k + 0.5) / (double)TS_C);
  const double sv = ((double)i + 0.5 * (theta + 1.0)) * h;
  const double t = sqrt(n * expm1(sv * sv));
  const double f = log(pt_val_tab(tc, t, n, CF_TABLE_LEN, nullptr, true));
  // c_j = (2 / 8) sum_k f_k cos(j pi (k + 1/2) / 8), c_0 halved: lane j of the group gathers the 8 node values
  double c = 0.0;
  const int lane = threadIdx.x & 31, base = lane & ~(TS_C - 1);
#pragma unroll
  for (int kk = 0; kk < TS_C; ++kk) {
    const double fk = __shfl_sync(0xffffffffu, f, base + kk);
    c += fk * cospi((double)k * ((double)kk + 0.5) / (double)TS_C);
  }
  c *= (k == 0 ? 1.0 : 2.0) / (double)TS_C;
  if (i < TS_M) sp[(size_t)i * TS_C + k] = c;
}
// val from the table; *in_range = false (and nothing else) when |t| lies beyond it
__device__ __forceinline__ double pt_val_spline(const TailConsts& tc, double x, bool* in_range) {
  const double s = sqrt(log1p(x * x * tc.sp_inv_n));
  const double u = s * tc.sp_inv_h;
  if (!(u < (double)TS_M)) {  // (also inf / NaN)
    *in_range = false;
    return 0.0;
  }
  const int i = (int)u;
  const double th = 2.0 * (u - (double)i) - 1.0, th2 = th + th;
  const double2* __restrict__ c2 = reinterpret_cast<const double2*>(tc.sp + (size_t)i * TS_C);
  const double2 c01 = __ldg(c2), c23 = __ldg(c2 + 1), c45 = __ldg(c2 + 2), c67 = __ldg(c2 + 3);
  double b1 = c67.y, b0 = fma(th2, b1, c67.x), t2;  // Clenshaw: b_k = c_k + 2 th b_{k+1} - b_{k+2}
  t2 = fma(th2, b0, c45.y - b1); b1 = b0; b0 = t2;
  t2 = fma(th2, b0, c45.x - b1); b1 = b0; b0 = t2;
  t2 = fma(th2, b0, c23.y - b1); b1 = b0; b0 = t2;
  t2 = fma(th2, b0, c23.x - b1); b1 = b0; b0 = t2;
  t2 = fma(th2, b0, c01.y - b1); b1 = b0; b0 = t2;
  const double g = fma(th, b0, c01.x - b1);
  *in_range = true;
  return exp(g);
}

// 2 * (1 - pt(|t|, n)) with tc = {0.5, n/2}: same composition as p_two_sided_ref_dev
__device__ inline double p_two_sided_ref_tab(const TailConsts& tc, double t, double n, int budget = CF_TABLE_LEN,
                                             bool* ok = nullptr) {
  if (isnan(t)) return nan("");
  double x = fabs(t), u;
  if (isinf(x) || x > 1e150) {  // (t^2 overflows; the tail is far below the 1.1e-16 grid of u for every n >= 1)
    u = 1.0;
  } else {
    double val = 0.0;
    bool tabled = false;
    if (tc.sp) val = pt_val_spline(tc, x, &tabled);
    if (!tabled) val = pt_val_tab(tc, x, n, budget, ok);
    if (tabled && val > 1.0) val = 1.0;  // (the interpolant may exceed 1 by an ulp at t -> 0)
    val *= 0.5;
    u = (x <= 0.0) ? val : __dadd_rn(__dsub_rn(0.5, val), 0.5);
  }
  return __dmul_rn(2.0, __dsub_rn(1.0, u));
}

// pf(f, 2, d2, lower.tail = FALSE) in closed form: with d1 = 2 the incomplete beta I_x(d2/2, 1) is x^(d2/2),
// x = d2 / (d2 + 2 f).  Three categories (the interaction F of aov has d1 = K - 1 = 2) therefore need no continued
// fraction at all.  log1p keeps full precision when x is close to 1 (small F), log(x) when it is not.
__device__ __forceinline__ double pf_upper_d1_2(double f, double d2) {
  const double den = d2 + 2.0 * f;
  const double x = d2 / den, y = 2.0 * f / den;  // x + y == 1
  const double lg = x < 0.5 ? log(x) : log1p(-y);
  return exp(0.5 * d2 * lg);
}

// pf(f, d1, d2, lower.tail = FALSE) with tc = {d1/2, d2/2}
__device__ inline double pf_upper_tab(const TailConsts& tc, double f, double d1, double d2, int budget = CF_TABLE_LEN,
                                      bool* ok = nullptr) {
  if (isnan(f)) return nan("");
  if (f <= 0.0) return 1.0;
  if (isinf(f)) return 0.0;
  if (d1 == 2.0) return pf_upper_d1_2(f, d2);
  double den = d2 + d1 * f;
  if (d1 * f > d2) return pbeta_xy_tab(tc, true, d2 / den, d1 * f / den, true, budget, ok);
  return pbeta_xy_tab(tc, false, d1 * f / den, d2 / den, false, budget, ok);
}

}  // namespace mdeggs

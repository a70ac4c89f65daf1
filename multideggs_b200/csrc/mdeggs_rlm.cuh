// mdeggs_rlm.cuh — K6: batched Huber IRLS for regression_method = "rlm", two categories.
//
// Replaces, per edge,  MASS::rlm(B ~ A * metadata)  followed by  sfsmisc::f.robftest(fit, var = 3)
// (R/core_functions.R:914-925).  MASS defaults: psi.huber k = 1.345, init = "ls", scale.est = "MAD"
// (median|r| / 0.6745 re-estimated every iteration), maxit = 20, acc = 1e-4 on irls.delta(resid).
//
// One warp per edge.  The 4-parameter weighted fit of  B ~ A * g  decomposes into two independent weighted
// simple regressions (one per category); only the MAD scale couples them.  Rows A and B are staged once in
// shared memory, residuals live in shared memory, so HBM/L2 bytes are paid once per edge and each IRLS
// iteration is: one pass for the Huber weights + 5 weighted moments per category, one pass for the new
// residuals + convergence norm, and one exact order statistic (median of |r|, 63-step key bisection).
// The Wald test needs only closed-form diagonal entries of (X'X)^-1 / (X'WX)^-1 in the same decomposition.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "mdeggs.h"
#include "mdeggs_kernels.cuh"
#include "mdeggs_front.cuh"  // warp_select_dbl: exact selection on doubles held in registers

namespace mdeggs {

// 1 / x for a positive normal double: hardware seed (MUFU.RCP64H, ~20 bits) + two Newton steps (4 FMAs), within an ulp or
// two of the correctly rounded quotient; the IEEE division costs ~4x as many instructions and sat on 19 % of the IRLS
// kernel's issue slots (one per sample and iteration, for the Huber weight k s / |r|)
__device__ __forceinline__ double rcp_newton(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  e = fma(-x, y, 1.0);
  return fma(y, e, y);
}

struct GroupFit {
  double W, ma, mb, sxx, sxy, slope, icpt;
};

// median of |r| over the real (non-padding) samples: exact order statistic (|r| >= 0, so the raw IEEE bit
// pattern is already order preserving); narrowing bisection + in-warp ranking (warp_select)
__device__ __forceinline__ double warp_median_abs(const double* r, const SegLayout& L, int lane,
                                                  unsigned long long* scratch32) {
  const int n = L.n, c0 = L.cnt[0], o0 = L.off[0], o1 = L.off[1];
  auto key_at = [&](int i) {
    int idx = i < c0 ? o0 + i : o1 + (i - c0);
    return (unsigned long long)__double_as_longlong(fabs(r[idx]));
  };
  unsigned long long kmin = ~0ull, kmax = 0ull;
  for (int i = lane; i < n; i += 32) {
    unsigned long long k = key_at(i);
    kmin = k < kmin ? k : kmin;
    kmax = k > kmax ? k : kmax;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    unsigned long long a = __shfl_xor_sync(0xffffffffu, kmin, o), b = __shfl_xor_sync(0xffffffffu, kmax, o);
    kmin = a < kmin ? a : kmin;
    kmax = b > kmax ? b : kmax;
  }
  if (n & 1) return __longlong_as_double((long long)warp_select(key_at, n, n / 2, kmin, kmax, n, scratch32, lane));
  unsigned long long k1 = warp_select(key_at, n, n / 2 - 1, kmin, kmax, n, scratch32, lane);
  int c_le = 0;
  unsigned long long nxt = ~0ull;
  for (int i = lane; i < n; i += 32) {
    unsigned long long kk = key_at(i);
    c_le += (kk <= k1);
    if (kk > k1 && kk < nxt) nxt = kk;
  }
  c_le = __reduce_add_sync(0xffffffffu, c_le);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    unsigned long long other = __shfl_xor_sync(0xffffffffu, nxt, o);
    nxt = other < nxt ? other : nxt;
  }
  unsigned long long k2 = (c_le > n / 2) ? k1 : nxt;
  return __ddiv_rn(__dadd_rn(__longlong_as_double((long long)k1), __longlong_as_double((long long)k2)), 2.0);
}

// The IRLS itself and the Wald test, one warp: `pa` / `pb` hold the pair's samples category by category as laid out by
// L (off[k], cnt[k]; L.n of them in total), `r` receives the residuals, ca / cb / sxa are the unweighted per-category
// means of A and B and the centred sum of squares of A.  Shared by k_fit_rlm (rows of the run's layout) and by
// k_fit_na_pairs (the complete cases of one pair, mdeggs_napair.cuh).
__device__ __forceinline__ void rlm_irls_warp(const double* pa, const double* pb, double* r, const SegLayout& L,
                                              const double (&ca)[2], const double (&cb)[2], const double (&sxa)[2],
                                              int tested_coef, int cov_mode, unsigned long long* s_cand32, int lane,
                                              double* F_out, int* status_out, int* iters_out, double (&beta)[4]) {
  const double kh = 1.345;
  const int n = L.n;
  GroupFit gf[2];
  // ---- init = "ls": ordinary least squares residuals ----
  for (int k = 0; k < 2; ++k) {
    double acc = 0.0;
    for (int i = L.off[k] + lane; i < L.off[k] + L.cnt[k]; i += 32) acc = fma(pa[i] - ca[k], pb[i] - cb[k], acc);
    acc = warp_sum(acc);
    gf[k].W = (double)L.cnt[k];
    gf[k].ma = ca[k];
    gf[k].mb = cb[k];
    gf[k].sxx = sxa[k];
    gf[k].sxy = acc;
    gf[k].slope = acc / gf[k].sxx;
    gf[k].icpt = cb[k] - gf[k].slope * ca[k];
    for (int i = L.off[k] + lane; i < L.off[k] + L.cnt[k]; i += 32) r[i] = pb[i] - gf[k].icpt - gf[k].slope * pa[i];
  }
  __syncwarp();
  double scale = 0.0;
  bool done = false, zero_scale = false;
  int it = 0;
  for (it = 1; it <= 20; ++it) {
    scale = warp_median_abs(r, L, lane, s_cand32) / 0.6745;
    if (scale == 0.0) {
      zero_scale = true;
      done = true;
      break;
    }
    double diff = 0.0, old_ss = 0.0;
    for (int k = 0; k < 2; ++k) {
      double sw = 0.0, swa = 0.0, swb = 0.0, swaa = 0.0, swab = 0.0;
      for (int i = L.off[k] + lane; i < L.off[k] + L.cnt[k]; i += 32) {
        double u = fabs(r[i] / scale);
        double w = (u <= kh) ? 1.0 : kh / u;  // psi.huber weight pmin(1, k/|u|)
        double da = pa[i] - ca[k], db = pb[i] - cb[k];
        sw += w;
        swa = fma(w, da, swa);
        swb = fma(w, db, swb);
        swaa = fma(w * da, da, swaa);
        swab = fma(w * da, db, swab);
      }
      sw = warp_sum(sw);
      swa = warp_sum(swa);
      swb = warp_sum(swb);
      swaa = warp_sum(swaa);
      swab = warp_sum(swab);
      GroupFit& g = gf[k];
      g.W = sw;
      double dma = swa / sw, dmb = swb / sw;
      g.ma = ca[k] + dma;
      g.mb = cb[k] + dmb;
      g.sxx = swaa - swa * dma;
      g.sxy = swab - swa * dmb;
      g.slope = g.sxy / g.sxx;
      g.icpt = g.mb - g.slope * g.ma;
      for (int i = L.off[k] + lane; i < L.off[k] + L.cnt[k]; i += 32) {
        double ro = r[i];
        double rn = pb[i] - g.icpt - g.slope * pa[i];
        double d = ro - rn;
        diff = fma(d, d, diff);
        old_ss = fma(ro, ro, old_ss);
        r[i] = rn;
      }
    }
    __syncwarp();
    diff = warp_sum(diff);
    old_ss = warp_sum(old_ss);
    double conv = sqrt(diff / fmax(old_ss, 1e-20));  // irls.delta
    if (conv <= 1e-4) {
      done = true;
      break;
    }
  }
  if (it > 20) it = 20;
  *iters_out = it;
  if (zero_scale) {
    *F_out = nan("");
    *status_out = MDEGGS_EDGE_ZERO_SCALE;
    return;
  }
  // ---- summary.rlm + Wald F of coefficient `tested_coef` (f.robftest) ----
  double S = 0.0, npp = 0.0, sumw = 0.0;
  double W2[2] = {0, 0}, wa2[2] = {0, 0}, waa2[2] = {0, 0};
  for (int k = 0; k < 2; ++k) {
    double sw = 0.0, swa = 0.0, swaa = 0.0;
    for (int i = L.off[k] + lane; i < L.off[k] + L.cnt[k]; i += 32) {
      double u = fabs(r[i] / scale);
      double w = (u <= kh) ? 1.0 : kh / u;
      double rw = r[i] * w;
      S = fma(rw, rw, S);
      npp += (u <= kh) ? 1.0 : 0.0;
      double da = pa[i] - ca[k];
      sw += w;
      swa = fma(w, da, swa);
      swaa = fma(w * da, da, swaa);
    }
    W2[k] = warp_sum(sw);
    wa2[k] = warp_sum(swa);
    waa2[k] = warp_sum(swaa);
    sumw += W2[k];
  }
  S = warp_sum(S);
  npp = warp_sum(npp);
  const double rdf = (double)(n - 4);
  S /= rdf;
  const double mn = npp / (double)n;
  const double var_pp = (double)n * mn * (1.0 - mn) / (double)(n - 1);
  const double kappa = 1.0 + 4.0 * var_pp / ((double)n * mn * mn);
  const double stddev = sqrt(S) * (kappa / mn);
  double v_icpt[2], v_slope[2];
  double cov_scale = 1.0;
  if (cov_mode == MDEGGS_RLM_COV_XTWX) {
    cov_scale = sumw / (double)n;  // X~ = X sqrt(w / mean(w))  =>  cov.unscaled = mean(w) (X'WX)^-1
    for (int k = 0; k < 2; ++k) {
      double dm = wa2[k] / W2[k];
      double mk = ca[k] + dm, sx = waa2[k] - wa2[k] * dm;
      v_icpt[k] = 1.0 / W2[k] + mk * mk / sx;
      v_slope[k] = 1.0 / sx;
    }
  } else {
    for (int k = 0; k < 2; ++k) {
      double sx = sxa[k];
      v_icpt[k] = 1.0 / (double)L.cnt[k] + ca[k] * ca[k] / sx;
      v_slope[k] = 1.0 / sx;
    }
  }
  beta[0] = gf[0].icpt;
  beta[1] = gf[0].slope;
  beta[2] = gf[1].icpt - gf[0].icpt;
  beta[3] = gf[1].slope - gf[0].slope;
  double cov[4] = {v_icpt[0], v_slope[0], v_icpt[0] + v_icpt[1], v_slope[0] + v_slope[1]};
  const int v = tested_coef - 1;
  const double var_v = stddev * stddev * cov_scale * cov[v];
  *F_out = beta[v] * beta[v] / var_v;
  *status_out = done ? MDEGGS_EDGE_OK : MDEGGS_EDGE_NOT_CONVERGED;
}

template <bool CACHE_AB>
__global__ void __launch_bounds__(256) k_fit_rlm(const double* __restrict__ D, SegLayout L,
                                                 const int32_t* __restrict__ ea, const int32_t* __restrict__ eb,
                                                 const double* __restrict__ mean, const double* __restrict__ sxx,
                                                 const uint8_t* __restrict__ bad, int64_t e_lo, int64_t e_hi,
                                                 int tested_coef, int cov_mode, double* __restrict__ stat,
                                                 int32_t* __restrict__ status, double* __restrict__ coef,
                                                 int32_t* __restrict__ iters, const int32_t* __restrict__ err) {
  extern __shared__ double smem[];
  if (*err) return;
  __shared__ unsigned long long s_cand[8][32];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const int per_warp = CACHE_AB ? 3 * L.npad : L.npad;
  double* r = smem + (size_t)wib * per_warp;
  double* sa = CACHE_AB ? r + L.npad : nullptr;
  double* sb = CACHE_AB ? r + 2 * L.npad : nullptr;
  for (int64_t e = e_lo + (int64_t)blockIdx.x * wpb + wib; e < e_hi; e += (int64_t)gridDim.x * wpb) {
    const int a = ea[e], b = eb[e];
    auto put = [&](double st_v, int code, int it) {
      if (lane == 0) {
        stat[e] = st_v;
        status[e] = code;
        iters[e] = it;
        if (coef)
          for (int j = 0; j < 4; ++j) coef[(size_t)e * 4 + j] = nan("");
      }
    };
    if (a == b || bad[a] || bad[b]) {
      bool self = (a == b) && !bad[a];
      put(self ? 0.0 : nan(""), self ? MDEGGS_EDGE_SELFLOOP : MDEGGS_EDGE_NONFINITE, 0);
      continue;
    }
    const double* ga = D + (size_t)a * L.npad;
    const double* gb = D + (size_t)b * L.npad;
    if (CACHE_AB) {
      for (int i = lane; i < L.npad; i += 32) {
        sa[i] = __ldg(ga + i);
        sb[i] = __ldg(gb + i);
      }
      __syncwarp();
    }
    const double* pa = CACHE_AB ? sa : ga;
    const double* pb = CACHE_AB ? sb : gb;
    // unweighted centring constants (from K2) keep every weighted moment well conditioned
    double ca[2], cb[2];
    bool singular = false;
    for (int k = 0; k < 2; ++k) {
      ca[k] = mean[(size_t)a * 2 + k];
      cb[k] = mean[(size_t)b * 2 + k];
      double s = sxx[(size_t)a * 2 + k];
      if (L.cnt[k] < 2 || !(s > 1e-14 * (s + (double)L.cnt[k] * ca[k] * ca[k]))) singular = true;
    }
    if (singular) {  // rlm(): "'x' is singular: singular fits are not implemented in 'rlm'"
      put(nan(""), MDEGGS_EDGE_SINGULAR, 0);
      continue;
    }
    if (sxx[(size_t)b * 2] == 0.0 && sxx[(size_t)b * 2 + 1] == 0.0) {  // constant response: exact fit
      put(0.0, MDEGGS_EDGE_SELFLOOP, 0);
      continue;
    }
    double F, beta[4];
    int st_code, it;
    {
      const double sxa[2] = {sxx[(size_t)a * 2], sxx[(size_t)a * 2 + 1]};
      rlm_irls_warp(pa, pb, r, L, ca, cb, sxa, tested_coef, cov_mode, s_cand[wib], lane, &F, &st_code, &it, beta);
    }
    if (st_code == MDEGGS_EDGE_ZERO_SCALE) {
      put(nan(""), MDEGGS_EDGE_ZERO_SCALE, it);
      continue;
    }
    const bool done = st_code == MDEGGS_EDGE_OK;
    if (lane == 0) {
      stat[e] = F;
      status[e] = done ? MDEGGS_EDGE_OK : MDEGGS_EDGE_NOT_CONVERGED;
      iters[e] = it;
      if (coef)
        for (int j = 0; j < 4; ++j) coef[(size_t)e * 4 + j] = beta[j];
    }
    __syncwarp();
  }
}

// Register-resident variant for npad <= 32 * ITEMS (n up to ~1,000): rows A and B are staged once in shared memory,
// |residual| keys live in registers, the MAD median is a register-only narrowing selection (warp_select_reg), and the
// irls.delta convergence norm needs NO pass: with r_old - r_new = (i_new - i_old) + (s_new - s_old) a inside a category,
//   sum (r_old - r_new)^2 = n di^2 + 2 di ds sum(a) + ds^2 sum(a^2),   sum r_old^2 = Syy - 2 s Sxy + s^2 Sxx + n (mb - i - s ma)^2
// are closed forms in the unweighted per-gene moments of K2.  Per IRLS iteration: one pass for the residual keys, the
// selection, one pass for the Huber weights + 5 weighted moments per category.
// warps per CTA of the register-resident kernel: as many as the register file holds (one CTA per SM at 32 items)
template <int ITEMS>
struct RlmRegWarps {
  static constexpr int value = ITEMS >= 32 ? 12 : (ITEMS >= 16 ? 12 : (ITEMS >= 8 ? 12 : 16));
};

template <int ITEMS>
__global__ void __launch_bounds__(RlmRegWarps<ITEMS>::value * 32) k_fit_rlm_reg(const double* __restrict__ D, SegLayout L,
                                                     const int32_t* __restrict__ ea, const int32_t* __restrict__ eb,
                                                     const double* __restrict__ mean, const double* __restrict__ sxx,
                                                     const uint8_t* __restrict__ bad, int64_t e_lo, int64_t e_hi,
                                                     int tested_coef, int cov_mode, double* __restrict__ stat,
                                                     int32_t* __restrict__ status, double* __restrict__ coef,
                                                     int32_t* __restrict__ iters, const int32_t* __restrict__ err) {
  extern __shared__ double smem[];  // per warp: A row, B row (npad doubles each)
  __shared__ unsigned long long s_cand[RlmRegWarps<ITEMS>::value][32];
  __shared__ unsigned int s_ncand[RlmRegWarps<ITEMS>::value];
  if (*err) return;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  double* sa = smem + (size_t)wib * 2 * L.npad;
  double* sb = sa + L.npad;
  const double kh = 1.345;
  const int n = L.n, c0 = L.cnt[0], c1 = L.cnt[1], o1 = L.off[1];
  // per-lane validity / category of the positions lane + 32 j (same for every edge)
  unsigned vmask = 0, cmask = 0;
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    const int pos = lane + 32 * j;
    const bool in1 = pos >= o1;
    const bool valid = in1 ? (pos < o1 + c1) : (pos < c0);
    vmask |= (unsigned)valid << j;
    cmask |= (unsigned)in1 << j;
  }
  for (int64_t e = e_lo + (int64_t)blockIdx.x * wpb + wib; e < e_hi; e += (int64_t)gridDim.x * wpb) {
    const int a = ea[e], b = eb[e];
    auto put = [&](double st_v, int code, int it) {
      if (lane == 0) {
        stat[e] = st_v;
        status[e] = code;
        iters[e] = it;
        if (coef)
          for (int j = 0; j < 4; ++j) coef[(size_t)e * 4 + j] = nan("");
      }
    };
    if (a == b || bad[a] || bad[b]) {
      bool self = (a == b) && !bad[a];
      put(self ? 0.0 : nan(""), self ? MDEGGS_EDGE_SELFLOOP : MDEGGS_EDGE_NONFINITE, 0);
      continue;
    }
    double ma[2], mb[2], sxa[2], syb[2];
    bool singular = false;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      ma[k] = mean[(size_t)a * 2 + k];
      mb[k] = mean[(size_t)b * 2 + k];
      sxa[k] = sxx[(size_t)a * 2 + k];
      syb[k] = sxx[(size_t)b * 2 + k];
      if (L.cnt[k] < 2 || !(sxa[k] > 1e-14 * (sxa[k] + (double)L.cnt[k] * ma[k] * ma[k]))) singular = true;
    }
    if (singular) {  // rlm(): "'x' is singular: singular fits are not implemented in 'rlm'"
      put(nan(""), MDEGGS_EDGE_SINGULAR, 0);
      continue;
    }
    if (syb[0] == 0.0 && syb[1] == 0.0) {  // constant response: exact fit
      put(0.0, MDEGGS_EDGE_SELFLOOP, 0);
      continue;
    }
    __syncwarp();
    {
      const double* ga = D + (size_t)a * L.npad;
      const double* gb = D + (size_t)b * L.npad;
      for (int i = lane; i < L.npad; i += 32) {
        sa[i] = __ldg(ga + i);
        sb[i] = __ldg(gb + i);
      }
    }
    __syncwarp();
    const double nk[2] = {(double)c0, (double)c1};
    // ---- init = "ls" ----
    double sxy[2];
    {
      double acc0 = 0.0, acc1 = 0.0;
#pragma unroll
      for (int j = 0; j < ITEMS; ++j) {
        if (!(vmask >> j & 1)) continue;
        const int pos = lane + 32 * j;
        const int c = cmask >> j & 1;
        const double t = (sa[pos] - ma[c]) * (sb[pos] - mb[c]);
        if (c) acc1 += t;
        else acc0 += t;
      }
      sxy[0] = warp_sum(acc0);
      sxy[1] = warp_sum(acc1);
    }
    double slope[2], icpt[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      slope[k] = sxy[k] / sxa[k];
      icpt[k] = mb[k] - slope[k] * ma[k];
    }
    // |residuals| of this lane's positions (NaN on padding: it compares false everywhere and sorts above everything)
    double ares[ITEMS];
    auto residual_keys = [&]() {  // |b - icpt_c - slope_c a|
#pragma unroll
      for (int j = 0; j < ITEMS; ++j) {
        double r = nan("");
        if (vmask >> j & 1) {
          const int pos = lane + 32 * j;
          const int c = cmask >> j & 1;
          r = fabs(sb[pos] - icpt[c] - slope[c] * sa[pos]);
        }
        ares[j] = r;
      }
    };
    // `guess` > 0: expected median of |r| (previous iteration's, or 0.6745 sd of the residuals at the start); the exact
    // selection (warp_select_dbl, mdeggs_front.cuh: FP64 compares, ballot quick-select among the last <= 32 candidates)
    // starts from the bracket guess * (1 -+ rel)
    auto median_abs = [&](double guess, double rel) -> double {
      const bool have = guess > 0.0 && isfinite(guess);
      double k2 = 0.0;
      const double k1 = warp_select_dbl<ITEMS>(ares, (n - 1) / 2, n, reinterpret_cast<double*>(s_cand[wib]), s_ncand + wib,
                                               lane, have, guess * (1.0 - rel), guess * (1.0 + rel), !(n & 1), &k2);
      return (n & 1) ? k1 : __ddiv_rn(__dadd_rn(k1, k2), 2.0);
    };
    double med_prev = 0.0;
    double scale = 0.0;
    bool done = false, zero_scale = false;
    int it = 0;
    for (it = 1; it <= 20; ++it) {
      residual_keys();
      {
        double guess = med_prev, rel = 0.1;        if (it == 1) {  // residual sd under the least-squares start (closed form) -> MAD of a normal sample
          double ss = 0.0;
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const double off = mb[k] - icpt[k] - slope[k] * ma[k];
            ss += syb[k] - 2.0 * slope[k] * sxy[k] + slope[k] * slope[k] * sxa[k] + nk[k] * off * off;
          }
          guess = 0.6745 * sqrt(fmax(ss, 0.0) / (double)n);
          rel = 0.25;
        }
        med_prev = median_abs(guess, rel);
      }
      scale = med_prev / 0.6745;
      if (scale == 0.0) {
        zero_scale = true;
        done = true;
        break;
      }
      // sum of squared residuals under the current parameters (closed form, unweighted moments)
      double old_ss = 0.0;
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const double off = mb[k] - icpt[k] - slope[k] * ma[k];
        old_ss += syb[k] - 2.0 * slope[k] * sxy[k] + slope[k] * slope[k] * sxa[k] + nk[k] * off * off;
      }
      double acc[2][5] = {{0, 0, 0, 0, 0}, {0, 0, 0, 0, 0}};
      const double ks = kh * scale;
#pragma unroll
      for (int j = 0; j < ITEMS; ++j) {
        if (!(vmask >> j & 1)) continue;
        const int pos = lane + 32 * j;
        const int c = cmask >> j & 1;
        // psi.huber weight pmin(1, k / |r / s|) with ONE division per sample: k s / |r|
        const double ar = ares[j];
        const double w = (ar <= ks) ? 1.0 : ks * rcp_newton(ar);
        const double da = sa[pos] - ma[c], db = sb[pos] - mb[c];
        const double wda = w * da;
        if (c) {
          acc[1][0] += w;
          acc[1][1] += wda;
          acc[1][2] = fma(w, db, acc[1][2]);
          acc[1][3] = fma(wda, da, acc[1][3]);
          acc[1][4] = fma(wda, db, acc[1][4]);
        } else {
          acc[0][0] += w;
          acc[0][1] += wda;
          acc[0][2] = fma(w, db, acc[0][2]);
          acc[0][3] = fma(wda, da, acc[0][3]);
          acc[0][4] = fma(wda, db, acc[0][4]);
        }
      }
      double diff = 0.0;
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const double sw = warp_sum(acc[k][0]), swa = warp_sum(acc[k][1]), swb = warp_sum(acc[k][2]);
        const double swaa = warp_sum(acc[k][3]), swab = warp_sum(acc[k][4]);
        const double dma = swa / sw, dmb = swb / sw;
        const double wsxx = swaa - swa * dma, wsxy = swab - swa * dmb;
        const double s_new = wsxy / wsxx;
        const double i_new = (mb[k] + dmb) - s_new * (ma[k] + dma);
        const double di = i_new - icpt[k], ds = s_new - slope[k];
        const double suma = nk[k] * ma[k], suma2 = sxa[k] + nk[k] * ma[k] * ma[k];
        diff += nk[k] * di * di + 2.0 * di * ds * suma + ds * ds * suma2;
        slope[k] = s_new;
        icpt[k] = i_new;
      }
      const double conv = sqrt(fmax(diff, 0.0) / fmax(old_ss, 1e-20));  // irls.delta
      if (conv <= 1e-4) {
        done = true;
        break;
      }
    }
    if (it > 20) it = 20;
    if (zero_scale) {
      put(nan(""), MDEGGS_EDGE_ZERO_SCALE, it);
      continue;
    }
    // ---- summary.rlm + Wald F of coefficient `tested_coef` (f.robftest) ----
    residual_keys();  // final residuals, scale of the last iteration
    double S = 0.0, npp = 0.0;
    double fw[2][3] = {{0, 0, 0}, {0, 0, 0}};
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
      if (!(vmask >> j & 1)) continue;
      const int pos = lane + 32 * j;
      const int c = cmask >> j & 1;
      const double ar = ares[j];
      const double u = ar / scale;
      const double w = (u <= kh) ? 1.0 : (kh * scale) * rcp_newton(ar);
      const double rw = ar * w;
      S = fma(rw, rw, S);
      npp += (u <= kh) ? 1.0 : 0.0;
      const double da = sa[pos] - ma[c];
      const double wda = w * da;
      if (c) {
        fw[1][0] += w;
        fw[1][1] += wda;
        fw[1][2] = fma(wda, da, fw[1][2]);
      } else {
        fw[0][0] += w;
        fw[0][1] += wda;
        fw[0][2] = fma(wda, da, fw[0][2]);
      }
    }
    S = warp_sum(S);
    npp = warp_sum(npp);
    double W2[2], wa2[2], waa2[2], sumw = 0.0;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      W2[k] = warp_sum(fw[k][0]);
      wa2[k] = warp_sum(fw[k][1]);
      waa2[k] = warp_sum(fw[k][2]);
      sumw += W2[k];
    }
    const double rdf = (double)(n - 4);
    S /= rdf;
    const double mn = npp / (double)n;
    const double var_pp = (double)n * mn * (1.0 - mn) / (double)(n - 1);
    const double kappa = 1.0 + 4.0 * var_pp / ((double)n * mn * mn);
    const double stddev = sqrt(S) * (kappa / mn);
    double v_icpt[2], v_slope[2];
    double cov_scale = 1.0;
    if (cov_mode == MDEGGS_RLM_COV_XTWX) {
      cov_scale = sumw / (double)n;  // X~ = X sqrt(w / mean(w))  =>  cov.unscaled = mean(w) (X'WX)^-1
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        double dm = wa2[k] / W2[k];
        double mk = ma[k] + dm, sx = waa2[k] - wa2[k] * dm;
        v_icpt[k] = 1.0 / W2[k] + mk * mk / sx;
        v_slope[k] = 1.0 / sx;
      }
    } else {
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        v_icpt[k] = 1.0 / nk[k] + ma[k] * ma[k] / sxa[k];
        v_slope[k] = 1.0 / sxa[k];
      }
    }
    const double beta[4] = {icpt[0], slope[0], icpt[1] - icpt[0], slope[1] - slope[0]};
    const double cov[4] = {v_icpt[0], v_slope[0], v_icpt[0] + v_icpt[1], v_slope[0] + v_slope[1]};
    const int v = tested_coef - 1;
    const double var_v = stddev * stddev * cov_scale * cov[v];
    const double F = beta[v] * beta[v] / var_v;
    if (lane == 0) {
      stat[e] = F;
      status[e] = done ? MDEGGS_EDGE_OK : MDEGGS_EDGE_NOT_CONVERGED;
      iters[e] = it;
      if (coef)
        for (int j = 0; j < 4; ++j) coef[(size_t)e * 4 + j] = beta[j];
    }
  }
}

template <int ITEMS>
inline void launch_fit_rlm_reg(cudaStream_t stream, const double* D, const SegLayout& L, const int32_t* ea,
                               const int32_t* eb, const double* mean, const double* sxx,
                               const uint8_t* bad, int64_t e_lo, int64_t e_hi, int tested_coef, int cov_mode, double* stat,
                               int32_t* status, double* coef, int32_t* iters, const int32_t* err) {
  const size_t per_warp = (size_t)L.npad * 16;
  const int wpb = RlmRegWarps<ITEMS>::value;
  const size_t smem = per_warp * wpb;
  int64_t want = (e_hi - e_lo + wpb - 1) / wpb;
  unsigned grid = (unsigned)(want < 148 * 32 ? want : 148 * 32);
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(k_fit_rlm_reg<ITEMS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k_fit_rlm_reg<ITEMS><<<grid, wpb * 32, smem, stream>>>(D, L, ea, eb, mean, sxx, bad, e_lo, e_hi,
                                                        tested_coef, cov_mode, stat, status, coef, iters, err);
}

// returns 0 on success, non-zero if even the residual-only variant does not fit in shared memory
inline int launch_fit_rlm(int64_t& launches, cudaStream_t stream, const double* D, const SegLayout& L,
                          const int32_t* ea, const int32_t* eb, const double* mean,
                          const double* sxx, const uint8_t* bad, int64_t e_lo, int64_t e_hi, int tested_coef,
                          int cov_mode, double* stat, int32_t* status, double* coef, int32_t* iters, const int32_t* err) {
  if (L.npad <= 32 * 32) {  // register-resident keys: n up to ~1,000 samples
#define MDEGGS_RLM(IT)                                                                                                 \
  launch_fit_rlm_reg<IT>(stream, D, L, ea, eb, mean, sxx, bad, e_lo, e_hi, tested_coef, cov_mode, stat, \
                         status, coef, iters, err)
    if (L.npad <= 32 * 4) MDEGGS_RLM(4);
    else if (L.npad <= 32 * 8) MDEGGS_RLM(8);
    else if (L.npad <= 32 * 16) MDEGGS_RLM(16);
    else MDEGGS_RLM(32);
#undef MDEGGS_RLM
    ++launches;
    return 0;
  }
  const size_t budget = 200 * 1024;
  const size_t row = (size_t)L.npad * 8;
  bool cache = 3 * row * 2 <= budget;  // at least two warps per CTA with A, B, r resident
  size_t per_warp = cache ? 3 * row : row;
  if (per_warp > budget) return 1;
  int wpb = (int)(budget / per_warp);
  if (wpb > 8) wpb = 8;
  size_t smem = per_warp * (size_t)wpb;
  int64_t ne = e_hi - e_lo;
  int64_t want = (ne + wpb - 1) / wpb;
  unsigned grid = (unsigned)(want < 148 * 16 ? want : 148 * 16);
  if (cache) {
    if (smem > 48 * 1024)
      cudaFuncSetAttribute(k_fit_rlm<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget);
    k_fit_rlm<true><<<grid, wpb * 32, smem, stream>>>(D, L, ea, eb, mean, sxx, bad, e_lo, e_hi,
                                                     tested_coef, cov_mode, stat, status, coef, iters, err);
  } else {
    if (smem > 48 * 1024)
      cudaFuncSetAttribute(k_fit_rlm<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget);
    k_fit_rlm<false><<<grid, wpb * 32, smem, stream>>>(D, L, ea, eb, mean, sxx, bad, e_lo, e_hi,
                                                      tested_coef, cov_mode, stat, status, coef, iters, err);
  }
  ++launches;
  return 0;
}

}  // namespace mdeggs

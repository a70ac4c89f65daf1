// mdeggs_napair.cuh — pairs that touch a gene row with missing values, for the two branches of the reference that go
// through a model frame.
//
// rlm(B ~ A * metadata) and aov(B ~ A * metadata) (R/core_functions.R:916-917, 966-967) call model.frame with
// na.action = na.omit: the samples where EITHER gene is NA/NaN are dropped for that pair, and the fit, its residual
// degrees of freedom and the tail probability are those of the n' complete cases (quirk q3 of SURVEY section 8).
// fit_lm_interaction (K == 2, "lm") hands the rows to .lm.fit as they are, which throws: those edges keep
// MDEGGS_EDGE_NONFINITE and never reach this kernel.
//
// Such pairs are rare (a handful of proteomics / Olink rows with gaps), so they do not ride in the streaming kernels:
// k_finish queues every edge whose end points carry the "has NaN" bit and this kernel fits the queue,
// one warp per edge: the complete cases are compacted, category by category, into shared memory; their per-category
// counts, means and centred moments give the same closed forms as the main path (finish_pair) or feed the same IRLS
// (rlm_irls_warp), now with the pair's own counts and degrees of freedom; the tail probability uses the generic
// continued fraction because the per-run tables are tied to the run's degrees of freedom.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "mdeggs.h"
#include "mdeggs_bsort.cuh"
#include "mdeggs_kernels.cuh"
#include "mdeggs_math.cuh"
#include "mdeggs_rlm.cuh"

namespace mdeggs {

constexpr int NA_WARPS = 4;  // warps per CTA at most (shared memory: 2 or 3 rows of n doubles per warp)

// bits of bad[node] (set by K2 / the fused front kernel)
constexpr int BAD_NAN = 1;   // the row holds NA / NaN
constexpr int BAD_INF = 2;   // the row holds +-Inf

template <int K>
__global__ void __launch_bounds__(NA_WARPS * 32) k_fit_na_pairs(const double* __restrict__ D, SegLayout L,
                                                               const int32_t* __restrict__ ea,
                                                               const int32_t* __restrict__ eb,
                                                               const int32_t* __restrict__ na_q,
                                                               const unsigned int* __restrict__ na_nq, int64_t e_lo,
                                                               int is_rlm, int tested_coef, int cov_mode,
                                                               double* __restrict__ stat, int32_t* __restrict__ status,
                                                               double* __restrict__ coef, int32_t* __restrict__ iters,
                                                               double* __restrict__ p_out,
                                                               const int32_t* __restrict__ err, int do_bs, BsArgs B) {
  extern __shared__ double na_smem[];
  __shared__ unsigned long long s_cand[NA_WARPS][32];
  if (*err) return;
  const unsigned int nq = *na_nq;
  if (nq == 0u) return;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const int rows = is_rlm ? 3 : 2;
  double* sa = na_smem + (size_t)wib * rows * L.n;
  double* sb = sa + L.n;
  double* sr = is_rlm ? sb + L.n : nullptr;
  for (unsigned int q = blockIdx.x * wpb + wib; q < nq; q += gridDim.x * wpb) {
    const int64_t e = e_lo + na_q[q];
    const int a = ea[e], b = eb[e];
    const double* __restrict__ ga = D + (size_t)a * L.npad;
    const double* __restrict__ gb = D + (size_t)b * L.npad;
    // ---- complete cases, category by category ----
    SegLayout Lp;
    Lp.K = K;
    int m = 0, Ke = 0, any_inf = 0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      Lp.off[k] = m;
      for (int i0 = 0; i0 < L.cnt[k]; i0 += 32) {
        const int i = i0 + lane;
        double x = 0.0, y = 0.0;
        bool keep = false;
        if (i < L.cnt[k]) {
          x = __ldg(ga + L.off[k] + i);
          y = __ldg(gb + L.off[k] + i);
          keep = !isnan(x) && !isnan(y);
          any_inf |= keep && (isinf(x) || isinf(y));
        }
        const unsigned mk = __ballot_sync(0xffffffffu, keep);
        if (keep) {
          const int at = m + __popc(mk & ((1u << lane) - 1u));
          sa[at] = x;
          sb[at] = y;
        }
        m += __popc(mk);
      }
      Lp.cnt[k] = m - Lp.off[k];
      Ke += Lp.cnt[k] > 0;
    }
    Lp.off[K] = m;
    Lp.n = m;
    Lp.npad = m;
    any_inf = __any_sync(0xffffffffu, any_inf);
    __syncwarp();
    double st_v = nan(""), p = nan("");
    int code = MDEGGS_EDGE_OK, it = 0;
    bool coef_written = false;
    if (any_inf) {
      code = MDEGGS_EDGE_NONFINITE;  // lm.fit / lm.wfit: "NA/NaN/Inf in 'x'"
    } else if (m == 0 || Ke < 2 || m <= 2 * Ke) {
      code = MDEGGS_EDGE_SINGULAR;   // no complete case, a single level left, or no residual degree of freedom
    } else {
      // per-category means and centred moments of the complete cases (two passes, first-order mean correction)
      double mean_l[2 * K], sxx_l[2 * K];
      PairMoments<K> pm;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const int o = Lp.off[k], c = Lp.cnt[k];
        double s1 = 0.0, s2 = 0.0;
        for (int i = lane; i < c; i += 32) {
          s1 += sa[o + i];
          s2 += sb[o + i];
        }
        s1 = warp_sum(s1);
        s2 = warp_sum(s2);
        double mu_a = c > 0 ? s1 / (double)c : nan(""), mu_b = c > 0 ? s2 / (double)c : nan("");
        double da1 = 0.0, db1 = 0.0, daa = 0.0, dbb = 0.0, dab = 0.0;
        double amin = INFINITY, amax = -INFINITY, bmin = INFINITY, bmax = -INFINITY;
        for (int i = lane; i < c; i += 32) {
          const double x = sa[o + i], y = sb[o + i];
          const double dx = x - mu_a, dy = y - mu_b;
          da1 += dx;
          db1 += dy;
          daa = fma(dx, dx, daa);
          dbb = fma(dy, dy, dbb);
          dab = fma(dx, dy, dab);
          amin = fmin(amin, x);
          amax = fmax(amax, x);
          bmin = fmin(bmin, y);
          bmax = fmax(bmax, y);
        }
        da1 = warp_sum(da1);
        db1 = warp_sum(db1);
        daa = warp_sum(daa);
        dbb = warp_sum(dbb);
        dab = warp_sum(dab);
#pragma unroll
        for (int o2 = 16; o2 > 0; o2 >>= 1) {
          amin = fmin(amin, __shfl_xor_sync(0xffffffffu, amin, o2));
          amax = fmax(amax, __shfl_xor_sync(0xffffffffu, amax, o2));
          bmin = fmin(bmin, __shfl_xor_sync(0xffffffffu, bmin, o2));
          bmax = fmax(bmax, __shfl_xor_sync(0xffffffffu, bmax, o2));
        }
        if (c > 0) {
          const double ca_ = da1 / (double)c, cb_ = db1 / (double)c;
          mu_a += ca_;
          mu_b += cb_;
          daa -= da1 * ca_;
          dbb -= db1 * cb_;
          dab -= da1 * cb_;
          if (amin == amax) {  // constant inside the category: exact zero spread, like K2
            mu_a = amin;
            daa = 0.0;
            dab = 0.0;
          }
          if (bmin == bmax) {
            mu_b = bmin;
            dbb = 0.0;
            dab = 0.0;
          }
        }
        mean_l[k] = mu_a;
        mean_l[K + k] = mu_b;
        sxx_l[k] = daa;
        sxx_l[K + k] = dbb;
        pm.sxy[k] = dab;
      }
      if (K == 2 && is_rlm) {
        bool singular = false;
#pragma unroll
        for (int k = 0; k < 2; ++k)
          if (Lp.cnt[k] < 2 || !(sxx_l[k] > 1e-14 * (sxx_l[k] + (double)Lp.cnt[k] * mean_l[k] * mean_l[k]))) singular = true;
        if (singular) {
          code = MDEGGS_EDGE_SINGULAR;  // rlm(): "'x' is singular"
        } else if (sxx_l[K] == 0.0 && sxx_l[K + 1] == 0.0) {
          code = MDEGGS_EDGE_SELFLOOP;  // constant response: exact fit
          st_v = 0.0;
          p = 1.0;
        } else {
          const double ca[2] = {mean_l[0], mean_l[1]}, cb[2] = {mean_l[K], mean_l[K + 1]}, sxa[2] = {sxx_l[0], sxx_l[1]};
          double beta[4];
          rlm_irls_warp(sa, sb, sr, Lp, ca, cb, sxa, tested_coef, cov_mode, s_cand[wib], lane, &st_v, &code, &it, beta);
          if (code != MDEGGS_EDGE_ZERO_SCALE) {
            p = pf_upper_dev(st_v, 1.0, (double)(m - 4));
            if (coef && lane == 0)
              for (int j = 0; j < 4; ++j) coef[(size_t)e * 4 + j] = beta[j];
            coef_written = true;
          }
          if (code == MDEGGS_EDGE_OK) code = MDEGGS_EDGE_NA_OMIT;
        }
      } else {
        int n_alias = 0;
        int32_t code32 = 0;
        double coef_l[2 * K];
        finish_pair<K>(Lp, mean_l, sxx_l, 0, 1, pm, &st_v, &code32, coef ? coef_l : nullptr, &n_alias);
        code = code32;
        if (code == MDEGGS_EDGE_SELFLOOP) p = 1.0;
        else if (code != MDEGGS_EDGE_SINGULAR && !isnan(st_v))
          p = pf_upper_dev(st_v, (double)(Ke - 1 - n_alias), (double)(m - 2 * Ke + n_alias));
        if (code == MDEGGS_EDGE_OK) code = MDEGGS_EDGE_NA_OMIT;
        if (coef && lane == 0)
          for (int j = 0; j < 2 * K; ++j) coef[(size_t)e * 2 * K + j] = coef_l[j];
        coef_written = true;
      }
    }
    if (lane == 0) {
      stat[e] = st_v;
      status[e] = code;
      p_out[e] = p;
      if (iters) iters[e] = it;
      if (coef && !coef_written)
        for (int j = 0; j < 2 * K; ++j) coef[(size_t)e * 2 * K + j] = nan("");
    }
    if (do_bs) bs_count(B, lane == 0, e, p);
    __syncwarp();
  }
}

}  // namespace mdeggs

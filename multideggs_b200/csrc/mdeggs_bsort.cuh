// mdeggs_bsort.cuh — K8 front end for small and medium edge counts: ordering all edges by p-value without a
// multi-pass device radix sort.
//
// stats::p.adjust (core_functions.R:1027-1035) needs each (percentile, category) segment in p order.  The raw p of
// an edge does not depend on the percentile, so all edges are ordered once.  A 64-bit LSD radix sort is 8 digit
// passes plus histogram/scan kernels: ~11 dependent launches that, for the 10^4..10^6 edges of a network, are pure
// launch latency (cfg3: 53,035 edges, ~100 us).  Here instead:
//   1. the kernel that produces p (k_finish) drops every edge into one of NB monotone buckets of its p-value
//      (bs_bucket: two buckets per binade below 2^-10, equal-width buckets above — balanced for uniform null p-values
//      and for the heavy left tail of true signals) with a warp-aggregated atomic that also hands the edge its slot
//      inside the bucket;
//   2. k_bsort_scatter turns the bucket counts into offsets (every CTA scans the few thousand counts itself) and
//      writes (key, edge) to offset[bucket] + slot;
//   3. k_bsort_sort: one CTA per run of whole buckets (~BS_CAP..2 BS_CAP elements) orders the run in shared memory
//      (a warp ranks each small bucket by (key, edge id), so the order is deterministic) and emits what the BH passes read: p in ascending
//      order, the node ids of the two end points, the edge id.
// Three short launches after k_finish.  A bucket that alone exceeds the shared-memory capacity (a pathological
// concentration of distinct p-values; a group of IDENTICAL p-values needs no ordering at all since every p.adjust
// method is invariant to the order of ties) is sorted by its CTA in global memory (slow, correct).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "mdeggs_kernels.cuh"

namespace mdeggs {

constexpr int BS_NLOG = 2026;       // buckets 0..2025: (biased exponent, top mantissa bit) of p < 2^-10 (0 also holds p <= 0)
constexpr int BS_CAP = 512;         // run granularity; shared memory holds 2 * BS_CAP elements
constexpr int BS_SORT_THREADS = 256;
constexpr int BS_WARP_MAX = 32;     // largest bucket whose elements are ranked by counting
constexpr int64_t BS_MAX_E = 1 << 26;  // beyond this the CUB radix sort is used

__host__ __device__ __forceinline__ int bs_nlin(int64_t E) {
  int64_t v = E / 16;
  v = v < 1024 ? 1024 : (v > (1 << 20) ? (1 << 20) : v);
  return (int)v;
}
// total buckets: BS_NLOG + nlin + 1 (the last one holds NaN)
__host__ __device__ __forceinline__ int bs_nbuckets(int64_t E) { return BS_NLOG + bs_nlin(E) + 1; }

__device__ __forceinline__ int bs_bucket(double p, int nlin, double scale) {
  if (isnan(p)) return BS_NLOG + nlin;
  if (!(p > 0.0)) return 0;
  if (p < 0x1p-10) return (int)((__double_as_longlong(p) >> 51) & 0xfff);  // 0..2025
  int b = (int)((p - 0x1p-10) * scale);                                    // monotone in p
  b = b > nlin - 1 ? nlin - 1 : b;
  return BS_NLOG + b;
}
__host__ __device__ __forceinline__ double bs_scale(int nlin) { return (double)nlin / (1.0 - 0x1p-10); }

struct BsArgs {
  int* count;        // [NB + 1]: bucket counts (zeroed before the run)
  int* base;         // [NB + 1]: exclusive prefix of the counts
  int32_t* slot;     // [E]: position of the edge inside its bucket; -1: the edge is not part of the order
  unsigned long long* n_valid;  // ordered edges - #NaN, for the BH passes
  unsigned long long* n_order;  // ordered edges (NaN p-values included)
  unsigned long long* n_low;    // ordered edges in the buckets up to the one of p = 0.05 (count-only K8 stops there)
  int nlin;
  double scale;
  int64_t E;
  // Only edges that are tested at SOME percentile can be a member of a (percentile, category) segment: the others are
  // left out of the order altogether (neither counted, scattered nor sorted).  act_t[node][nch] holds the activity
  // masks of a node for 16 percentiles per 16-byte chunk (k_actmask); null: every edge is ordered.
  const uint4* act_t;
  int nch;
  const int32_t *ea, *eb;
  // run_first[r] = first bucket whose offset is >= r BS_CAP, for every window start r BS_CAP <= n_order (written by
  // k_bsort_scatter, read by k_bsort_sort instead of a 19-step binary search over the offsets per CTA)
  int* run_first;
};

// One call per thread of a kernel that covers edges; `live` threads contribute p, `tested` = the edge is tested at some
// percentile (the others stay out of the order).  Must be reached by whole warps.
__device__ __forceinline__ void bs_count_flag(const BsArgs& B, bool live, bool tested, int64_t e, double p) {
  const int lane = threadIdx.x & 31;
  if (live && !tested) {
    B.slot[e] = -1;
    live = false;
  }
  int b = live ? bs_bucket(p, B.nlin, B.scale) : -1 - lane;  // dead lanes never match anyone
  const unsigned peers = __match_any_sync(0xffffffffu, b);
  if (live) {
    const int leader = __ffs(peers) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(&B.count[b], __popc(peers));
    base = __shfl_sync(peers, base, leader);
    B.slot[e] = base + __popc(peers & ((1u << lane) - 1u));
  }
}
__device__ __forceinline__ void bs_count(const BsArgs& B, bool live, int64_t e, double p) {
  const bool tested = !(live && B.act_t) || edge_tested_any(B.act_t, B.nch, B.ea[e], B.eb[e]);
  bs_count_flag(B, live, tested, e, p);
}

// exclusive prefix of count[0..NB) by one CTA into `out` (shared or global); returns the total to every thread.
// `s_warp_tot` is a 32-int shared scratch.
__device__ __forceinline__ int bs_scan_counts(const int* __restrict__ count, int NB, int* __restrict__ out,
                                              int* __restrict__ s_warp_tot) {
  const int nthr = blockDim.x, nwarp = nthr >> 5, wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int ITEMS = 4;  // one 16-byte load / store per thread and round (count and out are 16-byte aligned and
                            // padded to a multiple of 4 entries)
  int carry = 0;
  for (int c0 = 0; c0 < NB; c0 += nthr * ITEMS) {
    const int i0 = c0 + threadIdx.x * ITEMS;
    int v[ITEMS] = {0, 0, 0, 0}, mine = 0;
    if (i0 < NB) {
      const int4 q = __ldcg(reinterpret_cast<const int4*>(count + i0));
      v[0] = q.x;
      v[1] = q.y;
      v[2] = q.z;
      v[3] = q.w;
    }
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
      v[j] = i0 + j < NB ? v[j] : 0;
      mine += v[j];
    }
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp_tot[wid] = incl;
    __syncthreads();
    int woff = 0, tot = 0;
    for (int w = 0; w < nwarp; ++w) {
      const int t = s_warp_tot[w];
      woff += w < wid ? t : 0;
      tot += t;
    }
    const int run = carry + woff + incl - mine;
    if (i0 < NB) *reinterpret_cast<int4*>(out + i0) = make_int4(run, run + v[0], run + v[0] + v[1], run + v[0] + v[1] + v[2]);
    carry += tot;
    __syncthreads();
  }
  return carry;
}

// tail of k_percolate in multi-GPU runs: the bucket of every gathered p-value
struct PercolateBucketTail {
  BsArgs B;
  __device__ __forceinline__ void operator()(bool live, bool tested, int64_t e, double p) const {
    bs_count_flag(B, live, tested, e, p);
  }
};

// stand-alone counting pass (multi-GPU: the p-values arrive through the all-gather, after k_finish)
__global__ void __launch_bounds__(256) k_bsort_count(BsArgs B, const double* __restrict__ p_edge,
                                                     const int32_t* __restrict__ err) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = e < B.E && !*err;
  bs_count(B, live, e, live ? p_edge[e] : 0.0);
}

// ------------------------------------------------------------------------------------------------
// finish, one thread per edge: closed-form statistic from the cross moments, then the tail probability
//   src 0: moments from k_fit_stream (sxy[e][K]);  src 1: moments from the dense matrices S[k][i][j];
//   src 2: statistic and status already written by k_fit_rlm
//   mode 0: 2 * (1 - pt(|t|, df2))            core_functions.R:1072
//   mode 1: pf(F, df1, df2, lower.tail=FALSE)  core_functions.R:968 (aov) / 920 (f.robftest)
//   do_bs: also drop the edge into its p-value bucket (single-GPU runs: every edge is finished here)
// ------------------------------------------------------------------------------------------------
constexpr int FIN_THREADS = 256;  // (<= 256: the queue holds thread ids as bytes)
constexpr int FIN_CF_BUDGET = 24;  // terms of the first try (6 rounds of 4)
template <int K>
__global__ void __launch_bounds__(FIN_THREADS) k_finish(int src, const double* __restrict__ sxy, const double* __restrict__ S,
                                                int nn, SegLayout L, const int32_t* __restrict__ ea,
                                                const int32_t* __restrict__ eb,
                                                const double* __restrict__ mean, const double* __restrict__ sxx,
                                                const uint8_t* __restrict__ bad, int64_t e_lo, int64_t e_hi, int mode,
                                                double df1, double df2, TailConsts tc,
                                                const double* __restrict__ lbeta_p, double* __restrict__ stat,
                                                int32_t* __restrict__ status, double* __restrict__ coef,
                                                double* __restrict__ p_out, const int32_t* __restrict__ err,
                                                int do_bs, BsArgs B, int32_t* __restrict__ fin_q,
                                                unsigned int* __restrict__ fin_nq, int32_t* __restrict__ na_q,
                                                unsigned int* __restrict__ na_nq) {
  const int64_t e = e_lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = e < e_hi && !*err;
  double p = 0.0;
  bool deferred = false;
  bool na_pair = false;  // rlm / aov: a row of this pair has NA -> complete-case fit by k_fit_na_pairs
  int n_alias = 0;
  tc.lbeta_ab = *lbeta_p;
  if (live) {
  if (src != 2) {
    const int a = ea[e], b = eb[e];
    const int bb = bad[a] | bad[b];
    if (na_q && a != b && (bb & 1)) {  // (an Inf may sit in a sample that the other gene's NA removes)
      na_pair = true;
    } else if (a == b || bb) {
      const bool self = (a == b) && !bad[a];
      stat[e] = self ? 0.0 : nan("");
      status[e] = self ? MDEGGS_EDGE_SELFLOOP : MDEGGS_EDGE_NONFINITE;
      if (coef)
        for (int j = 0; j < 2 * K; ++j) coef[(size_t)e * 2 * K + j] = nan("");
    } else {
      PairMoments<K> pm;
      if (src == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) pm.sxy[k] = sxy[(size_t)e * K + k];
      } else {
        const int i = a < b ? a : b, j = a < b ? b : a;
#pragma unroll
        for (int k = 0; k < K; ++k) pm.sxy[k] = L.cnt[k] ? __ldg(S + ((size_t)k * nn + i) * nn + j) : 0.0;
      }
      finish_pair<K>(L, mean, sxx, a, b, pm, stat + e, status + e, coef ? coef + (size_t)e * 2 * K : nullptr, &n_alias);
    }
  } else if (na_q && status[e] == MDEGGS_EDGE_NONFINITE) {  // (k_fit_rlm skipped the pair)
    const int a = ea[e], b = eb[e];
    na_pair = a != b && ((bad[a] | bad[b]) & 1);
  }
  const int st = na_pair ? MDEGGS_EDGE_NONFINITE : status[e];
  if (na_pair) p = nan("");
  else if (st == MDEGGS_EDGE_SINGULAR || st == MDEGGS_EDGE_NONFINITE || st == MDEGGS_EDGE_ZERO_SCALE) p = nan("");
  else if (st == MDEGGS_EDGE_SELFLOOP) p = 1.0;
  else if (st == MDEGGS_EDGE_ALIASED) {  // rare: this edge's own degrees of freedom, generic continued fraction
    const double sv = stat[e];
    p = isnan(sv) ? nan("") : pf_upper_dev(sv, df1 - (double)n_alias, df2 + (double)n_alias);
  } else {
    // fin_q: first try a short continued fraction (enough away from the mode of the distribution); the edges that
    // need the long one (O(sqrt(max(a, b))) terms, a deep dependent FP64 chain) are queued for k_finish_slow, where
    // every lane is such an edge, instead of holding their whole warps here
    bool ok = true;
    const double sv = stat[e];
    const int budget = fin_q ? FIN_CF_BUDGET : CF_TABLE_LEN;
    p = (mode == 0) ? p_two_sided_ref_tab(tc, sv, df2, budget, fin_q ? &ok : nullptr)
                    : pf_upper_tab(tc, sv, df1, df2, budget, fin_q ? &ok : nullptr);
    deferred = !ok;
  }
  }
  if (fin_q) {  // warp-aggregated append (whole warps reach this point)
    const unsigned m = __ballot_sync(0xffffffffu, deferred);
    if (m) {
      const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
      unsigned int base = 0;
      if (lane == leader) base = atomicAdd(fin_nq, (unsigned int)__popc(m));
      base = __shfl_sync(0xffffffffu, base, leader);
      if (deferred) fin_q[base + __popc(m & ((1u << lane) - 1u))] = (int32_t)(e - e_lo);
    }
  }
  if (na_q) {  // same for the pairs with missing values
    const unsigned m = __ballot_sync(0xffffffffu, na_pair);
    if (m) {
      const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
      unsigned int base = 0;
      if (lane == leader) base = atomicAdd(na_nq, (unsigned int)__popc(m));
      base = __shfl_sync(0xffffffffu, base, leader);
      if (na_pair) na_q[base + __popc(m & ((1u << lane) - 1u))] = (int32_t)(e - e_lo);
    }
  }
  const bool done = live && !deferred && !na_pair;
  if (done) p_out[e] = p;
  if (do_bs) bs_count(B, done, e, p);  // reached by every thread of the CTA
}

// the queued edges of k_finish: full-length continued fraction, every lane busy
__global__ void __launch_bounds__(256) k_finish_slow(const int32_t* __restrict__ fin_q,
                                                     const unsigned int* __restrict__ fin_nq, int64_t e_lo, int mode,
                                                     double df1, double df2, TailConsts tc,
                                                     const double* __restrict__ lbeta_p,
                                                     const double* __restrict__ stat, double* __restrict__ p_out,
                                                     const int32_t* __restrict__ err, int do_bs, BsArgs B) {
  if (*err) return;
  tc.lbeta_ab = *lbeta_p;
  const unsigned int nq = *fin_nq;
  for (unsigned int i0 = blockIdx.x * blockDim.x; i0 < nq; i0 += gridDim.x * blockDim.x) {
    const unsigned int i = i0 + threadIdx.x;
    const bool live = i < nq;
    int64_t e = 0;
    double p = 0.0;
    if (live) {
      e = e_lo + fin_q[i];
      const double sv = stat[e];
      p = (mode == 0) ? p_two_sided_ref_tab(tc, sv, df2) : pf_upper_tab(tc, sv, df1, df2);
      p_out[e] = p;
    }
    if (do_bs) bs_count(B, live, e, p);
  }
}


// scan_in_cta: the bucket offsets are not in B.base yet: every CTA computes them from the counts in shared memory
// (NB <= BS_SCAN_SMEM), CTA 0 publishes them for k_bsort_sort and the BH passes
constexpr int BS_SCAN_SMEM = 8192;
__global__ void __launch_bounds__(256) k_bsort_scatter(BsArgs B, int scan_in_cta, const double* __restrict__ p_edge,
                                                       unsigned long long* __restrict__ bkey,
                                                       int32_t* __restrict__ bidx, const int32_t* __restrict__ err) {
  extern __shared__ int s_base[];  // NB ints when scan_in_cta
  __shared__ int s_warp_tot[32];
  if (*err) return;
  const int* base = B.base;
  const int NB = BS_NLOG + B.nlin + 1;
  int total = 0;
  if (!scan_in_cta && blockIdx.x == 0 && threadIdx.x == 0) {  // offsets came from a device-wide scan of count[0..NB]
    *B.n_valid = (unsigned long long)(B.base[NB] - __ldcg(B.count + NB - 1));
    *B.n_order = (unsigned long long)B.base[NB];
    *B.n_low = (unsigned long long)B.base[bs_bucket(0.05, B.nlin, B.scale) + 1];
  }
  if (scan_in_cta) {
    total = bs_scan_counts(B.count, NB, s_base, s_warp_tot);
    if (blockIdx.x == 0) {
      for (int i = threadIdx.x; i < NB; i += blockDim.x) B.base[i] = s_base[i];
      if (threadIdx.x == 0) {
        B.base[NB] = total;
        *B.n_valid = (unsigned long long)(total - __ldcg(B.count + NB - 1));
        *B.n_order = (unsigned long long)total;
        const int bq = bs_bucket(0.05, B.nlin, B.scale) + 1;
        *B.n_low = (unsigned long long)(bq < NB ? s_base[bq] : total);
      }
    }
    base = s_base;
  }
  // window starts: bucket b (b == NB: the end of the order) is the first one at or beyond r BS_CAP for every r with
  // base[b - 1] < r BS_CAP <= base[b]
  {
    auto base_at = [&](int b) { return (scan_in_cta && b == NB) ? total : base[b]; };
    // (every CTA holds the offsets - its own scan in shared memory, or the global array -: the buckets are spread over
    // all threads of the grid)
    const long long nthr = (long long)gridDim.x * blockDim.x;
    for (long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x; b <= NB; b += nthr) {
      const int hi = base_at((int)b), lo_excl = b == 0 ? -1 : base_at((int)b - 1);
      if (hi == lo_excl) continue;
      for (int r = lo_excl < 0 ? 0 : lo_excl / BS_CAP + 1; r <= hi / BS_CAP; ++r) B.run_first[r] = (int)b;
    }
  }
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= B.E) return;
  const int sl = B.slot[e];
  if (sl < 0) return;  // not part of the order
  const double p = p_edge[e];
  const int pos = base[bs_bucket(p, B.nlin, B.scale)] + sl;
  bkey[pos] = dkey(p);
  bidx[pos] = (int32_t)e;  // (measured: carrying the end points along as a 16-byte payload costs the scatter 50 us on
                           // 8 M edges and saves the sort 6)
}

__device__ __forceinline__ bool bs_less(unsigned long long ka, int ia, unsigned long long kb, int ib) {
  return ka < kb || (ka == kb && ia < ib);
}

// Bitonic sort in the "flip" formulation: every compare-exchange leaves the smaller element at the lower index, so
// virtual +infinity padding above n never has to move and the network works for any n (pairs whose upper index is
// >= n are skipped).  Sorts by (key, edge id).
template <class Mem>
__device__ __forceinline__ void bs_bitonic(Mem mem, int n, int n2, int tid0 = -1, int nthr0 = 0) {
  const int tid = tid0 < 0 ? (int)threadIdx.x : tid0, nthr = tid0 < 0 ? (int)blockDim.x : nthr0;
  for (int k = 2; k <= n2; k <<= 1) {
    const int hk = k >> 1;
    for (int t = tid; t < (n2 >> 1); t += nthr) {  // flip step: i <-> mirrored partner in its k-block
      const int blk = t / hk, w = t - blk * hk;
      const int i = blk * k + w, x = blk * k + (k - 1 - w);
      if (x < n) mem.cas(i, x);
    }
    mem.sync();
    for (int j = hk >> 1; j > 0; j >>= 1) {
      for (int t = tid; t < (n2 >> 1); t += nthr) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));  // index with bit j clear
        const int x = i | j;
        if (x < n) mem.cas(i, x);
      }
      mem.sync();
    }
  }
}

struct BsSmem {
  unsigned long long* k;
  int* i;
  __device__ __forceinline__ void cas(int a, int b) const {
    const unsigned long long ka = k[a], kb = k[b];
    const int ia = i[a], ib = i[b];
    if (bs_less(kb, ib, ka, ia)) {
      k[a] = kb;
      k[b] = ka;
      i[a] = ib;
      i[b] = ia;
    }
  }
  __device__ __forceinline__ void sync() const { __syncthreads(); }
};
struct BsSmemWarp {  // one warp sorting its own range of shared memory
  unsigned long long* k;
  int* i;
  __device__ __forceinline__ void cas(int a, int b) const { BsSmem{k, i}.cas(a, b); }
  __device__ __forceinline__ void sync() const { __syncwarp(); }
};
struct BsGmem {
  unsigned long long* k;
  int32_t* i;
  __device__ __forceinline__ void cas(int a, int b) const {
    const unsigned long long ka = __ldcg(k + a), kb = __ldcg(k + b);
    const int ia = __ldcg(i + a), ib = __ldcg(i + b);
    if (bs_less(kb, ib, ka, ia)) {
      __stcg(k + a, kb);
      __stcg(k + b, ka);
      __stcg(i + a, ib);
      __stcg(i + b, ia);
    }
  }
  __device__ __forceinline__ void sync() const {
    __threadfence_block();
    __syncthreads();
  }
};

// slow path for one oversize bucket [lo, hi): nothing to do when all keys are equal (any order of a tie group gives
// the same adjusted p-values), else the same network run by this CTA directly on global memory
__device__ void bs_sort_global(unsigned long long* __restrict__ bkey, int32_t* __restrict__ bidx, int lo, int hi) {
  __shared__ unsigned long long s_min, s_max;
  if (threadIdx.x == 0) {
    s_min = ~0ull;
    s_max = 0ull;
  }
  __syncthreads();
  unsigned long long mn = ~0ull, mx = 0ull;
  for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    const unsigned long long k = bkey[i];
    mn = k < mn ? k : mn;
    mx = k > mx ? k : mx;
  }
  atomicMin(&s_min, mn);
  atomicMax(&s_max, mx);
  __syncthreads();
  const bool tie_group = s_min == s_max;
  __syncthreads();
  if (tie_group) return;
  const int n = hi - lo;
  int n2 = 1;
  while (n2 < n) n2 <<= 1;
  BsGmem g{bkey + lo, bidx + lo};
  bs_bitonic(g, n, n2);
}

// one CTA per run r: the whole buckets whose first element lies in [r BS_CAP, (r+1) BS_CAP)
__global__ void __launch_bounds__(BS_SORT_THREADS) k_bsort_sort(BsArgs B, unsigned long long* __restrict__ bkey,
                                                               int32_t* __restrict__ bidx,
                                                               double* __restrict__ ps, int32_t* __restrict__ sa,
                                                               int32_t* __restrict__ sb, int32_t* __restrict__ sidx,
                                                               const int32_t* __restrict__ err) {
  __shared__ unsigned long long sk[2 * BS_CAP];
  __shared__ int si[2 * BS_CAP];
  if (*err) return;
  const int NB = BS_NLOG + B.nlin + 1;
  const int r = blockIdx.x;
  const int total = B.base[NB];
  if ((long long)r * BS_CAP > (long long)total) return;  // beyond the order (the grid is sized for E edges)
  // first bucket whose offset is >= r BS_CAP, and the same for the next window (k_bsort_scatter: run_first)
  const int bf = B.run_first[r];
  const int bl = (long long)(r + 1) * BS_CAP > (long long)total ? NB : B.run_first[r + 1];
  if (bf >= bl) return;  // no bucket starts in this window (an oversize bucket of an earlier run covers it)
  const int lo = B.base[bf], hi = B.base[bl];
  const int n = hi - lo;
  if (n <= 0) return;
  if (n <= 2 * BS_CAP) {
    // The run fits in shared memory.  Buckets are tiny (16 edges on average, 1-3 in the binade buckets of the small
    // p-values), so the work is spread per ELEMENT, not per bucket: a thread per bucket marks its elements with the
    // bucket's range, then a thread per element ranks its (key, edge id) inside that range with a short loop of
    // shared-memory reads.  No barrier-separated sorting network, no serial walk over buckets.  The rare bucket of
    // more than BS_WARP_MAX elements is sorted in place by the CTA (below).
    __shared__ unsigned short s_st[2 * BS_CAP], s_cn[2 * BS_CAP];
    __shared__ int s_big[2 * BS_CAP / BS_WARP_MAX];
    __shared__ int s_nbig;
    if (threadIdx.x == 0) s_nbig = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      sk[i] = bkey[lo + i];
      si[i] = bidx[lo + i];
      s_cn[i] = 0;
    }
    __syncthreads();
    for (int b = bf + threadIdx.x; b < bl; b += blockDim.x) {
      const int o = B.base[b], bn = B.base[b + 1] - o;
      if (bn <= 0) continue;
      const int blo = o - lo;
      if (bn <= BS_WARP_MAX) {
        for (int q = blo; q < blo + bn; ++q) {
          s_st[q] = (unsigned short)blo;
          s_cn[q] = (unsigned short)bn;
        }
      } else {
        s_big[atomicAdd(&s_nbig, 1)] = b;
      }
    }
    __syncthreads();
    auto emit = [&](int pos, unsigned long long k, int e) {
      ps[pos] = k != ~0ull ? dkey_inv(k) : nan("");
      sa[pos] = B.ea[e];
      sb[pos] = B.eb[e];
      sidx[pos] = e;
    };
    for (int q = threadIdx.x; q < n; q += blockDim.x) {
      const int bn = s_cn[q];
      if (bn == 0) continue;  // member of a larger bucket
      const int blo = s_st[q];
      const unsigned long long mk = sk[q];
      const int me = si[q];
      int rk = 0;
      for (int t = 0; t < bn; ++t) rk += bs_less(sk[blo + t], si[blo + t], mk, me);
      emit(lo + blo + rk, mk, me);
    }
    // Larger buckets (the binade buckets just below 2^-10 hold hundreds of null p-values on a multi-million edge list):
    // the whole CTA runs the network on one bucket at a time; a single warp would keep its CTA - and, at the end of
    // the grid, the whole GPU - waiting for tens of microseconds.
    const int nbig = s_nbig;
    __syncthreads();  // (every thread has read its neighbours' keys: the network may now move them)
    for (int i = 0; i < nbig; ++i) {
      const int b = s_big[i];
      const int blo = B.base[b] - lo, bn = B.base[b + 1] - B.base[b];
      int n2 = 1;
      while (n2 < bn) n2 <<= 1;
      bs_bitonic(BsSmem{sk + blo, si + blo}, bn, n2);  // (ends with a CTA barrier)
      for (int t = threadIdx.x; t < bn; t += blockDim.x) emit(lo + blo + t, sk[blo + t], si[blo + t]);
    }
    return;
  }
  // the run holds an oversize bucket: bucket by bucket
  for (int b = bf; b < bl; ++b) {
    const int blo = B.base[b], bhi = B.base[b + 1];
    const int bn = bhi - blo;
    if (bn <= 0) continue;  // uniform across the CTA
    if (bn <= 2 * BS_CAP) {
      int n2 = 1;
      while (n2 < bn) n2 <<= 1;
      for (int i = threadIdx.x; i < bn; i += blockDim.x) {
        sk[i] = bkey[blo + i];
        si[i] = bidx[blo + i];
      }
      __syncthreads();
      bs_bitonic(BsSmem{sk, si}, bn, n2);
      for (int i = threadIdx.x; i < bn; i += blockDim.x) {
        bkey[blo + i] = sk[i];
        bidx[blo + i] = si[i];
      }
      __syncthreads();
    } else {
      bs_sort_global(bkey, bidx, blo, bhi);
      __syncthreads();
    }
  }
  __threadfence_block();
  __syncthreads();
  for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    const unsigned long long k = bkey[i];
    const int e = bidx[i];
    ps[i] = k != ~0ull ? dkey_inv(k) : nan("");
    sa[i] = B.ea[e];
    sb[i] = B.eb[e];
    sidx[i] = e;
  }
}

}  // namespace mdeggs

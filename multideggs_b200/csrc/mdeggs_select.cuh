// mdeggs_select.cuh — K3: exact multi-quantile selection over the whole node-filtered matrix, no sort.
//
// Replaces  cut_off <- stats::quantile(as.matrix(assayData), prob = percentile, na.rm = TRUE)
// (R/core_functions.R:718), which the reference recomputes for each of the 13-15 percentiles.  Type 7 needs
// two order statistics per percentile (floor/ceil of 1 + (N-1)p): T = 2P target ranks out of N ~ 8.4e6
// values.  A most-significant-digit radix SELECT finds all T at once: six digit passes (12,12,12,12,12,4 key
// bits), each histogramming only the elements whose already-fixed key prefix equals the prefix of some
// target, each followed by a small kernel (one CTA per active prefix) that walks the histograms and narrows
// every target.  Only the first two passes read the matrix D; after 24 bits are fixed the surviving
// candidates (a few thousand per target) are compacted once (third and last read of D) and the remaining
// four passes run on that list.  Versus a full 64-bit radix sort of the matrix: 3 reads of D instead of ~16
// read+write sweeps, and no key/scratch arrays.  If the candidate list overflows (huge tie groups) the
// remaining passes fall back to scanning D.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cub/block/block_scan.cuh>

#include "mdeggs_kernels.cuh"

namespace mdeggs {

constexpr int RSEL_MAX_T = 256;       // 2 * P
constexpr int RSEL_BINS = 4096;
constexpr int RSEL_SMEM_SLOTS = 2;    // slots whose histogram is privatised in shared memory (32 KB)
constexpr int RSEL_PASSES = 6;

struct RselState {
  long long N;                               // non-NaN values
  int T;                                     // targets
  int U[2];                                  // distinct active prefixes; pass d reads U[d & 1], writes U[(d+1) & 1]
  unsigned int done;                         // CTAs of the current narrowing step that have finished
  unsigned long long prefix[RSEL_MAX_T];     // per target: key bits fixed so far (right aligned)
  long long rank[RSEL_MAX_T];                // per target: rank inside its prefix group (0-based)
  int slot[2][RSEL_MAX_T];                   // per target: index into slot_prefix (double buffered like U)
  unsigned long long slot_prefix[2][RSEL_MAX_T];  // sorted, unique
};

__host__ __device__ __forceinline__ int rsel_width(int pass) { return pass < 5 ? 12 : 4; }
__host__ __device__ __forceinline__ int rsel_consumed(int pass) { return 12 * pass; }

struct RselCand {
  unsigned long long* keys;   // compacted candidate keys (24-bit prefix matches an active slot)
  unsigned int* count;        // number of candidates appended
  unsigned int capacity;
};

__device__ __forceinline__ int rsel_find_slot(const unsigned long long* s_pref, int U, unsigned long long pre) {
  int lo = 0, hi = U - 1;
  while (lo <= hi) {
    int mid = (lo + hi) >> 1;
    unsigned long long q = s_pref[mid];
    if (q == pre) return mid;
    if (q < pre) lo = mid + 1;
    else hi = mid - 1;
  }
  return -1;
}

// MODE 0: histogram pass over D;  MODE 1: compaction of D into cand (after pass 1);  MODE 2: histogram pass over
// the compacted candidates (falls back to D when the compaction overflowed)
template <int MODE>
__global__ void __launch_bounds__(256) k_rsel_pass(const double* __restrict__ D, const int32_t* __restrict__ src_of_pos,
                                                   int npad, const int32_t* __restrict__ n_nodes_p, int pass,
                                                   const RselState* __restrict__ st, unsigned int* __restrict__ hist,
                                                   unsigned long long* __restrict__ nan_count, RselCand cand) {
  extern __shared__ unsigned int sh[];  // [RSEL_SMEM_SLOTS][RSEL_BINS]
  __shared__ unsigned long long s_pref[RSEL_MAX_T];
  __shared__ int s_U;
  const int width = rsel_width(pass), consumed = rsel_consumed(pass);
  const int nbins = 1 << width;
  if (threadIdx.x == 0) s_U = pass == 0 ? 1 : st->U[pass & 1];
  __syncthreads();
  const int U = s_U;
  const int usm = (MODE == 1) ? 0 : (U < RSEL_SMEM_SLOTS ? U : RSEL_SMEM_SLOTS);
  for (int i = threadIdx.x; i < usm * nbins; i += blockDim.x) sh[i] = 0;
  for (int i = threadIdx.x; i < U; i += blockDim.x) s_pref[i] = pass == 0 ? 0ull : st->slot_prefix[pass & 1][i];
  __syncthreads();
  const int shift = 64 - consumed - width;
  const unsigned long long dmask = (unsigned long long)(nbins - 1);
  const unsigned long long pmin = s_pref[0], pmax = s_pref[U - 1];

  auto visit = [&](unsigned long long key) {  // histogram one key (MODE 0 / 2)
    int slot = 0;
    if (pass > 0) {
      const unsigned long long pre = key >> (64 - consumed);
      if (pre < pmin || pre > pmax) return;  // outside every active prefix: the common case after pass 1
      slot = rsel_find_slot(s_pref, U, pre);
      if (slot < 0) return;
    }
    const unsigned int digit = (unsigned int)((key >> shift) & dmask);
    if (slot < RSEL_SMEM_SLOTS) atomicAdd(&sh[slot * nbins + digit], 1u);
    else atomicAdd(&hist[(size_t)slot * RSEL_BINS + digit], 1u);
  };

  const bool overflow = (MODE == 2) && (*cand.count > cand.capacity);
  if (MODE == 2 && !overflow) {
    const unsigned int nc = *cand.count;
    for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < nc; i += gridDim.x * blockDim.x)
      visit(cand.keys[i]);
  } else {
    const int n_nodes = *n_nodes_p;
    const int half = npad >> 1;  // double2 units per row (npad is a multiple of 4)
    const int lane = threadIdx.x & 31;
    unsigned long long nans = 0;
    // each CTA takes 4 rows at a time so that every thread has 4 independent 16-byte loads in flight
    for (int row0 = blockIdx.x * 4; row0 < n_nodes; row0 += gridDim.x * 4) {
      for (int c = threadIdx.x; c < half; c += blockDim.x) {
        double2 v[4];
#pragma unroll
        for (int r = 0; r < 4; ++r)
          v[r] = (row0 + r < n_nodes) ? __ldg(reinterpret_cast<const double2*>(D + (size_t)(row0 + r) * npad) + c)
                                      : make_double2(0.0, 0.0);
        const bool ok0 = src_of_pos[2 * c] >= 0, ok1 = src_of_pos[2 * c + 1] >= 0;  // padding slots are skipped
        if (MODE == 1) {
          // compaction: keep keys whose fixed prefix matches an active slot; one warp-aggregated append per
          // 8 elements (exclusive warp scan of the per-thread counts + a single atomicAdd by the last lane)
          unsigned int mask = 0;  // bit (2r + h) set: element kept
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            if (row0 + r >= n_nodes) break;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              if (!(h ? ok1 : ok0)) continue;
              const unsigned long long pre = dkey(h ? v[r].y : v[r].x) >> (64 - consumed);
              if (pre >= pmin && pre <= pmax && rsel_find_slot(s_pref, U, pre) >= 0) mask |= 1u << (2 * r + h);
            }
          }
          const int nk = __popc(mask);
          const unsigned am = __activemask();
          if (__any_sync(am, nk > 0)) {
            int incl = nk;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
              int t = __shfl_up_sync(am, incl, o);
              if (lane >= o) incl += t;
            }
            const int last = 31 - __clz(am);
            unsigned int base = 0;
            if (lane == last) base = atomicAdd(cand.count, (unsigned int)incl);
            base = __shfl_sync(am, base, last);
            const unsigned int idx = base + (unsigned int)(incl - nk);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              if (mask & (1u << q)) {
                const unsigned int at = idx + __popc(mask & ((1u << q) - 1u));
                if (at < cand.capacity) cand.keys[at] = dkey((q & 1) ? v[q >> 1].y : v[q >> 1].x);
              }
            }
          }
          continue;
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          if (row0 + r >= n_nodes) break;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            if (!(h ? ok1 : ok0)) continue;
            const unsigned long long key = dkey(h ? v[r].y : v[r].x);
            if (pass == 0 && key == ~0ull) ++nans;
            visit(key);
          }
        }
      }
    }
    if (MODE == 0 && pass == 0) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) nans += __shfl_xor_sync(0xffffffffu, nans, o);
      if (lane == 0 && nans) atomicAdd(nan_count, nans);
    }
  }
  if (MODE == 1) return;
  __syncthreads();
  for (int i = threadIdx.x; i < usm * nbins; i += blockDim.x) {
    unsigned int c = sh[i];
    if (c) atomicAdd(&hist[(size_t)(i / nbins) * RSEL_BINS + (i % nbins)], c);
  }
}

// one CTA per active prefix slot: narrow the targets of that slot by one digit; the last CTA to finish rebuilds
// the (sorted, unique) slot table for the next pass and, after the final pass, turns the selected keys into
// type-7 cut-offs  q = (1-h) x[lo] + h x[hi]  in R's operation order.
__global__ void __launch_bounds__(256) k_rsel_scan(int pass, RselState* __restrict__ st, unsigned int* __restrict__ hist,
                                                   const unsigned long long* __restrict__ nan_count,
                                                   const int32_t* __restrict__ n_nodes_p, int n,
                                                   const double* __restrict__ pct, int P,
                                                   double* __restrict__ cutoff) {
  typedef cub::BlockScan<long long, 256> Scan;
  __shared__ typename Scan::TempStorage tmp;
  __shared__ long long cum[RSEL_BINS + 1];
  __shared__ int s_last;
  const int width = rsel_width(pass);
  const int nbins = 1 << width;
  const int tid = threadIdx.x, s = blockIdx.x;
  const int cur = pass & 1, nxt = cur ^ 1;
  if (pass == 0) {
    if (s != 0) return;
    __shared__ long long sN;
    if (tid == 0) {
      sN = (long long)(*n_nodes_p) * n - (long long)(*nan_count);
      st->N = sN;
      st->T = 2 * P;
      st->U[0] = 1;
      st->slot_prefix[0][0] = 0;
    }
    __syncthreads();
    const long long N = sN;
    for (int p = tid; p < P; p += blockDim.x) {
      long long ilo = 1, ihi = 1;
      if (N > 0) {
        double idx = __dadd_rn(1.0, __dmul_rn((double)(N - 1), pct[p]));
        ilo = (long long)floor(idx);
        ihi = (long long)ceil(idx);
        ilo = ilo < 1 ? 1 : (ilo > N ? N : ilo);
        ihi = ihi < 1 ? 1 : (ihi > N ? N : ihi);
      }
      st->prefix[2 * p] = 0;
      st->prefix[2 * p + 1] = 0;
      st->rank[2 * p] = ilo - 1;
      st->rank[2 * p + 1] = ihi - 1;
      st->slot[0][2 * p] = 0;
      st->slot[0][2 * p + 1] = 0;
    }
    __threadfence();
    __syncthreads();
  }
  const int T = 2 * P;
  const int U = pass == 0 ? 1 : st->U[cur];
  if (s >= U) return;
  // exclusive scan of hist[s][0..nbins): 256 threads x 16 bins
  unsigned int* h = hist + (size_t)s * RSEL_BINS;
  constexpr int PER = RSEL_BINS / 256;
  long long v[PER], ex[PER];
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    int b = tid * PER + j;
    v[j] = b < nbins ? (long long)h[b] : 0;
  }
  long long agg;
  Scan(tmp).ExclusiveSum(v, ex, agg);
#pragma unroll
  for (int j = 0; j < PER; ++j) cum[tid * PER + j] = ex[j];
  if (tid == 0) cum[RSEL_BINS] = agg;
  __syncthreads();
  for (int t = tid; t < T; t += blockDim.x) {
    if (st->slot[cur][t] != s) continue;
    long long r = st->rank[t];
    int lo = 0, hi = nbins - 1;  // largest b with cum[b] <= r  (then hist[b] > 0 and r < cum[b+1])
    while (lo < hi) {
      int mid = (lo + hi + 1) >> 1;
      if (cum[mid] <= r) lo = mid;
      else hi = mid - 1;
    }
    st->prefix[t] = (st->prefix[t] << width) | (unsigned long long)lo;
    st->rank[t] = r - cum[lo];
  }
#pragma unroll
  for (int j = 0; j < PER; ++j) h[tid * PER + j] = 0;  // clean slate for the next pass
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(&st->done, 1u) == (unsigned)(U - 1));
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // rebuild the sorted, unique slot table in shared memory (T <= 256 == blockDim.x), O(T^2) in parallel
  __shared__ unsigned long long s_p[RSEL_MAX_T];
  __shared__ int s_first[RSEL_MAX_T];
  {
    volatile unsigned long long* pref = st->prefix;
    if (tid < T) s_p[tid] = pref[tid];
    __syncthreads();
    if (tid < T) {
      int first = 1;
      for (int v = 0; v < tid; ++v) first &= (s_p[v] != s_p[tid]);
      s_first[tid] = first;
    }
    __syncthreads();
    if (tid < T) {
      int sl = 0;
      for (int u = 0; u < T; ++u) sl += (s_first[u] && s_p[u] < s_p[tid]);
      st->slot[nxt][tid] = sl;
      if (s_first[tid]) st->slot_prefix[nxt][sl] = s_p[tid];
    }
    if (tid == 0) {
      int Un = 0;
      for (int u = 0; u < T; ++u) Un += s_first[u];
      st->U[nxt] = Un;
      st->done = 0;
    }
  }
  __syncthreads();
  if (pass == RSEL_PASSES - 1) {
    const long long N = st->N;
    const unsigned long long* pref = s_p;
    for (int p = tid; p < P; p += blockDim.x) {
      if (N <= 0) {
        cutoff[p] = nan("");
        continue;
      }
      double idx = __dadd_rn(1.0, __dmul_rn((double)(N - 1), pct[p]));
      double lo = floor(idx);
      double q = dkey_inv(pref[2 * p]);
      double xh = dkey_inv(pref[2 * p + 1]);
      if (idx > lo && xh != q) {
        double hfrac = __dsub_rn(idx, lo);
        q = __dadd_rn(__dmul_rn(__dsub_rn(1.0, hfrac), q), __dmul_rn(hfrac, xh));
      }
      cutoff[p] = q;
    }
  }
}

// After two digit passes (24 key bits fixed) and the compaction, ONE kernel finishes every target: a CTA per active
// prefix pulls that prefix's candidates (typically a few thousand) into shared memory, sorts them (bitonic) and reads
// each target at its rank; the last CTA to finish evaluates the type-7 cut-offs.  Groups that do not fit in shared
// memory (massive ties) or a compaction overflow take a slower in-CTA digit-by-digit path over the same source.
constexpr int RSEL_FIN_CAP = 8192;  // keys per CTA in shared memory (64 KB)
constexpr int RSEL_FIN_THREADS = 512;

__global__ void __launch_bounds__(RSEL_FIN_THREADS) k_rsel_finish(const double* __restrict__ D, const int32_t* __restrict__ src_of_pos,
                                                     int npad, const int32_t* __restrict__ n_nodes_p,
                                                     RselState* __restrict__ st, RselCand cand,
                                                     const double* __restrict__ pct, int P, double* __restrict__ cutoff) {
  extern __shared__ unsigned long long fk[];  // RSEL_FIN_CAP keys; reused as a 4096-bin histogram on the slow path
  __shared__ unsigned int s_cnt;
  __shared__ int s_last;
  __shared__ long long s_scan[RSEL_FIN_THREADS];
  __shared__ unsigned long long s_pfx;
  __shared__ long long s_rank;
  const int tid = threadIdx.x, s = blockIdx.x;
  const int T = 2 * P;
  const int U = st->U[0];
  if (s >= U) return;
  const unsigned long long prefix = st->slot_prefix[0][s];  // 24 bits
  const bool overflow = *cand.count > cand.capacity;
  const unsigned int nc = overflow ? 0u : *cand.count;
  const int n_nodes = *n_nodes_p;
  const int half = npad >> 1;

  // visit every source key whose top `bits` bits equal `pfx`
  auto for_each_match = [&](unsigned long long pfx, int bits, auto&& fn) {
    if (!overflow) {
      // 8 independent loads in flight per thread: the candidate list is read by every CTA
      for (unsigned int i0 = tid; i0 < nc; i0 += RSEL_FIN_THREADS * 8) {
        unsigned long long k[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          unsigned int i = i0 + u * RSEL_FIN_THREADS;
          k[u] = i < nc ? __ldg(cand.keys + i) : 0ull;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (i0 + u * RSEL_FIN_THREADS < nc && (k[u] >> (64 - bits)) == pfx) fn(k[u]);
      }
    } else {
      for (int row = 0; row < n_nodes; ++row) {
        const double2* R2 = reinterpret_cast<const double2*>(D + (size_t)row * npad);
        for (int c = tid; c < half; c += RSEL_FIN_THREADS) {
          double2 v = __ldg(R2 + c);
          if (src_of_pos[2 * c] >= 0) {
            unsigned long long k = dkey(v.x);
            if ((k >> (64 - bits)) == pfx) fn(k);
          }
          if (src_of_pos[2 * c + 1] >= 0) {
            unsigned long long k = dkey(v.y);
            if ((k >> (64 - bits)) == pfx) fn(k);
          }
        }
      }
    }
  };

  if (tid == 0) s_cnt = 0;
  __syncthreads();
  for_each_match(prefix, 24, [&](unsigned long long k) {
    unsigned int pos = atomicAdd(&s_cnt, 1u);
    if (pos < (unsigned)RSEL_FIN_CAP) fk[pos] = k;
  });
  __syncthreads();
  const unsigned int c = s_cnt;
  if (c <= (unsigned)RSEL_FIN_CAP) {
    unsigned int n2 = 1;
    while (n2 < c) n2 <<= 1;
    for (unsigned int i = c + tid; i < n2; i += RSEL_FIN_THREADS) fk[i] = ~0ull;
    __syncthreads();
    for (unsigned int k = 2; k <= n2; k <<= 1) {
      for (unsigned int j = k >> 1; j > 0; j >>= 1) {
        for (unsigned int i = tid; i < n2; i += RSEL_FIN_THREADS) {
          unsigned int ixj = i ^ j;
          if (ixj > i) {
            unsigned long long a = fk[i], b = fk[ixj];
            bool up = (i & k) == 0;
            if ((a > b) == up) {
              fk[i] = b;
              fk[ixj] = a;
            }
          }
        }
        __syncthreads();
      }
    }
    for (int t = tid; t < T; t += RSEL_FIN_THREADS)
      if (st->slot[0][t] == s) st->prefix[t] = fk[st->rank[t]];
  } else {
    // slow path: one target at a time, digit by digit, histogram in shared memory
    unsigned int* hist = reinterpret_cast<unsigned int*>(fk);
    for (int t = 0; t < T; ++t) {
      if (st->slot[0][t] != s) continue;  // uniform across the CTA
      if (tid == 0) {
        s_pfx = prefix;
        s_rank = st->rank[t];
      }
      __syncthreads();
      for (int pass = 2; pass < RSEL_PASSES; ++pass) {
        const int width = rsel_width(pass), consumed = rsel_consumed(pass);
        const int nbins = 1 << width;
        const int shift = 64 - consumed - width;
        for (int i = tid; i < RSEL_BINS; i += RSEL_FIN_THREADS) hist[i] = 0;
        __syncthreads();
        const unsigned long long pfx = s_pfx;
        for_each_match(pfx, consumed, [&](unsigned long long k) {
          atomicAdd(&hist[(unsigned int)((k >> shift) & (unsigned long long)(nbins - 1))], 1u);
        });
        __syncthreads();
        // exclusive scan of the bins (16 per thread), then locate the bin holding the rank
        long long loc = 0;
        for (int j = 0; j < RSEL_BINS / RSEL_FIN_THREADS; ++j) loc += hist[tid * (RSEL_BINS / RSEL_FIN_THREADS) + j];
        s_scan[tid] = loc;
        __syncthreads();
        if (tid == 0) {
          long long run = 0;
          for (int i = 0; i < RSEL_FIN_THREADS; ++i) {
            long long v = s_scan[i];
            s_scan[i] = run;
            run += v;
          }
        }
        __syncthreads();
        const long long r = s_rank;
        long long run = s_scan[tid];
        int found = -1;
        long long found_base = 0;
        for (int j = 0; j < RSEL_BINS / RSEL_FIN_THREADS; ++j) {
          int b = tid * (RSEL_BINS / RSEL_FIN_THREADS) + j;
          long long h = hist[b];
          if (b < nbins && r >= run && r < run + h) {
            found = b;
            found_base = run;
          }
          run += h;
        }
        __syncthreads();
        if (found >= 0) {  // exactly one thread
          s_pfx = (pfx << width) | (unsigned long long)found;
          s_rank = r - found_base;
        }
        __syncthreads();
      }
      if (tid == 0) st->prefix[t] = s_pfx;
      __syncthreads();
    }
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(&st->done, 1u) == (unsigned)(U - 1));
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (tid == 0) st->done = 0;
  const long long N = st->N;
  volatile unsigned long long* pref = st->prefix;
  for (int p = tid; p < P; p += RSEL_FIN_THREADS) {
    if (N <= 0) {
      cutoff[p] = nan("");
      continue;
    }
    double idx = __dadd_rn(1.0, __dmul_rn((double)(N - 1), pct[p]));
    double lo = floor(idx);
    double q = dkey_inv(pref[2 * p]);
    double xh = dkey_inv(pref[2 * p + 1]);
    if (idx > lo && xh != q) {
      double hfrac = __dsub_rn(idx, lo);
      q = __dadd_rn(__dmul_rn(__dsub_rn(1.0, hfrac), q), __dmul_rn(hfrac, xh));
    }
    cutoff[p] = q;
  }
}

}  // namespace mdeggs

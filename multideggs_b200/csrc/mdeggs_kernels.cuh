// mdeggs_kernels.cuh — hand-written sm_100a kernels of the multiDEGGs hot path (see include/mdeggs.h).
//
// Data layout in HBM (all FP64):
//   D[node][npad]   gene-major rows of the network-node genes; the samples of each category form one
//                   contiguous segment, every segment padded to a multiple of 4 doubles (32 B) so that a
//                   warp reads rows as coalesced 16-byte vectors; padding slots hold the segment mean, so
//                   they contribute exactly 0 to every centred sum and no kernel needs a tail predicate.
//   mean/sxx/med[node][K]   per gene, per category: mean, centred sum of squares, exact median.
// Each kernel cites the reference lines it replaces.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cub/block/block_reduce.cuh>
#include <cub/block/block_scan.cuh>

#include "mdeggs.h"
#include "mdeggs_math.cuh"
#include "mdeggs_qvalue_tab.h"

namespace mdeggs {

struct SegLayout {
  int K, n, npad;
  int off[MDEGGS_MAX_K + 1];  // padded segment offsets (doubles); off[K] == npad
  int cnt[MDEGGS_MAX_K];      // samples per category
};

// order-preserving map double -> uint64; every NaN maps to the maximum key
__device__ __forceinline__ unsigned long long dkey(double v) {
  if (isnan(v)) return ~0ull;
  unsigned long long b = (unsigned long long)__double_as_longlong(__dadd_rn(v, 0.0));  // -0.0 -> +0.0
  return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double dkey_inv(unsigned long long k) {
  unsigned long long b = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)b);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Streamed input copy (engine: stage_inputs).  The band of a host matrix crosses PCIe in row chunks on a copy stream
// while the compute phase already runs; behind every chunk the copy stream performs a stream-ordered 32-bit write
// (cuStreamWriteValue32: fenced after the chunk's DMA) of gate[c] = 1.  A kernel that is about to read rows of chunk c
// polls the word with a system-scope acquire load.  Everything a kernel waits for was enqueued BEFORE the kernel, so
// the wait always ends; the time-out (10 s of the global timer) only guards against a failed copy: it raises the
// device error flag (value 2) and lets the kernel run on.
constexpr int GATE_MAX_CHUNKS = 32;
constexpr int GATE_ERR = 2;
__device__ __forceinline__ void gate_wait(const unsigned int* __restrict__ gate, int c, int32_t* __restrict__ err) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(gate + c) : "memory");
  if (v) return;
  unsigned long long t0, t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (;;) {
    __nanosleep(200);
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(gate + c) : "memory");
    if (v) return;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if (t - t0 > 10000000000ull) {
      if (err) atomicMax(err, GATE_ERR);
      return;
    }
  }
}

// Are kernel launches asynchronous in this process?  The host raises `flag` (mapped pinned memory) right after the
// launch call returns; a kernel that still sees it down after 2 ms was launched synchronously (a profiler that
// serialises launches, CUDA_LAUNCH_BLOCKING=1): work that a running kernel waits for must then be enqueued BEFORE it.
__global__ void k_async_probe(const volatile unsigned int* flag, unsigned int* result) {
  unsigned long long t0, t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (;;) {
    if (*flag) {
      *result = 1u;
      return;
    }
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if (t - t0 > 2000000ull) {
      *result = 0u;
      return;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// K0: node filter   (core_functions.R:333-336: nodes <- unique(c(from,to)); assayData[rownames %in% nodes,])
//   k_mark_nodes : every CTA marks the end points of its edges in a shared-memory bitmap of the G rows (G/8 bytes;
//                  direct global atomicOr when G > K0_SMEM_ROWS), ORs the non-zero words into the global bitmap;
//                  the LAST CTA to finish turns the bitmap into per-word exclusive popcount prefixes + n_nodes.
//   k_build_nodes: thread per row -> node_of_row / rowlist;  thread per edge -> node ids (ea, eb) of its end
//                  points, read by every later edge kernel instead of from/to + a dependent node_of_row look-up.
// ------------------------------------------------------------------------------------------------
// bad[node]: bit 0 = the row holds NA / NaN, bit 1 = it holds +-Inf.  Several warps (one per category) may flag the same
// node, so the byte is set through an atomic OR on its 32-bit word (the array is 4-byte aligned and padded).
__device__ __forceinline__ int nonfinite_bits(double v) { return isnan(v) ? 1 : (isinf(v) ? 2 : 0); }
__device__ __forceinline__ void mark_bad(uint8_t* __restrict__ bad, int node, int bits) {
  atomicOr(reinterpret_cast<unsigned int*>(bad) + (node >> 2), (unsigned int)bits << ((node & 3) * 8));
}

constexpr int K0_SMEM_ROWS = 1 << 18;  // 32 KB bitmap

__global__ void __launch_bounds__(256) k_mark_nodes(const int32_t* __restrict__ from, const int32_t* __restrict__ to,
                                                    int64_t E, int G, unsigned int* __restrict__ bitmap,
                                                    int32_t* __restrict__ word_base, int32_t* __restrict__ n_nodes,
                                                    unsigned int* __restrict__ done, int32_t* __restrict__ err) {
  extern __shared__ unsigned int s_bits[];
  __shared__ int s_last;
  const int W = (G + 31) >> 5;
  const bool use_smem = G <= K0_SMEM_ROWS;
  if (use_smem) {
    for (int w = threadIdx.x; w < W; w += blockDim.x) s_bits[w] = 0u;
    __syncthreads();
  }
  unsigned int* bits = use_smem ? s_bits : bitmap;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  bool bad_idx = false;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < E; e += stride) {
    const int a = from[e] - 1, b = to[e] - 1;
    if (a < 0 || a >= G || b < 0 || b >= G) {
      bad_idx = true;
      continue;
    }
    const unsigned ma = 1u << (a & 31), mb = 1u << (b & 31);
    if (!(bits[a >> 5] & ma)) atomicOr(&bits[a >> 5], ma);
    if (!(bits[b >> 5] & mb)) atomicOr(&bits[b >> 5], mb);
  }
  if (bad_idx) *err = 1;
  if (use_smem) {
    __syncthreads();
    for (int w = threadIdx.x; w < W; w += blockDim.x) {
      const unsigned v = s_bits[w];
      if (v) atomicOr(&bitmap[w], v);
    }
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(done, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // last CTA: exclusive prefix of the per-word popcounts (each thread owns a contiguous run of words)
  typedef cub::BlockScan<int, 256> Scan;
  __shared__ typename Scan::TempStorage tmp;
  const int per = (W + 255) / 256;
  const int w0 = threadIdx.x * per, w1 = min(W, w0 + per);
  int mine = 0;
  for (int w = w0; w < w1; ++w) mine += __popc(__ldcg(bitmap + w));
  int ex, total;
  Scan(tmp).ExclusiveSum(mine, ex, total);
  for (int w = w0; w < w1; ++w) {
    word_base[w] = ex;
    ex += __popc(__ldcg(bitmap + w));
  }
  if (threadIdx.x == 0) {
    *n_nodes = total;
    *done = 0u;
  }
}

__device__ __forceinline__ int node_index(const unsigned int* __restrict__ bitmap,
                                          const int32_t* __restrict__ word_base, int row) {
  return word_base[row >> 5] + __popc(bitmap[row >> 5] & ((1u << (row & 31)) - 1u));
}

__global__ void __launch_bounds__(256) k_build_nodes(const unsigned int* __restrict__ bitmap,
                                                     const int32_t* __restrict__ word_base, int G,
                                                     const int32_t* __restrict__ from, const int32_t* __restrict__ to,
                                                     int64_t E, int32_t* __restrict__ node_of_row,
                                                     int32_t* __restrict__ rowlist, int32_t* __restrict__ ea,
                                                     int32_t* __restrict__ eb, const int32_t* __restrict__ err) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < G) {
    const int g = (int)i;
    const bool f = (bitmap[g >> 5] >> (g & 31)) & 1u;
    const int pos = node_index(bitmap, word_base, g);
    node_of_row[g] = f ? pos : -1;
    if (f) rowlist[pos] = g;
  }
  if (i < E && !*err) {
    ea[i] = node_index(bitmap, word_base, from[i] - 1);
    eb[i] = node_index(bitmap, word_base, to[i] - 1);
  }
}

// (K1 gather kernels live in mdeggs_qsel.cuh: they also build the value histogram of the quantile selection)

// ------------------------------------------------------------------------------------------------
// K3 S3 (candidate compaction of the quantile selection, mdeggs_qsel.cuh) rides inside K2: K2 holds every value of D
// in registers anyway, so it reads the 2-byte bin id the gather kept for each value, looks the bin up in a 4 KB
// shared table of "slot" bins and appends the few values (~0.6 %) that belong to a slot to that slot's list.
// ------------------------------------------------------------------------------------------------
constexpr int QSC_BINS = 4096;
constexpr unsigned QSC_NO_BIN = 0xffffu;
struct QsCompactArgs {
  const unsigned short* B16;  // null: no quantile selection in this run
  const int* U;               // number of slots
  const int* slot_bin;
  const int* slot_fat;        // slots that are not copied (tie groups): only their key range is tracked
  const unsigned int* slot_count;
  const unsigned int* slot_off;
  unsigned int* slot_cursor;
  unsigned long long *slot_min, *slot_max;
  unsigned long long* cand;
};
// whole CTA; s_tab[QSC_BINS], 16-byte aligned
__device__ __forceinline__ void qsc_load_table(const QsCompactArgs& Q, unsigned char* __restrict__ s_tab) {
  for (int i = threadIdx.x; i < QSC_BINS / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(s_tab)[i] = make_uint4(~0u, ~0u, ~0u, ~0u);
  __syncthreads();
  const int U = *Q.U;
  for (int i = threadIdx.x; i < U; i += blockDim.x) s_tab[Q.slot_bin[i]] = (unsigned char)i;
  __syncthreads();
}
__device__ __forceinline__ void qsc_visit(const QsCompactArgs& Q, const unsigned char* __restrict__ s_tab, unsigned b,
                                          unsigned long long key) {
  if (b >= (unsigned)QSC_BINS) return;  // padding / NaN
  const unsigned sl = s_tab[b];
  if (sl == 0xff) return;  // ~99 % of the values leave here
  if (Q.slot_fat[sl]) {
    if (key < __ldcg(Q.slot_min + sl)) atomicMin(Q.slot_min + sl, key);
    if (key > __ldcg(Q.slot_max + sl)) atomicMax(Q.slot_max + sl, key);
    return;
  }
  const unsigned pos = atomicAdd(Q.slot_cursor + sl, 1u);
  if (pos < Q.slot_count[sl]) Q.cand[Q.slot_off[sl] + pos] = key;
}

// ------------------------------------------------------------------------------------------------
// K2: per-gene, per-category statistics
//   median  — apply(category_vec, 1, stats::median, na.rm = TRUE)   core_functions.R:353-365
//   mean, centred Sxx — the sufficient statistics of fit_lm_interaction / aov shared by every edge that
//   touches the gene (core_functions.R:1058-1074, 966-968)
// One CTA per gene, the row staged in shared memory; one warp per category. The exact median is an order
// statistic found by bisection in key space (no sort, no scratch).
// ------------------------------------------------------------------------------------------------
// Exact k-th smallest (0-based) of m keys held behind `key_at(i)`, one warp.  Keys are known to lie in
// [kmin, kmax] except NaN keys (== ~0), which sit above every valid key and are never selected (kth < n_valid).
// Bisection in KEY space: [lo, hi) always holds the answer, cnt_lo / cnt_hi = number of keys below lo / hi; it
// starts from the optional guess bracket [g_lo, g_hi], halves the key interval until <= 32 keys remain and ranks
// those from shared memory.  At most ~64 counting passes whatever the data; a handful for anything well spread.
template <class KeyAt>
__device__ __forceinline__ unsigned long long warp_select(KeyAt key_at, int m, int kth, unsigned long long kmin,
                                                          unsigned long long kmax, int n_valid,
                                                          unsigned long long* scratch32, int lane,
                                                          unsigned long long g_lo = 0ull,
                                                          unsigned long long g_hi = 0ull) {
  if (kmin == kmax) return kmin;
  unsigned long long lo = kmin, hi = kmax + 1ull;  // kmax is a valid key, hence < ~0
  int cnt_lo = 0, cnt_hi = n_valid;
  auto narrow = [&](unsigned long long pivot) {
    if (pivot <= lo || pivot >= hi) return;
    int c = 0;
    for (int i = lane; i < m; i += 32) c += (key_at(i) < pivot);
    c = __reduce_add_sync(0xffffffffu, c);
    if (c <= kth) {
      lo = pivot;
      cnt_lo = c;
    } else {
      hi = pivot;
      cnt_hi = c;
    }
  };
  if (g_hi > g_lo) {
    narrow(g_lo);
    narrow(g_hi);
  }
  while (cnt_hi - cnt_lo > 32 && hi - lo > 1ull) narrow(lo + ((hi - lo) >> 1));
  if (cnt_hi - cnt_lo > 32) return lo;  // hi == lo + 1: more than 32 keys equal to lo
  int base = 0;
  for (int i0 = 0; i0 < m; i0 += 32) {
    int i = i0 + lane;
    unsigned long long k = i < m ? key_at(i) : ~0ull;
    bool in = i < m && k >= lo && k < hi;
    unsigned mask = __ballot_sync(0xffffffffu, in);
    if (in) scratch32[base + __popc(mask & ((1u << lane) - 1u))] = k;
    base += __popc(mask);
  }
  __syncwarp();
  const int ncand = cnt_hi - cnt_lo;
  unsigned long long mine = lane < ncand ? scratch32[lane] : ~0ull;
  int r = 0;
  for (int j = 0; j < ncand; ++j) {  // broadcast reads of the <= 32 candidates
    const unsigned long long other = scratch32[j];
    r += (other < mine) || (other == mine && j < lane);
  }
  unsigned hit = __ballot_sync(0xffffffffu, lane < ncand && r == kth - cnt_lo);
  unsigned long long res = __shfl_sync(0xffffffffu, mine, __ffs(hit) - 1);
  __syncwarp();
  return res;
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS) k_gene_stats(double* __restrict__ D, SegLayout L,
                                                       const int32_t* __restrict__ n_nodes_p,
                                                       double* __restrict__ mean, double* __restrict__ sxx,
                                                       double* __restrict__ med, uint8_t* __restrict__ bad,
                                                       int use_smem, QsCompactArgs Q) {
  extern __shared__ double srow[];
  __shared__ int s_bad;
  __shared__ unsigned long long s_cand[THREADS / 32][32];
  __shared__ __align__(16) unsigned char s_qtab[QSC_BINS];
  if (Q.B16) qsc_load_table(Q, s_qtab);
  const int node = blockIdx.x;
  if (node >= *n_nodes_p) return;
  double* grow = D + (size_t)node * L.npad;
  const double* r = grow;
  if (threadIdx.x == 0) s_bad = 0;
  if (use_smem) {
    for (int k = 0; k < L.K; ++k)
      for (int i = threadIdx.x; i < L.cnt[k]; i += THREADS) srow[L.off[k] + i] = grow[L.off[k] + i];
    r = srow;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int k = warp; k < L.K; k += THREADS / 32) {
    const double* seg = r + L.off[k];
    const int m = L.cnt[k];
    double s = 0.0, vmin = INFINITY, vmax = -INFINITY;
    int nn = 0, nf = 0;
    for (int i = lane; i < m; i += 32) {
      double v = seg[i];
      s += v;
      vmin = fmin(vmin, v);
      vmax = fmax(vmax, v);
      nn += !isnan(v);
      nf |= nonfinite_bits(v);
      if (Q.B16) qsc_visit(Q, s_qtab, Q.B16[(size_t)node * L.npad + L.off[k] + i], dkey(v));
    }
    s = warp_sum(s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      vmin = fmin(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
      vmax = fmax(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
    }
    nn = __reduce_add_sync(0xffffffffu, nn);
    nf = __reduce_or_sync(0xffffffffu, nf);
    double mu = m > 0 ? s / (double)m : nan("");
    double d1 = 0.0, d2 = 0.0;
    for (int i = lane; i < m; i += 32) {
      double d = seg[i] - mu;
      d1 += d;
      d2 = fma(d, d, d2);
    }
    d1 = warp_sum(d1);
    d2 = warp_sum(d2);
    if (m > 0) {
      double c = d1 / (double)m;  // first-order correction of the mean
      mu += c;
      d2 -= d1 * c;
    }
    if (m > 0 && vmin == vmax) {  // constant inside the category: exact zero spread (aliased column in R)
      mu = vmin;
      d2 = 0.0;
    }
    // exact median of the non-NaN values (fmin/fmax ignore NaN, so vmin/vmax bound every valid key)
    double median = nan("");
    if (nn > 0) {
      auto key_at = [&](int i) { return dkey(seg[i]); };
      const unsigned long long kmin = dkey(vmin), kmax = dkey(vmax);
      if (nn & 1) {
        median = dkey_inv(warp_select(key_at, m, nn / 2, kmin, kmax, nn, s_cand[warp], lane));
      } else {
        unsigned long long k1 = warp_select(key_at, m, nn / 2 - 1, kmin, kmax, nn, s_cand[warp], lane);
        int c_le = 0;
        unsigned long long nxt = ~0ull;
        for (int i = lane; i < m; i += 32) {
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
        unsigned long long k2 = (c_le > nn / 2) ? k1 : nxt;
        double a = dkey_inv(k1), b = dkey_inv(k2);
        median = __ddiv_rn(__dadd_rn(a, b), 2.0);
      }
    }
    if (lane == 0) {
      mean[(size_t)node * L.K + k] = mu;
      sxx[(size_t)node * L.K + k] = d2;
      med[(size_t)node * L.K + k] = median;
      if (nf) atomicOr(&s_bad, nf);
    }
    for (int i = L.off[k] + m + lane; i < L.off[k + 1]; i += 32) grow[i] = mu;  // padding := segment mean
  }
  __syncthreads();
  if (threadIdx.x == 0) bad[node] = (uint8_t)s_bad;
}

// Register-resident variant of K2 for segments of up to 32*ITEMS samples: one warp per (gene, category), the
// segment lives in registers as sortable keys, so every counting pass of the selection is ITEMS compares per lane
// with no shared-memory traffic.  `bad` must be zero-initialised by the caller.
// Exact k-th smallest (0-based) of the keys a warp holds in registers (ITEMS per lane; unused slots carry ~0, which
// sorts above every valid key).  Bisection in KEY space with arbitrary pivots: the interval [lo, hi) always holds the
// answer, cnt_lo / cnt_hi = number of keys below lo / hi.  It starts from the caller's guess bracket [g_lo, g_hi]
// (e.g. mean -+ 0.3 sd for a median: one counting pass usually leaves a few dozen candidates), halves the key
// interval until <= 32 keys remain, and ranks those with 32 shuffles.  At most ~64 counting passes whatever the data.
template <int ITEMS>
__device__ __forceinline__ int warp_count_below(const unsigned long long (&keys)[ITEMS], unsigned long long pivot) {
  int c = 0;
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) c += (keys[j] < pivot);
  return __reduce_add_sync(0xffffffffu, c);
}

template <int ITEMS>
__device__ __forceinline__ unsigned long long warp_select_reg(const unsigned long long (&keys)[ITEMS], int kth,
                                                              unsigned long long kmin, unsigned long long kmax,
                                                              int n_valid, unsigned long long* scratch32, int lane,
                                                              unsigned long long g_lo = 0ull,
                                                              unsigned long long g_hi = 0ull) {
  if (kmin == kmax) return kmin;
  unsigned long long lo = kmin, hi = kmax + 1ull;  // kmax is a valid key, hence < ~0
  int cnt_lo = 0, cnt_hi = n_valid;
  auto narrow = [&](unsigned long long pivot) {
    if (pivot <= lo || pivot >= hi) return;
    const int c = warp_count_below<ITEMS>(keys, pivot);
    if (c <= kth) {
      lo = pivot;
      cnt_lo = c;
    } else {
      hi = pivot;
      cnt_hi = c;
    }
  };
  if (g_hi > g_lo) {
    narrow(g_lo);
    narrow(g_hi);
  }
  while (cnt_hi - cnt_lo > 32 && hi - lo > 1ull) narrow(lo + ((hi - lo) >> 1));
  if (cnt_hi - cnt_lo > 32) return lo;  // hi == lo + 1: more than 32 keys equal to lo
  // the <= 32 keys left go to scratch (their order there is irrelevant): each lane counts its own, one warp prefix
  // sum hands out the slots (a ballot per item cost 8 instructions x ITEMS)
  int mine_n = 0;
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) mine_n += (keys[j] >= lo && keys[j] < hi);
  int at = mine_n;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, at, o);
    if (lane >= o) at += t;
  }
  at -= mine_n;
  if (mine_n) {
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
      const unsigned long long k = keys[j];
      if (k >= lo && k < hi) scratch32[at++] = k;
    }
  }
  __syncwarp();
  const int ncand = cnt_hi - cnt_lo;
  unsigned long long mine = lane < ncand ? scratch32[lane] : ~0ull;
  int r = 0;
  for (int j = 0; j < ncand; ++j) {  // broadcast reads of the <= 32 candidates
    const unsigned long long other = scratch32[j];
    r += (other < mine) || (other == mine && j < lane);
  }
  unsigned hit = __ballot_sync(0xffffffffu, lane < ncand && r == kth - cnt_lo);
  unsigned long long res = __shfl_sync(0xffffffffu, mine, __ffs(hit) - 1);
  __syncwarp();
  return res;
}

template <int ITEMS>
__global__ void __launch_bounds__(256) k_gene_stats_reg(double* __restrict__ D, SegLayout L,
                                                        const int32_t* __restrict__ n_nodes_p,
                                                        double* __restrict__ mean, double* __restrict__ sxx,
                                                        double* __restrict__ med, uint8_t* __restrict__ bad,
                                                        QsCompactArgs Q) {
  __shared__ unsigned long long s_cand[8][32];
  __shared__ __align__(16) unsigned char s_qtab[QSC_BINS];
  if (Q.B16) qsc_load_table(Q, s_qtab);
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const long long w = (long long)blockIdx.x * 8 + wib;
  const int node = (int)(w / L.K), k = (int)(w % L.K);
  if (node >= *n_nodes_p) return;
  double* grow = D + (size_t)node * L.npad;
  const double* seg = grow + L.off[k];
  const int m = L.cnt[k];
  unsigned long long keys[ITEMS];
  double s = 0.0, vmin = INFINITY, vmax = -INFINITY;
  int nn = 0, nf = 0;
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    const int i = lane + 32 * j;
    const bool valid = i < m;
    const double v = valid ? seg[i] : nan("");
    keys[j] = dkey(v);  // out-of-range slots carry the NaN key: never counted, never selected
    if (valid) {
      s += v;
      vmin = fmin(vmin, v);
      vmax = fmax(vmax, v);
      nn += !isnan(v);
      nf |= nonfinite_bits(v);
    }
  }
  if (Q.B16) {  // K3 S3: the bin ids of this segment (2 bytes per value), slot values appended to their lists
    const unsigned short* bseg = Q.B16 + (size_t)node * L.npad + L.off[k];
    unsigned bins[ITEMS];
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) bins[j] = lane + 32 * j < m ? (unsigned)bseg[lane + 32 * j] : QSC_NO_BIN;
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) qsc_visit(Q, s_qtab, bins[j], keys[j]);
  }
  s = warp_sum(s);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    vmin = fmin(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
    vmax = fmax(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
  }
  nn = __reduce_add_sync(0xffffffffu, nn);
  nf = __reduce_or_sync(0xffffffffu, nf);
  double mu = m > 0 ? s / (double)m : nan("");
  double d1 = 0.0, d2 = 0.0;
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    if (lane + 32 * j < m) {
      double d = dkey_inv(keys[j]) - mu;  // (-0.0 was canonicalised to +0.0 by dkey: same value)
      d1 += d;
      d2 = fma(d, d, d2);
    }
  }
  d1 = warp_sum(d1);
  d2 = warp_sum(d2);
  if (m > 0) {
    double c = d1 / (double)m;
    mu += c;
    d2 -= d1 * c;
  }
  if (m > 0 && vmin == vmax) {
    mu = vmin;
    d2 = 0.0;
  }
  double median = nan("");
  if (nn > 0) {
    const unsigned long long kmin = dkey(vmin), kmax = dkey(vmax);
    // guess bracket for the median: mean -+ 0.3 sd (holds it for anything close to symmetric; otherwise the bisection
    // simply continues from whichever side the counts point to)
    unsigned long long g_lo = 0ull, g_hi = 0ull;
    if (m > 2 && d2 > 0.0 && isfinite(d2)) {
      const double w = 0.3 * sqrt(d2 / (double)(m - 1));
      g_lo = dkey(mu - w);
      g_hi = dkey(mu + w);
    }
    if (nn & 1) {
      median = dkey_inv(warp_select_reg<ITEMS>(keys, nn / 2, kmin, kmax, nn, s_cand[wib], lane, g_lo, g_hi));
    } else {
      unsigned long long k1 = warp_select_reg<ITEMS>(keys, nn / 2 - 1, kmin, kmax, nn, s_cand[wib], lane, g_lo, g_hi);
      int c_le = 0;
      unsigned long long nxt = ~0ull;
#pragma unroll
      for (int j = 0; j < ITEMS; ++j) {
        c_le += (keys[j] <= k1);
        if (keys[j] > k1 && keys[j] < nxt) nxt = keys[j];
      }
      c_le = __reduce_add_sync(0xffffffffu, c_le);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        unsigned long long other = __shfl_xor_sync(0xffffffffu, nxt, o);
        nxt = other < nxt ? other : nxt;
      }
      unsigned long long k2 = (c_le > nn / 2) ? k1 : nxt;
      median = __ddiv_rn(__dadd_rn(dkey_inv(k1), dkey_inv(k2)), 2.0);
    }
  }
  if (lane == 0) {
    mean[(size_t)node * L.K + k] = mu;
    sxx[(size_t)node * L.K + k] = d2;
    med[(size_t)node * L.K + k] = median;
    if (nf) mark_bad(bad, node, nf);
  }
  for (int i = L.off[k] + m + lane; i < L.off[k + 1]; i += 32) grow[i] = mu;  // padding := segment mean
}

// ------------------------------------------------------------------------------------------------
// K4 / K5: streaming pair fit, one warp per edge
//   K == 2 : fit_lm_interaction  (core_functions.R:1058-1074)  -> t statistic of the A:category coefficient
//   K >= 3 : aov(B ~ A * metadata), row "A:metadata" (core_functions.R:966-968) -> interaction F statistic
// With the per-gene statistics of K2 a pair needs only the K centred cross moments
//   Sxy_k = sum_{i in k} (a_i - mean_a,k)(b_i - mean_b,k)
// i.e. one pass over two rows (16 n bytes), two subtractions and one FMA per sample; the 4x4 (2K x 2K)
// normal equations collapse to closed forms in (Sxx_k, Syy_k, Sxy_k):
//   slope_k = Sxy_k / Sxx_k;  RSS_full = sum_k (Syy_k - Sxy_k slope_k)
//   K == 2: t = (slope_2 - slope_1) / sqrt(RSS_full/(n-4) * (1/Sxx_1 + 1/Sxx_2))
//   K >= 3: RSS_add = sum Syy - (sum Sxy)^2 / sum Sxx;  F = ((RSS_add-RSS_full)/(K-1)) / (RSS_full/(n-2K))
// ------------------------------------------------------------------------------------------------
// k_fit_stream build parameters (tools/fit_variants.sh measures alternatives): resident CTAs per SM and unroll depth
#ifndef FIT_CTAS
#define FIT_CTAS 8
#endif
#ifndef FIT_UNROLL
#define FIT_UNROLL 1
#endif
constexpr int kFitUnroll = FIT_UNROLL;
template <int K>
struct PairMoments {
  double sxy[K];
};

// `n_aliased` (K >= 3): categories whose A:metadata column aov() drops because the predictor is constant inside them;
// the caller evaluates the tail with (Ke - 1 - n_aliased, n - 2 Ke + n_aliased) degrees of freedom.
template <int K>
__device__ __forceinline__ void finish_pair(const SegLayout& L, const double* __restrict__ mean,
                                            const double* __restrict__ sxx, int a, int b,
                                            const PairMoments<K>& pm, double* __restrict__ stat_out,
                                            int32_t* __restrict__ status_out, double* __restrict__ coef_out,
                                            int* __restrict__ n_aliased) {
  double slope[K], icpt[K];
  bool alias[K];
  double rss = 0.0, s_xy = 0.0, s_xx = 0.0, s_yy = 0.0, inv_sum = 0.0;
  int Ke = 0;  // categories that have samples in this omic (an empty level is an aliased column in R, quirk q2)
  int na = 0, last_ok = -1;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    slope[k] = icpt[k] = nan("");
    alias[k] = false;
    if (L.cnt[k] == 0) continue;
    ++Ke;
    double ma = mean[(size_t)a * K + k], mb = mean[(size_t)b * K + k];
    double sa = sxx[(size_t)a * K + k], sb = sxx[(size_t)b * K + k];
    double c = pm.sxy[k];
    // LINPACK dqrdc2 declares a column aliased when its residual norm drops below 1e-7 x its norm
    alias[k] = L.cnt[k] < 2 || !(sa > 1e-14 * (sa + (double)L.cnt[k] * ma * ma));
    s_xy += c;
    s_xx += sa;
    s_yy += sb;
    if (alias[k]) {  // no slope of its own: the category only contributes its spread around the mean
      ++na;
      rss += sb;
      continue;
    }
    last_ok = k;
    slope[k] = c / sa;
    icpt[k] = mb - slope[k] * ma;
    rss += sb - c * slope[k];
    inv_sum += 1.0 / sa;
  }
  // K == 2 (lm, core_functions.R:1068): solve(t(X) %*% X) throws on an aliased column -> hard failure.
  // K >= 3 (aov, core_functions.R:966-968): lm() pivots aliased columns out and the F test keeps the columns left.
  const bool singular = Ke < 2 || (K == 2 && na > 0);
  const bool exact_fit = !singular && s_yy == 0.0;  // response constant in every category: R returns noise/noise
  const bool no_test = K > 2 && !singular && Ke - 1 - na <= 0;  // no interaction column left: Pr(>F)[3] is NA
  double stat;
  if (K == 2) {
    double dof = (double)(L.n - 4);
    double se = sqrt(rss / dof * inv_sum);
    stat = (slope[1] - slope[0]) / se;
  } else {
    double rss_add = s_yy - s_xy * s_xy / s_xx;
    double d1 = (double)(Ke - 1 - na), d2 = (double)(L.n - 2 * Ke + na);
    stat = ((rss_add - rss) / d1) / (rss / d2);
  }
  if (singular || no_test) stat = nan("");
  if (exact_fit) stat = 0.0;
  *stat_out = stat;
  *status_out = singular ? MDEGGS_EDGE_SINGULAR
                         : (exact_fit ? MDEGGS_EDGE_SELFLOOP : (na > 0 ? MDEGGS_EDGE_ALIASED : MDEGGS_EDGE_OK));
  *n_aliased = na;
  if (coef_out) {  // R order: (Intercept), A, metadata2..K, A:metadata2..K; NA when the reference level is empty
    const bool ok = !singular && !exact_fit && !no_test && L.cnt[0] > 0;
    // aliased columns (NA coefficients): A:metadata_j of every category j >= 2 where A is constant; when that is the
    // reference category itself, lm() instead finds the LAST remaining interaction column dependent, so the common
    // slope `A` is that category's and its own interaction coefficient is the NA one
    const int ref = alias[0] ? last_ok : 0;
    const double b1 = ok ? slope[ref] : nan("");
    const double b0 = ok ? mean[(size_t)b * K] - b1 * mean[(size_t)a * K] : nan("");
    coef_out[0] = b0;
    coef_out[1] = b1;
#pragma unroll
    for (int k = 1; k < K; ++k) {
      const bool own = ok && L.cnt[k] > 0 && !alias[k] && k != ref;
      const double sl = own ? slope[k] : b1;
      coef_out[1 + k] = (ok && L.cnt[k] > 0) ? (mean[(size_t)b * K + k] - sl * mean[(size_t)a * K + k]) - b0 : nan("");
      coef_out[K + k] = own ? slope[k] - b1 : nan("");
    }
  }
}

// The streaming kernel only produces the K centred cross moments of each edge (sxy[e*K + k]); the closed forms
// (divisions, sqrt) and the tail probability are done one THREAD per edge by k_finish, so no warp ever idles 31
// lanes behind a scalar epilogue.
template <int K>
__global__ void __launch_bounds__(256, FIT_CTAS) k_fit_stream(const double* __restrict__ D, SegLayout L,
                                                       const int32_t* __restrict__ ea,
                                                       const int32_t* __restrict__ eb,
                                                       const double* __restrict__ mean,
                                                       const uint8_t* __restrict__ bad, int64_t e_lo, int64_t e_hi,
                                                       double* __restrict__ sxy, const int32_t* __restrict__ err) {
  if (*err) return;  // an edge index was out of range: nothing below may touch from/to
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t e = e_lo + warp0; e < e_hi; e += nwarps) {
    // No classification here (self-loops, non-finite rows): k_finish does it and never reads their moments, so the
    // row loads depend on nothing but the two node ids.
    const int a = ea[e], b = eb[e];
    const double2* __restrict__ pa = reinterpret_cast<const double2*>(D + (size_t)a * L.npad);
    const double2* __restrict__ pb = reinterpret_cast<const double2*>(D + (size_t)b * L.npad);
    double acc[K];
    // all K segments are streamed back to back; the K warp reductions come after the last load has been issued.
    // Few live registers on purpose: the kernel is bound by L2 latency, and what hides it is resident warps (measured:
    // 64 warps/SM at 32 registers beat 32 warps/SM at 64 registers with deeper unrolling by 12 %).
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const double mak = mean[(size_t)a * K + k], mbk = mean[(size_t)b * K + k];
      const int beg = L.off[k] >> 1, end = L.off[k + 1] >> 1;
      double acc0 = 0.0, acc1 = 0.0;
#pragma unroll kFitUnroll
      for (int i = beg + lane; i < end; i += 32) {
        const double2 x = __ldg(pa + i), y = __ldg(pb + i);
        acc0 = fma(x.x - mak, y.x - mbk, acc0);
        acc1 = fma(x.y - mak, y.y - mbk, acc1);
      }
      acc[k] = acc0 + acc1;
    }
    double mine = 0.0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const double tot = warp_sum(acc[k]);
      if (lane == k) mine = tot;
    }
    if (lane < K) sxy[(size_t)e * K + lane] = mine;
  }
}

// (k_finish itself lives in mdeggs_bsort.cuh: it also feeds the p-value bucket counts of the K8 ordering)

// ------------------------------------------------------------------------------------------------
// K7: percolation masks   (calc_pvalues_percentile, core_functions.R:723-779)
//   act[p][node] bit k  <=>  median_k(node) > cut_off_p           (core_functions.R:724, strict >)
//   an edge is tested at percentile p iff both ends are active in EXACTLY one category (745-770)
// ------------------------------------------------------------------------------------------------
// thread per (node, chunk of 16 percentiles): writes act[p][node] and the transposed, packed act_t[node][chunk]
__global__ void k_actmask(const double* __restrict__ med, const double* __restrict__ cutoff,
                          const int32_t* __restrict__ n_nodes_p, int P, int K, int act_stride,
                          uint8_t* __restrict__ act, uint4* __restrict__ act_t, int nch) {
  const int n_nodes = *n_nodes_p;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)nch * n_nodes) return;
  const int node = (int)(i / nch), ch = (int)(i - (int64_t)node * nch);
  double m_k[MDEGGS_MAX_K];
  for (int k = 0; k < K; ++k) m_k[k] = med[(size_t)node * K + k];
  unsigned w[4] = {0u, 0u, 0u, 0u};
  for (int j = 0; j < 16; ++j) {
    const int p = ch * 16 + j;
    if (p >= P) break;
    const double c = cutoff[p];
    unsigned m = 0;
    for (int k = 0; k < K; ++k) m |= (unsigned)(m_k[k] > c) << k;  // NaN median: never active
    act[(size_t)p * act_stride + node] = (uint8_t)m;
    w[j >> 2] |= m << (8 * (j & 3));
  }
  act_t[(size_t)node * nch + ch] = make_uint4(w[0], w[1], w[2], w[3]);
}

__device__ __forceinline__ int edge_category(unsigned ma, unsigned mb) {
  unsigned m = ma & mb;
  return (m != 0 && (m & (m - 1)) == 0) ? __ffs(m) : 0;  // 1-based category, 0 = none or shared
}

// bytes of x with exactly one bit set -> 0xff in that byte (SIMD within a register)
__device__ __forceinline__ unsigned single_bit_bytes(unsigned x) {
  return __vcmpeq4(x & __vsub4(x, 0x01010101u), 0u) & __vcmpne4(x, 0u);
}
// is the edge (a, b) specific to exactly one category at any percentile (core_functions.R:745-770)?
__device__ __forceinline__ bool edge_tested_any(const uint4* __restrict__ act_t, int nch, int a, int b) {
  unsigned r = 0u;
  for (int ch = 0; ch < nch; ++ch) {
    const uint4 x = __ldg(act_t + (size_t)a * nch + ch), y = __ldg(act_t + (size_t)b * nch + ch);
    r |= single_bit_bytes(x.x & y.x) | single_bit_bytes(x.y & y.y) | single_bit_bytes(x.z & y.z) |
         single_bit_bytes(x.w & y.w);
  }
  return r != 0u;
}

// counts per percentile: tot_edges (773-779), hard failures, and for padj "none" the raw p < 0.05 count (796-808).
// A thread reads the packed masks of its edge's two end points (16 percentiles per 16-byte load) and walks the
// percentiles in registers; a percentile where no edge of the warp is tested costs one ballot, the others one ballot per
// category (the per-edge properties - hard failure, raw p < 0.05, p not NA - are warp masks computed once).
// `Tail` is called once per thread at the end with (live, tested at some percentile, edge, p): multi-GPU runs drop the
// edge into its p-value bucket there (mdeggs_bsort.cuh), one pass over the gathered p-values instead of two.
struct PercolateNoTail {
  __device__ __forceinline__ void operator()(bool, bool, int64_t, double) const {}
};
template <class Tail>
__global__ void __launch_bounds__(256) k_percolate(Tail tail, const int32_t* __restrict__ ea, const int32_t* __restrict__ eb,
                                                   const int32_t* __restrict__ status,
                                                   const double* __restrict__ p_edge,
                                                   const uint4* __restrict__ act_t, int nch, int64_t E, int P,
                                                   unsigned long long* __restrict__ tot,
                                                   unsigned long long* __restrict__ fail,
                                                   unsigned long long* __restrict__ sig_raw,
                                                   int8_t* __restrict__ cat_out, double* __restrict__ padj_nan,
                                                   const int32_t* __restrict__ err, int K,
                                                   unsigned int* __restrict__ seg_cnt) {
  extern __shared__ unsigned int s_cnt[];  // [3][P], then [P][K] when seg_cnt
  if (*err) return;
  const int n_cnt = 3 * P + (seg_cnt ? P * K : 0);
  for (int i = threadIdx.x; i < n_cnt; i += blockDim.x) s_cnt[i] = 0;
  __syncthreads();
  // (a capped grid walks the edge list with a grid stride: on long lists the kernel then leaves room on every SM for the
  // memory-bound ordering kernels of the main stream instead of running ahead of them)
  const int lane = threadIdx.x & 31;
  for (int64_t e0 = (int64_t)blockIdx.x * blockDim.x; e0 < E; e0 += (int64_t)gridDim.x * blockDim.x) {
  const int64_t e = e0 + threadIdx.x;
  const bool live = e < E;
  int a = 0, b = 0, st = 0;
  double pv = 1.0;
  if (live) {
    a = ea[e];
    b = eb[e];
    st = status[e];
    pv = p_edge[e];
  }
  const bool hard = (st == MDEGGS_EDGE_SINGULAR || st == MDEGGS_EDGE_NONFINITE);
  const unsigned w_hard = __ballot_sync(0xffffffffu, hard);
  const unsigned w_sig = __ballot_sync(0xffffffffu, pv < 0.05);
  const unsigned w_val = __ballot_sync(0xffffffffu, !isnan(pv));
  unsigned tested_bits = 0u;
  for (int ch = 0; ch < nch; ++ch) {
    uint4 x = make_uint4(0u, 0u, 0u, 0u);
    if (live) {
      const uint4 ma = __ldg(act_t + (size_t)a * nch + ch), mb = __ldg(act_t + (size_t)b * nch + ch);
      x = make_uint4(ma.x & mb.x, ma.y & mb.y, ma.z & mb.z, ma.w & mb.w);
    }
    const unsigned xw[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int p0 = ch * 16 + q * 4;
      if (p0 >= P) break;
      // nothing tested in these four percentiles anywhere in the warp (and no row to write): next word
      const unsigned sb4 = single_bit_bytes(xw[q]);
      tested_bits |= sb4;
      const bool any4 = __any_sync(0xffffffffu, sb4 != 0u);
      if (!any4 && !cat_out && !padj_nan) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int p = p0 + j;
        if (p >= P) break;
        const unsigned m = (xw[q] >> (8 * j)) & 0xffu;
        const int c = (m != 0u && (m & (m - 1u)) == 0u) ? __ffs(m) : 0;
        if (cat_out && live) cat_out[(size_t)p * E + e] = (int8_t)c;
        if (padj_nan && live && (c == 0 || isnan(pv))) padj_nan[(size_t)p * E + e] = nan("");  // K8 writes the rest
        if (!any4) continue;
        const unsigned m_t = __ballot_sync(0xffffffffu, c != 0);
        if (!m_t) continue;
        if (lane == 0) {
          atomicAdd(&s_cnt[p], __popc(m_t));
          if (m_t & w_hard) atomicAdd(&s_cnt[P + p], __popc(m_t & w_hard));
          if (m_t & w_sig) atomicAdd(&s_cnt[2 * P + p], __popc(m_t & w_sig));
        }
        if (seg_cnt) {  // members of every (percentile, category) segment = n of p.adjust (non-NA p only)
          for (int k = 1; k <= K; ++k) {
            const unsigned m_k = __ballot_sync(0xffffffffu, c == k) & w_val;
            if (lane == 0 && m_k) atomicAdd(&s_cnt[3 * P + p * K + k - 1], __popc(m_k));
          }
        }
      }
    }
  }
  tail(live, tested_bits != 0u, e, pv);  // (reached by whole warps)
  }
  __syncthreads();
  for (int i = threadIdx.x; i < P; i += blockDim.x) {
    if (s_cnt[i]) atomicAdd(&tot[i], (unsigned long long)s_cnt[i]);
    if (s_cnt[P + i]) atomicAdd(&fail[i], (unsigned long long)s_cnt[P + i]);
    if (s_cnt[2 * P + i]) atomicAdd(&sig_raw[i], (unsigned long long)s_cnt[2 * P + i]);
  }
  if (seg_cnt)
    for (int i = threadIdx.x; i < P * K; i += blockDim.x)
      if (s_cnt[3 * P + i]) atomicAdd(&seg_cnt[i], s_cnt[3 * P + i]);
}

// rows of the device-chosen best percentile: category row (recomputed from the masks) and, when the adjusted
// values of every percentile are still in the scratch matrix, its p.adj row (else a BH pass is rerun for it)
__global__ void k_best_rows(const int32_t* __restrict__ ea, const int32_t* __restrict__ eb,
                            const uint8_t* __restrict__ act, int act_stride,
                            int64_t E, const int32_t* __restrict__ best, int8_t* __restrict__ cat_out,
                            const double* __restrict__ padj_all, double* __restrict__ padj_best,
                            const int32_t* __restrict__ err) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E || *err) return;
  const int p = *best - 1;
  if (cat_out) {
    int c = 0;
    if (p >= 0) {
      c = edge_category(act[(size_t)p * act_stride + ea[e]], act[(size_t)p * act_stride + eb[e]]);
    }
    cat_out[e] = (int8_t)c;
  }
  if (padj_best) padj_best[e] = (p >= 0 && padj_all) ? padj_all[(size_t)p * E + e] : nan("");
}

// ------------------------------------------------------------------------------------------------
// K8: multiple-testing adjustment per (percentile, category) segment + significant-pair count
//   stats::p.adjust(p, "bonferroni" | "BH")  core_functions.R:1027-1035;  count < 0.05  796-808
// The raw p-value of an edge does not depend on the percentile, so ALL edges are sorted once by p
// (CUB radix sort on sortable keys, NaN last); a (percentile, category) segment is a sub-sequence of that
// order. For each segment: rank i by a prefix count of its members, BH value (n/i)*p in R's operation
// order, cummin from the largest p (a suffix min), pmin(1, .), scatter back to network order.
// Three passes over 1024-element tiles with tile-level partials (count -> ranks -> suffix-min).
// ------------------------------------------------------------------------------------------------
constexpr int BH_THREADS = 256;
#ifndef BH_ITEMS_N
#define BH_ITEMS_N 8
#endif
constexpr int BH_ITEMS = BH_ITEMS_N;
constexpr int BH_TILE = BH_THREADS * BH_ITEMS;

__global__ void k_sort_keys(const double* __restrict__ p_edge, int64_t E, unsigned long long* __restrict__ keys,
                            int32_t* __restrict__ idx) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  keys[e] = dkey(p_edge[e]);
  idx[e] = (int32_t)e;
}

__global__ void k_gather_sorted(const unsigned long long* __restrict__ keys_sorted,
                                const int32_t* __restrict__ idx_sorted, const int32_t* __restrict__ ea,
                                const int32_t* __restrict__ eb, int64_t E,
                                double* __restrict__ ps, int32_t* __restrict__ sa, int32_t* __restrict__ sb,
                                unsigned long long* __restrict__ n_valid, unsigned long long* __restrict__ n_order,
                                const int32_t* __restrict__ err) {
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j == 0) *n_order = (unsigned long long)E;  // the radix sort orders every edge
  bool valid = false;
  if (j < E && !*err) {
    unsigned long long k = keys_sorted[j];
    int32_t e = idx_sorted[j];
    valid = (k != ~0ull);
    ps[j] = valid ? dkey_inv(k) : nan("");
    sa[j] = ea[e];
    sb[j] = eb[e];
  }
  unsigned m = __ballot_sync(0xffffffffu, valid);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(n_valid, (unsigned long long)__popc(m));
}

struct BhArgs {
  const double* ps;
  const int32_t *sa, *sb, *sidx;
  const uint8_t* act;
  int act_stride;
  int64_t E;
  const unsigned long long* n_valid;
  const unsigned long long* n_order;  // edges in the order (NaN p-values included, sorted last): <= E
  int K, P, ntiles;
  int p0, p_stride;     // segment y -> percentile p = p0 + (y / K) * p_stride, category c = y % K + 1
                        // (multi-GPU: rank r adjusts percentiles r, r + world, ...)
  const int32_t* best;  // when non-null: p := *best - 1 (device-chosen percentile), grid.y == K
  const unsigned long long* tot;  // tot_edges[P] from k_percolate: percentiles without tested edges are skipped
  int reuse_tiles;      // best-percentile pass: reuse the tile partials of the all-percentile pass (segment index
                        // of percentile p is ((p - p0) / p_stride) * K + c - 1 there)
  const double* seg_pi0;  // q.value: pi0 of each segment (k_qv_pi0), indexed like the tile partials
  // count-only passes (k_bh_count + k_bh_sigrank): no adjusted value below 0.05 can come from p >= 0.05 (every method's
  // factor n/i, n + 1 - i, q n/i is >= 1), so the tiles of the p order that start at or beyond `skip_thr` are neither
  // loaded nor counted (see bh_tiles_visited); the segment sizes then come from k_percolate (`seg_cnt[p * K + c - 1]`) and `seg_low[seg]` keeps
  // the members that were counted (holm needs the rank of the first member beyond).  skip_thr <= 0: every tile is visited.
  double skip_thr;
  const unsigned long long* n_low;
  const unsigned int* seg_cnt;
  long long* seg_low;
};

// Count-only passes with skip_thr > 0 visit only the first bh_tiles_visited() tiles of the p order: `n_low` = ordered
// edges in the p-value buckets up to the one that holds skip_thr (k_bsort_scatter), so every p < skip_thr lies in a
// visited tile.  The other CTAs leave before any barrier or counter (same answer in every thread of a CTA).
__device__ __forceinline__ int bh_tiles_visited(const BhArgs& A) {
  if (!(A.skip_thr > 0.0)) return A.ntiles;
  const long long t = ((long long)*A.n_low + BH_TILE - 1) / BH_TILE;
  return t < (long long)A.ntiles ? (int)t : A.ntiles;
}

// maps blockIdx.y to (percentile p, category c, segment index `seg` into the tile-partial arrays)
__device__ __forceinline__ bool bh_segment(const BhArgs& A, int& p, int& c, size_t& seg) {
  c = (int)blockIdx.y % A.K + 1;
  if (A.best) {
    p = *A.best - 1;
    if (p < 0) return false;
    seg = A.reuse_tiles ? (size_t)((p - A.p0) / A.p_stride) * A.K + (c - 1) : (size_t)blockIdx.y;
  } else {
    p = A.p0 + ((int)blockIdx.y / A.K) * A.p_stride;
    seg = blockIdx.y;
  }
  return true;
}

__device__ __forceinline__ void bh_load(const BhArgs& A, int p, int c, int64_t tile0, int (&flag)[BH_ITEMS],
                                        double (&pv)[BH_ITEMS]) {
  const int64_t nv = (int64_t)*A.n_valid;
  const int64_t j0 = tile0 + (int64_t)threadIdx.x * BH_ITEMS;
  const uint8_t* __restrict__ row = A.act + (size_t)p * A.act_stride;
  static_assert(BH_ITEMS % 4 == 0, "vector path below loads quads");
  if (j0 + BH_ITEMS <= nv) {  // all items valid: 16-byte loads (j0 is a multiple of 4, the arrays are 256-byte aligned)
#pragma unroll
    for (int q = 0; q < BH_ITEMS / 4; ++q) {
      const int4 a4 = *reinterpret_cast<const int4*>(A.sa + j0 + 4 * q);
      const int4 b4 = *reinterpret_cast<const int4*>(A.sb + j0 + 4 * q);
      const double2 p01 = *reinterpret_cast<const double2*>(A.ps + j0 + 4 * q);
      const double2 p23 = *reinterpret_cast<const double2*>(A.ps + j0 + 4 * q + 2);
      const int a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
      const double pp[4] = {p01.x, p01.y, p23.x, p23.y};
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        flag[4 * q + it] = edge_category(row[a[it]], row[b[it]]) == c;
        pv[4 * q + it] = pp[it];
      }
    }
    return;
  }
#pragma unroll
  for (int it = 0; it < BH_ITEMS; ++it) {
    const int64_t j = j0 + it;
    flag[it] = 0;
    pv[it] = 0.0;
    if (j < nv) {
      flag[it] = edge_category(row[A.sa[j]], row[A.sb[j]]) == c;
      pv[it] = A.ps[j];
    }
  }
}

// exclusive scan of the tile counts of one segment (whole CTA of BH_THREADS); seg_n = members of the segment
__device__ __forceinline__ void bh_scan_tiles_body(const int32_t* __restrict__ tile_cnt, int ntiles, size_t seg,
                                                   long long* __restrict__ tile_base, long long* __restrict__ seg_n,
                                                   int n_scan = -1) {
  typedef cub::BlockScan<long long, BH_THREADS> Scan;
  __shared__ typename Scan::TempStorage tmp;
  __shared__ long long carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const size_t base = seg * ntiles;
  const int nt = n_scan >= 0 ? n_scan : ntiles;  // (count-only passes: only the visited tiles hold a count)
  for (int t0 = 0; t0 < nt; t0 += BH_THREADS) {
    int t = t0 + threadIdx.x;
    long long v = t < nt ? (long long)__ldcg(tile_cnt + base + t) : 0, ex, agg;
    Scan(tmp).ExclusiveSum(v, ex, agg);
    if (t < nt) tile_base[base + t] = carry + ex;
    __syncthreads();
    if (threadIdx.x == 0) carry += agg;
    __syncthreads();
  }
  if (threadIdx.x == 0) seg_n[seg] = carry;
}

// "last CTA of the segment" election: every CTA of a segment calls it once after publishing its tile partial;
// returns true in exactly one CTA per segment (and re-arms the counter for the next launch)
__device__ __forceinline__ bool bh_last_of_segment(int* __restrict__ seg_done, size_t seg, int ntiles) {
  __shared__ int s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const int prev = atomicAdd(&seg_done[seg], 1);
    s_last = prev == ntiles - 1;
    if (s_last) seg_done[seg] = 0;
  }
  __syncthreads();
  if (s_last) __threadfence();
  return s_last != 0;
}

// pass A: members per tile; the last CTA of each segment turns the tile counts into tile offsets (ranks)
__global__ void __launch_bounds__(BH_THREADS) k_bh_count(BhArgs A, int32_t* __restrict__ tile_cnt,
                                                        long long* __restrict__ tile_base,
                                                        long long* __restrict__ seg_n, int* __restrict__ seg_done) {
  int p, c;
  size_t seg;
  if (!bh_segment(A, p, c, seg)) return;
  if (A.tot && A.tot[p] == 0) {  // nothing is tested at this percentile (whole segment: no scan needed either)
    if (threadIdx.x == 0) tile_cnt[seg * A.ntiles + blockIdx.x] = 0;
    return;
  }
  // a CTA walks the visited tiles with the grid's stride (count-only passes are launched with a capped grid: the
  // 100,000 CTAs of a full grid that would only find out they have nothing to do cost more than the counting itself)
  const int n_vis = bh_tiles_visited(A);
  const int n_cta = n_vis < (int)gridDim.x ? n_vis : (int)gridDim.x;
  if ((int)blockIdx.x >= n_cta) return;
  typedef cub::BlockReduce<int, BH_THREADS> Reduce;
  __shared__ typename Reduce::TempStorage tmp;
  for (int tile = blockIdx.x; tile < n_vis; tile += gridDim.x) {
    int flag[BH_ITEMS];
    double pv[BH_ITEMS];
    bh_load(A, p, c, (int64_t)tile * BH_TILE, flag, pv);
    int s = 0;
#pragma unroll
    for (int it = 0; it < BH_ITEMS; ++it) s += flag[it];
    const int tot = Reduce(tmp).Sum(s);
    if (threadIdx.x == 0) tile_cnt[seg * A.ntiles + tile] = tot;
    __syncthreads();
  }
  if (bh_last_of_segment(seg_done, seg, n_cta)) {
    bh_scan_tiles_body(tile_cnt, A.ntiles, seg, tile_base, seg_n, n_vis);
    if (A.skip_thr > 0.0 && threadIdx.x == 0) {  // (thread 0 wrote seg_n[seg] = members counted)
      A.seg_low[seg] = seg_n[seg];
      seg_n[seg] = (long long)A.seg_cnt[(size_t)p * A.K + (c - 1)];
    }
  }
}

// q = sum(1 / (1:n)) of stats::p.adjust(method = "BY"): summed left to right for small n, asymptotic beyond
__device__ __forceinline__ double harmonic_number(long long n) {
  if (n < 128) {
    double s = 0.0;
    for (long long j = 1; j <= n; ++j) s += 1.0 / (double)j;
    return s;
  }
  const double x = (double)n, r = 1.0 / (x * x);
  return log(x) + 0.57721566490153286061 + 0.5 / x - r * (1.0 / 12.0 - r * (1.0 / 120.0 - r / 252.0));
}

// value that enters the cumulative min/max of stats::p.adjust, i = ascending rank inside the segment
// `num` = n for BH, q * n for BY (q = sum(1/(1:n)), hoisted out of the per-element path by adj_numerator)
__device__ __forceinline__ double adj_numerator(int mode, double n) {
  return mode == MDEGGS_PADJ_BY ? __dmul_rn(harmonic_number((long long)n), n) : n;
}
__device__ __forceinline__ double adj_value(int mode, double n, double num, long long i, double p) {
  if (mode == MDEGGS_PADJ_BH || mode == MDEGGS_PADJ_BY) return __dmul_rn(__ddiv_rn(num, (double)i), p);  // (n/i) p[o]
  if (mode == MDEGGS_PADJ_QVALUE) return __ddiv_rn(__dmul_rn(p, n), (double)i);  // qvalue: p[o] * m / i
  return __dmul_rn(__dsub_rn(__dadd_rn(n, 1.0), (double)i), p);  // holm / hochberg: (n + 1 - i) * p[o]
}

struct MinOp {
  __device__ __forceinline__ double operator()(double a, double b) const { return a < b ? a : b; }
};
struct MaxOp {
  __device__ __forceinline__ double operator()(double a, double b) const { return a > b ? a : b; }
};

// carry over the tiles of one segment (whole CTA): exclusive suffix-min walking the tiles from the end (BH, hochberg,
// BY) or exclusive prefix-max walking them from the start (holm)
__device__ __forceinline__ void bh_suffix_body(const double* __restrict__ tile_min, int ntiles, int mode, size_t seg,
                                               double* __restrict__ tile_carry) {
  typedef cub::BlockScan<double, BH_THREADS> Scan;
  __shared__ typename Scan::TempStorage tmp;
  __shared__ double carry;
  const bool fwd_max = mode == MDEGGS_PADJ_HOLM;
  const double ident = fwd_max ? -INFINITY : INFINITY;
  if (threadIdx.x == 0) carry = ident;
  __syncthreads();
  const size_t base = seg * ntiles;
  for (int t0 = 0; t0 < ntiles; t0 += BH_THREADS) {
    int r = t0 + threadIdx.x;                    // position in walking order
    int t = fwd_max ? r : ntiles - 1 - r;        // tile index
    double v = r < ntiles ? __ldcg(tile_min + base + t) : ident, ex, agg;
    if (fwd_max) Scan(tmp).ExclusiveScan(v, ex, ident, MaxOp(), agg);
    else Scan(tmp).ExclusiveScan(v, ex, ident, MinOp(), agg);
    if (r < ntiles) tile_carry[base + t] = fwd_max ? fmax(carry, ex) : fmin(carry, ex);
    __syncthreads();
    if (threadIdx.x == 0) carry = fwd_max ? fmax(carry, agg) : fmin(carry, agg);
    __syncthreads();
  }
}

// pass B: BH values with global ranks -> minimum of each tile; the last CTA of each segment carries them over
__global__ void __launch_bounds__(BH_THREADS) k_bh_tilemin(BhArgs A, int mode, const int32_t* __restrict__ tile_cnt,
                                                          const long long* __restrict__ tile_base,
                                                          const long long* __restrict__ seg_n,
                                                          double* __restrict__ tile_min,
                                                          double* __restrict__ tile_carry, int* __restrict__ seg_done) {
  int p, c;
  size_t seg;
  if (!bh_segment(A, p, c, seg)) return;
  if (A.tot && A.tot[p] == 0) return;  // (same whole-segment exit as pass A)
  typedef cub::BlockScan<int, BH_THREADS> Scan;
  typedef cub::BlockReduce<double, BH_THREADS> Reduce;
  __shared__ union {
    typename Scan::TempStorage scan;
    typename Reduce::TempStorage red;
  } tmp;
  const bool fwd_max = mode == MDEGGS_PADJ_HOLM;  // holm: cummax over ascending p; others: cummin from the largest p
  if (tile_cnt[seg * A.ntiles + blockIdx.x] == 0) {  // no member of this segment in the tile
    if (threadIdx.x == 0) tile_min[seg * A.ntiles + blockIdx.x] = fwd_max ? -INFINITY : INFINITY;
  } else {
    int flag[BH_ITEMS], rank[BH_ITEMS];
    double pv[BH_ITEMS];
    bh_load(A, p, c, (int64_t)blockIdx.x * BH_TILE, flag, pv);
    Scan(tmp.scan).InclusiveSum(flag, rank);
    __syncthreads();
    const long long base = tile_base[seg * A.ntiles + blockIdx.x];
    const double n = (double)seg_n[seg];
    double mn = fwd_max ? -INFINITY : INFINITY;
    const double num = adj_numerator(mode, n);
#pragma unroll
    for (int it = 0; it < BH_ITEMS; ++it)
      if (flag[it]) {
        double v = adj_value(mode, n, num, base + rank[it], pv[it]);
        mn = fwd_max ? fmax(mn, v) : fmin(mn, v);
      }
    double tmn = fwd_max ? Reduce(tmp.red).Reduce(mn, MaxOp()) : Reduce(tmp.red).Reduce(mn, MinOp());
    if (threadIdx.x == 0) tile_min[seg * A.ntiles + blockIdx.x] = tmn;
  }
  if (bh_last_of_segment(seg_done, seg, A.ntiles)) bh_suffix_body(tile_min, A.ntiles, mode, seg, tile_carry);
}

// Count-only adjustment (no adjusted value leaves the device for this percentile block: big jobs keep only the best
// percentile's row, which is recomputed afterwards).  The number of members with p.adj < 0.05 needs no cumulative
// min/max at all: with v_i the value that enters the cummin (BH, BY, hochberg, q.value; i = ascending rank) the
// adjusted value q_i = f(min_{j >= i} v_j) with f monotone (pmin(1, .), times pi0 for q.value), hence
//   q_i < 0.05  <=>  some j >= i has f(v_j) < 0.05  <=>  i <= max{ j : f(v_j) < 0.05 }    (the BH step-up rule),
// and for holm (cummax from the smallest p)  q_i < 0.05  <=>  i < min{ j : f(v_j) >= 0.05 }.
// So one pass after the rank pass finds, per segment, t_max = max rank with f(v) < 0.05 (holm: max of n + 1 - rank
// over the members with f(v) >= 0.05, count = n - t_max); the last CTA of the segment adds the count to sig[p] and
// re-arms the accumulator.  bonferroni: v = n p, same rule.  f(v) is evaluated exactly as k_bh_final evaluates q.
__global__ void __launch_bounds__(BH_THREADS) k_bh_sigrank(BhArgs A, int mode, const int32_t* __restrict__ tile_cnt,
                                                          const long long* __restrict__ tile_base,
                                                          const long long* __restrict__ seg_n,
                                                          unsigned long long* __restrict__ seg_rank,
                                                          unsigned long long* __restrict__ sig,
                                                          int* __restrict__ seg_done) {
  int p, c;
  size_t seg;
  if (!bh_segment(A, p, c, seg)) return;
  if (A.tot && A.tot[p] == 0) return;  // (same whole-segment exit as pass A)
  const int n_vis = bh_tiles_visited(A);
  const int n_cta = n_vis < (int)gridDim.x ? n_vis : (int)gridDim.x;  // (tile-stride walk, see k_bh_count)
  if ((int)blockIdx.x >= n_cta) return;
  typedef cub::BlockScan<int, BH_THREADS> Scan;
  typedef cub::BlockReduce<long long, BH_THREADS> Reduce;
  __shared__ union {
    typename Scan::TempStorage scan;
    typename Reduce::TempStorage red;
  } tmp;
  const bool holm = mode == MDEGGS_PADJ_HOLM;
  const long long n_ll = seg_n[seg];
  const double n = (double)n_ll;
  const double num = adj_numerator(mode, n);
  const double pi0 = mode == MDEGGS_PADJ_QVALUE ? A.seg_pi0[seg] : 1.0;
  long long t = 0;
  for (int tile = blockIdx.x; tile < n_vis; tile += gridDim.x) {
    if (tile_cnt[seg * A.ntiles + tile] == 0) continue;  // (same for every thread of the CTA)
    int flag[BH_ITEMS], rank[BH_ITEMS];
    double pv[BH_ITEMS];
    bh_load(A, p, c, (int64_t)tile * BH_TILE, flag, pv);
    Scan(tmp.scan).InclusiveSum(flag, rank);
    __syncthreads();
    const long long base = tile_base[seg * A.ntiles + tile];
#pragma unroll
    for (int it = 0; it < BH_ITEMS; ++it)
      if (flag[it]) {
        const long long i = base + rank[it];
        double q = mode == MDEGGS_PADJ_BONFERRONI ? __dmul_rn(n, pv[it]) : adj_value(mode, n, num, i, pv[it]);
        q = fmin(1.0, q);
        if (mode == MDEGGS_PADJ_QVALUE) q = __dmul_rn(pi0, q);
        const bool good = q < 0.05;  // (NaN pi0: never significant, as in k_bh_final)
        if (holm) {
          if (!good) t = max(t, n_ll + 1 - i);
        } else if (good) {
          t = max(t, i);
        }
      }
  }
  t = Reduce(tmp.red).Reduce(t, cub::Max());
  if (threadIdx.x == 0 && t > 0) atomicMax(seg_rank + seg, (unsigned long long)t);
  if (bh_last_of_segment(seg_done, seg, n_cta)) {
    if (threadIdx.x == 0) {
      long long tmax = (long long)__ldcg(seg_rank + seg);
      if (holm && A.skip_thr > 0.0) {  // the first member of the skipped tiles fails the test: rank counted + 1
        const long long low = A.seg_low[seg];
        if (low < n_ll) tmax = max(tmax, n_ll - low);
      }
      const long long cnt = holm ? n_ll - tmax : tmax;
      seg_rank[seg] = 0ull;
      if (cnt > 0) atomicAdd(&sig[p], (unsigned long long)cnt);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// hommel (stats::p.adjust, core_functions.R:1027-1035).  With p ascending inside a segment of n members R computes
//   pa = rep(min(n p / i), n);  for m = n-1 .. 2:  q1 = min_{j=2..m} m p[n-m+j] / j,
//        q[i] = min(m p[i], q1) for i <= n-m+1,  q[i] = q[n-m+1] beyond,  pa = pmax(pa, q);   result pmax(pa, p).
// The m iterations only meet in the running maximum, so they are independent:
//   pa[i] = max( init,  max_{m=2..min(n-1, n-i+1)} min(m p[i], q1[m]),  max_{m=n-i+2..n-1} c[m] ),
//   c[m] = min(m p[n-m+1], q1[m])   (1-based indices)
// O(n^2) work per segment in two kernels (q1/c per m, then pa per i, a warp each), a suffix maximum of c in between;
// the members of every segment are first copied into a dense array in rank order (k_hm_compact).
// ------------------------------------------------------------------------------------------------
// members of segment `seg` (pass A skips percentiles without tested edges: their seg_n entry is stale)
__device__ __forceinline__ long long hm_seg_n(const BhArgs& A, const long long* __restrict__ seg_n, int seg) {
  const int p = A.p0 + (seg / A.K) * A.p_stride;
  return (A.tot && A.tot[p] == 0) ? 0 : seg_n[seg];
}
__global__ void k_hm_offsets(BhArgs A, const long long* __restrict__ seg_n, int n_segments,
                             long long* __restrict__ seg_off) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  long long acc = 0;
  for (int sgi = 0; sgi < n_segments; ++sgi) {
    seg_off[sgi] = acc;
    acc += hm_seg_n(A, seg_n, sgi);
  }
  seg_off[n_segments] = acc;
}

__global__ void __launch_bounds__(BH_THREADS) k_hm_compact(BhArgs A, const int32_t* __restrict__ tile_cnt,
                                                          const long long* __restrict__ tile_base,
                                                          const long long* __restrict__ seg_off,
                                                          double* __restrict__ hp, int32_t* __restrict__ hidx) {
  int p, c;
  size_t seg;
  if (!bh_segment(A, p, c, seg)) return;
  if (A.tot && A.tot[p] == 0) return;
  if (tile_cnt[seg * A.ntiles + blockIdx.x] == 0) return;
  typedef cub::BlockScan<int, BH_THREADS> Scan;
  __shared__ typename Scan::TempStorage tmp;
  int flag[BH_ITEMS], rank[BH_ITEMS];
  double pv[BH_ITEMS];
  const int64_t tile0 = (int64_t)blockIdx.x * BH_TILE;
  bh_load(A, p, c, tile0, flag, pv);
  Scan(tmp).InclusiveSum(flag, rank);
  const long long base = seg_off[seg] + tile_base[seg * A.ntiles + blockIdx.x];
#pragma unroll
  for (int it = 0; it < BH_ITEMS; ++it)
    if (flag[it]) {
      const long long pos = base + rank[it] - 1;
      hp[pos] = pv[it];
      hidx[pos] = A.sidx[tile0 + (int64_t)threadIdx.x * BH_ITEMS + it];
    }
}

__device__ __forceinline__ double warp_min_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_max_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// one warp per m = 2 .. n-1 of a segment: q1[m], c[m]; CTA 0 of the segment also computes init = min(n p / i)
__global__ void __launch_bounds__(256) k_hm_q1(BhArgs A, const long long* __restrict__ seg_n,
                                               const long long* __restrict__ seg_off, const double* __restrict__ hp,
                                               double* __restrict__ hq1, double* __restrict__ hc,
                                               double* __restrict__ hinit) {
  const int seg = blockIdx.y;
  const long long n = hm_seg_n(A, seg_n, seg);
  const double* __restrict__ p = hp + seg_off[seg];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const long long m = 2 + (long long)blockIdx.x * 8 + wib;
  if (m <= n - 1) {
    const double md = (double)m;
    double q1 = INFINITY;
    for (long long j = 2 + lane; j <= m; j += 32)  // p[i2[j-2]], i2 = (n-m+2):n, 0-based index n-m+j-1
      q1 = fmin(q1, __ddiv_rn(__dmul_rn(md, p[n - m + j - 1]), (double)j));
    q1 = warp_min_d(q1);
    if (lane == 0) {
      hq1[seg_off[seg] + m] = q1;
      hc[seg_off[seg] + m] = fmin(__dmul_rn(md, p[n - m]), q1);  // q[n-m+1] of this m
    }
  }
  if (blockIdx.x == 0) {
    __shared__ double s_w[8];
    const double nd = (double)n;
    double v = INFINITY;
    for (long long i = threadIdx.x; i < n; i += 256) v = fmin(v, __ddiv_rn(__dmul_rn(nd, p[i]), (double)(i + 1)));
    v = warp_min_d(v);
    if (lane == 0) s_w[wib] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < 8; ++w) v = fmin(v, s_w[w]);
      hinit[seg] = v;
    }
  }
}

// suffix maximum of c[m] over m (one CTA per segment): hsc[m] = max_{m' = m .. n-1} c[m']
__global__ void __launch_bounds__(BH_THREADS) k_hm_sufmax(BhArgs A, const long long* __restrict__ seg_n,
                                                         const long long* __restrict__ seg_off,
                                                         const double* __restrict__ hc, double* __restrict__ hsc) {
  typedef cub::BlockScan<double, BH_THREADS> Scan;
  __shared__ typename Scan::TempStorage tmp;
  __shared__ double carry;
  const int seg = blockIdx.x;
  const long long n = hm_seg_n(A, seg_n, seg), off = seg_off[seg];
  if (threadIdx.x == 0) carry = -INFINITY;
  __syncthreads();
  const long long cnt = n - 2;  // m = 2 .. n-1, walked from n-1 down
  for (long long r0 = 0; r0 < cnt; r0 += BH_THREADS) {
    const long long r = r0 + threadIdx.x, m = n - 1 - r;
    double v = r < cnt ? hc[off + m] : -INFINITY, inc, agg;
    Scan(tmp).InclusiveScan(v, inc, MaxOp(), agg);
    if (r < cnt) hsc[off + m] = fmax(carry, inc);
    __syncthreads();
    if (threadIdx.x == 0) carry = fmax(carry, agg);
    __syncthreads();
  }
}

// one warp per member i: pa[i], the adjusted value, its count < 0.05 and the scatter to network order
__global__ void __launch_bounds__(256) k_hm_pa(BhArgs A, const long long* __restrict__ seg_n,
                                               const long long* __restrict__ seg_off, const double* __restrict__ hp,
                                               const double* __restrict__ hq1, const double* __restrict__ hsc,
                                               const double* __restrict__ hinit, const int32_t* __restrict__ hidx,
                                               unsigned long long* __restrict__ sig, double* __restrict__ padj_out,
                                               int64_t padj_stride) {
  const int seg = blockIdx.y;
  const int pct = A.p0 + (seg / A.K) * A.p_stride;
  const long long n = hm_seg_n(A, seg_n, seg), off = seg_off[seg];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const long long i0 = (long long)blockIdx.x * 8 + wib;  // 0-based member
  if (i0 >= n) return;
  const double pi = hp[off + i0];
  double res = pi;  // n <= 1: p.adjust returns p unchanged
  if (n > 1) {
    const long long i = i0 + 1;
    const long long m_hi = (n - 1 < n - i + 1) ? n - 1 : n - i + 1;
    double t = -INFINITY;
    for (long long m = 2 + lane; m <= m_hi; m += 32) t = fmax(t, fmin(__dmul_rn((double)m, pi), hq1[off + m]));
    t = warp_max_d(t);
    double pa = fmax(hinit[seg], t);
    if (n - i + 2 <= n - 1 && n - i + 2 >= 2) pa = fmax(pa, hsc[off + n - i + 2]);
    res = fmax(pa, pi);
  }
  if (lane == 0) {
    if (padj_out) padj_out[(int64_t)pct * padj_stride + hidx[off + i0]] = res;
    if (res < 0.05) atomicAdd(&sig[pct], 1ull);
  }
}

// q.value: pi0 of every segment (qvalue::pi0est, smoother method; core_functions.R:1012-1025)
//   pi0(lambda_j) = mean(p >= lambda_j) / (1 - lambda_j), lambda = seq(0.05, 0.95, 0.05);
//   pi0 = min(1, value at the last lambda of smooth.spline(lambda, pi0(lambda), df = 3)) = QV_ROW . pi0(lambda);
//   qvalue() stops when there are fewer than two p-values, when max(p) < max(lambda) or when pi0 <= 0: the reference
//   then retries with pi0 = 1 if the category network has more than one row, else returns NA.
// The members of a segment below lambda_j are counted from the tile offsets of pass A plus one partial tile.
__global__ void __launch_bounds__(BH_THREADS) k_qv_pi0(BhArgs A, const long long* __restrict__ tile_base,
                                                      const long long* __restrict__ seg_n,
                                                      double* __restrict__ seg_pi0) {
  int p, c;
  size_t seg;
  if (!bh_segment(A, p, c, seg)) return;
  if (A.tot && A.tot[p] == 0) {
    if (threadIdx.x == 0) seg_pi0[seg] = nan("");
    return;
  }
  __shared__ long long s_J[QV_NLAMBDA], s_lt[QV_NLAMBDA];
  __shared__ int s_part;
  const int64_t nv = (int64_t)*A.n_valid;
  if (threadIdx.x < QV_NLAMBDA) {  // first sorted position whose p is >= lambda_j
    const double lam = QV_LAMBDA[threadIdx.x];
    int64_t lo = 0, hi = nv;
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if (A.ps[mid] < lam) lo = mid + 1;
      else hi = mid;
    }
    s_J[threadIdx.x] = lo;
  }
  if (threadIdx.x == 0) s_part = 0;
  __syncthreads();
  auto member = [&](int64_t q) {
    return edge_category(A.act[(size_t)p * A.act_stride + A.sa[q]], A.act[(size_t)p * A.act_stride + A.sb[q]]) == c;
  };
  for (int j = 0; j <= QV_NLAMBDA; ++j) {  // j == QV_NLAMBDA: members whose p is NaN (sorted last)
    const int64_t from = j < QV_NLAMBDA ? (s_J[j] / BH_TILE) * BH_TILE : nv;
    const int64_t to = j < QV_NLAMBDA ? s_J[j] : (int64_t)*A.n_order;
    int cnt = 0;
    for (int64_t q = from + threadIdx.x; q < to; q += BH_THREADS) cnt += member(q);
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(&s_part, cnt);
    __syncthreads();
    if (threadIdx.x == 0) {
      if (j < QV_NLAMBDA) {
        const int64_t tile = s_J[j] / BH_TILE;
        s_lt[j] = (tile < A.ntiles ? tile_base[seg * A.ntiles + tile] : seg_n[seg]) + s_part;
      } else {
        // thread 0 finishes: pi0 and the fallback rule
        const long long m = seg_n[seg], m_all = m + s_part;
        double pi0 = 0.0;
        bool ok = m >= 2 && (m - s_lt[QV_NLAMBDA - 1]) > 0;  // at least two p-values, max(p) >= max(lambda)
        if (ok) {
          for (int l = 0; l < QV_NLAMBDA; ++l) {
            const double frac = __ddiv_rn((double)(m - s_lt[l]), (double)m);
            pi0 = fma(QV_ROW[l], __ddiv_rn(frac, __dsub_rn(1.0, QV_LAMBDA[l])), pi0);
          }
          pi0 = fmin(pi0, 1.0);
          ok = pi0 > 0.0;
        }
        if (!ok) pi0 = m_all > 1 ? 1.0 : nan("");
        seg_pi0[seg] = pi0;
      }
      s_part = 0;
    }
    __syncthreads();
  }
}

// which.max(sig_pvalues) over non-NULL percentiles (core_functions.R:475-495): first maximum.  Also packs every small
// result (counts, cut-offs, best, node count, error flag, degrees of freedom) into one contiguous block so that they
// leave the device in a single copy:  i64 tot[P] sig[P] fail[P] | f64 cutoff[P] | i32 best nn err 0 | f64 df1 df2
struct SelArgs {
  const unsigned long long *tot, *fail, *sig;
  int P, host_adjust;
  const double* cutoff;
  const int32_t *n_nodes, *err;
  double df1, df2;
  long long *blk, *host_blk;
  int32_t* best;
  unsigned int* done;  // k_bh_final: "CTAs done" counter of the fused tail (nullptr: selection is a launch of its own)
};

// one warp: lanes stride over the percentiles, the first maximum is a shuffle arg-max; every word of the block is
// written to device memory and (host_blk: mapped pinned memory) straight into the caller's process by its lane.
// Loads bypass L1: inside k_bh_final the counts were produced by other CTAs of the same launch.
__device__ __forceinline__ void select_best_warp(const SelArgs& S) {
  const int lane = threadIdx.x & 31;
  const int P = S.P, Pn = P > 0 ? P : 1;
  auto put = [&](int word, long long v) {
    S.blk[word] = v;
    if (S.host_blk) S.host_blk[word] = v;
  };
  long long bv = -1;
  int bp = 0x7fffffff;
  for (int p = lane; p < P; p += 32) {
    const long long t = (long long)__ldcg(S.tot + p), f = (long long)__ldcg(S.fail + p);
    long long s = t > 0 ? (long long)__ldcg(S.sig + p) : -1;
    if (S.host_adjust) s = -1;
    put(p, t);
    put(Pn + p, s);
    put(2 * Pn + p, f);
    put(3 * Pn + p, __double_as_longlong(__ldcg(S.cutoff + p)));
    if (!S.host_adjust && t > 0 && f == 0 && s > bv) {  // p ascending per lane: keeps the lane's first maximum
      bv = s;
      bp = p;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {  // maximum count, ties -> smallest percentile (which.max: first maximum)
    const long long ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int op = __shfl_xor_sync(0xffffffffu, bp, o);
    if (ov > bv || (ov == bv && op < bp)) {
      bv = ov;
      bp = op;
    }
  }
  const int b = bv >= 0 ? bp + 1 : 0;
  if (lane == 0) {
    *S.best = b;
    put(4 * Pn, (long long)(unsigned int)b | ((long long)(unsigned int)__ldcg(S.n_nodes) << 32));  // i32 best, nn
    put(4 * Pn + 1, (long long)(unsigned int)__ldcg(S.err));                                        // i32 err, 0
    put(4 * Pn + 2, __double_as_longlong(S.df1));
    put(4 * Pn + 3, __double_as_longlong(S.df2));
  }
  if (S.host_blk) __threadfence_system();
}

__global__ void k_select_best(SelArgs S) {
  if (blockIdx.x != 0 || threadIdx.x >= 32) return;
  select_best_warp(S);
}

// pass C: adjusted values, count < 0.05, optional scatter to network order
//   mode 1 bonferroni: pmin(1, n * p);  mode 2 BH: pmin(1, cummin((n/i) * p[o]))[ro]
__device__ __forceinline__ void bh_final_body(const BhArgs& A, int mode, const int32_t* __restrict__ tile_cnt,
                                              const long long* __restrict__ tile_base,
                                              const long long* __restrict__ seg_n,
                                              const double* __restrict__ tile_carry,
                                              unsigned long long* __restrict__ sig, double* __restrict__ padj_out,
                                              int64_t padj_stride) {
  int p, c;
  size_t seg;
  if (!bh_segment(A, p, c, seg)) return;
  if (tile_cnt[seg * A.ntiles + blockIdx.x] == 0) return;
  typedef cub::BlockScan<int, BH_THREADS> ScanI;
  typedef cub::BlockScan<double, BH_THREADS> ScanD;
  typedef cub::BlockReduce<int, BH_THREADS> Reduce;
  __shared__ union {
    typename ScanI::TempStorage si;
    typename ScanD::TempStorage sd;
    typename Reduce::TempStorage red;
  } tmp;
  int flag[BH_ITEMS], rank[BH_ITEMS];
  double pv[BH_ITEMS], q[BH_ITEMS];
  const int64_t tile0 = (int64_t)blockIdx.x * BH_TILE;
  bh_load(A, p, c, tile0, flag, pv);
  const double n = (double)seg_n[seg];
  if (mode != MDEGGS_PADJ_BONFERRONI) {
    const bool fwd_max = mode == MDEGGS_PADJ_HOLM;
    const double ident = fwd_max ? -INFINITY : INFINITY;
    ScanI(tmp.si).InclusiveSum(flag, rank);
    __syncthreads();
    const long long base = tile_base[seg * A.ntiles + blockIdx.x];
    double v[BH_ITEMS], run_it[BH_ITEMS];
    const double num = adj_numerator(mode, n);
#pragma unroll
    for (int it = 0; it < BH_ITEMS; ++it) v[it] = flag[it] ? adj_value(mode, n, num, base + rank[it], pv[it]) : ident;
    // running min (from the thread's last item backwards) or max (forwards) over the thread's own items
    double run = ident;
    if (fwd_max) {
#pragma unroll
      for (int it = 0; it < BH_ITEMS; ++it) {
        run = fmax(run, v[it]);
        run_it[it] = run;
      }
    } else {
#pragma unroll
      for (int it = BH_ITEMS - 1; it >= 0; --it) {
        run = fmin(run, v[it]);
        run_it[it] = run;
      }
    }
    // aggregate of the threads before (holm) / after (others) this one: exclusive scan in walking order
    __shared__ double tagg_s[BH_THREADS];
    const int wpos = fwd_max ? threadIdx.x : BH_THREADS - 1 - threadIdx.x;
    tagg_s[wpos] = run;
    __syncthreads();
    double mine = tagg_s[threadIdx.x], ex;
    if (fwd_max) ScanD(tmp.sd).ExclusiveScan(mine, ex, ident, MaxOp());
    else ScanD(tmp.sd).ExclusiveScan(mine, ex, ident, MinOp());
    __syncthreads();
    tagg_s[threadIdx.x] = ex;
    __syncthreads();
    const double others = tagg_s[wpos];
    const double carry = tile_carry[seg * A.ntiles + blockIdx.x];
#pragma unroll
    for (int it = 0; it < BH_ITEMS; ++it)
      q[it] = fwd_max ? fmin(1.0, fmax(run_it[it], fmax(others, carry))) : fmin(1.0, fmin(run_it[it], fmin(others, carry)));
    if (mode == MDEGGS_PADJ_QVALUE) {  // qvalues = pi0 * pmin(1, cummin(p[o] m / i)); pi0 NaN: the reference returns NA
      const double pi0 = A.seg_pi0[seg];
#pragma unroll
      for (int it = 0; it < BH_ITEMS; ++it) q[it] = __dmul_rn(pi0, q[it]);
    }
    __syncthreads();
  } else {
#pragma unroll
    for (int it = 0; it < BH_ITEMS; ++it) q[it] = fmin(1.0, __dmul_rn(n, pv[it]));
  }
  int cnt = 0;
#pragma unroll
  for (int it = 0; it < BH_ITEMS; ++it) {
    if (!flag[it]) continue;
    cnt += (q[it] < 0.05);
    if (padj_out) {
      int64_t j = tile0 + (int64_t)threadIdx.x * BH_ITEMS + it;
      int64_t row = A.best ? 0 : p;  // full matrix rows are indexed by the absolute percentile
      padj_out[row * padj_stride + A.sidx[j]] = q[it];
    }
  }
  if (sig) {
    int tot = Reduce(tmp.red).Sum(cnt);
    if (threadIdx.x == 0 && tot) atomicAdd(&sig[p], (unsigned long long)tot);
  }
}

// pass C; with S.done the CTA that finishes last (all significant counts are in) also runs the percentile selection
__global__ void __launch_bounds__(BH_THREADS) k_bh_final(BhArgs A, int mode, const int32_t* __restrict__ tile_cnt,
                                                        const long long* __restrict__ tile_base,
                                                        const long long* __restrict__ seg_n,
                                                        const double* __restrict__ tile_carry,
                                                        unsigned long long* __restrict__ sig,
                                                        double* __restrict__ padj_out, int64_t padj_stride, SelArgs S) {
  bh_final_body(A, mode, tile_cnt, tile_base, seg_n, tile_carry, sig, padj_out, padj_stride);
  if (!S.done) return;
  __shared__ int s_last_cta;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int prev = atomicAdd(S.done, 1u);
    s_last_cta = prev == gridDim.x * gridDim.y - 1;
  }
  __syncthreads();
  if (!s_last_cta || threadIdx.x >= 32) return;
  __threadfence();
  select_best_warp(S);
}

// K8 in ONE launch for edge lists of up to BH1_MAX_TILES tiles (config 3: 26): count, rank, adjusted value, cumulative
// min / max and the significant count of every (percentile, category) segment, with the percentile selection in the
// CTA that finishes last.  The three-launch form above costs three dependent grid launches plus three "last CTA of the
// segment" tails of pure latency.  Here the segment sizes n come from k_percolate, tiles are walked from the END of the
// p order (from the start for holm), and a CTA only needs the aggregates of the tiles walked before it: it publishes
// (members + 1) and later the sortable key of its extreme value in one word each (0 = not there yet), and one warp
// reads all predecessors' words at once.  A CTA waits only for CTAs with a smaller blockIdx.x, which the hardware
// dispatched earlier: no deadlock whatever the residency.
constexpr int BH1_MAX_TILES = 64;
__global__ void __launch_bounds__(BH_THREADS) k_bh_onepass(BhArgs A, int mode, const unsigned int* __restrict__ seg_cnt,
                                                          unsigned int* __restrict__ st_cnt,
                                                          unsigned long long* __restrict__ st_ext,
                                                          unsigned long long* __restrict__ sig,
                                                          double* __restrict__ padj_out, int64_t padj_stride,
                                                          SelArgs S) {
  typedef cub::BlockScan<int, BH_THREADS> ScanI;
  typedef cub::BlockScan<double, BH_THREADS> ScanD;
  typedef cub::BlockReduce<int, BH_THREADS> ReduceI;
  typedef cub::BlockReduce<double, BH_THREADS> ReduceD;
  __shared__ union {
    typename ScanI::TempStorage si;
    typename ScanD::TempStorage sd;
    typename ReduceI::TempStorage ri;
    typename ReduceD::TempStorage rd;
  } tmp;
  __shared__ double tagg_s[BH_THREADS];
  __shared__ long long s_before;
  __shared__ double s_carry;
  __shared__ int s_tile_cnt;
  const bool fwd = mode == MDEGGS_PADJ_HOLM;  // holm: cummax over ascending p; others: cummin from the largest p
  const double ident = fwd ? -INFINITY : INFINITY;
  const int w = blockIdx.x;                       // position in walking order
  const int tile = fwd ? w : A.ntiles - 1 - w;    // tile in p order
  const int p = A.p0 + ((int)blockIdx.y / A.K) * A.p_stride, c = (int)blockIdx.y % A.K + 1;
  const size_t sbase = (size_t)blockIdx.y * A.ntiles;
  const long long n_ll = (long long)seg_cnt[p * A.K + c - 1];
  if (A.tot[p] != 0 && n_ll > 0) {
    int flag[BH_ITEMS], rank[BH_ITEMS];
    double pv[BH_ITEMS], q[BH_ITEMS];
    const int64_t tile0 = (int64_t)tile * BH_TILE;
    bh_load(A, p, c, tile0, flag, pv);
    int tcnt;
    ScanI(tmp.si).InclusiveSum(flag, rank, tcnt);
    if (threadIdx.x == 0) {
      s_tile_cnt = tcnt;
      __stcg(st_cnt + sbase + w, (unsigned int)tcnt + 1u);
      if (tcnt == 0) __stcg(st_ext + sbase + w, dkey(ident));  // nothing to wait for in this tile
    }
    __syncthreads();
    if (s_tile_cnt != 0) {
      const double n = (double)n_ll;
      if (mode != MDEGGS_PADJ_BONFERRONI) {
        // members of the tiles walked before this one
        if (threadIdx.x < 32) {
          long long acc = 0;
          for (int j = threadIdx.x; j < w; j += 32) {
            unsigned int v;
            while ((v = __ldcg(st_cnt + sbase + j)) == 0u) __nanosleep(20);
            acc += (long long)(v - 1u);
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
          if (threadIdx.x == 0) s_before = acc;
        }
        __syncthreads();
        // ascending rank inside the segment
        const long long base = fwd ? s_before : n_ll - s_before - (long long)s_tile_cnt;
        double v[BH_ITEMS], run_it[BH_ITEMS];
        const double num = adj_numerator(mode, n);
#pragma unroll
        for (int it = 0; it < BH_ITEMS; ++it) v[it] = flag[it] ? adj_value(mode, n, num, base + rank[it], pv[it]) : ident;
        double run = ident;
        if (fwd) {
#pragma unroll
          for (int it = 0; it < BH_ITEMS; ++it) {
            run = fmax(run, v[it]);
            run_it[it] = run;
          }
        } else {
#pragma unroll
          for (int it = BH_ITEMS - 1; it >= 0; --it) {
            run = fmin(run, v[it]);
            run_it[it] = run;
          }
        }
        // this tile's extreme value -> published; exclusive scan of the thread aggregates in walking order
        const int wpos = fwd ? threadIdx.x : BH_THREADS - 1 - threadIdx.x;
        tagg_s[wpos] = run;
        __syncthreads();
        double mine = tagg_s[threadIdx.x], ex, agg;
        if (fwd) ScanD(tmp.sd).ExclusiveScan(mine, ex, ident, MaxOp(), agg);
        else ScanD(tmp.sd).ExclusiveScan(mine, ex, ident, MinOp(), agg);
        if (threadIdx.x == 0) __stcg(st_ext + sbase + w, dkey(agg));
        __syncthreads();
        tagg_s[threadIdx.x] = ex;
        __syncthreads();
        const double others = tagg_s[wpos];
        // extreme value of the tiles walked before this one
        if (threadIdx.x < 32) {
          double acc = ident;
          for (int j = threadIdx.x; j < w; j += 32) {
            unsigned long long k;
            while ((k = __ldcg(st_ext + sbase + j)) == 0ull) __nanosleep(20);
            const double x = dkey_inv(k);
            acc = fwd ? fmax(acc, x) : fmin(acc, x);
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            const double x = __shfl_xor_sync(0xffffffffu, acc, o);
            acc = fwd ? fmax(acc, x) : fmin(acc, x);
          }
          if (threadIdx.x == 0) s_carry = acc;
        }
        __syncthreads();
        const double carry = s_carry;
#pragma unroll
        for (int it = 0; it < BH_ITEMS; ++it)
          q[it] = fwd ? fmin(1.0, fmax(run_it[it], fmax(others, carry))) : fmin(1.0, fmin(run_it[it], fmin(others, carry)));
      } else {
#pragma unroll
        for (int it = 0; it < BH_ITEMS; ++it) q[it] = fmin(1.0, __dmul_rn(n, pv[it]));
      }
      int cnt = 0;
#pragma unroll
      for (int it = 0; it < BH_ITEMS; ++it) {
        if (!flag[it]) continue;
        cnt += (q[it] < 0.05);
        if (padj_out) padj_out[(int64_t)p * padj_stride + A.sidx[tile0 + (int64_t)threadIdx.x * BH_ITEMS + it]] = q[it];
      }
      __syncthreads();
      const int tot = ReduceI(tmp.ri).Sum(cnt);
      if (threadIdx.x == 0 && tot) atomicAdd(&sig[p], (unsigned long long)tot);
    }
  }
  if (!S.done) return;
  __shared__ int s_last_cta;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int prev = atomicAdd(S.done, 1u);
    s_last_cta = prev == gridDim.x * gridDim.y - 1;
  }
  __syncthreads();
  if (!s_last_cta || threadIdx.x >= 32) return;
  __threadfence();
  select_best_warp(S);
}

// results -> caller's DEVICE buffers: every requested array in one launch (entry = blockIdx.y; src == nullptr: zeros)
constexpr int EMIT_MAX = 16;
enum { EMIT_COPY = 0, EMIT_CAT_BEST = 1, EMIT_PADJ_BEST = 2 };
struct EmitTab {
  void* dst[EMIT_MAX];
  const void* src[EMIT_MAX];
  long long bytes[EMIT_MAX];
  int kind[EMIT_MAX];
  // rows of the device-chosen percentile, produced straight into the caller's buffers (what k_best_rows does for
  // host callers): EMIT_CAT_BEST from the masks, EMIT_PADJ_BEST from row best-1 of the P x E scratch (src)
  const int32_t *ea, *eb, *best, *err;
  const uint8_t* act;
  int act_stride;
  long long E;
};
__global__ void __launch_bounds__(256) k_emit(EmitTab T) {
  const int ent = blockIdx.y;
  char* dst = static_cast<char*>(T.dst[ent]);
  const char* src = static_cast<const char*>(T.src[ent]);
  const long long bytes = T.bytes[ent];
  const long long stride = (long long)gridDim.x * blockDim.x, t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (T.kind[ent] != EMIT_COPY) {
    if (*T.err) return;
    const int p = *T.best - 1;
    if (T.kind[ent] == EMIT_CAT_BEST) {
      const uint8_t* row = T.act + (size_t)(p < 0 ? 0 : p) * T.act_stride;
      for (long long e = t0; e < T.E; e += stride)
        dst[e] = p >= 0 ? (char)edge_category(row[T.ea[e]], row[T.eb[e]]) : (char)0;
    } else {
      const double* all = reinterpret_cast<const double*>(src);
      for (long long e = t0; e < T.E; e += stride)
        reinterpret_cast<double*>(dst)[e] = (p >= 0 && all) ? all[(size_t)p * T.E + e] : nan("");
    }
    return;
  }
  const bool vec = ((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src)) & 15) == 0;
  long long done = 0;
  if (vec) {
    const long long nv = bytes >> 4;
    for (long long i = t0; i < nv; i += stride)
      reinterpret_cast<uint4*>(dst)[i] = src ? reinterpret_cast<const uint4*>(src)[i] : make_uint4(0, 0, 0, 0);
    done = nv << 4;
  }
  for (long long i = done + t0; i < bytes; i += stride) dst[i] = src ? src[i] : 0;
}

__global__ void k_fill_f64(double* __restrict__ x, int64_t n, double v) {
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) x[i] = v;
}

// median[k*G + g] for the caller (NaN for rows that are not nodes)
__global__ void k_scatter_median(const double* __restrict__ med, const int32_t* __restrict__ node_of_row, int G,
                                 int K, double* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)G * K) return;
  int k = (int)(i / G), g = (int)(i % G);
  int node = node_of_row[g];
  out[i] = node >= 0 ? med[(size_t)node * K + k] : nan("");
}

// predict.multiDEGGs_filter feature construction (feature_selection_for_ML.R:423-431)
__global__ void k_pair_features(const double* __restrict__ x, int64_t ldx, int n, const int32_t* __restrict__ a,
                                const int32_t* __restrict__ b, int n_pairs, int ratio, double* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)n * n_pairs) return;
  int j = (int)(i / n), s = (int)(i % n);
  double va = x[s + (int64_t)a[j] * ldx], vb = x[s + (int64_t)b[j] * ldx];
  double r = ratio ? __ddiv_rn(va, vb) : __dmul_rn(va, vb);
  out[i] = isfinite(r) ? r : 0.0;
}

// per-edge status codes travel between the ranks as bytes (a quarter of the all-gather's status payload)
__global__ void k_pack_status(const int32_t* __restrict__ status, int64_t n, int8_t* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (int8_t)status[i];
}
__global__ void k_unpack_status(const int8_t* __restrict__ in, int64_t n, int32_t* __restrict__ status) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) status[i] = (int32_t)in[i];
}

// Per-category sufficient statistics of gene pairs from the gene-major matrix and the per-gene moments that the last
// run left on the device: what the per-category regression lines of plot_regressions need
// (plotting_functions.R:125-162: lm(y ~ x) inside each category, predict(interval = "confidence")).  One warp per pair.
// out[(j K + k) 6 + 0..5] = n_k, mean_A, mean_B, Sxx_A, Syy_B, Sxy;  NaN when a row is no network node of that run.
__global__ void __launch_bounds__(256) k_pair_moments(const double* __restrict__ D, SegLayout L,
                                                      const int32_t* __restrict__ node_of_row, int G,
                                                      const int32_t* __restrict__ a_rows,
                                                      const int32_t* __restrict__ b_rows, int n_pairs,
                                                      const double* __restrict__ mean, const double* __restrict__ sxx,
                                                      double* __restrict__ out) {
  const int j = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (j >= n_pairs) return;
  const int ra = a_rows[j] - 1, rb = b_rows[j] - 1;
  const int na = (ra >= 0 && ra < G) ? node_of_row[ra] : -1, nb = (rb >= 0 && rb < G) ? node_of_row[rb] : -1;
  for (int k = 0; k < L.K; ++k) {
    double* o = out + ((size_t)j * L.K + k) * 6;
    if (na < 0 || nb < 0) {
      if (lane < 6) o[lane] = nan("");
      continue;
    }
    const double ma = mean[(size_t)na * L.K + k], mb = mean[(size_t)nb * L.K + k];
    const double* __restrict__ pa = D + (size_t)na * L.npad;
    const double* __restrict__ pb = D + (size_t)nb * L.npad;
    double acc = 0.0;
    for (int i = L.off[k] + lane; i < L.off[k] + L.cnt[k]; i += 32) acc = fma(pa[i] - ma, pb[i] - mb, acc);
    acc = warp_sum(acc);
    if (lane == 0) {
      o[0] = (double)L.cnt[k];
      o[1] = ma;
      o[2] = mb;
      o[3] = sxx[(size_t)na * L.K + k];
      o[4] = sxx[(size_t)nb * L.K + k];
      o[5] = acc;
    }
  }
}

// FP64 FMA peak of this GPU (denominator of the rlm roofline, SURVEY §8 d): 16 independent DFMA chains per thread
__global__ void __launch_bounds__(256) k_dfma_peak(int iters, double seed, double* __restrict__ sink) {
  double a[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) a[j] = seed + (double)(threadIdx.x + j);
  const double m = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 16; ++j) a[j] = fma(a[j], m, c);
  }
  double t = 0.0;
#pragma unroll
  for (int j = 0; j < 16; ++j) t += a[j];
  if (t == 12345.678) sink[0] = t;  // never true: keeps the chains alive
}

// L2 read bandwidth of this GPU for an L2-resident working set (denominator of the pair-fit roofline: the gene-major
// matrix D, <= 67 MB, stays in the 126 MB L2 while k_fit_stream re-reads it ~12 times).  Every thread keeps 4
// independent 16-byte loads in flight and bypasses L1 (ld.global.cg), so the figure is L2 -> SM throughput only.
__device__ __forceinline__ uint4 l2_ld_cg(const uint4* p) {
  uint4 v;
  asm volatile("ld.global.cg.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__global__ void __launch_bounds__(256) k_l2_read(const uint4* __restrict__ buf, size_t n_vec, int passes,
                                                 unsigned int* __restrict__ sink) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  unsigned int acc = 0;
  for (int r = 0; r < passes; ++r) {
    // every pass starts at another CTA's offset: an SM does not re-read what it read in the pass before
    size_t i = (((size_t)blockIdx.x + (size_t)r * 37) % gridDim.x) * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n_vec; i += 4 * stride) {
      const uint4 a = l2_ld_cg(buf + i), b = l2_ld_cg(buf + i + stride), c = l2_ld_cg(buf + i + 2 * stride),
                  d = l2_ld_cg(buf + i + 3 * stride);
      acc += a.x ^ b.y ^ c.z ^ d.w;
    }
    for (; i < n_vec; i += stride) acc += l2_ld_cg(buf + i).x;
  }
  if (acc == 0x12345678u) sink[0] = acc;  // (practically) never true: keeps the loads alive
}

// scalar math twins for tests
__global__ void k_dev_math(int which, const double* __restrict__ x, const double* __restrict__ a,
                           const double* __restrict__ b, int64_t n, double* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (which == 0) out[i] = pt_dev(x[i], a[i]);
  else if (which == 1) out[i] = pf_upper_dev(x[i], a[i], b[i]);
  else out[i] = p_two_sided_ref_dev(x[i], a[i]);
}

// the run-table form of the two-sided t tail (series / continued fractions from k_cf_tables, and - tc.sp set - the
// interpolation table of k_tail_spline), for tests
__global__ void k_dev_tail_tab(TailConsts tc, const double* __restrict__ lbeta_p, double df, const double* __restrict__ x,
                               int64_t n, double* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  tc.lbeta_ab = *lbeta_p;
  out[i] = p_two_sided_ref_tab(tc, x[i], df);
}

}  // namespace mdeggs

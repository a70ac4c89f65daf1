// mdeggs_qsel.cuh — K3: exact multi-quantile selection over the whole node-filtered matrix without a sort, and the
// K1 gather kernels that carry its first pass.
//
// Replaces  cut_off <- stats::quantile(as.matrix(assayData), prob = percentile, na.rm = TRUE)
// (R/core_functions.R:718), which the reference recomputes for each of the 13-15 percentiles.  Type 7 needs two
// order statistics per percentile (floor/ceil of 1 + (N-1) p): T = 2P target ranks out of N ~ 8.4e6 values.
//
//   S0 k_qs_sample  one CTA: a stratified sample of the node rows of the input matrix, sorted in shared memory ->
//                   65 knots at sample quantiles (uniform in the middle, halving towards both tails).  Knot cell c,
//                   split into 64 equal-width sub-bins, defines a MONOTONE map value -> bin in [0, 4096) evaluated
//                   in FP32 on (x - sample minimum); bins hold roughly N/4096 values whatever the distribution, and
//                   a value that fills a whole cell of the sample gets a bin of its own.
//   S1 (fused into the K1 gather kernels) exact 4,096-bin histogram of everything that is written to D,
//                   shared-memory privatised per persistent CTA; NaN skipped (na.rm = TRUE).
//   S2 qs_scan_body run by the last CTA of the gather: prefix sums; each target rank -> (bin, rank inside the bin); the <= 2P distinct bins
//                   become "slots" with exact candidate counts and offsets in the candidate buffer.
//   S3 k_qs_compact reads the 2-byte bin id that the gather kept for every value (a quarter of D) and appends the values
//                   of slot bins to their slot's list (staged per CTA in shared memory: one global atomic per CTA and
//                   slot).  Slots too fat to copy (tie groups) are not copied; their min/max key is tracked instead.
//   S4 k_qs_finish  one CTA per slot: radix select (12-bit digits below the common key prefix) of every target rank
//                   over the slot's candidates in shared memory; a fat slot whose min == max (a pure tie group)
//                   answers immediately, any other fat slot runs the same select straight over D (slow, correct).
//                   The last CTA evaluates the type-7 interpolation  q = (1-h) x[lo] + h x[hi]  in R's order.
// Versus a full 64-bit radix sort of the matrix: one read of a 2-byte side array instead of ~16 read+write sweeps of D.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "mdeggs_kernels.cuh"

namespace mdeggs {

constexpr int QS_MAX_T = 256;        // 2 * P
constexpr int QS_CELLS = 64;         // knot cells
constexpr int QS_SUB = 64;           // equal-width sub-bins per cell
constexpr int QS_BINS = QS_CELLS * QS_SUB;
constexpr int QS_SAMPLE = 1024;      // sample size (a single-CTA sort: kept small; it only shapes the bins)
constexpr int QS_SAMPLE_MAX = QS_SAMPLE;
constexpr int QS_SAMPLE_THREADS = QS_SAMPLE / 2;  // one compare-exchange pair per thread in every sorting step
static_assert(QS_SAMPLE_THREADS <= 1024 && QS_SAMPLE_THREADS % 32 == 0, "single CTA");
constexpr int QS_FIN_CAP = 16384;    // keys a finish CTA keeps in shared memory (128 KB); longer lists stay in L2
constexpr unsigned int QS_SLOT_MAX = 1u << 20;  // slots with more values are not compacted ("fat": tie groups)
constexpr unsigned int QS_CAND_MAX = 8u << 20;  // candidate buffer, keys
constexpr int QS_FIN_THREADS = 512;
constexpr int QS_DIRECT = 256;        // keys ranked directly once a digit pass leaves this few
constexpr int QS_SCAN_THREADS = 256;  // >= QS_MAX_T
constexpr int QS_STAGE = 2048;       // per-CTA staging entries of the compaction pass

constexpr int QS_NO_BIN = 0xffff;    // bin id of padding slots and NaN in the per-value bin array
constexpr int QS_TAB = 4096;         // uniform value cells of the knot look-up table
struct QsKnots {
  double base;                 // sample minimum; bins are evaluated on xf = (float)(x - base)
  float knot[QS_CELLS + 1];
  float scale[QS_CELLS];
  float inv_w;                 // QS_TAB / (sample range): uniform cell index = (int)(xf * inv_w)
  // per uniform cell: low 6 bits = first knot cell that can hold its values, top 2 bits = 0: exactly that cell,
  // 1: that cell or the next, 2: anything (binary search)
  unsigned char tab[QS_TAB];
};
static_assert(sizeof(QsKnots) % 8 == 0, "QsKnots is copied as 8-byte words");

struct QsState {
  long long N;                         // non-NaN values
  int U;                               // slots (distinct target bins)
  unsigned int done;                   // finish CTAs completed
  unsigned int gather_done;            // gather CTAs that flushed their histogram (zero between calls)
  int t_slot[QS_MAX_T];                // per target
  long long t_rank[QS_MAX_T];          // per target: 0-based rank inside its bin
  unsigned long long t_key[QS_MAX_T];  // per target: selected key
  int slot_bin[QS_MAX_T];
  unsigned int slot_count[QS_MAX_T];   // values in the slot's bin
  unsigned int slot_off[QS_MAX_T];     // offset of the slot's list in the candidate buffer
  unsigned int slot_cursor[QS_MAX_T];  // append cursor
  int slot_fat[QS_MAX_T];              // 1: not compacted (too many values, or the buffer would overflow)
  unsigned long long slot_min[QS_MAX_T], slot_max[QS_MAX_T];  // fat slots: key range
  QsKnots knots;
};

// monotone map value -> bin.  `kn` points at a shared-memory copy of the knots.  NaN never reaches here.
__device__ __forceinline__ int qs_cell_search(const QsKnots* __restrict__ kn, float xf) {
  int c = 0;
#pragma unroll
  for (int step = QS_CELLS / 2; step > 0; step >>= 1)
    if (xf >= kn->knot[c + step]) c += step;  // last c in [0, 63] with knot[c] <= xf (c = 0 below knot[0])
  return c;
}
__device__ __forceinline__ int qs_tab_index(float xf, float inv_w, int cells) {
  int idx = (int)(xf * inv_w);  // saturating conversion; NaN (inf * 0) -> 0
  return idx < 0 ? 0 : (idx > cells - 1 ? cells - 1 : idx);
}
__device__ __forceinline__ int qs_bin(const QsKnots* __restrict__ kn, double x) {
  const float xf = (float)(x - kn->base);  // monotone; +-inf and overflow stay ordered
  const unsigned e = kn->tab[qs_tab_index(xf, kn->inv_w, QS_TAB)];
  int c = (int)(e & 63u);
  // (an entry without the 0x40 flag has every value of its uniform cell below knot[c + 1], so the comparison is simply
  // made for all of them: no second test; the shared-memory copy carries knot[QS_CELLS] = NaN, see qs_load_knots)
  if (e & 0x80u) c = qs_cell_search(kn, xf);
  else c += (xf >= kn->knot[c + 1]);
  int sub = (int)((xf - kn->knot[c]) * kn->scale[c]);  // NaN products (inf * 0) convert to 0
  sub = sub < 0 ? 0 : (sub > QS_SUB - 1 ? QS_SUB - 1 : sub);
  return c * QS_SUB + sub;
}

__device__ __forceinline__ void qs_load_knots(QsKnots* __restrict__ dst, const QsState* __restrict__ st) {
  // (knot[QS_CELLS], the sample maximum, is not needed by qs_bin: the copy holds NaN there so that the unconditional
  // "next cell?" comparison of qs_bin can never leave the last cell; written by the thread that copies its word)
  constexpr int last_word = (int)(offsetof(QsKnots, knot) / 8) + QS_CELLS / 2;  // knot[QS_CELLS] is the low float of it
  static_assert(offsetof(QsKnots, knot) % 8 == 0 && QS_CELLS % 2 == 0, "knot[QS_CELLS] sits in the low half of a word");
  for (int i = threadIdx.x; i < (int)(sizeof(QsKnots) / 8); i += blockDim.x) {
    double w = reinterpret_cast<const double*>(&st->knots)[i];
    if (i == last_word) {
      unsigned long long b = (unsigned long long)__double_as_longlong(w);
      b = (b & 0xffffffff00000000ull) | 0x7fc00000ull;  // low float := NaN (">=" is false for every value, +inf included)
      w = __longlong_as_double((long long)b);
    }
    reinterpret_cast<double*>(dst)[i] = w;
  }
}
__device__ __forceinline__ void qs_flush_hist(const unsigned int* __restrict__ s_hist, unsigned int* __restrict__ hist) {
  for (int i = threadIdx.x; i < QS_BINS; i += blockDim.x) {
    const unsigned int c = s_hist[i];
    if (c) atomicAdd(&hist[i], c);
  }
}

// position of knot c on the sample-rank axis, in 1/1024 of (m - 1): 0 1 2 4 8 | 26 44 ... 998 | 1016 1020 1022 1023 1024
__device__ __forceinline__ int qs_knot_pos(int c) {
  if (c < 5) return c == 0 ? 0 : 1 << (c - 1);
  if (c > 59) return c == 64 ? 1024 : 1024 - (1 << (63 - c));
  return 8 + (c - 4) * 18;
}

// ---- S0 ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(QS_SAMPLE_THREADS) k_qs_sample(const double* __restrict__ assay, int64_t ld,
                                                                 int64_t row_off, int colmajor,
                                                                 const int32_t* __restrict__ from,
                                                                 const int32_t* __restrict__ to, int64_t E, int G,
                                                                 int n, int S, QsState* __restrict__ st,
                                                                 const unsigned int* __restrict__ gate,
                                                                 int64_t row_limit, int32_t* __restrict__ err) {
  // Rows are drawn through the edge list (end points of edges ARE the node rows), so the kernel needs nothing from
  // K0 and runs beside it.  Hub genes are drawn more often: harmless, the sample only shapes the bins.
  // Streamed input copy (`gate` != null): only the first chunk of rows, [0, row_limit) of the device copy, is waited
  // for and sampled; draws that land elsewhere are dropped and the drawing is repeated (other edges of the same strata)
  // until the sample is three quarters full.  A sample that is not representative only unbalances the bins.
  __shared__ unsigned long long sk[QS_SAMPLE_MAX];
  __shared__ int s_m;
  const long long ends = 2 * E;
  const long long total = ends * n;
  if (threadIdx.x == 0) {
    s_m = 0;
    if (gate) gate_wait(gate, 0, err);
  }
  for (int i = threadIdx.x; i < S; i += blockDim.x) sk[i] = ~0ull;
  __syncthreads();
  auto row_of_end = [&](long long j) -> int64_t {  // j in [0, 2E): from[0..E) then to[0..E); -1 when out of range
    const int r = (j < E ? from[j] : to[j - E]) - 1;
    if (r < 0 || r >= G) return -1;
    const int64_t rel = (int64_t)r - row_off;
    return (gate && (rel < 0 || rel >= row_limit)) ? -1 : rel;
  };
  if (total > S) {
    const unsigned jit = (unsigned)(ends / S + 1);
    const int rounds = gate ? 12 : 1;
    for (int rd = 0; rd < rounds; ++rd) {
      if (rd > 0) {  // (uniform decision: nobody adds to s_m between the two barriers)
        __syncthreads();
        const int have = s_m;
        __syncthreads();
        if (have >= S - S / 4) break;
      }
      for (int i0 = 0; i0 < S; i0 += 2 * QS_SAMPLE_THREADS) {  // 2 independent draws per thread in flight
        int64_t off[2];
        double v[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int i = i0 + u * QS_SAMPLE_THREADS + threadIdx.x;
          const unsigned h = (unsigned)(i + rd * S) * 2654435761u;
          long long j = ((long long)i * ends) / S + (long long)((h >> 20) % jit);
          j = j < ends ? j : ends - 1;
          const int smp = (int)((h >> 4) % (unsigned)n);
          const int64_t row = i < S ? row_of_end(j) : -1;
          off[u] = row < 0 ? -1 : (colmajor ? row + (int64_t)smp * ld : row * ld + smp);
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) v[u] = off[u] >= 0 ? __ldg(assay + off[u]) : nan("");
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const unsigned long long k = dkey(v[u]);
          if (k != ~0ull) {
            const int at = atomicAdd(&s_m, 1);
            if (at < S) sk[at] = k;
          }
        }
      }
    }
  } else {
    for (int li = threadIdx.x; li < (int)total; li += blockDim.x) {
      const int64_t row = row_of_end(li / n);
      const int smp = li % n;
      const unsigned long long k =
          row < 0 ? ~0ull : dkey(colmajor ? assay[row + (int64_t)smp * ld] : assay[row * ld + smp]);
      if (k != ~0ull) sk[atomicAdd(&s_m, 1)] = k;
    }
  }
  __syncthreads();
  const int m = s_m < S ? s_m : S;
  // bitonic sort of the S slots (unused slots hold the maximum key and stay at the end), one pair per thread and
  // step.  For j <= 32 the 32 pairs of a warp live in its own 64 consecutive slots, so those steps (45 of the 55)
  // only need a warp barrier; a CTA barrier is needed when the step that follows reaches across warps.
  for (int k = 2; k <= S; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < S / 2; t += QS_SAMPLE_THREADS) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const unsigned long long a = sk[i], b = sk[i | j];
        if ((a > b) == ((i & k) == 0)) {
          sk[i] = b;
          sk[i | j] = a;
        }
      }
      const int next_j = j > 1 ? (j >> 1) : k;  // first step of the next merge: j' = (2k) / 2
      if (S / 2 <= QS_SAMPLE_THREADS && j <= 32 && next_j <= 32) __syncwarp();
      else __syncthreads();
    }
  }
  const double base = m > 0 ? dkey_inv(sk[0]) : 0.0;
  auto knot_at = [&](int c) -> float {
    if (m <= 0) return 0.f;
    const int r = (int)(((long long)qs_knot_pos(c) * (m - 1) + 512) >> 10);
    return (float)(dkey_inv(sk[r]) - base);
  };
  __shared__ float s_knot[QS_CELLS + 1];
  if (threadIdx.x <= QS_CELLS) {
    const float v = knot_at(threadIdx.x);
    s_knot[threadIdx.x] = v;
    st->knots.knot[threadIdx.x] = v;
  }
  __syncthreads();
  if (threadIdx.x < QS_CELLS) {
    const float lo = s_knot[threadIdx.x], hi = s_knot[threadIdx.x + 1];
    float sc = (float)QS_SUB / (hi - lo);
    if (!(hi > lo) || !isfinite(sc)) sc = 0.f;
    st->knots.scale[threadIdx.x] = sc;
  }
  // look-up table over QS_TAB uniform cells of [0, sample range]: knot cells that the values of uniform cell i can
  // fall into, padded by one uniform cell on both sides against rounding of the index computation
  const float range = s_knot[QS_CELLS];
  float inv_w = range > 0.f ? (float)QS_TAB / range : 0.f;
  if (!isfinite(inv_w)) inv_w = 0.f;
  const float w = inv_w > 0.f ? 1.f / inv_w : 0.f;
  if (threadIdx.x == 0) {
    st->knots.base = base;
    st->knots.inv_w = inv_w;
  }
  auto cell_of = [&](float xf) {
    int c = 0;
#pragma unroll
    for (int step = QS_CELLS / 2; step > 0; step >>= 1)
      if (xf >= s_knot[c + step]) c += step;
    return c;
  };
  for (int i = threadIdx.x; i < QS_TAB; i += blockDim.x) {
    const int c_lo = inv_w > 0.f ? cell_of(((float)i - 1.f) * w) : 0;
    const int c_hi = inv_w > 0.f && i < QS_TAB - 1 ? cell_of(((float)i + 2.f) * w) : QS_CELLS - 1;
    const int d = c_hi - c_lo;
    st->knots.tab[i] = (unsigned char)(c_lo | (d == 0 ? 0 : (d == 1 ? 0x40 : 0x80)));
  }
}

__device__ __forceinline__ void qs_flush_and_scan(unsigned int* __restrict__ s_hist, unsigned int* __restrict__ hist,
                                                  QsState* __restrict__ st, const double* __restrict__ pct, int P,
                                                  unsigned int cand_capacity);

// ------------------------------------------------------------------------------------------------
// K1: gather + transposition  (replaces the per-pair `assayData[gene, ]` row look-ups, core_functions.R:880-881)
//   in : R column-major G x n (gene fastest)  or row-major (sample fastest)
//   out: D[node][pos] with category-contiguous samples
// Persistent CTAs (a multiple of the SM count) walk the tiles; every value written to D is also counted in the
// CTA's shared-memory copy of the quantile histogram (S1), flushed once per CTA.  Column-major input goes through a
// 32x32 shared-memory transposition tile, coalesced on both sides.  Padding positions (src < 0) are written by K2.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 6) k_gather_colmajor(const double* __restrict__ assay, int64_t ld,
                                                         int64_t row_off, const int32_t* __restrict__ rowlist,
                                                         const int32_t* __restrict__ n_nodes_p,
                                                         const int32_t* __restrict__ src_of_pos, int npad, int n,
                                                         double* __restrict__ D, QsState* __restrict__ st,
                                                         unsigned int* __restrict__ hist,
                                                         unsigned short* __restrict__ B16,
                                                         const double* __restrict__ pct, int P,
                                                         unsigned int cand_capacity) {
  __shared__ double tile[32][33];
  __shared__ unsigned short tile16[32][34];
  __shared__ QsKnots kn;
  __shared__ __align__(16) unsigned int s_hist[QS_BINS];
  const bool do_hist = hist != nullptr;
  if (do_hist) {
    qs_load_knots(&kn, st);
    for (int i = threadIdx.x; i < QS_BINS; i += blockDim.x) s_hist[i] = 0u;
  }
  __syncthreads();
  const int n_nodes = *n_nodes_p;
  const int n_tx = (npad + 31) / 32, n_ty = (n_nodes + 31) / 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const long long n_tiles = (long long)n_tx * n_ty;
  const bool small = n_tiles < (1ll << 31);  // 32-bit tile arithmetic (a 64-bit division costs ~15 % of this kernel)
  for (long long tile_id = blockIdx.x; tile_id < n_tiles; tile_id += gridDim.x) {
    int node0, pos0;
    if (small) {
      const unsigned t = (unsigned)tile_id, q = t / (unsigned)n_tx;
      node0 = (int)q * 32;
      pos0 = (int)(t - q * (unsigned)n_tx) * 32;
    } else {
      node0 = (int)(tile_id / n_tx) * 32;
      pos0 = (int)(tile_id % n_tx) * 32;
    }
    const int node_r = node0 + tx;
    const int64_t row = node_r < n_nodes ? rowlist[node_r] - row_off : -1;  // device copy starts at host row row_off
    double v[4];
    bool ok[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {  // the 4 loads of this thread are issued before any of them is binned
      const int pos = pos0 + ty + 8 * q;
      const int s = pos < npad ? src_of_pos[pos] : -1;
      ok[q] = s >= 0 && row >= 0;
      v[q] = ok[q] ? __ldg(assay + row + (int64_t)s * ld) : 0.0;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) tile[ty + 8 * q][tx] = v[q];
    if (do_hist) {  // bin of every value: counted here, and kept (2 bytes per value) for the compaction pass
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int b = (ok[q] && !isnan(v[q])) ? qs_bin(&kn, v[q]) : QS_NO_BIN;
        tile16[ty + 8 * q][tx] = (unsigned short)b;
        if (b != QS_NO_BIN) atomicAdd(&s_hist[b], 1u);
      }
    }
    __syncthreads();
    const int pos_w = pos0 + tx;
    const int s_w = pos_w < npad ? src_of_pos[pos_w] : -1;
#pragma unroll
    for (int i = ty; i < 32; i += 8) {
      const int node = node0 + i;
      if (node < n_nodes && pos_w < npad) {
        if (s_w >= 0) D[(size_t)node * npad + pos_w] = tile[tx][i];
        if (do_hist) B16[(size_t)node * npad + pos_w] = s_w >= 0 ? tile16[tx][i] : (unsigned short)QS_NO_BIN;
      }
    }
    __syncthreads();
  }
  if (do_hist) qs_flush_and_scan(s_hist, hist, st, pct, P, cand_capacity);  // the last CTA also runs S2
}

__global__ void __launch_bounds__(256, 6) k_gather_rowmajor(const double* __restrict__ assay, int64_t ld,
                                                         int64_t row_off, const int32_t* __restrict__ rowlist,
                                                         const int32_t* __restrict__ n_nodes_p,
                                                         const int32_t* __restrict__ src_of_pos, int npad, int n,
                                                         double* __restrict__ D, QsState* __restrict__ st,
                                                         unsigned int* __restrict__ hist,
                                                         unsigned short* __restrict__ B16,
                                                         const double* __restrict__ pct, int P,
                                                         unsigned int cand_capacity) {
  __shared__ QsKnots kn;
  __shared__ __align__(16) unsigned int s_hist[QS_BINS];
  const bool do_hist = hist != nullptr;
  if (do_hist) {
    qs_load_knots(&kn, st);
    for (int i = threadIdx.x; i < QS_BINS; i += blockDim.x) s_hist[i] = 0u;
  }
  __syncthreads();
  const int n_nodes = *n_nodes_p;
  for (int node = blockIdx.x; node < n_nodes; node += gridDim.x) {
    const int64_t row = rowlist[node] - row_off;
    for (int pos = threadIdx.x; pos < npad; pos += blockDim.x) {
      const int s = src_of_pos[pos];
      int b = QS_NO_BIN;
      if (s >= 0) {
        const double v = __ldg(assay + row * ld + s);
        if (do_hist && !isnan(v)) {
          b = qs_bin(&kn, v);
          atomicAdd(&s_hist[b], 1u);
        }
        D[(size_t)node * npad + pos] = v;
      }
      if (do_hist) B16[(size_t)node * npad + pos] = (unsigned short)b;
    }
  }
  __syncthreads();
  if (do_hist) qs_flush_and_scan(s_hist, hist, st, pct, P, cand_capacity);  // the last CTA also runs S2
}

// ---- S2 ------------------------------------------------------------------------------------------
// Run by the LAST CTA of the gather kernel (256 threads) once every CTA has flushed its histogram: prefix sums, each
// target rank -> (bin, rank inside the bin), distinct target bins -> slots with exact counts and list offsets.  The
// bins' prefix lives in registers (16 bins per thread); a thread answers the targets that fall into its own range.
struct QsScanSmem {  // scratch of qs_scan_body; lives in the CTA's (flushed) shared-memory histogram
  unsigned long long warp[QS_SCAN_THREADS / 32];
  unsigned long long rank[QS_MAX_T];
  int tbin[QS_MAX_T], first[QS_MAX_T];
  unsigned int tcnt[QS_MAX_T], slot_cnt[QS_MAX_T];
};
static_assert(sizeof(QsScanSmem) <= (size_t)QS_BINS * 4, "the scan scratch aliases the histogram");

// Called by every thread of the CTA (blockDim.x >= QS_SCAN_THREADS, a multiple of 32); the first QS_SCAN_THREADS work.
__device__ __forceinline__ void qs_scan_body(QsState* __restrict__ st, unsigned int* __restrict__ hist,
                                             const double* __restrict__ pct, int P, unsigned int cand_capacity,
                                             QsScanSmem* __restrict__ sm) {
  unsigned long long* s_warp = sm->warp;
  unsigned long long* s_rank = sm->rank;
  int *s_tbin = sm->tbin, *s_first = sm->first;
  unsigned int *s_tcnt = sm->tcnt, *s_slot_cnt = sm->slot_cnt;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const bool act = tid < QS_SCAN_THREADS;
  constexpr int PER = QS_BINS / QS_SCAN_THREADS;
  unsigned long long v[PER], mine = 0;
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    v[j] = 0ull;
    if (act) {
      v[j] = __ldcg(hist + tid * PER + j);
      mine += v[j];
      hist[tid * PER + j] = 0u;  // clean slate for the next call
    }
  }
  unsigned long long incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (act && lane == 31) s_warp[wid] = incl;
  __syncthreads();
  unsigned long long woff = 0, total = 0;
  for (int w = 0; w < QS_SCAN_THREADS / 32; ++w) {
    woff += w < wid ? s_warp[w] : 0ull;
    total += s_warp[w];
  }
  const unsigned long long my_start = woff + incl - mine, my_end = my_start + mine;
  const long long N = (long long)total;  // == number of non-NaN node-row values
  const int T = 2 * P;
  if (tid == 0) {
    st->N = N;
    st->done = 0u;
  }
  if (tid < T) {
    const int p = tid >> 1;
    long long ilo = 1, ihi = 1;
    if (N > 0) {
      const double idx = __dadd_rn(1.0, __dmul_rn((double)(N - 1), pct[p]));
      ilo = (long long)floor(idx);
      ihi = (long long)ceil(idx);
      ilo = ilo < 1 ? 1 : (ilo > N ? N : ilo);
      ihi = ihi < 1 ? 1 : (ihi > N ? N : ihi);
    }
    s_rank[tid] = (unsigned long long)(((tid & 1) ? ihi : ilo) - 1);
    s_tbin[tid] = QS_BINS - 1;  // N == 0: no thread owns the rank; the last bin (empty) stands in
    s_tcnt[tid] = 0u;
    st->t_rank[tid] = 0;
    st->t_key[tid] = 0ull;
  }
  __syncthreads();
  if (mine > 0) {
    for (int t = 0; t < T; ++t) {
      const unsigned long long r = s_rank[t];
      if (r < my_start || r >= my_end) continue;
      unsigned long long run = my_start;
#pragma unroll
      for (int j = 0; j < PER; ++j) {
        if (r >= run && r < run + v[j]) {
          s_tbin[t] = tid * PER + j;
          s_tcnt[t] = (unsigned int)v[j];
          st->t_rank[t] = (long long)(r - run);
        }
        run += v[j];
      }
    }
  }
  __syncthreads();
  // distinct target bins -> slots, in ascending bin order (T <= 256, O(T^2) in parallel)
  if (tid < T) {
    int first = 1;
    for (int u = 0; u < tid; ++u) first &= (s_tbin[u] != s_tbin[tid]);
    s_first[tid] = first;
  }
  __syncthreads();
  if (tid < T) {
    int sl = 0;
    for (int u = 0; u < T; ++u) sl += (s_first[u] && s_tbin[u] < s_tbin[tid]);
    st->t_slot[tid] = sl;
    if (s_first[tid]) {
      st->slot_bin[sl] = s_tbin[tid];
      s_slot_cnt[sl] = s_tcnt[tid];
      st->slot_count[sl] = s_tcnt[tid];
    }
  }
  __threadfence_block();
  __syncthreads();
  if (tid == 0) {
    int U = 0;
    for (int u = 0; u < T; ++u) U += s_first[u];
    st->U = U;
    unsigned long long off = 0;
    for (int s = 0; s < U; ++s) {
      const unsigned int c = s_slot_cnt[s];
      const bool fat = c > QS_SLOT_MAX || off + c > (unsigned long long)cand_capacity;
      st->slot_fat[s] = fat ? 1 : 0;
      st->slot_off[s] = (unsigned int)off;
      st->slot_cursor[s] = 0u;
      st->slot_min[s] = ~0ull;
      st->slot_max[s] = 0ull;
      if (!fat) off += c;
    }
  }
}

// whole CTA, at the end of a gather kernel: flush this CTA's histogram; the last CTA to do so runs S2
__device__ __forceinline__ void qs_flush_and_scan(unsigned int* __restrict__ s_hist, unsigned int* __restrict__ hist,
                                                  QsState* __restrict__ st, const double* __restrict__ pct, int P,
                                                  unsigned int cand_capacity) {
  __shared__ int s_last;
  qs_flush_hist(s_hist, hist);
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    s_last = (atomicAdd(&st->gather_done, 1u) == gridDim.x - 1);
    if (s_last) st->gather_done = 0u;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  qs_scan_body(st, hist, pct, P, cand_capacity, reinterpret_cast<QsScanSmem*>(s_hist));  // (histogram flushed: reuse)
}

// ---- S4 ------------------------------------------------------------------------------------------
// Key of 0-based rank `rank` among the keys visited by for_each(fn), all inside [kmin, kmax]: MSD radix select with
// 12-bit digits starting below the common prefix of kmin / kmax.  Called by the whole CTA.
template <class ForEach>
__device__ __forceinline__ unsigned long long qs_digit_select(ForEach&& for_each, unsigned long long kmin,
                                                              unsigned long long kmax, long long rank,
                                                              unsigned int* hist /* [4096] shared */,
                                                              long long* s_scan /* [blockDim.x] shared */,
                                                              unsigned long long* s_pfx, long long* s_rank,
                                                              long long* s_direct /* [2] shared */,
                                                              unsigned long long* s_list /* [QS_DIRECT] shared */) {
  const int tid = threadIdx.x, nthr = blockDim.x;
  if (kmin == kmax) return kmin;
  const int top = 63 - __clzll((long long)(kmin ^ kmax));  // highest differing bit
  if (tid == 0) {
    *s_pfx = top == 63 ? 0ull : (kmin >> (top + 1));  // fixed high bits, right aligned
    *s_rank = rank;
  }
  __syncthreads();
  int remaining = top + 1;  // undecided low bits
  while (remaining > 0) {
    const int width = remaining < 12 ? remaining : 12;
    const int nbins = 1 << width;
    const int shift = remaining - width;
    for (int i = tid; i < 4096; i += nthr) hist[i] = 0;
    __syncthreads();
    const unsigned long long pfx = *s_pfx;
    const bool no_fixed = remaining == 64;
    for_each([&](unsigned long long k) {
      if (no_fixed || (k >> remaining) == pfx)
        atomicAdd(&hist[(unsigned int)((k >> shift) & (unsigned long long)(nbins - 1))], 1u);
    });
    __syncthreads();
    const int per = 4096 / nthr;
    long long loc = 0;
    for (int j = 0; j < per; ++j) loc += hist[tid * per + j];
    s_scan[tid] = loc;
    __syncthreads();
    if (tid < 32) {  // exclusive scan of the per-thread totals by one warp
      long long carry = 0;
      for (int i0 = 0; i0 < nthr; i0 += 32) {
        const long long v = s_scan[i0 + tid];
        long long incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const long long t = __shfl_up_sync(0xffffffffu, incl, o);
          if (tid >= o) incl += t;
        }
        s_scan[i0 + tid] = carry + incl - v;
        carry += __shfl_sync(0xffffffffu, incl, 31);
      }
    }
    __syncthreads();
    const long long r = *s_rank;
    long long run = s_scan[tid];
    int found = -1;
    long long found_base = 0;
    for (int j = 0; j < per; ++j) {
      const int b = tid * per + j;
      const long long h = hist[b];
      if (b < nbins && r >= run && r < run + h) {
        found = b;
        found_base = run;
      }
      run += h;
    }
    __syncthreads();
    if (found >= 0) {  // exactly one thread
      *s_pfx = (pfx << width) | (unsigned long long)found;
      *s_rank = r - found_base;
      s_direct[0] = (long long)hist[found];
    }
    __syncthreads();
    remaining -= width;
    // few keys left under the new prefix (the usual case after ONE pass: a slot holds a few thousand keys, the
    // histogram 4,096 bins): collect them and rank them directly instead of further digit passes
    if (remaining > 0 && s_direct[0] <= (long long)QS_DIRECT) {
      const unsigned long long npfx = *s_pfx;
      const long long rr = *s_rank;
      if (tid == 0) s_direct[1] = 0;
      __syncthreads();
      for_each([&](unsigned long long k) {
        if ((k >> remaining) == npfx) {
          const long long at = (long long)atomicAdd(reinterpret_cast<unsigned long long*>(&s_direct[1]), 1ull);
          if (at < QS_DIRECT) s_list[at] = k;
        }
      });
      __syncthreads();
      const int cnt = (int)s_direct[1];
      for (int t = tid; t < cnt; t += nthr) {
        const unsigned long long mine = s_list[t];
        int rk = 0;
        for (int u = 0; u < cnt; ++u) rk += (s_list[u] < mine) || (s_list[u] == mine && u < t);
        if (rk == (int)rr) *s_pfx = mine;
      }
      __syncthreads();
      remaining = 0;
    }
  }
  const unsigned long long res = *s_pfx;
  __syncthreads();
  return res;
}

__global__ void __launch_bounds__(QS_FIN_THREADS) k_qs_finish(const double* __restrict__ D,
                                                             const unsigned short* __restrict__ B16, int npad,
                                                             const int32_t* __restrict__ n_nodes_p,
                                                             QsState* __restrict__ st,
                                                             const unsigned long long* __restrict__ cand,
                                                             const double* __restrict__ pct, int P,
                                                             double* __restrict__ cutoff) {
  extern __shared__ unsigned long long fk[];  // QS_FIN_CAP keys
  __shared__ unsigned int hist[4096];
  __shared__ long long s_scan[QS_FIN_THREADS];
  __shared__ unsigned long long s_pfx, s_min, s_max;
  __shared__ long long s_rank;
  __shared__ long long s_direct[2];
  __shared__ unsigned long long s_list[QS_DIRECT];
  __shared__ int s_last;
  const int tid = threadIdx.x, s = blockIdx.x;
  const int T = 2 * P;
  // everything this CTA needs from the selection state in ONE round of loads (a loop of dependent global reads per
  // target cost more than the selection itself)
  __shared__ int sh_tslot[QS_MAX_T];
  __shared__ long long sh_trank[QS_MAX_T];
  for (int t = tid; t < T; t += QS_FIN_THREADS) {
    sh_tslot[t] = st->t_slot[t];
    sh_trank[t] = st->t_rank[t];
  }
  const int U = st->U;
  const unsigned int c = s < U ? st->slot_count[s] : 0u;
  const int fat = s < U ? st->slot_fat[s] : 0;
  const unsigned int list_off = s < U ? st->slot_off[s] : 0u;
  if (s >= U) return;
  __syncthreads();
  if (!fat) {
    const unsigned long long* list = cand + list_off;
    if (tid == 0) {
      s_min = ~0ull;
      s_max = 0ull;
    }
    __syncthreads();
    const bool in_smem = c <= (unsigned)QS_FIN_CAP;
    unsigned long long mn = ~0ull, mx = 0ull;
    for (unsigned int i = tid; i < c; i += QS_FIN_THREADS) {
      const unsigned long long k = list[i];
      if (in_smem) fk[i] = k;
      mn = k < mn ? k : mn;
      mx = k > mx ? k : mx;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long a = __shfl_xor_sync(0xffffffffu, mn, o), b = __shfl_xor_sync(0xffffffffu, mx, o);
      mn = a < mn ? a : mn;
      mx = b > mx ? b : mx;
    }
    if ((tid & 31) == 0 && mn <= mx) {
      atomicMin(&s_min, mn);
      atomicMax(&s_max, mx);
    }
    __syncthreads();
    const unsigned long long kmin = s_min, kmax = s_max;
    auto for_each_cand = [&](auto&& fn) {
      if (in_smem) {
        for (unsigned int i = tid; i < c; i += QS_FIN_THREADS) fn(fk[i]);
      } else {
        for (unsigned int i = tid; i < c; i += QS_FIN_THREADS) fn(__ldg(list + i));
      }
    };
    for (int t = 0; t < T; ++t) {
      if (sh_tslot[t] != s) continue;  // uniform across the CTA
      const unsigned long long key = c > 0 ? qs_digit_select(for_each_cand, kmin, kmax, sh_trank[t], hist, s_scan, &s_pfx, &s_rank, s_direct, s_list) : ~0ull;
      if (tid == 0) st->t_key[t] = key;
    }
  } else if (st->slot_min[s] == st->slot_max[s]) {
    // a pure tie group: every rank inside it is the same key
    for (int t = tid; t < T; t += QS_FIN_THREADS)
      if (sh_tslot[t] == s) st->t_key[t] = st->slot_min[s];
  } else {
    // slow path: the same select straight over D, filtered by the bin ids
    const unsigned my_bin = (unsigned)st->slot_bin[s];
    const long long total = (long long)(*n_nodes_p) * npad;
    auto for_each_in_bin = [&](auto&& fn) {
      for (long long i = tid; i < total; i += QS_FIN_THREADS)
        if (B16[i] == my_bin) fn(dkey(D[i]));
    };
    for (int t = 0; t < T; ++t) {
      if (sh_tslot[t] != s) continue;  // uniform across the CTA
      const unsigned long long key = qs_digit_select(for_each_in_bin, st->slot_min[s], st->slot_max[s], sh_trank[t], hist, s_scan, &s_pfx, &s_rank, s_direct, s_list);
      if (tid == 0) st->t_key[t] = key;
    }
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(&st->done, 1u) == (unsigned)(U - 1));
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const long long N = st->N;
  volatile unsigned long long* tk = st->t_key;
  for (int p = tid; p < P; p += QS_FIN_THREADS) {
    if (N <= 0) {
      cutoff[p] = nan("");
      continue;
    }
    const double idx = __dadd_rn(1.0, __dmul_rn((double)(N - 1), pct[p]));
    const double lo = floor(idx);
    double q = dkey_inv(tk[2 * p]);
    const double xh = dkey_inv(tk[2 * p + 1]);
    if (idx > lo && xh != q) {
      const double hfrac = __dsub_rn(idx, lo);
      q = __dadd_rn(__dmul_rn(__dsub_rn(1.0, hfrac), q), __dmul_rn(hfrac, xh));
    }
    cutoff[p] = q;
  }
}

}  // namespace mdeggs

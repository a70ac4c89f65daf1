// mdeggs_front.cuh — K1 + K2 as ONE kernel: the node rows of the omic matrix are pulled into shared memory by TMA
// (cp.async.bulk.tensor 2-D boxes, hardware swizzle, mbarrier pipeline) and everything that used to take two kernels
// and a round trip of D through L2/HBM happens on the staged tile:
//   gather + transposition into the gene-major, category-contiguous matrix D   (core_functions.R:336, 880-881)
//   per (gene, category) mean, centred Sxx and the exact median                (core_functions.R:353-365)
//   the 4,096-bin value histogram of the quantile selection + the 2-byte bin id of every value (core_functions.R:718)
// D (and the bin ids) are written once; nothing is re-read.  The quantile candidates are collected afterwards by
// k_qs_compact from the bin ids (a quarter of D), because the slot bins are only known once the histogram is complete.
//
// Tiling.  A tile is GB consecutive ROWS of the input matrix (GB = 16 / 8 / 4: a 128 / 64 / 32-byte TMA inner box) by
// all n samples, `nbox` boxes of `box1` samples each, landing as stage[sample][gene] (gene fastest) with the
// SWIZZLE_128B / 64B / 32B pattern.  Which of the GB rows are network nodes comes from K0's row bitmap; tiles without a
// node are skipped, so scattered node rows cost DRAM sectors that would be touched anyway and never a different code
// path.  One persistent CTA per SM walks the tiles through a ring of `stages` buffers: 4 producer warps (thread 0
// issues the TMA; all of them run the plain cp.async loader when the matrix cannot be described by a tensor map: odd
// leading dimension, unaligned base, row-major) and NW consumer warps, one (gene, category) unit at a time, the
// segment held in registers as doubles.  A lane reads stage[src_of_pos[pos]][g]: the swizzle spreads the rows of one
// gene over the 16-byte bank groups, so the transposition costs a ~3-way bank conflict instead of a second buffer.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "mdeggs_kernels.cuh"
#include "mdeggs_qsel.cuh"

namespace mdeggs {

constexpr int FR_PROD_WARPS = 4;
constexpr int FR_MAX_STAGES = 4;
constexpr int FR_SMEM_MAX = 227 * 1024;

struct FrontArgs {
  const double* assay;   // device copy of the matrix; its first row is absolute row `row_off`
  int64_t ld, row_off, rows_avail;
  int colmajor, use_tma;
  const unsigned int* bitmap;  // K0: bit r set <=> absolute row r is a network node
  const int32_t* word_base;    // K0: node id of the first set bit of every 32-row word
  const int32_t* src_of_pos;   // category-contiguous position -> sample column (-1: padding)
  SegLayout L;
  double* D;
  unsigned short* B16;         // null: no quantile selection in this run
  double *mean, *sxx, *med;
  uint8_t* bad;
  QsState* st;
  unsigned int* hist;
  const double* pct;
  int P;
  unsigned int cand_capacity;
  int tile_lo, tile_hi;        // tiles (blocks of GB absolute rows) to visit
  int nbox, box1, stages;
  uint32_t stage_bytes;
  // streamed input copy: rows of chunk c = [row_off + c gate_rows, row_off + (c + 1) gate_rows) are on the device once
  // gate[c] != 0 (null: the whole matrix is there); only the TMA loader honours it (it reads through L2, never L1)
  const unsigned int* gate;
  int gate_rows, gate_chunks;
  int32_t* err;
};

// ---- PTX helpers (mbarrier pipeline, TMA) --------------------------------------------------------
__device__ __forceinline__ uint32_t fr_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void fr_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fr_mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void fr_mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void fr_mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void fr_tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void fr_cp_async8(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void fr_cp_async_arrive(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}

// byte offset of element (sample s, gene g) inside a stage whose rows are GB doubles wide, TMA swizzle applied:
// the 16-byte chunk index (address bits 4..6) is XORed with address bits 7..9, masked to the swizzle span
template <int GB>
__device__ __forceinline__ uint32_t fr_swz(int s) {
  constexpr uint32_t ROWB = GB * 8;
  constexpr uint32_t MASK = ROWB / 16 - 1;  // 7 / 3 / 1
  return (((uint32_t)s * ROWB) >> 7) & MASK;
}

// NaN-ignoring running minimum / maximum (a NaN operand compares false and is never taken; `acc` starts at +-inf).
// fmin / fmax on FP64 cost seven instructions each on sm_100a, a compare and a 64-bit select three.
__device__ __forceinline__ double fr_min(double acc, double x) { return x < acc ? x : acc; }
__device__ __forceinline__ double fr_max(double acc, double x) { return x > acc ? x : acc; }
__device__ __forceinline__ double fr_lds(uint32_t addr) {
  double x;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(x) : "r"(addr));
  return x;
}

// Exact k-th smallest (0-based) of the non-NaN values a warp holds in registers (ITEMS per lane, unused slots NaN),
// and, when `want_next`, the (k+1)-th as well (the two middle values of an even-sized segment).
// Bisection in KEY space, the values stay doubles: a counting pass is one FP64 compare per item (NaN compares false =
// sorts above everything), only the pivot travels between key and value space.  The bracket starts from the caller's
// guess [g_lo, g_hi); the minimum and maximum of the segment are only computed when the guess misses (or there is
// none): they bound the key range the halvings have to cross.  Once <= 32 values remain in [lo, hi) each goes to one
// lane (through shared memory, order irrelevant) and the warp quick-selects among its lanes with ballots: a pivot is
// the value of the lowest lane still in play, ~2 ln(32) rounds of a dozen instructions instead of 32 x 32 comparisons.
template <int ITEMS>
__device__ __forceinline__ double warp_select_dbl(const double (&v)[ITEMS], int kth, int n_valid, double* scratch32,
                                                  unsigned int* scratch_n, int lane, bool have_guess, double g_lo,
                                                  double g_hi, bool want_next, double* next_out) {
  unsigned long long lo = 0ull, hi = ~0ull;  // every non-NaN key lies in [lo, hi): dkey(NaN) == ~0
  double lo_d = -INFINITY, hi_d = 0.0;
  bool hi_open = true, lo_open = true;  // no value is below lo / at or above hi yet: no test needed on that side
  int cnt_lo = 0, cnt_hi = n_valid;
  auto narrow = [&](unsigned long long pk) {
    if (pk <= lo || pk >= hi) return;
    const double pd = dkey_inv(pk);
    int c = 0;
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) c += (v[j] < pd);
    c = __reduce_add_sync(0xffffffffu, c);
    if (c <= kth) {
      lo = pk;
      lo_d = pd;
      lo_open = false;
      cnt_lo = c;
    } else {
      hi = pk;
      hi_d = pd;
      hi_open = false;
      cnt_hi = c;
    }
  };
  if (have_guess) {
    narrow(dkey(g_lo));
    narrow(dkey(g_hi));
  }
  if (cnt_hi - cnt_lo > 32 && (lo_open || hi_open)) {
    // the guess missed on one side (or there was none): bound that side by the segment's extreme value
    double mn = INFINITY, mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
      mn = fr_min(mn, v[j]);
      mx = fr_max(mx, v[j]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn = fr_min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
      mx = fr_max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if (lo_open) {  // every value is >= mn: count below is 0
      lo = dkey(mn);
      lo_d = mn;
      lo_open = false;
    }
    if (hi_open) hi = dkey(mx) + 1ull;  // (stays "open": every value is <= mx, no upper test needed)
  }
  while (cnt_hi - cnt_lo > 32 && hi - lo > 1ull) narrow(lo + ((hi - lo) >> 1));
  double res, nxt;
  bool have_next;
  if (cnt_hi - cnt_lo > 32) {  // hi == lo + 1: more than 32 values equal to lo
    res = lo_d;
    nxt = lo_d;
    have_next = kth + 1 < cnt_hi;
  } else {
    if (lane == 0) *scratch_n = 0u;
    __syncwarp();
    if (!lo_open && !hi_open) {  // the usual case: both sides of the bracket are values (two compares per item)
#pragma unroll
      for (int j = 0; j < ITEMS; ++j) {
        const double x = v[j];
        if (x >= lo_d && x < hi_d) scratch32[atomicAdd(scratch_n, 1u)] = x;
      }
    } else {
#pragma unroll
      for (int j = 0; j < ITEMS; ++j) {
        const double x = v[j];
        if ((lo_open ? x == x : x >= lo_d) && (hi_open || x < hi_d)) scratch32[atomicAdd(scratch_n, 1u)] = x;
      }
    }
    __syncwarp();
    const int ncand = cnt_hi - cnt_lo;
    const double mine = lane < ncand ? scratch32[lane] : 0.0;
    // quick-select among the lanes: `act` = lanes still in play, `want` = rank wanted among them
    unsigned act = ncand >= 32 ? 0xffffffffu : ((1u << ncand) - 1u);
    int want = kth - cnt_lo;
    res = 0.0;
    unsigned above = 0u;  // lanes known to hold values > res when the loop ends
    for (;;) {
      const double pv = __shfl_sync(0xffffffffu, mine, __ffs(act) - 1);
      const bool in = (act >> lane) & 1u;
      const unsigned lt = __ballot_sync(0xffffffffu, in && mine < pv);
      const unsigned eq = __ballot_sync(0xffffffffu, in && mine == pv);
      const int nlt = __popc(lt), neq = __popc(eq);
      if (want < nlt) {
        above |= act & ~lt;
        act = lt;
      } else if (want < nlt + neq) {
        res = pv;
        above |= act & ~(lt | eq);
        have_next = want + 1 < nlt + neq;  // another copy of the same value follows
        break;
      } else {
        want -= nlt + neq;
        act &= ~(lt | eq);
      }
    }
    nxt = res;
    if (want_next && !have_next) {  // smallest candidate above res (if any)
      double mn = ((above >> lane) & 1u) ? mine : INFINITY;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mn = fr_min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
      have_next = above != 0u;
      nxt = mn;
    }
    __syncwarp();
  }
  if (want_next) {
    if (!have_next) {  // the k-th value is the largest below hi: the next one is the smallest value >= hi (rare)
      double mn = INFINITY;
#pragma unroll
      for (int j = 0; j < ITEMS; ++j)
        if (v[j] >= hi_d) mn = fr_min(mn, v[j]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mn = fr_min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
      nxt = mn;
    }
    *next_out = nxt;
  }
  return res;
}

// one (gene, category) unit of a staged tile, whole warp.  `tab` = this category's table of stage offsets (per
// position, gene 0, swizzle included; positions past the segment point at the stage's NaN row), `sbase` = shared
// address of the stage.
template <int ITEMS>
__device__ __forceinline__ void front_unit(const FrontArgs& A, uint32_t sbase, const uint32_t* __restrict__ tab,
                                           const QsKnots* __restrict__ kn, unsigned int* __restrict__ s_hist,
                                           double* __restrict__ scratch32, unsigned int* __restrict__ scratch_n, int g,
                                           int node, int k, int lane) {
  const SegLayout& L = A.L;
  const int m = L.cnt[k], off = L.off[k];
  const uint32_t g8 = (uint32_t)g * 8u;
  double v[ITEMS];
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) v[j] = fr_lds(sbase + (tab[lane + 32 * j] ^ g8));  // NaN past the segment
  double* __restrict__ drow = A.D + (size_t)node * L.npad + off;
  // `same`: every value of the segment equals the first one (a gene that is constant inside the category: exact zero
  // spread, an aliased column in R).  One equality test per value; the minimum and maximum are not needed here.
  const double v00 = __shfl_sync(0xffffffffu, v[0], 0);
  double s = 0.0;
  bool same = m > 0;
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    const double x = v[j];
    if (lane + 32 * j < m) {
      drow[lane + 32 * j] = x;
      s += x;
      same = same && (x == v00);
    }
  }
  s = warp_sum(s);
  same = __all_sync(0xffffffffu, same);
  int nn = m, nf = 0;
  if (!isfinite(s)) {  // a NaN / Inf in the segment (or an overflowing sum): count exactly
    nn = 0;
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
      if (lane + 32 * j < m) {
        nn += !isnan(v[j]);
        nf |= nonfinite_bits(v[j]);
      }
    }
    nn = __reduce_add_sync(0xffffffffu, nn);
    nf = __reduce_or_sync(0xffffffffu, nf);
  }
  if (A.B16) {  // K3 S1: bin of every value, counted in the CTA's histogram and kept (2 bytes) for the compaction
    unsigned short* __restrict__ brow = A.B16 + (size_t)node * L.npad + off;
    const bool any_nan = nn != m;
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
      if (lane + 32 * j < m) {
        const int b = (any_nan && isnan(v[j])) ? QS_NO_BIN : qs_bin(kn, v[j]);
        brow[lane + 32 * j] = (unsigned short)b;
        if (b != QS_NO_BIN) atomicAdd(&s_hist[b], 1u);
      }
    }
    if (m + lane < L.off[k + 1] - off) brow[m + lane] = (unsigned short)QS_NO_BIN;
  }
  double mu = m > 0 ? s / (double)m : nan("");
  double d1 = 0.0, d2 = 0.0;
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    if (lane + 32 * j < m) {
      const double d = v[j] - mu;
      d1 += d;
      d2 = fma(d, d, d2);
    }
  }
  d1 = warp_sum(d1);
  d2 = warp_sum(d2);
  if (m > 0) {
    const double c = d1 / (double)m;  // first-order correction of the mean
    mu += c;
    d2 -= d1 * c;
  }
  if (same) {  // constant inside the category: exact zero spread (aliased column in R)
    mu = v00;
    d2 = 0.0;
  }
  double median = nan("");
  if (same) {
    median = v00;
  } else if (nn > 0) {
    // guess bracket for the median: mean -+ c sd with c ~ 40 / m (about 32 values of a bell-shaped segment), at least
    // 2.5 standard errors of the median; a miss only costs further halvings
    bool guess = false;
    double g_lo = 0.0, g_hi = 0.0;
    if (m > 2 && d2 > 0.0 && isfinite(d2)) {
      const double c = fmax(36.0 / (double)m, 3.2 * (double)rsqrtf((float)m));
      const double w = fmin(c, 0.45) * sqrt(d2 / (double)(m - 1));
      g_lo = mu - w;
      g_hi = mu + w;
      guess = g_hi > g_lo;
    }
    double k2;
    const double k1 = warp_select_dbl<ITEMS>(v, (nn - 1) / 2, nn, scratch32, scratch_n, lane, guess, g_lo, g_hi,
                                             !(nn & 1), &k2);
    median = (nn & 1) ? k1 : __ddiv_rn(__dadd_rn(k1, k2), 2.0);
  }
  if (lane == 0) {
    A.mean[(size_t)node * L.K + k] = mu;
    A.sxx[(size_t)node * L.K + k] = d2;
    A.med[(size_t)node * L.K + k] = median;
    if (nf) mark_bad(A.bad, node, nf);
  }
  if (m + lane < L.off[k + 1] - off) drow[m + lane] = mu;  // padding := segment mean
}

template <int GB, int ITEMS, int NW>
__global__ void __launch_bounds__((NW + FR_PROD_WARPS) * 32, 1)
    k_front_fused(const __grid_constant__ CUtensorMap tmap, const FrontArgs A) {
  extern __shared__ unsigned char fr_smem_raw[];
  // the swizzle pattern is a function of the shared-memory ADDRESS: stages start on a 1,024-byte boundary
  unsigned char* fr_smem = fr_smem_raw + ((1024u - (fr_smem_u32(fr_smem_raw) & 1023u)) & 1023u);
  constexpr uint32_t ROWB = GB * 8;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int S = A.stages, K = A.L.K;
  const uint32_t stride = A.stage_bytes + 1024;  // a stage, then its NaN row (what positions past a segment read)
  // carve: [stages][stride] | histogram | knots | candidates[NW][32] | per-warp counters | barriers | ticket | tables
  unsigned char* p = fr_smem + (size_t)S * stride;
  unsigned int* s_hist = reinterpret_cast<unsigned int*>(p);
  p += QS_BINS * 4;
  QsKnots* kn = reinterpret_cast<QsKnots*>(p);
  p += sizeof(QsKnots);
  double* s_cand = reinterpret_cast<double*>(p);
  p += NW * 32 * 8;
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(p);
  p += 2 * FR_MAX_STAGES * 8;
  unsigned int* s_ncand = reinterpret_cast<unsigned int*>(p);
  p += NW * 4;
  unsigned int* s_ticket = reinterpret_cast<unsigned int*>(p);
  p += 16;
  uint32_t* tabK = reinterpret_cast<uint32_t*>(p);  // [K][ITEMS * 32]
  const bool do_hist = A.B16 != nullptr;
  if (do_hist) {
    qs_load_knots(kn, A.st);
    for (int i = tid; i < QS_BINS; i += blockDim.x) s_hist[i] = 0u;
  }
  for (int e = tid; e < K * ITEMS * 32; e += blockDim.x) {
    const int k = e / (ITEMS * 32), i = e - k * (ITEMS * 32);
    uint32_t t = A.stage_bytes;  // the NaN row
    if (i < A.L.cnt[k]) {
      const int s = A.src_of_pos[A.L.off[k] + i];
      t = ((uint32_t)s * ROWB) | (fr_swz<GB>(s) << 4);
    }
    tabK[e] = t;
  }
  for (int e = tid; e < S * (int)(ROWB / 8); e += blockDim.x)
    reinterpret_cast<double*>(fr_smem + (size_t)(e / (ROWB / 8)) * stride + A.stage_bytes)[e % (ROWB / 8)] = nan("");
  const uint32_t bar0 = fr_smem_u32(bars);
  if (tid == 0) {
    *s_ticket = 0u;
    for (int i = 0; i < S; ++i) {
      // full[i]: the TMA thread alone, or every thread of the plain loader; empty[i]: one arrival per unit ticket
      fr_mbar_init(bar0 + 8 * i, A.use_tma ? 1u : (uint32_t)(FR_PROD_WARPS * 32));
      fr_mbar_init(bar0 + 8 * (FR_MAX_STAGES + i), (uint32_t)(GB * K));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  auto tile_bits = [&](int t, int* node_base) -> unsigned {
    const int r0 = t * GB, w = r0 >> 5, sh = r0 & 31;
    const unsigned bits = __ldg(A.bitmap + w);
    *node_base = __ldg(A.word_base + w) + __popc(bits & ((1u << sh) - 1u));
    return (bits >> sh) & ((1u << GB) - 1u);
  };
  if (warp >= NW) {
    // ===== producers =====
    const int ptid = tid - NW * 32;
    if (!A.use_tma || ptid == 0) {
      int it = 0, chunk_ready = -1;
      for (int t = A.tile_lo + (int)blockIdx.x; t < A.tile_hi; t += (int)gridDim.x) {
        int nb;
        if (!tile_bits(t, &nb)) continue;
        if (A.gate) {  // the chunk that holds the tile's last row (chunks land in order)
          int c = (int)(((int64_t)t * GB + (GB - 1) - A.row_off) / A.gate_rows);
          c = c < A.gate_chunks ? c : A.gate_chunks - 1;
          if (c > chunk_ready) {
            gate_wait(A.gate, c, A.err);
            asm volatile("fence.proxy.async.global;" ::: "memory");  // the TMA reads what the copy engine wrote
            chunk_ready = c;
          }
        }
        const int stg = it % S, use = it / S;
        const uint32_t full = bar0 + 8 * stg, empty = bar0 + 8 * (FR_MAX_STAGES + stg);
        if (use > 0) fr_mbar_wait(empty, (uint32_t)(use - 1) & 1u);
        const uint32_t dst = fr_smem_u32(fr_smem) + (uint32_t)stg * stride;
        if (A.use_tma) {
          fr_mbar_arrive_expect_tx(full, A.stage_bytes);
          const int c0 = t * GB - (int)A.row_off;
          for (int b = 0; b < A.nbox; ++b)
            fr_tma_load_2d(dst + (uint32_t)b * (uint32_t)A.box1 * ROWB, &tmap, c0, b * A.box1, full);
        } else {
          // plain loader: one 8-byte cp.async per element into the same swizzled layout
          const int n = A.L.n;
          const int64_t r0 = (int64_t)t * GB - A.row_off;
          if (A.colmajor) {
            for (int e = ptid; e < GB * n; e += FR_PROD_WARPS * 32) {
              const int g = e % GB, s = e / GB;
              if (r0 + g < A.rows_avail)
                fr_cp_async8(dst + ((((uint32_t)s * ROWB) | (fr_swz<GB>(s) << 4)) ^ ((uint32_t)g * 8u)),
                             A.assay + (r0 + g) + (int64_t)s * A.ld);
            }
          } else {
            for (int e = ptid; e < GB * n; e += FR_PROD_WARPS * 32) {
              const int s = e % n, g = e / n;
              if (r0 + g < A.rows_avail)
                fr_cp_async8(dst + ((((uint32_t)s * ROWB) | (fr_swz<GB>(s) << 4)) ^ ((uint32_t)g * 8u)),
                             A.assay + (r0 + g) * A.ld + s);
            }
          }
          fr_cp_async_arrive(full);
        }
        ++it;
      }
    }
  } else if (warp < S * GB * K) {
    // ===== consumers =====  (at most `stages` tiles' worth of tickets in flight: a stage's barriers then never see
    // arrivals of two phases at once)
    // Units (row of the tile, category) are handed out by a CTA-wide ticket counter in tile order, so whichever warp is
    // free takes the next one: no warp waits for the slowest of its tile.  A ticket for a row that is no node is a
    // no-op; every ticket arrives once on its stage's "empty" barrier.
    const int UPT = GB * K;
    int it_cur = -1, t_cur = A.tile_lo + (int)blockIdx.x - (int)gridDim.x, node_base = 0;
    unsigned mask = 0u;
    for (;;) {
      unsigned ticket = 0u;
      if (lane == 0) ticket = atomicAdd(s_ticket, 1u);
      ticket = __shfl_sync(0xffffffffu, ticket, 0);
      const int it = (int)(ticket / (unsigned)UPT), u = (int)(ticket - (unsigned)it * (unsigned)UPT);
      while (it_cur < it) {  // walk to the it-th non-empty tile of this CTA
        t_cur += (int)gridDim.x;
        if (t_cur >= A.tile_hi) break;
        mask = tile_bits(t_cur, &node_base);
        if (mask) ++it_cur;
      }
      if (t_cur >= A.tile_hi) break;
      const int stg = it % S, use = it / S;
      const int g = u / K, k = u - g * K;
      fr_mbar_wait(bar0 + 8 * stg, (uint32_t)use & 1u);  // (also for a no-op ticket: its arrival belongs to this phase)
      if ((mask >> g) & 1u) {
        const int node = node_base + __popc(mask & ((1u << g) - 1u));
        front_unit<ITEMS>(A, fr_smem_u32(fr_smem) + (uint32_t)stg * stride, tabK + k * (ITEMS * 32), kn, s_hist,
                          s_cand + warp * 32, s_ncand + warp, g, node, k, lane);
      }
      __syncwarp();
      if (lane == 0) fr_mbar_arrive(bar0 + 8 * (FR_MAX_STAGES + stg));
    }
  }
  __syncthreads();
  // K3 S1/S2: flush this CTA's histogram; the last CTA turns the histogram into slots
  if (do_hist) qs_flush_and_scan(s_hist, A.hist, A.st, A.pct, A.P, A.cand_capacity);
}

// K3 S3: candidates of the quantile slots from the bin ids (8 per 16-byte load).  The ~0.6 % of values that fall into
// a slot are looked up in D (just written, L2-resident) and staged in shared memory; a CTA reserves room in a slot's
// list with ONE global atomic per slot and flush (50,000 same-address atomics would serialise for tens of
// microseconds), then copies its keys.  Fat slots (tie groups, never copied) only track their key range.
constexpr int QSC_STAGE = 2048;  // staged hits per CTA (a CTA sees ~350 on continuous data)
constexpr int QSC_THREADS = 1024;  // one fat CTA per SM: the flush costs one same-address global atomic per slot and CTA
__global__ void __launch_bounds__(QSC_THREADS) k_qs_compact(const double* __restrict__ D,
                                                    const unsigned short* __restrict__ B16, int npad,
                                                    const int32_t* __restrict__ n_nodes_p, QsCompactArgs Q) {
  __shared__ __align__(16) unsigned char s_tab[QSC_BINS + 16];  // [QSC_BINS] = 0xff: where padding / NaN ids land
  __shared__ unsigned long long st_key[QSC_STAGE];              // phase 1: position of the hit in D; phase 2: its key
  __shared__ unsigned char st_slot[QSC_STAGE];
  __shared__ unsigned char s_fat[QS_MAX_T];
  __shared__ unsigned int s_n, s_cnt[QS_MAX_T], s_base[QS_MAX_T];
  const long long total = (long long)(*n_nodes_p) * npad;
  const int U = *Q.U;
  if (threadIdx.x == 0) s_n = 0u;
  if (threadIdx.x < 16) s_tab[QSC_BINS + threadIdx.x] = 0xff;
  for (int i = threadIdx.x; i < QS_MAX_T; i += blockDim.x) {
    s_cnt[i] = 0u;
    s_fat[i] = i < U ? (unsigned char)(Q.slot_fat[i] != 0) : 0;
  }
  qsc_load_table(Q, s_tab);  // (ends with a CTA barrier)
  const long long nvec = (total + 7) >> 3;
  const uint4* __restrict__ bv = reinterpret_cast<const uint4*>(B16);
  // Phase 1 only RECORDS the hits (position + slot): the look-up of the value in D is a dependent global load, and a
  // warp that does it inside the scan pays its latency once per hit position (~6 of the 32 ids a lane visits per round
  // have a hit somewhere in the warp).  Phase 2 fetches all recorded values at once, one hit per thread.
  auto hit_direct = [&](unsigned sl, long long idx) {  // staging full (a slot bin that holds a large share of the data)
    const unsigned long long key = dkey(D[idx]);
    if (s_fat[sl]) {
      if (key < __ldcg(Q.slot_min + sl)) atomicMin(Q.slot_min + sl, key);
      if (key > __ldcg(Q.slot_max + sl)) atomicMax(Q.slot_max + sl, key);
    } else {
      const unsigned pos = atomicAdd(Q.slot_cursor + sl, 1u);
      if (pos < Q.slot_count[sl]) Q.cand[Q.slot_off[sl] + pos] = key;
    }
  };
  auto hit = [&](unsigned sl, long long idx) {  // ~0.6 % of the values
    const unsigned at = atomicAdd(&s_n, 1u);
    if (at < (unsigned)QSC_STAGE) {
      st_key[at] = (unsigned long long)idx;
      st_slot[at] = (unsigned char)sl;
    } else {
      hit_direct(sl, idx);
    }
  };
  auto visit = [&](const uint4& w, long long i) {
    const unsigned ws[4] = {w.x, w.y, w.z, w.w};
    const bool whole = (i + 1) * 8 <= total;
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      // bin ids >= 4096 (padding, NaN) index the table's spill entry 4096..: clamp with one min instead of a branch
      const unsigned b0 = min(ws[h] & 0xffffu, (unsigned)QSC_BINS), b1 = min(ws[h] >> 16, (unsigned)QSC_BINS);
      const unsigned s0 = s_tab[b0], s1 = s_tab[b1];
      if ((s0 & s1) != 0xffu) {  // at least one of the two is a slot value
        if (s0 != 0xffu && (whole || i * 8 + 2 * h < total)) hit(s0, i * 8 + 2 * h);
        if (s1 != 0xffu && (whole || i * 8 + 2 * h + 1 < total)) hit(s1, i * 8 + 2 * h + 1);
      }
    }
  };
  constexpr int VPT = 4;  // 16-byte loads in flight per thread
  const long long step = (long long)gridDim.x * blockDim.x * VPT;
  for (long long i0 = (long long)blockIdx.x * blockDim.x * VPT + threadIdx.x; i0 < nvec; i0 += step) {
    uint4 w[VPT];
#pragma unroll
    for (int r = 0; r < VPT; ++r) {
      const long long i = i0 + r * (long long)blockDim.x;
      w[r] = i < nvec ? __ldg(bv + i) : make_uint4(~0u, ~0u, ~0u, ~0u);
    }
#pragma unroll
    for (int r = 0; r < VPT; ++r) visit(w[r], i0 + r * (long long)blockDim.x);
  }
  __syncthreads();
  const unsigned n = s_n < (unsigned)QSC_STAGE ? s_n : (unsigned)QSC_STAGE;
  if (n == 0u) return;
  // Phase 2: the values of the recorded hits; fat slots (tie groups, never copied) only track their key range
  for (unsigned i = threadIdx.x; i < n; i += blockDim.x) {
    const unsigned sl = st_slot[i];
    const unsigned long long key = dkey(D[(long long)st_key[i]]);
    st_key[i] = key;
    if (s_fat[sl]) {
      if (key < __ldcg(Q.slot_min + sl)) atomicMin(Q.slot_min + sl, key);
      if (key > __ldcg(Q.slot_max + sl)) atomicMax(Q.slot_max + sl, key);
      st_slot[i] = 0xff;  // nothing to copy
    } else {
      atomicAdd(&s_cnt[sl], 1u);
    }
  }
  __syncthreads();
  // flush: ONE global atomic per slot and CTA reserves the room, then the staged keys are copied
  for (int sl = threadIdx.x; sl < U; sl += blockDim.x) {
    const unsigned c = s_cnt[sl];
    s_base[sl] = c ? atomicAdd(Q.slot_cursor + sl, c) : 0u;
    s_cnt[sl] = 0u;  // reused as the running offset inside the reservation
  }
  __syncthreads();
  for (unsigned i = threadIdx.x; i < n; i += blockDim.x) {
    const unsigned sl = st_slot[i];
    if (sl == 0xffu) continue;
    const unsigned pos = s_base[sl] + atomicAdd(&s_cnt[sl], 1u);
    if (pos < Q.slot_count[sl]) Q.cand[Q.slot_off[sl] + pos] = st_key[i];
  }
}

// host side: shape of the staging ring for a run, or ok == false when the fused kernel does not apply
struct FrontPlan {
  bool ok;
  int GB, items, nbox, box1, stages;
  uint32_t stage_bytes;
  size_t smem;
};
inline size_t front_fixed_smem(int K, int items, int nw) {
  return (size_t)QS_BINS * 4 + sizeof(QsKnots) + (size_t)nw * 32 * 8 + 2 * FR_MAX_STAGES * 8 + (size_t)nw * 4 + 16 +
         (size_t)K * items * 32 * 4 + 64;
}
inline int front_items(int max_cnt) { return max_cnt <= 128 ? 4 : (max_cnt <= 384 ? 12 : (max_cnt <= 640 ? 20 : 32)); }
inline int front_warps(int items) { return items <= 12 ? 24 : 16; }
inline FrontPlan front_plan(int n, int K, int max_cnt, size_t static_smem) {
  FrontPlan P{};
  if (max_cnt > 1024 || n < 1) return P;
  P.items = front_items(max_cnt);
  const int nw = front_warps(P.items);
  // boxes of <= 256 samples, a multiple of 8 (the swizzle pattern repeats every 8 rows): least padding wins
  const int nb0 = (n + 255) / 256;
  int best_rows = 1 << 30;
  for (int nb = nb0; nb <= nb0 + 4; ++nb) {
    const int b1 = (((n + nb - 1) / nb) + 7) & ~7;
    if (b1 > 256) continue;
    if (nb * b1 < best_rows) {
      best_rows = nb * b1;
      P.nbox = nb;
      P.box1 = b1;
    }
  }
  const size_t fixed = front_fixed_smem(K, P.items, nw) + static_smem + 1024;
  if (fixed >= (size_t)FR_SMEM_MAX) return P;
  const size_t avail = (size_t)FR_SMEM_MAX - fixed;
  for (int gb : {16, 8, 4}) {
    const size_t sb = (size_t)best_rows * gb * 8;
    const int st = (int)(avail / (sb + 1024));  // (+ the stage's NaN row)
    if (st >= 2 || (gb == 4 && st >= 1)) {
      P.ok = true;
      P.GB = gb;
      P.stage_bytes = (uint32_t)sb;
      P.stages = st > FR_MAX_STAGES ? FR_MAX_STAGES : st;
      P.smem = (size_t)P.stages * (sb + 1024) + front_fixed_smem(K, P.items, nw) + 1024;  // + alignment slack
      return P;
    }
  }
  return P;
}

}  // namespace mdeggs

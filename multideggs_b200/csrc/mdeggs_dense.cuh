// mdeggs_dense.cuh — K4D: pair fits for DENSE networks (e.g. the user-supplied all-pairs network of
// BASELINE config 5: 4,000 features, 7,998,000 edges).
//
// Same arithmetic as k_fit_stream (fit_lm_interaction, R/core_functions.R:1058-1074; aov row 3, 966-968):
// every edge needs the K within-category centred cross moments Sxy_k(a,b).  When the edge list covers a
// large fraction of all node pairs, streaming two 16n-byte rows per edge re-reads each gene row thousands of
// times (config 5: 256 GB of L1/L2 traffic).  Here a CTA stages a 128-gene x 8-sample tile of two row blocks
// in shared memory (centred on the fly) and accumulates the 128x128 block of cross moments in registers with the FP64
// tensor-core instruction (mma.sync.m8n8k4.f64, 64 accumulators per thread; the 8x8-FMA form stays behind
// DENSE_USE_MMA = 0), so each row is read nn/128 times instead of ~nn times.  (tcgen05 has no FP64 type.)
// Only tile pairs with ti <= tj are computed (Sxy is symmetric); k_finish (src 1) then finishes each edge
// from Sxy[min(a,b)][max(a,b)] with the same closed forms as the streaming path (finish_pair).
// The kernel the engine launches is the TMA-fed form at the end of this file (k_center_rows + k_dense_tma: 16-sample
// stages of a centred copy, mbarrier ring, no CTA barrier in the sample loop); k_dense_moments below is its fallback.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "mdeggs_kernels.cuh"

namespace mdeggs {

constexpr int DT = 128;   // tile edge (genes)
constexpr int DKC = 8;    // samples per shared-memory stage (2 stages x 2 tiles = 33 KB static smem)
// DENSE_USE_MMA = 1: the 128 x 128 tile is computed with the FP64 tensor-core instruction mma.sync.m8n8k4 (a warp owns
// 32 x 64 = 4 x 8 MMA tiles; same 64 accumulators per thread as the FMA form, 8x fewer instructions per flop);
// 0: 8 x DCOLS FMA thread tiles.
#ifndef DENSE_USE_MMA
#define DENSE_USE_MMA 1
#endif
// padded leading dimension of the transposed stage [DKC][DLD]: DLD = 8 (mod 16) keeps the 8-row x 4-sample fragment
// loads of the MMA form at the minimum of two shared-memory wavefronts
constexpr int DLD = DENSE_USE_MMA ? DT + 8 : DT + 2;

// Thread tile 8 rows x DCOLS columns: DCOLS = 8 -> 256 threads, 64 accumulators (one 234-register CTA per SM, two
// warps per scheduler); DCOLS = 4 -> 512 threads, 32 accumulators, four warps per scheduler.  Measured on config 5:
// 1.80 ms with 8, 1.87 ms with 4 (more warps do not raise the FP64 issue rate, the extra operand loads cost more).
#ifndef DENSE_COLS
#define DENSE_COLS 8
#endif
constexpr int DCOLS = DENSE_COLS;
constexpr int DTHREADS = (DT / 8) * (DT / DCOLS);  // 256 or 512
constexpr int DQ = 512 / DTHREADS;                 // double2 per matrix and stage that a thread stages (2 or 1)
constexpr int DENSE_DYN_SMEM = 2 * DQ * DTHREADS * 16;  // raw staging slots (cp.async targets), dynamic shared memory

// S[k][i][j] (row-major nn x nn per category), written for tile pairs ti <= tj
__global__ void __launch_bounds__(DTHREADS, 1) k_dense_moments(const double* __restrict__ D, SegLayout L,
                                                               const double* __restrict__ mean, int nn, int ntiles,
                                                               const int32_t* __restrict__ row_range,
                                                               double* __restrict__ S) {
  __shared__ __align__(16) double As[2][DKC][DLD];
  __shared__ __align__(16) double Bs[2][DKC][DLD];
  // blockIdx.x -> (ti, tj) with ti <= tj  (row-wise enumeration of the upper triangle)
  int ti = 0, rem = blockIdx.x;
  while (rem >= ntiles - ti) {
    rem -= ntiles - ti;
    ++ti;
  }
  const int tj = ti + rem;
  // multi-GPU: this rank only needs the tile rows that its edge block touches (row = min(a, b) of an edge)
  if (row_range && (ti < row_range[0] / DT || ti > row_range[1] / DT)) return;
  const int k = blockIdx.y;
  const int seg_beg = L.off[k], seg_len = L.off[k + 1] - L.off[k];
  if (L.cnt[k] == 0) return;
  const int i0 = ti * DT, j0 = tj * DT;
  constexpr int NX = DT / DCOLS;  // threads along the columns (16 or 32)
  constexpr int CG = DCOLS / 2;   // column pairs per thread, NX * 2 columns apart
  const int tid = threadIdx.x, tx = tid % NX, ty = tid / NX;
  // loader mapping: 128 rows x 8 samples = 512 double2 per matrix; thread handles rows lr + (128 / DQ) q, sample pair lc
  const int lc = tid & 3, lr = tid >> 2;
  double ma[DQ], mb[DQ];
  const double* pa[DQ];
  const double* pb[DQ];
#pragma unroll
  for (int q = 0; q < DQ; ++q) {
    int ra = i0 + lr + (DT / DQ) * q, rb = j0 + lr + (DT / DQ) * q;
    bool va = ra < nn, vb = rb < nn;
    ma[q] = va ? mean[(size_t)ra * L.K + k] : 0.0;
    mb[q] = vb ? mean[(size_t)rb * L.K + k] : 0.0;
    pa[q] = va ? D + (size_t)ra * L.npad + seg_beg + 2 * lc : nullptr;
    pb[q] = vb ? D + (size_t)rb * L.npad + seg_beg + 2 * lc : nullptr;
  }
#if DENSE_USE_MMA
  static_assert(DTHREADS == 256, "MMA form: 8 warps, warp tile 32 x 64");
  const int lane = tid & 31, wid = tid >> 5;
  const int wrow = (wid >> 1) * 32, wcol = (wid & 1) * 64;  // warp tile origin inside the CTA tile
  const int fr = lane >> 2, fk = lane & 3;                  // fragment row / sample of this lane
  double acc[4][8][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
#else
  double acc[8][DCOLS];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < DCOLS; ++j) acc[i][j] = 0.0;
#endif

  // Staging of the next 8-sample stage: global -> the thread's OWN 16-byte slots of `raw` by cp.async (no registers
  // are held across the FMAs of a stage; with register staging ptxas sank the loads behind the FMA block), then
  // centred + transposed into As/Bs by the same thread.
  extern __shared__ __align__(16) double2 raw[];  // [2 DQ][DTHREADS]: slot u of thread t at raw[u * DTHREADS + t]
  auto gload = [&](int kk) {  // segment length is a multiple of 4, DKC = 8: guard the tail per double2
    const bool in = kk + 2 * lc < seg_len;
#pragma unroll
    for (int q = 0; q < DQ; ++q) {
      double2* da = raw + (2 * q) * DTHREADS + tid;
      double2* db = raw + (2 * q + 1) * DTHREADS + tid;
      if (in && pa[q]) {
        const unsigned sa = (unsigned)__cvta_generic_to_shared(da);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(pa[q] + kk) : "memory");
      } else {
        *da = make_double2(ma[q], ma[q]);
      }
      if (in && pb[q]) {
        const unsigned sb = (unsigned)__cvta_generic_to_shared(db);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(sb), "l"(pb[q] + kk) : "memory");
      } else {
        *db = make_double2(mb[q], mb[q]);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  auto sstore = [&](int buf) {  // transposed + centred
    asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
    for (int q = 0; q < DQ; ++q) {
      const double2 va = raw[(2 * q) * DTHREADS + tid], vb = raw[(2 * q + 1) * DTHREADS + tid];
      const int row = lr + (DT / DQ) * q;
      As[buf][2 * lc][row] = va.x - ma[q];
      As[buf][2 * lc + 1][row] = va.y - ma[q];
      Bs[buf][2 * lc][row] = vb.x - mb[q];
      Bs[buf][2 * lc + 1][row] = vb.y - mb[q];
    }
  };
  gload(0);
  sstore(0);
  __syncthreads();
  int buf = 0;
  for (int kk = 0; kk < seg_len; kk += DKC) {
    const bool more = kk + DKC < seg_len;
    if (more) gload(kk + DKC);
#if DENSE_USE_MMA
#pragma unroll
    for (int k0 = 0; k0 < DKC; k0 += 4) {  // D[8x8] += A[8 rows x 4 samples] * B[4 samples x 8 rows]
      double a[4], b[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[buf][k0 + fk][wrow + 8 * i + fr];
#pragma unroll
      for (int j = 0; j < 8; ++j) b[j] = Bs[buf][k0 + fk][wcol + 8 * j + fr];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j)
          asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                       : "+d"(acc[i][j][0]), "+d"(acc[i][j][1])
                       : "d"(a[i]), "d"(b[j]));
    }
#else
#pragma unroll
    for (int s = 0; s < DKC; ++s) {
      double a[8], b[DCOLS];
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        double2 va = *reinterpret_cast<const double2*>(&As[buf][s][ty * 2 + 32 * g]);
        a[2 * g] = va.x;
        a[2 * g + 1] = va.y;
      }
#pragma unroll
      for (int g = 0; g < CG; ++g) {
        double2 vb = *reinterpret_cast<const double2*>(&Bs[buf][s][tx * 2 + 2 * NX * g]);
        b[2 * g] = vb.x;
        b[2 * g + 1] = vb.y;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < DCOLS; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
    }
#endif
    if (more) {
      sstore(buf ^ 1);
      __syncthreads();
      buf ^= 1;
    }
  }
  double* Sk = S + (size_t)k * nn * nn;
#if DENSE_USE_MMA
  (void)tx;
  (void)ty;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = i0 + wrow + 8 * i + fr;
    if (r >= nn) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = j0 + wcol + 8 * j + 2 * fk;
      if (c + 1 < nn && !(nn & 1)) {  // 16-byte aligned only when the row pitch is even
        *reinterpret_cast<double2*>(Sk + (size_t)r * nn + c) = make_double2(acc[i][j][0], acc[i][j][1]);
      } else {
        if (c < nn) Sk[(size_t)r * nn + c] = acc[i][j][0];
        if (c + 1 < nn) Sk[(size_t)r * nn + c + 1] = acc[i][j][1];
      }
    }
  }
#else
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = i0 + ty * 2 + 32 * (i >> 1) + (i & 1);
    if (r >= nn) continue;
#pragma unroll
    for (int g = 0; g < CG; ++g) {
      const int c = j0 + tx * 2 + 2 * NX * g;
      if (c + 1 < nn && !(nn & 1)) {  // 16-byte aligned only when the row pitch is even
        *reinterpret_cast<double2*>(Sk + (size_t)r * nn + c) = make_double2(acc[i][2 * g], acc[i][2 * g + 1]);
      } else {
        if (c < nn) Sk[(size_t)r * nn + c] = acc[i][2 * g];
        if (c + 1 < nn) Sk[(size_t)r * nn + c + 1] = acc[i][2 * g + 1];
      }
    }
  }
#endif
}

// ------------------------------------------------------------------------------------------------
// TMA-fed form of the same kernel (the one the engine launches; the cp.async form above stays as the fallback for a
// driver without cuTensorMapEncodeTiled).
//
// Operands come from Dc, a CENTRED gene-major copy of D written by k_center_rows: Dc[node][npadc], every category
// segment starting on a multiple of DS samples and padded with zeros (x - mean = 0 adds nothing to a cross moment), so
//   * a stage is two TMA boxes (DS samples x 128 genes of the row block and of the column block, 16 KB each; one on
//     the diagonal), landing gene-major with the 128-byte hardware swizzle -- no per-thread loads, no transposition,
//     no centring and no CTA barrier inside the sample loop: a producer warp keeps a ring of DSTAGES stages full
//     (mbarrier full / empty pairs), the 8 MMA warps only wait on the stage they are about to read;
//   * a lane fetches its A / B fragments as 16-byte pairs of consecutive samples (s, s + 1) of one gene: the .x halves
//     feed one m8n8k4 step and the .y halves the next (a sum over samples does not care which four go together), i.e.
//     12 LDS.128 per 64 MMAs, each at the 4-wavefront minimum thanks to the swizzle.
// Multi-GPU: bit ti * ntiles + tj of `tile_bits` (k_mark_tiles over the rank's edge block) lets a CTA whose tile holds no
// edge of this rank exit at once.
// ------------------------------------------------------------------------------------------------
constexpr int DS = 16;                      // samples per stage (128-byte rows: the SWIZZLE_128B span)
constexpr int DSTAGES = 4;
constexpr int DTILE_BYTES = DT * DS * 8;    // 16 KB per operand and stage
constexpr int DENSE_TMA_SMEM = DSTAGES * 2 * DTILE_BYTES + 1024 + 2 * DSTAGES * 8 + 64;
constexpr int DENSE_TMA_THREADS = 256 + 32;  // 8 MMA warps + the producer warp

struct DenseLayout {  // layout of the centred copy
  int K, npadc;
  int off[MDEGGS_MAX_K + 1];  // multiples of DS
};

// Dc[node][Lc.off[k] + i] = D[node][L.off[k] + i] - mean[node][k] for i < cnt[k], 0 beyond; one warp per (node, k)
__global__ void __launch_bounds__(256) k_center_rows(const double* __restrict__ D, SegLayout L, DenseLayout Lc,
                                                     const double* __restrict__ mean,
                                                     const int32_t* __restrict__ n_nodes_p, double* __restrict__ Dc) {
  const int nn = *n_nodes_p;
  const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= (long long)nn * L.K) return;
  const int node = (int)(w / L.K), k = (int)(w - (long long)node * L.K);
  const double mu = mean[(size_t)node * L.K + k];
  const double2* __restrict__ src = reinterpret_cast<const double2*>(D + (size_t)node * L.npad + L.off[k]);
  double2* __restrict__ dst = reinterpret_cast<double2*>(Dc + (size_t)node * Lc.npadc + Lc.off[k]);
  const int m = L.cnt[k], len2 = (Lc.off[k + 1] - Lc.off[k]) >> 1, src2 = (L.off[k + 1] - L.off[k]) >> 1;
  for (int i = lane; i < len2; i += 32) {
    double2 v = make_double2(0.0, 0.0);
    if (i < src2) {
      const double2 x = __ldg(src + i);
      v.x = 2 * i < m ? x.x - mu : 0.0;
      v.y = 2 * i + 1 < m ? x.y - mu : 0.0;
    }
    dst[i] = v;
  }
}

__device__ __forceinline__ uint32_t dn_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void dn_mbar_wait(uint32_t bar, uint32_t parity) {
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

__global__ void __launch_bounds__(DENSE_TMA_THREADS, 1)
    k_dense_tma(const __grid_constant__ CUtensorMap tmap, DenseLayout Lc, SegLayout L, int nn, int ntiles,
                const unsigned int* __restrict__ tile_bits, double* __restrict__ S) {
  extern __shared__ unsigned char dn_raw[];
  // blockIdx.x -> (ti, tj), ti <= tj: the off-diagonal pairs row by row, then the diagonal tiles.  A diagonal tile only
  // needs its upper triangle (k_finish reads S[min][max]), so the two warps that own rows 64.. x columns ..63 idle
  // there; with the diagonal units at the END of the grid the last, partial wave consists of these cheaper units.
  const int n_off = ntiles * (ntiles - 1) / 2;
  int ti, tj;
  if ((int)blockIdx.x >= n_off) {
    ti = tj = (int)blockIdx.x - n_off;
  } else {
    ti = 0;
    int rem = blockIdx.x;
    while (rem >= ntiles - 1 - ti) {
      rem -= ntiles - 1 - ti;
      ++ti;
    }
    tj = ti + 1 + rem;
  }
  if (tile_bits && !((tile_bits[(ti * ntiles + tj) >> 5] >> ((ti * ntiles + tj) & 31)) & 1u)) return;
  const int k = blockIdx.y;
  if (L.cnt[k] == 0) return;
  const int seg_beg = Lc.off[k], nst = (Lc.off[k + 1] - Lc.off[k]) / DS;
  const int i0 = ti * DT, j0 = tj * DT;
  const bool diag = ti == tj;
  unsigned char* base = dn_raw + ((1024u - (dn_smem_u32(dn_raw) & 1023u)) & 1023u);  // swizzle atoms are 1,024 bytes
  const uint32_t sb = dn_smem_u32(base);
  const uint32_t bar0 = sb + DSTAGES * 2 * DTILE_BYTES;  // full[DSTAGES], empty[DSTAGES]
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) {
    for (int i = 0; i < DSTAGES; ++i) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar0 + 8 * i), "r"(1u) : "memory");
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar0 + 8 * (DSTAGES + i)), "r"(8u) : "memory");
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (wid == 8) {
    // ===== producer =====
    if (lane == 0) {
      const uint32_t bytes = diag ? DTILE_BYTES : 2 * DTILE_BYTES;
      for (int st = 0; st < nst; ++st) {
        const int slot = st % DSTAGES, use = st / DSTAGES;
        const uint32_t full = bar0 + 8 * slot, empty = bar0 + 8 * (DSTAGES + slot);
        if (use > 0) dn_mbar_wait(empty, (uint32_t)(use - 1) & 1u);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full), "r"(bytes) : "memory");
        const uint32_t dstA = sb + (uint32_t)slot * 2 * DTILE_BYTES;
        const int c0 = seg_beg + st * DS;
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dstA),
            "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(c0), "r"(i0), "r"(full)
            : "memory");
        if (!diag)
          asm volatile(
              "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                  dstA + DTILE_BYTES),
              "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(c0), "r"(j0), "r"(full)
              : "memory");
      }
    }
    return;
  }
  // ===== MMA warps =====
  const int wrow = (wid >> 1) * 32, wcol = (wid & 1) * 64;  // warp tile origin inside the CTA tile
  const int fr = lane >> 2, fk = lane & 3;                  // fragment row / sample pair of this lane
  // byte offset of this lane's 16-byte pair inside a 128-byte gene row, per 8-sample half h: chunk (4 h + fk) ^ (gene & 7)
  const uint32_t swz0 = (uint32_t)((fk ^ fr) << 4), swz1 = (uint32_t)(((4 + fk) ^ fr) << 4);
  const uint32_t rowA = (uint32_t)(wrow + fr) * 128u, rowB = (uint32_t)(wcol + fr) * 128u + (diag ? 0u : (uint32_t)DTILE_BYTES);
  double acc[4][8][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  const bool idle = diag && wrow >= 64 && wcol == 0;  // (lower-left quadrant of a diagonal tile: never read)
  for (int st = 0; st < nst; ++st) {
    const int slot = st % DSTAGES, use = st / DSTAGES;
    dn_mbar_wait(bar0 + 8 * slot, (uint32_t)use & 1u);
    const uint32_t stg = sb + (uint32_t)slot * 2 * DTILE_BYTES;
#pragma unroll
    for (int h = 0; h < 2 && !idle; ++h) {
      const uint32_t sw = h ? swz1 : swz0;
      double2 a[4], b[8];
#pragma unroll
      for (int i = 0; i < 4; ++i)
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(a[i].x), "=d"(a[i].y) : "r"(stg + rowA + (uint32_t)i * 1024u + sw));
#pragma unroll
      for (int j = 0; j < 8; ++j)
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(b[j].x), "=d"(b[j].y) : "r"(stg + rowB + (uint32_t)j * 1024u + sw));
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j)
          asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                       : "+d"(acc[i][j][0]), "+d"(acc[i][j][1])
                       : "d"(a[i].x), "d"(b[j].x));
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j)
          asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                       : "+d"(acc[i][j][0]), "+d"(acc[i][j][1])
                       : "d"(a[i].y), "d"(b[j].y));
    }
    __syncwarp();
    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar0 + 8 * (DSTAGES + slot)) : "memory");
  }
  if (idle) return;
  double* Sk = S + (size_t)k * nn * nn;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = i0 + wrow + 8 * i + fr;
    if (r >= nn) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = j0 + wcol + 8 * j + 2 * fk;
      if (c + 1 < nn && !(nn & 1)) {  // 16-byte aligned only when the row pitch is even
        *reinterpret_cast<double2*>(Sk + (size_t)r * nn + c) = make_double2(acc[i][j][0], acc[i][j][1]);
      } else {
        if (c < nn) Sk[(size_t)r * nn + c] = acc[i][j][0];
        if (c + 1 < nn) Sk[(size_t)r * nn + c + 1] = acc[i][j][1];
      }
    }
  }
}

// multi-GPU: the 128 x 128 tiles of S that hold an edge of this rank's block.  Each CTA marks a shared-memory bitmap and
// publishes it with one atomic OR per non-zero word (plain stores of a million threads to the same few bytes serialise
// in L2: measured 175 us for 10^6 edges).
constexpr int DENSE_MARK_WORDS = 2048;  // up to 65,536 tile pairs in shared memory (256 x 256 tiles = 32,768 nodes)
__global__ void __launch_bounds__(256) k_mark_tiles(const int32_t* __restrict__ ea, const int32_t* __restrict__ eb,
                                                    int64_t e_lo, int64_t e_hi, int ntiles,
                                                    unsigned int* __restrict__ tile_bits, const int32_t* __restrict__ err) {
  __shared__ unsigned int s_bits[DENSE_MARK_WORDS];
  if (*err) return;
  const int W = (ntiles * ntiles + 31) >> 5;
  const bool in_smem = W <= DENSE_MARK_WORDS;
  if (in_smem) {
    for (int w = threadIdx.x; w < W; w += blockDim.x) s_bits[w] = 0u;
    __syncthreads();
  }
  unsigned int* bits = in_smem ? s_bits : tile_bits;
  for (int64_t e = e_lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < e_hi; e += (int64_t)gridDim.x * blockDim.x) {
    const int a = ea[e], b = eb[e];
    const int t = ((a < b ? a : b) / DT) * ntiles + (a < b ? b : a) / DT;
    const unsigned m = 1u << (t & 31);
    if (!(bits[t >> 5] & m)) atomicOr(&bits[t >> 5], m);
  }
  if (in_smem) {
    __syncthreads();
    for (int w = threadIdx.x; w < W; w += blockDim.x)
      if (s_bits[w]) atomicOr(&tile_bits[w], s_bits[w]);
  }
}
// [min, max] over this rank's edges of the smaller node index of the edge: the rows of S it will read
__global__ void k_edge_row_range(const int32_t* __restrict__ ea, const int32_t* __restrict__ eb,
                                 int64_t e_lo, int64_t e_hi,
                                 int32_t* __restrict__ row_range, const int32_t* __restrict__ err) {
  if (*err) return;
  int lo = 0x7fffffff, hi = -1;
  for (int64_t e = e_lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < e_hi; e += (int64_t)gridDim.x * blockDim.x) {
    int a = ea[e], b = eb[e];
    int i = a < b ? a : b;
    lo = i < lo ? i : lo;
    hi = i > hi ? i : hi;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0) {
    if (lo != 0x7fffffff) atomicMin(&row_range[0], lo);
    if (hi >= 0) atomicMax(&row_range[1], hi);
  }
}

// dense path pays off when the edge list is a sizeable fraction of all node pairs:
// streaming moves 16 n E bytes through L1/L2, the tiles do ~nn^2 n FP64 FMAs; at comparable rates 16 E > nn^2
inline bool dense_path_candidate(int64_t E, int64_t nn) { return nn >= 256 && 16 * E >= nn * nn; }

}  // namespace mdeggs

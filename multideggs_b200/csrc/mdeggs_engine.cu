// mdeggs_engine.cu — context, pipeline orchestration and the extern "C" ABI of include/mdeggs.h.
//
// One `mdeggs_run_omic` call = one omic layer of get_diffNetworks_singleOmic (R/core_functions.R:258-529)
// from "edges selected" to "best percentile chosen".  Pipeline per GPU (stream s_main unless noted):
//   H2D -> K0 node filter -> K1 gather/transposition (+ K3 value histogram) -> { K3 quantile selection on s_sort }
//        -> K2 gene stats -> K4/K5/K6 pair fits on this rank's edge block -> tail probabilities
//        -> [NCCL all-gather of p/status over NVLink when world > 1] -> K7 percolation masks
//        -> K8 sort-once + per-(percentile,category) Bonferroni/BH + counts -> best percentile -> D2H
// Every unique edge is fitted ONCE (its p-value does not depend on the percentile); the
// reference refits it for each of the 13-15 percentiles.
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <chrono>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "mdeggs.h"
#include "mdeggs_bsort.cuh"
#include "mdeggs_dense.cuh"
#include "mdeggs_kernels.cuh"
#include "mdeggs_rlm.cuh"
#include "mdeggs_napair.cuh"
#include "mdeggs_qsel.cuh"
#include "mdeggs_front.cuh"
#include "mdeggs_hostpack.h"
#include "mdeggs_p2p.cuh"
#include "nccl_dl.h"

using namespace mdeggs;

constexpr int64_t TAIL_SPLINE_MIN_EDGES = 16384;  // below, building the table costs more than it saves

namespace {

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  int gen = 0;  // (re)allocations so far: the same on every rank of a multi-rank job (same sizes, same history)
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    ++gen;
    if (p && cap) cudaFree(p);  // cap == 0 marks a borrowed caller pointer
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() {
    if (p && cap) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  template <class T>
  T* as() const {
    return reinterpret_cast<T*>(p);
  }
};

enum Ev { EV_START, EV_H2D, EV_GATHER, EV_STATS, EV_FIT, EV_TAIL, EV_COLL, EV_PERC, EV_ADJ, EV_END, EV_Q0, EV_Q1, EV_COUNT };

// scalar slots in the `scal` buffer (8-byte each)
enum Scal { SC_NNODES = 0, SC_ERR = 1, SC_BEST = 2, SC_NVALID = 3, SC_ROWLO = 4, SC_ROWHI = 5, SC_DONE = 6, SC_FINQ = 7, SC_NAQ = 8, SC_NORDER = 9, SC_NLOW = 10, SC_COUNT = 12 };
constexpr int64_t FIN_DEFER_MIN_EDGES = 1 << 16;  // below this a second launch costs more than the slow lanes

struct Shard {
  int device = 0, rank = 0, world = 1;
  cudaStream_t s_main = nullptr, s_sort = nullptr, s_copy = nullptr, s_copy2 = nullptr;
  cudaEvent_t ev[EV_COUNT] = {};
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_sample = nullptr, ev_cf = nullptr;
  cudaEvent_t ev_copy_go = nullptr, ev_copy_done = nullptr, ev_copy2_done = nullptr;
  // streamed input copy (stage_inputs): the band of a host matrix crosses PCIe in row chunks on s_copy while the
  // compute phase already runs; chunk c is complete once gate[c] != 0 (a stream-ordered 32-bit write behind its copy)
  // peer-memory all-gather (mdeggs_p2p.cuh), one process per GPU: mappings of every rank's p_edge / status bytes / flags
  struct P2P {
    bool flags_ok = false, ok = false, failed = false;
    int gen_p = -1, gen_st = -1;
    void* peer_p[P2P_MAX] = {};
    void* peer_st[P2P_MAX] = {};
    void* peer_flags[P2P_MAX] = {};
    bool used = false;  // this call pushed its block (k_signal_done closes it)
    PushArgs args;
  } p2p;
  bool tail_sp_on = false;  // this call's t tails come from the interpolation table (k_tail_spline)
  bool gated = false;
  int gate_chunks = 0;
  int64_t gate_rows = 0;  // rows per chunk (a multiple of 16)
  nccl_comm_t comm = nullptr;
  bool smem_attr_set = false, rsel_attr_set = false, dense_attr_set = false;  // cudaFuncSetAttribute is per device
  unsigned front_attr_mask = 0;  // ... per instantiation of k_front_fused
  unsigned na_attr_mask = 0;     // ... per instantiation of k_fit_na_pairs
  int front_bits = 0;            // this call's front path (see stats.fit_path)
  int sm_count = 0;
  // tensor map of the device copy of the assay (k_front_fused), rebuilt when any of its ingredients changes
  CUtensorMap tmap;
  struct TmapKey {
    const void* base;
    int64_t ld, rows;
    int n, gb, box1;
  } tmap_key = {nullptr, 0, 0, 0, 0, 0};
  CUtensorMap dense_tmap;  // ... and of the centred copy Dc (k_dense_tma)
  TmapKey dense_tmap_key = {nullptr, 0, 0, 0, 0, 0};
  bool dense_tma_attr_set = false;
  int act_nch = 1;                 // 16-percentile chunks of the packed activity masks act_t[node][act_nch]
  bool order_tested_only = false;  // this call's K8 order holds only the edges tested at some percentile
  bool select_fused = false;  // this call's percentile selection ran inside k_bh_final
  bool partials_complete = true;  // the all-percentile K8 left the tile carries behind (false: count-only pass)
  bool tiles_complete = true;     // ... and the member counts of every tile (false: count-only pass that skipped tiles)
  bool best_rows_in_emit = false;  // this call's best-percentile rows are produced by k_emit (device callers)
  const double* best_row_src = nullptr;  // ... its p.adj row comes from row best-1 of this P x E matrix
  int64_t row_off = 0, ld_eff = 0;  // the device copy of the assay starts at host row `row_off`
  int64_t rows_avail = 0;           // ... and holds this many rows
  std::vector<char> stage_shadow;   // host copy of what stage_dev currently holds
  const void* stage_shadow_dev = nullptr;
  void* h_stage = nullptr;          // pinned host staging for the tiny per-call inputs (src_of_pos, pct, df)
  size_t h_stage_cap = 0;
  cudaError_t ensure_stage(size_t bytes) {
    if (bytes <= h_stage_cap) return cudaSuccess;
    if (h_stage) cudaFreeHost(h_stage);
    h_stage = nullptr;
    h_stage_cap = 0;
    cudaError_t e = cudaMallocHost(&h_stage, bytes + 256);
    if (e == cudaSuccess) h_stage_cap = bytes + 256;
    return e;
  }
  DevBuf p2p_flags, p2p_x, run_first, tail_sp, gate, st8, act_t, Dc, tile_need, na_q, bh1_state, assay, from, to, ea, eb, stage_dev, bs_count, bs_slot, seg_done, seg_rank, fin_q, hm_p, hm_q1, hm_c, hm_sc, hm_idx, hm_off, seg_pi0, B16, stamps, excl, node_of_row, rowlist, scal, src_of_pos, pct, cutoff, D, cub_tmp, mean, sxx, med, bad, stat, status, p_edge, coef, iters, act, counters, sortk, sortk2, idx,
      idx2, ps, sa, sb, tile_cnt, tile_base, seg_n, tile_min, tile_carry, padj_best, cat_best, cat, padj, med_out,
      small_out, rsel_hist, rsel_state, rsel_cand, dense_S, cf_tab, sxy;
  void release_all() {
    DevBuf* all[] = {&p2p_flags, &p2p_x, &run_first, &tail_sp, &gate, &st8, &act_t, &Dc, &tile_need, &na_q, &bh1_state, &assay, &from, &to, &ea, &eb, &stage_dev, &bs_count, &bs_slot, &seg_done, &seg_rank, &fin_q, &hm_p, &hm_q1, &hm_c, &hm_sc, &hm_idx, &hm_off, &seg_pi0, &B16, &stamps, &excl, &node_of_row, &rowlist, &scal, &src_of_pos, &pct, &cutoff,
                     &D, &cub_tmp, &mean, &sxx, &med, &bad, &stat, &status, &p_edge,
                     &coef, &iters, &act, &counters, &sortk, &sortk2, &idx, &idx2, &ps, &sa, &sb, &tile_cnt,
                     &tile_base, &seg_n, &tile_min, &tile_carry, &padj_best, &cat_best, &cat, &padj, &med_out,
                     &small_out, &rsel_hist, &rsel_state, &rsel_cand, &dense_S, &cf_tab, &sxy};
    for (DevBuf* b : all) b->release();
  }
};

}  // namespace

// Everything that is baked into a captured compute graph: if any of it changes the graph is rebuilt.
struct GraphKey {
  mdeggs_problem in;  // group / pct pointers zeroed (their content reaches the device through staging)
  mdeggs_result out;
  SegLayout L;
  const void *d_assay, *d_from, *d_to;
  int64_t row_off, ld_eff;
  int64_t gate_rows;  // 0: the input copy is complete before the compute phase starts
  size_t dense_cap;
  const void* h_stage;  // k_select_best writes the packed small results into this pinned block
  size_t h_stage_cap;
};

struct mdeggs_ctx {
  std::vector<Shard> shards;
  std::string err;
  std::mutex err_mu;  // (the packing thread of a streamed host-packed call may report an error too)
  mdeggs_stats stats;
  int64_t launches = 0;
  // CUDA-graph replay of the compute phase (K0..K8) for repeated calls with identical shapes and buffers
  int debug_sync = 0;        // MDEGGS_DEBUG_SYNC=1
  int use_graph = 1;         // option "graph" / MDEGGS_GRAPH
  bool capturing = false;
  bool key_valid = false;
  GraphKey key;
  int key_hits = 0;
  cudaGraphExec_t graph_exec = nullptr;
  int64_t graph_launches = 0;
  int graph_fit_path = 0;    // stats.fit_path of the captured call (run_shard_main does not run on replay)
  bool graph_p2p = false;    // ... and whether it pushed its block over peer memory (k_signal_done follows every call)
  int64_t rsel_cap_override = 0;  // option "rsel_cap": shrink the candidate list to exercise the overflow path (tests)
  int fit_path_override = -1;  // -1 auto, 0 force streaming, 1 force dense (MDEGGS_FIT_PATH, tests)
  int bh_onepass = 1;          // option "bh_onepass": 0 keeps the three-launch K8 for small edge lists too (tests)
  int order_filter = 1;        // option "order_filter": 0 = K8 orders every edge, tested somewhere or not (tests)
  int dense_path = -1;         // option "dense_path": 0 = the cp.async form of the dense moments kernel, else the TMA-fed one
  int front_path = -1;         // -1 auto, 0 separate K1 + K2 kernels, 1 fused (TMA when possible), 2 fused with the plain loader
  bool last_zc = false;        // the last call read its assay through zero copy
  // host-side packing of the node rows into pinned staging (mdeggs_hostpack.h)
  int tail_spline = 1;         // option "tail_spline": 0 off, 1 = table of the t tail on edge lists >= 16,384 per rank, 2 always
  int host_pack = 1;           // option "host_pack": 0 off, 1 auto (pageable or scattered node rows), 2 whenever possible
  struct HostPack {
    bool active = false;       // this call runs on the packed problem `in2`
    bool deferred = false;     // ... and its chunks still have to be packed + sent (submit decides when, see there)
    bool last_packed = false;  // the last completed call was packed: mdeggs_pair_moments translates row numbers
    std::vector<int32_t> rows, rank;
    // the edge list the node rows above were derived from: repeated calls with the same network (nestedcv folds,
    // benchmarks) compare (exactly, memcmp) instead of analysing again - 30 instead of 100 us in front of the first DMA
    std::vector<int32_t> seen_from, seen_to;
    int32_t seen_G = -1;
    bool rows_valid = false;   // rows / rank describe the edge list of THIS call (stage_inputs reads the band from them)
    bool idx_ok = false;       // ... and every index was inside 1..G
    std::vector<double> med_tmp;   // medians of the packed problem [K x nodes]; mdeggs_wait spreads them over [K x G]
    double* user_median = nullptr;
    int32_t user_G = 0;
    mdeggs_result out2;
    mdeggs_problem in2;
    void* h_buf = nullptr;     // pinned: packed matrix | from' | to'
    size_t h_cap = 0;
    PackJob job;
  } pack;
  int p2p = 1;                 // option "p2p": 1 = peer-memory push all-gather with one process per GPU, 0 = NCCL
  int launch_async = -1;       // -1 not probed yet, 1 kernel launches return before the kernel ends, 0 they do not
  int copy_streams = 2;        // option "copy_streams": streams the chunks of a streamed input copy alternate between
  int h2d_pipeline = 1;        // option "h2d_pipeline": 0 = the whole band is copied before the compute phase starts
  int zero_copy = 0;           // option "zero_copy": pinned host assays are read by the gather straight over PCIe
  int64_t padj_keep_cap = (int64_t)256 << 20;  // option "padj_keep_cap": largest P x E x 8 scratch kept for the best row
  int count_skip = 1;          // option "count_skip": 0 = the count-only K8 visits every tile of the p order (tests)
  int count_only_adjust = 1;   // option "count_only": 0 keeps the three-pass K8 even when no adjusted value is kept (tests)
  int sort_path_override = -1; // -1 auto, 0 force the CUB radix sort, 1 force the bucket ordering (option "sort_path")
  // MDEGGS_TRACE=1 / option "trace": an event after every eager launch, timeline printed to stderr after the call
  int trace = 0;
  struct TraceRec {
    const char* name;
    cudaStream_t stream;
    cudaEvent_t ev;
  };
  std::vector<TraceRec> trace_recs;
  std::vector<cudaEvent_t> trace_pool;
  size_t trace_used = 0;
  // what the last completed call left on the device (mdeggs_pair_moments reads it)
  bool last_valid = false;
  SegLayout last_L;
  int last_G = 0;
  // a submitted call that has not been waited for yet (mdeggs_submit_omic / mdeggs_wait)
  struct Pending {
    bool active = false;
    mdeggs_problem in;
    mdeggs_result out;
    int32_t n_nodes = 0;
    SegLayout L;
    bool replayed = false;
    int src_device = 0;
    std::chrono::steady_clock::time_point t0, t1, t2, tA;
  } pend;
};

namespace {

void set_err(mdeggs_ctx* c, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  std::lock_guard<std::mutex> lk(c->err_mu);
  c->err = buf;
}

#define CU(call)                                                                               \
  do {                                                                                         \
    cudaError_t e_ = (call);                                                                   \
    if (e_ != cudaSuccess) {                                                                   \
      set_err(ctx, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_));         \
      return e_ == cudaErrorMemoryAllocation ? MDEGGS_ENOMEM : MDEGGS_ECUDA;                   \
    }                                                                                          \
  } while (0)

#define NC(call)                                                                               \
  do {                                                                                         \
    int r_ = (call);                                                                           \
    if (r_ != 0) {                                                                             \
      NcclApi& api_ = nccl_api();                                                              \
      set_err(ctx, "%s:%d %s: %s", __FILE__, __LINE__, #call,                                  \
              api_.GetErrorString ? api_.GetErrorString(r_) : "nccl error");                  \
      return MDEGGS_ENCCL;                                                                     \
    }                                                                                          \
  } while (0)

// stage-boundary timing events are not recorded while the compute phase is being captured into a graph
#define EVREC(ev, stream)                                  \
  do {                                                     \
    if (!ctx->capturing) CU(cudaEventRecord((ev), (stream))); \
  } while (0)

void trace_mark(mdeggs_ctx* ctx, const char* name, cudaStream_t stream) {
  if (!ctx->trace || ctx->capturing) return;
  if (ctx->trace_used == ctx->trace_pool.size()) {
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    ctx->trace_pool.push_back(e);
  }
  cudaEvent_t e = ctx->trace_pool[ctx->trace_used++];
  cudaEventRecord(e, stream);
  ctx->trace_recs.push_back({name, stream, e});
}

#define LAUNCH(kernel, grid, block, smem, stream, ...)                                         \
  do {                                                                                         \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                                \
    ++ctx->launches;                                                                           \
    if (ctx->trace) trace_mark(ctx, #kernel, stream);                                      \
    if (ctx->debug_sync && !ctx->capturing) { /* MDEGGS_DEBUG_SYNC=1: pinpoint a failing launch */ \
      cudaError_t le_ = cudaGetLastError();                                                    \
      if (le_ == cudaSuccess) le_ = cudaStreamSynchronize(stream);                             \
      if (le_ != cudaSuccess)                                                                  \
        fprintf(stderr, "[mdeggs] %s:%d launch of %s failed: %s\n", __FILE__, __LINE__, #kernel, \
                cudaGetErrorString(le_));                                                      \
    }                                                                                          \
  } while (0)

// layout of the packed small results (see k_select_best)
size_t small_block_bytes(int P) { return (size_t)4 * (P > 0 ? P : 1) * 8 + 16 + 16; }
inline unsigned blocks_for(int64_t n, int threads) { return (unsigned)((n + threads - 1) / threads); }

int init_shard(mdeggs_ctx* ctx, Shard& s) {
  CU(cudaSetDevice(s.device));
  CU(cudaStreamCreate(&s.s_main));
  {  // the side stream carries short kernels that must not queue behind the big grids of the main stream
    int prio_lo = 0, prio_hi = 0;
    CU(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    CU(cudaStreamCreateWithPriority(&s.s_sort, cudaStreamDefault, prio_hi));
  }
  for (int i = 0; i < EV_COUNT; ++i) CU(cudaEventCreate(&s.ev[i]));
  CU(cudaEventCreateWithFlags(&s.ev_fork, cudaEventDisableTiming));
  CU(cudaEventCreateWithFlags(&s.ev_join, cudaEventDisableTiming));
  CU(cudaEventCreateWithFlags(&s.ev_sample, cudaEventDisableTiming));
  CU(cudaEventCreateWithFlags(&s.ev_cf, cudaEventDisableTiming));
  CU(cudaStreamCreateWithFlags(&s.s_copy, cudaStreamNonBlocking));
  CU(cudaStreamCreateWithFlags(&s.s_copy2, cudaStreamNonBlocking));
  CU(cudaEventCreateWithFlags(&s.ev_copy2_done, cudaEventDisableTiming));
  CU(cudaEventCreateWithFlags(&s.ev_copy_go, cudaEventDisableTiming));
  CU(cudaEventCreateWithFlags(&s.ev_copy_done, cudaEventDisableTiming));
  CU(s.gate.ensure(GATE_MAX_CHUNKS * 4));
  CU(s.scal.ensure(SC_COUNT * 8));
  CU(cudaDeviceGetAttribute(&s.sm_count, cudaDevAttrMultiProcessorCount, s.device));
  return MDEGGS_OK;
}

struct HostLayout {
  SegLayout L;
  std::vector<int32_t> src_of_pos;
};

void result_df(const mdeggs_problem* in, const HostLayout& H, double* dfh) {
  if (in->K == 2) {
    dfh[0] = 1.0;
    dfh[1] = (double)(in->n - 4);
  } else {
    int Ke = 0;
    for (int k = 0; k < in->K; ++k) Ke += H.L.cnt[k] > 0;
    dfh[0] = (double)(Ke - 1);
    dfh[1] = (double)(in->n - 2 * Ke);
  }
}

int build_layout(mdeggs_ctx* ctx, const mdeggs_problem* in, HostLayout& H) {
  SegLayout& L = H.L;
  memset(&L, 0, sizeof L);
  L.K = in->K;
  L.n = in->n;
  for (int s = 0; s < in->n; ++s) {
    int g = in->group[s];
    if (g < 1 || g > in->K) {
      set_err(ctx, "group[%d] = %d outside 1..K", s, g);
      return MDEGGS_EINVAL;
    }
    L.cnt[g - 1]++;
  }
  int off = 0;
  for (int k = 0; k < in->K; ++k) {
    L.off[k] = off;
    off += (L.cnt[k] + 3) & ~3;
  }
  L.off[in->K] = off;
  L.npad = off;
  H.src_of_pos.assign((size_t)(off > 0 ? off : 1), -1);
  int fill[MDEGGS_MAX_K] = {0};
  for (int s = 0; s < in->n; ++s) {  // stable: samples keep their order inside a category
    int k = in->group[s] - 1;
    H.src_of_pos[(size_t)L.off[k] + fill[k]++] = s;
  }
  return MDEGGS_OK;
}

unsigned int qs_cand_capacity(const mdeggs_ctx* ctx, const mdeggs_problem* in) {
  // candidates of all slots together: a few thousand per slot for continuous data; bounded by the matrix itself
  const long long n_upper = (long long)(2 * in->E < (int64_t)in->G ? 2 * in->E : (int64_t)in->G) * in->n;
  unsigned int cap = (unsigned int)(n_upper < (long long)QS_CAND_MAX ? n_upper : (long long)QS_CAND_MAX);
  if (cap < 1024) cap = 1024;
  if (ctx->rsel_cap_override > 0 && (unsigned int)ctx->rsel_cap_override < cap) cap = (unsigned int)ctx->rsel_cap_override;
  return cap;
}
bool device_adjust(const mdeggs_problem* in) {
  return in->padj == MDEGGS_PADJ_BONFERRONI || in->padj == MDEGGS_PADJ_BH || in->padj == MDEGGS_PADJ_HOLM ||
         in->padj == MDEGGS_PADJ_HOCHBERG || in->padj == MDEGGS_PADJ_BY || in->padj == MDEGGS_PADJ_QVALUE ||
         in->padj == MDEGGS_PADJ_HOMMEL;
}
// K8 ordering by p-value buckets (mdeggs_bsort.cuh) instead of the multi-pass radix sort
bool use_bucket_order(const mdeggs_ctx* ctx, const mdeggs_problem* in, int32_t n_nodes) {
  if (!(device_adjust(in) && in->E > 0 && in->P > 0 && n_nodes > 0)) return false;
  if (ctx->sort_path_override >= 0) return ctx->sort_path_override == 1 && in->E <= BS_MAX_E;
  return in->E <= BS_MAX_E;
}
BsArgs make_bs_args(Shard& s, int64_t E);
inline int bs_padded(int NB) { return ((NB + 1 + 3) & ~3) + 4; }  // entries per array: 16-byte vector access past NB

__global__ void k_stamp(unsigned long long* slot) {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  *slot = t;
}

// complete-case fits of the pairs with missing values (rlm and aov only; see mdeggs_napair.cuh)
struct NaPairPlan {
  bool on = false;
  int wpb = 0, is_rlm = 0, tested_coef = 3, cov_mode = 0;
  size_t smem = 0;
};
NaPairPlan na_pair_plan(const mdeggs_problem* in) {
  NaPairPlan NP;
  const bool is_rlm = in->K == 2 && in->method == MDEGGS_METHOD_RLM;
  if (in->K == 2 && !is_rlm) return NP;  // fit_lm_interaction: .lm.fit throws on NA (core_functions.R:1065)
  const size_t per_warp = (size_t)(is_rlm ? 3 : 2) * (size_t)in->n * 8;
  const size_t budget = 200 * 1024;
  if (per_warp > budget) return NP;      // rows too long for shared memory: such pairs keep MDEGGS_EDGE_NONFINITE
  NP.on = true;
  NP.wpb = (int)(budget / per_warp) < NA_WARPS ? (int)(budget / per_warp) : NA_WARPS;
  NP.smem = per_warp * (size_t)NP.wpb;
  NP.is_rlm = is_rlm ? 1 : 0;
  NP.tested_coef = in->tested_coef;
  NP.cov_mode = in->rlm_cov_mode;
  return NP;
}

template <int K>
void launch_fit_stream(mdeggs_ctx* ctx, Shard& s, const SegLayout& L, int64_t e_lo, int64_t e_hi) {
  int64_t ne = e_hi - e_lo;
  if (ne <= 0) return;
  int64_t want_blocks = (ne + 7) / 8;  // 8 warps per CTA, one edge per warp per trip
  int64_t cap = 148 * 8 * 16;          // at most 16 waves of 8 CTAs/SM (64 resident warps/SM); warps loop beyond that
  unsigned grid = (unsigned)(want_blocks < cap ? want_blocks : cap);
  LAUNCH(k_fit_stream<K>, grid, 256, 0, s.s_main, s.D.as<double>(), L, s.ea.as<int32_t>(), s.eb.as<int32_t>(),
         s.mean.as<double>(), s.bad.as<uint8_t>(), e_lo, e_hi, s.sxy.as<double>(),
         s.scal.as<int32_t>() + 2 * SC_ERR);
}

template <int K>
void launch_finish(mdeggs_ctx* ctx, Shard& s, const SegLayout& L, int src, int nn, int64_t e_lo, int64_t e_hi, int mode,
                   double df1, double df2, const TailConsts& tc, const double* d_lbeta, double* coef, int do_bs,
                   const BsArgs& B, const NaPairPlan& NP) {
  // tails that need the continued fraction (everything but the closed-form F with 2 numerator degrees of freedom):
  // on big edge lists the few slow lanes are queued and finished by a second, dense launch
  const int64_t ne = e_hi - e_lo;
  const bool defer = ne >= FIN_DEFER_MIN_EDGES && !(mode == 1 && df1 == 2.0);
  int32_t* fin_q = nullptr;
  unsigned int* fin_nq = s.scal.as<unsigned int>() + 2 * SC_FINQ;
  if (defer) fin_q = s.fin_q.as<int32_t>();  // (sized by the caller)
  LAUNCH(k_finish<K>, blocks_for(ne, FIN_THREADS), FIN_THREADS, 0, s.s_main, src, s.sxy.as<double>(), s.dense_S.as<double>(),
         nn, L, s.ea.as<int32_t>(), s.eb.as<int32_t>(), s.mean.as<double>(),
         s.sxx.as<double>(), s.bad.as<uint8_t>(), e_lo, e_hi, mode, df1, df2, tc, d_lbeta, s.stat.as<double>(),
         s.status.as<int32_t>(), coef, s.p_edge.as<double>(), s.scal.as<int32_t>() + 2 * SC_ERR, do_bs, B, fin_q, fin_nq,
         NP.on ? s.na_q.as<int32_t>() : nullptr, s.scal.as<unsigned int>() + 2 * SC_NAQ);
  if (defer) {
    const unsigned want = blocks_for(ne, 256);
    LAUNCH(k_finish_slow, want < 148u * 8 ? want : 148u * 8, 256, 0, s.s_main, fin_q, fin_nq, e_lo, mode, df1, df2, tc,
           d_lbeta, s.stat.as<double>(), s.p_edge.as<double>(), s.scal.as<int32_t>() + 2 * SC_ERR, do_bs, B);
  }
  if (NP.on) {
    // rlm / aov: complete-case fits of the queued pairs that touch a row with NA (mdeggs_napair.cuh).  The grid is
    // sized for a modest queue and strides over a longer one; without NA in the data it finds an empty queue and exits.
    const int64_t want = (ne + NP.wpb - 1) / NP.wpb;
    const unsigned grid = (unsigned)(want < 148 * 2 ? want : 148 * 2);
    LAUNCH(k_fit_na_pairs<K>, grid, NP.wpb * 32, NP.smem, s.s_main, s.D.as<double>(), L, s.ea.as<int32_t>(),
           s.eb.as<int32_t>(), s.na_q.as<int32_t>(), s.scal.as<unsigned int>() + 2 * SC_NAQ, e_lo, NP.is_rlm,
           NP.tested_coef, NP.cov_mode, s.stat.as<double>(), s.status.as<int32_t>(), coef,
           NP.is_rlm ? s.iters.as<int32_t>() : nullptr, s.p_edge.as<double>(), s.scal.as<int32_t>() + 2 * SC_ERR, do_bs, B);
  }
}

BsArgs make_bs_args(Shard& s, int64_t E) {
  BsArgs B;
  const int NB = bs_nbuckets(E);
  B.count = s.bs_count.as<int>();
  B.base = B.count + bs_padded(NB);
  B.slot = s.bs_slot.as<int32_t>();
  B.n_valid = s.scal.as<unsigned long long>() + SC_NVALID;
  B.n_order = s.scal.as<unsigned long long>() + SC_NORDER;
  B.n_low = s.scal.as<unsigned long long>() + SC_NLOW;
  B.nlin = bs_nlin(E);
  B.scale = bs_scale(B.nlin);
  B.E = E;
  B.act_t = s.order_tested_only ? s.act_t.as<uint4>() : nullptr;
  B.nch = s.act_nch;
  B.run_first = nullptr;  // (set where the order is built)
  B.ea = s.ea.as<int32_t>();
  B.eb = s.eb.as<int32_t>();
  return B;
}

// arguments of the percentile selection (k_select_best, or the tail of k_bh_final)
SelArgs make_sel_args(Shard& s, const mdeggs_problem* in, const HostLayout& H, bool emit,
                      const unsigned long long* sig_src) {
  const int P = in->P, Pn = P > 0 ? P : 1;
  double dfh[2];
  result_df(in, H, dfh);
  SelArgs S{};
  S.tot = s.counters.as<unsigned long long>();
  S.fail = S.tot + Pn;
  S.sig = sig_src;
  S.P = P;
  S.host_adjust = in->padj == MDEGGS_PADJ_HOST ? 1 : 0;
  S.cutoff = s.cutoff.as<double>();
  S.n_nodes = s.scal.as<int32_t>() + 2 * SC_NNODES;
  S.err = s.scal.as<int32_t>() + 2 * SC_ERR;
  S.df1 = dfh[0];
  S.df2 = dfh[1];
  S.blk = s.small_out.as<long long>();
  S.host_blk =
      emit ? reinterpret_cast<long long*>((char*)s.h_stage + s.h_stage_cap - 64 - small_block_bytes(P)) : nullptr;
  S.best = s.scal.as<int32_t>() + 2 * SC_BEST;
  S.done = nullptr;
  return S;
}

// ---- BH / Bonferroni over a block of percentiles (or the device-chosen best one) -----------------
int run_adjust(mdeggs_ctx* ctx, Shard& s, const mdeggs_problem* in, int n_segments, int p0, int p_stride,
               const int32_t* best_dev, unsigned long long* sig_dev, double* padj_out, int64_t padj_stride,
               int act_stride, bool reuse_tiles = false, const SelArgs* sel = nullptr, bool redo_minima = false,
               double skip_thr = 0.0) {
  if (n_segments <= 0) return MDEGGS_OK;
  const int64_t E = in->E;
  const int ntiles = (int)((E + BH_TILE - 1) / BH_TILE);
  BhArgs A;
  A.ps = s.ps.as<double>();
  A.sa = s.sa.as<int32_t>();
  A.sb = s.sb.as<int32_t>();
  A.sidx = s.idx2.as<int32_t>();
  A.act = s.act.as<uint8_t>();
  A.act_stride = act_stride;
  A.E = E;
  A.n_valid = s.scal.as<unsigned long long>() + SC_NVALID;
  A.n_order = s.scal.as<unsigned long long>() + SC_NORDER;
  A.K = in->K;
  A.P = in->P;
  A.ntiles = ntiles;
  A.p0 = p0;
  A.p_stride = p_stride;
  A.best = best_dev;
  A.tot = s.counters.as<unsigned long long>();  // c_tot[P], filled by k_percolate
  A.reuse_tiles = reuse_tiles ? 1 : 0;
  CU(s.seg_pi0.ensure((size_t)in->P * in->K * 8));
  A.seg_pi0 = s.seg_pi0.as<double>();
  A.skip_thr = 0.0;
  A.n_low = s.scal.as<unsigned long long>() + SC_NLOW;
  A.seg_cnt = nullptr;
  A.seg_low = nullptr;
  dim3 grid((unsigned)ntiles, (unsigned)n_segments);
  // per-segment "CTAs done" counters of the fused tile passes: zero when first allocated, re-armed by the kernels
  {
    const size_t need = (size_t)in->P * in->K * 2 * sizeof(int);
    if (need > s.seg_done.cap) {
      CU(s.seg_done.ensure(need));
      CU(cudaMemsetAsync(s.seg_done.p, 0, s.seg_done.cap, s.s_main));
    }
  }
  int* done_a = s.seg_done.as<int>();
  int* done_b = done_a + (size_t)in->P * in->K;
  // count-only block (no adjusted value is kept): the significant counts come from one rank pass, see k_bh_sigrank
  const bool count_only = sig_dev && !padj_out && !best_dev && !reuse_tiles && ctx->count_only_adjust &&
                          in->padj != MDEGGS_PADJ_HOMMEL;
  if (count_only) {
    const size_t need = (size_t)in->P * in->K * sizeof(unsigned long long);
    if (2 * need > s.seg_rank.cap) {  // [P x K] rank accumulators (zero between launches), then [P x K] seg_low
      CU(s.seg_rank.ensure(2 * need));
      CU(cudaMemsetAsync(s.seg_rank.p, 0, s.seg_rank.cap, s.s_main));
    }
    if (skip_thr > 0.0 && in->padj != MDEGGS_PADJ_QVALUE && use_bucket_order(ctx, in, 1)) {  // (q.value scales by pi0 <= 1: its q can fall below p)
      const int Pn = in->P > 0 ? in->P : 1;
      A.skip_thr = skip_thr;
      A.seg_cnt = reinterpret_cast<const unsigned int*>(s.counters.as<char>() + (size_t)Pn * 8 * 7 + 8);
      A.seg_low = s.seg_rank.as<long long>() + (size_t)in->P * in->K;
    }
  }
  const bool simple_mode = in->padj == MDEGGS_PADJ_BONFERRONI || in->padj == MDEGGS_PADJ_BH || in->padj == MDEGGS_PADJ_BY ||
                           in->padj == MDEGGS_PADJ_HOLM || in->padj == MDEGGS_PADJ_HOCHBERG;
  if (ctx->bh_onepass && simple_mode && !reuse_tiles && !best_dev && !count_only && sig_dev && padj_out &&
      ntiles <= BH1_MAX_TILES && s.bh1_state.p) {  // (padj_out: no later pass wants this pass's tile partials)
    // small edge lists: the whole adjustment in one launch (k_bh_onepass); segment sizes came from k_percolate
    const int Pn = in->P > 0 ? in->P : 1;
    const unsigned int* seg_cnt = reinterpret_cast<const unsigned int*>(s.counters.as<char>() + (size_t)Pn * 8 * 7 + 8);
    const size_t words = (size_t)Pn * in->K * ntiles;
    unsigned long long* st_ext = s.bh1_state.as<unsigned long long>();
    unsigned int* st_cnt = reinterpret_cast<unsigned int*>(st_ext + words);
    LAUNCH(k_bh_onepass, grid, BH_THREADS, 0, s.s_main, A, in->padj, seg_cnt, st_cnt, st_ext, sig_dev, padj_out,
           padj_stride, sel ? *sel : SelArgs{});
    return MDEGGS_OK;
  }
  // count-only passes walk the tiles with the grid's stride: a capped grid (most tiles are skipped, see bh_tiles_visited)
  const dim3 grid_c((unsigned)(ntiles < 2 * s.sm_count ? ntiles : 2 * s.sm_count), (unsigned)n_segments);
  if (!reuse_tiles)
    LAUNCH(k_bh_count, (count_only ? grid_c : grid), BH_THREADS, 0, s.s_main, A, s.tile_cnt.as<int32_t>(),
           s.tile_base.as<long long>(), s.seg_n.as<long long>(), done_a);
  if (in->padj == MDEGGS_PADJ_QVALUE && !reuse_tiles)
    LAUNCH(k_qv_pi0, dim3(1, (unsigned)n_segments), BH_THREADS, 0, s.s_main, A, s.tile_base.as<long long>(),
           s.seg_n.as<long long>(), s.seg_pi0.as<double>());
  if (in->padj == MDEGGS_PADJ_HOMMEL) {  // O(n^2) per segment on dense copies of the segments (all-percentile pass only)
    if (best_dev || reuse_tiles || !sig_dev) {
      set_err(ctx, "internal: hommel is adjusted in the all-percentile pass only");
      return MDEGGS_EINVAL;
    }
    const size_t tot_cap = (size_t)((n_segments + in->K - 1) / in->K) * (size_t)E;  // sum of the segment sizes at most
    CU(s.hm_p.ensure(tot_cap * 8));
    CU(s.hm_q1.ensure(tot_cap * 8));
    CU(s.hm_c.ensure(tot_cap * 8));
    CU(s.hm_sc.ensure(tot_cap * 8));
    CU(s.hm_idx.ensure(tot_cap * 4));
    CU(s.hm_off.ensure(((size_t)n_segments + 1) * 8 + (size_t)n_segments * 8));
    long long* seg_off = s.hm_off.as<long long>();
    double* hinit = reinterpret_cast<double*>(seg_off + n_segments + 1);
    const unsigned gx = (unsigned)((E + 7) / 8);
    LAUNCH(k_hm_offsets, 1, 32, 0, s.s_main, A, s.seg_n.as<long long>(), n_segments, seg_off);
    LAUNCH(k_hm_compact, grid, BH_THREADS, 0, s.s_main, A, s.tile_cnt.as<int32_t>(), s.tile_base.as<long long>(), seg_off,
           s.hm_p.as<double>(), s.hm_idx.as<int32_t>());
    LAUNCH(k_hm_q1, dim3(gx, (unsigned)n_segments), 256, 0, s.s_main, A, s.seg_n.as<long long>(), seg_off,
           s.hm_p.as<double>(), s.hm_q1.as<double>(), s.hm_c.as<double>(), hinit);
    LAUNCH(k_hm_sufmax, (unsigned)n_segments, BH_THREADS, 0, s.s_main, A, s.seg_n.as<long long>(), seg_off,
           s.hm_c.as<double>(), s.hm_sc.as<double>());
    LAUNCH(k_hm_pa, dim3(gx, (unsigned)n_segments), 256, 0, s.s_main, A, s.seg_n.as<long long>(), seg_off,
           s.hm_p.as<double>(), s.hm_q1.as<double>(), s.hm_sc.as<double>(), hinit, s.hm_idx.as<int32_t>(), sig_dev,
           padj_out, padj_stride);
    return MDEGGS_OK;
  }
  if (count_only) {
    LAUNCH(k_bh_sigrank, grid_c, BH_THREADS, 0, s.s_main, A, in->padj, s.tile_cnt.as<int32_t>(),
           s.tile_base.as<long long>(), s.seg_n.as<long long>(), s.seg_rank.as<unsigned long long>(), sig_dev, done_b);
    return MDEGGS_OK;
  }
  // (redo_minima: a count-only pass left the tile counts / ranks of every segment but no tile minima)
  if (in->padj != MDEGGS_PADJ_BONFERRONI && (!reuse_tiles || redo_minima))
    LAUNCH(k_bh_tilemin, grid, BH_THREADS, 0, s.s_main, A, in->padj, s.tile_cnt.as<int32_t>(),
           s.tile_base.as<long long>(), s.seg_n.as<long long>(), s.tile_min.as<double>(),
           s.tile_carry.as<double>(), done_b);
  LAUNCH(k_bh_final, grid, BH_THREADS, 0, s.s_main, A, in->padj, s.tile_cnt.as<int32_t>(), s.tile_base.as<long long>(),
         s.seg_n.as<long long>(),
         s.tile_carry.as<double>(), sig_dev, padj_out, padj_stride, sel ? *sel : SelArgs{});
  return MDEGGS_OK;
}

int front_prepare(mdeggs_ctx* ctx, Shard& s, const mdeggs_problem* in, const SegLayout& L, int32_t n_nodes, FrontPlan* Pout,
                  bool* use_tma_out, bool* ok);

// cuStreamWriteValue32 (a stream-ordered, fenced 32-bit write), resolved through the runtime like the tensor-map encoder
typedef CUresult (*stream_write32_fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
stream_write32_fn stream_write32() {
  static stream_write32_fn fn = []() -> stream_write32_fn {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuStreamWriteValue32", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess) {
      cudaGetLastError();
      return nullptr;
    }
    return reinterpret_cast<stream_write32_fn>(p);
  }();
  return fn;
}


// ---- streamed input copy: chunk geometry is in the shard (gate_rows, gate_chunks), see stage_inputs ----
int copy_chunks_begin(mdeggs_ctx* ctx, Shard& s) {
  // (the copy streams do not wait for s_main: the previous call of this context has been waited for, so nothing reads
  // the device copy or the gate words any more)
  CU(cudaMemsetAsync(s.gate.p, 0, GATE_MAX_CHUNKS * 4, s.s_copy));
  CU(cudaEventRecord(s.ev_copy_go, s.s_copy));
  if (ctx->copy_streams >= 2) CU(cudaStreamWaitEvent(s.s_copy2, s.ev_copy_go, 0));
  // the gates are closed before any kernel of this call can look at them (and the memset never has to find room on an
  // SM beside the waiting front kernel)
  CU(cudaStreamWaitEvent(s.s_main, s.ev_copy_go, 0));
  return MDEGGS_OK;
}
// chunk c = rows [c gate_rows, ...) of a column-major host block whose row 0 is `src` (leading dimension ld_src)
int copy_chunk(mdeggs_ctx* ctx, Shard& s, int c, const double* src, int64_t ld_src, int64_t total_rows, int n) {
  // chunks alternate between two copy streams: the set-up of one chunk's DMA and the flag write behind the other
  // (~10 us per chunk on one stream) no longer leave the link idle; chunks may then land slightly out of order,
  // which the loader tolerates (it waits for the chunk of its own tile)
  cudaStream_t st = (ctx->copy_streams >= 2 && (c & 1)) ? s.s_copy2 : s.s_copy;
  const int64_t r0 = (int64_t)c * s.gate_rows;
  const int64_t rows = total_rows - r0 < s.gate_rows ? total_rows - r0 : s.gate_rows;
  CU(cudaMemcpy2DAsync(s.assay.as<double>() + r0, (size_t)s.ld_eff * 8, src + r0, (size_t)ld_src * 8, (size_t)rows * 8,
                       (size_t)n, cudaMemcpyHostToDevice, st));
  if (stream_write32()((CUstream)st, (CUdeviceptr)(s.gate.as<unsigned int>() + c), 1u, 0u) != CUDA_SUCCESS) {
    set_err(ctx, "cuStreamWriteValue32 failed");
    return MDEGGS_ECUDA;
  }
  return MDEGGS_OK;
}
int copy_chunks_end(mdeggs_ctx* ctx, Shard& s) {
  if (ctx->copy_streams >= 2) {
    CU(cudaEventRecord(s.ev_copy2_done, s.s_copy2));
    CU(cudaStreamWaitEvent(s.s_copy, s.ev_copy2_done, 0));
  }
  CU(cudaEventRecord(s.ev_copy_done, s.s_copy));
  return MDEGGS_OK;
}

// ---- host-side packing (mdeggs_hostpack.h) ----
// Decides whether this call runs on the packed problem and prepares it: node rows, renumbered edge list in pinned
// staging, the packing job.  The matrix itself is packed by run_host_pack.
int plan_host_pack(mdeggs_ctx* ctx, const mdeggs_problem* in) {
  mdeggs_ctx::HostPack& K = ctx->pack;
  K.active = false;
  K.deferred = false;
  K.rows_valid = false;
  if (!ctx->host_pack || in->mem != MDEGGS_MEM_HOST || in->layout != MDEGGS_LAYOUT_COLMAJOR || in->E <= 0 ||
      in->E > ((int64_t)1 << 20) || ctx->shards.size() != 1 || ctx->zero_copy)
    return MDEGGS_OK;
  const size_t min_bytes = ctx->host_pack >= 2 ? 0 : ((size_t)2 << 20);
  if ((size_t)in->G * (size_t)in->n * 8 < min_bytes) return MDEGGS_OK;  // small matrices: the plain copy is as good
  const size_t eb = (size_t)in->E * 4;
  if (!(K.seen_G == in->G && K.seen_from.size() == (size_t)in->E && memcmp(K.seen_from.data(), in->from, eb) == 0 &&
        memcmp(K.seen_to.data(), in->to, eb) == 0)) {
    K.idx_ok = pack_node_rows(in->from, in->to, in->E, in->G, K.rows, K.rank);
    K.seen_from.assign(in->from, in->from + in->E);
    K.seen_to.assign(in->to, in->to + in->E);
    K.seen_G = in->G;
  }
  K.rows_valid = true;
  const int64_t nn = (int64_t)K.rows.size();
  if (nn == 0 || (size_t)nn * (size_t)in->n * 8 < min_bytes) return MDEGGS_OK;
  const int64_t band = (int64_t)K.rows.back() - K.rows.front() + 1;
  if (ctx->host_pack < 2) {
    bool pageable = true;
    cudaPointerAttributes pa;
    if (cudaPointerGetAttributes(&pa, in->assay) == cudaSuccess) pageable = pa.type == cudaMemoryTypeUnregistered;
    else cudaGetLastError();
    // pinned and (nearly) contiguous node rows: the streamed band copy moves the same bytes without touching them
    if (!pageable && 4 * nn > 3 * band) return MDEGGS_OK;
  }
  const int64_t ld_p = (nn + 1) & ~(int64_t)1;
  const size_t mat_bytes = (size_t)ld_p * (size_t)in->n * 8;
  const size_t need = mat_bytes + 2 * (size_t)in->E * 4 + 64;
  if (need > K.h_cap) {
    if (K.h_buf) cudaFreeHost(K.h_buf);
    K.h_buf = nullptr;
    K.h_cap = 0;
    CU(cudaMallocHost(&K.h_buf, need + need / 8));
    K.h_cap = need + need / 8;
  }
  int32_t* hf = reinterpret_cast<int32_t*>(static_cast<char*>(K.h_buf) + mat_bytes);
  int32_t* ht = hf + in->E;
  const int32_t G = in->G;
  for (int64_t e = 0; e < in->E; ++e) {  // 1-based node ranks; an index outside 1..G stays invalid (0): the device reports it
    const int32_t a = in->from[e], b = in->to[e];
    hf[e] = (a >= 1 && a <= G) ? K.rank[(size_t)a - 1] + 1 : 0;
    ht[e] = (b >= 1 && b <= G) ? K.rank[(size_t)b - 1] + 1 : 0;
  }
  K.in2 = *in;
  K.in2.G = (int32_t)nn;
  K.in2.ld = ld_p;
  K.in2.assay = static_cast<const double*>(K.h_buf);
  K.in2.from = hf;
  K.in2.to = ht;
  K.job = PackJob();
  K.job.src = in->assay;
  K.job.ld_src = in->ld;
  K.job.dst = static_cast<double*>(K.h_buf);
  K.job.ld_dst = ld_p;
  K.job.rows = K.rows.data();
  K.job.nn = (int)nn;
  K.job.n_cols = in->n;
  K.job.contiguous = band == nn;
  K.active = true;
  return MDEGGS_OK;
}

// True when a kernel launch returns to the host before the kernel has run (probed once per context): only then may a
// launched kernel wait for copies the host enqueues AFTER the launch (streamed host packing under graph replay).
bool launches_are_async(mdeggs_ctx* ctx, Shard& s) {
  if (ctx->launch_async >= 0) return ctx->launch_async == 1;
  ctx->launch_async = 0;
  unsigned int* h = nullptr;
  if (cudaHostAlloc(reinterpret_cast<void**>(&h), 64, cudaHostAllocMapped) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  volatile unsigned int* flag = h;
  flag[0] = 0u;
  h[8] = 0u;
  unsigned int *d_flag = nullptr;
  if (cudaHostGetDevicePointer(reinterpret_cast<void**>(&d_flag), h, 0) == cudaSuccess) {
    k_async_probe<<<1, 1, 0, s.s_main>>>(d_flag, d_flag + 8);
    flag[0] = 1u;
    if (cudaStreamSynchronize(s.s_main) == cudaSuccess) ctx->launch_async = h[8] == 1u ? 1 : 0;
  }
  cudaGetLastError();
  cudaFreeHost(h);
  return ctx->launch_async == 1;
}

// Packs the node rows into the pinned staging with the pool.  streamed: chunk by chunk with the shard's gate geometry,
// every finished chunk is sent at once (the compute phase is already enqueued and waits on the gates); otherwise the
// staging is simply filled and the caller copies it.
int run_host_pack(mdeggs_ctx* ctx, Shard& s, bool streamed) {
  mdeggs_ctx::HostPack& K = ctx->pack;
  PackJob& J = K.job;
  J.chunk_rows = streamed ? s.gate_rows : (int64_t)J.nn;
  J.chunks = streamed ? s.gate_chunks : 1;
  J.units_per_chunk = (J.n_cols + PACK_COL_BLOCK - 1) / PACK_COL_BLOCK;
  HostPool& pool = HostPool::get();
  struct Guard {
    HostPool& p;
    ~Guard() { p.finish(); }
  };
  int rc = MDEGGS_OK;
  pool.start(J);  // (streamed: stage_inputs has closed the gates, copy_chunks_begin)
  Guard g{pool};
  for (int c = 0; c < J.chunks; ++c) {
    while (!pool.chunk_done(c))
      if (!pool.help()) _mm_pause();
    if (streamed && (rc = copy_chunk(ctx, s, c, J.dst, J.ld_dst, J.nn, J.n_cols))) return rc;
  }
  if (streamed) rc = copy_chunks_end(ctx, s);
  return rc;
}

// Phase A (always eager): caller inputs -> device, tiny per-call tables through pinned staging.
int stage_inputs(mdeggs_ctx* ctx, Shard& s, const mdeggs_problem* in, const HostLayout& H, int src_device) {
  const int G = in->G, n = in->n;
  const int64_t E = in->E;
  CU(cudaSetDevice(s.device));
  CU(cudaEventRecord(s.ev[EV_START], s.s_main));
  s.gated = false;
  // ---- inputs ----
  const size_t assay_elems = (in->layout == MDEGGS_LAYOUT_COLMAJOR) ? (size_t)in->ld * (size_t)(n - 1) + (size_t)G
                                                                   : (size_t)in->ld * (size_t)(G - 1) + (size_t)n;
  const double* d_assay;
  const int32_t *d_from, *d_to;
  bool zc = false;
  if (in->mem == MDEGGS_MEM_HOST && ctx->zero_copy) {  // is the assay in pinned (device-mapped) host memory?
    cudaPointerAttributes pa;
    if (cudaPointerGetAttributes(&pa, in->assay) == cudaSuccess && pa.type == cudaMemoryTypeHost && pa.devicePointer)
      zc = true;
    else
      cudaGetLastError();
  }
  ctx->last_zc = zc;
  if (zc) {
    // zero copy: the gather (and the quantile sample) read the caller's pinned matrix over PCIe themselves; only the
    // node rows are touched, the transfer overlaps the transposition and no staging copy exists
    s.row_off = 0;
    s.ld_eff = in->ld;
    s.rows_avail = G;
    CU(s.from.ensure((size_t)E * 4 + 4));
    CU(s.to.ensure((size_t)E * 4 + 4));
    if (E > 0) {
      CU(cudaMemcpyAsync(s.from.p, in->from, (size_t)E * 4, cudaMemcpyHostToDevice, s.s_main));
      CU(cudaMemcpyAsync(s.to.p, in->to, (size_t)E * 4, cudaMemcpyHostToDevice, s.s_main));
    }
    ctx->stats.h2d_bytes += (int64_t)((size_t)E * 8);
    d_assay = in->assay;
    d_from = s.from.as<int32_t>();
    d_to = s.to.as<int32_t>();
    s.assay.release();
    s.assay.p = (void*)d_assay;  // borrowed (cap stays 0)
  } else if (in->mem == MDEGGS_MEM_HOST) {
    // Only rows that are network nodes are ever read (core_functions.R:336). For small edge lists the host finds
    // the node-row RANGE [r_lo, r_hi] in one pass over from/to and only that band crosses PCIe (config 3: the
    // 8,405 node genes are the first rows of 20,000 -> 67 MB instead of 160 MB).
    int64_t r_lo = 0, r_hi = G - 1;
    if (E > 0 && E <= (int64_t)1 << 20 && ctx->pack.rows_valid && !ctx->pack.active) {
      // (plan_host_pack has just analysed this very edge list: the band is its first and last node row)
      if (ctx->pack.idx_ok && !ctx->pack.rows.empty()) {
        r_lo = (int64_t)ctx->pack.rows.front() & ~(int64_t)15;
        r_hi = ctx->pack.rows.back();
      }
    } else if (E > 0 && E <= (int64_t)1 << 20) {
      int32_t lo = INT32_MAX, hi = 0;
      for (const int32_t* v : {in->from, in->to}) {  // two plain min/max reductions: the host compiler vectorises them
        int32_t l = INT32_MAX, h = 0;
        for (int64_t e = 0; e < E; ++e) {
          l = v[e] < l ? v[e] : l;
          h = v[e] > h ? v[e] : h;
        }
        lo = l < lo ? l : lo;
        hi = h > hi ? h : hi;
      }
      if (lo >= 1 && hi <= G) {  // out-of-range indices are reported by the device check below
        r_lo = (lo - 1) & ~(int64_t)15;  // the fused front kernel walks 16-row blocks of absolute rows
        r_hi = hi - 1;
      }
    }
    const int64_t band = r_hi - r_lo + 1;
    s.row_off = r_lo;
    s.rows_avail = band;
    CU(s.from.ensure((size_t)E * 4 + 4));
    CU(s.to.ensure((size_t)E * 4 + 4));
    size_t copied;
    s.gated = false;
    if (E > 0) {  // (the edge list first: K0 only needs it, and it is what a streamed matrix copy is hidden behind)
      CU(cudaMemcpyAsync(s.from.p, in->from, (size_t)E * 4, cudaMemcpyHostToDevice, s.s_main));
      CU(cudaMemcpyAsync(s.to.p, in->to, (size_t)E * 4, cudaMemcpyHostToDevice, s.s_main));
    }
    if (in->layout == MDEGGS_LAYOUT_COLMAJOR) {
      s.ld_eff = (band + 1) & ~(int64_t)1;  // even: a TMA tensor map needs 16-byte strides
      CU(s.assay.ensure((size_t)s.ld_eff * (size_t)n * 8));
      copied = (size_t)band * n * 8;
      // Streamed copy: when the TMA-fed front kernel serves the call, the band crosses PCIe in row chunks on s_copy and
      // the compute phase starts at once: K0 and the graph launch hide behind the first chunk, the front kernel works
      // on chunk c while chunk c + 1 is in flight (its loader waits on gate[c], mdeggs_kernels.cuh: gate_wait).
      int64_t chunk_rows = 0;
      int chunks = 0;
      if (ctx->h2d_pipeline && E > 0 && stream_write32()) {
        // (a packed matrix is produced by the host at about the speed of the link: finer chunks, earlier first DMA)
        const size_t chunk_bytes = (size_t)(ctx->h2d_pipeline > 1 ? ctx->h2d_pipeline : (ctx->pack.active ? 4 : 12)) << 20;
        int64_t c = (int64_t)(copied / chunk_bytes);
        c = c > GATE_MAX_CHUNKS ? GATE_MAX_CHUNKS : c;
        if (c >= 2) {
          chunk_rows = (((band + c - 1) / c) + 15) & ~(int64_t)15;
          chunks = (int)((band + chunk_rows - 1) / chunk_rows);
          FrontPlan FP;
          bool tma = false, ok = false;
          const int64_t cap = 2 * E < (int64_t)G ? 2 * E : (int64_t)G;
          int rc = front_prepare(ctx, s, in, H.L, (int32_t)cap, &FP, &tma, &ok);
          if (rc) return rc;
          if (!ok || !tma || chunks < 2) chunks = 0;
        }
      }
      if (ctx->pack.active && chunks < 2) {  // packed, but not streamed: fill the staging now, the plain copy moves it
        int rc = run_host_pack(ctx, s, false);
        if (rc) return rc;
      }
      if (chunks >= 2) {
        s.gated = true;
        s.gate_rows = chunk_rows;
        s.gate_chunks = chunks;
        int rc = copy_chunks_begin(ctx, s);
        if (rc) return rc;
        if (ctx->pack.active) {
          ctx->pack.deferred = true;  // packed + sent chunk by chunk once the compute phase is enqueued (submit)
        } else {
          for (int c = 0; c < chunks && !rc; ++c) rc = copy_chunk(ctx, s, c, in->assay + r_lo, in->ld, band, n);
          if (!rc) rc = copy_chunks_end(ctx, s);
          if (rc) return rc;
        }
      } else if (band == G && in->ld == G && s.ld_eff == G) {
        CU(cudaMemcpyAsync(s.assay.p, in->assay, (size_t)G * n * 8, cudaMemcpyHostToDevice, s.s_main));
      } else {
        CU(cudaMemcpy2DAsync(s.assay.p, (size_t)s.ld_eff * 8, in->assay + r_lo, (size_t)in->ld * 8, (size_t)band * 8,
                             (size_t)n, cudaMemcpyHostToDevice, s.s_main));
      }
    } else {
      s.ld_eff = in->ld;
      copied = ((size_t)in->ld * (size_t)(band - 1) + (size_t)n) * 8;
      CU(s.assay.ensure(copied));
      CU(cudaMemcpyAsync(s.assay.p, in->assay + (size_t)r_lo * in->ld, copied, cudaMemcpyHostToDevice, s.s_main));
    }
    ctx->stats.h2d_bytes += (int64_t)(copied + (size_t)E * 8);
    d_assay = s.assay.as<double>();
    d_from = s.from.as<int32_t>();
    d_to = s.to.as<int32_t>();
  } else if (s.device == src_device) {
    s.row_off = 0;
    s.ld_eff = in->ld;
    s.rows_avail = G;
    d_assay = in->assay;
    d_from = in->from;
    d_to = in->to;
  } else {
    s.row_off = 0;
    s.ld_eff = in->ld;
    s.rows_avail = G;
    CU(s.assay.ensure(assay_elems * 8));
    CU(s.from.ensure((size_t)E * 4 + 4));
    CU(s.to.ensure((size_t)E * 4 + 4));
    CU(cudaMemcpyPeerAsync(s.assay.p, s.device, in->assay, src_device, assay_elems * 8, s.s_main));
    if (E > 0) {
      CU(cudaMemcpyPeerAsync(s.from.p, s.device, in->from, src_device, (size_t)E * 4, s.s_main));
      CU(cudaMemcpyPeerAsync(s.to.p, s.device, in->to, src_device, (size_t)E * 4, s.s_main));
    }
    d_assay = s.assay.as<double>();
    d_from = s.from.as<int32_t>();
    d_to = s.to.as<int32_t>();
  }
  // borrowed device pointers are stashed in the DevBuf slots without ownership for later stages
  if (in->mem == MDEGGS_MEM_DEVICE && s.device == src_device) {
    s.assay.release();
    s.from.release();
    s.to.release();
    s.assay.p = (void*)d_assay;
    s.from.p = (void*)d_from;
    s.to.p = (void*)d_to;  // cap stays 0 => never freed by us
  }
  // src_of_pos and the percentile vector travel as ONE small block, and only when they differ from what the device
  // already holds (repeated calls with the same sample layout: nestedcv folds of equal size, benchmarks)
  const size_t sop_bytes = H.src_of_pos.size() * 4, pct_bytes = (size_t)(in->P > 0 ? in->P : 0) * 8;
  const size_t pct_off = (sop_bytes + 15) & ~(size_t)15;
  const size_t blk = pct_off + (pct_bytes ? pct_bytes : 8);
  CU(s.ensure_stage(blk + small_block_bytes(in->P) + 128));  // inputs at the front, packed small results at the back
  CU(s.stage_dev.ensure(blk));
  std::vector<char> cur(blk, 0);
  memcpy(cur.data(), H.src_of_pos.data(), sop_bytes);
  if (pct_bytes) memcpy(cur.data() + pct_off, in->pct, pct_bytes);
  if (s.stage_shadow_dev != s.stage_dev.p || cur != s.stage_shadow) {
    memcpy(s.h_stage, cur.data(), blk);
    CU(cudaMemcpyAsync(s.stage_dev.p, s.h_stage, blk, cudaMemcpyHostToDevice, s.s_main));
    s.stage_shadow.swap(cur);
    s.stage_shadow_dev = s.stage_dev.p;
  }
  s.src_of_pos.p = s.stage_dev.p;  // borrowed views (cap == 0: never freed through these slots)
  s.pct.p = s.stage_dev.as<char>() + pct_off;
  CU(cudaEventRecord(s.ev[EV_H2D], s.s_main));
  trace_mark(ctx, "(inputs staged)", s.s_main);
  return MDEGGS_OK;
}

// Phase B starts here (capturable): K0 node filter
int run_shard_front(mdeggs_ctx* ctx, Shard& s, const mdeggs_problem* in, int32_t* n_nodes_host) {
  const int G = in->G;
  const int64_t E = in->E;
  CU(cudaSetDevice(s.device));
  const int32_t* d_from = s.from.as<int32_t>();
  const int32_t* d_to = s.to.as<int32_t>();
  if (getenv("MDEGGS_HOSTTIME")) {  // diagnostics: GPU-side time stamps around the compute phase
    CU(s.stamps.ensure(64));
    LAUNCH(k_stamp, 1, 1, 0, s.s_main, s.stamps.as<unsigned long long>());
  }
  // K3 S0 on the side stream, beside K0: sample -> knots of the quantile bins (needs only the inputs)
  if (E > 0 && in->P > 0) {
    const int n = in->n;
    if ((size_t)QS_BINS * 4 > s.rsel_hist.cap) {  // the scan kernel leaves the histogram zeroed for the next call
      CU(s.rsel_hist.ensure((size_t)QS_BINS * 4));
      CU(cudaMemsetAsync(s.rsel_hist.p, 0, (size_t)QS_BINS * 4, s.s_main));
    }
    if (sizeof(QsState) > s.rsel_state.cap) {  // its "CTAs done" counters are zero between calls
      CU(s.rsel_state.ensure(sizeof(QsState)));
      CU(cudaMemsetAsync(s.rsel_state.p, 0, sizeof(QsState), s.s_main));
    }
    CU(s.rsel_cand.ensure((size_t)qs_cand_capacity(ctx, in) * 8));
    if (!s.rsel_attr_set) {
      CU(cudaFuncSetAttribute(k_qs_finish, cudaFuncAttributeMaxDynamicSharedMemorySize, QS_FIN_CAP * 8));
      s.rsel_attr_set = true;
    }
    CU(cudaEventRecord(s.ev_fork, s.s_main));
    CU(cudaStreamWaitEvent(s.s_sort, s.ev_fork, 0));
    const int S = QS_SAMPLE;
    LAUNCH(k_qs_sample, 1, QS_SAMPLE_THREADS, 0, s.s_sort, s.assay.as<double>(), s.ld_eff, s.row_off,
           in->layout == MDEGGS_LAYOUT_COLMAJOR ? 1 : 0, d_from, d_to, E, G, n, S, s.rsel_state.as<QsState>(),
           s.gated ? s.gate.as<unsigned int>() : (const unsigned int*)nullptr, s.gate_rows,
           (int32_t*)nullptr);  // (K0's memset of the error flag runs beside this kernel: the front kernel reports)
    CU(cudaEventRecord(s.ev_sample, s.s_sort));
  }
  // ---- K0 node filter ----
  // one memset clears the scalars (node count, error flag, "done" counter) and the row bitmap behind them
  const int W = (G + 31) >> 5;
  CU(s.scal.ensure(SC_COUNT * 8 + (size_t)W * 4));
  CU(s.excl.ensure((size_t)W * 4));
  CU(s.node_of_row.ensure((size_t)G * 4));
  CU(s.rowlist.ensure((size_t)G * 4));
  CU(s.ea.ensure((size_t)(E > 0 ? E : 1) * 4));
  CU(s.eb.ensure((size_t)(E > 0 ? E : 1) * 4));
  CU(cudaMemsetAsync(s.scal.p, 0, SC_COUNT * 8 + (size_t)W * 4, s.s_main));
  unsigned int* d_bitmap = reinterpret_cast<unsigned int*>(s.scal.as<char>() + SC_COUNT * 8);
  {
    // each CTA flushes its shared-memory bitmap (W atomicOr at most): few, fat CTAs
    int64_t nb = (E + 1023) / 1024;
    unsigned grid = (unsigned)(nb < 1 ? 1 : (nb < 148 * 4 ? nb : 148 * 4));
    const size_t smem = G <= K0_SMEM_ROWS ? (size_t)W * 4 : 0;
    LAUNCH(k_mark_nodes, grid, 256, smem, s.s_main, d_from, d_to, E, G, d_bitmap, s.excl.as<int32_t>(),
           s.scal.as<int32_t>() + 2 * SC_NNODES, s.scal.as<unsigned int>() + 2 * SC_DONE,
           s.scal.as<int32_t>() + 2 * SC_ERR);
    const int64_t nt = (int64_t)G > E ? (int64_t)G : E;
    LAUNCH(k_build_nodes, blocks_for(nt, 256), 256, 0, s.s_main, d_bitmap, s.excl.as<int32_t>(), G, d_from, d_to, E,
           s.node_of_row.as<int32_t>(), s.rowlist.as<int32_t>(), s.ea.as<int32_t>(), s.eb.as<int32_t>(),
           s.scal.as<int32_t>() + 2 * SC_ERR);
  }
  // No mid-pipeline host read: buffers and grids are sized for the upper bound min(G, 2E) of the node count,
  // kernels read the real count from device memory, and an out-of-range edge index raises a device flag that
  // makes every later edge kernel exit at once; the flag and the count are read after the final sync.
  int64_t cap = 2 * E < (int64_t)G ? 2 * E : (int64_t)G;
  *n_nodes_host = (int32_t)cap;
  return MDEGGS_OK;
}

// ---- K1 + K2 fused (mdeggs_front.cuh) --------------------------------------------------------------
typedef CUresult (*tmap_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
tmap_encode_fn tmap_encoder() {  // the driver entry point is resolved through the runtime: no link against libcuda
  static tmap_encode_fn fn = []() -> tmap_encode_fn {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess) {
      cudaGetLastError();
      return nullptr;
    }
    return reinterpret_cast<tmap_encode_fn>(p);
  }();
  return fn;
}

template <int GB, int ITEMS, int NW>
int launch_front_inst(mdeggs_ctx* ctx, Shard& s, const FrontArgs& A, const FrontPlan& P, unsigned grid, int inst) {
  if (!(s.front_attr_mask & (1u << inst))) {
    CU(cudaFuncSetAttribute(k_front_fused<GB, ITEMS, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, FR_SMEM_MAX - 2048));
    s.front_attr_mask |= 1u << inst;
  }
  LAUNCH((k_front_fused<GB, ITEMS, NW>), grid, (NW + FR_PROD_WARPS) * 32, P.smem, s.s_main, s.tmap, A);
  return MDEGGS_OK;
}

// Shape of the fused front kernel for this call and, when the device copy of the matrix can be described by one, its
// tensor map (cached in the shard).  *ok = false: the separate K1 + K2 kernels serve the call.  Called by stage_inputs
// (the streamed input copy needs to know whether the TMA-fed kernel will run) and by launch_front.
int front_prepare(mdeggs_ctx* ctx, Shard& s, const mdeggs_problem* in, const SegLayout& L, int32_t n_nodes, FrontPlan* Pout,
                  bool* use_tma_out, bool* ok) {
  *ok = false;
  *use_tma_out = false;
  if (ctx->front_path == 0 || ctx->last_zc || n_nodes <= 0) return MDEGGS_OK;
  int max_cnt = 0;
  for (int k = 0; k < L.K; ++k) max_cnt = L.cnt[k] > max_cnt ? L.cnt[k] : max_cnt;
  const FrontPlan P = front_plan(L.n, L.K, max_cnt, 256);
  if (!P.ok) return MDEGGS_OK;
  const bool colmajor = in->layout == MDEGGS_LAYOUT_COLMAJOR;
  bool use_tma = ctx->front_path != 2 && colmajor && (s.ld_eff % 2 == 0) && ((uintptr_t)s.assay.p % 16 == 0) &&
                 tmap_encoder() != nullptr;
  if (use_tma) {
    Shard::TmapKey key = {s.assay.p, s.ld_eff, s.rows_avail, L.n, P.GB, P.box1};
    if (memcmp(&key, &s.tmap_key, sizeof key) != 0) {
      const cuuint64_t gdim[2] = {(cuuint64_t)s.rows_avail, (cuuint64_t)L.n};
      const cuuint64_t gstr[1] = {(cuuint64_t)s.ld_eff * 8};
      const cuuint32_t box[2] = {(cuuint32_t)P.GB, (cuuint32_t)P.box1};
      const cuuint32_t estr[2] = {1, 1};
      const CUtensorMapSwizzle sw = P.GB == 16 ? CU_TENSOR_MAP_SWIZZLE_128B
                                               : (P.GB == 8 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
      const CUresult r = tmap_encoder()(&s.tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, s.assay.p, gdim, gstr, box, estr,
                                        CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) {
        use_tma = false;  // (shape outside what a tensor map can describe: the plain loader serves it)
        memset(&s.tmap_key, 0, sizeof s.tmap_key);
      } else {
        s.tmap_key = key;
      }
    }
  }
  *Pout = P;
  *use_tma_out = use_tma;
  *ok = true;
  return MDEGGS_OK;
}

// returns MDEGGS_OK and *done = true when the fused kernel was launched; *done = false: use the separate K1 + K2
int launch_front(mdeggs_ctx* ctx, Shard& s, const mdeggs_problem* in, const SegLayout& L, int32_t n_nodes, bool do_q,
                 unsigned int cand_cap, bool* done) {
  *done = false;
  FrontPlan P{};
  bool use_tma = false, ok = false;
  {
    int rc = front_prepare(ctx, s, in, L, n_nodes, &P, &use_tma, &ok);
    if (rc) return rc;
  }
  if (!ok) return MDEGGS_OK;
  const bool colmajor = in->layout == MDEGGS_LAYOUT_COLMAJOR;
  FrontArgs A;
  memset(&A, 0, sizeof A);
  A.assay = s.assay.as<double>();
  A.ld = s.ld_eff;
  A.row_off = s.row_off;
  A.rows_avail = s.rows_avail;
  A.colmajor = colmajor ? 1 : 0;
  A.use_tma = use_tma ? 1 : 0;
  A.bitmap = reinterpret_cast<const unsigned int*>(s.scal.as<char>() + SC_COUNT * 8);
  A.word_base = s.excl.as<int32_t>();
  A.src_of_pos = s.src_of_pos.as<int32_t>();
  A.L = L;
  A.D = s.D.as<double>();
  A.B16 = do_q ? s.B16.as<unsigned short>() : nullptr;
  A.mean = s.mean.as<double>();
  A.sxx = s.sxx.as<double>();
  A.med = s.med.as<double>();
  A.bad = s.bad.as<uint8_t>();
  A.st = do_q ? s.rsel_state.as<QsState>() : nullptr;
  A.hist = do_q ? s.rsel_hist.as<unsigned int>() : nullptr;
  A.pct = s.pct.as<double>();
  A.P = in->P;
  A.cand_capacity = cand_cap;
  A.tile_lo = (int)(s.row_off / P.GB);  // (row_off is a multiple of 16)
  A.tile_hi = (int)((s.row_off + s.rows_avail + P.GB - 1) / P.GB);
  const int64_t g_tiles = ((int64_t)in->G + P.GB - 1) / P.GB;
  if (A.tile_hi > g_tiles) A.tile_hi = (int)g_tiles;
  A.nbox = P.nbox;
  A.box1 = P.box1;
  A.stages = P.stages;
  A.stage_bytes = P.stage_bytes;
  if (s.gated && use_tma) {  // (stage_inputs only streams the copy when this kernel, TMA-fed, serves the call)
    A.gate = s.gate.as<unsigned int>();
    A.gate_rows = (int)s.gate_rows;
    A.gate_chunks = s.gate_chunks;
    A.err = s.scal.as<int32_t>() + 2 * SC_ERR;
  }
  const int n_tiles = A.tile_hi - A.tile_lo;
  if (n_tiles <= 0) return MDEGGS_OK;
  const unsigned grid = (unsigned)(n_tiles < s.sm_count ? n_tiles : s.sm_count);
  CU(cudaMemsetAsync(s.bad.p, 0, (size_t)n_nodes, s.s_main));
  int rc = MDEGGS_OK;
#define MDEGGS_FR(GB_, IT_, NW_, ID_) rc = launch_front_inst<GB_, IT_, NW_>(ctx, s, A, P, grid, ID_)
#define MDEGGS_FR_GB(IT_, NW_, ID0_)                       \
  do {                                                     \
    if (P.GB == 16) MDEGGS_FR(16, IT_, NW_, ID0_);         \
    else if (P.GB == 8) MDEGGS_FR(8, IT_, NW_, ID0_ + 1);  \
    else MDEGGS_FR(4, IT_, NW_, ID0_ + 2);                 \
  } while (0)
  switch (P.items) {
    case 4: MDEGGS_FR_GB(4, 24, 0); break;
    case 12: MDEGGS_FR_GB(12, 24, 3); break;
    case 20: MDEGGS_FR_GB(20, 16, 6); break;
    default: MDEGGS_FR_GB(32, 16, 9); break;
  }
#undef MDEGGS_FR_GB
#undef MDEGGS_FR
  if (rc) return rc;
  s.front_bits = use_tma ? 0x400 : 0x800;  // fit_path bit 10: fused front kernel fed by TMA, bit 11: by the plain loader
  if (A.gate) s.front_bits |= 0x1000;      // bit 12: ... while the input copy was still streaming in (stage_inputs)
  *done = true;
  return MDEGGS_OK;
}

int run_shard_main(mdeggs_ctx* ctx, Shard& s, const mdeggs_problem* in, const mdeggs_result* out, const HostLayout& H,
                   int32_t n_nodes) {
  const SegLayout& L = H.L;
  const int n = in->n, K = in->K, P = in->P;
  const int64_t E = in->E;
  CU(cudaSetDevice(s.device));
  const int32_t* d_nn = s.scal.as<int32_t>() + 2 * SC_NNODES;
  const size_t nn = (size_t)(n_nodes > 0 ? n_nodes : 1);
  // ---- K1 gather ----
  CU(s.D.ensure(nn * (size_t)L.npad * 8));
  if (P > 0) CU(s.B16.ensure(nn * (size_t)L.npad * 2 + 16));
  CU(s.mean.ensure(nn * K * 8));
  CU(s.sxx.ensure(nn * K * 8));
  CU(s.med.ensure(nn * K * 8));
  CU(s.bad.ensure(nn));
  CU(s.cutoff.ensure((size_t)(P > 0 ? P : 1) * 8));
  // K3 (exact quantiles of the whole node-filtered matrix, mdeggs_qsel.cuh) is threaded through the pipeline:
  //   S0 sample -> knots (one CTA) ; S1 histogram fused into the gather ; S2 scan (one CTA) ; S3 compaction (the
  //   only extra read of D) ; S4 finish on the side stream.  Full-grid kernels are NOT overlapped with each other:
  //   measured, two of them side by side only slow each other down (resident CTAs = loads in flight).
  const bool do_q = n_nodes > 0 && P > 0;  // == the condition under which run_shard_front launched S0
  QsState* d_st = do_q ? s.rsel_state.as<QsState>() : nullptr;
  const unsigned int cand_cap = do_q ? qs_cand_capacity(ctx, in) : 0;
  if (do_q) CU(cudaStreamWaitEvent(s.s_main, s.ev_sample, 0));  // knots ready
  // K1 + K2 as one TMA-fed kernel whenever a tile of the matrix fits the staging ring (mdeggs_front.cuh)
  bool fused = false;
  s.front_bits = 0;
  {
    int rc = launch_front(ctx, s, in, L, n_nodes, do_q, cand_cap, &fused);
    if (rc) return rc;
  }
  if (n_nodes > 0 && !fused) {
    unsigned int* d_hist = do_q ? s.rsel_hist.as<unsigned int>() : nullptr;
    const unsigned grid = 148 * 6;  // persistent CTAs (6 fit an SM's shared memory): one histogram flush each
    if (in->layout == MDEGGS_LAYOUT_COLMAJOR) {
      LAUNCH(k_gather_colmajor, grid, 256, 0, s.s_main, s.assay.as<double>(), s.ld_eff, s.row_off, s.rowlist.as<int32_t>(), d_nn,
             s.src_of_pos.as<int32_t>(), L.npad, n, s.D.as<double>(), d_st, d_hist, s.B16.as<unsigned short>(),
             s.pct.as<double>(), P, cand_cap);
    } else {
      LAUNCH(k_gather_rowmajor, grid, 256, 0, s.s_main, s.assay.as<double>(), s.ld_eff, s.row_off, s.rowlist.as<int32_t>(), d_nn,
             s.src_of_pos.as<int32_t>(), L.npad, n, s.D.as<double>(), d_st, d_hist, s.B16.as<unsigned short>(),
             s.pct.as<double>(), P, cand_cap);
    }
  }
  EVREC(s.ev[EV_GATHER], s.s_main);
  EVREC(s.ev[EV_Q0], s.s_main);
  // ---- per-run tail tables (depend only on the degrees of freedom) ----
  {
    const bool rlm2 = (K == 2 && in->method == MDEGGS_METHOD_RLM);
    int Ke = 0;
    for (int k = 0; k < K; ++k) Ke += L.cnt[k] > 0;
    const double tdf1 = K == 2 ? 1.0 : (double)(Ke - 1), tdf2 = K == 2 ? (double)(n - 4) : (double)(n - 2 * Ke);
    const double ta = (K == 2 && !rlm2) ? 0.5 : tdf1 / 2.0;
    CU(s.cf_tab.ensure((size_t)CF_TABLE_DOUBLES * 8));
    // a one-CTA kernel nothing depends on until the tails: side stream (idle between the quantile sample and finish)
    CU(cudaEventRecord(s.ev_fork, s.s_main));
    CU(cudaStreamWaitEvent(s.s_sort, s.ev_fork, 0));
    LAUNCH(k_cf_tables, 1, 256, 0, s.s_sort, ta, tdf2 / 2.0, s.cf_tab.as<double>(), s.cf_tab.as<double>() + CF_TABLE_LEN,
           s.cf_tab.as<double>() + 2 * CF_TABLE_LEN, s.cf_tab.as<double>() + 2 * CF_TABLE_LEN + 2,
           s.cf_tab.as<double>() + 3 * CF_TABLE_LEN + 2);
    // two-category lm on a long edge list: the t tail of the run as one interpolation table (mdeggs_math.cuh)
    const int64_t e_rank = (E + s.world - 1) / s.world;
    s.tail_sp_on = K == 2 && !rlm2 && tdf2 >= 1.0 && ctx->tail_spline &&
                   (ctx->tail_spline >= 2 || e_rank >= TAIL_SPLINE_MIN_EDGES);
    if (s.tail_sp_on) {
      CU(s.tail_sp.ensure((size_t)TS_M * TS_C * 8));
      TailConsts tc;
      memset(&tc, 0, sizeof tc);
      tc.a = 0.5;
      tc.b = tdf2 / 2.0;
      tc.e_ab = s.cf_tab.as<double>();
      tc.e_ba = s.cf_tab.as<double>() + CF_TABLE_LEN;
      tc.c_ab = s.cf_tab.as<double>() + 2 * CF_TABLE_LEN + 2;
      tc.c_ba = s.cf_tab.as<double>() + 3 * CF_TABLE_LEN + 2;
      LAUNCH(k_tail_spline, TS_M * TS_C / 256, 256, 0, s.s_sort, tc, s.cf_tab.as<double>() + 2 * CF_TABLE_LEN, tdf2,
             tail_spline_smax(tdf2) / (double)TS_M, s.tail_sp.as<double>());
    }
    CU(cudaEventRecord(s.ev_cf, s.s_sort));
  }
  // ---- K2 gene stats (+ K3 S3: the values of the quantile slots are appended to their lists here) ----
  QsCompactArgs QC;
  memset(&QC, 0, sizeof QC);
  if (do_q) {
    QC.B16 = s.B16.as<unsigned short>();
    QC.U = &d_st->U;
    QC.slot_bin = d_st->slot_bin;
    QC.slot_fat = d_st->slot_fat;
    QC.slot_count = d_st->slot_count;
    QC.slot_off = d_st->slot_off;
    QC.slot_cursor = d_st->slot_cursor;
    QC.slot_min = d_st->slot_min;
    QC.slot_max = d_st->slot_max;
    QC.cand = s.rsel_cand.as<unsigned long long>();
  }
  if (n_nodes > 0 && !fused) {
    int max_cnt = 0;
    for (int k = 0; k < K; ++k) max_cnt = L.cnt[k] > max_cnt ? L.cnt[k] : max_cnt;
    if (max_cnt <= 32 * 32) {  // segments fit in registers: one warp per (gene, category)
      CU(cudaMemsetAsync(s.bad.p, 0, nn, s.s_main));
      const unsigned grid = (unsigned)(((long long)n_nodes * K + 7) / 8);
#define MDEGGS_GS(IT)                                                                                          \
  LAUNCH(k_gene_stats_reg<IT>, grid, 256, 0, s.s_main, s.D.as<double>(), L, d_nn, s.mean.as<double>(),        \
         s.sxx.as<double>(), s.med.as<double>(), s.bad.as<uint8_t>(), QC)
      if (max_cnt <= 32 * 2) MDEGGS_GS(2);
      else if (max_cnt <= 32 * 4) MDEGGS_GS(4);
      else if (max_cnt <= 32 * 8) MDEGGS_GS(8);
      else if (max_cnt <= 32 * 12) MDEGGS_GS(12);
      else if (max_cnt <= 32 * 16) MDEGGS_GS(16);
      else if (max_cnt <= 32 * 24) MDEGGS_GS(24);
      else MDEGGS_GS(32);
#undef MDEGGS_GS
    } else {
      size_t smem = (size_t)L.npad * 8;
      int use_smem = smem <= 200 * 1024;
      if (use_smem && smem > 48 * 1024 && !s.smem_attr_set) {
        CU(cudaFuncSetAttribute(k_gene_stats<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        s.smem_attr_set = true;
      }
      LAUNCH(k_gene_stats<128>, (unsigned)n_nodes, 128, use_smem ? smem : 0, s.s_main, s.D.as<double>(), L, d_nn,
             s.mean.as<double>(), s.sxx.as<double>(), s.med.as<double>(), s.bad.as<uint8_t>(), use_smem, QC);
    }
  }
  if (do_q && fused) {
    // K3 S3 of the fused front: slot candidates from the bin ids (the separate K2 collects them itself).  On the MAIN
    // stream, before the pair fits: beside them it takes SM slots from the fit kernel (measured: fits 61 -> 90 us) and
    // lengthens the quantile chain that K7 waits for.
    const int64_t nvec = ((int64_t)n_nodes * L.npad + 7) / 8;
    const int64_t nb = (nvec + 4 * QSC_THREADS - 1) / (4 * QSC_THREADS);  // 4 vectors per thread and round
    LAUNCH(k_qs_compact, (unsigned)(nb < s.sm_count ? nb : s.sm_count), QSC_THREADS, 0, s.s_main, s.D.as<double>(),
           s.B16.as<unsigned short>(), L.npad, d_nn, QC);
  }
  EVREC(s.ev[EV_STATS], s.s_main);
  // S4 (one small CTA per slot) runs on the side stream next to the pair fits; K7 joins it
  CU(cudaEventRecord(s.ev_fork, s.s_main));
  CU(cudaStreamWaitEvent(s.s_sort, s.ev_fork, 0));
  if (do_q) {
    LAUNCH(k_qs_finish, (unsigned)(2 * P), QS_FIN_THREADS, (size_t)QS_FIN_CAP * 8, s.s_sort, s.D.as<double>(),
           s.B16.as<unsigned short>(), L.npad, d_nn, d_st, s.rsel_cand.as<unsigned long long>(), s.pct.as<double>(), P,
           s.cutoff.as<double>());
  } else if (P > 0) {
    LAUNCH(k_fill_f64, 1, 64, 0, s.s_sort, s.cutoff.as<double>(), (int64_t)P, nan(""));
  }
  // K7 activity masks need the medians (K2, this stream) and the cut-offs (S4, side stream): they are computed on
  // the side stream while the pair fits run
  CU(s.act.ensure((size_t)(P > 0 ? P : 1) * (n_nodes > 0 ? n_nodes : 1)));
  s.act_nch = P > 0 ? (P + 15) / 16 : 1;
  CU(s.act_t.ensure((size_t)s.act_nch * (n_nodes > 0 ? n_nodes : 1) * 16));
  s.order_tested_only = false;
  if (n_nodes > 0 && P > 0) {
    CU(cudaEventRecord(s.ev_sample, s.s_main));  // (event reused: "medians ready")
    CU(cudaStreamWaitEvent(s.s_sort, s.ev_sample, 0));
    LAUNCH(k_actmask, blocks_for((int64_t)s.act_nch * n_nodes, 256), 256, 0, s.s_sort, s.med.as<double>(),
           s.cutoff.as<double>(), d_nn, P, K, n_nodes, s.act.as<uint8_t>(), s.act_t.as<uint4>(), s.act_nch);
    s.order_tested_only = ctx->order_filter != 0;
  }
  EVREC(s.ev[EV_Q1], s.s_sort);
  CU(cudaEventRecord(s.ev_join, s.s_sort));
  // ---- pair fits on this rank's contiguous edge block ----
  const int64_t chunk = (E + s.world - 1) / s.world;
  const int64_t e_lo = (int64_t)s.rank * chunk < E ? (int64_t)s.rank * chunk : E;
  const int64_t e_hi = e_lo + chunk < E ? e_lo + chunk : E;
  const size_t epad = (size_t)(chunk * s.world > 0 ? chunk * s.world : 1);
  const bool want_coef = out->coef != nullptr;
  const bool is_rlm = (K == 2 && in->method == MDEGGS_METHOD_RLM);
  CU(s.stat.ensure(epad * 8));
  CU(s.status.ensure(epad * 4));
  CU(s.p_edge.ensure(epad * 8));
  if (want_coef) CU(s.coef.ensure(epad * 2 * K * 8));
  if (is_rlm) CU(s.iters.ensure(epad * 4));
  if (!is_rlm) CU(s.sxy.ensure(epad * K * 8));
  double* d_coef = want_coef ? s.coef.as<double>() : nullptr;
  bool use_dense = false;
  if (!is_rlm && ctx->fit_path_override != 0) {
    use_dense = ctx->fit_path_override == 1;
    const size_t need = (size_t)K * (size_t)n_nodes * (size_t)n_nodes * 8;
    if (ctx->fit_path_override < 0 && dense_path_candidate(E, n_nodes)) {
      use_dense = need <= s.dense_S.cap;
      if (!use_dense) {  // cudaMemGetInfo costs ~1 ms: only asked when the moment matrices must be (re)allocated
        size_t free_b = 0, total_b = 0;
        CU(cudaMemGetInfo(&free_b, &total_b));
        use_dense = need < (free_b + s.dense_S.cap) / 2;
      }
    }
    if (use_dense) CU(s.dense_S.ensure(need));
  }
  ctx->stats.fit_path = (is_rlm ? 2 : (use_dense ? 1 : 0)) | s.front_bits;
  if (e_hi > e_lo) {
    if (is_rlm) {
      int rc = launch_fit_rlm(ctx->launches, s.s_main, s.D.as<double>(), L, s.ea.as<int32_t>(), s.eb.as<int32_t>(),
                              s.mean.as<double>(), s.sxx.as<double>(),
                              s.bad.as<uint8_t>(), e_lo, e_hi, in->tested_coef, in->rlm_cov_mode,
                              s.stat.as<double>(), s.status.as<int32_t>(), d_coef, s.iters.as<int32_t>(),
                              s.scal.as<int32_t>() + 2 * SC_ERR);
      trace_mark(ctx, "k_fit_rlm", s.s_main);
      if (rc != 0) {
        set_err(ctx, "rlm kernel configuration failed (n too large for shared memory?)");
        return MDEGGS_EUNSUPPORTED;
      }
    } else if (use_dense) {
      // one GPU: all upper-triangular tile pairs; several: only the tiles that hold an edge of this rank's block
      const int ntl = (n_nodes + DT - 1) / DT;
      dim3 grid((unsigned)(ntl * (ntl + 1) / 2), (unsigned)K);
      const bool tma = ctx->dense_path != 0 && tmap_encoder() != nullptr;
      const unsigned int* d_need = nullptr;
      const int32_t* d_range = nullptr;
      if (s.world > 1 && tma) {
        const size_t wbytes = (((size_t)ntl * ntl + 31) / 32) * 4;
        CU(s.tile_need.ensure(wbytes));
        CU(cudaMemsetAsync(s.tile_need.p, 0, wbytes, s.s_main));
        int64_t nb = (e_hi - e_lo + 1023) / 1024;
        LAUNCH(k_mark_tiles, (unsigned)(nb < 1 ? 1 : (nb < 148 * 4 ? nb : 148 * 4)), 256, 0, s.s_main, s.ea.as<int32_t>(),
               s.eb.as<int32_t>(), e_lo, e_hi, ntl, s.tile_need.as<unsigned int>(), s.scal.as<int32_t>() + 2 * SC_ERR);
        d_need = s.tile_need.as<unsigned int>();
      } else if (s.world > 1) {  // (cp.async form: the tile rows the block reads)
        int32_t* rr = s.scal.as<int32_t>() + 2 * SC_ROWLO;
        CU(cudaMemsetAsync(rr, 0x7f, 4, s.s_main));                     // row_range[0] = huge
        CU(cudaMemsetAsync(rr + 1, 0xff, 4, s.s_main));                 // row_range[1] = -1
        int64_t nb = (e_hi - e_lo + 255) / 256;
        LAUNCH(k_edge_row_range, (unsigned)(nb < 148 * 8 ? nb : 148 * 8), 256, 0, s.s_main, s.ea.as<int32_t>(),
               s.eb.as<int32_t>(), e_lo, e_hi, rr, s.scal.as<int32_t>() + 2 * SC_ERR);
        d_range = rr;
      }
      bool launched = false;
      if (tma) {
        // centred copy of D with 16-sample aligned, zero-padded category segments, then the TMA-fed MMA kernel
        DenseLayout Lc;
        memset(&Lc, 0, sizeof Lc);
        Lc.K = K;
        int off = 0;
        for (int k = 0; k < K; ++k) {
          Lc.off[k] = off;
          off += (L.cnt[k] + DS - 1) / DS * DS;
        }
        Lc.off[K] = off;
        Lc.npadc = off > 0 ? off : DS;
        CU(s.Dc.ensure(nn * (size_t)Lc.npadc * 8));
        Shard::TmapKey key = {s.Dc.p, Lc.npadc, (int64_t)n_nodes, 0, DT, DS};
        bool ok = true;
        if (memcmp(&key, &s.dense_tmap_key, sizeof key) != 0) {
          const cuuint64_t gdim[2] = {(cuuint64_t)Lc.npadc, (cuuint64_t)n_nodes};
          const cuuint64_t gstr[1] = {(cuuint64_t)Lc.npadc * 8};
          const cuuint32_t box[2] = {(cuuint32_t)DS, (cuuint32_t)DT};
          const cuuint32_t estr[2] = {1, 1};
          const CUresult r = tmap_encoder()(&s.dense_tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, s.Dc.p, gdim, gstr, box, estr,
                                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
          ok = r == CUDA_SUCCESS;
          if (ok) s.dense_tmap_key = key;
          else memset(&s.dense_tmap_key, 0, sizeof s.dense_tmap_key);
        }
        if (ok) {
          if (!s.dense_tma_attr_set) {
            CU(cudaFuncSetAttribute(k_dense_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, DENSE_TMA_SMEM));
            s.dense_tma_attr_set = true;
          }
          LAUNCH(k_center_rows, blocks_for((int64_t)n_nodes * K * 32, 256), 256, 0, s.s_main, s.D.as<double>(), L, Lc,
                 s.mean.as<double>(), d_nn, s.Dc.as<double>());
          LAUNCH(k_dense_tma, grid, DENSE_TMA_THREADS, DENSE_TMA_SMEM, s.s_main, s.dense_tmap, Lc, L, (int)n_nodes, ntl,
                 d_need, s.dense_S.as<double>());
          launched = true;
        }
      }
      if (!launched) {
        if (d_need && !d_range) {
          set_err(ctx, "internal: dense fallback without its row range");
          return MDEGGS_ECUDA;
        }
        if (!s.dense_attr_set) {  // static tiles + dynamic staging slots exceed the 48 KB default
          CU(cudaFuncSetAttribute(k_dense_moments, cudaFuncAttributeMaxDynamicSharedMemorySize, DENSE_DYN_SMEM));
          s.dense_attr_set = true;
        }
        LAUNCH(k_dense_moments, grid, DTHREADS, DENSE_DYN_SMEM, s.s_main, s.D.as<double>(), L, s.mean.as<double>(),
               (int)n_nodes, ntl, d_range, s.dense_S.as<double>());
      }
    } else {
      switch (K) {
        case 2: launch_fit_stream<2>(ctx, s, L, e_lo, e_hi); break;
        case 3: launch_fit_stream<3>(ctx, s, L, e_lo, e_hi); break;
        case 4: launch_fit_stream<4>(ctx, s, L, e_lo, e_hi); break;
        case 5: launch_fit_stream<5>(ctx, s, L, e_lo, e_hi); break;
        case 6: launch_fit_stream<6>(ctx, s, L, e_lo, e_hi); break;
        case 7: launch_fit_stream<7>(ctx, s, L, e_lo, e_hi); break;
        default: launch_fit_stream<8>(ctx, s, L, e_lo, e_hi); break;
      }
    }
  }
  EVREC(s.ev[EV_FIT], s.s_main);
  double df1, df2;
  int tail_mode;
  if (K == 2) {
    df1 = 1.0;
    df2 = (double)(n - 4);
    tail_mode = is_rlm ? 1 : 0;
  } else {
    int Ke = 0;
    for (int k = 0; k < K; ++k) Ke += L.cnt[k] > 0;
    df1 = (double)(Ke - 1);
    df2 = (double)(n - 2 * Ke);
    tail_mode = 1;
  }
  if (e_hi > e_lo) {
    CU(s.cf_tab.ensure((size_t)CF_TABLE_DOUBLES * 8));
    TailConsts tc;
    tc.a = tail_mode == 0 ? 0.5 : df1 / 2.0;
    tc.b = df2 / 2.0;
    tc.lbeta_ab = 0.0;
    tc.e_ab = s.cf_tab.as<double>();
    tc.e_ba = s.cf_tab.as<double>() + CF_TABLE_LEN;
    tc.c_ab = s.cf_tab.as<double>() + 2 * CF_TABLE_LEN + 2;
    tc.c_ba = s.cf_tab.as<double>() + 3 * CF_TABLE_LEN + 2;
    tc.sp = (s.tail_sp_on && tail_mode == 0) ? s.tail_sp.as<double>() : nullptr;
    tc.sp_inv_h = (double)TS_M / tail_spline_smax(df2);
    tc.sp_inv_n = 1.0 / df2;
    const double* d_lbeta = s.cf_tab.as<double>() + 2 * CF_TABLE_LEN;
    const int src = is_rlm ? 2 : (use_dense ? 1 : 0);
    CU(cudaStreamWaitEvent(s.s_main, s.ev_cf, 0));  // continued-fraction tables (side stream)
    if (s.order_tested_only) CU(cudaStreamWaitEvent(s.s_main, s.ev_join, 0));  // activity masks (k_finish filters the order)
    // single GPU: every edge passes through this launch, which then also feeds the p-value buckets of K8
    const bool bs = use_bucket_order(ctx, in, n_nodes);
    const int do_bs = bs && s.world == 1;
    BsArgs B;
    memset(&B, 0, sizeof B);
    if (bs) {
      const int NB = bs_nbuckets(E);
      CU(s.bs_count.ensure((size_t)2 * bs_padded(NB) * 4));
      CU(s.bs_slot.ensure((size_t)E * 4));
      CU(cudaMemsetAsync(s.bs_count.p, 0, (size_t)(NB + 1) * 4, s.s_main));
      B = make_bs_args(s, E);
    }
    if (e_hi - e_lo >= FIN_DEFER_MIN_EDGES) CU(s.fin_q.ensure((size_t)(e_hi - e_lo) * 4));  // queue of slow tails
    const NaPairPlan NP = na_pair_plan(in);
    if (NP.on) {
      CU(s.na_q.ensure((size_t)(e_hi - e_lo) * 4));
      if (NP.smem > 48 * 1024 && !(s.na_attr_mask & (1u << K))) {
        cudaError_t ae = cudaSuccess;
        switch (K) {
          case 2: ae = cudaFuncSetAttribute(k_fit_na_pairs<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); break;
          case 3: ae = cudaFuncSetAttribute(k_fit_na_pairs<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); break;
          case 4: ae = cudaFuncSetAttribute(k_fit_na_pairs<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); break;
          case 5: ae = cudaFuncSetAttribute(k_fit_na_pairs<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); break;
          case 6: ae = cudaFuncSetAttribute(k_fit_na_pairs<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); break;
          case 7: ae = cudaFuncSetAttribute(k_fit_na_pairs<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); break;
          default: ae = cudaFuncSetAttribute(k_fit_na_pairs<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); break;
        }
        CU(ae);
        s.na_attr_mask |= 1u << K;
      }
    }
    switch (K) {
      case 2: launch_finish<2>(ctx, s, L, src, n_nodes, e_lo, e_hi, tail_mode, df1, df2, tc, d_lbeta, d_coef, do_bs, B, NP); break;
      case 3: launch_finish<3>(ctx, s, L, src, n_nodes, e_lo, e_hi, tail_mode, df1, df2, tc, d_lbeta, d_coef, do_bs, B, NP); break;
      case 4: launch_finish<4>(ctx, s, L, src, n_nodes, e_lo, e_hi, tail_mode, df1, df2, tc, d_lbeta, d_coef, do_bs, B, NP); break;
      case 5: launch_finish<5>(ctx, s, L, src, n_nodes, e_lo, e_hi, tail_mode, df1, df2, tc, d_lbeta, d_coef, do_bs, B, NP); break;
      case 6: launch_finish<6>(ctx, s, L, src, n_nodes, e_lo, e_hi, tail_mode, df1, df2, tc, d_lbeta, d_coef, do_bs, B, NP); break;
      case 7: launch_finish<7>(ctx, s, L, src, n_nodes, e_lo, e_hi, tail_mode, df1, df2, tc, d_lbeta, d_coef, do_bs, B, NP); break;
      default: launch_finish<8>(ctx, s, L, src, n_nodes, e_lo, e_hi, tail_mode, df1, df2, tc, d_lbeta, d_coef, do_bs, B, NP); break;
    }
  }
  EVREC(s.ev[EV_TAIL], s.s_main);
  return MDEGGS_OK;
}


// ---- peer-memory all-gather set-up (one process per GPU) ------------------------------------------------------------
// All ranks call this at the same point of the same call (buffer generations are identical everywhere).  Handles travel
// through the NCCL communicator; the outcome is agreed on with an all-reduce, so either every rank pushes or none does.
int p2p_prepare(mdeggs_ctx* ctx, Shard& s) {
  Shard::P2P& X = s.p2p;
  X.used = false;
  if (!ctx->p2p || X.failed || s.world < 2 || s.world > P2P_MAX || ctx->shards.size() != 1) return MDEGGS_OK;
  if (X.ok && X.gen_p == s.p_edge.gen && X.gen_st == s.st8.gen) return MDEGGS_OK;
  if (ctx->capturing) {  // (cannot happen: captured calls find every buffer in place) - NCCL for this graph
    X.ok = false;
    return MDEGGS_OK;
  }
  NcclApi& api = nccl_api();
  const int W = s.world;
  struct Rec {
    cudaIpcMemHandle_t p, st, fl;
  };
  static_assert(sizeof(Rec) == 192, "three 64-byte handles");
  bool good = true;
  if (!X.flags_ok) {
    if (s.p2p_flags.ensure(64 * 4) != cudaSuccess || cudaMemsetAsync(s.p2p_flags.p, 0, 64 * 4, s.s_main) != cudaSuccess)
      good = false;
  }
  Rec mine;
  memset(&mine, 0, sizeof mine);
  if (good && (cudaIpcGetMemHandle(&mine.p, s.p_edge.p) != cudaSuccess ||
               cudaIpcGetMemHandle(&mine.st, s.st8.p) != cudaSuccess ||
               cudaIpcGetMemHandle(&mine.fl, s.p2p_flags.p) != cudaSuccess))
    good = false;
  cudaGetLastError();
  CU(s.p2p_x.ensure((size_t)W * sizeof(Rec) + 64));
  std::vector<Rec> all((size_t)W);
  char* dx = s.p2p_x.as<char>();
  CU(cudaMemcpyAsync(dx + (size_t)s.rank * sizeof(Rec), &mine, sizeof(Rec), cudaMemcpyHostToDevice, s.s_main));
  NC(api.AllGather(dx + (size_t)s.rank * sizeof(Rec), dx, sizeof(Rec), NCCL_INT8, s.comm, s.s_main));
  CU(cudaMemcpyAsync(all.data(), dx, (size_t)W * sizeof(Rec), cudaMemcpyDeviceToHost, s.s_main));
  CU(cudaStreamSynchronize(s.s_main));
  // (old mappings of the data buffers go first: the owners have reallocated them)
  for (int r = 0; r < W; ++r) {
    if (r == s.rank) continue;
    if (X.peer_p[r]) cudaIpcCloseMemHandle(X.peer_p[r]);
    if (X.peer_st[r]) cudaIpcCloseMemHandle(X.peer_st[r]);
    X.peer_p[r] = X.peer_st[r] = nullptr;
  }
  cudaGetLastError();
  for (int r = 0; r < W && good; ++r) {
    if (r == s.rank) {
      X.peer_p[r] = s.p_edge.p;
      X.peer_st[r] = s.st8.p;
      X.peer_flags[r] = s.p2p_flags.p;
      continue;
    }
    if (cudaIpcOpenMemHandle(&X.peer_p[r], all[(size_t)r].p, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess ||
        cudaIpcOpenMemHandle(&X.peer_st[r], all[(size_t)r].st, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess)
      good = false;
    if (good && !X.flags_ok &&
        cudaIpcOpenMemHandle(&X.peer_flags[r], all[(size_t)r].fl, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess)
      good = false;
  }
  cudaGetLastError();
  // agreement: the minimum over ranks of "all my mappings are in place"
  int32_t flag = good ? 1 : 0;
  int32_t* dflag = reinterpret_cast<int32_t*>(dx + (size_t)W * sizeof(Rec));
  CU(cudaMemcpyAsync(dflag, &flag, 4, cudaMemcpyHostToDevice, s.s_main));
  NC(api.AllReduce(dflag, dflag, 1, NCCL_INT32, NCCL_MIN, s.comm, s.s_main));
  CU(cudaMemcpyAsync(&flag, dflag, 4, cudaMemcpyDeviceToHost, s.s_main));
  CU(cudaStreamSynchronize(s.s_main));
  if (flag != 1) {
    X.ok = false;
    X.failed = true;  // stay on NCCL (every rank takes the same decision)
    return MDEGGS_OK;
  }
  X.flags_ok = true;
  X.ok = true;
  X.gen_p = s.p_edge.gen;
  X.gen_st = s.st8.gen;
  return MDEGGS_OK;
}


// one process, several GPUs: the same push, with the peers' buffers addressed directly (peer access enabled once)
int p2p_prepare_local(mdeggs_ctx* ctx) {
  const int W = (int)ctx->shards.size();
  Shard& s0 = ctx->shards[0];
  for (Shard& s : ctx->shards) s.p2p.used = false;
  if (!ctx->p2p || s0.p2p.failed || W < 2 || W > P2P_MAX) return MDEGGS_OK;
  bool same = s0.p2p.ok;
  for (Shard& s : ctx->shards) same = same && s.p2p.gen_p == s.p_edge.gen && s.p2p.gen_st == s.st8.gen;
  if (same) return MDEGGS_OK;
  auto fail = [&]() {
    for (Shard& s : ctx->shards) s.p2p.ok = false;
    s0.p2p.failed = true;
    cudaGetLastError();
    return MDEGGS_OK;
  };
  if (!s0.p2p.flags_ok) {
    // (a wait kernel of one GPU is enqueued before the push kernels of the next ones: launches must not block)
    if (cudaSetDevice(s0.device) != cudaSuccess || !launches_are_async(ctx, s0)) return fail();
    for (int i = 0; i < W; ++i)
      for (int j = 0; j < W; ++j) {
        if (i == j) continue;
        int can = 0;
        if (cudaDeviceCanAccessPeer(&can, ctx->shards[i].device, ctx->shards[j].device) != cudaSuccess || !can) return fail();
        if (cudaSetDevice(ctx->shards[i].device) != cudaSuccess) return fail();
        const cudaError_t e = cudaDeviceEnablePeerAccess(ctx->shards[j].device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail();
        cudaGetLastError();
      }
    for (Shard& s : ctx->shards) {
      if (cudaSetDevice(s.device) != cudaSuccess || s.p2p_flags.ensure(64 * 4) != cudaSuccess ||
          cudaMemset(s.p2p_flags.p, 0, 64 * 4) != cudaSuccess)
        return fail();
    }
    for (Shard& s : ctx->shards) s.p2p.flags_ok = true;
  }
  for (Shard& s : ctx->shards) {
    for (int r = 0; r < W; ++r) {
      s.p2p.peer_p[r] = ctx->shards[r].p_edge.p;
      s.p2p.peer_st[r] = ctx->shards[r].st8.p;
      s.p2p.peer_flags[r] = ctx->shards[r].p2p_flags.p;
    }
    s.p2p.gen_p = s.p_edge.gen;
    s.p2p.gen_st = s.st8.gen;
    s.p2p.ok = true;
  }
  return MDEGGS_OK;
}

// all-gather of the per-rank edge blocks (p-values, status and whatever per-edge output was requested): peer-memory
// push (mdeggs_p2p.cuh), NCCL as the fallback and for the optional per-edge outputs
int run_collective(mdeggs_ctx* ctx, const mdeggs_problem* in, const mdeggs_result* out) {
  const int64_t E = in->E;
  Shard& s0 = ctx->shards[0];
  if (s0.world > 1 && E > 0) {
    NcclApi& api = nccl_api();
    const int64_t chunk = (E + s0.world - 1) / s0.world;
    const bool is_rlm = (in->K == 2 && in->method == MDEGGS_METHOD_RLM);
    for (Shard& s : ctx->shards) {
      CU(cudaSetDevice(s.device));
      CU(s.st8.ensure((size_t)chunk * s0.world));
    }
    bool pushed = false;
    {
      int rc = ctx->shards.size() == 1 ? p2p_prepare(ctx, s0) : p2p_prepare_local(ctx);
      if (rc) return rc;
      if (ctx->p2p && s0.p2p.ok) {  // (option p2p = 0 keeps valid mappings but sends through NCCL)
        for (Shard& s : ctx->shards) {  // every rank pushes its block ...
          CU(cudaSetDevice(s.device));
          PushArgs& A = s.p2p.args;
          memset(&A, 0, sizeof A);
          for (int r = 0; r < s.world; ++r) {
            A.p_dst[r] = static_cast<double*>(s.p2p.peer_p[r]);
            A.st_dst[r] = static_cast<int8_t*>(s.p2p.peer_st[r]);
            A.flag_dst[r] = static_cast<unsigned int*>(s.p2p.peer_flags[r]);
          }
          A.status = s.status.as<int32_t>();
          A.world = s.world;
          A.rank = s.rank;
          A.lo = (long long)((int64_t)s.rank * chunk < E ? (int64_t)s.rank * chunk : E);
          A.n_own = (long long)(A.lo + chunk < E ? chunk : E - A.lo);
          A.ctl = s.p2p_flags.as<unsigned int>() + 2 * P2P_MAX;
          A.err = s.scal.as<int32_t>() + 2 * SC_ERR;
          LAUNCH(k_push_gather, (unsigned)(s.sm_count * 4), 256, 0, s.s_main, A);
          s.p2p.used = true;
        }
        for (Shard& s : ctx->shards) {  // ... and waits for everybody else's
          CU(cudaSetDevice(s.device));
          LAUNCH(k_wait_gather, 1, 32, 0, s.s_main, s.p2p.args);
        }
        pushed = true;
      }
    }
    if (!pushed)
      for (Shard& s : ctx->shards) {  // status codes as bytes
        CU(cudaSetDevice(s.device));
        const int64_t lo = (int64_t)s.rank * chunk < E ? (int64_t)s.rank * chunk : E;
        const int64_t n_own = lo + chunk < E ? chunk : E - lo;
        if (n_own > 0)
          LAUNCH(k_pack_status, blocks_for(n_own, 256), 256, 0, s.s_main, s.status.as<int32_t>() + lo, n_own,
                 s.st8.as<int8_t>() + lo);
      }
    NC(api.GroupStart());
    for (Shard& s : ctx->shards) {
      CU(cudaSetDevice(s.device));
      char* pe = (char*)s.p_edge.p;
      char* st = (char*)s.st8.p;
      if (!pushed) {
        NC(api.AllGather(pe + (size_t)s.rank * chunk * 8, pe, (size_t)chunk, NCCL_FLOAT64, s.comm, s.s_main));
        NC(api.AllGather(st + (size_t)s.rank * chunk, st, (size_t)chunk, NCCL_INT8, s.comm, s.s_main));
      }
      if (out->stat) {
        char* x = (char*)s.stat.p;
        NC(api.AllGather(x + (size_t)s.rank * chunk * 8, x, (size_t)chunk, NCCL_FLOAT64, s.comm, s.s_main));
      }
      if (out->coef) {
        char* x = (char*)s.coef.p;
        size_t per = (size_t)chunk * 2 * in->K;
        NC(api.AllGather(x + (size_t)s.rank * per * 8, x, per, NCCL_FLOAT64, s.comm, s.s_main));
      }
      if (out->iters && is_rlm) {
        char* x = (char*)s.iters.p;
        NC(api.AllGather(x + (size_t)s.rank * chunk * 4, x, (size_t)chunk, NCCL_INT32, s.comm, s.s_main));
      }
    }
    NC(api.GroupEnd());
    for (Shard& s : ctx->shards) {
      CU(cudaSetDevice(s.device));
      LAUNCH(k_unpack_status, blocks_for(E, 256), 256, 0, s.s_main, s.st8.as<int8_t>(), E, s.status.as<int32_t>());
    }
  }
  for (Shard& s : ctx->shards) {
    CU(cudaSetDevice(s.device));
    EVREC(s.ev[EV_COLL], s.s_main);
    if (s0.world > 1 && E > 0) trace_mark(ctx, s0.p2p.used ? "(peer push p, status)" : "(ncclAllGather p, status)", s.s_main);
  }
  return MDEGGS_OK;
}

int run_shard_back(mdeggs_ctx* ctx, Shard& s, const mdeggs_problem* in, const mdeggs_result* out, const HostLayout& H,
                   int32_t n_nodes, bool emit) {
  const int K = in->K, P = in->P;
  const int64_t E = in->E;
  CU(cudaSetDevice(s.device));
  const int32_t* d_nn = s.scal.as<int32_t>() + 2 * SC_NNODES;
  int32_t* d_best = s.scal.as<int32_t>() + 2 * SC_BEST;
  const int act_stride = n_nodes > 0 ? n_nodes : 1;
  const int Pn = P > 0 ? P : 1;
  const bool dev_adjust = device_adjust(in);
  const bool do_adjust = dev_adjust && E > 0 && P > 0 && n_nodes > 0;
  s.select_fused = false;
  // K7 percolation (category of every edge at every percentile, counts) does not need the p order: it runs on the side
  // stream, after the activity masks, while the main stream orders the edges; both join before K8
  // several ranks: the p-values arrived through the all-gather, so the bucket counting could not ride in k_finish; it
  // rides in k_percolate instead (its buffers are cleared before the side stream forks off)
  // several ranks: the gathered p-values still have to be dropped into their buckets: by k_bsort_count on the main stream,
  // beside the (capped) k_percolate of the side stream [default: config 5 on 2 GPUs 1.91 -> 1.87 ms], or -
  // MDEGGS_COUNT_IN_PERC=1 - inside k_percolate (one pass over the edges instead of two, but then the ordering waits for it)
  static const int count_in_perc_env = getenv("MDEGGS_COUNT_IN_PERC") ? atoi(getenv("MDEGGS_COUNT_IN_PERC")) : 0;
  const bool count_in_percolate = count_in_perc_env && do_adjust && s.world > 1 && use_bucket_order(ctx, in, n_nodes) && s.order_tested_only;
  if (count_in_percolate) {
    CU(s.bs_count.ensure((size_t)2 * bs_padded(bs_nbuckets(E)) * 4));
    CU(s.bs_slot.ensure((size_t)E * 4));
    CU(cudaMemsetAsync(s.bs_count.p, 0, (size_t)(bs_nbuckets(E) + 1) * 4, s.s_main));
  }
  // single GPU, many buckets: the device-wide scan of the bucket counts (they are final: k_finish wrote them) runs BEFORE
  // the side stream forks off - beside the compute-bound k_percolate its look-back CTAs crawl (137 us instead of 8)
  bool scan_done_early = false;
  if (do_adjust && s.world == 1 && use_bucket_order(ctx, in, n_nodes) && bs_nbuckets(E) > BS_SCAN_SMEM) {
    BsArgs B0 = make_bs_args(s, E);
    const int NB0 = bs_nbuckets(E);
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, B0.count, B0.base, NB0 + 1, s.s_main);
    CU(s.cub_tmp.ensure(tb));
    CU(cub::DeviceScan::ExclusiveSum(s.cub_tmp.p, tb, B0.count, B0.base, NB0 + 1, s.s_main));
    ctx->launches += 2;
    trace_mark(ctx, "cub::ExclusiveSum", s.s_main);
    scan_done_early = true;
  }
  CU(cudaEventRecord(s.ev_fork, s.s_main));  // p-values and status of every edge are final
  CU(cudaStreamWaitEvent(s.s_sort, s.ev_fork, 0));
  // counters: tot[P] fail[P] sig[P] sig_raw[P] (u64), then out: tot_out[P] sig_out[P] fail_out[P] (i64)
  // ... and one word: "CTAs done" counter of the k_bh_final + selection tail
  // ... then seg_cnt[P][K] (u32): members of every (percentile, category) segment, for the one-launch K8
  const size_t counters_bytes = (size_t)Pn * 8 * 7 + 8 + (size_t)Pn * K * 4;
  CU(s.counters.ensure(counters_bytes));
  CU(cudaMemsetAsync(s.counters.p, 0, counters_bytes, s.s_sort));
  unsigned int* d_seg_cnt = reinterpret_cast<unsigned int*>(s.counters.as<char>() + (size_t)Pn * 8 * 7 + 8);
  const int bh_ntiles = (int)((E + BH_TILE - 1) / BH_TILE);
  const bool onepass_shape = ctx->bh_onepass && E > 0 && bh_ntiles <= BH1_MAX_TILES;
  // count-only K8 of big jobs skips the tiles of the p order beyond 0.05: the segment sizes then come from k_percolate
  const bool skip_shape = ctx->count_skip && do_adjust && in->padj != MDEGGS_PADJ_QVALUE && in->padj != MDEGGS_PADJ_HOMMEL &&
                          use_bucket_order(ctx, in, n_nodes);  // (n_low comes from the bucket offsets)
  const bool want_seg_cnt = onepass_shape || skip_shape;
  if (onepass_shape) {  // tile words of k_bh_onepass: 0 = not published yet
    const size_t words = (size_t)Pn * K * bh_ntiles;
    CU(s.bh1_state.ensure(words * 12));
    CU(cudaMemsetAsync(s.bh1_state.p, 0, words * 12, s.s_sort));
  }
  unsigned long long* c_tot = s.counters.as<unsigned long long>();
  unsigned long long* c_fail = c_tot + Pn;
  unsigned long long* c_sig = c_tot + 2 * Pn;
  unsigned long long* c_raw = c_tot + 3 * Pn;
  CU(s.act.ensure((size_t)Pn * act_stride));
  const bool dev_out = in->mem == MDEGGS_MEM_DEVICE;
  int8_t* d_cat = nullptr;
  if (out->cat && E > 0 && P > 0) {
    if (dev_out && emit) d_cat = out->cat;
    else {
      CU(s.cat.ensure((size_t)P * E));
      d_cat = s.cat.as<int8_t>();
    }
  }
  // adjusted values of every percentile are kept when they fit a modest scratch matrix (the best row is then a copy);
  // multi-GPU: the (percentile, category) segments are independent, so rank r adjusts percentiles r, r+world, ...
  // and only the per-percentile significant counts are all-reduced; a full P x E p.adj output keeps it replicated
  double* d_padj = nullptr;
  const bool want_padj_best = out->padj_best && do_adjust;
  const bool shard_p = do_adjust && s.world > 1 && !out->padj && in->padj != MDEGGS_PADJ_HOMMEL;
  const bool keep_all_padj = do_adjust && !shard_p &&
                             (out->padj || (want_padj_best && ((size_t)P * (size_t)E * 8 <= (size_t)ctx->padj_keep_cap ||
                                                                in->padj == MDEGGS_PADJ_HOMMEL)));
  if (keep_all_padj) {
    if (out->padj && dev_out && emit) d_padj = out->padj;
    else {
      CU(s.padj.ensure((size_t)P * E * 8));
      d_padj = s.padj.as<double>();
    }
  }
  if (n_nodes > 0 && P > 0 && E > 0) {
    // (also writes NaN into the p.adj rows of the edges that are not tested at a percentile)
    const size_t pc_smem = (size_t)(3 * P + (want_seg_cnt ? P * K : 0)) * sizeof(unsigned);
    // long edge lists: a capped grid (half of every SM's thread slots) walks the list with a grid stride, so that the
    // ordering kernels of the main stream run beside this compute-bound kernel instead of queueing behind it
    unsigned pc_grid = blocks_for(E, 256);
    {
      static const int cap_env = getenv("MDEGGS_PERC_CAP") ? atoi(getenv("MDEGGS_PERC_CAP")) : 4;
      const unsigned cap = (unsigned)(s.sm_count * cap_env);
      if (do_adjust && !count_in_percolate && cap_env > 0 && pc_grid > 4 * cap) pc_grid = cap;
    }
    if (count_in_percolate) {  // several ranks: the gathered p-values are dropped into their buckets in the same pass
      PercolateBucketTail tail{make_bs_args(s, E)};
      LAUNCH(k_percolate<PercolateBucketTail>, pc_grid, 256, pc_smem, s.s_sort, tail, s.ea.as<int32_t>(),
             s.eb.as<int32_t>(), s.status.as<int32_t>(), s.p_edge.as<double>(), s.act_t.as<uint4>(), s.act_nch, E, P, c_tot,
             c_fail, c_raw, d_cat, d_padj, s.scal.as<int32_t>() + 2 * SC_ERR, K, want_seg_cnt ? d_seg_cnt : nullptr);
    } else {
      LAUNCH(k_percolate<PercolateNoTail>, pc_grid, 256, pc_smem, s.s_sort, PercolateNoTail{}, s.ea.as<int32_t>(),
             s.eb.as<int32_t>(), s.status.as<int32_t>(), s.p_edge.as<double>(), s.act_t.as<uint4>(), s.act_nch, E, P, c_tot,
             c_fail, c_raw, d_cat, d_padj, s.scal.as<int32_t>() + 2 * SC_ERR, K, want_seg_cnt ? d_seg_cnt : nullptr);
    }
  }
  CU(cudaEventRecord(s.ev_join, s.s_sort));
  if (do_adjust) {
    // The raw p-value of an edge does not depend on the percentile: sort ALL edges by p once. This needs no
    // cut-off, so it is enqueued BEFORE joining the quantile stream and overlaps the tail of the radix select.
    CU(s.sortk.ensure((size_t)E * 8));
    CU(s.sortk2.ensure((size_t)E * 8));
    CU(s.idx.ensure((size_t)E * 4));
    CU(s.idx2.ensure((size_t)E * 4));
    CU(s.ps.ensure((size_t)E * 8));
    CU(s.sa.ensure((size_t)E * 4));
    CU(s.sb.ensure((size_t)E * 4));
    if (use_bucket_order(ctx, in, n_nodes)) {
      // bucket ordering (mdeggs_bsort.cuh): counts/slots came out of k_finish (single GPU) or from a counting pass
      // over the all-gathered p-values; scatter into buckets, then one CTA per run of buckets sorts in shared memory
      BsArgs B;
      if (count_in_percolate) {
        B = make_bs_args(s, E);
        CU(cudaStreamWaitEvent(s.s_main, s.ev_join, 0));  // bucket counts (k_percolate on the side stream)
      } else if (s.world > 1) {
        CU(s.bs_count.ensure((size_t)2 * bs_padded(bs_nbuckets(E)) * 4));
        CU(s.bs_slot.ensure((size_t)E * 4));
        CU(cudaMemsetAsync(s.bs_count.p, 0, (size_t)(bs_nbuckets(E) + 1) * 4, s.s_main));
        B = make_bs_args(s, E);
        LAUNCH(k_bsort_count, blocks_for(E, 256), 256, 0, s.s_main, B, s.p_edge.as<double>(),
               s.scal.as<int32_t>() + 2 * SC_ERR);
      } else {
        B = make_bs_args(s, E);
      }
      const int NB = bs_nbuckets(E);
      const int scan_in_cta = NB <= BS_SCAN_SMEM;
      if (!scan_in_cta && !scan_done_early) {  // many buckets: device-wide scan of count[0..NB] (count[NB] == 0, so base[NB] = total)
        size_t tb = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tb, B.count, B.base, NB + 1, s.s_main);
        CU(s.cub_tmp.ensure(tb));
        CU(cub::DeviceScan::ExclusiveSum(s.cub_tmp.p, tb, B.count, B.base, NB + 1, s.s_main));
        ctx->launches += 2;
        trace_mark(ctx, "cub::ExclusiveSum", s.s_main);
      }
      CU(s.run_first.ensure((size_t)(E / BS_CAP + 4) * 4));
      B.run_first = s.run_first.as<int>();
      LAUNCH(k_bsort_scatter, blocks_for(E, 256), 256, scan_in_cta ? (size_t)bs_padded(NB) * 4 : 0, s.s_main, B, scan_in_cta,
             s.p_edge.as<double>(), s.sortk.as<unsigned long long>(), s.idx.as<int32_t>(),
             s.scal.as<int32_t>() + 2 * SC_ERR);
      LAUNCH(k_bsort_sort, blocks_for(E, BS_CAP), BS_SORT_THREADS, 0, s.s_main, B, s.sortk.as<unsigned long long>(),
             s.idx.as<int32_t>(), s.ps.as<double>(), s.sa.as<int32_t>(),
             s.sb.as<int32_t>(), s.idx2.as<int32_t>(), s.scal.as<int32_t>() + 2 * SC_ERR);
    } else {
      LAUNCH(k_sort_keys, blocks_for(E, 256), 256, 0, s.s_main, s.p_edge.as<double>(), E,
             s.sortk.as<unsigned long long>(), s.idx.as<int32_t>());
      size_t tb = 0;
      cub::DeviceRadixSort::SortPairs(nullptr, tb, s.sortk.as<unsigned long long>(), s.sortk2.as<unsigned long long>(),
                                      s.idx.as<int32_t>(), s.idx2.as<int32_t>(), E, 0, 64, s.s_main);
      CU(s.cub_tmp.ensure(tb));
      CU(cub::DeviceRadixSort::SortPairs(s.cub_tmp.p, tb, s.sortk.as<unsigned long long>(),
                                         s.sortk2.as<unsigned long long>(), s.idx.as<int32_t>(), s.idx2.as<int32_t>(), E,
                                         0, 64, s.s_main));
      ctx->launches += 10;
      trace_mark(ctx, "cub::SortPairs", s.s_main);
      LAUNCH(k_gather_sorted, blocks_for(E, 256), 256, 0, s.s_main, s.sortk2.as<unsigned long long>(),
             s.idx2.as<int32_t>(), s.ea.as<int32_t>(), s.eb.as<int32_t>(), E,
             s.ps.as<double>(), s.sa.as<int32_t>(), s.sb.as<int32_t>(), s.scal.as<unsigned long long>() + SC_NVALID,
             s.scal.as<unsigned long long>() + SC_NORDER, s.scal.as<int32_t>() + 2 * SC_ERR);
    }
  }
  CU(cudaStreamWaitEvent(s.s_main, s.ev_join, 0));  // categories, counts (and NaN rows of the p.adj scratch) ready
  trace_mark(ctx, "(join side stream)", s.s_main);
  EVREC(s.ev[EV_PERC], s.s_main);
  // ---- K8 ----
  if (do_adjust) {
    const int ntiles = (int)((E + BH_TILE - 1) / BH_TILE);
    const size_t nseg = (size_t)P * K;
    CU(s.tile_cnt.ensure(nseg * ntiles * 4));
    CU(s.tile_base.ensure(nseg * ntiles * 8));
    CU(s.tile_min.ensure(nseg * ntiles * 8));
    CU(s.tile_carry.ensure(nseg * ntiles * 8));
    CU(s.seg_n.ensure(nseg * 8));
    const int p0 = shard_p ? s.rank : 0, stride = shard_p ? s.world : 1;
    const int n_loc = p0 < P ? (P - p0 + stride - 1) / stride : 0;
    // one rank: the percentile selection rides in the last CTA of k_bh_final (with more ranks it follows the all-reduce)
    const bool count_only = !d_padj && ctx->count_only_adjust;  // (run_adjust takes the k_bh_sigrank path)
    s.select_fused = s.world == 1 && n_loc > 0 && !count_only && in->padj != MDEGGS_PADJ_HOMMEL;
    s.partials_complete = !count_only;
    const bool skip = count_only && skip_shape;
    s.tiles_complete = !skip;
    CU(s.small_out.ensure(small_block_bytes(P)));
    SelArgs sel = make_sel_args(s, in, H, emit, c_sig);
    sel.done = reinterpret_cast<unsigned int*>(c_tot + 7 * Pn);
    int rc = run_adjust(ctx, s, in, n_loc * K, p0, stride, nullptr, c_sig, d_padj, E, act_stride, false,
                        s.select_fused ? &sel : nullptr, false, skip ? 0.05 : 0.0);
    if (rc) return rc;
  }
  return MDEGGS_OK;
}

// all-reduce of the per-percentile significant counts when the adjustment was sharded over the ranks
int run_sig_allreduce(mdeggs_ctx* ctx, const mdeggs_problem* in, const mdeggs_result* out) {
  Shard& s0 = ctx->shards[0];
  const bool dev_adjust = device_adjust(in);
  if (!(dev_adjust && s0.world > 1 && !out->padj && in->E > 0 && in->P > 0) || in->padj == MDEGGS_PADJ_HOMMEL)
    return MDEGGS_OK;  // (hommel is not sharded by percentile: every rank already holds every count)
  NcclApi& api = nccl_api();
  const int Pn = in->P;
  NC(api.GroupStart());
  for (Shard& s : ctx->shards) {
    CU(cudaSetDevice(s.device));
    unsigned long long* c_sig = s.counters.as<unsigned long long>() + 2 * Pn;
    NC(api.AllReduce(c_sig, c_sig, (size_t)Pn, NCCL_UINT64, NCCL_SUM, s.comm, s.s_main));
  }
  NC(api.GroupEnd());
  for (Shard& s : ctx->shards) trace_mark(ctx, "(ncclAllReduce sig counts)", s.s_main);
  return MDEGGS_OK;
}

int run_shard_back2(mdeggs_ctx* ctx, Shard& s, const mdeggs_problem* in, const mdeggs_result* out, const HostLayout& H,
                    int32_t n_nodes, bool emit) {
  const int G = in->G, K = in->K, P = in->P;
  const int64_t E = in->E;
  (void)H;
  CU(cudaSetDevice(s.device));
  int32_t* d_best = s.scal.as<int32_t>() + 2 * SC_BEST;
  const int act_stride = n_nodes > 0 ? n_nodes : 1;
  const int Pn = P > 0 ? P : 1;
  unsigned long long* c_tot = s.counters.as<unsigned long long>();
  unsigned long long* c_sig = c_tot + 2 * Pn;
  unsigned long long* c_raw = c_tot + 3 * Pn;
  const bool dev_out = in->mem == MDEGGS_MEM_DEVICE;
  const bool dev_adjust = device_adjust(in);
  const bool do_adjust = dev_adjust && E > 0 && P > 0 && n_nodes > 0;
  const bool want_padj_best = out->padj_best && do_adjust;
  const bool shard_p = do_adjust && s.world > 1 && !out->padj && in->padj != MDEGGS_PADJ_HOMMEL;
  const bool keep_all_padj = do_adjust && !shard_p &&
                             (out->padj || (want_padj_best && ((size_t)P * (size_t)E * 8 <= (size_t)ctx->padj_keep_cap ||
                                                                in->padj == MDEGGS_PADJ_HOMMEL)));
  double* d_padj = nullptr;
  if (keep_all_padj) d_padj = (out->padj && dev_out && emit) ? out->padj : s.padj.as<double>();
  const unsigned long long* sig_src = dev_adjust ? c_sig : c_raw;
  if (!s.select_fused) {
    CU(s.small_out.ensure(small_block_bytes(P)));
    LAUNCH(k_select_best, 1, 32, 0, s.s_main, make_sel_args(s, in, H, emit, sig_src));
  }
  // rows of the chosen percentile
  const bool want_cat_best = out->cat_best && E > 0 && in->padj != MDEGGS_PADJ_HOST;
  if (want_cat_best) CU(s.cat_best.ensure((size_t)E));
  if (want_padj_best) CU(s.padj_best.ensure((size_t)E * 8));
  const bool row_copy = want_padj_best && keep_all_padj;
  if (want_padj_best && !row_copy) {  // P x E too large to keep: rerun the adjustment for the best percentile only
    LAUNCH(k_fill_f64, (unsigned)(blocks_for(E, 256) < 148u * 16 ? blocks_for(E, 256) : 148u * 16), 256, 0, s.s_main,
           s.padj_best.as<double>(), E, nan(""));
    // single rank: the tile partials of every percentile are still there, only the final pass is repeated;
    // sharded K8: this rank may not own the best percentile, so the whole adjustment is rerun for it
    int rc = (shard_p || !s.tiles_complete)
                 ? run_adjust(ctx, s, in, K, 0, 1, d_best, nullptr, s.padj_best.as<double>(), E, act_stride)
                     : run_adjust(ctx, s, in, K, 0, 1, d_best, nullptr, s.padj_best.as<double>(), E, act_stride, true,
                                  nullptr, !s.partials_complete);
    if (rc) return rc;
  }
  // device callers: these two rows are produced by k_emit straight into the caller's buffers
  s.best_rows_in_emit = dev_out && emit;
  s.best_row_src = row_copy ? d_padj : nullptr;
  if ((want_cat_best || row_copy) && !s.best_rows_in_emit)
    LAUNCH(k_best_rows, blocks_for(E, 256), 256, 0, s.s_main, s.ea.as<int32_t>(), s.eb.as<int32_t>(),
           s.act.as<uint8_t>(), act_stride, E, d_best,
           want_cat_best ? s.cat_best.as<int8_t>() : nullptr, row_copy ? d_padj : nullptr,
           row_copy ? s.padj_best.as<double>() : nullptr, s.scal.as<int32_t>() + 2 * SC_ERR);
  if (emit && out->median) {
    CU(s.med_out.ensure((size_t)G * K * 8));
    LAUNCH(k_scatter_median, blocks_for((int64_t)G * K, 256), 256, 0, s.s_main, s.med.as<double>(),
           s.node_of_row.as<int32_t>(), G, K, s.med_out.as<double>());
  }
  if (getenv("MDEGGS_HOSTTIME")) LAUNCH(k_stamp, 1, 1, 0, s.s_main, s.stamps.as<unsigned long long>() + 1);
  EVREC(s.ev[EV_ADJ], s.s_main);
  return MDEGGS_OK;
}

// Phase C (always eager): results -> caller buffers.  The small results travel as one packed block (k_select_best):
// device callers get everything through ONE copy kernel, host callers through one D2H per large array plus one for
// the block, which finish_small_outputs() scatters after the final synchronisation.
int emit_outputs(mdeggs_ctx* ctx, Shard& s, const mdeggs_problem* in, const mdeggs_result* out, const HostLayout& H,
                 int32_t n_nodes, bool record_end = true) {
  const int G = in->G, K = in->K, P = in->P;
  const int64_t E = in->E;
  (void)H;
  CU(cudaSetDevice(s.device));
  const int Pn = P > 0 ? P : 1;
  const bool dev_out = in->mem == MDEGGS_MEM_DEVICE;
  const bool dev_adjust = device_adjust(in);
  const int8_t* d_cat = (out->cat && E > 0 && P > 0) ? (dev_out ? out->cat : s.cat.as<int8_t>()) : nullptr;
  const double* d_padj = (dev_adjust && E > 0 && P > 0 && n_nodes > 0 && out->padj)
                             ? (dev_out ? out->padj : s.padj.as<double>())
                             : nullptr;  // (an internal scratch copy used only for the best row is never emitted)
  const bool want_cat_best = out->cat_best && E > 0 && in->padj != MDEGGS_PADJ_HOST;
  const bool want_padj_best = out->padj_best && E > 0 && dev_adjust && n_nodes > 0 && P > 0;
  const bool is_rlm = K == 2 && in->method == MDEGGS_METHOD_RLM;
  const char* blk = s.small_out.as<char>();
  const size_t blk_bytes = small_block_bytes(P);
  struct Item {
    void* dst;
    const void* src;
    size_t bytes;
    int kind;
  };
  std::vector<Item> big, small;
  auto add = [](std::vector<Item>& v, void* dst, const void* src, size_t bytes, int kind = EMIT_COPY) {
    if (dst && bytes) v.push_back({dst, src, bytes, kind});
  };
  const bool rows_here = dev_out && s.best_rows_in_emit;
  add(small, out->tot_edges, blk, (size_t)P * 8);
  add(small, out->sig_count, blk + (size_t)Pn * 8, (size_t)P * 8);
  add(small, out->fail_count, blk + (size_t)2 * Pn * 8, (size_t)P * 8);
  add(small, out->cutoff, blk + (size_t)3 * Pn * 8, (size_t)P * 8);
  add(small, out->best, blk + (size_t)4 * Pn * 8, 4);
  add(small, out->df, blk + (size_t)4 * Pn * 8 + 16, 16);
  add(big, out->median, s.med_out.p, (size_t)G * K * 8);
  add(big, out->p_edge, s.p_edge.p, (size_t)E * 8);
  add(big, out->stat, s.stat.p, (size_t)E * 8);
  add(big, out->coef, s.coef.p, (size_t)E * 2 * K * 8);
  add(big, out->status, s.status.p, (size_t)E * 4);
  add(big, out->iters, is_rlm ? s.iters.p : nullptr, (size_t)E * 4);  // null source: zeros
  if (d_cat && d_cat != out->cat) add(big, out->cat, d_cat, (size_t)P * E);
  if (d_padj && d_padj != out->padj) add(big, out->padj, d_padj, (size_t)P * E * 8);
  if (want_cat_best) add(big, out->cat_best, s.cat_best.p, (size_t)E, rows_here ? EMIT_CAT_BEST : EMIT_COPY);
  if (want_padj_best) {
    if (rows_here && s.best_row_src) add(big, out->padj_best, s.best_row_src, (size_t)E * 8, EMIT_PADJ_BEST);
    else add(big, out->padj_best, s.padj_best.p, (size_t)E * 8);
  }
  if (dev_out) {
    EmitTab T;
    auto reset = [&]() {
      memset(&T, 0, sizeof T);
      T.ea = s.ea.as<int32_t>();
      T.eb = s.eb.as<int32_t>();
      T.best = s.scal.as<int32_t>() + 2 * SC_BEST;
      T.err = s.scal.as<int32_t>() + 2 * SC_ERR;
      T.act = s.act.as<uint8_t>();
      T.act_stride = n_nodes > 0 ? n_nodes : 1;
      T.E = (long long)E;
    };
    reset();
    int ne = 0;
    size_t max_bytes = 0;
    auto flush = [&]() {
      if (ne == 0) return;
      const unsigned gx = (unsigned)((max_bytes / 16 + 255) / 256);
      dim3 grid(gx < 1 ? 1 : (gx > 148 * 8 ? 148 * 8 : gx), (unsigned)ne);
      LAUNCH(k_emit, grid, 256, 0, s.s_main, T);
      reset();
      ne = 0;
      max_bytes = 0;
    };
    for (const std::vector<Item>* v : {&small, &big})
      for (const Item& it : *v) {
        T.dst[ne] = it.dst;
        T.src[ne] = it.src;
        T.bytes[ne] = (long long)it.bytes;
        T.kind[ne] = it.kind;
        max_bytes = it.bytes > max_bytes ? it.bytes : max_bytes;
        if (++ne == EMIT_MAX) flush();
      }
    flush();
  } else {
    for (const Item& it : big) {
      if (it.src) {
        CU(cudaMemcpyAsync(it.dst, it.src, it.bytes, cudaMemcpyDeviceToHost, s.s_main));
        ctx->stats.d2h_bytes += (int64_t)it.bytes;
      } else {
        memset(it.dst, 0, it.bytes);
      }
    }
  }
  // the packed block was also written straight into pinned host memory by k_select_best (mapped, no copy): node
  // count + error flag are read from it after the final sync, and host callers get their small results from it
  // (finish_small_outputs)
  if (!dev_out) ctx->stats.d2h_bytes += (int64_t)blk_bytes;
  if (record_end) CU(cudaEventRecord(s.ev[EV_END], s.s_main));
  trace_mark(ctx, "(outputs copied)", s.s_main);
  return MDEGGS_OK;
}

// after the final synchronisation: small results from the pinned copy of the packed block -> host caller buffers;
// returns the node count and the error flag
void finish_small_outputs(Shard& s, const mdeggs_problem* in, const mdeggs_result* out, int32_t* n_nodes,
                          int32_t* err_flag) {
  const int P = in->P, Pn = P > 0 ? P : 1;
  const size_t blk_bytes = small_block_bytes(P);
  const char* blk = (const char*)s.h_stage + s.h_stage_cap - 64 - blk_bytes;
  const int32_t* tail = reinterpret_cast<const int32_t*>(blk + (size_t)4 * Pn * 8);
  *n_nodes = tail[1];
  *err_flag = tail[2];
  if (in->mem != MDEGGS_MEM_HOST) return;
  if (out->tot_edges) memcpy(out->tot_edges, blk, (size_t)P * 8);
  if (out->sig_count) memcpy(out->sig_count, blk + (size_t)Pn * 8, (size_t)P * 8);
  if (out->fail_count) memcpy(out->fail_count, blk + (size_t)2 * Pn * 8, (size_t)P * 8);
  if (out->cutoff) memcpy(out->cutoff, blk + (size_t)3 * Pn * 8, (size_t)P * 8);
  if (out->best) memcpy(out->best, tail, 4);
  if (out->df) memcpy(out->df, tail + 4, 16);
}

float ev_ms(cudaEvent_t a, cudaEvent_t b) {
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, a, b) != cudaSuccess) return 0.f;
  return ms;
}

}  // namespace

// =================================================================================================
// extern "C" ABI
// =================================================================================================
extern "C" {

int mdeggs_abi_version(void) { return MDEGGS_ABI_VERSION; }

const char* mdeggs_strerror(int code) {
  switch (code) {
    case MDEGGS_OK: return "ok";
    case MDEGGS_EINVAL: return "invalid argument";
    case MDEGGS_ECUDA: return "CUDA error";
    case MDEGGS_ENCCL: return "NCCL error";
    case MDEGGS_ENOMEM: return "out of device memory";
    case MDEGGS_EUNSUPPORTED: return "unsupported configuration";
    default: return "unknown error";
  }
}

const char* mdeggs_last_error(const mdeggs_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int mdeggs_create(mdeggs_ctx** out, const int* device_ids, int n_dev) {
  if (!out || n_dev < 1 || !device_ids) return MDEGGS_EINVAL;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count < 1) return MDEGGS_ECUDA;
  for (int i = 0; i < n_dev; ++i)
    if (device_ids[i] < 0 || device_ids[i] >= count) return MDEGGS_EINVAL;
  mdeggs_ctx* ctx = new mdeggs_ctx();
  memset(&ctx->stats, 0, sizeof ctx->stats);
  if (const char* fp = getenv("MDEGGS_FIT_PATH")) ctx->fit_path_override = atoi(fp);
  if (const char* g = getenv("MDEGGS_GRAPH")) ctx->use_graph = atoi(g);
  if (const char* g = getenv("MDEGGS_DEBUG_SYNC")) ctx->debug_sync = atoi(g);
  if (const char* g = getenv("MDEGGS_TRACE")) ctx->trace = atoi(g);
  if (const char* g = getenv("MDEGGS_FRONT_PATH")) ctx->front_path = atoi(g);
  ctx->shards.resize((size_t)n_dev);
  for (int i = 0; i < n_dev; ++i) {
    ctx->shards[(size_t)i].device = device_ids[i];
    ctx->shards[(size_t)i].rank = i;
    ctx->shards[(size_t)i].world = n_dev;
    int rc = init_shard(ctx, ctx->shards[(size_t)i]);
    if (rc) {
      delete ctx;
      return rc;
    }
  }
  if (n_dev > 1) {
    NcclApi& api = nccl_api();
    if (!api.load()) {
      delete ctx;
      return MDEGGS_ENCCL;
    }
    std::vector<nccl_comm_t> comms((size_t)n_dev);
    if (api.CommInitAll(comms.data(), n_dev, device_ids) != 0) {
      delete ctx;
      return MDEGGS_ENCCL;
    }
    for (int i = 0; i < n_dev; ++i) ctx->shards[(size_t)i].comm = comms[(size_t)i];
  }
  *out = ctx;
  return MDEGGS_OK;
}

int mdeggs_nccl_unique_id(void* out128) {
  NcclApi& api = nccl_api();
  if (!out128 || !api.load()) return MDEGGS_ENCCL;
  nccl_unique_id id;
  if (api.GetUniqueId(&id) != 0) return MDEGGS_ENCCL;
  memcpy(out128, &id, 128);
  return MDEGGS_OK;
}

int mdeggs_create_rank(mdeggs_ctx** out, int device_id, int rank, int world, const void* unique_id) {
  if (!out || world < 1 || rank < 0 || rank >= world) return MDEGGS_EINVAL;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count < 1) return MDEGGS_ECUDA;
  if (device_id < 0 || device_id >= count) return MDEGGS_EINVAL;
  mdeggs_ctx* ctx = new mdeggs_ctx();
  memset(&ctx->stats, 0, sizeof ctx->stats);
  if (const char* fp = getenv("MDEGGS_FIT_PATH")) ctx->fit_path_override = atoi(fp);
  if (const char* g = getenv("MDEGGS_GRAPH")) ctx->use_graph = atoi(g);
  if (const char* g = getenv("MDEGGS_DEBUG_SYNC")) ctx->debug_sync = atoi(g);
  if (const char* g = getenv("MDEGGS_TRACE")) ctx->trace = atoi(g);
  if (const char* g = getenv("MDEGGS_FRONT_PATH")) ctx->front_path = atoi(g);
  ctx->shards.resize(1);
  Shard& s = ctx->shards[0];
  s.device = device_id;
  s.rank = rank;
  s.world = world;
  int rc = init_shard(ctx, s);
  if (rc) {
    delete ctx;
    return rc;
  }
  if (world > 1) {
    NcclApi& api = nccl_api();
    if (!unique_id || !api.load()) {
      delete ctx;
      return MDEGGS_ENCCL;
    }
    nccl_unique_id id;
    memcpy(&id, unique_id, 128);
    if (api.CommInitRank(&s.comm, world, id, rank) != 0) {
      delete ctx;
      return MDEGGS_ENCCL;
    }
  }
  *out = ctx;
  return MDEGGS_OK;
}

int mdeggs_destroy(mdeggs_ctx* ctx) {
  if (!ctx) return MDEGGS_OK;
  if (ctx->graph_exec) cudaGraphExecDestroy(ctx->graph_exec);
  if (ctx->pack.h_buf) cudaFreeHost(ctx->pack.h_buf);
  for (cudaEvent_t e : ctx->trace_pool) cudaEventDestroy(e);
  for (Shard& s : ctx->shards) {
    cudaSetDevice(s.device);
    cudaDeviceSynchronize();
    if (s.h_stage) cudaFreeHost(s.h_stage);
    for (int r = 0; r < P2P_MAX && ctx->shards.size() == 1; ++r)  // (IPC mappings exist with one process per GPU only)
      if (r != s.rank) {
        if (s.p2p.peer_p[r]) cudaIpcCloseMemHandle(s.p2p.peer_p[r]);
        if (s.p2p.peer_st[r]) cudaIpcCloseMemHandle(s.p2p.peer_st[r]);
        if (s.p2p.peer_flags[r]) cudaIpcCloseMemHandle(s.p2p.peer_flags[r]);
      }
    cudaGetLastError();
    if (s.comm) nccl_api().CommDestroy(s.comm);
    s.release_all();
    for (int i = 0; i < EV_COUNT; ++i)
      if (s.ev[i]) cudaEventDestroy(s.ev[i]);
    if (s.ev_fork) cudaEventDestroy(s.ev_fork);
    if (s.ev_join) cudaEventDestroy(s.ev_join);
    if (s.ev_sample) cudaEventDestroy(s.ev_sample);
    if (s.ev_cf) cudaEventDestroy(s.ev_cf);
    if (s.s_main) cudaStreamDestroy(s.s_main);
    if (s.s_sort) cudaStreamDestroy(s.s_sort);
    if (s.s_copy) cudaStreamDestroy(s.s_copy);
    if (s.s_copy2) cudaStreamDestroy(s.s_copy2);
  }
  delete ctx;
  return MDEGGS_OK;
}

int mdeggs_run_omic(mdeggs_ctx* ctx, const mdeggs_problem* in, mdeggs_result* out) {
  int rc = mdeggs_submit_omic(ctx, in, out);
  if (rc) return rc;
  return mdeggs_wait(ctx);
}

int mdeggs_run_omics(mdeggs_ctx* const* ctxs, int n_layers, const mdeggs_problem* in, mdeggs_result* out,
                     int* layer_rc) {
  if (n_layers < 0 || (n_layers > 0 && (!ctxs || !in || !out))) return MDEGGS_EINVAL;
  for (int i = 0; i < n_layers; ++i)
    for (int j = 0; j < i; ++j)
      if (ctxs[i] == ctxs[j]) return MDEGGS_EINVAL;  // one context (streams, scratch, staging) per layer in flight
  int first = MDEGGS_OK;
  std::vector<int> rcs((size_t)n_layers, MDEGGS_OK);
  for (int i = 0; i < n_layers; ++i) rcs[i] = mdeggs_submit_omic(ctxs[i], in + i, out + i);
  for (int i = 0; i < n_layers; ++i) {
    if (rcs[i] == MDEGGS_OK) rcs[i] = mdeggs_wait(ctxs[i]);
    if (rcs[i] != MDEGGS_OK && first == MDEGGS_OK) first = rcs[i];
    if (layer_rc) layer_rc[i] = rcs[i];
  }
  return first;
}

int mdeggs_submit_omic(mdeggs_ctx* ctx, const mdeggs_problem* in, mdeggs_result* out) {
  if (!ctx || !in || !out) return MDEGGS_EINVAL;
  if (ctx->pend.active) {
    set_err(ctx, "a submitted call is still pending: call mdeggs_wait first");
    return MDEGGS_EINVAL;
  }
  ctx->err.clear();
  if (in->K < 2 || in->K > MDEGGS_MAX_K) {
    set_err(ctx, "K = %d outside 2..%d", in->K, MDEGGS_MAX_K);
    return MDEGGS_EINVAL;
  }
  if (in->G < 1 || in->n < 1 || in->E < 0 || in->P < 0 || !in->assay || !in->group || (in->P > 0 && !in->pct) ||
      (in->E > 0 && (!in->from || !in->to)) || in->E > 2147483647LL) {
    set_err(ctx, "bad problem dimensions or null pointer");
    return MDEGGS_EINVAL;
  }
  if ((in->layout == MDEGGS_LAYOUT_COLMAJOR && in->ld < in->G) || (in->layout == MDEGGS_LAYOUT_ROWMAJOR && in->ld < in->n) ||
      (in->layout != MDEGGS_LAYOUT_COLMAJOR && in->layout != MDEGGS_LAYOUT_ROWMAJOR)) {
    set_err(ctx, "bad layout / leading dimension");
    return MDEGGS_EINVAL;
  }
  if (in->K == 2 && in->method == MDEGGS_METHOD_RLM && (in->tested_coef < 1 || in->tested_coef > 4)) {
    set_err(ctx, "tested_coef must be 1..4");
    return MDEGGS_EINVAL;
  }
  if (2 * in->P >= QS_MAX_T) {  // slot ids travel as bytes, 0xff = none
    set_err(ctx, "more than %d percentiles are not supported", QS_MAX_T / 2 - 1);
    return MDEGGS_EUNSUPPORTED;
  }
  if (in->padj == MDEGGS_PADJ_HOMMEL && in->E > MDEGGS_HOMMEL_MAX_E) {
    set_err(ctx, "hommel on the device is limited to %d edges (O(n^2) per segment): use MDEGGS_PADJ_HOST", MDEGGS_HOMMEL_MAX_E);
    return MDEGGS_EUNSUPPORTED;
  }
  if (in->n <= 2 * in->K) {
    set_err(ctx, "n = %d leaves no residual degrees of freedom for K = %d", in->n, in->K);
    return MDEGGS_EINVAL;
  }
  const auto t_host0 = std::chrono::steady_clock::now();
  int rc = plan_host_pack(ctx, in);
  if (rc) return rc;
  if (ctx->pack.active) {  // from here on the call runs on the packed problem (mdeggs_hostpack.h)
    mdeggs_ctx::HostPack& K = ctx->pack;
    K.user_median = out->median;
    K.user_G = in->G;
    in = &K.in2;
    if (out->median) {  // the only output indexed by matrix row: produced for the packed rows, spread by mdeggs_wait
      K.med_tmp.resize((size_t)in->K * (size_t)in->G);
      K.out2 = *out;
      K.out2.median = K.med_tmp.data();
      out = &K.out2;
    }
  }
  HostLayout H;
  rc = build_layout(ctx, in, H);
  if (rc) return rc;
  ctx->launches = 0;
  memset(&ctx->stats, 0, sizeof ctx->stats);
  ctx->trace_recs.clear();
  ctx->trace_used = 0;
  const int src_device = ctx->shards[0].device;
  Shard& sh0 = ctx->shards[0];
  for (Shard& s : ctx->shards) s.p2p.used = false;
  // ---- phase A: inputs (eager) ----
  for (Shard& s : ctx->shards) {
    rc = stage_inputs(ctx, s, in, H, src_device);
    if (rc) return rc;
  }
  const auto t_hostA = std::chrono::steady_clock::now();
  if (getenv("MDEGGS_HOSTTIME") && sh0.stamps.p) LAUNCH(k_stamp, 1, 1, 0, sh0.s_main, sh0.stamps.as<unsigned long long>() + 2);
  // ---- phase B: compute, replayed from a CUDA graph when this exact call shape repeats ----
  auto enqueue_compute = [&](int32_t* n_cap) -> int {
    int r = MDEGGS_OK;
    for (Shard& s : ctx->shards)
      if ((r = run_shard_front(ctx, s, in, n_cap))) return r;
    for (Shard& s : ctx->shards)
      if ((r = run_shard_main(ctx, s, in, out, H, *n_cap))) return r;
    if ((r = run_collective(ctx, in, out))) return r;
    for (size_t i = 0; i < ctx->shards.size(); ++i)
      if ((r = run_shard_back(ctx, ctx->shards[i], in, out, H, *n_cap, i == 0))) return r;
    if ((r = run_sig_allreduce(ctx, in, out))) return r;
    for (size_t i = 0; i < ctx->shards.size(); ++i)
      if ((r = run_shard_back2(ctx, ctx->shards[i], in, out, H, *n_cap, i == 0))) return r;
    // device callers: the output kernel is part of the captured graph (host callers get eager copies in phase C)
    if (ctx->capturing && in->mem == MDEGGS_MEM_DEVICE) r = emit_outputs(ctx, ctx->shards[0], in, out, H, *n_cap, false);
    return r;
  };
  int32_t n_nodes = (int32_t)(2 * in->E < (int64_t)in->G ? 2 * in->E : (int64_t)in->G);
  // (one process per GPU: the NCCL collectives of the call are captured with it, every rank captures at the same call)
  bool graph_ok = ctx->use_graph && !ctx->trace && ctx->shards.size() == 1;
  bool replayed = false;
  if (graph_ok) {
    GraphKey key;
    memset(&key, 0, sizeof key);
    key.in = *in;
    key.in.group = nullptr;
    key.in.pct = nullptr;
    key.out = *out;
    key.L = H.L;
    key.d_assay = sh0.assay.p;
    key.d_from = sh0.from.p;
    key.d_to = sh0.to.p;
    key.row_off = sh0.row_off;
    key.ld_eff = sh0.ld_eff;
    key.gate_rows = sh0.gated ? sh0.gate_rows : 0;
    key.dense_cap = sh0.dense_S.cap;
    key.h_stage = sh0.h_stage;
    key.h_stage_cap = sh0.h_stage_cap;
    if (in->mem == MDEGGS_MEM_HOST) {  // host buffers are only touched by the eager phases A and C
      key.in.assay = nullptr;
      key.in.from = key.in.to = nullptr;
      memset(&key.out, 0, sizeof key.out);
      const void* const* op = reinterpret_cast<const void* const*>(out);
      void** kp = reinterpret_cast<void**>(&key.out);
      for (size_t i = 0; i < sizeof(mdeggs_result) / sizeof(void*); ++i) kp[i] = op[i] ? (void*)1 : nullptr;
    }
    if (ctx->key_valid && memcmp(&key, &ctx->key, sizeof key) == 0) {
      ++ctx->key_hits;
    } else {
      if (ctx->graph_exec) {
        cudaGraphExecDestroy(ctx->graph_exec);
        ctx->graph_exec = nullptr;
      }
      ctx->key = key;
      ctx->key_valid = true;
      ctx->key_hits = 0;
    }
    if (ctx->key_hits >= 1 && !ctx->graph_exec) {  // second identical call: every buffer already has its size
      cudaGraph_t graph = nullptr;
      CU(cudaSetDevice(sh0.device));
      CU(cudaStreamBeginCapture(sh0.s_main, cudaStreamCaptureModeRelaxed));
      ctx->capturing = true;
      ctx->launches = 0;
      int32_t cap = 0;
      int r = enqueue_compute(&cap);
      ctx->capturing = false;
      cudaError_t ce = cudaStreamEndCapture(sh0.s_main, &graph);
      if (r != MDEGGS_OK || ce != cudaSuccess || !graph) {
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        ctx->use_graph = 0;  // capture not possible in this configuration: stay eager from now on
        graph_ok = false;
      } else {
        ce = cudaGraphInstantiate(&ctx->graph_exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ce != cudaSuccess) {
          ctx->graph_exec = nullptr;
          cudaGetLastError();
          ctx->use_graph = 0;
          graph_ok = false;
        } else {
          ctx->graph_launches = ctx->launches;
          ctx->graph_fit_path = ctx->stats.fit_path;
          ctx->graph_p2p = sh0.p2p.used;
        }
      }
    }
    if (graph_ok && ctx->graph_exec) {
      if (ctx->pack.deferred && !launches_are_async(ctx, sh0)) {  // (a serialising profiler, CUDA_LAUNCH_BLOCKING)
        ctx->pack.deferred = false;
        rc = run_host_pack(ctx, sh0, true);
        if (rc) return rc;
      }
      CU(cudaGraphLaunch(ctx->graph_exec, sh0.s_main));
      CU(cudaEventRecord(sh0.ev[EV_ADJ], sh0.s_main));
      ctx->launches = ctx->graph_launches;
      ctx->stats.fit_path = ctx->graph_fit_path;
      sh0.p2p.used = ctx->graph_p2p;
      replayed = true;
    }
  }
  // Streamed host packing.  A replayed graph is ONE launch with nothing after it that could block this thread, so the
  // packing runs behind it: the kernels wait on the gates and the front kernel works on the chunks as they land.
  // Eager launches may block half-way (a growing scratch buffer is a cudaFree, which waits for the device - and the
  // device would wait for the gates): there the chunks are packed and sent first, overlapping each other only.
  if (!replayed) {
    if (ctx->pack.deferred) {
      ctx->pack.deferred = false;
      rc = run_host_pack(ctx, sh0, true);
      if (rc) return rc;
    }
    ctx->launches = 0;
    rc = enqueue_compute(&n_nodes);
    if (rc) return rc;
  }
  if (ctx->pack.deferred) {
    rc = run_host_pack(ctx, sh0, true);
    if (rc) return rc;
  }
  for (Shard& s : ctx->shards)  // a streamed input copy ends inside the compute phase: the call is not over before it
    if (s.gated) {
      CU(cudaSetDevice(s.device));
      CU(cudaStreamWaitEvent(s.s_main, s.ev_copy_done, 0));
    }
  CU(cudaSetDevice(sh0.device));
  const auto t_host1 = std::chrono::steady_clock::now();
  // ---- phase C: outputs (eager; device callers of a replayed graph already have them) ----
  if (replayed && in->mem == MDEGGS_MEM_DEVICE) {
    CU(cudaEventRecord(sh0.ev[EV_END], sh0.s_main));
  } else {
    rc = emit_outputs(ctx, sh0, in, out, H, n_nodes);
    if (rc) return rc;
  }
  // peer-memory all-gather: this rank no longer reads the gathered values (its outputs are out): the peers' next push
  // may overwrite them
  for (Shard& s : ctx->shards)
    if (s.p2p.used) {
      CU(cudaSetDevice(s.device));
      LAUNCH(k_signal_done, 1, 32, 0, s.s_main, s.p2p.args);
    }
  CU(cudaSetDevice(sh0.device));
  if (getenv("MDEGGS_HOSTTIME") && sh0.stamps.p) LAUNCH(k_stamp, 1, 1, 0, sh0.s_main, sh0.stamps.as<unsigned long long>() + 3);
  const auto t_host2 = std::chrono::steady_clock::now();
  for (size_t i = 1; i < ctx->shards.size(); ++i) {
    CU(cudaSetDevice(ctx->shards[i].device));
    CU(cudaEventRecord(ctx->shards[i].ev[EV_END], ctx->shards[i].s_main));
  }
  ctx->pend.in = *in;
  ctx->pend.out = *out;
  ctx->pend.n_nodes = n_nodes;
  ctx->pend.L = H.L;
  ctx->pend.replayed = replayed;
  ctx->last_valid = false;
  ctx->pend.src_device = src_device;
  ctx->pend.t0 = t_host0;
  ctx->pend.t1 = t_host1;
  ctx->pend.t2 = t_host2;
  ctx->pend.tA = t_hostA;
  ctx->pend.active = true;
  return MDEGGS_OK;
}

int mdeggs_wait(mdeggs_ctx* ctx) {
  if (!ctx) return MDEGGS_EINVAL;
  if (!ctx->pend.active) {
    set_err(ctx, "mdeggs_wait without a submitted call");
    return MDEGGS_EINVAL;
  }
  ctx->pend.active = false;
  const mdeggs_problem* in = &ctx->pend.in;
  mdeggs_result* out = &ctx->pend.out;
  int32_t n_nodes = ctx->pend.n_nodes;
  const bool replayed = ctx->pend.replayed;
  const int src_device = ctx->pend.src_device;
  const auto t_host0 = ctx->pend.t0, t_host1 = ctx->pend.t1, t_host2 = ctx->pend.t2;
  for (Shard& s : ctx->shards) {
    CU(cudaSetDevice(s.device));
    CU(cudaStreamSynchronize(s.s_sort));
    CU(cudaStreamSynchronize(s.s_main));
  }
  CU(cudaSetDevice(src_device));
  CU(cudaGetLastError());  // kernel launches are not checked one by one: surface any launch failure here
  Shard& s0 = ctx->shards[0];
  {
    int32_t err_flag = 0;
    finish_small_outputs(s0, in, out, &n_nodes, &err_flag);
    if (err_flag == P2P_ERR) {
      set_err(ctx, "a peer's p-value block did not arrive (peer-memory all-gather; option p2p = 0 selects NCCL)");
      return MDEGGS_ECUDA;
    }
    if (err_flag == GATE_ERR) {
      set_err(ctx, "the streamed input copy did not arrive (option h2d_pipeline = 0 disables it)");
      return MDEGGS_ECUDA;
    }
    if (err_flag != 0) {
      set_err(ctx, "from/to contain an index outside 1..G");
      return MDEGGS_EINVAL;
    }
  }
  if (ctx->trace == 2 || getenv("MDEGGS_HOSTTIME")) {
    const auto t_host3 = std::chrono::steady_clock::now();
    auto us = [](auto a, auto b) { return std::chrono::duration<double, std::micro>(b - a).count(); };
    unsigned long long st2[4] = {0, 0, 0, 0};
    if (s0.stamps.p) cudaMemcpy(st2, s0.stamps.p, 32, cudaMemcpyDeviceToHost);
    fprintf(stderr, "[mdeggs host] GPU: pre-stamp -> first node %.1f us, compute %.1f us, last node -> after emit %.1f us\n",
            (double)(st2[0] - st2[2]) * 1e-3, (double)(st2[1] - st2[0]) * 1e-3, (double)(st2[3] - st2[1]) * 1e-3);
    fprintf(stderr, "[mdeggs host] layout+staging %.1f us of the prep; ", us(t_host0, ctx->pend.tA));
    fprintf(stderr, "[mdeggs host] prep+enqueue %.1f us, emit enqueue %.1f us, sync+finish %.1f us, total %.1f us; "
            "GPU compute span %.1f us (%s)\n",
            us(t_host0, t_host1), us(t_host1, t_host2), us(t_host2, t_host3), us(t_host0, t_host3),
            (double)(st2[1] - st2[0]) * 1e-3, replayed ? "graph" : "eager");
  }
  if (ctx->trace && ctx->trace != 3 && !ctx->trace_recs.empty()) {  // timeline of this call: end time of every launch, per stream
    float prev[2] = {0.f, 0.f};
    fprintf(stderr, "[mdeggs trace] %-34s %-6s %10s %10s\n", "launch", "stream", "end_us", "delta_us");
    for (const auto& r : ctx->trace_recs) {
      const int sid = r.stream == s0.s_sort ? 1 : 0;
      const float t = ev_ms(s0.ev[EV_START], r.ev) * 1e3f;
      fprintf(stderr, "[mdeggs trace] %-34s %-6s %10.1f %10.1f\n", r.name, sid ? "side" : "main", t, t - prev[sid]);
      prev[sid] = t;
    }
  }
  if (ctx->pack.active && ctx->pack.user_median) {
    const mdeggs_ctx::HostPack& K = ctx->pack;
    const size_t Gu = (size_t)K.user_G, nn = K.rows.size();
    for (int k = 0; k < in->K; ++k) {
      double* dst = K.user_median + (size_t)k * Gu;
      for (size_t g = 0; g < Gu; ++g) dst[g] = nan("");
      for (size_t i = 0; i < nn; ++i) dst[(size_t)K.rows[i]] = K.med_tmp[(size_t)k * nn + i];
    }
  }
  ctx->last_valid = n_nodes > 0;
  ctx->pack.last_packed = ctx->pack.active;
  ctx->last_L = ctx->pend.L;
  ctx->last_G = in->G;
  // ---- stats ----
  (void)replayed;
  mdeggs_stats& st = ctx->stats;
  st.n_nodes = n_nodes;
  if (ctx->last_zc) st.h2d_bytes += (int64_t)n_nodes * in->n * 8;  // node rows read over PCIe by the gather kernel
  st.unique_fits = in->E;
  st.kernel_launches = ctx->launches;
  st.n_ranks = s0.world;
  st.fit_bytes = in->E * (16LL * in->n + 16);
  st.ms_total = ev_ms(s0.ev[EV_START], s0.ev[EV_END]);
  st.ms_h2d = ev_ms(s0.ev[EV_START], s0.ev[EV_H2D]);
  st.ms_d2h = ev_ms(s0.ev[EV_ADJ], s0.ev[EV_END]);
  if (ctx->pack.active) st.fit_path |= 0x2000;  // bit 13: the call ran on host-packed node rows (mdeggs_hostpack.h)
  if (s0.p2p.used) st.fit_path |= 0x4000;       // bit 14: the p-value blocks were pushed over peer memory (mdeggs_p2p.cuh)
  if (use_bucket_order(ctx, in, n_nodes)) st.fit_path |= 0x200;  // bit 9: K8 ordering by p-value buckets (no radix sort)
  if (replayed) {  // graph replay: no stage boundaries inside the graph; everything between H2D and D2H is compute
    st.fit_path |= 0x100;  // bit 8 flags "replayed from a CUDA graph"
  } else {
    st.ms_gather = ev_ms(s0.ev[EV_H2D], s0.ev[EV_GATHER]);
    st.ms_stats = ev_ms(s0.ev[EV_GATHER], s0.ev[EV_STATS]);
    st.ms_quantile = ev_ms(s0.ev[EV_Q0], s0.ev[EV_Q1]);
    st.ms_fit = ev_ms(s0.ev[EV_STATS], s0.ev[EV_FIT]);
    st.ms_tail = ev_ms(s0.ev[EV_FIT], s0.ev[EV_TAIL]);
    st.ms_collective = ev_ms(s0.ev[EV_TAIL], s0.ev[EV_COLL]);
    st.ms_percolate = ev_ms(s0.ev[EV_COLL], s0.ev[EV_PERC]);
    st.ms_adjust = ev_ms(s0.ev[EV_PERC], s0.ev[EV_ADJ]);
  }
  if (out->tot_edges && in->mem == MDEGGS_MEM_HOST)
    for (int p = 0; p < in->P; ++p) st.ref_equiv_tests += out->tot_edges[p];
  return MDEGGS_OK;
}

int mdeggs_set_option(mdeggs_ctx* ctx, const char* name, int64_t value) {
  if (!ctx || !name) return MDEGGS_EINVAL;
  if (!strcmp(name, "graph")) {
    ctx->use_graph = (int)value;
    if (!value && ctx->graph_exec) {
      cudaGraphExecDestroy(ctx->graph_exec);
      ctx->graph_exec = nullptr;
    }
    ctx->key_valid = false;
    return MDEGGS_OK;
  }
  if (!strcmp(name, "trace")) {
    ctx->trace = (int)value;
    return MDEGGS_OK;
  }
  if (!strcmp(name, "rsel_cap")) {
    ctx->rsel_cap_override = value;
    ctx->key_valid = false;
    return MDEGGS_OK;
  }
  if (!strcmp(name, "h2d_pipeline")) {  // 0 off, 1 on (12 MB chunks), > 1: chunk size in MB
    ctx->h2d_pipeline = (int)value;
    ctx->key_valid = false;  // (a captured compute graph has the choice baked in)
    return MDEGGS_OK;
  }
  if (!strcmp(name, "p2p")) {
    ctx->p2p = (int)value;
    ctx->key_valid = false;
    return MDEGGS_OK;
  }
  if (!strcmp(name, "tail_spline")) {
    ctx->tail_spline = (int)value;
    ctx->key_valid = false;
    return MDEGGS_OK;
  }
  if (!strcmp(name, "host_pack")) {
    ctx->host_pack = (int)value;
    ctx->key_valid = false;
    return MDEGGS_OK;
  }
  if (!strcmp(name, "copy_streams")) {
    ctx->copy_streams = (int)value;
    return MDEGGS_OK;
  }
  if (!strcmp(name, "zero_copy")) {
    ctx->zero_copy = (int)value;
    ctx->key_valid = false;
    return MDEGGS_OK;
  }
  if (!strcmp(name, "padj_keep_cap")) {
    ctx->padj_keep_cap = value < 0 ? 0 : value;
    ctx->key_valid = false;
    return MDEGGS_OK;
  }
  if (!strcmp(name, "count_skip")) {
    ctx->count_skip = (int)value;
    ctx->key_valid = false;
    return MDEGGS_OK;
  }
  if (!strcmp(name, "count_only")) {
    ctx->count_only_adjust = (int)value;
    ctx->key_valid = false;
    return MDEGGS_OK;
  }
  if (!strcmp(name, "sort_path")) {
    ctx->sort_path_override = (int)value;
    ctx->key_valid = false;
    return MDEGGS_OK;
  }
  if (!strcmp(name, "bh_onepass")) {
    ctx->bh_onepass = (int)value;
    ctx->key_valid = false;
    return MDEGGS_OK;
  }
  if (!strcmp(name, "order_filter")) {
    ctx->order_filter = (int)value;
    ctx->key_valid = false;
    return MDEGGS_OK;
  }
  if (!strcmp(name, "dense_path")) {
    ctx->dense_path = (int)value;
    ctx->key_valid = false;
    return MDEGGS_OK;
  }
  if (!strcmp(name, "front_path")) {
    ctx->front_path = (int)value;
    ctx->key_valid = false;
    return MDEGGS_OK;
  }
  if (!strcmp(name, "fit_path")) {
    ctx->fit_path_override = (int)value;
    ctx->key_valid = false;
    return MDEGGS_OK;
  }
  return MDEGGS_EINVAL;
}

int mdeggs_trace_count(const mdeggs_ctx* ctx) { return ctx ? (int)ctx->trace_recs.size() : 0; }

int mdeggs_trace_get(const mdeggs_ctx* ctx, int i, char* name, int name_cap, int* stream_id, float* end_us) {
  if (!ctx || i < 0 || (size_t)i >= ctx->trace_recs.size() || !name || name_cap < 1 || !stream_id || !end_us)
    return MDEGGS_EINVAL;
  const auto& r = ctx->trace_recs[(size_t)i];
  const Shard& s0 = ctx->shards[0];
  snprintf(name, (size_t)name_cap, "%s", r.name);
  *stream_id = r.stream == s0.s_sort ? 1 : 0;
  *end_us = ev_ms(s0.ev[EV_START], r.ev) * 1e3f;
  return MDEGGS_OK;
}

int mdeggs_last_stats(const mdeggs_ctx* ctx, mdeggs_stats* out) {
  if (!ctx || !out) return MDEGGS_EINVAL;
  *out = ctx->stats;
  return MDEGGS_OK;
}

int mdeggs_pair_features(mdeggs_ctx* ctx, const double* x, int64_t ldx, int32_t n, const int32_t* a, const int32_t* b,
                         int32_t n_pairs, int32_t ratio, int32_t mem, double* out) {
  if (!ctx || !x || !out || n < 0 || n_pairs < 0 || ldx < n) return MDEGGS_EINVAL;
  if (n == 0 || n_pairs == 0) return MDEGGS_OK;
  Shard& s = ctx->shards[0];
  CU(cudaSetDevice(s.device));
  const size_t total = (size_t)n * n_pairs;
  if (mem == MDEGGS_MEM_DEVICE) {
    LAUNCH(k_pair_features, blocks_for((int64_t)total, 256), 256, 0, s.s_main, x, ldx, n, a, b, n_pairs, ratio, out);
    CU(cudaStreamSynchronize(s.s_main));
    return MDEGGS_OK;
  }
  // host: only the referenced columns cross PCIe
  std::vector<int32_t> cols;
  std::vector<int32_t> ra((size_t)n_pairs), rb((size_t)n_pairs);
  {
    std::vector<std::pair<int32_t, int32_t>> seen;
    auto remap = [&](int32_t c) {
      for (auto& pr : seen)
        if (pr.first == c) return pr.second;
      int32_t id = (int32_t)cols.size();
      cols.push_back(c);
      seen.push_back({c, id});
      return id;
    };
    for (int j = 0; j < n_pairs; ++j) {
      if (a[j] < 0 || b[j] < 0) return MDEGGS_EINVAL;
      ra[(size_t)j] = remap(a[j]);
      rb[(size_t)j] = remap(b[j]);
    }
  }
  DevBuf dx, da, db, dout;
  int rc = MDEGGS_OK;
  do {
    if (dx.ensure(cols.size() * (size_t)n * 8) != cudaSuccess || da.ensure((size_t)n_pairs * 4) != cudaSuccess ||
        db.ensure((size_t)n_pairs * 4) != cudaSuccess || dout.ensure(total * 8) != cudaSuccess) {
      rc = MDEGGS_ENOMEM;
      break;
    }
    for (size_t c = 0; c < cols.size(); ++c)
      cudaMemcpyAsync(dx.as<double>() + c * (size_t)n, x + (size_t)cols[c] * ldx, (size_t)n * 8, cudaMemcpyHostToDevice,
                      s.s_main);
    cudaMemcpyAsync(da.p, ra.data(), (size_t)n_pairs * 4, cudaMemcpyHostToDevice, s.s_main);
    cudaMemcpyAsync(db.p, rb.data(), (size_t)n_pairs * 4, cudaMemcpyHostToDevice, s.s_main);
    LAUNCH(k_pair_features, blocks_for((int64_t)total, 256), 256, 0, s.s_main, dx.as<double>(), (int64_t)n, n,
           da.as<int32_t>(), db.as<int32_t>(), n_pairs, ratio, dout.as<double>());
    cudaMemcpyAsync(out, dout.p, total * 8, cudaMemcpyDeviceToHost, s.s_main);
    if (cudaStreamSynchronize(s.s_main) != cudaSuccess) rc = MDEGGS_ECUDA;
  } while (0);
  dx.release();
  da.release();
  db.release();
  dout.release();
  return rc;
}

int mdeggs_host_pack_rows(const double* src, int64_t ld, int32_t G, int32_t n, const int32_t* from, const int32_t* to,
                          int64_t E, int32_t chunk_rows, double* dst, int64_t ld_dst, int32_t* rows_out, int32_t* n_nodes) {
  if (!src || !from || !to || !n_nodes || G < 1 || n < 1 || E < 0 || ld < G) return MDEGGS_EINVAL;
  std::vector<int32_t> rows, rank;
  pack_node_rows(from, to, E, G, rows, rank);
  *n_nodes = (int32_t)rows.size();
  if (!dst) return MDEGGS_OK;  // (count only)
  if (ld_dst < (int64_t)rows.size() || rows.empty()) return rows.empty() ? MDEGGS_OK : MDEGGS_EINVAL;
  if (rows_out) memcpy(rows_out, rows.data(), rows.size() * 4);
  PackJob J;
  J.src = src;
  J.ld_src = ld;
  J.dst = dst;
  J.ld_dst = ld_dst;
  J.rows = rows.data();
  J.nn = (int)rows.size();
  J.n_cols = n;
  J.contiguous = (int64_t)rows.back() - rows.front() + 1 == (int64_t)rows.size();
  J.chunk_rows = chunk_rows > 0 ? chunk_rows : J.nn;
  J.chunks = (int)((J.nn + J.chunk_rows - 1) / J.chunk_rows);
  if (J.chunks > PACK_MAX_CHUNKS) return MDEGGS_EINVAL;
  J.units_per_chunk = (n + PACK_COL_BLOCK - 1) / PACK_COL_BLOCK;
  HostPool& pool = HostPool::get();
  pool.start(J);
  for (int c = 0; c < J.chunks; ++c)
    while (!pool.chunk_done(c))
      if (!pool.help()) _mm_pause();
  pool.finish();
  return MDEGGS_OK;
}

int mdeggs_pair_moments(mdeggs_ctx* ctx, const int32_t* a_rows, const int32_t* b_rows, int32_t n_pairs, double* out) {
  if (!ctx || !a_rows || !b_rows || !out || n_pairs < 0) return MDEGGS_EINVAL;
  if (ctx->pend.active) {
    set_err(ctx, "a submitted call is still pending: call mdeggs_wait first");
    return MDEGGS_EINVAL;
  }
  if (!ctx->last_valid) {
    set_err(ctx, "mdeggs_pair_moments needs a completed mdeggs_run_omic on this context");
    return MDEGGS_EINVAL;
  }
  if (n_pairs == 0) return MDEGGS_OK;
  Shard& s = ctx->shards[0];
  CU(cudaSetDevice(s.device));
  const SegLayout& L = ctx->last_L;
  const size_t n_out = (size_t)n_pairs * L.K * 6;
  DevBuf da, db, dout;
  int rc = MDEGGS_OK;
  if (da.ensure((size_t)n_pairs * 4) != cudaSuccess || db.ensure((size_t)n_pairs * 4) != cudaSuccess ||
      dout.ensure(n_out * 8) != cudaSuccess) {
    rc = MDEGGS_ENOMEM;
  } else {
    std::vector<int32_t> ta, tb;
    if (ctx->pack.last_packed) {  // the last call ran on the packed problem: rows of the caller's matrix -> node ranks
      const std::vector<int32_t>& rk = ctx->pack.rank;
      ta.resize((size_t)n_pairs);
      tb.resize((size_t)n_pairs);
      for (int32_t j = 0; j < n_pairs; ++j) {
        // (1-based on both sides; a row that was no node maps to 0 = invalid)
        ta[(size_t)j] = (a_rows[j] >= 1 && (size_t)a_rows[j] <= rk.size()) ? rk[(size_t)a_rows[j] - 1] + 1 : 0;
        tb[(size_t)j] = (b_rows[j] >= 1 && (size_t)b_rows[j] <= rk.size()) ? rk[(size_t)b_rows[j] - 1] + 1 : 0;
      }
      a_rows = ta.data();
      b_rows = tb.data();
    }
    cudaMemcpyAsync(da.p, a_rows, (size_t)n_pairs * 4, cudaMemcpyHostToDevice, s.s_main);
    cudaMemcpyAsync(db.p, b_rows, (size_t)n_pairs * 4, cudaMemcpyHostToDevice, s.s_main);
    LAUNCH(k_pair_moments, blocks_for((int64_t)n_pairs * 32, 256), 256, 0, s.s_main, s.D.as<double>(), L,
           s.node_of_row.as<int32_t>(), ctx->last_G, da.as<int32_t>(), db.as<int32_t>(), (int)n_pairs, s.mean.as<double>(),
           s.sxx.as<double>(), dout.as<double>());
    cudaMemcpyAsync(out, dout.p, n_out * 8, cudaMemcpyDeviceToHost, s.s_main);
    if (cudaStreamSynchronize(s.s_main) != cudaSuccess) rc = MDEGGS_ECUDA;
  }
  da.release();
  db.release();
  dout.release();
  return rc;
}

static int dev_math(mdeggs_ctx* ctx, int which, const double* x, const double* a, const double* b, int64_t n,
                    double* out) {
  if (!ctx || !x || !a || !out || n < 0) return MDEGGS_EINVAL;
  if (n == 0) return MDEGGS_OK;
  Shard& s = ctx->shards[0];
  CU(cudaSetDevice(s.device));
  DevBuf dx, da, db, dout;
  int rc = MDEGGS_OK;
  if (dx.ensure((size_t)n * 8) != cudaSuccess || da.ensure((size_t)n * 8) != cudaSuccess ||
      db.ensure((size_t)n * 8) != cudaSuccess || dout.ensure((size_t)n * 8) != cudaSuccess) {
    rc = MDEGGS_ENOMEM;
  } else {
    cudaMemcpyAsync(dx.p, x, (size_t)n * 8, cudaMemcpyHostToDevice, s.s_main);
    cudaMemcpyAsync(da.p, a, (size_t)n * 8, cudaMemcpyHostToDevice, s.s_main);
    cudaMemcpyAsync(db.p, b ? b : a, (size_t)n * 8, cudaMemcpyHostToDevice, s.s_main);
    LAUNCH(k_dev_math, blocks_for(n, 128), 128, 0, s.s_main, which, dx.as<double>(), da.as<double>(), db.as<double>(), n,
           dout.as<double>());
    cudaMemcpyAsync(out, dout.p, (size_t)n * 8, cudaMemcpyDeviceToHost, s.s_main);
    if (cudaStreamSynchronize(s.s_main) != cudaSuccess) rc = MDEGGS_ECUDA;
  }
  dx.release();
  da.release();
  db.release();
  dout.release();
  return rc;
}

int mdeggs_measure_fp64_peak(mdeggs_ctx* ctx, double* tflops_out) {
  if (!ctx || !tflops_out) return MDEGGS_EINVAL;
  Shard& s = ctx->shards[0];
  CU(cudaSetDevice(s.device));
  CU(s.small_out.ensure(64));
  const int iters = 8192, blocks = 148 * 8, threads = 256;
  float best_ms = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {  // first repetition warms up
    CU(cudaEventRecord(s.ev[EV_Q0], s.s_main));
    LAUNCH(k_dfma_peak, blocks, threads, 0, s.s_main, iters, 1.0 + rep, s.small_out.as<double>());
    CU(cudaEventRecord(s.ev[EV_Q1], s.s_main));
    CU(cudaStreamSynchronize(s.s_main));
    float ms = ev_ms(s.ev[EV_Q0], s.ev[EV_Q1]);
    if (rep > 0 && ms < best_ms) best_ms = ms;
  }
  const double flops = 2.0 * 16.0 * (double)iters * (double)blocks * (double)threads;
  *tflops_out = flops / ((double)best_ms * 1e-3) / 1e12;
  return MDEGGS_OK;
}

int mdeggs_measure_l2_peak(mdeggs_ctx* ctx, int64_t bytes, double* gbs_out) {
  if (!ctx || !gbs_out || bytes < (1 << 20)) return MDEGGS_EINVAL;
  Shard& s = ctx->shards[0];
  CU(cudaSetDevice(s.device));
  DevBuf buf;
  const size_t n_vec = (size_t)bytes / 16;
  if (buf.ensure(n_vec * 16) != cudaSuccess) return MDEGGS_ENOMEM;
  CU(s.small_out.ensure(64));
  cudaMemsetAsync(buf.p, 0x5a, n_vec * 16, s.s_main);
  // enough passes that the launch latency vanishes (>= 4 GB read from L2 per launch)
  const int passes = (int)(((size_t)4 << 30) / (n_vec * 16)) + 8;
  float best_ms = 1e30f;
  int rc = MDEGGS_OK;
  for (int ctas_per_sm : {4, 8}) {
    for (int rep = 0; rep < 4; ++rep) {  // the first repetition also brings the buffer into L2
      cudaEventRecord(s.ev[EV_Q0], s.s_main);
      LAUNCH(k_l2_read, 148 * ctas_per_sm, 256, 0, s.s_main, buf.as<uint4>(), n_vec, passes,
             s.small_out.as<unsigned int>());
      cudaEventRecord(s.ev[EV_Q1], s.s_main);
      if (cudaStreamSynchronize(s.s_main) != cudaSuccess) rc = MDEGGS_ECUDA;
      const float ms = ev_ms(s.ev[EV_Q0], s.ev[EV_Q1]);
      if (rep > 0 && ms > 0.f && ms < best_ms) best_ms = ms;
    }
  }
  buf.release();
  if (rc) return rc;
  *gbs_out = (double)(n_vec * 16) * passes / ((double)best_ms * 1e-3) / 1e9;
  return MDEGGS_OK;
}

int mdeggs_dev_pt(mdeggs_ctx* ctx, const double* x, const double* df, int64_t n, double* out) {
  return dev_math(ctx, 0, x, df, nullptr, n, out);
}
int mdeggs_dev_pf_upper(mdeggs_ctx* ctx, const double* f, const double* df1, const double* df2, int64_t n, double* out) {
  return dev_math(ctx, 1, f, df1, df2, n, out);
}
int mdeggs_dev_p_two_sided_ref(mdeggs_ctx* ctx, const double* t, const double* df, int64_t n, double* out) {
  return dev_math(ctx, 2, t, df, nullptr, n, out);
}

int mdeggs_dev_p_two_sided_tab(mdeggs_ctx* ctx, const double* t, double df, int32_t use_spline, int64_t n, double* out) {
  if (!ctx || !t || !out || n < 0 || !(df >= 1.0)) return MDEGGS_EINVAL;
  if (n == 0) return MDEGGS_OK;
  Shard& s = ctx->shards[0];
  CU(cudaSetDevice(s.device));
  CU(s.cf_tab.ensure((size_t)CF_TABLE_DOUBLES * 8));
  CU(s.tail_sp.ensure((size_t)TS_M * TS_C * 8));
  DevBuf dx, dout;
  int rc = MDEGGS_OK;
  if (dx.ensure((size_t)n * 8) != cudaSuccess || dout.ensure((size_t)n * 8) != cudaSuccess) {
    rc = MDEGGS_ENOMEM;
  } else {
    double* cf = s.cf_tab.as<double>();
    TailConsts tc;
    memset(&tc, 0, sizeof tc);
    tc.a = 0.5;
    tc.b = df / 2.0;
    tc.e_ab = cf;
    tc.e_ba = cf + CF_TABLE_LEN;
    tc.c_ab = cf + 2 * CF_TABLE_LEN + 2;
    tc.c_ba = cf + 3 * CF_TABLE_LEN + 2;
    LAUNCH(k_cf_tables, 1, 256, 0, s.s_main, 0.5, df / 2.0, cf, cf + CF_TABLE_LEN, cf + 2 * CF_TABLE_LEN,
           cf + 2 * CF_TABLE_LEN + 2, cf + 3 * CF_TABLE_LEN + 2);
    if (use_spline) {
      LAUNCH(k_tail_spline, TS_M * TS_C / 256, 256, 0, s.s_main, tc, cf + 2 * CF_TABLE_LEN, df,
             tail_spline_smax(df) / (double)TS_M, s.tail_sp.as<double>());
      tc.sp = s.tail_sp.as<double>();
      tc.sp_inv_h = (double)TS_M / tail_spline_smax(df);
      tc.sp_inv_n = 1.0 / df;
    }
    cudaMemcpyAsync(dx.p, t, (size_t)n * 8, cudaMemcpyHostToDevice, s.s_main);
    LAUNCH(k_dev_tail_tab, blocks_for(n, 128), 128, 0, s.s_main, tc, cf + 2 * CF_TABLE_LEN, df, dx.as<double>(), n,
           dout.as<double>());
    cudaMemcpyAsync(out, dout.p, (size_t)n * 8, cudaMemcpyDeviceToHost, s.s_main);
    if (cudaStreamSynchronize(s.s_main) != cudaSuccess) rc = MDEGGS_ECUDA;
  }
  dx.release();
  dout.release();
  return rc;
}

}  // extern "C"

// mdeggs_hostpack.h — host side of the input path for matrices that a single DMA serves badly: node rows scattered over a
// column-major matrix (the whole matrix would have to cross PCIe: config 3 with shuffled rows moves 160 MB for 67 MB of
// node rows) and PAGEABLE matrices (an R numeric matrix is never pinned; the driver stages such copies through its own
// small bounce buffers).  Only the rows that are network nodes are ever read (core_functions.R:333-336), so the host
// gathers exactly those rows into pinned staging, row chunk by row chunk with a small pool of worker threads, and every
// finished chunk goes out at once as one 2-D DMA while the next one is being packed; on the device the chunks open the
// gates the front kernel's loader waits on (mdeggs_kernels.cuh: gate_wait).  The packed problem (G' = number of nodes,
// edge end points renumbered by node rank) gives bit-identical results: node order is ascending row order either way.
#pragma once
#include <emmintrin.h>
#include <sched.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

namespace mdeggs {

constexpr int PACK_MAX_CHUNKS = 32;  // (== GATE_MAX_CHUNKS)
constexpr int PACK_COL_BLOCK = 8;    // columns per work unit

struct PackJob {
  const double* src = nullptr;  // column-major G x n, leading dimension ld_src
  int64_t ld_src = 0;
  double* dst = nullptr;        // column-major nn x n, leading dimension ld_dst (pinned staging)
  int64_t ld_dst = 0;
  const int32_t* rows = nullptr;  // packed row -> source row, ascending
  int nn = 0, n_cols = 0;
  int64_t chunk_rows = 0;
  int chunks = 0, units_per_chunk = 0;
  bool contiguous = false;  // rows[i] == rows[0] + i: plain memcpy per column segment
};

// one unit = (row chunk, block of PACK_COL_BLOCK columns); units are numbered chunk-major so chunk 0 is finished first
inline void pack_unit(const PackJob& J, int u) {
  const int c = u / J.units_per_chunk, b = u - c * J.units_per_chunk;
  const int64_t r0 = (int64_t)c * J.chunk_rows;
  const int64_t r1 = r0 + J.chunk_rows < J.nn ? r0 + J.chunk_rows : J.nn;
  const int col0 = b * PACK_COL_BLOCK, col1 = col0 + PACK_COL_BLOCK < J.n_cols ? col0 + PACK_COL_BLOCK : J.n_cols;
  const int32_t* __restrict__ rows = J.rows;
  for (int col = col0; col < col1; ++col) {
    const double* __restrict__ sc = J.src + (size_t)col * J.ld_src;
    double* __restrict__ dc = J.dst + (size_t)col * J.ld_dst;
    if (J.contiguous) {  // a plain copy - with streaming stores as well: a cached memcpy reads the destination lines first
      const double* __restrict__ s0 = sc + rows[r0];
      int64_t i = r0;
      if ((reinterpret_cast<uintptr_t>(dc + i) & 15) && i < r1) {
        dc[i] = s0[i - r0];
        ++i;
      }
      for (; i + 8 <= r1; i += 8) {
        const __m128d a = _mm_loadu_pd(s0 + (i - r0)), b = _mm_loadu_pd(s0 + (i - r0) + 2);
        const __m128d c = _mm_loadu_pd(s0 + (i - r0) + 4), d = _mm_loadu_pd(s0 + (i - r0) + 6);
        _mm_stream_pd(dc + i, a);
        _mm_stream_pd(dc + i + 2, b);
        _mm_stream_pd(dc + i + 4, c);
        _mm_stream_pd(dc + i + 6, d);
      }
      for (; i < r1; ++i) dc[i] = s0[i - r0];
      continue;
    }
    int64_t i = r0;
    // streaming stores (the staging is written once and read by the DMA engine only): no read-for-ownership traffic
    if ((reinterpret_cast<uintptr_t>(dc + i) & 15) && i < r1) {
      dc[i] = sc[rows[i]];
      ++i;
    }
    for (; i + 2 <= r1; i += 2)
      _mm_stream_pd(dc + i, _mm_set_pd(sc[rows[i + 1]], sc[rows[i]]));
    if (i < r1) dc[i] = sc[rows[i]];
  }
  _mm_sfence();
}

// Process-wide pool of packing threads (created on first use; the calling thread works too).  One job at a time.
class HostPool {
 public:
  static HostPool& get() {
    static HostPool* p = new HostPool();  // (leaked on purpose: worker threads must not outlive a static destructor)
    return *p;
  }
  int threads() const { return (int)workers_.size() + 1; }

  // Starts job J (kept by reference until finish()).  The caller then alternates between help() and polling done(c).
  void start(const PackJob& J) {
    job_mu_.lock();  // one job at a time, process-wide (contexts of several omic layers share the pool)
    // workers that woke up late for the previous job still have to check in (they need mu_ to wake: not held here)
    while (active_.load(std::memory_order_acquire) != 0) std::this_thread::yield();
    std::unique_lock<std::mutex> lk(mu_);
    job_ = J;
    total_ = J.chunks * J.units_per_chunk;
    for (int c = 0; c < PACK_MAX_CHUNKS; ++c) done_[c].store(0, std::memory_order_relaxed);
    next_.store(0, std::memory_order_release);
    active_.store((int)workers_.size(), std::memory_order_release);
    ++gen_;
    lk.unlock();
    cv_.notify_all();
  }
  // the calling thread packs one unit; false when none is left to hand out
  bool help() {
    const int u = next_.fetch_add(1, std::memory_order_acq_rel);
    if (u >= total_) return false;
    pack_unit(job_, u);
    done_[u / job_.units_per_chunk].fetch_add(1, std::memory_order_release);
    return true;
  }
  bool chunk_done(int c) const { return done_[c].load(std::memory_order_acquire) >= job_.units_per_chunk; }
  void finish() { job_mu_.unlock(); }

 private:
  HostPool() {
    int ncpu = (int)std::thread::hardware_concurrency();
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof set, &set) == 0) {
      const int a = CPU_COUNT(&set);
      if (a > 0 && (ncpu <= 0 || a < ncpu)) ncpu = a;
    }
    if (const char* e = getenv("MDEGGS_PACK_THREADS")) ncpu = atoi(e);
    ncpu = ncpu < 1 ? 1 : (ncpu > 32 ? 32 : ncpu);
    for (int i = 0; i < PACK_MAX_CHUNKS; ++i) done_[i].store(0);
    for (int t = 1; t < ncpu; ++t) workers_.emplace_back([this]() { worker(); });
    for (auto& w : workers_) w.detach();
  }
  void worker() {
    uint64_t seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&]() { return gen_ != seen; });
        seen = gen_;
      }
      for (;;) {
        const int u = next_.fetch_add(1, std::memory_order_acq_rel);
        if (u >= total_) break;
        pack_unit(job_, u);
        done_[u / job_.units_per_chunk].fetch_add(1, std::memory_order_release);
      }
      active_.fetch_sub(1, std::memory_order_acq_rel);
    }
  }
  std::vector<std::thread> workers_;
  std::mutex mu_, job_mu_;
  std::condition_variable cv_;
  uint64_t gen_ = 0;
  PackJob job_;
  int total_ = 0;
  std::atomic<int> next_{1 << 30};
  std::atomic<int> active_{0};
  std::atomic<int> done_[PACK_MAX_CHUNKS];
};

// Node rows of an edge list on the host: rows[] (ascending source rows, 0-based) and rank[] (source row -> packed row,
// -1: no node).  Indices outside 1..G are skipped here (the device check reports them); returns false when there are any.
inline bool pack_node_rows(const int32_t* from, const int32_t* to, int64_t E, int G, std::vector<int32_t>& rows,
                           std::vector<int32_t>& rank) {
  rank.assign((size_t)G, -1);
  bool ok = true;
  for (const int32_t* v : {from, to})
    for (int64_t e = 0; e < E; ++e) {
      const int32_t r = v[e];
      if (r >= 1 && r <= G) rank[(size_t)r - 1] = 0;
      else ok = false;
    }
  rows.clear();
  for (int r = 0; r < G; ++r)
    if (rank[(size_t)r] == 0) {
      rank[(size_t)r] = (int32_t)rows.size();
      rows.push_back(r);
    }
  return ok;
}

}  // namespace mdeggs

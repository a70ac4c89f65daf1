// mdeggs_p2p.cuh — K9 over peer memory: the all-gather of the per-rank p-value / status blocks as ONE push kernel.
//
// p.adjust couples all edges of a (percentile, category) segment, so every rank needs every rank's p-values
// (core_functions.R:1027-1035 on the whole edge list).  With one process per GPU the ranks map each other's p_edge /
// status-byte / flag buffers once (cudaIpcGetMemHandle / cudaIpcOpenMemHandle, handles exchanged through the NCCL
// communicator) and from then on a rank's k_push_gather writes its own block straight into every peer's buffer over
// NVLink - 16-byte stores of two p-values, 8-byte stores of eight status codes packed to bytes on the way (the separate
// packing kernel is gone) - and raises a flag word at every peer; k_wait_gather holds the stream until every peer's flag has arrived.
// Two NCCL launches (64 + 8 MB per rank on config 5) with their staging and launch latency become one kernel that
// runs at link speed.  Reuse of the buffers by the next call is fenced the other way round: k_signal_done tells every
// peer "I have read your values of call c", and the push of call c + 1 waits for those words first.
// All waits are bounded (10 s of the global timer) and raise the device error flag instead of hanging.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mdeggs {

constexpr int P2P_MAX = 8;
constexpr int P2P_ERR = 3;

struct PushArgs {
  double* p_dst[P2P_MAX];           // p_edge of every rank (own entry: the local buffer)
  int8_t* st_dst[P2P_MAX];          // status bytes of every rank
  unsigned int* flag_dst[P2P_MAX];  // flag words of every rank: arrive[P2P_MAX] then done[P2P_MAX]
  const int32_t* status;            // local status codes (int32)
  int world, rank;
  long long lo, n_own;              // this rank's block of the edge list
  unsigned int* ctl;                // local: [0] calls completed (epoch), [1] CTA counter of the push kernel
  int32_t* err;
};

__device__ __forceinline__ unsigned int p2p_ld_acquire(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void p2p_st_release(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// waits until *p >= want (epochs only grow)
__device__ __forceinline__ void p2p_wait_ge(const unsigned int* p, unsigned int want, int32_t* err) {
  if ((int)(p2p_ld_acquire(p) - want) >= 0) return;
  unsigned long long t0, t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (;;) {
    __nanosleep(100);
    if ((int)(p2p_ld_acquire(p) - want) >= 0) return;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if (t - t0 > 10000000000ull) {
      if (err) atomicMax(err, P2P_ERR);
      return;
    }
  }
}

__global__ void __launch_bounds__(256) k_push_gather(PushArgs A) {
  const unsigned int ep = A.ctl[0];
  // every peer has finished reading what the previous call put into its buffers
  if ((int)threadIdx.x < A.world && (int)threadIdx.x != A.rank)
    p2p_wait_ge(A.flag_dst[A.rank] + P2P_MAX + threadIdx.x, ep, A.err);
  __syncthreads();
  const double* __restrict__ src = A.p_dst[A.rank] + A.lo;
  const int32_t* __restrict__ st = A.status + A.lo;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nthr = (long long)gridDim.x * blockDim.x;
  if ((A.lo & 7) == 0) {
    // aligned block (the usual case): 16-byte stores of two p-values, 8-byte stores of eight status bytes
    const long long n2 = A.n_own >> 1;
    const double2* __restrict__ src2 = reinterpret_cast<const double2*>(src);
    for (long long i = tid; i < n2; i += nthr) {
      const double2 v = src2[i];
#pragma unroll
      for (int r = 0; r < P2P_MAX; ++r) {
        if (r >= A.world) break;
        if (r != A.rank) reinterpret_cast<double2*>(A.p_dst[r] + A.lo)[i] = v;
      }
    }
    const long long n8 = A.n_own >> 3;
    const int4* __restrict__ st4 = reinterpret_cast<const int4*>(st);
    for (long long i = tid; i < n8; i += nthr) {
      const int4 a = st4[2 * i], b = st4[2 * i + 1];
      const unsigned int w0 = (unsigned)(a.x & 0xff) | ((unsigned)(a.y & 0xff) << 8) | ((unsigned)(a.z & 0xff) << 16) |
                              ((unsigned)(a.w & 0xff) << 24);
      const unsigned int w1 = (unsigned)(b.x & 0xff) | ((unsigned)(b.y & 0xff) << 8) | ((unsigned)(b.z & 0xff) << 16) |
                              ((unsigned)(b.w & 0xff) << 24);
      const uint2 w = make_uint2(w0, w1);
#pragma unroll
      for (int r = 0; r < P2P_MAX; ++r) {
        if (r >= A.world) break;
        reinterpret_cast<uint2*>(A.st_dst[r] + A.lo)[i] = w;  // (own entry too: the local status bytes of the own block)
      }
    }
    // tails
    for (long long i = 2 * n2 + tid; i < A.n_own; i += nthr)
      for (int r = 0; r < A.world; ++r)
        if (r != A.rank) A.p_dst[r][A.lo + i] = src[i];
    for (long long i = 8 * n8 + tid; i < A.n_own; i += nthr)
      for (int r = 0; r < A.world; ++r) A.st_dst[r][A.lo + i] = (int8_t)st[i];
  } else {
    for (long long i = tid; i < A.n_own; i += nthr) {
      const double v = src[i];
      const int8_t c = (int8_t)st[i];
#pragma unroll
      for (int r = 0; r < P2P_MAX; ++r) {
        if (r >= A.world) break;
        if (r != A.rank) A.p_dst[r][A.lo + i] = v;
        A.st_dst[r][A.lo + i] = c;  // (own entry too: the local status bytes of the own block)
      }
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int prev = atomicAdd(A.ctl + 1, 1u);
    if (prev == gridDim.x - 1) {  // every CTA's stores are out: raise this rank's flag at every peer
      A.ctl[1] = 0u;
      __threadfence_system();
      for (int r = 0; r < A.world; ++r)
        if (r != A.rank) p2p_st_release(A.flag_dst[r] + A.rank, ep + 1u);
    }
  }
}

// holds the stream until the blocks of every peer have landed in the local buffers
__global__ void k_wait_gather(PushArgs A) {
  const unsigned int ep = A.ctl[0];
  if ((int)threadIdx.x < A.world && (int)threadIdx.x != A.rank)
    p2p_wait_ge(A.flag_dst[A.rank] + threadIdx.x, ep + 1u, A.err);
}

// end of the call: nothing on this rank reads the gathered values any more
__global__ void k_signal_done(PushArgs A) {
  const unsigned int ep = A.ctl[0];
  __syncthreads();
  if ((int)threadIdx.x < A.world && (int)threadIdx.x != A.rank)
    p2p_st_release(A.flag_dst[threadIdx.x] + P2P_MAX + A.rank, ep + 1u);
  __syncthreads();
  if (threadIdx.x == 0) A.ctl[0] = ep + 1u;
}

}  // namespace mdeggs

// nccl_dl.h — NCCL entry points resolved at run time (dlopen), so that libmdeggs.so has no link-time
// dependency on a particular libnccl and single-GPU use needs no NCCL at all.  When torch is already
// in the process its bundled libnccl.so.2 (2.28.x here) is the one the soname resolves to; otherwise the
// system libnccl (2.27.x) is used.  Only the stable subset needed for the p-value all-gather is declared.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stddef.h>
#include <stdlib.h>

namespace mdeggs {

typedef struct ncclComm* nccl_comm_t;
typedef struct {
  char internal[128];
} nccl_unique_id;
enum { NCCL_INT8 = 0, NCCL_INT32 = 2, NCCL_UINT64 = 5, NCCL_FLOAT64 = 8 };
enum { NCCL_SUM = 0, NCCL_MIN = 3 };

struct NcclApi {
  void* handle = nullptr;
  int (*GetUniqueId)(nccl_unique_id*) = nullptr;
  int (*CommInitRank)(nccl_comm_t*, int, nccl_unique_id, int) = nullptr;
  int (*CommInitAll)(nccl_comm_t*, int, const int*) = nullptr;
  int (*CommDestroy)(nccl_comm_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, nccl_comm_t, cudaStream_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;

  bool load() {
    if (handle) return true;
    const char* env = getenv("MDEGGS_NCCL_LIB");
    const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
      if (!nm) continue;
      handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
      if (handle) break;
    }
    if (!handle) return false;
#define MDEGGS_SYM(field, name) *(void**)(&field) = dlsym(handle, name)
    MDEGGS_SYM(GetUniqueId, "ncclGetUniqueId");
    MDEGGS_SYM(CommInitRank, "ncclCommInitRank");
    MDEGGS_SYM(CommInitAll, "ncclCommInitAll");
    MDEGGS_SYM(CommDestroy, "ncclCommDestroy");
    MDEGGS_SYM(AllGather, "ncclAllGather");
    MDEGGS_SYM(AllReduce, "ncclAllReduce");
    MDEGGS_SYM(GroupStart, "ncclGroupStart");
    MDEGGS_SYM(GroupEnd, "ncclGroupEnd");
    MDEGGS_SYM(GetErrorString, "ncclGetErrorString");
#undef MDEGGS_SYM
    return GetUniqueId && CommInitRank && CommInitAll && CommDestroy && AllGather && AllReduce && GroupStart && GroupEnd;
  }
};

inline NcclApi& nccl_api() {
  static NcclApi api;
  return api;
}

}  // namespace mdeggs

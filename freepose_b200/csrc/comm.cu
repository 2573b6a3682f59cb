// The ONE collective of the hot path: an all-gather of the per-hypothesis fp32 scores, issued
// from the C ABI on the compute stream, right behind the score kernel that wrote this rank's slice of the buffer.
//
// NCCL is resolved at run time (dlopen of the libnccl.so.2 that PyTorch already mapped into the process; no link-time
// dependency, so the library still loads on a box without NCCL and the single-GPU path never touches it).  Only the
// five entry points used here are declared, with the types of nccl.h 2.x spelled out (ncclUniqueId = 128 opaque bytes,
// ncclFloat32 = 7, results are ints with 0 = success).
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>

#include <mutex>

#include "common.cuh"
#include "freepose_b200.h"
#include "kernels.h"

namespace fp {

namespace {

struct NcclApi {
  int (*GetUniqueId)(void* id);
  int (*CommInitRank)(void** comm, int nranks, fp_comm_id id, int rank);
  int (*AllGather)(const void* send, void* recv, size_t count, int dtype, void* comm, cudaStream_t stream);
  int (*CommDestroy)(void* comm);
  const char* (*GetErrorString)(int);
  bool ok = false;
  char why[256] = "";
};

NcclApi& nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    void* h = nullptr;
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      h = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (h) break;
    }
    if (!h) {
      snprintf(api.why, sizeof(api.why), "libnccl.so.2 not found (%s); import torch first or set LD_LIBRARY_PATH", dlerror());
      return;
    }
    auto sym = [&](const char* n) {
      void* p = dlsym(h, n);
      if (!p) snprintf(api.why, sizeof(api.why), "NCCL symbol %s missing", n);
      return p;
    };
    api.GetUniqueId = reinterpret_cast<int (*)(void*)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<int (*)(void**, int, fp_comm_id, int)>(sym("ncclCommInitRank"));
    api.AllGather = reinterpret_cast<int (*)(const void*, void*, size_t, int, void*, cudaStream_t)>(sym("ncclAllGather"));
    api.CommDestroy = reinterpret_cast<int (*)(void*)>(sym("ncclCommDestroy"));
    api.GetErrorString = reinterpret_cast<const char* (*)(int)>(sym("ncclGetErrorString"));
    api.ok = api.GetUniqueId && api.CommInitRank && api.AllGather && api.CommDestroy && api.GetErrorString;
  });
  return api;
}

struct Comm {
  void* nccl_comm;
  int rank, world;
};

#define FP_NCCL(expr)                                                                           \
  do {                                                                                          \
    int _r = (expr);                                                                            \
    if (_r != 0) {                                                                              \
      set_error("NCCL error %d (%s) at: %s", _r, nccl().GetErrorString(_r), #expr);             \
      return -3;                                                                                \
    }                                                                                           \
  } while (0)

}  // namespace

int comm_unique_id(fp_comm_id* out) {
  FP_REQUIRE(out != nullptr, "comm: null id");
  NcclApi& api = nccl();
  FP_REQUIRE(api.ok, "comm: %s", api.why);
  FP_NCCL(api.GetUniqueId(out));
  return 0;
}

int comm_create(const fp_comm_id* id, int rank, int world, void** comm_out) {
  FP_REQUIRE(id && comm_out, "comm: null argument");
  FP_REQUIRE(world >= 1 && rank >= 0 && rank < world, "comm: rank %d outside world %d", rank, world);
  NcclApi& api = nccl();
  FP_REQUIRE(api.ok, "comm: %s", api.why);
  void* c = nullptr;
  FP_NCCL(api.CommInitRank(&c, world, *id, rank));
  *comm_out = new Comm{c, rank, world};
  return 0;
}

int comm_allgather_scores(void* comm, float* scores, int per_rank, cudaStream_t stream) {
  FP_REQUIRE(comm && scores, "allgather: null argument");
  FP_REQUIRE(per_rank > 0, "allgather: per_rank must be positive");
  Comm* c = reinterpret_cast<Comm*>(comm);
  // in place: this rank's scores already sit at scores + rank * per_rank (the score kernel wrote them there)
  FP_NCCL(nccl().AllGather(scores + size_t(c->rank) * per_rank, scores, size_t(per_rank), /*ncclFloat32*/ 7, c->nccl_comm,
                           stream));
  return 0;
}

// ---- peer memory through CUDA IPC (one process per GPU on one node: NVLink / NVSwitch peer access) --------------------
int p2p_alloc(size_t bytes, void** ptr, void* handle64) {
  FP_REQUIRE(ptr && handle64 && bytes > 0, "p2p_alloc: bad arguments");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  void* p = nullptr;
  FP_CUDA(cudaMalloc(&p, bytes));
  FP_CUDA(cudaMemset(p, 0, bytes));
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) { cudaFree(p); return cuda_fail(e, "cudaIpcGetMemHandle"); }
  memcpy(handle64, &h, 64);
  *ptr = p;
  return 0;
}
int p2p_open(const void* handle64, void** ptr) {
  FP_REQUIRE(ptr && handle64, "p2p_open: bad arguments");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  FP_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}
int p2p_close(void* ptr) {
  if (ptr) FP_CUDA(cudaIpcCloseMemHandle(ptr));
  return 0;
}
int p2p_free(void* ptr) {
  if (ptr) FP_CUDA(cudaFree(ptr));
  return 0;
}

int comm_destroy(void* comm) {
  if (!comm) return 0;
  Comm* c = reinterpret_cast<Comm*>(comm);
  int r = nccl().CommDestroy(c->nccl_comm);
  delete c;
  FP_REQUIRE(r == 0, "NCCL error %d in ncclCommDestroy", r);
  return 0;
}

}  // namespace fp

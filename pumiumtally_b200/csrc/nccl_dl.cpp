// NCCL bindings resolved with dlopen/dlsym so libpumitally.so has no link-time
// dependency on a particular libnccl (the process may already carry the copy
// bundled with PyTorch; a single-GPU run needs none at all).
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <dlfcn.h>

#include <cuda_runtime.h>

#include "engine.hpp"

namespace ptb {
namespace {

// ABI of the handful of NCCL 2.x entry points used (stable since 2.0).
struct UniqueId { char internal[128]; };
using Comm = void *;
constexpr int kNcclFloat64 = 8;  // ncclDouble
constexpr int kNcclSum = 0;      // ncclSum

struct Api {
  void *handle = nullptr;
  int (*GetUniqueId)(UniqueId *) = nullptr;
  int (*CommInitRank)(Comm *, int, UniqueId, int) = nullptr;
  int (*AllReduce)(const void *, void *, size_t, int, int, Comm, cudaStream_t) = nullptr;
  int (*ReduceScatter)(const void *, void *, size_t, int, int, Comm, cudaStream_t) = nullptr;
  int (*AllGather)(const void *, void *, size_t, int, Comm, cudaStream_t) = nullptr;
  int (*CommDestroy)(Comm) = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
};

Api *api() {
  static Api a;
  static bool tried = false;
  if (tried) return a.handle ? &a : nullptr;
  tried = true;
  for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
    a.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
    if (a.handle) break;
  }
  if (!a.handle) {
    fprintf(stderr, "[pumitally] ERROR: cannot load libnccl: %s\n", dlerror());
    return nullptr;
  }
  a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(dlsym(a.handle, "ncclGetUniqueId"));
  a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(dlsym(a.handle, "ncclCommInitRank"));
  a.AllReduce = reinterpret_cast<decltype(a.AllReduce)>(dlsym(a.handle, "ncclAllReduce"));
  a.ReduceScatter = reinterpret_cast<decltype(a.ReduceScatter)>(dlsym(a.handle, "ncclReduceScatter"));
  a.AllGather = reinterpret_cast<decltype(a.AllGather)>(dlsym(a.handle, "ncclAllGather"));
  a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(dlsym(a.handle, "ncclCommDestroy"));
  a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(dlsym(a.handle, "ncclGetErrorString"));
  if (!a.GetUniqueId || !a.CommInitRank || !a.AllReduce || !a.CommDestroy) {
    fprintf(stderr, "[pumitally] ERROR: libnccl lacks required symbols\n");
    a.handle = nullptr;
    return nullptr;
  }
  return &a;
}

int check(Api *a, int rc, const char *what) {
  if (rc == 0) return 0;
  fprintf(stderr, "[pumitally] NCCL error in %s: %s\n", what,
          a->GetErrorString ? a->GetErrorString(rc) : "?");
  return 1;
}

}  // namespace

int nccl_get_unique_id(uint8_t out[128]) {
  Api *a = api();
  if (!a) return 1;
  UniqueId id;
  if (check(a, a->GetUniqueId(&id), "ncclGetUniqueId")) return 1;
  std::memcpy(out, id.internal, 128);
  return 0;
}

int nccl_comm_init_rank(void **comm, int nranks, const uint8_t idb[128], int rank) {
  Api *a = api();
  if (!a) return 1;
  UniqueId id;
  std::memcpy(id.internal, idb, 128);
  return check(a, a->CommInitRank(comm, nranks, id, rank), "ncclCommInitRank");
}

int nccl_allreduce_sum_f64(void *comm, const double *send, double *recv, size_t count, cudaStream_t stream) {
  Api *a = api();
  if (!a) return 1;
  return check(a, a->AllReduce(send, recv, count, kNcclFloat64, kNcclSum, comm, stream), "ncclAllReduce");
}

// recv[0..count) = sum over ranks of send[rank*count .. (rank+1)*count)
int nccl_reduce_scatter_sum_f64(void *comm, const double *send, double *recv, size_t count, cudaStream_t stream) {
  Api *a = api();
  if (!a || !a->ReduceScatter) return 1;
  return check(a, a->ReduceScatter(send, recv, count, kNcclFloat64, kNcclSum, comm, stream), "ncclReduceScatter");
}

// recv[r*count .. (r+1)*count) = rank r's send[0..count); in place when send == recv + rank*count
int nccl_all_gather_f64(void *comm, const double *send, double *recv, size_t count, cudaStream_t stream) {
  Api *a = api();
  if (!a || !a->AllGather) return 1;
  return check(a, a->AllGather(send, recv, count, kNcclFloat64, comm, stream), "ncclAllGather");
}

void nccl_comm_destroy(void *comm) {
  Api *a = api();
  if (a && comm) a->CommDestroy(comm);
}

}  // namespace ptb

// NCCL behind the C ABI (include/b200gan.h: b200_nccl_*): the data-parallel exchange that replaces the reference's
// CPU-side `average_gradients` (util.py:118-147) -- an in-place sum all-reduce of a slice of the flat fp32 gradient
// bucket over NVLink, asynchronous on the caller's stream (capturable in a CUDA graph).  libnccl is resolved at run
// time with dlopen (the library torch bundles, or the system one): the shared object has no link-time dependency
// on it and a single-GPU process never loads it.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>
#include <string>

#include "../../include/b200gan.h"

namespace {

// the subset of nccl.h this file needs (stable ABI since NCCL 2.0)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclSum = 0 };
enum { ncclFloat32 = 7 };

struct Api {
  void* handle = nullptr;
  int (*GetUniqueId)(ncclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  int (*GetVersion)(int*) = nullptr;
} api;

thread_local std::string g_nccl_err;

int nfail(const std::string& m) { g_nccl_err = m; return -1; }
int ncheck(int rc, const char* what) {
  if (rc == ncclSuccess) return 0;
  return nfail(std::string(what) + ": " + (api.GetErrorString ? api.GetErrorString(rc) : "nccl error"));
}

template <typename F>
bool sym(F& f, const char* name) {
  f = reinterpret_cast<F>(dlsym(api.handle, name));
  return f != nullptr;
}

}  // namespace

extern "C" const char* b200_nccl_last_error(void) { return g_nccl_err.c_str(); }

extern "C" int b200_nccl_load(const char* path) {
  if (api.handle) return 0;
  const char* candidates[] = {path, "libnccl.so.2", "libnccl.so"};
  for (const char* c : candidates) {
    if (!c || !*c) continue;
    api.handle = dlopen(c, RTLD_NOW | RTLD_GLOBAL);
    if (api.handle) break;
  }
  if (!api.handle) return nfail(std::string("cannot load libnccl: ") + (dlerror() ? dlerror() : "not found"));
  bool ok = sym(api.GetUniqueId, "ncclGetUniqueId") && sym(api.CommInitRank, "ncclCommInitRank") &&
            sym(api.CommDestroy, "ncclCommDestroy") && sym(api.AllReduce, "ncclAllReduce") &&
            sym(api.Broadcast, "ncclBroadcast") && sym(api.GetErrorString, "ncclGetErrorString") &&
            sym(api.GetVersion, "ncclGetVersion");
  if (!ok) { dlclose(api.handle); api.handle = nullptr; return nfail("libnccl lacks a required symbol"); }
  return 0;
}

extern "C" int b200_nccl_version(void) {
  int v = 0;
  if (!api.handle || api.GetVersion(&v) != ncclSuccess) return -1;
  return v;
}

extern "C" int b200_nccl_unique_id(void* out128) {
  if (!api.handle) return nfail("b200_nccl_load has not been called");
  ncclUniqueId id;
  if (ncheck(api.GetUniqueId(&id), "ncclGetUniqueId")) return -1;
  memcpy(out128, id.internal, 128);
  return 0;
}

extern "C" int b200_nccl_init(const void* id128, int rank, int world, void** comm_out) {
  if (!api.handle) return nfail("b200_nccl_load has not been called");
  if (!id128 || !comm_out || world < 1 || rank < 0 || rank >= world) return nfail("b200_nccl_init: bad arguments");
  ncclUniqueId id;
  memcpy(id.internal, id128, 128);
  ncclComm_t comm = nullptr;
  if (ncheck(api.CommInitRank(&comm, world, id, rank), "ncclCommInitRank")) return -1;
  *comm_out = comm;
  return 0;
}

extern "C" int b200_nccl_allreduce_f32(void* comm, float* buf, long long n, b200_stream s) {
  if (!api.handle || !comm) return nfail("b200_nccl_allreduce_f32: no communicator");
  if (n <= 0) return 0;
  return ncheck(api.AllReduce(buf, buf, (size_t)n, ncclFloat32, ncclSum, (ncclComm_t)comm, (cudaStream_t)s), "ncclAllReduce");
}

extern "C" int b200_nccl_broadcast_f32(void* comm, float* buf, long long n, int root, b200_stream s) {
  if (!api.handle || !comm) return nfail("b200_nccl_broadcast_f32: no communicator");
  if (n <= 0) return 0;
  return ncheck(api.Broadcast(buf, buf, (size_t)n, ncclFloat32, root, (ncclComm_t)comm, (cudaStream_t)s), "ncclBroadcast");
}

extern "C" int b200_nccl_destroy(void* comm) {
  if (!api.handle || !comm) return 0;
  return ncheck(api.CommDestroy((ncclComm_t)comm), "ncclCommDestroy");
}

// comm.cpp — the collective side of the C ABI: one NCCL communicator per context (one process per GPU) and
// the table of the cluster-wide split ids that gives every split a global tie-break rank.
//
// The root of a Quickwit cluster merges the leaves' LeafSearchResponses (quickwit-search/src/root.rs:836-853,
// collector.rs:914-974). With every split of a query resident on one of the GPUs of a box, that merge is a
// single all-gather of fixed-size per-GPU records over NVLink; `qwgpu_leaf_search_allgather`
// (leaf.cpp) runs it on the device. NCCL is bound at run time (dlopen of libnccl.so.2 — the one the process
// already loaded, e.g. torch's), so libqwgpu.so has no link-time dependency on it.
#include <cuda_runtime_api.h>
#include <dlfcn.h>

#include <algorithm>
#include <map>

#include "comm.h"

namespace qw {

namespace {
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, NcclUniqueId, int) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, void*, void*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
NcclApi& nccl() {
  static NcclApi api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    const char* names[] = {getenv("QWGPU_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      if (!n || !*n) continue;
      api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (api.lib) break;
    }
    if (api.lib) {
      api.GetUniqueId = (int (*)(void*))dlsym(api.lib, "ncclGetUniqueId");
      api.CommInitRank = (int (*)(void**, int, NcclUniqueId, int))dlsym(api.lib, "ncclCommInitRank");
      api.CommDestroy = (int (*)(void*))dlsym(api.lib, "ncclCommDestroy");
      api.AllGather = (int (*)(const void*, void*, size_t, int, void*, void*))dlsym(api.lib, "ncclAllGather");
      api.GetErrorString = (const char* (*)(int))dlsym(api.lib, "ncclGetErrorString");
    }
  }
  if (!api.lib || !api.GetUniqueId || !api.CommInitRank || !api.AllGather)
    fail(QWGPU_EUNSUPPORTED, "NCCL is not available in this process (libnccl.so.2 could not be loaded: %s)", dlerror() ? dlerror() : "missing symbols");
  return api;
}
void check_nccl(int rc, const char* what) {
  if (rc != 0) {
    const NcclApi& a = nccl();
    fail(QWGPU_EINTERNAL, "%s failed: %s", what, a.GetErrorString ? a.GetErrorString(rc) : "NCCL error");
  }
}
}  // namespace

int comm_allgather(void* comm, const void* send, void* recv, size_t bytes, void* stream) {
  return nccl().AllGather(send, recv, bytes, /*ncclChar*/ 0, comm, stream);
}

void comm_unique_id(uint8_t out[128]) {
  NcclUniqueId id;
  check_nccl(nccl().GetUniqueId(&id), "ncclGetUniqueId");
  memcpy(out, id.internal, 128);
}

Comm* comm_create(int device, const uint8_t id_bytes[128], int rank, int world) {
  if (world < 1 || rank < 0 || rank >= world) fail(QWGPU_EINVALID_ARG, "bad rank %d / world %d", rank, world);
  NcclUniqueId id;
  memcpy(id.internal, id_bytes, 128);
  std::unique_ptr<Comm> c(new Comm());
  c->rank = rank;
  c->world = world;
  c->device = device;
  if (cudaSetDevice(device) != cudaSuccess) fail(QWGPU_EINTERNAL, "cudaSetDevice(%d) failed", device);
  check_nccl(nccl().CommInitRank(&c->comm, world, id, rank), "ncclCommInitRank");
  return c.release();
}

void comm_allgather_host(Comm* c, const uint8_t* send, uint8_t* recv, size_t bytes) {
  auto cu = [](cudaError_t e, const char* what) { if (e != cudaSuccess) fail(QWGPU_EINTERNAL, "%s: %s", what, cudaGetErrorString(e)); };
  cu(cudaSetDevice(c->device), "cudaSetDevice");
  if (!c->stream) { cudaStream_t s; cu(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking), "cudaStreamCreate"); c->stream = s; }
  const size_t need = bytes * (size_t)(c->world + 1);
  if (need > c->stage_cap) {
    if (c->d_stage) cudaFree(c->d_stage);
    c->d_stage = nullptr; c->stage_cap = 0;
    cu(cudaMalloc((void**)&c->d_stage, need), "cudaMalloc (exchange staging)");
    c->stage_cap = need;
  }
  cudaStream_t st = (cudaStream_t)c->stream;
  uint8_t *d_send = c->d_stage, *d_recv = c->d_stage + bytes;
  cu(cudaMemcpyAsync(d_send, send, bytes, cudaMemcpyHostToDevice, st), "H2D of the partial");
  check_nccl(nccl().AllGather(d_send, d_recv, bytes, /*ncclChar*/ 0, c->comm, st), "ncclAllGather");
  cu(cudaMemcpyAsync(recv, d_recv, bytes * (size_t)c->world, cudaMemcpyDeviceToHost, st), "D2H of the gathered partials");
  cu(cudaStreamSynchronize(st), "exchange stream");
}

void comm_destroy(Comm* c) {
  if (!c) return;
  if (c->d_stage) cudaFree(c->d_stage);
  if (c->stream) cudaStreamDestroy((cudaStream_t)c->stream);
  if (c->comm && nccl().CommDestroy) nccl().CommDestroy(c->comm);
  delete c;
}

void comm_set_split_table(Comm* c, uint32_t n, const char* const* split_ids) {
  c->split_ids.assign(split_ids, split_ids + n);
  std::sort(c->split_ids.begin(), c->split_ids.end());
  c->split_ids.erase(std::unique(c->split_ids.begin(), c->split_ids.end()), c->split_ids.end());
}

int comm_split_rank(const Comm* c, const std::string& split_id) {
  auto it = std::lower_bound(c->split_ids.begin(), c->split_ids.end(), split_id);
  if (it == c->split_ids.end() || *it != split_id) return -1;
  return (int)(it - c->split_ids.begin());
}

}  // namespace qw

qwgpu_ctx::~qwgpu_ctx() {
  for (qw::Comm*& c : lanes) { qw::comm_destroy(c); c = nullptr; }
  comm = nullptr;
}

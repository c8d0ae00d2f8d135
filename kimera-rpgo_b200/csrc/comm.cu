/* comm.cu — the multi-GPU data plane of the PCM path behind the C ABI.
 *
 * The reference is single-threaded C++ (Pcm.h has no distributed code at all); what is sharded here is
 *   - the pairwise consistency matrix of a group, by row chunks (rpgo_lc_append, K3), followed by ONE exchange
 *     step: the all-gather of the adjacency row chunks, and
 *   - the root candidates of the clique searches (findCliqueHeu.cpp:32-117 candidate loop), with an all-reduce of
 *     the incumbent between epochs.
 * Both collectives are NCCL calls issued from C++ on the handle's own stream (NVLink 5 / NVSwitch on a B200 box),
 * so a C++ integrator needs nothing but rpgo_comm_unique_id / rpgo_comm_init.
 *
 * NCCL is bound at run time (dlopen of libnccl.so.2): the single-GPU library has no link-time dependency on it,
 * a process that already carries an NCCL (e.g. PyTorch's) shares that copy, and rpgo_comm_init fails loudly when
 * no NCCL is installed.  Only the long-stable v2 entry points are used.
 */
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>

#include "comm.h"

namespace rpgo {

namespace {

/* the slice of nccl.h this file needs (ABI-stable since NCCL 2.0) */
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclSuccess = 0 };
enum { ncclUint8 = 1, ncclInt64 = 4 };
enum { ncclMax = 2, ncclMin = 3 };

struct Api {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*CommAbort)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
  std::string err;
};

Api* api() {
  static Api* a = nullptr;  /* never freed: no library teardown at process exit */
  static std::once_flag once;
  std::call_once(once, []() {
    Api* x = new Api();
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
      x->lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
      if (x->lib) break;
    }
    if (!x->lib) {
      const char* e = dlerror();
      x->err = std::string("NCCL not found (dlopen libnccl.so.2): ") + (e ? e : "?");
      a = x;
      return;
    }
    bool ok = true;
    auto sym = [&](const char* nm) {
      void* p = dlsym(x->lib, nm);
      if (!p) { ok = false; x->err = std::string("NCCL symbol missing: ") + nm; }
      return p;
    };
    x->GetUniqueId = (decltype(x->GetUniqueId))sym("ncclGetUniqueId");
    x->CommInitRank = (decltype(x->CommInitRank))sym("ncclCommInitRank");
    x->CommDestroy = (decltype(x->CommDestroy))sym("ncclCommDestroy");
    x->CommAbort = (decltype(x->CommAbort))sym("ncclCommAbort");
    x->AllGather = (decltype(x->AllGather))sym("ncclAllGather");
    x->AllReduce = (decltype(x->AllReduce))sym("ncclAllReduce");
    x->Broadcast = (decltype(x->Broadcast))sym("ncclBroadcast");
    x->GroupStart = (decltype(x->GroupStart))sym("ncclGroupStart");
    x->GroupEnd = (decltype(x->GroupEnd))sym("ncclGroupEnd");
    x->GetErrorString = (decltype(x->GetErrorString))sym("ncclGetErrorString");
    x->GetVersion = (decltype(x->GetVersion))sym("ncclGetVersion");
    if (!ok) x->lib = nullptr;
    a = x;
  });
  return a;
}

}  // namespace

struct Comm {
  ncclComm_t nccl = nullptr;
  int rank = 0, world = 1;
  /* staging for the small host-side exchanges (incumbent words, batched results) */
  void* d_stage = nullptr;
  void* h_stage = nullptr;
  size_t stage_cap = 0;
  std::string err;
};

#define NCCL_TRY(c, call)                                                                      \
  do {                                                                                         \
    ncclResult_t r_ = (call);                                                                  \
    if (r_ != ncclSuccess) {                                                                   \
      (c)->err = std::string(#call) + ": " + (api()->GetErrorString ? api()->GetErrorString(r_) : "?"); \
      return 1;                                                                                \
    }                                                                                          \
  } while (0)
#define CUDA_TRY(c, call)                                                   \
  do {                                                                      \
    cudaError_t e_ = (call);                                                \
    if (e_ != cudaSuccess) {                                                \
      (c)->err = std::string(#call) + ": " + cudaGetErrorString(e_);        \
      return 2;                                                             \
    }                                                                       \
  } while (0)

int comm_unique_id(void* id_out, std::string* err) {
  Api* a = api();
  if (!a->lib) { if (err) *err = a->err; return 1; }
  ncclUniqueId id;
  const ncclResult_t r = a->GetUniqueId(&id);
  if (r != ncclSuccess) { if (err) *err = std::string("ncclGetUniqueId: ") + a->GetErrorString(r); return 1; }
  memcpy(id_out, id.internal, sizeof(id.internal));
  return 0;
}

int comm_create(Comm** out, const void* id128, int rank, int world, std::string* err) {
  Api* a = api();
  if (!a->lib) { if (err) *err = a->err; return 1; }
  Comm* c = new Comm();
  c->rank = rank;
  c->world = world;
  ncclUniqueId id;
  memcpy(id.internal, id128, sizeof(id.internal));
  const ncclResult_t r = a->CommInitRank(&c->nccl, world, id, rank);
  if (r != ncclSuccess) {
    if (err) *err = std::string("ncclCommInitRank: ") + a->GetErrorString(r);
    delete c;
    return 1;
  }
  *out = c;
  return 0;
}

void comm_destroy(Comm* c) {
  if (!c) return;
  if (c->nccl) api()->CommDestroy(c->nccl);
  if (c->d_stage) cudaFree(c->d_stage);
  if (c->h_stage) cudaFreeHost(c->h_stage);
  delete c;
}

const char* comm_error(const Comm* c) { return c ? c->err.c_str() : "no communicator"; }
int comm_rank(const Comm* c) { return c ? c->rank : 0; }
int comm_world(const Comm* c) { return c ? c->world : 1; }

int comm_nccl_version() {
  Api* a = api();
  int v = 0;
  if (a->lib && a->GetVersion) a->GetVersion(&v);
  return v;
}

/* Row chunks of one group's adjacency (capi.cu::group_shard): rows are cut into 2*world chunks of chunk_bytes;
 * rank r computed chunks r and 2*world-1-r.  The low half is a plain in-place all-gather (rank r's chunk sits at slot r);
 * the high half is stored in mirrored rank order, so it goes as `world` in-place broadcasts fused in one NCCL group. */
int comm_allgather_row_chunks(Comm* c, void* bits, size_t chunk_bytes, cudaStream_t st) {
  Api* a = api();
  char* base = (char*)bits;
  const int W = c->world;
  NCCL_TRY(c, a->AllGather(base + (size_t)c->rank * chunk_bytes, base, chunk_bytes, ncclUint8, c->nccl, st));
  NCCL_TRY(c, a->GroupStart());
  for (int q = 0; q < W; ++q) {
    char* p = base + (size_t)(2 * W - 1 - q) * chunk_bytes;
    const ncclResult_t r = a->Broadcast(p, p, chunk_bytes, ncclUint8, q, c->nccl, st);
    if (r != ncclSuccess) {
      a->GroupEnd();
      c->err = std::string("ncclBroadcast: ") + a->GetErrorString(r);
      return 1;
    }
  }
  NCCL_TRY(c, a->GroupEnd());
  return 0;
}

int comm_allreduce_i64_device(Comm* c, long long* dev, size_t count, bool is_max, cudaStream_t st) {
  NCCL_TRY(c, api()->AllReduce(dev, dev, count, ncclInt64, is_max ? ncclMax : ncclMin, c->nccl, st));
  return 0;
}

int comm_bcast_device(Comm* c, void* dev, size_t bytes, int root, cudaStream_t st) {
  NCCL_TRY(c, api()->Broadcast(dev, dev, bytes, ncclUint8, root, c->nccl, st));
  return 0;
}

static int ensure_stage(Comm* c, size_t bytes) {
  if (bytes <= c->stage_cap) return 0;
  size_t cap = c->stage_cap ? c->stage_cap : 4096;
  while (cap < bytes) cap *= 2;
  if (c->d_stage) cudaFree(c->d_stage);
  if (c->h_stage) cudaFreeHost(c->h_stage);
  c->d_stage = c->h_stage = nullptr;
  c->stage_cap = 0;
  CUDA_TRY(c, cudaMalloc(&c->d_stage, cap));
  CUDA_TRY(c, cudaMallocHost(&c->h_stage, cap));
  c->stage_cap = cap;
  return 0;
}

/* host-buffer forms (blocking): stage through pinned memory, run the collective on `st`, copy back */
int comm_allreduce_i64_host(Comm* c, long long* host, size_t count, bool is_max, cudaStream_t st) {
  const size_t bytes = count * sizeof(long long);
  if (int rc = ensure_stage(c, bytes)) return rc;
  memcpy(c->h_stage, host, bytes);
  CUDA_TRY(c, cudaMemcpyAsync(c->d_stage, c->h_stage, bytes, cudaMemcpyHostToDevice, st));
  if (int rc = comm_allreduce_i64_device(c, (long long*)c->d_stage, count, is_max, st)) return rc;
  CUDA_TRY(c, cudaMemcpyAsync(c->h_stage, c->d_stage, bytes, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(c, cudaStreamSynchronize(st));
  memcpy(host, c->h_stage, bytes);
  return 0;
}

int comm_bcast_host(Comm* c, void* host, size_t bytes, int root, cudaStream_t st) {
  if (int rc = ensure_stage(c, bytes)) return rc;
  if (c->rank == root) {
    memcpy(c->h_stage, host, bytes);
    CUDA_TRY(c, cudaMemcpyAsync(c->d_stage, c->h_stage, bytes, cudaMemcpyHostToDevice, st));
  }
  if (int rc = comm_bcast_device(c, c->d_stage, bytes, root, st)) return rc;
  CUDA_TRY(c, cudaMemcpyAsync(c->h_stage, c->d_stage, bytes, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(c, cudaStreamSynchronize(st));
  memcpy(host, c->h_stage, bytes);
  return 0;
}

}  // namespace rpgo

// The one exchange step of the path, inside the library: the per-cycle all-reduce(sum) of the flat
// gradient buffer across the GPUs of a box (replaces the parameter-server push of the reference,
// main.py:60-62 variable placement + agent.py:321 apply_gradients on the ps).  One process per
// GPU, one communicator per process.  NCCL is bound at run time (dlopen of libnccl.so.2 -- inside
// a PyTorch process that is the copy torch already loaded), so the library itself has no
// link-time dependency and still loads on a machine without NCCL; arl_comm_* then fail loudly.
#include <dlfcn.h>
#include <string.h>

#include "common.cuh"

namespace arl {
namespace {

// the few NCCL prototypes used (stable since NCCL 2.0); ncclUniqueId is 128 opaque bytes passed BY VALUE
struct NcclId { char internal[ARL_COMM_ID_BYTES]; };
typedef void* NcclComm;
typedef int (*GetUniqueIdFn)(NcclId*);
typedef int (*CommInitRankFn)(NcclComm*, int, NcclId, int);
typedef int (*CommDestroyFn)(NcclComm);
typedef int (*AllReduceFn)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t);
typedef const char* (*GetErrorStringFn)(int);
typedef int (*GetVersionFn)(int*);
constexpr int kNcclFloat32 = 7, kNcclSum = 0;

struct Nccl {
  void* handle = nullptr;
  GetUniqueIdFn get_unique_id = nullptr;
  CommInitRankFn comm_init_rank = nullptr;
  CommDestroyFn comm_destroy = nullptr;
  AllReduceFn all_reduce = nullptr;
  GetErrorStringFn error_string = nullptr;
  GetVersionFn get_version = nullptr;
};
Nccl g_nccl;

struct Comm {
  NcclComm comm = nullptr;
  int rank = 0, nranks = 0, device = -1;
  cudaStream_t side = nullptr;           // bucket all-reduces overlap the rest of the backward here
  cudaEvent_t ready = nullptr, done = nullptr;
  bool pending = false;
};
Comm g_comm;

int load_nccl() {
  if (g_nccl.handle) return ARL_OK;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (const char* n : names) {
    h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  if (!h) {
    set_error("arl_comm: cannot load libnccl.so.2 (%s)", dlerror());
    return ARL_ERR_UNSUPPORTED;
  }
  Nccl n;
  n.handle = h;
  n.get_unique_id = (GetUniqueIdFn)dlsym(h, "ncclGetUniqueId");
  n.comm_init_rank = (CommInitRankFn)dlsym(h, "ncclCommInitRank");
  n.comm_destroy = (CommDestroyFn)dlsym(h, "ncclCommDestroy");
  n.all_reduce = (AllReduceFn)dlsym(h, "ncclAllReduce");
  n.error_string = (GetErrorStringFn)dlsym(h, "ncclGetErrorString");
  n.get_version = (GetVersionFn)dlsym(h, "ncclGetVersion");
  if (!n.get_unique_id || !n.comm_init_rank || !n.comm_destroy || !n.all_reduce || !n.error_string) {
    set_error("arl_comm: libnccl.so.2 lacks an expected symbol");
    return ARL_ERR_UNSUPPORTED;
  }
  g_nccl = n;
  return ARL_OK;
}

int nccl_fail(int rc, const char* what) {
  set_error("NCCL error %d (%s) at %s", rc, g_nccl.error_string ? g_nccl.error_string(rc) : "?", what);
  return ARL_ERR_CUDA;
}
#define ARL_NCCL(expr)                                   \
  do {                                                   \
    const int r__ = (expr);                              \
    if (r__ != 0) return nccl_fail(r__, #expr);          \
  } while (0)

}  // namespace
}  // namespace arl

using namespace arl;

extern "C" int arl_comm_unique_id(uint8_t* id_out) {
  ARL_REQUIRE(id_out, "arl_comm_unique_id: null pointer");
  int rc = load_nccl();
  if (rc) return rc;
  NcclId id;
  ARL_NCCL(g_nccl.get_unique_id(&id));
  memcpy(id_out, id.internal, ARL_COMM_ID_BYTES);
  return ARL_OK;
}

extern "C" int arl_comm_init(const uint8_t* id_bytes, int rank, int nranks) {
  ARL_REQUIRE(id_bytes, "arl_comm_init: null pointer");
  ARL_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "arl_comm_init: rank %d not in [0,%d)", rank, nranks);
  ARL_REQUIRE(g_comm.comm == nullptr, "arl_comm_init: a communicator already exists (one per process)");
  int rc = load_nccl();
  if (rc) return rc;
  NcclId id;
  memcpy(id.internal, id_bytes, ARL_COMM_ID_BYTES);
  Comm c;
  ARL_CUDA(cudaGetDevice(&c.device));
  ARL_NCCL(g_nccl.comm_init_rank(&c.comm, nranks, id, rank));
  c.rank = rank;
  c.nranks = nranks;
  ARL_CUDA(cudaStreamCreateWithFlags(&c.side, cudaStreamNonBlocking));
  ARL_CUDA(cudaEventCreateWithFlags(&c.ready, cudaEventDisableTiming));
  ARL_CUDA(cudaEventCreateWithFlags(&c.done, cudaEventDisableTiming));
  g_comm = c;
  return ARL_OK;
}

extern "C" int arl_comm_size(void) { return g_comm.comm ? g_comm.nranks : 0; }

extern "C" int arl_comm_nccl_version(void) {
  int v = 0;
  if (load_nccl() != ARL_OK || !g_nccl.get_version || g_nccl.get_version(&v) != 0) return 0;
  return v;
}

extern "C" int arl_comm_destroy(void) {
  if (!g_comm.comm) return ARL_OK;
  cudaStreamSynchronize(g_comm.side);
  g_nccl.comm_destroy(g_comm.comm);
  cudaEventDestroy(g_comm.ready);
  cudaEventDestroy(g_comm.done);
  cudaStreamDestroy(g_comm.side);
  g_comm = Comm();
  return ARL_OK;
}

extern "C" int arl_allreduce_grads(float* grads, int64_t count, void* stream) {
  ARL_REQUIRE(grads && count >= 0, "arl_allreduce_grads: bad arguments");
  ARL_REQUIRE(g_comm.comm, "arl_allreduce_grads: arl_comm_init has not been called");
  if (count == 0 || g_comm.nranks == 1) return ARL_OK;
  ARL_NCCL(g_nccl.all_reduce(grads, grads, (size_t)count, kNcclFloat32, kNcclSum, g_comm.comm,
                             (cudaStream_t)stream));
  count_launch();
  return ARL_OK;
}

// Bucketed form: `begin` queues the all-reduce of grads[offset, offset+count) on the library's
// side stream, ordered after everything already queued on `stream` (the kernels that produced
// that slice); the caller keeps launching the rest of the backward on `stream`.  `end` makes
// `stream` wait for every bucket begun since the last `end`.
extern "C" int arl_allreduce_begin(float* grads, int64_t offset, int64_t count, void* stream) {
  ARL_REQUIRE(grads && offset >= 0 && count >= 0, "arl_allreduce_begin: bad arguments");
  ARL_REQUIRE(g_comm.comm, "arl_allreduce_begin: arl_comm_init has not been called");
  if (count == 0 || g_comm.nranks == 1) return ARL_OK;
  ARL_CUDA(cudaEventRecord(g_comm.ready, (cudaStream_t)stream));
  ARL_CUDA(cudaStreamWaitEvent(g_comm.side, g_comm.ready, 0));
  ARL_NCCL(g_nccl.all_reduce(grads + offset, grads + offset, (size_t)count, kNcclFloat32, kNcclSum,
                             g_comm.comm, g_comm.side));
  count_launch();
  g_comm.pending = true;
  return ARL_OK;
}

extern "C" int arl_allreduce_end(void* stream) {
  if (!g_comm.comm || !g_comm.pending) return ARL_OK;
  ARL_CUDA(cudaEventRecord(g_comm.done, g_comm.side));
  ARL_CUDA(cudaStreamWaitEvent((cudaStream_t)stream, g_comm.done, 0));
  g_comm.pending = false;
  return ARL_OK;
}

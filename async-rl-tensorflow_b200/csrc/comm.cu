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
typedef int (*AllGatherFn)(const void*, void*, size_t, int, NcclComm, cudaStream_t);
constexpr int kNcclInt8 = 0;
constexpr int kNcclFloat32 = 7, kNcclSum = 0;

struct Nccl {
  void* handle = nullptr;
  GetUniqueIdFn get_unique_id = nullptr;
  CommInitRankFn comm_init_rank = nullptr;
  CommDestroyFn comm_destroy = nullptr;
  AllReduceFn all_reduce = nullptr;
  GetErrorStringFn error_string = nullptr;
  GetVersionFn get_version = nullptr;
  AllGatherFn all_gather = nullptr;
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

// peer-memory exchange (arl_comm_enable_p2p)
struct P2P {
  bool on = false;
  void* local = nullptr;                 // this rank's cudaMalloc'd buffer
  void* opened[kMaxRanks] = {nullptr};   // peers' buffers as mapped here (cudaIpcOpenMemHandle)
  P2PView view;
};
P2P g_p2p;

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
  n.all_gather = (AllGatherFn)dlsym(h, "ncclAllGather");
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

const P2PView* p2p_view() { return g_p2p.on ? &g_p2p.view : nullptr; }

}  // namespace arl

using namespace arl;

// One-shot all-reduce over NVLink peer memory, fused into the update (update.cu): every rank
// publishes its gradient into one of two slots of its own IPC-shared buffer and raises a flag word
// in every peer's memory; the norm pass of the update then READS all ranks' slots (plain loads over
// NVLink), adds them in rank order -- bit-identical on every rank -- writes the summed gradient
// locally and accumulates the per-tensor sums of squares in the same pass.  No NCCL kernel, no
// extra launch, nothing for the persistent backward kernels to wait behind.  Two slots alternate so
// that no second barrier is needed: a rank overwrites slot p two cycles later, after its own
// reduction of the cycle in between has seen every peer's flag for that cycle.
extern "C" int arl_comm_enable_p2p(int64_t count) {
  ARL_REQUIRE(g_comm.comm, "arl_comm_enable_p2p: arl_comm_init has not been called");
  ARL_REQUIRE(count > 0, "arl_comm_enable_p2p: count must be > 0");
  ARL_REQUIRE(g_comm.nranks <= kMaxRanks, "arl_comm_enable_p2p: at most %d ranks", kMaxRanks);
  ARL_REQUIRE(g_nccl.all_gather, "arl_comm_enable_p2p: libnccl lacks ncclAllGather");
  if (g_p2p.on) {
    ARL_REQUIRE(g_p2p.view.count == count, "arl_comm_enable_p2p: already enabled for %lld floats",
                (long long)g_p2p.view.count);
    return ARL_OK;
  }
  const int n = g_comm.nranks, me = g_comm.rank;
  const size_t slot_bytes = ((size_t)count * sizeof(float) + 255) / 256 * 256;
  const size_t ctl_off = 2 * slot_bytes, bytes = ctl_off + 4096;
  ARL_CUDA(cudaMalloc(&g_p2p.local, bytes));
  ARL_CUDA(cudaMemset(g_p2p.local, 0, bytes));
  // exchange the IPC handles through the communicator that already exists
  cudaIpcMemHandle_t mine;
  ARL_CUDA(cudaIpcGetMemHandle(&mine, g_p2p.local));
  void* stage = nullptr;
  ARL_CUDA(cudaMalloc(&stage, (size_t)(n + 1) * sizeof(mine)));
  char* all_dev = (char*)stage + sizeof(mine);
  ARL_CUDA(cudaMemcpy(stage, &mine, sizeof(mine), cudaMemcpyHostToDevice));
  ARL_NCCL(g_nccl.all_gather(stage, all_dev, sizeof(mine), kNcclInt8, g_comm.comm, g_comm.side));
  ARL_CUDA(cudaStreamSynchronize(g_comm.side));
  cudaIpcMemHandle_t all[kMaxRanks];
  ARL_CUDA(cudaMemcpy(all, all_dev, (size_t)n * sizeof(mine), cudaMemcpyDeviceToHost));
  ARL_CUDA(cudaFree(stage));
  P2PView v = {};
  for (int r = 0; r < n; ++r) {
    void* base = g_p2p.local;
    if (r != me) {
      ARL_CUDA(cudaIpcOpenMemHandle(&base, all[r], cudaIpcMemLazyEnablePeerAccess));
      g_p2p.opened[r] = base;
    }
    v.slot[r] = (float*)base;
    v.flags[r] = (unsigned long long*)((char*)base + ctl_off);
  }
  char* ctl = (char*)g_p2p.local + ctl_off;
  v.cycle = (unsigned long long*)(ctl + 1024);
  v.done = (unsigned int*)(ctl + 1024 + 64);
  v.error = (int*)(ctl + 1024 + 128);
  v.count = (long long)count;
  v.stride = (long long)(slot_bytes / sizeof(float));
  v.rank = me; v.nranks = n;
  g_p2p.view = v;
  g_p2p.on = true;
  // nobody publishes before every rank has mapped every buffer: one tiny collective as a barrier
  void* tok = nullptr;
  ARL_CUDA(cudaMalloc(&tok, 256));
  ARL_CUDA(cudaMemset(tok, 0, 256));
  ARL_NCCL(g_nccl.all_reduce(tok, tok, 1, kNcclFloat32, kNcclSum, g_comm.comm, g_comm.side));
  ARL_CUDA(cudaStreamSynchronize(g_comm.side));
  ARL_CUDA(cudaFree(tok));
  return ARL_OK;
}

/* 1 when the peer-memory exchange is active */
extern "C" int arl_comm_p2p_enabled(void) { return g_p2p.on ? 1 : 0; }

/* the error word of the peer-memory exchange: 0 = fine, 1 = a peer's flag did not arrive within
 * the time limit of the reduction kernel (its result is then garbage; the caller should stop) */
extern "C" int arl_comm_p2p_error(void) {
  if (!g_p2p.on) return 0;
  int e = 0;
  if (cudaMemcpy(&e, g_p2p.view.error, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  return e;
}

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
  if (g_p2p.on) {
    cudaDeviceSynchronize();
    for (int r = 0; r < kMaxRanks; ++r)
      if (g_p2p.opened[r]) cudaIpcCloseMemHandle(g_p2p.opened[r]);
    cudaFree(g_p2p.local);
    g_p2p = P2P();
  }
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

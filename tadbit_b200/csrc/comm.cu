// Multi-GPU exchange: one process per GPU, one NCCL communicator over NVLink/NVSwitch.
// NCCL is resolved at run time (dlopen) so single-GPU use never needs it and the library
// shares the libnccl.so.2 that the host process (e.g. PyTorch) has already loaded.
//
// The path has ONE exchange per ICE pass: the row-sum vector s.  By default it does not go
// through a collective at all: every rank maps the row-sum buffers of its peers (CUDA IPC,
// exchange_setup below) and k_ice_exchange (kernels.cu) stores each rank's rows straight into
// every peer's buffer over NVLink, followed by a flag barrier.  NCCL carries the set-up (IPC
// handles), the integer reductions of the filters / distance sums, and -- TB_EXCHANGE=nccl, or
// when the peers cannot be mapped -- the row sums as an all-reduce (each rank fills the rows it
// owns, zeros elsewhere; x + 0 is exact, so every rank ends with identical bits either way and
// takes the same loop decision without a second exchange).
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace tb {

namespace {

typedef struct { char internal[128]; } NcclUniqueId;
typedef void *NcclComm;
enum { kNcclInt8 = 0, kNcclInt64 = 4, kNcclFloat64 = 8, kNcclSum = 0 };

struct NcclApi {
  void *lib = nullptr;
  int (*GetUniqueId)(NcclUniqueId *) = nullptr;
  int (*CommInitRank)(NcclComm *, int, NcclUniqueId, int) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  int (*AllReduce)(const void *, void *, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*AllGather)(const void *, void *, size_t, int, NcclComm, cudaStream_t) = nullptr;
  int (*CommAbort)(NcclComm) = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
};

NcclApi &api() {
  static NcclApi a;
  if (a.lib) return a;
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char *n : names) {
    a.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (a.lib) break;
  }
  if (!a.lib) {
    set_error("NCCL not loadable: %s", dlerror());
    throw Fail{TB_ERR_NCCL};
  }
  a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(a.lib, "ncclGetUniqueId");
  a.CommInitRank = (decltype(a.CommInitRank))dlsym(a.lib, "ncclCommInitRank");
  a.CommDestroy = (decltype(a.CommDestroy))dlsym(a.lib, "ncclCommDestroy");
  a.AllReduce = (decltype(a.AllReduce))dlsym(a.lib, "ncclAllReduce");
  a.AllGather = (decltype(a.AllGather))dlsym(a.lib, "ncclAllGather");
  a.CommAbort = (decltype(a.CommAbort))dlsym(a.lib, "ncclCommAbort");
  a.GetErrorString = (decltype(a.GetErrorString))dlsym(a.lib, "ncclGetErrorString");
  if (!a.GetUniqueId || !a.CommInitRank || !a.CommDestroy || !a.AllReduce) {
    set_error("NCCL symbols missing");
    throw Fail{TB_ERR_NCCL};
  }
  return a;
}

void check(int rc, const char *what) {
  if (rc == 0) return;
  NcclApi &a = api();
  set_error("NCCL %s failed: %s", what, a.GetErrorString ? a.GetErrorString(rc) : "?");
  throw Fail{TB_ERR_NCCL};
}

// a failed collective leaves the peers blocked inside it: abort the communicator so that they
// fail too instead of hanging (the group has to be torn down and rebuilt by the caller)
void check_abort(tb_store *s, int rc, const char *what) {
  if (rc == 0) return;
  NcclApi &a = api();
  if (a.CommAbort && s->comm) {
    a.CommAbort((NcclComm)s->comm);
    s->comm = nullptr;
    if (s->group) s->group->nccl = nullptr;
  }
  check(rc, what);
}

}  // namespace

void comm_unique_id(uint8_t id[128]) {
  NcclUniqueId u;
  check(api().GetUniqueId(&u), "GetUniqueId");
  memcpy(id, u.internal, 128);
}

// A group = one NCCL communicator of this process (one rank per GPU).  Creating it costs most of a
// second, so it can outlive the stores: a long-running process creates it once and attaches every
// matrix it normalises (tb_comm_create / tb_store_attach); tb_comm_init is the one-shot form that
// creates a group owned by the store.
tb_comm *group_create(int rank, int world, const uint8_t id[128], int device) {
  TB_REQUIRE(world >= 1 && rank >= 0 && rank < world, "comm: bad rank/world");
  TB_REQUIRE(world <= kMaxPeers, "at most %d GPUs per group", kMaxPeers);
  tb_comm *g = new tb_comm();
  g->rank = rank;
  g->world = world;
  g->device = device;
  if (world > 1) {
    NcclUniqueId u;
    memcpy(u.internal, id, 128);
    NcclComm c = nullptr;
    int rc = api().CommInitRank(&c, world, u, rank);
    if (rc != 0) {
      delete g;
      check(rc, "CommInitRank");
    }
    g->nccl = c;
  }
  return g;
}

void group_destroy(tb_comm *g) {
  if (!g) return;
  if (g->nccl) api().CommDestroy((NcclComm)g->nccl);
  delete g;
}

void store_attach(tb_store *s, tb_comm *g, bool owned) {
  TB_REQUIRE(g && g->device == s->device, "attach: the group lives on another device");
  comm_destroy(s);
  s->group = g;
  s->owns_group = owned;
  s->rank = g->rank;
  s->world = g->world;
  s->comm = g->nccl;
  exchange_setup(s);
}

void comm_init(tb_store *s, int rank, int world, const uint8_t id[128]) {
  comm_destroy(s);
  tb_comm *g = group_create(rank, world, id, s->device);
  try {
    store_attach(s, g, true);
  } catch (...) {
    if (s->group != g) group_destroy(g);
    throw;
  }
}

// Row-sum exchange buffers (common.cuh:PeerExchange).  Plain cudaMalloc memory (the stream-ordered
// pool cannot be exported); with several ranks every rank's block is opened by all the others
// through CUDA IPC.  Whether that worked is agreed on by all ranks (an all-reduce of the failure
// count): either every rank uses the peer-mapped exchange or every rank uses the all-reduce.
void exchange_setup(tb_store *s) {
  exchange_release(s);
  PeerExchange &px = s->px;
  TB_REQUIRE(s->world <= kMaxPeers, "at most %d GPUs per group", kMaxPeers);
  px.stride = (((size_t)s->nbins + 63) / 64) * 64;
  // an entry is 16 bytes: two tagged words of the LL protocol (the flag protocol uses 8 of them)
  const size_t flag_off = 2 * px.stride * 2 * sizeof(double);
  const size_t bytes = flag_off + kMaxPeers * sizeof(unsigned long long);
  TB_CUDA(cudaMalloc(&px.base, bytes));
  TB_CUDA(cudaMemset(px.base, 0, bytes));
  px.epoch = 0;
  PeerView &v = px.view;
  v = PeerView{};
  v.rank = s->rank;
  v.world = 1;                                   // until the peers are mapped
  auto place = [&](int q, void *base) {
    v.s[q][0] = reinterpret_cast<double *>(base);
    v.s[q][1] = reinterpret_cast<double *>(base) + 2 * px.stride;
    v.flag[q] = reinterpret_cast<unsigned long long *>(reinterpret_cast<uint8_t *>(base) + flag_off);
  };
  // one GPU: the only "peer" is the rank itself, slot 0
  v.rank = s->world == 1 ? 0 : s->rank;
  place(v.rank, px.base);
  v.my_flags = v.flag[v.rank];
  v.mine[0] = v.s[v.rank][0];
  v.mine[1] = v.s[v.rank][1];
  px.ready = true;
  if (s->world == 1) return;
  const char *e = getenv("TB_EXCHANGE");
  const bool want = !(e && !strcmp(e, "nccl"));
  // every rank takes part in the two collectives below whatever it wants, so nobody waits alone
  cudaIpcMemHandle_t mine;
  memset(&mine, 0, sizeof(mine));
  long long failed = 0;
  if (want && cudaIpcGetMemHandle(&mine, px.base) != cudaSuccess) { cudaGetLastError(); failed = 1; }
  if (!want) failed = 1;
  DevBuf<uint8_t> d_handles;
  DevBuf<long long> d_fail;
  const size_t hb = sizeof(cudaIpcMemHandle_t);
  d_handles.alloc(hb * (s->world + 1));
  d_fail.alloc(2);
  TB_CUDA(cudaMemcpyAsync(d_handles.p + hb * s->world, &mine, hb, cudaMemcpyHostToDevice, s->stream));
  check_abort(s, api().AllGather ? api().AllGather(d_handles.p + hb * s->world, d_handles.p, hb, kNcclInt8,
                                                    (NcclComm)s->comm, s->stream) : 0, "AllGather(ipc handles)");
  std::vector<cudaIpcMemHandle_t> all(s->world);
  TB_CUDA(cudaMemcpyAsync(all.data(), d_handles.p, hb * s->world, cudaMemcpyDeviceToHost, s->stream));
  TB_CUDA(cudaStreamSynchronize(s->stream));
  if (!api().AllGather) failed = 1;
  if (!failed) {
    for (int q = 0; q < s->world && !failed; ++q) {
      if (q == s->rank) continue;
      void *pb = nullptr;
      if (cudaIpcOpenMemHandle(&pb, all[q], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        cudaGetLastError();
        failed = 1;
        break;
      }
      px.peer_base[q] = pb;
    }
  }
  TB_CUDA(cudaMemcpyAsync(d_fail.p, &failed, sizeof(long long), cudaMemcpyHostToDevice, s->stream));
  allreduce_i64(s, d_fail.p, d_fail.p + 1, 1);
  long long any_failed = 0;
  TB_CUDA(cudaMemcpyAsync(&any_failed, d_fail.p + 1, sizeof(long long), cudaMemcpyDeviceToHost, s->stream));
  TB_CUDA(cudaStreamSynchronize(s->stream));
  if (any_failed) {
    for (int q = 0; q < kMaxPeers; ++q)
      if (px.peer_base[q]) { cudaIpcCloseMemHandle(px.peer_base[q]); px.peer_base[q] = nullptr; }
    if (want && s->rank == 0)
      fprintf(stderr, "[tadbit_b200] peer-mapped row-sum exchange not available on this box "
                      "(CUDA IPC failed on %lld rank(s)); using the NCCL all-reduce\n", any_failed);
    return;
  }
  for (int q = 0; q < s->world; ++q)
    if (q != s->rank) place(q, px.peer_base[q]);
  v.world = s->world;
  px.mapped = true;
}

void exchange_release(tb_store *s) {
  PeerExchange &px = s->px;
  for (int q = 0; q < kMaxPeers; ++q)
    if (px.peer_base[q]) { cudaIpcCloseMemHandle(px.peer_base[q]); px.peer_base[q] = nullptr; }
  if (px.base) cudaFree(px.base);
  px = PeerExchange{};
}

void comm_destroy(tb_store *s) {
  if (s->px.mapped) {
    // peers may still be reading this rank's flags / writing nothing: all exchanges of a call are
    // complete on every rank before the call returns, so only the mappings are left to drop
    cudaStreamSynchronize(s->stream);
  }
  exchange_release(s);
  if (s->group && s->owns_group) group_destroy(s->group);
  s->group = nullptr;
  s->owns_group = false;
  s->comm = nullptr;
  s->world = 1;
  s->rank = 0;
}

void allreduce_f64(tb_store *s, const double *send, double *recv, size_t n) {
  if (!s->comm) { set_error("communicator not initialised"); throw Fail{TB_ERR_STATE}; }
  check_abort(s, api().AllReduce(send, recv, n, kNcclFloat64, kNcclSum, (NcclComm)s->comm, s->stream),
              "AllReduce(f64)");
}

void allreduce_i64(tb_store *s, const long long *send, long long *recv, size_t n) {
  if (!s->comm) { set_error("communicator not initialised"); throw Fail{TB_ERR_STATE}; }
  check_abort(s, api().AllReduce(send, recv, n, kNcclInt64, kNcclSum, (NcclComm)s->comm, s->stream),
              "AllReduce(i64)");
}

}  // namespace tb

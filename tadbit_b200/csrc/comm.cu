// Multi-GPU exchange: one process per GPU, one NCCL communicator over NVLink/NVSwitch.
// NCCL is resolved at run time (dlopen) so single-GPU use never needs it and the library
// shares the libnccl.so.2 that the host process (e.g. PyTorch) has already loaded.
//
// The path has ONE exchange per ICE pass: the row-sum vector s (each rank fills the rows
// it owns, zeros elsewhere; x + 0 is exact, so every rank ends with identical bits and
// takes the same loop decision without a second collective).
#include <dlfcn.h>
#include <string.h>

#include "common.cuh"

namespace tb {

namespace {

typedef struct { char internal[128]; } NcclUniqueId;
typedef void *NcclComm;
enum { kNcclInt64 = 4, kNcclFloat64 = 8, kNcclSum = 0 };

struct NcclApi {
  void *lib = nullptr;
  int (*GetUniqueId)(NcclUniqueId *) = nullptr;
  int (*CommInitRank)(NcclComm *, int, NcclUniqueId, int) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  int (*AllReduce)(const void *, void *, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
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

}  // namespace

void comm_unique_id(uint8_t id[128]) {
  NcclUniqueId u;
  check(api().GetUniqueId(&u), "GetUniqueId");
  memcpy(id, u.internal, 128);
}

void comm_init(tb_store *s, int rank, int world, const uint8_t id[128]) {
  TB_REQUIRE(world >= 1 && rank >= 0 && rank < world, "comm_init: bad rank/world");
  comm_destroy(s);
  s->rank = rank;
  s->world = world;
  if (world == 1) return;
  NcclUniqueId u;
  memcpy(u.internal, id, 128);
  NcclComm c = nullptr;
  check(api().CommInitRank(&c, world, u, rank), "CommInitRank");
  s->comm = c;
}

void comm_destroy(tb_store *s) {
  if (s->comm) api().CommDestroy((NcclComm)s->comm);
  s->comm = nullptr;
  s->world = 1;
  s->rank = 0;
}

void allreduce_f64(tb_store *s, const double *send, double *recv, size_t n) {
  if (!s->comm) { set_error("communicator not initialised"); throw Fail{TB_ERR_STATE}; }
  check(api().AllReduce(send, recv, n, kNcclFloat64, kNcclSum, (NcclComm)s->comm, s->stream),
        "AllReduce(f64)");
}

void allreduce_i64(tb_store *s, const long long *send, long long *recv, size_t n) {
  if (!s->comm) { set_error("communicator not initialised"); throw Fail{TB_ERR_STATE}; }
  check(api().AllReduce(send, recv, n, kNcclInt64, kNcclSum, (NcclComm)s->comm, s->stream),
        "AllReduce(i64)");
}

}  // namespace tb

// allreduce.cu -- the data-parallel exchange step behind the C ABI: one NCCL communicator per process,
// in-place fp32 sum all-reduce of a slice of the flat gradient buffer on a caller-chosen stream.
//
// NCCL is bound at run time (dlsym): the library is the one PyTorch ships and has already loaded into the process
// (nvidia/nccl/lib/libnccl.so.2), so there is no link-time dependency and the CPU-only build needs no NCCL headers.
// The few prototypes used are restated from nccl.h (2.x ABI: ncclUniqueId is 128 opaque bytes passed BY VALUE to
// ncclCommInitRank; ncclFloat32 = 7; ncclSum = 0).
#include <dlfcn.h>
#include <string.h>
#include "common.cuh"

namespace {

struct UniqueId { char internal[128]; };
typedef void* Comm;
typedef int (*GetUniqueIdFn)(UniqueId*);
typedef int (*CommInitRankFn)(Comm*, int, UniqueId, int);
typedef int (*AllReduceFn)(const void*, void*, size_t, int, int, Comm, cudaStream_t);
typedef int (*CommDestroyFn)(Comm);
typedef const char* (*GetErrorStringFn)(int);

struct Api {
  GetUniqueIdFn get_unique_id = nullptr;
  CommInitRankFn comm_init_rank = nullptr;
  AllReduceFn all_reduce = nullptr;
  CommDestroyFn comm_destroy = nullptr;
  GetErrorStringFn error_string = nullptr;
  bool ok = false;
};

void* find(const char* name) {
  void* p = dlsym(RTLD_DEFAULT, name);                 // already loaded by PyTorch (and made global by the host side)
  if (!p) {
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (h) p = dlsym(h, name);
  }
  return p;
}

Api& api() {
  static Api a = [] {
    Api x;
    x.get_unique_id = reinterpret_cast<GetUniqueIdFn>(find("ncclGetUniqueId"));
    x.comm_init_rank = reinterpret_cast<CommInitRankFn>(find("ncclCommInitRank"));
    x.all_reduce = reinterpret_cast<AllReduceFn>(find("ncclAllReduce"));
    x.comm_destroy = reinterpret_cast<CommDestroyFn>(find("ncclCommDestroy"));
    x.error_string = reinterpret_cast<GetErrorStringFn>(find("ncclGetErrorString"));
    x.ok = x.get_unique_id && x.comm_init_rank && x.all_reduce && x.comm_destroy;
    return x;
  }();
  return a;
}

int fail(const char* what, int rc) {
  Api& a = api();
  myolo::set_error("%s: NCCL error %d (%s)", what, rc, a.error_string ? a.error_string(rc) : "?");
  return MYOLO_ERR_CUDA;
}

constexpr int kNcclFloat32 = 7, kNcclSum = 0;

}  // namespace

#define MYOLO_NEED_NCCL()                                                                                    \
  do {                                                                                                       \
    if (!api().ok) {                                                                                         \
      myolo::set_error("libnccl.so.2 is not loaded in this process (import torch first, or set LD_LIBRARY_PATH)"); \
      return MYOLO_ERR_CUDA;                                                                                 \
    }                                                                                                        \
  } while (0)

extern "C" int myolo_allreduce_unique_id(char* id128) {
  MYOLO_CHECK_ARG(id128);
  MYOLO_NEED_NCCL();
  UniqueId id;
  const int rc = api().get_unique_id(&id);
  if (rc) return fail("ncclGetUniqueId", rc);
  memcpy(id128, id.internal, sizeof(id.internal));
  return MYOLO_OK;
}

extern "C" int myolo_allreduce_init(const char* id128, int rank, int world, void** comm_out) {
  MYOLO_CHECK_ARG(id128 && comm_out && world >= 1 && rank >= 0 && rank < world);
  MYOLO_NEED_NCCL();
  UniqueId id;
  memcpy(id.internal, id128, sizeof(id.internal));
  Comm c = nullptr;
  const int rc = api().comm_init_rank(&c, world, id, rank);       // uses the calling thread's current CUDA device
  if (rc) return fail("ncclCommInitRank", rc);
  *comm_out = c;
  return MYOLO_OK;
}

extern "C" int myolo_allreduce_run(void* comm, float* buf, long long n, myolo_stream stream) {
  MYOLO_CHECK_ARG(comm && buf && n > 0);
  MYOLO_NEED_NCCL();
  const int rc = api().all_reduce(buf, buf, (size_t)n, kNcclFloat32, kNcclSum, comm, myolo::as_stream(stream));
  if (rc) return fail("ncclAllReduce", rc);
  return MYOLO_OK;
}

extern "C" int myolo_allreduce_destroy(void* comm) {
  MYOLO_CHECK_ARG(comm);
  MYOLO_NEED_NCCL();
  const int rc = api().comm_destroy(comm);
  if (rc) return fail("ncclCommDestroy", rc);
  return MYOLO_OK;
}

// The one collective of the path: sum all-reduce of the 112-double Gram pack [Y^T Y | Y^T f | f^T f | n] across the ranks
// that each accumulated a shard (the reference has no counterpart, loggers/loggers.py:129 is single-process).
//
// NCCL is bound at run time with dlopen/dlsym (no link-time dependency, so the library loads on machines without NCCL and the
// process shares whatever libnccl.so.2 is already resident, e.g. the one PyTorch bundles).  The communicator is created from
// a 128-byte unique id that the host application distributes by any means it likes (torch.distributed broadcast in
// rigid_body_manipulation_b200/distributed.py).
#include <dlfcn.h>

#include <cstdlib>
#include <cstring>

#include "rbm_internal.h"

namespace {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;                 // ncclSuccess == 0
enum { kNcclDouble = 8, kNcclSum = 0 };   // ncclFloat64, ncclSum (nccl.h, stable across 2.x)

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string error;
};

NcclApi& api() {
  static NcclApi a = [] {
    NcclApi x;
    const char* override_name = getenv("RBM_NCCL_LIB");  // e.g. a path that does not exist: exercises the NCCL-less behaviour
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    if (override_name && override_name[0]) {
      x.handle = dlopen(override_name, RTLD_NOW | RTLD_GLOBAL);
    } else {
      for (const char* n : names) {
        x.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (x.handle) break;
      }
    }
    if (!x.handle) {
      const char* e = dlerror();  // one call: dlerror() clears the message it returns
      x.error = std::string("cannot dlopen libnccl.so.2: ") + (e ? e : "unknown");
      return x;
    }
    auto sym = [&](const char* s) { return dlsym(x.handle, s); };
    x.GetUniqueId = reinterpret_cast<decltype(x.GetUniqueId)>(sym("ncclGetUniqueId"));
    x.CommInitRank = reinterpret_cast<decltype(x.CommInitRank)>(sym("ncclCommInitRank"));
    x.CommDestroy = reinterpret_cast<decltype(x.CommDestroy)>(sym("ncclCommDestroy"));
    x.AllReduce = reinterpret_cast<decltype(x.AllReduce)>(sym("ncclAllReduce"));
    x.GetErrorString = reinterpret_cast<decltype(x.GetErrorString)>(sym("ncclGetErrorString"));
    if (!x.GetUniqueId || !x.CommInitRank || !x.CommDestroy || !x.AllReduce) x.error = "libnccl is missing a required symbol";
    return x;
  }();
  return a;
}

int nccl_fail(const char* what, ncclResult_t r) {
  NcclApi& a = api();
  rbm::set_error(std::string("NCCL error in ") + what + ": " + (a.GetErrorString ? a.GetErrorString(r) : "?"));
  return RBM_ERR_NCCL;
}

int nccl_ready() {
  NcclApi& a = api();
  if (!a.error.empty()) {
    rbm::set_error(a.error);
    return RBM_ERR_NCCL;
  }
  return RBM_OK;
}

}  // namespace

extern "C" {

int rbm_nccl_available(void) { return nccl_ready() == RBM_OK ? 1 : 0; }

int rbm_nccl_unique_id(void* id128) {
  if (!id128) {
    rbm::set_error("rbm_nccl_unique_id: NULL output");
    return RBM_ERR_INVALID;
  }
  if (int rc = nccl_ready()) return rc;
  ncclUniqueId id;
  ncclResult_t r = api().GetUniqueId(&id);
  if (r != 0) return nccl_fail("ncclGetUniqueId", r);
  std::memcpy(id128, id.internal, 128);
  return RBM_OK;
}

int rbm_nccl_comm_create(const void* id128, int nranks, int rank, int device, void** comm) {
  if (!id128 || !comm || nranks < 1 || rank < 0 || rank >= nranks) {
    rbm::set_error("rbm_nccl_comm_create: bad argument");
    return RBM_ERR_INVALID;
  }
  if (int rc = nccl_ready()) return rc;
  rbm::DeviceGuard guard(device);
  RBM_CUDA_TRY(guard.status());
  ncclUniqueId id;
  std::memcpy(id.internal, id128, 128);
  ncclComm_t c = nullptr;
  ncclResult_t r = api().CommInitRank(&c, nranks, id, rank);
  if (r != 0) return nccl_fail("ncclCommInitRank", r);
  *comm = c;
  return RBM_OK;
}

int rbm_nccl_comm_destroy(void* comm) {
  if (!comm) return RBM_OK;
  if (int rc = nccl_ready()) return rc;
  ncclResult_t r = api().CommDestroy(static_cast<ncclComm_t>(comm));
  return r == 0 ? RBM_OK : nccl_fail("ncclCommDestroy", r);
}

int rbm_allreduce_gram_n(void* comm, double* gram_packs, int64_t count, void* stream) {
  if (!comm || !gram_packs || count < 1) {
    rbm::set_error("rbm_allreduce_gram: NULL communicator / pack or count < 1");
    return RBM_ERR_INVALID;
  }
  if (int rc = nccl_ready()) return rc;
  ncclResult_t r = api().AllReduce(gram_packs, gram_packs, (size_t)count * 112, kNcclDouble, kNcclSum, static_cast<ncclComm_t>(comm),
                                   static_cast<cudaStream_t>(stream));
  return r == 0 ? RBM_OK : nccl_fail("ncclAllReduce", r);
}

int rbm_allreduce_gram(void* comm, double* gram_pack, void* stream) { return rbm_allreduce_gram_n(comm, gram_pack, 1, stream); }

}  // extern "C"

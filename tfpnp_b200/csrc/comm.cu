// NCCL behind the C ABI: the one exchange of the data path -- an all-gather of the per-image PSNR vector
// (tfpnp/env/base.py:237-242) after the last iteration -- for hosts that do not run torch.distributed.  NCCL is resolved
// at run time with dlopen("libnccl.so.2") (the copy PyTorch ships is already mapped when the Python host is used), so the
// library has no link-time dependency on it and single-GPU users never touch it.
#include "common.cuh"
#include <dlfcn.h>
#include <cstring>
#include <mutex>

namespace tfpnp {
namespace {

typedef struct { char internal[128]; } NcclUniqueId;     // ncclUniqueId: 128 opaque bytes (NCCL_UNIQUE_ID_BYTES)
typedef void* NcclComm;
enum { kNcclFloat32 = 7 };                                // ncclDataType_t: ncclFloat32 = ncclFloat = 7

struct NcclApi {
  void* handle = nullptr;
  int (*GetUniqueId)(NcclUniqueId*) = nullptr;
  int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok = false;
};

NcclApi& nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (api.handle) break;
    }
    if (!api.handle) return;
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(api.handle, "ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(api.handle, "ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(api.handle, "ncclCommDestroy"));
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(dlsym(api.handle, "ncclAllGather"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(api.handle, "ncclGetErrorString"));
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllGather;
  });
  return api;
}

struct Comm {
  NcclComm comm = nullptr;
  int rank = 0, world = 1;
};

int nccl_fail(const char* what, int rc) {
  NcclApi& a = nccl();
  set_error("%s failed: %s", what, a.GetErrorString ? a.GetErrorString(rc) : "NCCL error");
  return TFPNP_ERR_CUDA;
}

}  // namespace
}  // namespace tfpnp

using namespace tfpnp;

extern "C" {

int tfpnp_comm_unique_id(void* id_out, size_t id_bytes) {
  TFPNP_CHECK(id_out && id_bytes >= sizeof(NcclUniqueId), "tfpnp_comm_unique_id needs a %zu-byte buffer", sizeof(NcclUniqueId));
  NcclApi& a = nccl();
  if (!a.ok) { set_error("libnccl.so.2 not found (dlopen): %s", dlerror() ? dlerror() : "?"); return TFPNP_ERR_UNSUPPORTED; }
  NcclUniqueId id;
  const int rc = a.GetUniqueId(&id);
  if (rc != 0) return nccl_fail("ncclGetUniqueId", rc);
  memcpy(id_out, &id, sizeof(id));
  return 0;
}

int tfpnp_comm_init(const void* id, size_t id_bytes, int rank, int world, void** out) {
  TFPNP_CHECK(id && out && id_bytes >= sizeof(NcclUniqueId) && world >= 1 && rank >= 0 && rank < world, "bad argument");
  NcclApi& a = nccl();
  if (!a.ok) { set_error("libnccl.so.2 not found (dlopen)"); return TFPNP_ERR_UNSUPPORTED; }
  NcclUniqueId uid;
  memcpy(&uid, id, sizeof(uid));
  Comm* c = new Comm();
  c->rank = rank; c->world = world;
  const int rc = a.CommInitRank(&c->comm, world, uid, rank);     // binds to the calling thread's current device
  if (rc != 0) { delete c; return nccl_fail("ncclCommInitRank", rc); }
  *out = c;
  return 0;
}

int tfpnp_comm_destroy(void* h) {
  Comm* c = static_cast<Comm*>(h);
  if (c) {
    if (c->comm) nccl().CommDestroy(c->comm);
    delete c;
  }
  return 0;
}

int tfpnp_comm_allgather_psnr(void* h, const float* local, int n_local, float* out, void* stream) {
  TFPNP_CHECK(h && local && out && n_local > 0, "bad argument");
  Comm* c = static_cast<Comm*>(h);
  const int rc = nccl().AllGather(local, out, (size_t)n_local, kNcclFloat32, c->comm, static_cast<cudaStream_t>(stream));
  if (rc != 0) return nccl_fail("ncclAllGather", rc);
  return 0;
}

}  // extern "C"

// Thin NCCL binding for the row-sharded step (one process per GPU).  libnccl is the copy PyTorch ships
// (nvidia/nccl/lib/libnccl.so.2); it is bound at run time with dlopen so that libvipant_b200.so has no link-time
// dependency on it and single-GPU users never load it.  Only what the hot path needs: communicator setup from a
// unique id (exchanged by the host side, e.g. with torch.distributed), all-gather and all-reduce on a stream.
#include <dlfcn.h>

#include <cstring>

#include "common.cuh"

namespace vpa {

struct NcclApi {
  void* handle = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, ncclUniqueIdBlob, int) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;

static int nccl_fail(const char* what, int rc) {
  return set_error(VPA_E_COMM, "%s failed: %s (%d)", what, g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?", rc);
}

int comm_load(const char* path) {
  if (g_nccl.handle) return 0;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);          // already loaded by the host framework?
  if (!h && path && *path) h = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return set_error(VPA_E_COMM, "cannot load libnccl.so.2 (%s)", dlerror());
#define VPA_SYM(field, name)                                                        \
  g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(h, name));          \
  if (!g_nccl.field) return set_error(VPA_E_COMM, "libnccl: symbol %s not found", name)
  VPA_SYM(GetUniqueId, "ncclGetUniqueId");
  VPA_SYM(CommInitRank, "ncclCommInitRank");
  VPA_SYM(CommDestroy, "ncclCommDestroy");
  VPA_SYM(AllGather, "ncclAllGather");
  VPA_SYM(AllReduce, "ncclAllReduce");
  VPA_SYM(GetErrorString, "ncclGetErrorString");
#undef VPA_SYM
  g_nccl.handle = h;
  return 0;
}

int comm_unique_id(void* out128) {
  if (!g_nccl.handle) return set_error(VPA_E_COMM, "vpa_comm_load has not been called");
  if (int rc = g_nccl.GetUniqueId(out128)) return nccl_fail("ncclGetUniqueId", rc);
  return 0;
}

int comm_init(const void* id128, int rank, int world, void** comm_out) {
  if (!g_nccl.handle) return set_error(VPA_E_COMM, "vpa_comm_load has not been called");
  ncclUniqueIdBlob id;
  memcpy(id.internal, id128, sizeof(id.internal));
  void* comm = nullptr;
  if (int rc = g_nccl.CommInitRank(&comm, world, id, rank)) return nccl_fail("ncclCommInitRank", rc);
  *comm_out = comm;
  return 0;
}

int comm_destroy(void* comm) {
  if (comm && g_nccl.handle) g_nccl.CommDestroy(comm);
  return 0;
}

// dtype: 7 = float32, 9 = bfloat16 (ncclDataType_t); in-place when send == recv + rank * count elements
int comm_all_gather(void* comm, const void* send, void* recv, size_t count, int nccl_dtype, cudaStream_t st) {
  if (int rc = g_nccl.AllGather(send, recv, count, nccl_dtype, comm, st)) return nccl_fail("ncclAllGather", rc);
  return 0;
}
int comm_all_reduce_sum_f32(void* comm, const void* send, void* recv, size_t count, cudaStream_t st) {
  if (int rc = g_nccl.AllReduce(send, recv, count, 7 /* ncclFloat32 */, 0 /* ncclSum */, comm, st)) return nccl_fail("ncclAllReduce", rc);
  return 0;
}

}  // namespace vpa

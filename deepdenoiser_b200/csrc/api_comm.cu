// C-ABI: the one exchange step of data-parallel training - NCCL sum all-reduce of the flat fp32 gradient buffer
// over NVLink.  NCCL is resolved at run time (dlopen of the libnccl.so.2 that PyTorch already loaded, or any on the loader
// path), so libdd_b200.so has no link-time dependency on it and single-GPU use never touches it.  The reference has no
// counterpart (it is single-device TensorFlow); this replaces what a tf.distribute strategy would have done for
// optimizer.minimize (Training.py:700-702).
#include <dlfcn.h>
#include <string.h>

#include "dd_internal.h"

namespace {

typedef struct ncclComm* ncclComm_t;
struct NcclUniqueId { char internal[128]; };
typedef int (*GetUniqueIdFn)(NcclUniqueId*);
typedef int (*CommInitRankFn)(ncclComm_t*, int, NcclUniqueId, int);
typedef int (*AllReduceFn)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t);
typedef int (*CommDestroyFn)(ncclComm_t);
typedef const char* (*GetErrorStringFn)(int);

struct NcclApi {
  void* handle = nullptr;
  GetUniqueIdFn get_unique_id = nullptr;
  CommInitRankFn comm_init_rank = nullptr;
  AllReduceFn all_reduce = nullptr;
  CommDestroyFn comm_destroy = nullptr;
  GetErrorStringFn get_error_string = nullptr;
  bool ok = false;
};

NcclApi& nccl() {
  static NcclApi api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    api.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);          // the copy PyTorch loaded, if any
    if (!api.handle) api.handle = dlopen("libnccl.so.2", RTLD_NOW);
    if (!api.handle) api.handle = dlopen("libnccl.so", RTLD_NOW);
    if (api.handle) {
      api.get_unique_id = reinterpret_cast<GetUniqueIdFn>(dlsym(api.handle, "ncclGetUniqueId"));
      api.comm_init_rank = reinterpret_cast<CommInitRankFn>(dlsym(api.handle, "ncclCommInitRank"));
      api.all_reduce = reinterpret_cast<AllReduceFn>(dlsym(api.handle, "ncclAllReduce"));
      api.comm_destroy = reinterpret_cast<CommDestroyFn>(dlsym(api.handle, "ncclCommDestroy"));
      api.get_error_string = reinterpret_cast<GetErrorStringFn>(dlsym(api.handle, "ncclGetErrorString"));
      api.ok = api.get_unique_id && api.comm_init_rank && api.all_reduce && api.comm_destroy && api.get_error_string;
    }
  }
  return api;
}

constexpr int kNcclFloat32 = 7, kNcclSum = 0;   // ncclDataType_t / ncclRedOp_t values of nccl.h (stable ABI since NCCL 2.0)

int nccl_check(int rc, const char* what) {
  if (rc == 0) return DD_OK;
  dd::set_error("%s failed: %s", what, nccl().get_error_string ? nccl().get_error_string(rc) : "NCCL error");
  return DD_ERR_CUDA;
}

}  // namespace

struct dd_comm {
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1, device = 0;
};

extern "C" {

int dd_comm_unique_id(void* id128) {
  if (!id128) { dd::set_error("dd_comm_unique_id: NULL buffer"); return DD_ERR_INVALID; }
  if (!nccl().ok) { dd::set_error("NCCL (libnccl.so.2) could not be loaded"); return DD_ERR_UNSUPPORTED; }
  NcclUniqueId id;
  int rc = nccl_check(nccl().get_unique_id(&id), "ncclGetUniqueId");
  if (rc) return rc;
  memcpy(id128, &id, sizeof(id));
  return DD_OK;
}

int dd_comm_init(dd_ctx* ctx, const void* id128, int rank, int world, dd_comm** out) {
  if (!ctx || !id128 || !out || world < 1 || rank < 0 || rank >= world) { dd::set_error("dd_comm_init: bad argument"); return DD_ERR_INVALID; }
  if (!nccl().ok) { dd::set_error("NCCL (libnccl.so.2) could not be loaded"); return DD_ERR_UNSUPPORTED; }
  DD_CUDA(cudaSetDevice(ctx->device));
  NcclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  dd_comm* c = new dd_comm();
  c->rank = rank; c->world = world; c->device = ctx->device;
  int rc = nccl_check(nccl().comm_init_rank(&c->comm, world, id, rank), "ncclCommInitRank");
  if (rc) { delete c; return rc; }
  *out = c;
  return DD_OK;
}

int dd_comm_allreduce_sum_f32(dd_comm* comm, float* buf_dev, size_t count, void* stream) {
  if (!comm || !buf_dev) { dd::set_error("dd_comm_allreduce_sum_f32: bad argument"); return DD_ERR_INVALID; }
  if (count == 0 || comm->world == 1) return DD_OK;
  return nccl_check(nccl().all_reduce(buf_dev, buf_dev, count, kNcclFloat32, kNcclSum, comm->comm, static_cast<cudaStream_t>(stream)),
                    "ncclAllReduce");
}

int dd_comm_destroy(dd_comm* comm) {
  if (!comm) return DD_OK;
  int rc = DD_OK;
  if (comm->comm) rc = nccl_check(nccl().comm_destroy(comm->comm), "ncclCommDestroy");
  delete comm;
  return rc;
}

}  // extern "C"

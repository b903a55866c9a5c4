// Data-parallel collectives INSIDE the C-ABI (SURVEY §8(b)4 `b2_ctx` / `b2_allreduce_grads`): a binder in any language gets the
// multi-GPU path of SURVEY §8(e) — all-reduce(sum) of the flat gradient bucket, all-gather of the narrow dense operands of the
// row-sharded aggregate — without going through Python's torch.distributed.  NCCL is resolved at run time with dlopen
// ("libnccl.so.2": the copy already loaded by the host process — e.g. torch's — or the system one), so the library itself has no
// link-time dependency and still loads on machines without NCCL.  The reference has no distributed path.
#include "common.cuh"

#include <dlfcn.h>
#include <string.h>

namespace b2 {
namespace {

typedef struct { char internal[128]; } nccl_uid;        // ncclUniqueId (NCCL_UNIQUE_ID_BYTES = 128)
typedef void* nccl_comm;
// ncclDataType_t: ncclFloat32 = 7, ncclInt32 = 2;  ncclRedOp_t: ncclSum = 0  (nccl.h, stable since 2.0)
struct Api {
  void* handle = nullptr;
  int (*GetUniqueId)(nccl_uid*) = nullptr;
  int (*CommInitRank)(nccl_comm*, int, nccl_uid, int) = nullptr;
  int (*CommDestroy)(nccl_comm) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, nccl_comm, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, nccl_comm, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  int (*GetVersion)(int*) = nullptr;
};

Api* api() {
  static Api a;
  static bool tried = false;
  if (!tried) {
    tried = true;
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      a.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (a.handle) break;
    }
    if (a.handle) {
      a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(a.handle, "ncclGetUniqueId");
      a.CommInitRank = (decltype(a.CommInitRank))dlsym(a.handle, "ncclCommInitRank");
      a.CommDestroy = (decltype(a.CommDestroy))dlsym(a.handle, "ncclCommDestroy");
      a.AllReduce = (decltype(a.AllReduce))dlsym(a.handle, "ncclAllReduce");
      a.AllGather = (decltype(a.AllGather))dlsym(a.handle, "ncclAllGather");
      a.GetErrorString = (decltype(a.GetErrorString))dlsym(a.handle, "ncclGetErrorString");
      a.GetVersion = (decltype(a.GetVersion))dlsym(a.handle, "ncclGetVersion");
      if (!a.GetUniqueId || !a.CommInitRank || !a.CommDestroy || !a.AllReduce || !a.AllGather) a.handle = nullptr;
    }
  }
  return a.handle ? &a : nullptr;
}

int nccl_fail(Api* a, int rc, const char* what) {
  set_error("NCCL error %d (%s) at %s", rc, a->GetErrorString ? a->GetErrorString(rc) : "?", what);
  return B2_ERR_CUDA;
}

}  // namespace
}  // namespace b2

struct b2_comm {
  b2::nccl_comm comm;
  int world, rank;
};

extern "C" int b2_comm_available(void) { return b2::api() != nullptr; }

extern "C" int b2_comm_version(void) {
  b2::Api* a = b2::api();
  int v = 0;
  if (a && a->GetVersion) a->GetVersion(&v);
  return v;
}

extern "C" int b2_comm_unique_id(void* id128) {
  using namespace b2;
  B2_REQUIRE(id128, "b2_comm_unique_id: null pointer");
  Api* a = api();
  B2_REQUIRE(a, "b2_comm_unique_id: libnccl.so.2 could not be loaded");
  nccl_uid id;
  const int rc = a->GetUniqueId(&id);
  if (rc) return nccl_fail(a, rc, "ncclGetUniqueId");
  memcpy(id128, &id, sizeof(id));
  return B2_OK;
}

extern "C" int b2_comm_init_rank(b2_comm** out, const void* id128, int world, int rank) {
  using namespace b2;
  B2_REQUIRE(out && id128 && world >= 1 && rank >= 0 && rank < world, "b2_comm_init_rank: bad arguments");
  Api* a = api();
  B2_REQUIRE(a, "b2_comm_init_rank: libnccl.so.2 could not be loaded");
  nccl_uid id;
  memcpy(&id, id128, sizeof(id));
  nccl_comm c = nullptr;
  const int rc = a->CommInitRank(&c, world, id, rank);     // uses the calling thread's current CUDA device
  if (rc) return nccl_fail(a, rc, "ncclCommInitRank");
  *out = new b2_comm{c, world, rank};
  return B2_OK;
}

extern "C" int b2_comm_destroy(b2_comm* c) {
  using namespace b2;
  if (!c) return B2_OK;
  Api* a = api();
  if (a && c->comm) a->CommDestroy(c->comm);
  delete c;
  return B2_OK;
}

extern "C" int b2_comm_world(const b2_comm* c) { return c ? c->world : 1; }
extern "C" int b2_comm_rank(const b2_comm* c) { return c ? c->rank : 0; }

// in-place sum over ranks of a flat fp32 buffer (the gradient bucket, the partial dz of the pair-sharded decoder, the loss scalar)
extern "C" int b2_allreduce_sum_f32(b2_comm* c, float* buf, int64_t n, void* stream) {
  using namespace b2;
  B2_REQUIRE(c && buf && n >= 0, "b2_allreduce_sum_f32: bad arguments");
  if (n == 0 || c->world == 1) return B2_OK;
  Api* a = api();
  const int rc = a->AllReduce(buf, buf, (size_t)n, /*ncclFloat32*/ 7, /*ncclSum*/ 0, c->comm, as_stream(stream));
  if (rc) return nccl_fail(a, rc, "ncclAllReduce");
  return B2_OK;
}

// full[r·count … (r+1)·count) ← rank r's local[0 … count): equal-sized row blocks (pad the last shard), fp32
extern "C" int b2_allgather_f32(b2_comm* c, const float* local, float* full, int64_t count, void* stream) {
  using namespace b2;
  B2_REQUIRE(c && local && full && count >= 0, "b2_allgather_f32: bad arguments");
  if (count == 0) return B2_OK;
  if (c->world == 1) {
    if (full != local) B2_CHECK_CUDA(cudaMemcpyAsync(full, local, sizeof(float) * (size_t)count, cudaMemcpyDeviceToDevice, as_stream(stream)));
    return B2_OK;
  }
  Api* a = api();
  const int rc = a->AllGather(local, full, (size_t)count, /*ncclFloat32*/ 7, c->comm, as_stream(stream));
  if (rc) return nccl_fail(a, rc, "ncclAllGather");
  return B2_OK;
}

// Data-parallel gradient exchange behind the C ABI: NCCL all-reduce of one gradient bucket on the caller's stream.
//
// Replaces: the torch DistributedDataParallel bucket all-reduce Composer's Trainer sets up around the model
// (reference diffusion/train.py:40, `hydra.utils.instantiate(config.trainer, model=model, ...)`; batch rows are the only
// sharded dimension, so this is the one collective of the hot path - SURVEY.md section 8e).
//
// libnccl is resolved at run time (dlopen of the copy the process already has - torch's bundled one - or the system
// library), so the shared object neither links NCCL nor needs it for single-GPU use or for loading on a CPU box.
#include <dlfcn.h>

#include "common.cuh"
#include "host.h"

namespace {

typedef struct ncclComm* ncclComm_t;
struct NcclUniqueId {
  char internal[128];
};
typedef int ncclResult_t;  // 0 = ncclSuccess
// values of ncclDataType_t / ncclRedOp_t (stable across NCCL 2.x)
enum { NCCL_FLOAT16 = 6, NCCL_FLOAT32 = 7, NCCL_BFLOAT16 = 9 };
enum { NCCL_SUM = 0, NCCL_AVG = 4 };

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(NcclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, NcclUniqueId, int) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
  std::string err;
};

NcclApi* nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return &api;
  tried = true;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (api.handle) break;
  }
  if (!api.handle) {
    api.err = std::string("libnccl not found: ") + (dlerror() ? dlerror() : "");
    return &api;
  }
#define SD2_NCCL_SYM(field, sym)                                           \
  api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.handle, sym)); \
  if (!api.field) {                                                        \
    api.err = std::string("libnccl lacks ") + sym;                         \
    api.handle = nullptr;                                                  \
    return &api;                                                           \
  }
  SD2_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
  SD2_NCCL_SYM(CommInitRank, "ncclCommInitRank")
  SD2_NCCL_SYM(AllReduce, "ncclAllReduce")
  SD2_NCCL_SYM(CommDestroy, "ncclCommDestroy")
  SD2_NCCL_SYM(GetErrorString, "ncclGetErrorString")
  SD2_NCCL_SYM(GetVersion, "ncclGetVersion")
#undef SD2_NCCL_SYM
  return &api;
}

}  // namespace

using namespace sd2;

extern "C" {

int sd2_ddp_unique_id(sd2_ctx* ctx, void* out128) {
  if (!ctx || !out128) return 1;
  NcclApi* a = nccl_api();
  if (!a->handle) return fail(ctx, "sd2_ddp_unique_id: " + a->err);
  NcclUniqueId id;
  ncclResult_t r = a->GetUniqueId(&id);
  if (r != 0) return fail(ctx, std::string("ncclGetUniqueId: ") + a->GetErrorString(r));
  memcpy(out128, &id, sizeof(id));
  return 0;
}

int sd2_ddp_init(sd2_ctx* ctx, const void* unique_id128, int rank, int world) {
  if (!ctx || !unique_id128) return 1;
  if (world < 1 || rank < 0 || rank >= world) return fail(ctx, "sd2_ddp_init: rank / world out of range");
  NcclApi* a = nccl_api();
  if (!a->handle) return fail(ctx, "sd2_ddp_init: " + a->err);
  if (ctx->nccl_comm) return fail(ctx, "sd2_ddp_init: this context already has a communicator (sd2_ddp_destroy first)");
  cudaError_t ce = cudaSetDevice(ctx->device);
  if (ce != cudaSuccess) return fail(ctx, std::string("sd2_ddp_init: ") + cudaGetErrorString(ce));
  NcclUniqueId id;
  memcpy(&id, unique_id128, sizeof(id));
  ncclComm_t comm = nullptr;
  ncclResult_t r = a->CommInitRank(&comm, world, id, rank);
  if (r != 0) return fail(ctx, std::string("ncclCommInitRank: ") + a->GetErrorString(r));
  ctx->nccl_comm = comm;
  ctx->ddp_rank = rank;
  ctx->ddp_world = world;
  return 0;
}

int sd2_ddp_world(sd2_ctx* ctx) { return ctx && ctx->nccl_comm ? ctx->ddp_world : 0; }

int sd2_ddp_allreduce_bucket(sd2_ctx* ctx, void* ptr, long long count, int dtype, int average, sd2_stream stream_) {
  if (!ctx) return 1;
  if (!ctx->nccl_comm) return fail(ctx, "sd2_ddp_allreduce_bucket: no communicator (sd2_ddp_init)");
  if (count < 0 || (count > 0 && !ptr)) return fail(ctx, "sd2_ddp_allreduce_bucket: bad buffer");
  if (count == 0) return 0;
  int dt;
  switch (dtype) {
    case SD2_DT_F32: dt = NCCL_FLOAT32; break;
    case SD2_DT_BF16: dt = NCCL_BFLOAT16; break;
    case SD2_DT_F16: dt = NCCL_FLOAT16; break;
    default: return fail(ctx, "sd2_ddp_allreduce_bucket: dtype must be SD2_DT_F32 / BF16 / F16");
  }
  NcclApi* a = nccl_api();
  ncclResult_t r = a->AllReduce(ptr, ptr, (size_t)count, dt, average ? NCCL_AVG : NCCL_SUM,
                                reinterpret_cast<ncclComm_t>(ctx->nccl_comm), reinterpret_cast<cudaStream_t>(stream_));
  if (r != 0) return fail(ctx, std::string("ncclAllReduce: ") + a->GetErrorString(r));
  return 0;
}

int sd2_ddp_destroy(sd2_ctx* ctx) {
  if (!ctx) return 1;
  if (ctx->nccl_comm) {
    NcclApi* a = nccl_api();
    if (a->handle) a->CommDestroy(reinterpret_cast<ncclComm_t>(ctx->nccl_comm));
    ctx->nccl_comm = nullptr;
    ctx->ddp_world = 0;
  }
  return 0;
}

/* Bytes of split-K scratch sd2_gemm / sd2_conv3x3_* can use for an [M, N] output cut into `splits` K ranges (fp32
 * partial tiles).  The planner never splits further than the workspace it is handed allows, so any size - including
 * none - is valid; this is the size that leaves it unconstrained. */
long long sd2_workspace_bytes(long long M, long long N, int splits) {
  if (M <= 0 || N <= 0 || splits <= 1) return 0;
  return M * N * 4 * (long long)splits;
}

}  // extern "C"

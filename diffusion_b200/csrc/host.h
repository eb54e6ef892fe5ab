// Internal host-side declarations shared by the .cu translation units (not part of the C ABI).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <stdlib.h>
#include <string.h>

#include <string>
#include <utility>

#include "../../include/sd2b200.h"

struct sd2_ctx {
  int device = 0;
  int num_sms = 148;
  long long launches = 0;
  std::string err;
  void* nccl_comm = nullptr;  // ncclComm_t of sd2_ddp_init (ddp.cu), or null
  int ddp_rank = 0, ddp_world = 0;
};

namespace sd2 {
struct GemmKParams;
typedef __nv_bfloat16 bf16;

// cuTensorMapEncodeTiled obtained through cudaGetDriverEntryPoint (no link-time libcuda dependency)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_tiled();
// 4-D bf16 tensor map, SWIZZLE_128B, zero OOB fill.  dims/box innermost first; strides in BYTES for dims 1..3.
bool encode_tmap_bf16_4d(CUtensorMap* out, const void* ptr, const uint64_t dims[4], const uint64_t strides_bytes[3],
                         const uint32_t box[4], std::string* err);

enum { SD2_DT_U64_INTERNAL = 100 };  // 64-bit integer elements (fixed-point accumulators), not part of the public ABI
// generic form: dtype SD2_DT_F32 / SD2_DT_BF16 / SD2_DT_U64_INTERNAL, swizzle_bytes 128 / 64 / 32 / 0 (inner box row must not exceed it)
bool encode_tmap_4d(CUtensorMap* out, int dtype, int swizzle_bytes, const void* ptr, const uint64_t dims[4],
                    const uint64_t strides_bytes[3], const uint32_t box[4], std::string* err);

int gemm_out_chunk(int BN);
// tensor map of a plain (possibly batched) bf16 operand; K-major box = (64, box_rows), MN-major box = (64, 64)
bool plain_tmap(CUtensorMap* tm, const sd2_operand& o, int box_rows, std::string* err);
// output map: dims (N, M, nb0, nb1), box (chunk_cols, 32 rows); inner box row 128 B (SWIZZLE_128B) or 64 B (SWIZZLE_64B)
bool out_tmap(CUtensorMap* tm, void* ptr, bool f32, int chunk_cols, int N, int M, long long ldo, long long nb0,
              long long nb1, long long bs0, long long bs1, std::string* err);
// cluster: run as 2-CTA clusters that share the B tile through TMA multicast (tmB of a K-major B then has BN/2-row boxes)
int gemm_use_cluster(const GemmKParams& p, int BN, bool a_mn, bool b_mn, int num_sms);  // 0 / 2 (multicast B) / 3 (cta_group::2)
cudaError_t launch_gemm_tc(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmO, const GemmKParams& p,
                           int BN, bool a_mn, bool b_mn, int cluster, int num_sms, cudaStream_t stream);
cudaError_t launch_splitk_finalize(const float* ws, int splits, long long M, int N, float alpha, const float* bias,
                                   const float* rowbias, int rows_per_group, long long ld_rowbias, const bf16* residual,
                                   long long ldr, void* out, long long ldo, int out_f32, cudaStream_t stream);

// Kernel launch with the programmatic-stream-serialization attribute (see pdl_grid_sync in common.cuh).
// SD2_NO_PDL=1 in the environment launches plainly (debugging aid).
inline bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SD2_NO_PDL");
    v = (e && e[0] == '1') ? 0 : 1;
  }
  return v == 1;
}
template <typename... P, typename... A>
inline cudaError_t launch_k(void (*kern)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, A&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<P>(std::forward<A>(args))...);
}

inline int fail(sd2_ctx* ctx, const std::string& msg) {
  if (ctx) ctx->err = msg;
  return 1;
}
inline int check_launch(sd2_ctx* ctx, const char* what, int nlaunch = 1) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(ctx, std::string(what) + ": " + cudaGetErrorString(e));
  if (ctx) ctx->launches += nlaunch;
  return 0;
}
// default cap of 24 blocks per SM: measured on the whole step (SD2_WAVES_PCT A/B, same box) 8 -> 24 is worth 0.2-0.5 %
inline int grid_for(long long work_items, int threads, int num_sms, int max_waves = 24) {
  static int scale_pct = -1;  // tuning aid: SD2_WAVES_PCT scales the grid cap of every grid-stride kernel (100 = as written)
  if (scale_pct < 0) {
    const char* e = getenv("SD2_WAVES_PCT");
    scale_pct = e ? atoi(e) : 100;
    if (scale_pct < 1) scale_pct = 100;
  }
  max_waves = (int)(((long long)max_waves * scale_pct + 99) / 100);
  long long b = (work_items + threads - 1) / threads;
  long long cap = (long long)num_sms * max_waves;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}
}  // namespace sd2

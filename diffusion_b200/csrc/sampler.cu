// Rows f2-f4 of the hot-path scope table: the kernels either side of the training step.
//   * cfg_ddim_step : classifier-free-guidance combine + one DDIM step (eta = 0) + re-layout of the new latents into the
//                     UNet's NHWC8 bf16 input, one launch per sampling step.  Replaces, per step, the chunk / guidance
//                     arithmetic / DDIMScheduler.step / torch.cat([latents] * 2) of reference
//                     diffusion/models/stable_diffusion.py:353-371 (and the cast at the UNet's first conv).
//   * ema_update    : ema = ema * s + p * (1 - s), reference diffusion/algorithms/ema.py:62-63 (compute_ema)
//   * cast_to_bf16  : fp16 / fp32 wire data -> the bf16 context buffer (reference dataset format laion.py:103-111)
//   * wire_gather   : host-side gather of n equally sized sample buffers into one (pinned) batch buffer
// All device kernels are HBM / latency bound elementwise passes.
#include "common.cuh"
#include "host.h"

#include <cuda_fp16.h>

namespace sd2 {

__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// One thread per latent pixel.  pred8: [(nb*B)][HW][8] bf16, nb = 2 with guidance (unconditional half first).
// Type promotion of the reference expression is reproduced operation by operation: the UNet output is bf16 (autocast),
// python / 0-dim scalars do not promote it, the latents are fp32:
//   eps   = bf16(u + bf16(gs * bf16(t - u)))                               (guidance, all in bf16)
//   x0    = (x - bf16(sqrt(1 - a_t) * eps)) * (1 / sqrt(a_t))              (fp32; tensor / scalar is a reciprocal multiply)
//   x'    = sqrt(a_prev) * x0 + bf16(sqrt(1 - a_prev) * eps)               (fp32)
__global__ void __launch_bounds__(256) cfg_ddim_step_kernel(const bf16* __restrict__ pred8, float* __restrict__ latents,
                                                            bf16* __restrict__ next8, int B, int HW, int nb, float gs,
                                                            float sqrt_beta_t, float sqrt_alpha_t, float sqrt_alpha_prev,
                                                            float dir_coef) {
  pdl_grid_sync();
  // torch divides a tensor by a (CPU) scalar as a multiplication with the fp32-rounded reciprocal
  const float inv_sqrt_alpha_t = __frcp_rn(sqrt_alpha_t);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)B * HW;
       i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / HW, hw = i % HW;
    const uint2 uu = *reinterpret_cast<const uint2*>(pred8 + i * 8);
    const float2 u01 = unpack_bf16x2(uu.x), u23 = unpack_bf16x2(uu.y);
    float eps[4] = {u01.x, u01.y, u23.x, u23.y};
    if (nb == 2) {
      const uint2 tt = *reinterpret_cast<const uint2*>(pred8 + ((long long)B * HW + i) * 8);
      const float2 t01 = unpack_bf16x2(tt.x), t23 = unpack_bf16x2(tt.y);
      const float tx[4] = {t01.x, t01.y, t23.x, t23.y};
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float d = bf16_round(__fsub_rn(tx[c], eps[c]));
        const float m = bf16_round(__fmul_rn(gs, d));
        eps[c] = bf16_round(__fadd_rn(eps[c], m));
      }
    }
    float xn[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const long long o = (b * 4 + c) * HW + hw;
      const float x = latents[o];
      const float x0 = __fmul_rn(__fsub_rn(x, bf16_round(__fmul_rn(sqrt_beta_t, eps[c]))), inv_sqrt_alpha_t);
      xn[c] = __fadd_rn(__fmul_rn(sqrt_alpha_prev, x0), bf16_round(__fmul_rn(dir_coef, eps[c])));
      latents[o] = xn[c];
    }
    uint4 o8;
    o8.x = pack_bf16x2(xn[0], xn[1]);
    o8.y = pack_bf16x2(xn[2], xn[3]);
    o8.z = 0u;
    o8.w = 0u;
    for (int h = 0; h < nb; ++h) *reinterpret_cast<uint4*>(next8 + ((long long)h * B * HW + i) * 8) = o8;
  }
}

// ema = ema * s + p * (1 - s): two rounded products and a rounded sum, like the reference's tensor expression
__global__ void __launch_bounds__(256) ema_update_kernel(float* __restrict__ ema, const float* __restrict__ p, long long n,
                                                         float s, float one_minus_s) {
  pdl_grid_sync();
  const long long n4 = n / 4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 e = reinterpret_cast<float4*>(ema)[i];
    const float4 q = reinterpret_cast<const float4*>(p)[i];
    e.x = __fadd_rn(__fmul_rn(e.x, s), __fmul_rn(q.x, one_minus_s));
    e.y = __fadd_rn(__fmul_rn(e.y, s), __fmul_rn(q.y, one_minus_s));
    e.z = __fadd_rn(__fmul_rn(e.z, s), __fmul_rn(q.z, one_minus_s));
    e.w = __fadd_rn(__fmul_rn(e.w, s), __fmul_rn(q.w, one_minus_s));
    reinterpret_cast<float4*>(ema)[i] = e;
  }
  const long long t = n4 * 4 + blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t < n) ema[t] = __fadd_rn(__fmul_rn(ema[t], s), __fmul_rn(p[t], one_minus_s));
}

template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <>
__device__ __forceinline__ float to_f32<bf16>(bf16 v) { return __bfloat162float(v); }

template <typename T>
__global__ void __launch_bounds__(256) cast_to_bf16_kernel(const T* __restrict__ src, bf16* __restrict__ dst, long long n) {
  pdl_grid_sync();
  const long long n2 = n / 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n2; i += (long long)gridDim.x * blockDim.x)
    reinterpret_cast<uint32_t*>(dst)[i] = pack_bf16x2(to_f32<T>(src[2 * i]), to_f32<T>(src[2 * i + 1]));
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) dst[n - 1] = __float2bfloat16_rn(to_f32<T>(src[n - 1]));
}

}  // namespace sd2

using namespace sd2;

extern "C" int sd2_cfg_ddim_step(sd2_ctx* ctx, const void* pred_nhwc8, float* latents, void* next_nhwc8, int B, int H, int W,
                                 int guidance, float guidance_scale, float sqrt_beta_t, float sqrt_alpha_t,
                                 float sqrt_alpha_prev, float dir_coef, sd2_stream stream_) {
  if (!ctx) return 1;
  if (B <= 0 || H <= 0 || W <= 0) return fail(ctx, "sd2_cfg_ddim_step: bad shape");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const long long n = (long long)B * H * W;
  launch_k(cfg_ddim_step_kernel, dim3(grid_for(n, 256, ctx->num_sms)), dim3(256), 0, stream,
           reinterpret_cast<const bf16*>(pred_nhwc8), latents, reinterpret_cast<bf16*>(next_nhwc8), B, H * W,
           guidance ? 2 : 1, guidance_scale, sqrt_beta_t, sqrt_alpha_t, sqrt_alpha_prev, dir_coef);
  return check_launch(ctx, "cfg_ddim_step");
}

extern "C" int sd2_ema_update(sd2_ctx* ctx, float* ema, const float* param, long long n, float smoothing,
                              float one_minus_smoothing, sd2_stream stream_) {
  if (!ctx) return 1;
  if (n < 0) return fail(ctx, "sd2_ema_update: n < 0");
  if (n == 0) return 0;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  launch_k(ema_update_kernel, dim3(grid_for(n / 4 + 1, 256, ctx->num_sms, 16)), dim3(256), 0, stream, ema, param, n, smoothing,
           one_minus_smoothing);
  return check_launch(ctx, "ema_update");
}

extern "C" int sd2_cast_to_bf16(sd2_ctx* ctx, const void* src, int src_dtype, void* dst, long long n, sd2_stream stream_) {
  if (!ctx) return 1;
  if (n <= 0) return n == 0 ? 0 : fail(ctx, "sd2_cast_to_bf16: n < 0");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const dim3 grid(grid_for(n / 2 + 1, 256, ctx->num_sms));
  bf16* d = reinterpret_cast<bf16*>(dst);
  if (src_dtype == SD2_DT_F32)
    launch_k(cast_to_bf16_kernel<float>, grid, dim3(256), 0, stream, reinterpret_cast<const float*>(src), d, n);
  else if (src_dtype == SD2_DT_F16)
    launch_k(cast_to_bf16_kernel<__half>, grid, dim3(256), 0, stream, reinterpret_cast<const __half*>(src), d, n);
  else if (src_dtype == SD2_DT_BF16)
    launch_k(cast_to_bf16_kernel<bf16>, grid, dim3(256), 0, stream, reinterpret_cast<const bf16*>(src), d, n);
  else
    return fail(ctx, "sd2_cast_to_bf16: unsupported dtype");
  return check_launch(ctx, "cast_to_bf16");
}

// Host function (no GPU work): dst[i * bytes_each ..] = src[i][0 .. bytes_each).  Null sources are an error.
extern "C" int sd2_wire_gather(const void* const* src, int n, long long bytes_each, void* dst) {
  if (n < 0 || bytes_each < 0 || (n > 0 && (!src || !dst))) return 1;
  uint8_t* d = reinterpret_cast<uint8_t*>(dst);
  for (int i = 0; i < n; ++i) {
    if (!src[i]) return 2;
    memcpy(d + (size_t)i * (size_t)bytes_each, src[i], (size_t)bytes_each);
  }
  return 0;
}

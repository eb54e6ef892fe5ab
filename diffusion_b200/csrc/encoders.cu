// Row f1 of the hot-path scope table: the element-wise / gather kernels of the in-loop VAE encoder and
// CLIP text encoder (reference diffusion/models/stable_diffusion.py:160-174).  The contractions of both encoders run on
// the tcgen05 GEMM / implicit-conv kernel (gemm_tc.cu) and the norm kernels (norm.cu); what is here is the glue those
// networks need and the UNet does not:
//   * nchw_to_nhwc8 / nhwc8_to_nchw : RGB images (3 channels) <-> the 8-channel bf16 NHWC rows the conv kernel reads
//   * vae_sample                   : quant_conv (1x1, 8 -> 8) + DiagonalGaussianDistribution.sample() + the 0.18215 scale
//   * embed_tokens                 : token_embedding[ids] + position_embedding
//   * softmax_causal               : row softmax of fp32 scores under the causal mask of the text transformer
//   * gelu                         : exact (erf) GELU of the CLIP MLP
// All HBM / latency bound.
#include "common.cuh"
#include "host.h"

#include <cuda_fp16.h>

namespace sd2 {

template <typename T>
struct Io;
template <>
struct Io<float> {
  static __device__ __forceinline__ float ld(const float* p, long long i) { return p[i]; }
  static __device__ __forceinline__ void st(float* p, long long i, float v) { p[i] = v; }
  static __device__ __forceinline__ float rt(float v) { return v; }
};
template <>
struct Io<__half> {
  static __device__ __forceinline__ float ld(const __half* p, long long i) { return __half2float(p[i]); }
  static __device__ __forceinline__ void st(__half* p, long long i, float v) { p[i] = __float2half_rn(v); }
  static __device__ __forceinline__ float rt(float v) { return __half2float(__float2half_rn(v)); }
};
template <>
struct Io<bf16> {
  static __device__ __forceinline__ float ld(const bf16* p, long long i) { return __bfloat162float(p[i]); }
  static __device__ __forceinline__ void st(bf16* p, long long i, float v) { p[i] = __float2bfloat16_rn(v); }
  static __device__ __forceinline__ float rt(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }
};

// one thread per pixel; Cc <= 8 real channels, the rest of the 8-wide row is zero
template <typename T>
__global__ void __launch_bounds__(256) nchw_to_nhwc8_kernel(const T* __restrict__ src, bf16* __restrict__ dst, int B, int Cc,
                                                            long long HW) {
  pdl_grid_sync();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)B * HW;
       i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / HW, hw = i % HW;
    float v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int c = 0; c < Cc; ++c) v[c] = Io<T>::ld(src, (b * Cc + c) * HW + hw);
    uint4 u;
    u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]); u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(dst + i * 8) = u;
  }
}
// dst[b][c][hw] = clamp(src8[b][hw][c] * scale + shift, lo, hi) for c < Cc
template <typename T>
__global__ void __launch_bounds__(256) nhwc8_to_nchw_kernel(const bf16* __restrict__ src, T* __restrict__ dst, int B, int Cc,
                                                            long long HW, float scale, float shift, float lo, float hi) {
  pdl_grid_sync();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)B * HW;
       i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / HW, hw = i % HW;
    const uint4 u = *reinterpret_cast<const uint4*>(src + i * 8);
    const float2 p0 = unpack_bf16x2(u.x), p1 = unpack_bf16x2(u.y), p2 = unpack_bf16x2(u.z), p3 = unpack_bf16x2(u.w);
    const float v[8] = {p0.x, p0.y, p1.x, p1.y, p2.x, p2.y, p3.x, p3.y};
    for (int c = 0; c < Cc; ++c) Io<T>::st(dst, (b * Cc + c) * HW + hw, fminf(fmaxf(fmaf(v[c], scale, shift), lo), hi));
  }
}

// moments8: bf16 [B*HW][8] = encoder.conv_out output (mean 0..3 | logvar 4..7 before quant_conv).
// m = Wq[8][8] * h + bq (quant_conv, 1x1);  mean = m[0..3];  logvar = clamp(m[4..7], -30, 20);  std = exp(0.5 logvar);
// z = (mean + std * noise) * scale, every step rounded in T like the reference's half-precision tensor ops.
template <typename T>
__global__ void __launch_bounds__(256) vae_sample_kernel(const bf16* __restrict__ moments8, const float* __restrict__ wq,
                                                         const float* __restrict__ bq, const T* __restrict__ noise,
                                                         T* __restrict__ latents, T* __restrict__ mean_out, int B, long long HW,
                                                         float scale) {
  pdl_grid_sync();
  __shared__ float w[64], bb[8];
  if (threadIdx.x < 64) w[threadIdx.x] = wq[threadIdx.x];
  if (threadIdx.x < 8) bb[threadIdx.x] = bq[threadIdx.x];
  __syncthreads();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)B * HW;
       i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / HW, hw = i % HW;
    const uint4 u = *reinterpret_cast<const uint4*>(moments8 + i * 8);
    const float2 p0 = unpack_bf16x2(u.x), p1 = unpack_bf16x2(u.y), p2 = unpack_bf16x2(u.z), p3 = unpack_bf16x2(u.w);
    const float h[8] = {p0.x, p0.y, p1.x, p1.y, p2.x, p2.y, p3.x, p3.y};
    float m[8];
#pragma unroll
    for (int o = 0; o < 8; ++o) {
      float a = bb[o];
#pragma unroll
      for (int k = 0; k < 8; ++k) a = fmaf(w[o * 8 + k], h[k], a);
      m[o] = Io<T>::rt(a);
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const long long o = (b * 4 + c) * HW + hw;
      const float logvar = fminf(fmaxf(m[4 + c], -30.f), 20.f);
      const float sd = Io<T>::rt(expf(Io<T>::rt(0.5f * logvar)));
      const float z = Io<T>::rt(m[c] + Io<T>::rt(sd * Io<T>::ld(noise, o)));
      Io<T>::st(latents, o, z * scale);
      if (mean_out) Io<T>::st(mean_out, o, m[c]);
    }
  }
}

// x[r][:] = bf16(tok[ids[r]][:] + pos[r % L][:]), one warp per row, D % 8 == 0; ids outside [0, vocab) trap to row 0
__global__ void __launch_bounds__(256) embed_tokens_kernel(const long long* __restrict__ ids, const float* __restrict__ tok,
                                                           const float* __restrict__ pos, bf16* __restrict__ x, long long rows,
                                                           int L, int D, int vocab) {
  pdl_grid_sync();
  const int lane = threadIdx.x & 31;
  const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long r = warp; r < rows; r += nwarps) {
    long long id = ids[r];
    if (id < 0 || id >= vocab) id = 0;
    const float* t = tok + id * D;
    const float* p = pos + (r % L) * D;
    for (int c = lane * 4; c < D; c += 128) {
      const float4 a = *reinterpret_cast<const float4*>(t + c), b = *reinterpret_cast<const float4*>(p + c);
      uint2 o;
      o.x = pack_bf16x2(a.x + b.x, a.y + b.y);
      o.y = pack_bf16x2(a.z + b.z, a.w + b.w);
      *reinterpret_cast<uint2*>(x + r * D + c) = o;
    }
  }
}

// P[row][0..ldp) = softmax over the first min(cols, (row % period) + 1) entries of S[row][:], zero elsewhere; warp per row
__global__ void __launch_bounds__(256) softmax_causal_kernel(const float* __restrict__ S, long long lds, bf16* __restrict__ P,
                                                             long long ldp, long long rows, int cols, int period) {
  pdl_grid_sync();
  const int lane = threadIdx.x & 31;
  const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long row = warp; row < rows; row += nwarps) {
    const float* s = S + row * lds;
    int n = (int)(row % period) + 1;
    if (n > cols) n = cols;
    float mx = -INFINITY;
    for (int c = lane; c < n; c += 32) mx = fmaxf(mx, s[c]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int c = lane; c < n; c += 32) sum += __expf(s[c] - mx);
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    bf16* o = P + row * ldp;
    for (int c = lane; c < (int)ldp; c += 32) o[c] = __float2bfloat16_rn(c < n ? __expf(s[c] - mx) * inv : 0.f);
  }
}

__global__ void __launch_bounds__(256) gelu_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, long long n8) {
  pdl_grid_sync();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    const uint4 u = ldg_stream16(x + i * 8);
    const float2 p0 = unpack_bf16x2(u.x), p1 = unpack_bf16x2(u.y), p2 = unpack_bf16x2(u.z), p3 = unpack_bf16x2(u.w);
    float v[8] = {p0.x, p0.y, p1.x, p1.y, p2.x, p2.y, p3.x, p3.y};
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = 0.5f * v[e] * (1.f + erff(v[e] * 0.70710678118654752f));
    uint4 o;
    o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]); o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(y + i * 8) = o;
  }
}

// out8[r][o] = bf16(b[o] + sum_k w[o][k] * in8[r][k]) for o < n_out (k < n_in), zero for o >= n_out: a 1x1 convolution over
// at most 8 channels (post_quant_conv of the VAE decoder), one thread per pixel
__global__ void __launch_bounds__(256) pixel_linear8_kernel(const bf16* __restrict__ in8, const float* __restrict__ w,
                                                            const float* __restrict__ b, bf16* __restrict__ out8, long long n,
                                                            int n_in, int n_out) {
  pdl_grid_sync();
  __shared__ float ws[64], bs[8];
  if (threadIdx.x < 64) ws[threadIdx.x] = (threadIdx.x / 8 < n_out && threadIdx.x % 8 < n_in) ? w[(threadIdx.x / 8) * n_in + threadIdx.x % 8] : 0.f;
  if (threadIdx.x < 8) bs[threadIdx.x] = threadIdx.x < n_out ? b[threadIdx.x] : 0.f;
  __syncthreads();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const uint4 u = *reinterpret_cast<const uint4*>(in8 + i * 8);
    const float2 p0 = unpack_bf16x2(u.x), p1 = unpack_bf16x2(u.y), p2 = unpack_bf16x2(u.z), p3 = unpack_bf16x2(u.w);
    const float h[8] = {p0.x, p0.y, p1.x, p1.y, p2.x, p2.y, p3.x, p3.y};
    float m[8];
#pragma unroll
    for (int o = 0; o < 8; ++o) {
      float a = bs[o];
#pragma unroll
      for (int k = 0; k < 8; ++k) a = fmaf(ws[o * 8 + k], h[k], a);
      m[o] = a;
    }
    uint4 r;
    r.x = pack_bf16x2(m[0], m[1]); r.y = pack_bf16x2(m[2], m[3]); r.z = pack_bf16x2(m[4], m[5]); r.w = pack_bf16x2(m[6], m[7]);
    *reinterpret_cast<uint4*>(out8 + i * 8) = r;
  }
}

}  // namespace sd2

using namespace sd2;

#define SD2_DT_SWITCH(dt, CALL, what)                                  \
  if ((dt) == SD2_DT_F32) { CALL(float); }                             \
  else if ((dt) == SD2_DT_BF16) { CALL(bf16); }                        \
  else if ((dt) == SD2_DT_F16) { CALL(__half); }                       \
  else return fail(ctx, what ": unsupported dtype");

extern "C" int sd2_nchw_to_nhwc8(sd2_ctx* ctx, const void* src, int src_dtype, void* dst_nhwc8, int B, int C, int H, int W,
                                 sd2_stream stream_) {
  if (!ctx) return 1;
  if (C < 1 || C > 8 || B < 1 || H < 1 || W < 1) return fail(ctx, "sd2_nchw_to_nhwc8: need 1 <= C <= 8");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const long long HW = (long long)H * W;
  const dim3 grid(grid_for((long long)B * HW, 256, ctx->num_sms));
#define CALL(T) launch_k(nchw_to_nhwc8_kernel<T>, grid, dim3(256), 0, stream, reinterpret_cast<const T*>(src), reinterpret_cast<bf16*>(dst_nhwc8), B, C, HW)
  SD2_DT_SWITCH(src_dtype, CALL, "sd2_nchw_to_nhwc8")
#undef CALL
  return check_launch(ctx, "nchw_to_nhwc8");
}

extern "C" int sd2_nhwc8_to_nchw(sd2_ctx* ctx, const void* src_nhwc8, void* dst, int dst_dtype, int B, int C, int H, int W,
                                 float scale, float shift, float lo, float hi, sd2_stream stream_) {
  if (!ctx) return 1;
  if (C < 1 || C > 8 || B < 1 || H < 1 || W < 1) return fail(ctx, "sd2_nhwc8_to_nchw: need 1 <= C <= 8");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const long long HW = (long long)H * W;
  const dim3 grid(grid_for((long long)B * HW, 256, ctx->num_sms));
#define CALL(T) launch_k(nhwc8_to_nchw_kernel<T>, grid, dim3(256), 0, stream, reinterpret_cast<const bf16*>(src_nhwc8), reinterpret_cast<T*>(dst), B, C, HW, scale, shift, lo, hi)
  SD2_DT_SWITCH(dst_dtype, CALL, "sd2_nhwc8_to_nchw")
#undef CALL
  return check_launch(ctx, "nhwc8_to_nchw");
}

extern "C" int sd2_vae_sample(sd2_ctx* ctx, const void* moments_nhwc8, const float* quant_w, const float* quant_b,
                              const void* noise, void* latents, void* mean_out, int dtype, int B, int H, int W, float scale,
                              sd2_stream stream_) {
  if (!ctx) return 1;
  if (B < 1 || H < 1 || W < 1) return fail(ctx, "sd2_vae_sample: bad shape");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const long long HW = (long long)H * W;
  const dim3 grid(grid_for((long long)B * HW, 256, ctx->num_sms));
#define CALL(T) launch_k(vae_sample_kernel<T>, grid, dim3(256), 0, stream, reinterpret_cast<const bf16*>(moments_nhwc8), quant_w, quant_b, reinterpret_cast<const T*>(noise), reinterpret_cast<T*>(latents), reinterpret_cast<T*>(mean_out), B, HW, scale)
  SD2_DT_SWITCH(dtype, CALL, "sd2_vae_sample")
#undef CALL
  return check_launch(ctx, "vae_sample");
}

extern "C" int sd2_embed_tokens(sd2_ctx* ctx, const int64_t* ids, const float* token_embedding, const float* position_embedding,
                                void* x, long long rows, int L, int D, int vocab, sd2_stream stream_) {
  if (!ctx) return 1;
  if (rows < 1 || L < 1 || D % 8 != 0 || vocab < 1) return fail(ctx, "sd2_embed_tokens: bad shape (D % 8)");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  launch_k(embed_tokens_kernel, dim3(grid_for(rows * 32, 256, ctx->num_sms)), dim3(256), 0, stream,
           reinterpret_cast<const long long*>(ids), token_embedding, position_embedding, reinterpret_cast<bf16*>(x), rows, L, D, vocab);
  return check_launch(ctx, "embed_tokens");
}

extern "C" int sd2_softmax_causal_fwd(sd2_ctx* ctx, const float* S, long long lds, void* P, long long ldp, long long rows,
                                      int cols, int period, sd2_stream stream_) {
  if (!ctx) return 1;
  if (rows < 1 || cols < 1 || period < 1 || ldp < cols) return fail(ctx, "sd2_softmax_causal_fwd: bad shape");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  launch_k(softmax_causal_kernel, dim3(grid_for(rows * 32, 256, ctx->num_sms, 16)), dim3(256), 0, stream, S, lds,
           reinterpret_cast<bf16*>(P), ldp, rows, cols, period);
  return check_launch(ctx, "softmax_causal_fwd");
}

extern "C" int sd2_gelu_fwd(sd2_ctx* ctx, const void* x, void* y, long long n, sd2_stream stream_) {
  if (!ctx) return 1;
  if (n < 0 || n % 8 != 0) return fail(ctx, "sd2_gelu_fwd: n % 8");
  if (n == 0) return 0;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  launch_k(gelu_kernel, dim3(grid_for(n / 8, 256, ctx->num_sms)), dim3(256), 0, stream, reinterpret_cast<const bf16*>(x),
           reinterpret_cast<bf16*>(y), n / 8);
  return check_launch(ctx, "gelu_fwd");
}

extern "C" int sd2_pixel_linear8(sd2_ctx* ctx, const void* in_nhwc8, const float* w, const float* b, void* out_nhwc8, long long n,
                                 int n_in, int n_out, sd2_stream stream_) {
  if (!ctx) return 1;
  if (n < 1 || n_in < 1 || n_in > 8 || n_out < 1 || n_out > 8) return fail(ctx, "sd2_pixel_linear8: need 1 <= n_in, n_out <= 8");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  launch_k(pixel_linear8_kernel, dim3(grid_for(n, 256, ctx->num_sms)), dim3(256), 0, stream, reinterpret_cast<const bf16*>(in_nhwc8), w, b,
           reinterpret_cast<bf16*>(out_nhwc8), n, n_in, n_out);
  return check_launch(ctx, "pixel_linear8");
}

// K1: fused timestep sampling + Gaussian noise + DDPM add_noise + sinusoidal timestep embedding.
//
// Replaces, in one launch, the reference's
//   timesteps = torch.randint(0, len(noise_scheduler), (B,), device)      stable_diffusion.py:177
//   noise     = torch.randn_like(latents)                                 stable_diffusion.py:179
//   noised    = noise_scheduler.add_noise(latents, noise, timesteps)      stable_diffusion.py:180  (diffusers DDPMScheduler)
//   t_emb     = get_timestep_embedding(timesteps, 320, flip_sin_to_cos)   first op of unet(...) at :183 (diffusers Timesteps)
// Bit-exact with torch's CUDA generator: the same curand Philox4_32_10 device functions, the same per-thread
// subsequence (= global thread id), the same grid policy and offset increments as ATen's
// distribution_elementwise_grid_stride_kernel (ATen/native/cuda/DistributionTemplates.h:50-87), randint first,
// randn second.  HBM-bound (a few hundred KB) - a single wave of 256-thread blocks.
#include <curand_kernel.h>

#include "common.cuh"
#include "host.h"

namespace sd2 {

template <typename T>
struct Cvt;
template <>
struct Cvt<float> {
  static __device__ __forceinline__ float rt(float x) { return x; }  // round-trip through T
  static __device__ __forceinline__ float ld(const float* p, long long i) { return p[i]; }
  static __device__ __forceinline__ void st(float* p, long long i, float v) { p[i] = v; }
};
template <>
struct Cvt<__nv_bfloat16> {
  static __device__ __forceinline__ float rt(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }
  static __device__ __forceinline__ float ld(const __nv_bfloat16* p, long long i) { return __bfloat162float(p[i]); }
  static __device__ __forceinline__ void st(__nv_bfloat16* p, long long i, float v) { p[i] = __float2bfloat16_rn(v); }
};
template <>
struct Cvt<__half> {
  static __device__ __forceinline__ float rt(float x) { return __half2float(__float2half_rn(x)); }
  static __device__ __forceinline__ float ld(const __half* p, long long i) { return __half2float(p[i]); }
  static __device__ __forceinline__ void st(__half* p, long long i, float v) { p[i] = __float2half_rn(v); }
};

// torch.randint(0, T, (B,)) element `b`: drawn by thread (b % nth1) of a grid of nth1 threads, component (b / nth1) % 4
// of that thread's (b / (4 nth1))-th curand4()
__device__ __forceinline__ int sample_timestep(uint64_t seed, uint64_t offset, int b, int nth1, int T) {
  curandStatePhilox4_32_10_t st;
  curand_init(seed, (unsigned long long)(b % nth1), offset, &st);
  const int round = b / (4 * nth1), comp = (b / nth1) % 4;
  uint4 r = curand4(&st);
  for (int i = 0; i < round; ++i) r = curand4(&st);
  const unsigned int v = comp == 0 ? r.x : comp == 1 ? r.y : comp == 2 ? r.z : r.w;
  return (int)(v % (unsigned int)T);
}

template <typename T>
__global__ void __launch_bounds__(256) k1_noise_sched_kernel(uint64_t seed, uint64_t off_randint, uint64_t off_randn,
                                                             int nth1, const T* __restrict__ latents, int B, int H, int W,
                                                             const float* __restrict__ alphas_cumprod, int num_t,
                                                             int64_t* __restrict__ out_t, T* __restrict__ out_noise,
                                                             T* __restrict__ out_noised_nchw,
                                                             __nv_bfloat16* __restrict__ out_nhwc8,
                                                             __nv_bfloat16* __restrict__ out_temb, int temb_dim) {
  const long long numel = (long long)B * 4 * H * W;
  const long long nthreads = (long long)gridDim.x * blockDim.x;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int HW = H * W;

  // ---- noise + add_noise: ATen grid-stride policy, unroll 4
  curandStatePhilox4_32_10_t st;
  curand_init(seed, (unsigned long long)idx, off_randn, &st);
  const long long rounded = ((numel - 1) / (nthreads * 4) + 1) * nthreads * 4;
  for (long long li0 = idx; li0 < rounded; li0 += nthreads * 4) {
    const float4 rnd = curand_normal4(&st);
    const float rv[4] = {rnd.x, rnd.y, rnd.z, rnd.w};
#pragma unroll
    for (int ii = 0; ii < 4; ++ii) {
      const long long li = li0 + nthreads * ii;
      if (li < numel) {
        // transformation::normal(val, mean=0, std=1) then static_cast<scalar_t>
        const float nz = Cvt<T>::rt(__fadd_rn(__fmul_rn(rv[ii], 1.0f), 0.0f));
        Cvt<T>::st(out_noise, li, nz);
        const int b = (int)(li / (4 * HW));
        const int c = (int)((li / HW) % 4);
        const int hw = (int)(li % HW);
        const int t = sample_timestep(seed, off_randint, b, nth1, num_t);
        if (c == 0 && hw == 0) out_t[b] = (int64_t)t;
        // DDPMScheduler.add_noise: alphas_cumprod cast to the sample dtype BEFORE the sqrt, every op rounded in T
        const float ac = Cvt<T>::rt(alphas_cumprod[t]);
        const float a = Cvt<T>::rt(sqrtf(ac));
        const float s = Cvt<T>::rt(sqrtf(Cvt<T>::rt(__fsub_rn(1.0f, ac))));
        const float x0 = Cvt<T>::ld(latents, li);
        const float nd = Cvt<T>::rt(__fadd_rn(Cvt<T>::rt(__fmul_rn(a, x0)), Cvt<T>::rt(__fmul_rn(s, nz))));
        if (out_noised_nchw != nullptr) Cvt<T>::st(out_noised_nchw, li, nd);
        __nv_bfloat16* px = out_nhwc8 + ((long long)b * HW + hw) * 8;
        px[c] = __float2bfloat16_rn(nd);
        px[4 + c] = __float2bfloat16_rn(0.f);
      }
    }
  }

  // ---- sinusoidal embedding [cos | sin], fp32 math as in diffusers get_timestep_embedding
  const int half = temb_dim / 2;
  for (long long e = idx; e < (long long)B * half; e += nthreads) {
    const int b = (int)(e / half), i = (int)(e % half);
    const int t = sample_timestep(seed, off_randint, b, nth1, num_t);
    const float exponent = __fdiv_rn(__fmul_rn(-9.210340371976184f, (float)i), (float)half);
    const float arg = __fmul_rn((float)t, expf(exponent));
    out_temb[(long long)b * temb_dim + i] = __float2bfloat16_rn(Cvt<T>::rt(cosf(arg)));
    out_temb[(long long)b * temb_dim + half + i] = __float2bfloat16_rn(Cvt<T>::rt(sinf(arg)));
  }
}

template <typename T>
__global__ void temb_kernel(const int64_t* __restrict__ ts, int B, __nv_bfloat16* __restrict__ out, int temb_dim) {
  const int half = temb_dim / 2;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < (long long)B * half;
       e += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(e / half), i = (int)(e % half);
    const float exponent = __fdiv_rn(__fmul_rn(-9.210340371976184f, (float)i), (float)half);
    const float arg = __fmul_rn((float)ts[b], expf(exponent));
    out[(long long)b * temb_dim + i] = __float2bfloat16_rn(Cvt<T>::rt(cosf(arg)));
    out[(long long)b * temb_dim + half + i] = __float2bfloat16_rn(Cvt<T>::rt(sinf(arg)));
  }
}

// one thread per pixel
template <typename T>
__global__ void nchw4_to_nhwc8_kernel(const T* __restrict__ src, __nv_bfloat16* __restrict__ dst, int B, int HW) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)B * HW;
       i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / HW, hw = i % HW;
    float v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int c = 0; c < 4; ++c) v[c] = Cvt<T>::ld(src, (b * 4 + c) * HW + hw);
    uint4 u;
    u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]); u.z = 0u; u.w = 0u;
    *reinterpret_cast<uint4*>(dst + i * 8) = u;
  }
}
template <typename T>
__global__ void nhwc8_to_nchw4_kernel(const __nv_bfloat16* __restrict__ src, T* __restrict__ dst, int B, int HW) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)B * HW;
       i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / HW, hw = i % HW;
    const uint4 u = *reinterpret_cast<const uint4*>(src + i * 8);
    const float2 a = unpack_bf16x2(u.x), c = unpack_bf16x2(u.y);
    Cvt<T>::st(dst, (b * 4 + 0) * HW + hw, a.x);
    Cvt<T>::st(dst, (b * 4 + 1) * HW + hw, a.y);
    Cvt<T>::st(dst, (b * 4 + 2) * HW + hw, c.x);
    Cvt<T>::st(dst, (b * 4 + 3) * HW + hw, c.y);
  }
}
__global__ void scale_by_scalar_kernel(__nv_bfloat16* __restrict__ x, long long n, const float* __restrict__ scalar) {
  const float s = *scalar;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    x[i] = __float2bfloat16_rn(__bfloat162float(x[i]) * s);
}
__global__ void fill_f32_kernel2(float* __restrict__ x, long long n, float v) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) x[i] = v;
}

}  // namespace sd2

using namespace sd2;

#define SD2_DISPATCH_DT(dt, CALL)                                  \
  if ((dt) == SD2_DT_F32) { CALL(float); }                         \
  else if ((dt) == SD2_DT_BF16) { CALL(__nv_bfloat16); }           \
  else if ((dt) == SD2_DT_F16) { CALL(__half); }                   \
  else return fail(ctx, "unsupported dtype");

extern "C" int sd2_timestep_embedding(sd2_ctx* ctx, const int64_t* timesteps, int B, void* out_temb, int temb_dim,
                                      int round_dtype, sd2_stream stream_) {
  if (!ctx) return 1;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const int blocks = grid_for((long long)B * temb_dim / 2, 256, ctx->num_sms);
#define CALL(T) temb_kernel<T><<<blocks, 256, 0, stream>>>(timesteps, B, reinterpret_cast<__nv_bfloat16*>(out_temb), temb_dim)
  SD2_DISPATCH_DT(round_dtype, CALL)
#undef CALL
  return check_launch(ctx, "timestep_embedding");
}
extern "C" int sd2_nchw4_to_nhwc8(sd2_ctx* ctx, const void* src, int src_dtype, void* dst, int B, int H, int W,
                                  sd2_stream stream_) {
  if (!ctx) return 1;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const int blocks = grid_for((long long)B * H * W, 256, ctx->num_sms);
#define CALL(T) nchw4_to_nhwc8_kernel<T><<<blocks, 256, 0, stream>>>(reinterpret_cast<const T*>(src), reinterpret_cast<__nv_bfloat16*>(dst), B, H * W)
  SD2_DISPATCH_DT(src_dtype, CALL)
#undef CALL
  return check_launch(ctx, "nchw4_to_nhwc8");
}
extern "C" int sd2_nhwc8_to_nchw4(sd2_ctx* ctx, const void* src, void* dst, int dst_dtype, int B, int H, int W,
                                  sd2_stream stream_) {
  if (!ctx) return 1;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const int blocks = grid_for((long long)B * H * W, 256, ctx->num_sms);
#define CALL(T) nhwc8_to_nchw4_kernel<T><<<blocks, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(src), reinterpret_cast<T*>(dst), B, H * W)
  SD2_DISPATCH_DT(dst_dtype, CALL)
#undef CALL
  return check_launch(ctx, "nhwc8_to_nchw4");
}
extern "C" int sd2_scale_by_scalar(sd2_ctx* ctx, void* x, long long n, const float* scalar, sd2_stream stream_) {
  if (!ctx) return 1;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  scale_by_scalar_kernel<<<grid_for(n, 256, ctx->num_sms), 256, 0, stream>>>(reinterpret_cast<__nv_bfloat16*>(x), n, scalar);
  return check_launch(ctx, "scale_by_scalar");
}
extern "C" int sd2_fill_f32(sd2_ctx* ctx, float* x, long long n, float value, sd2_stream stream_) {
  if (!ctx) return 1;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  fill_f32_kernel2<<<grid_for(n, 256, ctx->num_sms), 256, 0, stream>>>(x, n, value);
  return check_launch(ctx, "fill_f32");
}

extern "C" int sd2_noise_sched_fwd(sd2_ctx* ctx, uint64_t seed, uint64_t philox_offset, const void* latents,
                                   int lat_dtype, int B, int H, int W, const float* alphas_cumprod,
                                   int num_train_timesteps, int64_t* out_timesteps, void* out_noise,
                                   void* out_noised_nchw, void* out_noised_nhwc8, void* out_temb, int temb_dim,
                                   uint64_t* offset_used, sd2_stream stream_) {
  if (!ctx) return 1;
  if (B <= 0 || H <= 0 || W <= 0 || temb_dim % 2 != 0) return fail(ctx, "sd2_noise_sched_fwd: bad shape");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const unsigned int block = 256, blocks_per_sm = 2048 / block;
  const unsigned long long max_grid = (unsigned long long)ctx->num_sms * blocks_per_sm;
  // randint(B): ATen calc_execution_policy(numel=B, unroll=4)
  unsigned long long grid1 = ((unsigned long long)B + block - 1) / block;
  if (grid1 > max_grid) grid1 = max_grid;
  const uint64_t inc1 = (((uint64_t)B - 1) / (block * grid1 * 4) + 1) * 4;
  // randn_like(latents)
  const unsigned long long numel = (unsigned long long)B * 4 * H * W;
  unsigned long long grid2 = (numel + block - 1) / block;
  if (grid2 > max_grid) grid2 = max_grid;
  const uint64_t inc2 = ((numel - 1) / (block * grid2 * 4) + 1) * 4;
  const uint64_t off_randn = philox_offset + inc1;
  const int nth1 = (int)(grid1 * block);
  if (offset_used) *offset_used = inc1 + inc2;
#define K1_LAUNCH(T)                                                                                                 \
  k1_noise_sched_kernel<T><<<(unsigned int)grid2, block, 0, stream>>>(                                               \
      seed, philox_offset, off_randn, nth1, reinterpret_cast<const T*>(latents), B, H, W, alphas_cumprod,            \
      num_train_timesteps, out_timesteps, reinterpret_cast<T*>(out_noise), reinterpret_cast<T*>(out_noised_nchw),    \
      reinterpret_cast<__nv_bfloat16*>(out_noised_nhwc8), reinterpret_cast<__nv_bfloat16*>(out_temb), temb_dim)
  if (lat_dtype == SD2_DT_F32) {
    K1_LAUNCH(float);
  } else if (lat_dtype == SD2_DT_BF16) {
    K1_LAUNCH(__nv_bfloat16);
  } else if (lat_dtype == SD2_DT_F16) {
    K1_LAUNCH(__half);
  } else {
    return fail(ctx, "sd2_noise_sched_fwd: unsupported latent dtype");
  }
#undef K1_LAUNCH
  return check_launch(ctx, "k1_noise_sched");
}

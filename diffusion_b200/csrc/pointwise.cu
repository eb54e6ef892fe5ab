// HBM-bound pointwise / layout / reduction kernels of the UNet training step: softmax (fwd/bwd) for the
// materialised-score attention path, GEGLU, SiLU, axpby, strided copies (skip-concat), nearest-2x upsample,
// stride-2 phase split, column sums (bias gradients), dtype casts and the fused MSE loss head.
// All use 16-byte vector accesses coalesced along the channel dimension and grid-stride loops sized in
// multiples of the SM count.
//
// Replaces the corresponding ATen elementwise kernels behind diffusers' Attention/GEGLU/Upsample2D/torch.cat and
// F.mse_loss + torchmetrics MeanSquaredError (reference stable_diffusion.py:76,101,185-187,241-242).
#include "common.cuh"
#include "host.h"
#include <cstdlib>

namespace sd2 {

__device__ __forceinline__ void ld8(const bf16* p, float* v) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
}
__device__ __forceinline__ void st8(bf16* p, const float* v) {
  uint4 u;
  u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]); u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(p) = u;
}

// ---------------------------------------------------------------------------------------------- softmax
// one warp per row; S fp32 (already scaled), P bf16.  Columns [cols, ldp) of P are zero-filled so P can be a
// zero-padded GEMM operand.
__global__ void __launch_bounds__(256) softmax_fwd_kernel(const float* __restrict__ S, long long lds, bf16* __restrict__ P,
                                                          long long ldp, long long rows, int cols) {
  pdl_grid_sync();
  const int lane = threadIdx.x & 31;
  const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long row = warp; row < rows; row += nwarps) {
    const float* s = S + row * lds;
    float m = -INFINITY;
    for (int c = lane; c < cols; c += 32) m = fmaxf(m, s[c]);
    m = warp_max(m);
    float sum = 0.f;
    for (int c = lane; c < cols; c += 32) sum += __expf(s[c] - m);
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    bf16* p = P + row * ldp;
    for (int c = lane; c < (int)ldp; c += 32) p[c] = __float2bfloat16_rn(c < cols ? __expf(s[c] - m) * inv : 0.f);
  }
}

// dS = P * (dP - sum_j dP_j P_j) * scale
__global__ void __launch_bounds__(256) softmax_bwd_kernel(const bf16* __restrict__ P, long long ldp, const float* __restrict__ dP,
                                                          long long lddp, bf16* __restrict__ dS, long long ldds, long long rows,
                                                          int cols, float scale) {
  pdl_grid_sync();
  const int lane = threadIdx.x & 31;
  const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long row = warp; row < rows; row += nwarps) {
    const bf16* p = P + row * ldp;
    const float* dp = dP + row * lddp;
    float dot = 0.f;
    for (int c = lane; c < cols; c += 32) dot += __bfloat162float(p[c]) * dp[c];
    dot = warp_sum(dot);
    bf16* o = dS + row * ldds;
    for (int c = lane; c < (int)ldds; c += 32)
      o[c] = __float2bfloat16_rn(c < cols ? __bfloat162float(p[c]) * (dp[c] - dot) * scale : 0.f);
  }
}

// ---------------------------------------------------------------------------------------------- GEGLU / SiLU
// Exact-GELU pieces from one MUFU.RCP and one MUFU.EX2: erf by Abramowitz-Stegun 7.1.26 (|error| < 1.5e-7, two orders
// below the bf16 rounding of the results; libm's erff made both GEGLU kernels instruction-bound at ~55 % of HBM speed).
//   cdf = Phi(x) = 0.5 (1 + erf(x / sqrt 2)),  pdf = phi(x) = exp(-x^2 / 2) / sqrt(2 pi);  gelu = x cdf, gelu' = cdf + x pdf
// The negative branch returns the small tail directly (no 1 - (1 - tail) cancellation).
__device__ __forceinline__ void gelu_cdf_pdf(float x, float& cdf, float& pdf) {
  const float z = fabsf(x) * 0.70710678118654752f;
  const float t = __fdividef(1.f, fmaf(0.3275911f, z, 1.f));
  const float u = __expf(-0.5f * x * x);
  float poly = fmaf(t, 1.061405429f, -1.453152027f);
  poly = fmaf(t, poly, 1.421413741f);
  poly = fmaf(t, poly, -0.284496736f);
  poly = fmaf(t, poly, 0.254829592f);
  const float tail = 0.5f * poly * t * u;
  cdf = x >= 0.f ? 1.f - tail : tail;
  pdf = 0.3989422804014327f * u;
}
__device__ __forceinline__ void unpack8f(const uint4& u, float* v) {
  const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
}

__global__ void __launch_bounds__(256) geglu_fwd_kernel(const bf16* __restrict__ h, bf16* __restrict__ y, long long rows, int C) {
  pdl_grid_sync();
  const int V = C / 8;
  const long long n = rows * V;
  const long long stride = (long long)gridDim.x * blockDim.x;
  // two vectors per iteration: four 16-byte loads in flight per thread
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += 2 * stride) {
    const long long i2 = i + stride;
    const bool two = i2 < n;
    const long long r = i / V, r2 = two ? i2 / V : r;
    const int v = (int)(i % V), v2 = two ? (int)(i2 % V) : v;
    const uint4 ua = ldg_stream16(h + r * 2 * C + v * 8), ug = ldg_stream16(h + r * 2 * C + C + v * 8);
    const uint4 ub = ldg_stream16(h + r2 * 2 * C + v2 * 8), uh = ldg_stream16(h + r2 * 2 * C + C + v2 * 8);
    float a[8], g[8];
    unpack8f(ua, a);
    unpack8f(ug, g);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float cdf, pdf;
      gelu_cdf_pdf(g[e], cdf, pdf);
      a[e] *= g[e] * cdf;
    }
    st8(y + r * C + v * 8, a);
    if (two) {
      unpack8f(ub, a);
      unpack8f(uh, g);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        float cdf, pdf;
        gelu_cdf_pdf(g[e], cdf, pdf);
        a[e] *= g[e] * cdf;
      }
      st8(y + r2 * C + v2 * 8, a);
    }
  }
}

__global__ void __launch_bounds__(256) geglu_bwd_kernel(const bf16* __restrict__ h, const bf16* __restrict__ dy,
                                                        bf16* __restrict__ dh, long long rows, int C) {
  pdl_grid_sync();
  const int V = C / 8;
  const long long n = rows * V;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / V;
    const int v = (int)(i % V);
    const uint4 ua = ldg_stream16(h + r * 2 * C + v * 8), ug = ldg_stream16(h + r * 2 * C + C + v * 8);
    const uint4 ud = ldg_stream16(dy + r * C + v * 8);
    float a[8], g[8], d[8], da[8], dg[8];
    unpack8f(ua, a);
    unpack8f(ug, g);
    unpack8f(ud, d);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float cdf, pdf;
      gelu_cdf_pdf(g[e], cdf, pdf);
      da[e] = d[e] * (g[e] * cdf);
      dg[e] = d[e] * a[e] * fmaf(g[e], pdf, cdf);
    }
    st8(dh + r * 2 * C + v * 8, da);
    st8(dh + r * 2 * C + C + v * 8, dg);
  }
}

// GEGLU backward that also produces the bias gradient of the projection in front of it (dbias[2C] += column sums of dh),
// so the separate column-sum pass over dh - 46 % of all bias-gradient traffic of the step - disappears.  Layout: block =
// 32 column vectors x 8 row lanes over a chunk of rows (a warp still reads 512 contiguous bytes per operand), every thread
// keeps its 16 column sums in registers, the 8 row lanes meet in shared memory, one atomicAdd per column and block.
__global__ void __launch_bounds__(256) geglu_bwd_bias_kernel(const bf16* __restrict__ h, const bf16* __restrict__ dy,
                                                             bf16* __restrict__ dh, float* __restrict__ dbias, long long rows,
                                                             int C, long long rows_per_chunk) {
  pdl_grid_sync();
  __shared__ float sm[8][32][17];
  const int vx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int V = C / 8, v = blockIdx.x * 32 + vx;
  const bool act = v < V;
  const long long r0 = (long long)blockIdx.y * rows_per_chunk;
  long long r1 = r0 + rows_per_chunk;
  if (r1 > rows) r1 = rows;
  float sa[8], sg[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) sa[e] = sg[e] = 0.f;
  if (act) {
    for (long long r = r0 + ry; r < r1; r += 16) {
      const long long r2 = r + 8;
      const bool two = r2 < r1;
      const uint4 ua = ldg_stream16(h + r * 2 * C + v * 8), ug = ldg_stream16(h + r * 2 * C + C + v * 8);
      const uint4 ud = ldg_stream16(dy + r * C + v * 8);
      uint4 ua2 = ua, ug2 = ug, ud2 = ud;
      if (two) {
        ua2 = ldg_stream16(h + r2 * 2 * C + v * 8);
        ug2 = ldg_stream16(h + r2 * 2 * C + C + v * 8);
        ud2 = ldg_stream16(dy + r2 * C + v * 8);
      }
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        if (j == 1 && !two) break;
        float a[8], g[8], d[8], da[8], dg[8];
        unpack8f(j ? ua2 : ua, a);
        unpack8f(j ? ug2 : ug, g);
        unpack8f(j ? ud2 : ud, d);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          float cdf, pdf;
          gelu_cdf_pdf(g[e], cdf, pdf);
          da[e] = d[e] * (g[e] * cdf);
          dg[e] = d[e] * a[e] * fmaf(g[e], pdf, cdf);
          sa[e] += da[e];
          sg[e] += dg[e];
        }
        const long long rr = j ? r2 : r;
        st8(dh + rr * 2 * C + v * 8, da);
        st8(dh + rr * 2 * C + C + v * 8, dg);
      }
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    sm[ry][vx][e] = sa[e];
    sm[ry][vx][8 + e] = sg[e];
  }
  __syncthreads();
  // 512 column sums per block: thread t -> (vector t / 8 ... ) two passes of 256
  for (int i = threadIdx.x; i < 512; i += 256) {
    const int vv = i >> 4, e = i & 15;
    const int col_v = blockIdx.x * 32 + vv;
    if (col_v < V) {
      float t = 0.f;
#pragma unroll
      for (int r = 0; r < 8; ++r) t += sm[r][vv][e];
      atomicAdd(dbias + (e < 8 ? 0 : C) + col_v * 8 + (e & 7), t);
    }
  }
}

__global__ void silu_fwd_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, long long n8) {
  pdl_grid_sync();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    float f[8];
    ld8(x + i * 8, f);
#pragma unroll
    for (int e = 0; e < 8; ++e) f[e] = silu_f(f[e]);
    st8(y + i * 8, f);
  }
}
__global__ void silu_bwd_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dy, bf16* __restrict__ dx, long long n8) {
  pdl_grid_sync();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    float f[8], d[8];
    ld8(x + i * 8, f);
    ld8(dy + i * 8, d);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float s = sigmoid_f(f[e]);
      f[e] = d[e] * s * (1.f + f[e] * (1.f - s));
    }
    st8(dx + i * 8, f);
  }
}
__global__ void axpby_kernel(const bf16* __restrict__ a, float alpha, const bf16* __restrict__ b, float beta,
                             bf16* __restrict__ out, long long n8) {
  pdl_grid_sync();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    float f[8], g[8];
    ld8(a + i * 8, f);
    if (b) ld8(b + i * 8, g);
#pragma unroll
    for (int e = 0; e < 8; ++e) f[e] = alpha * f[e] + (b ? beta * g[e] : 0.f);
    st8(out + i * 8, f);
  }
}

// ---------------------------------------------------------------------------------------------- layout
__global__ void copy2d_kernel(const bf16* __restrict__ src, long long lds, bf16* __restrict__ dst, long long ldd,
                              long long rows, int cols, int accumulate) {
  pdl_grid_sync();
  const int V = cols / 8;
  const long long n = rows * V;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / V;
    const int v = (int)(i % V);
    if (accumulate) {
      float f[8], g[8];
      ld8(src + r * lds + v * 8, f);
      ld8(dst + r * ldd + v * 8, g);
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] += g[e];
      st8(dst + r * ldd + v * 8, f);
    } else {
      *reinterpret_cast<uint4*>(dst + r * ldd + v * 8) = *reinterpret_cast<const uint4*>(src + r * lds + v * 8);
    }
  }
}

// y[b][2h+i][2w+j][c] = x[b][h][w][c]
__global__ void upsample2x_fwd_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, int B, int H, int W, int C) {
  pdl_grid_sync();
  const int V = C / 8;
  const long long n = (long long)B * 2 * H * 2 * W * V;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % V);
    long long pix = i / V;
    const int ow = (int)(pix % (2 * W));
    pix /= 2 * W;
    const int oh = (int)(pix % (2 * H));
    const int b = (int)(pix / (2 * H));
    const long long src = (((long long)b * H + oh / 2) * W + ow / 2) * C + v * 8;
    *reinterpret_cast<uint4*>(y + (i / V) * C + v * 8) = *reinterpret_cast<const uint4*>(x + src);
  }
}
// dx[b][h][w][c] = sum_{i,j} dy[b][2h+i][2w+j][c]
__global__ void upsample2x_bwd_kernel(const bf16* __restrict__ dy, bf16* __restrict__ dx, int B, int H, int W, int C) {
  pdl_grid_sync();
  const int V = C / 8;
  const long long n = (long long)B * H * W * V;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % V);
    long long pix = i / V;
    const int w = (int)(pix % W);
    pix /= W;
    const int h = (int)(pix % H);
    const int b = (int)(pix / H);
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int di = 0; di < 2; ++di)
#pragma unroll
      for (int dj = 0; dj < 2; ++dj) {
        float f[8];
        ld8(dy + ((((long long)b * 2 * H + 2 * h + di) * 2 * W) + 2 * w + dj) * C + v * 8, f);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] += f[e];
      }
    st8(dx + (i / V) * C + v * 8, acc);
  }
}
// planes[(h%2)*2 + (w%2)][b][h/2][w/2][c] <-> x[b][h][w][c]
__global__ void phase_kernel(const bf16* __restrict__ src, bf16* __restrict__ dst, int B, int H, int W, int C, int merge) {
  pdl_grid_sync();
  const int V = C / 8;
  const long long n = (long long)B * H * W * V;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % V);
    long long pix = i / V;
    const int w = (int)(pix % W);
    pix /= W;
    const int h = (int)(pix % H);
    const int b = (int)(pix / H);
    const int plane = (h & 1) * 2 + (w & 1);
    const long long xoff = (i / V) * C + v * 8;
    const long long poff = ((((long long)plane * B + b) * (H / 2) + h / 2) * (W / 2) + w / 2) * C + v * 8;
    if (merge)
      *reinterpret_cast<uint4*>(dst + xoff) = *reinterpret_cast<const uint4*>(src + poff);
    else
      *reinterpret_cast<uint4*>(dst + poff) = *reinterpret_cast<const uint4*>(src + xoff);
  }
}

// ---------------------------------------------------------------------------------------------- column sums
// grid (ceil(N/64), groups, row_splits); block 256 = 8 column-vectors(8 ch) x 32 row lanes
__global__ void __launch_bounds__(256) colsum_kernel(const bf16* __restrict__ x, long long ldx, float* __restrict__ out,
                                                     long long ldo, long long rows_per_group, int N, int use_atomic) {
  pdl_grid_sync();
  __shared__ float sm[32][65];
  const int cv = threadIdx.x % 8, rl = threadIdx.x / 8;
  const int n0 = blockIdx.x * 64 + cv * 8;
  const int g = blockIdx.y;
  const long long chunk = (rows_per_group + gridDim.z - 1) / gridDim.z;
  const long long r0 = (long long)blockIdx.z * chunk;
  long long r1 = r0 + chunk;
  if (r1 > rows_per_group) r1 = rows_per_group;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (n0 < N) {
    const bf16* base = x + ((long long)g * rows_per_group) * ldx + n0;
    // four independent 16-byte loads in flight per thread (the loop is pure streaming: bytes in flight, not
    // arithmetic, set its speed)
    long long r = r0 + rl;
    for (; r + 96 < r1; r += 128) {
      uint4 u[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) u[j] = ldg_stream16(base + (r + 32 * j) * ldx);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 a = unpack_bf16x2(u[j].x), b = unpack_bf16x2(u[j].y), c = unpack_bf16x2(u[j].z), d = unpack_bf16x2(u[j].w);
        acc[0] += a.x; acc[1] += a.y; acc[2] += b.x; acc[3] += b.y;
        acc[4] += c.x; acc[5] += c.y; acc[6] += d.x; acc[7] += d.y;
      }
    }
    for (; r < r1; r += 32) {
      float f[8];
      ld8(base + r * ldx, f);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] += f[e];
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) sm[rl][cv * 8 + e] = acc[e];
  __syncthreads();
  if (threadIdx.x < 64) {
    float s = 0.f;
#pragma unroll 8
    for (int r = 0; r < 32; ++r) s += sm[r][threadIdx.x];
    const int n = blockIdx.x * 64 + threadIdx.x;
    if (n < N) {
      float* o = out + (long long)g * ldo + n;
      if (use_atomic)
        atomicAdd(o, s);
      else
        *o = s;
    }
  }
}

__global__ void fill_f32_kernel(float* p, long long n, float v) {
  pdl_grid_sync();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = v;
}

// ---------------------------------------------------------------------------------------------- casts
__global__ void cast_f32_bf16_kernel(const float* __restrict__ src, bf16* __restrict__ dst, long long n) {
  pdl_grid_sync();
  const long long n8 = n / 8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    const float4 a = *reinterpret_cast<const float4*>(src + i * 8), b = *reinterpret_cast<const float4*>(src + i * 8 + 4);
    uint4 u;
    u.x = pack_bf16x2(a.x, a.y); u.y = pack_bf16x2(a.z, a.w); u.z = pack_bf16x2(b.x, b.y); u.w = pack_bf16x2(b.z, b.w);
    *reinterpret_cast<uint4*>(dst + i * 8) = u;
  }
  if (blockIdx.x == 0 && threadIdx.x < (int)(n - n8 * 8)) dst[n8 * 8 + threadIdx.x] = __float2bfloat16_rn(src[n8 * 8 + threadIdx.x]);
}
__global__ void pad_cast_rows_kernel(const float* __restrict__ src, int cs, bf16* __restrict__ dst, int cd, long long rows) {
  pdl_grid_sync();
  const long long n = rows * cd;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cd;
    const int c = (int)(i % cd);
    dst[i] = __float2bfloat16_rn(c < cs ? src[r * cs + c] : 0.f);
  }
}
__global__ void unpad_accum_rows_kernel(const float* __restrict__ src, int cs, float* __restrict__ dst, int cd, long long rows,
                                        int accumulate) {
  pdl_grid_sync();
  const long long n = rows * cd;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cd;
    const int c = (int)(i % cd);
    const float v = src[r * cs + c];
    if (accumulate)
      atomicAdd(&dst[i], v);  // concurrent half-batch chains accumulate into the same gradient
    else
      dst[i] = v;
  }
}

// ---------------------------------------------------------------------------------------------- MSE head
template <typename T>
__device__ __forceinline__ float to_f(T v);
template <>
__device__ __forceinline__ float to_f<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <>
__device__ __forceinline__ float to_f<__half>(__half v) { return __half2float(v); }
template <typename T>
__device__ __forceinline__ T from_f(float v);
template <>
__device__ __forceinline__ float from_f<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <>
__device__ __forceinline__ __half from_f<__half>(float v) { return __float2half_rn(v); }

// one thread per pixel: reads the 4 valid channels of the conv_out tile, the 4 NCHW noise planes
template <typename T>
__global__ void __launch_bounds__(256) mse_head_kernel(const bf16* __restrict__ pred8, const T* __restrict__ noise,
                                                       T* __restrict__ pred_nchw, bf16* __restrict__ dpred8,
                                                       float* __restrict__ loss_acc, float gscale, int B, int HW) {
  pdl_grid_sync();
  const long long npix = (long long)B * HW;
  const float k = gscale * 2.f / (float)(npix * 4);
  float local = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < npix; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / HW, hw = i % HW;
    float p[8];
    ld8(pred8 + i * 8, p);
    float d[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const long long j = (b * 4 + c) * HW + hw;
      const float diff = p[c] - to_f<T>(noise[j]);
      local += diff * diff;
      d[c] = k * diff;
      if (pred_nchw) pred_nchw[j] = from_f<T>(p[c]);
    }
    if (dpred8) st8(dpred8 + i * 8, d);
  }
  local = warp_sum(local);
  __shared__ float sm[8];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += sm[w];
    atomicAdd(&loss_acc[0], s);
    if (blockIdx.x == 0) atomicAdd(&loss_acc[1], (float)(npix * 4));
  }
}

}  // namespace sd2

using namespace sd2;
#define SD2_STREAM cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_)
#define SD2_BF(p) reinterpret_cast<const bf16*>(p)
#define SD2_BFW(p) reinterpret_cast<bf16*>(p)

// ---- nearest-neighbour x2 upsample folded into the 3x3 convolution that follows it (Upsample2D)
// conv3x3(up2(x))[2h+py, 2w+px] only sees 2 x 2 distinct low-resolution pixels: the 3 kernel rows collapse to two row groups
// (py = 0: {0} at dh = -1, {1, 2} at dh = 0; py = 1: {0, 1} at dh = 0, {2} at dh = +1), likewise the columns.  Each of the four
// output phases is therefore a 4-tap convolution of x with summed weights: 16 instead of 36 tap-products per low-res pixel.
// weff[phase = py*2+px][tap = a*2+b][Cout][Cin] (bf16) = sum of the fp32 master taps of row group a x column group b.
namespace sd2 {
__device__ __forceinline__ int upconv_group(int parity, int k) { return parity == 0 ? (k >= 1) : (k >= 2); }
__global__ void __launch_bounds__(256) upconv_weff_build_kernel(const float* __restrict__ w9, bf16* __restrict__ weff, long long n) {
  pdl_grid_sync();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float w[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) w[t] = w9[t * n + i];
#pragma unroll
    for (int ph = 0; ph < 4; ++ph) {
      const int py = ph >> 1, px = ph & 1;
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) acc[upconv_group(py, ky) * 2 + upconv_group(px, kx)] += w[ky * 3 + kx];
#pragma unroll
      for (int t = 0; t < 4; ++t) weff[(ph * 4 + t) * n + i] = __float2bfloat16_rn(acc[t]);
    }
  }
}
// dw9[ky*3+kx] += sum over the four phases of dweff[phase][group tap that contains (ky, kx)]   (fp32, accumulating)
__global__ void __launch_bounds__(256) upconv_wgrad_scatter_kernel(const float* __restrict__ dweff, float* __restrict__ dw9, long long n) {
  pdl_grid_sync();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float d[16];
#pragma unroll
    for (int t = 0; t < 16; ++t) d[t] = dweff[t * n + i];
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        float g = 0.f;
#pragma unroll
        for (int ph = 0; ph < 4; ++ph) g += d[ph * 4 + upconv_group(ph >> 1, ky) * 2 + upconv_group(ph & 1, kx)];
        dw9[(ky * 3 + kx) * n + i] += g;
      }
  }
}
}  // namespace sd2

extern "C" {

int sd2_softmax_fwd(sd2_ctx* ctx, const float* S, long long lds, void* P, long long ldp, long long rows, int cols,
                    sd2_stream stream_) {
  if (!ctx) return 1;
  SD2_STREAM;
  launch_k(softmax_fwd_kernel, dim3(grid_for(rows * 32, 256, ctx->num_sms, 16)), dim3(256), 0, stream, S, lds, SD2_BFW(P), ldp, rows, cols);
  return check_launch(ctx, "softmax_fwd");
}
int sd2_softmax_bwd(sd2_ctx* ctx, const void* P, long long ldp, const float* dP, long long lddp, void* dS,
                    long long ldds, long long rows, int cols, float scale, sd2_stream stream_) {
  if (!ctx) return 1;
  SD2_STREAM;
  launch_k(softmax_bwd_kernel, dim3(grid_for(rows * 32, 256, ctx->num_sms, 16)), dim3(256), 0, stream, SD2_BF(P), ldp, dP, lddp, SD2_BFW(dS),
                                                                                    ldds, rows, cols, scale);
  return check_launch(ctx, "softmax_bwd");
}
// grid cap: 32 waves of blocks (measured, tools/sweep_hbm.sh: 8 -> 32 waves takes the backward from 84 % to 93 % and the
// forward from 67 % to 74 % of the HBM copy rate at the B=128 shapes)
int sd2_geglu_fwd(sd2_ctx* ctx, const void* h, void* y, long long rows, int C, sd2_stream stream_) {
  if (!ctx) return 1;
  if (C % 8) return fail(ctx, "geglu: C % 8");
  SD2_STREAM;
  launch_k(geglu_fwd_kernel, dim3(grid_for(rows * (C / 8), 256, ctx->num_sms, 32)), dim3(256), 0, stream, SD2_BF(h), SD2_BFW(y), rows, C);
  return check_launch(ctx, "geglu_fwd");
}
int sd2_geglu_bwd(sd2_ctx* ctx, const void* h, const void* dy, void* dh, float* dbias, long long rows, int C, sd2_stream stream_) {
  if (!ctx) return 1;
  if (C % 8) return fail(ctx, "geglu: C % 8");
  SD2_STREAM;
  if (dbias == nullptr) {
    launch_k(geglu_bwd_kernel, dim3(grid_for(rows * (C / 8), 256, ctx->num_sms, 32)), dim3(256), 0, stream, SD2_BF(h), SD2_BF(dy), SD2_BFW(dh), rows, C);
    return check_launch(ctx, "geglu_bwd");
  }
  const int ncb = (C / 8 + 31) / 32;
  long long nrc = (24LL * ctx->num_sms + ncb - 1) / ncb;  // ~24 blocks per SM like the plain kernel's measured optimum
  const long long max_rc = (rows + 31) / 32;              // at least 32 rows (two unrolled iterations of the 8 lanes) per chunk
  if (nrc > max_rc) nrc = max_rc;
  if (nrc < 1) nrc = 1;
  if (nrc > 65535) nrc = 65535;
  const long long rpc = ((rows + nrc - 1) / nrc + 7) / 8 * 8;
  nrc = (rows + rpc - 1) / rpc;
  launch_k(geglu_bwd_bias_kernel, dim3(ncb, (unsigned)nrc), dim3(256), 0, stream, SD2_BF(h), SD2_BF(dy), SD2_BFW(dh), dbias, rows, C, rpc);
  return check_launch(ctx, "geglu_bwd");
}
int sd2_silu_fwd(sd2_ctx* ctx, const void* x, void* y, long long n, sd2_stream stream_) {
  if (!ctx) return 1;
  if (n % 8) return fail(ctx, "silu: n % 8");
  SD2_STREAM;
  launch_k(silu_fwd_kernel, dim3(grid_for(n / 8, 256, ctx->num_sms)), dim3(256), 0, stream, SD2_BF(x), SD2_BFW(y), n / 8);
  return check_launch(ctx, "silu_fwd");
}
int sd2_silu_bwd(sd2_ctx* ctx, const void* x, const void* dy, void* dx, long long n, sd2_stream stream_) {
  if (!ctx) return 1;
  if (n % 8) return fail(ctx, "silu: n % 8");
  SD2_STREAM;
  launch_k(silu_bwd_kernel, dim3(grid_for(n / 8, 256, ctx->num_sms)), dim3(256), 0, stream, SD2_BF(x), SD2_BF(dy), SD2_BFW(dx), n / 8);
  return check_launch(ctx, "silu_bwd");
}
int sd2_axpby(sd2_ctx* ctx, const void* a, float alpha, const void* b, float beta, void* out, long long n,
              sd2_stream stream_) {
  if (!ctx) return 1;
  if (n % 8) return fail(ctx, "axpby: n % 8");
  SD2_STREAM;
  launch_k(axpby_kernel, dim3(grid_for(n / 8, 256, ctx->num_sms)), dim3(256), 0, stream, SD2_BF(a), alpha, SD2_BF(b), beta, SD2_BFW(out), n / 8);
  return check_launch(ctx, "axpby");
}
int sd2_copy2d(sd2_ctx* ctx, const void* src, long long lds, void* dst, long long ldd, long long rows, int cols,
               int accumulate, sd2_stream stream_) {
  if (!ctx) return 1;
  if (cols % 8 || lds % 8 || ldd % 8) return fail(ctx, "copy2d: cols/ld % 8");
  SD2_STREAM;
  launch_k(copy2d_kernel, dim3(grid_for(rows * (cols / 8), 256, ctx->num_sms)), dim3(256), 0, stream, SD2_BF(src), lds, SD2_BFW(dst), ldd, rows,
                                                                                  cols, accumulate);
  return check_launch(ctx, "copy2d");
}
int sd2_upsample2x_fwd(sd2_ctx* ctx, const void* x, void* y, int B, int H, int W, int C, sd2_stream stream_) {
  if (!ctx) return 1;
  if (C % 8) return fail(ctx, "upsample: C % 8");
  SD2_STREAM;
  launch_k(upsample2x_fwd_kernel, dim3(grid_for((long long)B * 4 * H * W * (C / 8), 256, ctx->num_sms)), dim3(256), 0, stream, SD2_BF(x), SD2_BFW(y), B, H, W, C);
  return check_launch(ctx, "upsample2x_fwd");
}
int sd2_upsample2x_bwd(sd2_ctx* ctx, const void* dy, void* dx, int B, int H, int W, int C, sd2_stream stream_) {
  if (!ctx) return 1;
  if (C % 8) return fail(ctx, "upsample: C % 8");
  SD2_STREAM;
  launch_k(upsample2x_bwd_kernel, dim3(grid_for((long long)B * H * W * (C / 8), 256, ctx->num_sms)), dim3(256), 0, stream, SD2_BF(dy), SD2_BFW(dx),
                                                                                                       B, H, W, C);
  return check_launch(ctx, "upsample2x_bwd");
}
int sd2_upconv_weff_build(sd2_ctx* ctx, const float* w9, void* weff16, long long n, sd2_stream stream_) {
  if (!ctx) return 1;
  if (n <= 0) return fail(ctx, "upconv_weff_build: n <= 0");
  SD2_STREAM;
  launch_k(upconv_weff_build_kernel, dim3(grid_for(n, 256, ctx->num_sms)), dim3(256), 0, stream, w9, SD2_BFW(weff16), n);
  return check_launch(ctx, "upconv_weff_build");
}
int sd2_upconv_wgrad_scatter(sd2_ctx* ctx, const float* dweff16, float* dw9, long long n, sd2_stream stream_) {
  if (!ctx) return 1;
  if (n <= 0) return fail(ctx, "upconv_wgrad_scatter: n <= 0");
  SD2_STREAM;
  launch_k(upconv_wgrad_scatter_kernel, dim3(grid_for(n, 256, ctx->num_sms)), dim3(256), 0, stream, dweff16, dw9, n);
  return check_launch(ctx, "upconv_wgrad_scatter");
}
int sd2_phase_split(sd2_ctx* ctx, const void* x, void* planes, int B, int H, int W, int C, sd2_stream stream_) {
  if (!ctx) return 1;
  if (C % 8 || H % 2 || W % 2) return fail(ctx, "phase_split: C % 8 or odd H/W");
  SD2_STREAM;
  launch_k(phase_kernel, dim3(grid_for((long long)B * H * W * (C / 8), 256, ctx->num_sms)), dim3(256), 0, stream, SD2_BF(x), SD2_BFW(planes), B, H,
                                                                                              W, C, 0);
  return check_launch(ctx, "phase_split");
}
int sd2_phase_merge(sd2_ctx* ctx, const void* planes, void* x, int B, int H, int W, int C, sd2_stream stream_) {
  if (!ctx) return 1;
  if (C % 8 || H % 2 || W % 2) return fail(ctx, "phase_merge: C % 8 or odd H/W");
  SD2_STREAM;
  launch_k(phase_kernel, dim3(grid_for((long long)B * H * W * (C / 8), 256, ctx->num_sms)), dim3(256), 0, stream, SD2_BF(planes), SD2_BFW(x), B, H,
                                                                                              W, C, 1);
  return check_launch(ctx, "phase_merge");
}
int sd2_colsum(sd2_ctx* ctx, const void* x, long long ldx, float* out, long long ldo, int groups,
               long long rows_per_group, int N, int accumulate, sd2_stream stream_) {
  if (!ctx) return 1;
  if (N % 8 || ldx % 8) return fail(ctx, "colsum: N/ldx % 8");
  SD2_STREAM;
  const int nblk = (N + 63) / 64;
  // blocks per SM: measured at the B=128 shapes (SD2_COLSUM_MULT sweep) 6 -> 24 takes the 131072 x 2560 sum (the GEGLU
  // projection bias, 46 % of all column-sum traffic) from 73 % to 98 % of the HBM copy rate and the C = 320 sums from 60 to 68 %
  static int mult = -1;
  if (mult < 0) {
    const char* e = getenv("SD2_COLSUM_MULT");
    mult = e ? atoi(e) : 24;
    if (mult < 1) mult = 24;
  }
  long long splits = ((long long)mult * ctx->num_sms) / ((long long)nblk * groups);
  const long long max_splits = (rows_per_group + 255) / 256;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  int launches = 1;
  if (!accumulate && splits > 1) {
    // zero the destination rows first so that the split partials can be combined with atomics
    if (ldo == N || groups == 1) {  // contiguous destination rows: one launch for all groups
      launch_k(fill_f32_kernel, dim3(grid_for((long long)groups * N, 256, ctx->num_sms)), dim3(256), 0, stream, out, (long long)groups * N, 0.f);
      ++launches;
    } else {
      for (int g = 0; g < groups; ++g) {
        launch_k(fill_f32_kernel, dim3(grid_for(N, 256, ctx->num_sms)), dim3(256), 0, stream, out + (long long)g * ldo, N, 0.f);
        ++launches;
      }
    }
  }
  const int use_atomic = (accumulate || splits > 1) ? 1 : 0;
  launch_k(colsum_kernel, dim3(dim3(nblk, groups, (unsigned)splits)), dim3(256), 0, stream, SD2_BF(x), ldx, out, ldo, rows_per_group, N, use_atomic);
  return check_launch(ctx, "colsum", launches);
}
int sd2_cast_f32_to_bf16(sd2_ctx* ctx, const float* src, void* dst, long long n, sd2_stream stream_) {
  if (!ctx) return 1;
  SD2_STREAM;
  launch_k(cast_f32_bf16_kernel, dim3(grid_for(n / 8 + 1, 256, ctx->num_sms)), dim3(256), 0, stream, src, SD2_BFW(dst), n);
  return check_launch(ctx, "cast_f32_to_bf16");
}
int sd2_pad_cast_rows(sd2_ctx* ctx, const float* src, int cols_src, void* dst, int cols_dst, long long rows,
                      sd2_stream stream_) {
  if (!ctx) return 1;
  SD2_STREAM;
  launch_k(pad_cast_rows_kernel, dim3(grid_for(rows * cols_dst, 256, ctx->num_sms)), dim3(256), 0, stream, src, cols_src, SD2_BFW(dst), cols_dst, rows);
  return check_launch(ctx, "pad_cast_rows");
}
int sd2_unpad_accum_rows(sd2_ctx* ctx, const float* src, int cols_src, float* dst, int cols_dst, long long rows,
                         int accumulate, sd2_stream stream_) {
  if (!ctx) return 1;
  SD2_STREAM;
  launch_k(unpad_accum_rows_kernel, dim3(grid_for(rows * cols_dst, 256, ctx->num_sms)), dim3(256), 0, stream, src, cols_src, dst, cols_dst, rows,
                                                                                          accumulate);
  return check_launch(ctx, "unpad_accum_rows");
}
int sd2_mse_head(sd2_ctx* ctx, const void* pred_nhwc8, const void* noise, int noise_dtype, void* pred_nchw,
                 void* dpred_nhwc8, float* loss_acc, float gscale, int B, int H, int W, sd2_stream stream_) {
  if (!ctx) return 1;
  SD2_STREAM;
  const int blocks = grid_for((long long)B * H * W, 256, ctx->num_sms, 4);
#define MSE_LAUNCH(T)                                                                                                  \
  launch_k(mse_head_kernel<T>, dim3(blocks), dim3(256), 0, stream, SD2_BF(pred_nhwc8), reinterpret_cast<const T*>(noise),                 \
                                                 reinterpret_cast<T*>(pred_nchw), SD2_BFW(dpred_nhwc8), loss_acc, gscale, B, H * W)
  if (noise_dtype == SD2_DT_F32) {
    MSE_LAUNCH(float);
  } else if (noise_dtype == SD2_DT_BF16) {
    MSE_LAUNCH(__nv_bfloat16);
  } else if (noise_dtype == SD2_DT_F16) {
    MSE_LAUNCH(__half);
  } else {
    return fail(ctx, "mse_head: dtype");
  }
#undef MSE_LAUNCH
  return check_launch(ctx, "mse_head");
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------- fused AdamW
// One pass over a flat fp32 parameter range: decoupled weight decay + Adam moments + update (torch.optim.AdamW
// semantics, reference yaml SD-2-base-256.yaml:55-58), the bf16 shadow copy the tensor-core kernels read, and the
// reset of the gradient buffer for the next accumulation.  HBM-bound: 16 B read + 18..22 B written per parameter.
namespace sd2 {
__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
                                                    float* __restrict__ v, bf16* __restrict__ p16, long long n, float lr,
                                                    float beta1, float beta2, float eps, float wd, float bc1, float rsqrt_bc2,
                                                    float gscale, int zero_grad) {
  pdl_grid_sync();
  const long long n4 = n / 4;
  const float step_size = lr / bc1, decay = 1.f - lr * wd;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    float4 gg = reinterpret_cast<const float4*>(g)[i];
    float4 mm = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    float* pf = reinterpret_cast<float*>(&pp);
    float* gf = reinterpret_cast<float*>(&gg);
    float* mf = reinterpret_cast<float*>(&mm);
    float* vf = reinterpret_cast<float*>(&vv);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float gr = gf[e] * gscale;
      pf[e] *= decay;
      mf[e] = beta1 * mf[e] + (1.f - beta1) * gr;
      vf[e] = beta2 * vf[e] + (1.f - beta2) * gr * gr;
      const float denom = sqrtf(vf[e]) * rsqrt_bc2 + eps;
      pf[e] -= step_size * (mf[e] / denom);
    }
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
    if (p16) {
      uint2 o;
      o.x = pack_bf16x2(pf[0], pf[1]);
      o.y = pack_bf16x2(pf[2], pf[3]);
      reinterpret_cast<uint2*>(p16)[i] = o;
    }
    if (zero_grad) reinterpret_cast<float4*>(g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  // tail (n not a multiple of 4)
  const long long t = n4 * 4 + blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t < n) {
    const float gr = g[t] * gscale;
    float pv = p[t] * decay;
    const float mv = beta1 * m[t] + (1.f - beta1) * gr;
    const float vv2 = beta2 * v[t] + (1.f - beta2) * gr * gr;
    pv -= step_size * (mv / (sqrtf(vv2) * rsqrt_bc2 + eps));
    p[t] = pv;
    m[t] = mv;
    v[t] = vv2;
    if (p16) p16[t] = __float2bfloat16_rn(pv);
    if (zero_grad) g[t] = 0.f;
  }
}
}  // namespace sd2

extern "C" int sd2_adamw_step(sd2_ctx* ctx, float* param, float* grad, float* exp_avg, float* exp_avg_sq, void* param_bf16,
                              long long n, float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                              float grad_scale, int zero_grad, sd2_stream stream_) {
  if (!ctx) return 1;
  if (n <= 0 || step < 1) return fail(ctx, "sd2_adamw_step: n <= 0 or step < 1");
  if ((reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(grad) | reinterpret_cast<uintptr_t>(exp_avg) |
       reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15)
    return fail(ctx, "sd2_adamw_step: buffers must be 16-byte aligned");
  SD2_STREAM;
  const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
  launch_k(adamw_kernel, dim3(grid_for(n / 4 + 1, 256, ctx->num_sms, 16)), dim3(256), 0, stream, param, grad, exp_avg, exp_avg_sq, SD2_BFW(param_bf16), n, lr, beta1, beta2, eps, weight_decay, (float)bc1,
      (float)(1.0 / sqrt(bc2)), grad_scale, zero_grad);
  return check_launch(ctx, "adamw_step");
}

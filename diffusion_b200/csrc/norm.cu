// GroupNorm(+SiLU) and LayerNorm, forward and backward, NHWC bf16 I/O with fp32 statistics.  HBM-bound kernels:
// 16-byte vector loads (8 bf16 channels per thread), coalesced along the channel dimension, deterministic
// two-level reductions through a caller-provided fp32 scratch (no atomics on global memory).
//
// Replaces ATen group_norm / layer_norm (+ their backward) as run by composer's LPGroupNorm / LPLayerNorm surgery
// (reference diffusion/train.py:91-108) and the F.silu that follows every ResnetBlock2D GroupNorm (diffusers).
#include "common.cuh"
#include "host.h"

namespace sd2 {

static constexpr int GN_THREADS = 512;
static constexpr int GN_MAXP = 64;  // pixel-chunk partials per image

__device__ __forceinline__ void load8(const bf16* p, float* v) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
}
__device__ __forceinline__ void store8(bf16* p, const float* v) {
  uint4 u;
  u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]); u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(p) = u;
}

// Pixel chunks per image: B * P blocks should fill two waves of (2 resident blocks per SM) without a ragged tail.
static inline int gn_chunks(int HW, int B, int C, int num_sms) {
  int p = (4 * num_sms) / (B > 0 ? B : 1);
  const int R = GN_THREADS / (C / 8);       // row lanes per block
  const int pmax = HW / (4 * (R > 0 ? R : 1));  // keep >= 4 rows per thread (the row loops are unrolled by 4)
  if (p > pmax) p = pmax;
  if (p > GN_MAXP) p = GN_MAXP;
  return p < 1 ? 1 : p;
}

// ------------------------------------------------------------------------------------------------ GroupNorm fwd
// Cross-row-lane reduction used by the two statistics kernels: every thread holds 2 x 8 partial sums for channels
// v*8..v*8+7; lanes r = 0..R-1 of the same channel vector are summed through shared memory without atomics:
// sm[r][c][2] staging, then thread i < 2C adds its R entries into red[i] (red = sm + R*2C).
__device__ __forceinline__ void gn_block_reduce(float* sm, const float* a, const float* b, int C, int R, int v, int r) {
  if (r < R) {
    float* o = sm + ((size_t)r * C + v * 8) * 2;
#pragma unroll
    for (int e = 0; e < 8; e += 2) *reinterpret_cast<float4*>(o + 2 * e) = make_float4(a[e], b[e], a[e + 1], b[e + 1]);
  }
  __syncthreads();
  float* red = sm + (size_t)R * C * 2;
  for (int i = threadIdx.x; i < 2 * C; i += GN_THREADS) {
    float t = 0.f;
    for (int rr = 0; rr < R; ++rr) t += sm[(size_t)rr * 2 * C + i];
    red[i] = t;
  }
  __syncthreads();
}

// grid (P, B). Partial (sum, sumsq) per group over this block's pixel chunk -> ws[b][p][G][2]
__global__ void __launch_bounds__(GN_THREADS, 2) gn_stats_kernel(const bf16* __restrict__ x, long long ldx, float* __restrict__ ws,
                                                              int HW, int C, int G) {
  pdl_grid_sync();
  extern __shared__ float sm[];  // [R][C][2] + [C][2]
  const int P = gridDim.x, p = blockIdx.x, b = blockIdx.y;
  const int V = C / 8, R = GN_THREADS / V;
  const int v = threadIdx.x % V, r = threadIdx.x / V;
  const int row0 = (int)((long long)p * HW / P), row1 = (int)((long long)(p + 1) * HW / P);
  float s[8], q[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) s[e] = q[e] = 0.f;
  if (r < R) {
    const bf16* base = x + ((long long)b * HW) * ldx + v * 8;
#pragma unroll 4
    for (int row = row0 + r; row < row1; row += R) {
      float f[8];
      load8(base + (long long)row * ldx, f);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        s[e] += f[e];
        q[e] += f[e] * f[e];
      }
    }
  }
  gn_block_reduce(sm, s, q, C, R, v, r);
  const float* red = sm + (size_t)R * C * 2;
  const int cpg = C / G;
  for (int g = threadIdx.x; g < G; g += GN_THREADS) {
    float a = 0.f, c2 = 0.f;
    for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
      a += red[c * 2];
      c2 += red[c * 2 + 1];
    }
    float* o = ws + (((long long)b * P + p) * G + g) * 2;
    o[0] = a;
    o[1] = c2;
  }
}

// grid (P, B): finalize stats from the P partials, then y = [silu](x * scale + shift)
__global__ void __launch_bounds__(GN_THREADS, 2) gn_apply_kernel(const bf16* __restrict__ x, long long ldx,
                                                              const float* __restrict__ gamma, const float* __restrict__ beta,
                                                              bf16* __restrict__ y, long long ldy, float* __restrict__ stats,
                                                              const float* __restrict__ ws, int HW, int C, int G, float eps,
                                                              int silu) {
  pdl_grid_sync();
  __shared__ float s_mean[64], s_rstd[64];
  const int P = gridDim.x, p = blockIdx.x, b = blockIdx.y;
  const int cpg = C / G;
  if (threadIdx.x < G) {
    const int g = threadIdx.x;
    float s = 0.f, q = 0.f;
    for (int i = 0; i < P; ++i) {
      const float* o = ws + (((long long)b * P + i) * G + g) * 2;
      s += o[0];
      q += o[1];
    }
    const float n = (float)cpg * (float)HW;
    const float mean = s / n;
    const float var = fmaxf(q / n - mean * mean, 0.f);
    const float rstd = rsqrtf(var + eps);
    s_mean[g] = mean;
    s_rstd[g] = rstd;
    if (p == 0) {
      stats[((long long)b * G + g) * 2] = mean;
      stats[((long long)b * G + g) * 2 + 1] = rstd;
    }
  }
  __syncthreads();
  const int V = C / 8, R = GN_THREADS / V;
  const int v = threadIdx.x % V, r = threadIdx.x / V;
  if (r >= R) return;
  float sc[8], sh[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int c = v * 8 + e, g = c / cpg;
    sc[e] = gamma[c] * s_rstd[g];
    sh[e] = beta[c] - s_mean[g] * sc[e];
  }
  const int row0 = (int)((long long)p * HW / P), row1 = (int)((long long)(p + 1) * HW / P);
  const bf16* xb = x + ((long long)b * HW) * ldx + v * 8;
  bf16* yb = y + ((long long)b * HW) * ldy + v * 8;
#pragma unroll 4
  for (int row = row0 + r; row < row1; row += R) {
    float f[8];
    load8(xb + (long long)row * ldx, f);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float z = f[e] * sc[e] + sh[e];
      f[e] = silu ? silu_f(z) : z;
    }
    store8(yb + (long long)row * ldy, f);
  }
}

// ------------------------------------------------------------------------------------------------ GroupNorm bwd
__device__ __forceinline__ float silu_grad(float z) {
  const float s = sigmoid_f(z);
  return s * (1.f + z * (1.f - s));
}

// grid (P, B): per-channel partial sums of dyh and dyh * xhat -> ws[b][p][C][2]
__global__ void __launch_bounds__(GN_THREADS, 1) gn_bwd_stats_kernel(const bf16* __restrict__ dy, long long lddy,
                                                                  const bf16* __restrict__ x, long long ldx,
                                                                  const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                  const float* __restrict__ stats, float* __restrict__ ws, int HW,
                                                                  int C, int G, int silu) {
  pdl_grid_sync();
  extern __shared__ float sm[];  // [R][C][2] + [C][2]
  const int P = gridDim.x, p = blockIdx.x, b = blockIdx.y;
  const int V = C / 8, R = GN_THREADS / V, cpg = C / G;
  const int v = threadIdx.x % V, r = threadIdx.x / V;
  const int row0 = (int)((long long)p * HW / P), row1 = (int)((long long)(p + 1) * HW / P);
  float s1[8], s2[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) s1[e] = s2[e] = 0.f;
  if (r < R) {
    float mean[8], rstd[8], ga[8], be[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int c = v * 8 + e, g = c / cpg;
      mean[e] = stats[((long long)b * G + g) * 2];
      rstd[e] = stats[((long long)b * G + g) * 2 + 1];
      ga[e] = gamma[c];
      be[e] = beta[c];
    }
    const bf16* xb = x + ((long long)b * HW) * ldx + v * 8;
    const bf16* db = dy + ((long long)b * HW) * lddy + v * 8;
#pragma unroll 4
    for (int row = row0 + r; row < row1; row += R) {
      float f[8], d[8];
      load8(xb + (long long)row * ldx, f);
      load8(db + (long long)row * lddy, d);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float xh = (f[e] - mean[e]) * rstd[e];
        float dyh = d[e];
        if (silu) dyh *= silu_grad(xh * ga[e] + be[e]);
        s1[e] += dyh;
        s2[e] += dyh * xh;
      }
    }
  }
  gn_block_reduce(sm, s1, s2, C, R, v, r);
  const float* red = sm + (size_t)R * C * 2;
  float* o = ws + (((long long)b * P + p) * C) * 2;
  for (int i = threadIdx.x; i < 2 * C; i += GN_THREADS) o[i] = red[i];
}

// grid (G), block 256: one block per group.  Sums the per-chunk partials of its cpg channels for every image
// (T[b][cl] in shared memory), then (a) per image: db_g = sum gamma_c * T1, ds_g = sum gamma_c * T2 -> gstat[b][g]
// (consumed by the apply kernel), (b) per channel: dbeta_c += sum_b T1, dgamma_c += sum_b T2.
__global__ void __launch_bounds__(256) gn_bwd_reduce_kernel(const float* __restrict__ ws, const float* __restrict__ gamma,
                                                            float* __restrict__ gstat, float* __restrict__ dgamma,
                                                            float* __restrict__ dbeta, int B, int P, int C, int G) {
  pdl_grid_sync();
  extern __shared__ float T[];  // [B][cpg][2]
  const int g = blockIdx.x, cpg = C / G;
  for (int idx = threadIdx.x; idx < B * cpg; idx += blockDim.x) {
    const int b = idx / cpg, cl = idx % cpg;
    const float* o = ws + (((long long)b * P) * C + g * cpg + cl) * 2;
    float a1 = 0.f, a2 = 0.f;
    for (int pp = 0; pp < P; ++pp) {
      const float2 t = *reinterpret_cast<const float2*>(o + (long long)pp * C * 2);
      a1 += t.x;
      a2 += t.y;
    }
    T[idx * 2] = a1;
    T[idx * 2 + 1] = a2;
  }
  __syncthreads();
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    float dbv = 0.f, ds = 0.f;
    for (int cl = 0; cl < cpg; ++cl) {
      const float ga = gamma[g * cpg + cl];
      dbv += ga * T[(b * cpg + cl) * 2];
      ds += ga * T[(b * cpg + cl) * 2 + 1];
    }
    gstat[((long long)b * G + g) * 2] = dbv;
    gstat[((long long)b * G + g) * 2 + 1] = ds;
  }
  for (int cl = threadIdx.x; cl < cpg; cl += blockDim.x) {
    float a1 = 0.f, a2 = 0.f;
    for (int b = 0; b < B; ++b) {
      a1 += T[(b * cpg + cl) * 2];
      a2 += T[(b * cpg + cl) * 2 + 1];
    }
    atomicAdd(&dbeta[g * cpg + cl], a1);  // atomics: two half-batch chains may accumulate concurrently (DualEngine)
    atomicAdd(&dgamma[g * cpg + cl], a2);
  }
}

// grid (P, B): dx = rstd * (gamma*dyh - (db_g + xhat*ds_g)/n) (+ dx_add)
__global__ void __launch_bounds__(GN_THREADS, 1) gn_bwd_apply_kernel(const bf16* __restrict__ dy, long long lddy,
                                                                  const bf16* __restrict__ x, long long ldx,
                                                                  const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                  const float* __restrict__ stats, const float* __restrict__ gstat,
                                                                  const bf16* __restrict__ dx_add, long long ldadd,
                                                                  bf16* __restrict__ dx, long long lddx, int HW, int C, int G,
                                                                  int silu) {
  pdl_grid_sync();
  __shared__ float s_ds[64], s_db[64];
  const int P = gridDim.x, p = blockIdx.x, b = blockIdx.y;
  const int cpg = C / G;
  if (threadIdx.x < G) {
    s_db[threadIdx.x] = gstat[((long long)b * G + threadIdx.x) * 2];
    s_ds[threadIdx.x] = gstat[((long long)b * G + threadIdx.x) * 2 + 1];
  }
  __syncthreads();
  const int V = C / 8, R = GN_THREADS / V;
  const int v = threadIdx.x % V, r = threadIdx.x / V;
  if (r >= R) return;
  const float inv_n = 1.f / ((float)cpg * (float)HW);
  float mean[8], rstd[8], ga[8], be[8], k1[8], k2[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int c = v * 8 + e, g = c / cpg;
    mean[e] = stats[((long long)b * G + g) * 2];
    rstd[e] = stats[((long long)b * G + g) * 2 + 1];
    ga[e] = gamma[c];
    be[e] = beta[c];
    k1[e] = s_db[g] * inv_n;
    k2[e] = s_ds[g] * inv_n;
  }
  const int row0 = (int)((long long)p * HW / P), row1 = (int)((long long)(p + 1) * HW / P);
  const bf16* xb = x + ((long long)b * HW) * ldx + v * 8;
  const bf16* db = dy + ((long long)b * HW) * lddy + v * 8;
  bf16* ob = dx + ((long long)b * HW) * lddx + v * 8;
  const bf16* ab = dx_add ? dx_add + ((long long)b * HW) * ldadd + v * 8 : nullptr;
#pragma unroll 4
  for (int row = row0 + r; row < row1; row += R) {
    float f[8], d[8], a[8];
    load8(xb + (long long)row * ldx, f);
    load8(db + (long long)row * lddy, d);
    if (ab) load8(ab + (long long)row * ldadd, a);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float xh = (f[e] - mean[e]) * rstd[e];
      float dyh = d[e];
      if (silu) dyh *= silu_grad(xh * ga[e] + be[e]);
      float o = rstd[e] * (ga[e] * dyh - k1[e] - xh * k2[e]);
      if (ab) o += a[e];
      f[e] = o;
    }
    store8(ob + (long long)row * lddx, f);
  }
}

// dgamma[c] += sum_i ws[i][c][1], dbeta[c] += sum_i ws[i][c][0] over `n_part` partial slabs.
// grid (ceil(C/32)), block (32 channels, 8 partial lanes): coalesced float2 reads, smem tree over the 8 lanes.
__global__ void __launch_bounds__(256) affine_grad_reduce_kernel(const float* __restrict__ ws, int n_part, int C,
                                                                 float* __restrict__ dgamma, float* __restrict__ dbeta) {
  pdl_grid_sync();
  __shared__ float sm[8][32][2];
  const int cl = threadIdx.x & 31, pl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  float a1 = 0.f, a2 = 0.f;
  if (c < C) {
    for (int i = pl; i < n_part; i += 8) {
      const float2 t = *reinterpret_cast<const float2*>(ws + ((long long)i * C + c) * 2);
      a1 += t.x;
      a2 += t.y;
    }
  }
  sm[pl][cl][0] = a1;
  sm[pl][cl][1] = a2;
  __syncthreads();
  if (pl == 0 && c < C) {
#pragma unroll
    for (int k = 1; k < 8; ++k) {
      a1 += sm[k][cl][0];
      a2 += sm[k][cl][1];
    }
    atomicAdd(&dbeta[c], a1);
    atomicAdd(&dgamma[c], a2);
  }
}

// ------------------------------------------------------------------------------------------------ LayerNorm
static constexpr int LN_MAXV = 5;  // up to 5 x 32 lanes x 8 channels = 1280

__device__ __forceinline__ void unpack8(const uint4& u, float* v) {
  const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
}

// One warp per row, NV = ceil(C / 256) 16-byte vectors per lane (exactly unrolled).  The next row's vectors are
// fetched (raw bf16, 4 registers each) before the current row's reductions, so HBM latency overlaps the shuffles.
template <int NV>
__global__ void __launch_bounds__(256) ln_fwd_kernel(const bf16* __restrict__ x, const float* __restrict__ gamma,
                                                     const float* __restrict__ beta, bf16* __restrict__ y,
                                                     float* __restrict__ stats, long long rows, int C, float eps) {
  pdl_grid_sync();
  const int lane = threadIdx.x & 31;
  const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  const int V = C / 8;
  const float invC = 1.f / (float)C;
  uint4 nx[NV];
  if (warp < rows) {
#pragma unroll
    for (int j = 0; j < NV; ++j)
      if (lane + 32 * j < V) nx[j] = *reinterpret_cast<const uint4*>(x + warp * C + (lane + 32 * j) * 8);
  }
  for (long long row = warp; row < rows; row += nwarps) {
    float f[NV][8];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      if (lane + 32 * j < V) {
        unpack8(nx[j], f[j]);
#pragma unroll
        for (int e = 0; e < 8; ++e) s += f[j][e];
      }
    }
    if (row + nwarps < rows) {
#pragma unroll
      for (int j = 0; j < NV; ++j)
        if (lane + 32 * j < V) nx[j] = *reinterpret_cast<const uint4*>(x + (row + nwarps) * C + (lane + 32 * j) * 8);
    }
    const float mean = warp_sum(s) * invC;
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      if (lane + 32 * j < V) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float d = f[j][e] - mean;
          q += d * d;
        }
      }
    }
    const float rstd = rsqrtf(warp_sum(q) * invC + eps);
    if (lane == 0) *reinterpret_cast<float2*>(stats + row * 2) = make_float2(mean, rstd);
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int v = lane + 32 * j;
      if (v < V) {
        const float4 g0 = *reinterpret_cast<const float4*>(gamma + v * 8), g1 = *reinterpret_cast<const float4*>(gamma + v * 8 + 4);
        const float4 b0 = *reinterpret_cast<const float4*>(beta + v * 8), b1 = *reinterpret_cast<const float4*>(beta + v * 8 + 4);
        const float gv[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        float o[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = (f[j][e] - mean) * rstd * gv[e] + bv[e];
        store8(y + row * C + v * 8, o);
      }
    }
  }
}

// one warp per row; per-block partial (dbeta, dgamma) -> ws[block][C][2]
template <int NV>
__global__ void __launch_bounds__(256, (NV <= 2 ? 2 : 1)) ln_bwd_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x,
                                                                        const float* __restrict__ gamma,
                                                                        const float* __restrict__ stats,
                                                                        const bf16* __restrict__ dx_add, bf16* __restrict__ dx,
                                                                        float* __restrict__ ws, long long rows, int C) {
  pdl_grid_sync();
  extern __shared__ float sm[];  // [C][2]
  const int lane = threadIdx.x & 31;
  const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  const int V = C / 8;
  const float invC = 1.f / (float)C;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
  float ag[NV][8], ab[NV][8];
#pragma unroll
  for (int j = 0; j < NV; ++j)
#pragma unroll
    for (int e = 0; e < 8; ++e) ag[j][e] = ab[j][e] = 0.f;
  uint4 nx[NV], nd[NV], na[NV];
  float2 nst = make_float2(0.f, 0.f);
  if (warp < rows) {
    nst = *reinterpret_cast<const float2*>(stats + warp * 2);
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int v = lane + 32 * j;
      if (v < V) {
        nx[j] = *reinterpret_cast<const uint4*>(x + warp * C + v * 8);
        nd[j] = *reinterpret_cast<const uint4*>(dy + warp * C + v * 8);
        if (dx_add) na[j] = *reinterpret_cast<const uint4*>(dx_add + warp * C + v * 8);
      }
    }
  }
  for (long long row = warp; row < rows; row += nwarps) {
    const float mean = nst.x, rstd = nst.y;
    float xh[NV][8], g[NV][8];
    uint4 ca[NV];
    float c1 = 0.f, c2 = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int v = lane + 32 * j;
      if (v < V) {
        float d[8];
        unpack8(nx[j], xh[j]);
        unpack8(nd[j], d);
        ca[j] = na[j];
        const float4 g0 = *reinterpret_cast<const float4*>(gamma + v * 8), g1 = *reinterpret_cast<const float4*>(gamma + v * 8 + 4);
        const float gv[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          xh[j][e] = (xh[j][e] - mean) * rstd;
          g[j][e] = d[e] * gv[e];
          c1 += g[j][e];
          c2 += g[j][e] * xh[j][e];
          ab[j][e] += d[e];
          ag[j][e] += d[e] * xh[j][e];
        }
      }
    }
    if (row + nwarps < rows) {  // next row's operands: in flight during the reductions below
      const long long nr = row + nwarps;
      nst = *reinterpret_cast<const float2*>(stats + nr * 2);
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const int v = lane + 32 * j;
        if (v < V) {
          nx[j] = *reinterpret_cast<const uint4*>(x + nr * C + v * 8);
          nd[j] = *reinterpret_cast<const uint4*>(dy + nr * C + v * 8);
          if (dx_add) na[j] = *reinterpret_cast<const uint4*>(dx_add + nr * C + v * 8);
        }
      }
    }
    c1 = warp_sum(c1) * invC;
    c2 = warp_sum(c2) * invC;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int v = lane + 32 * j;
      if (v < V) {
        float o[8];
        if (dx_add) unpack8(ca[j], o);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float t = rstd * (g[j][e] - c1 - xh[j][e] * c2);
          o[e] = dx_add ? o[e] + t : t;
        }
        store8(dx + row * C + v * 8, o);
      }
    }
  }
  // block-level sum of the per-warp partials: warps take turns adding into shared memory (each lane owns distinct
  // channels, so the read-modify-writes are conflict-free and need no atomics)
  for (int wturn = 0; wturn < (int)(blockDim.x >> 5); ++wturn) {
    if ((int)(threadIdx.x >> 5) == wturn) {
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const int v = lane + 32 * j;
        if (v < V) {
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            sm[(v * 8 + e) * 2] += ab[j][e];
            sm[(v * 8 + e) * 2 + 1] += ag[j][e];
          }
        }
      }
    }
    __syncthreads();
  }
  float* o = ws + (long long)blockIdx.x * C * 2;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) o[i] = sm[i];
}

static int ln_blocks(long long rows, int num_sms) {
  long long b = (rows + 7) / 8;
  const long long cap = (long long)num_sms * 4;  // sd2_layernorm_ws_floats is sized for 4 blocks per SM
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace sd2

using namespace sd2;

extern "C" {

long long sd2_groupnorm_ws_floats(int B, int C) { return (long long)B * GN_MAXP * C * 2 + (long long)B * 64 * 2; }

int sd2_groupnorm_fwd(sd2_ctx* ctx, const void* x, long long ldx, const float* gamma, const float* beta, void* y,
                      long long ldy, float* stats, float* ws, int B, int HW, int C, int G, float eps, int silu,
                      sd2_stream stream_) {
  if (!ctx) return 1;
  if (C % 8 != 0 || C % G != 0 || G > 64 || C / 8 > GN_THREADS) return fail(ctx, "sd2_groupnorm_fwd: unsupported C/G");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const dim3 grid(gn_chunks(HW, B, C, ctx->num_sms), B);
  const size_t red_smem = ((size_t)(GN_THREADS / (C / 8)) * C * 2 + (size_t)C * 2) * sizeof(float);
  launch_k(gn_stats_kernel, dim3(grid), dim3(GN_THREADS), red_smem, stream, reinterpret_cast<const bf16*>(x), ldx, ws, HW, C, G);
  launch_k(gn_apply_kernel, dim3(grid), dim3(GN_THREADS), 0, stream, reinterpret_cast<const bf16*>(x), ldx, gamma, beta,
                                                   reinterpret_cast<bf16*>(y), ldy, stats, ws, HW, C, G, eps, silu);
  return check_launch(ctx, "groupnorm_fwd", 2);
}

int sd2_groupnorm_bwd(sd2_ctx* ctx, const void* dy, long long lddy, const void* x, long long ldx, const float* gamma,
                      const float* beta, const float* stats, const void* dx_add, long long ldadd, void* dx,
                      long long lddx, float* dgamma, float* dbeta, float* ws, int B, int HW, int C, int G, int silu,
                      sd2_stream stream_) {
  if (!ctx) return 1;
  if (C % 8 != 0 || C % G != 0 || G > 64 || C / 8 > GN_THREADS) return fail(ctx, "sd2_groupnorm_bwd: unsupported C/G");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const int P = gn_chunks(HW, B, C, ctx->num_sms);
  const dim3 grid(P, B);
  const size_t red_smem = ((size_t)(GN_THREADS / (C / 8)) * C * 2 + (size_t)C * 2) * sizeof(float);
  float* gstat = ws + (long long)B * GN_MAXP * C * 2;
  launch_k(gn_bwd_stats_kernel, dim3(grid), dim3(GN_THREADS), red_smem, stream, reinterpret_cast<const bf16*>(dy), lddy, reinterpret_cast<const bf16*>(x), ldx, gamma, beta, stats, ws, HW, C, G, silu);
  const size_t t_smem = (size_t)B * (C / G) * 2 * sizeof(float);
  if (t_smem > 48 * 1024) {  // large per-GPU batches: opt in to more than the default 48 KB of dynamic shared memory
    static size_t opted = 0;
    if (t_smem > 227 * 1024) return fail(ctx, "sd2_groupnorm_bwd: B * C / G too large for the per-group reduction");
    if (t_smem > opted) {
      if (cudaFuncSetAttribute(gn_bwd_reduce_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
        return fail(ctx, "sd2_groupnorm_bwd: cannot raise the shared-memory limit");
      opted = 227 * 1024;
    }
  }
  launch_k(gn_bwd_reduce_kernel, dim3(G), dim3(256), t_smem, stream, ws, gamma, gstat, dgamma, dbeta, B, P, C, G);
  launch_k(gn_bwd_apply_kernel, dim3(grid), dim3(GN_THREADS), 0, stream, reinterpret_cast<const bf16*>(dy), lddy,
                                                       reinterpret_cast<const bf16*>(x), ldx, gamma, beta, stats, gstat,
                                                       reinterpret_cast<const bf16*>(dx_add), ldadd,
                                                       reinterpret_cast<bf16*>(dx), lddx, HW, C, G, silu);
  return check_launch(ctx, "groupnorm_bwd", 3);
}

int sd2_layernorm_fwd(sd2_ctx* ctx, const void* x, const float* gamma, const float* beta, void* y, float* stats,
                      long long rows, int C, float eps, sd2_stream stream_) {
  if (!ctx) return 1;
  if (C % 8 != 0 || C > LN_MAXV * 256) return fail(ctx, "sd2_layernorm_fwd: unsupported C");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const int blocks = grid_for(rows * 32, 256, ctx->num_sms, 8);
  const int nv = (C / 8 + 31) / 32;
#define LN_FWD(NV)                                                                                                        \
  launch_k(ln_fwd_kernel<NV>, dim3(blocks), dim3(256), 0, stream, reinterpret_cast<const bf16*>(x), gamma, beta,           \
           reinterpret_cast<bf16*>(y), stats, rows, C, eps)
  if (nv == 1) LN_FWD(1); else if (nv == 2) LN_FWD(2); else if (nv == 3) LN_FWD(3); else if (nv == 4) LN_FWD(4); else LN_FWD(5);
#undef LN_FWD
  return check_launch(ctx, "layernorm_fwd");
}

long long sd2_layernorm_ws_floats(long long rows, int C) { return (long long)148 * 4 * C * 2; }

int sd2_layernorm_bwd(sd2_ctx* ctx, const void* dy, const void* x, const float* gamma, const float* stats,
                      const void* dx_add, void* dx, float* dgamma, float* dbeta, float* ws, long long rows, int C,
                      sd2_stream stream_) {
  if (!ctx) return 1;
  if (C % 8 != 0 || C > LN_MAXV * 256) return fail(ctx, "sd2_layernorm_bwd: unsupported C");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const int blocks = ln_blocks(rows, ctx->num_sms);
  const int nv = (C / 8 + 31) / 32;
#define LN_BWD(NV)                                                                                                        \
  launch_k(ln_bwd_kernel<NV>, dim3(blocks), dim3(256), 2 * C * sizeof(float), stream, reinterpret_cast<const bf16*>(dy),   \
           reinterpret_cast<const bf16*>(x), gamma, stats, reinterpret_cast<const bf16*>(dx_add),                          \
           reinterpret_cast<bf16*>(dx), ws, rows, C)
  if (nv == 1) LN_BWD(1); else if (nv == 2) LN_BWD(2); else if (nv == 3) LN_BWD(3); else if (nv == 4) LN_BWD(4); else LN_BWD(5);
#undef LN_BWD
  launch_k(affine_grad_reduce_kernel, dim3((C + 31) / 32), dim3(256), 0, stream, ws, blocks, C, dgamma, dbeta);
  return check_launch(ctx, "layernorm_bwd", 2);
}

}  // extern "C"

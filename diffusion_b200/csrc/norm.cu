// GroupNorm(+SiLU) and LayerNorm, forward and backward, NHWC bf16 I/O with fp32 statistics.  HBM-bound kernels:
// 16-byte vector loads (8 bf16 channels per thread), coalesced along the channel dimension, deterministic
// two-level reductions through a caller-provided fp32 scratch (no atomics on global memory).
//
// Replaces ATen group_norm / layer_norm (+ their backward) as run by composer's LPGroupNorm / LPLayerNorm surgery
// (reference diffusion/train.py:91-108) and the F.silu that follows every ResnetBlock2D GroupNorm (diffusers).
#include "common.cuh"
#include "host.h"
#include <cstdlib>

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
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// d/dz [z * sigmoid(z)] = s + z*s*(1-s), sigmoid through one MUFU.TANH
__device__ __forceinline__ float silu_grad_fast(float z) {
  const float s = fmaf(tanh_approx(0.5f * z), 0.5f, 0.5f);
  return fmaf(z, fmaf(-s, s, s), s);
}
__device__ __forceinline__ float silu_grad(float z) { return silu_grad_fast(z); }

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

// grid (P, B): dx = rstd * (gamma*dyh - (db_g + xhat*ds_g)/n) (+ dx_add); cs_ws != null: per-block column sums of dx
// -> cs_ws[b*P+p][C] (dynamic shared memory [R][C] floats then)
__global__ void __launch_bounds__(GN_THREADS, 1) gn_bwd_apply_kernel(const bf16* __restrict__ dy, long long lddy,
                                                                  const bf16* __restrict__ x, long long ldx,
                                                                  const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                  const float* __restrict__ stats, const float* __restrict__ gstat,
                                                                  const bf16* __restrict__ dx_add, long long ldadd,
                                                                  bf16* __restrict__ dx, long long lddx,
                                                                  float* __restrict__ cs_ws, int HW, int C, int G, int silu) {
  pdl_grid_sync();
  extern __shared__ float cs_sm[];  // [R][C] (only with cs_ws)
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
  float cs[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) cs[e] = 0.f;
  if (r < R) {
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
        cs[e] += o;
      }
      store8(ob + (long long)row * lddx, f);
    }
  }
  if (cs_ws != nullptr) {
    if (r < R) {
      float* o = cs_sm + (size_t)r * C + v * 8;
      *reinterpret_cast<float4*>(o) = make_float4(cs[0], cs[1], cs[2], cs[3]);
      *reinterpret_cast<float4*>(o + 4) = make_float4(cs[4], cs[5], cs[6], cs[7]);
    }
    __syncthreads();
    float* wo = cs_ws + ((long long)b * P + p) * C;
    for (int i = threadIdx.x; i < C; i += GN_THREADS) {
      float t = 0.f;
      for (int rr = 0; rr < R; ++rr) t += cs_sm[(size_t)rr * C + i];
      wo[i] = t;
    }
  }
}

// Per-CTA partials of the GroupNorm backward -> their consumers, one launch.  Slab i = b*NP + r (NP partial slabs per image):
//   dbeta[c] += sum_i ws[i][c][0],  dgamma[c] += sum_i ws[i][c][1]                    (affine gradients)
//   rowsum[b][c]  = sum_r cs_ws[b*NP+r][c]   (overwritten; per-image column sums of dx = gradient of a per-image bias)
//   dcs1/dcs2[c] += sum_b rowsum[b][c]       (bias gradients of the conv / linear that produced the norm's input)
// grid (ceil(C/32), slices over images), block (32 channels, 8 image lanes).
__global__ void __launch_bounds__(256) gn_bwd_finalize_kernel(const float* __restrict__ ws, const float* __restrict__ cs_ws, int B,
                                                              int NP, int C, float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                              float* __restrict__ rowsum, float* __restrict__ dcs1,
                                                              float* __restrict__ dcs2) {
  pdl_grid_sync();
  __shared__ float sm[8][32][3];
  const int cl = threadIdx.x & 31, pl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  float tot = 0.f, a1 = 0.f, a2 = 0.f;
  if (c < C) {
    for (int b = blockIdx.y * 8 + pl; b < B; b += 8 * gridDim.y) {
      const float* src = cs_ws + ((long long)b * NP) * C + c;
      const float* aff = ws + (((long long)b * NP) * C + c) * 2;
      float t = 0.f;
      int i = 0;
      for (; i + 3 < NP; i += 4) {
        const float t0 = src[(long long)i * C], t1 = src[(long long)(i + 1) * C], t2 = src[(long long)(i + 2) * C],
                    t3 = src[(long long)(i + 3) * C];
        const float2 g0 = *reinterpret_cast<const float2*>(aff + (long long)i * C * 2),
                     g1 = *reinterpret_cast<const float2*>(aff + (long long)(i + 1) * C * 2),
                     g2 = *reinterpret_cast<const float2*>(aff + (long long)(i + 2) * C * 2),
                     g3 = *reinterpret_cast<const float2*>(aff + (long long)(i + 3) * C * 2);
        t += (t0 + t1) + (t2 + t3);
        a1 += (g0.x + g1.x) + (g2.x + g3.x);
        a2 += (g0.y + g1.y) + (g2.y + g3.y);
      }
      for (; i < NP; ++i) {
        t += src[(long long)i * C];
        const float2 g = *reinterpret_cast<const float2*>(aff + (long long)i * C * 2);
        a1 += g.x;
        a2 += g.y;
      }
      if (rowsum != nullptr) rowsum[(long long)b * C + c] = t;
      tot += t;
    }
  }
  sm[pl][cl][0] = tot;
  sm[pl][cl][1] = a1;
  sm[pl][cl][2] = a2;
  __syncthreads();
  if (pl == 0 && c < C) {
#pragma unroll
    for (int j = 1; j < 8; ++j) {
      tot += sm[j][cl][0];
      a1 += sm[j][cl][1];
      a2 += sm[j][cl][2];
    }
    atomicAdd(&dbeta[c], a1);
    atomicAdd(&dgamma[c], a2);
    if (dcs1 != nullptr) atomicAdd(&dcs1[c], tot);
    if (dcs2 != nullptr) atomicAdd(&dcs2[c], tot);
  }
}

// column sums only (two-pass fallback path: the affine gradients are produced by gn_bwd_reduce_kernel there)
__global__ void __launch_bounds__(256) gn_colsum_only_kernel(const float* __restrict__ cs_ws, int B, int NP, int C,
                                                             float* __restrict__ rowsum, float* __restrict__ dcs1,
                                                             float* __restrict__ dcs2) {
  pdl_grid_sync();
  __shared__ float sm[8][32];
  const int cl = threadIdx.x & 31, pl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  float tot = 0.f;
  if (c < C) {
    for (int b = blockIdx.y * 8 + pl; b < B; b += 8 * gridDim.y) {
      const float* src = cs_ws + ((long long)b * NP) * C + c;
      float t = 0.f;
      for (int i = 0; i < NP; ++i) t += src[(long long)i * C];
      if (rowsum != nullptr) rowsum[(long long)b * C + c] = t;
      tot += t;
    }
  }
  sm[pl][cl] = tot;
  __syncthreads();
  if (pl == 0 && c < C && (dcs1 != nullptr || dcs2 != nullptr)) {
    float t = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) t += sm[j][cl];
    if (dcs1 != nullptr) atomicAdd(&dcs1[c], t);
    if (dcs2 != nullptr) atomicAdd(&dcs2[c], t);
  }
}

// ---- skip concatenation that also takes the GroupNorm statistics of what it writes
// out[row] = [a[row] | b[row]] (the up-path torch.cat of the running tensor and the skip); the kernel streams both halves
// anyway, so it accumulates per column the sum and sum of squares of its pixel chunk: part[(img * P + p)][Ca + Cb][2].
// The norm1 behind it (fused path, slab = HW / P) then needs no statistics pass.  grid (P, B), block GN_THREADS.
__global__ void __launch_bounds__(GN_THREADS, 2) concat_stats_kernel(const bf16* __restrict__ a, long long lda,
                                                                   const bf16* __restrict__ bsrc, long long ldb,
                                                                   bf16* __restrict__ out, float* __restrict__ part, int HW,
                                                                   int Ca, int Cb) {
  pdl_grid_sync();
  extern __shared__ float sm[];  // [R][C][2] + [C][2]
  const int P = gridDim.x, p = blockIdx.x, b = blockIdx.y;
  const int C = Ca + Cb, V = C / 8, Va = Ca / 8, R = GN_THREADS / V;
  const int v = threadIdx.x % V, r = threadIdx.x / V;
  const int row0 = (int)((long long)p * HW / P), row1 = (int)((long long)(p + 1) * HW / P);
  float s[8], q[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) s[e] = q[e] = 0.f;
  if (r < R) {
    const bf16* src = v < Va ? a + ((long long)b * HW) * lda + v * 8 : bsrc + ((long long)b * HW) * ldb + (v - Va) * 8;
    const long long lds = v < Va ? lda : ldb;
    bf16* dst = out + ((long long)b * HW) * C + v * 8;
#pragma unroll 4
    for (int row = row0 + r; row < row1; row += R) {
      const uint4 u = ldg_stream16(src + (long long)row * lds);
      *reinterpret_cast<uint4*>(dst + (long long)row * C) = u;
      const float2 x0 = unpack_bf16x2(u.x), x1 = unpack_bf16x2(u.y), x2 = unpack_bf16x2(u.z), x3 = unpack_bf16x2(u.w);
      const float f[8] = {x0.x, x0.y, x1.x, x1.y, x2.x, x2.y, x3.x, x3.y};
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        s[e] += f[e];
        q[e] = fmaf(f[e], f[e], q[e]);
      }
    }
  }
  gn_block_reduce(sm, s, q, C, R, v, r);
  const float* red = sm + (size_t)R * C * 2;
  float* o = part + (((long long)b * P + p) * C) * 2;
  for (int i = threadIdx.x; i < 2 * C; i += GN_THREADS) o[i] = red[i];
}

// ---- GroupNorm forward on statistics taken by the producing GEMM's epilogue (gemm_tc.cu::epi_gn_stats)
// part[slab][C][2] = (sum, sum of squares) of every column over `slab` consecutive rows; NS = HW / slab slabs per image.
// grid (B, ceil(G / 8)), block 256: one warp per group, lanes over (slab, channel of the group) -> stats[b][g] = (mean, rstd)
__global__ void __launch_bounds__(256) gn_part_finalize_kernel(const float* __restrict__ part, float* __restrict__ stats, int NS,
                                                               int HW, int C, int G, float eps) {
  pdl_grid_sync();
  const int b = blockIdx.x, cpg = C / G;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* pb = part + (long long)b * NS * C * 2;
  const int n = NS * cpg;
  for (int g = blockIdx.y * 8 + warp; g < G; g += 8 * gridDim.y) {
    float s = 0.f, q = 0.f;
    for (int i = lane; i < n; i += 32) {
      const int sl = i / cpg, cl = i % cpg;
      const float2 t = *reinterpret_cast<const float2*>(pb + ((long long)sl * C + g * cpg + cl) * 2);
      s += t.x;
      q += t.y;
    }
    s = warp_sum(s);
    q = warp_sum(q);
    if (lane == 0) {
      const float cnt = (float)cpg * (float)HW;
      const float mean = s / cnt;
      const float var = fmaxf(q / cnt - mean * mean, 0.f);
      *reinterpret_cast<float2*>(stats + ((long long)b * G + g) * 2) = make_float2(mean, rsqrtf(var + eps));
    }
  }
}

// grid (P, B): y = [silu](x * scale + shift) with the per-(image, channel) scale / shift from stats - one streaming pass
__global__ void __launch_bounds__(GN_THREADS, 2) gn_apply_stats_kernel(const bf16* __restrict__ x, const float* __restrict__ gamma,
                                                                     const float* __restrict__ beta, bf16* __restrict__ y,
                                                                     const float* __restrict__ stats, int HW, int C, int G,
                                                                     int silu) {
  pdl_grid_sync();
  const int P = gridDim.x, p = blockIdx.x, b = blockIdx.y;
  const int cpg = C / G;
  const int V = C / 8, R = GN_THREADS / V;
  const int v = threadIdx.x % V, r = threadIdx.x / V;
  if (r >= R) return;
  float sc[8], sh[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int c = v * 8 + e, g = c / cpg;
    const float2 t = *reinterpret_cast<const float2*>(stats + ((long long)b * G + g) * 2);
    sc[e] = gamma[c] * t.y;
    sh[e] = beta[c] - t.x * sc[e];
  }
  const int row0 = (int)((long long)p * HW / P), row1 = (int)((long long)(p + 1) * HW / P);
  const bf16* xb = x + ((long long)b * HW) * C + v * 8;
  bf16* yb = y + ((long long)b * HW) * C + v * 8;
#pragma unroll 4
  for (int row = row0 + r; row < row1; row += R) {
    float f[8];
    const uint4 u = ldg_stream16(xb + (long long)row * C);
    const float2 a = unpack_bf16x2(u.x), bb = unpack_bf16x2(u.y), cc = unpack_bf16x2(u.z), dd = unpack_bf16x2(u.w);
    f[0] = a.x; f[1] = a.y; f[2] = bb.x; f[3] = bb.y; f[4] = cc.x; f[5] = cc.y; f[6] = dd.x; f[7] = dd.y;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float z = fmaf(f[e], sc[e], sh[e]);
      f[e] = silu ? silu_f(z) : z;
    }
    store8(yb + (long long)row * C, f);
  }
}

// ---- single-pass GroupNorm backward: one thread-block CLUSTER per image
// The two-kernel path above reads x and dy twice from HBM/L2 and evaluates the SiLU derivative twice; at the UNet's
// shapes it ran at 1.5-2.5 TB/s and was half instruction-issue bound.  Here a cluster of NC CTAs owns one image: each
// CTA stages its HW/NC pixels of (x, dy) in shared memory with the bulk-copy engine (NSUB sub-tiles, each with its own
// mbarrier, so the first pass starts while later sub-tiles are still in flight),
//   pass 1: dyh = dy * silu'(.) (written back over dy in smem as bf16 - the same rounding point as the reference's
//           silu_backward output), per-channel sums of dyh and dyh*xhat -> block reduce -> ws[b*NC+rank][C][2]
//           (for dgamma/dbeta) and per-group partials,
//   cluster barrier + distributed-shared-memory reads: group sums of the whole image,
//   pass 2: dx = rstd*(gamma*dyh - (db_g + xhat*ds_g)/n) (+dx_add) from the smem copy -> global.
// x and dy cross HBM exactly once.
static constexpr int GNC_NSUB = 4;

// Split second barrier of the cluster kernels: a CTA may exit only after every peer has read its gpart.  Peers read it right
// after the first barrier, so each CTA ARRIVES (relaxed: no memory to publish) as soon as its own remote reads are done and
// WAITS only at the very end - by then everybody has long arrived, and the arrive no longer fences the apply pass's global
// stores (ncu: `membar` + `barrier` were the top stall reasons of both kernels with the full barrier at the end).
__device__ __forceinline__ void cluster_arrive_relaxed() { asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.aligned;" ::: "memory"); }
__device__ __forceinline__ float2 ld_dsmem_f2(const void* local_smem, uint32_t rank) {
  uint32_t ra;
  float2 v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(local_smem)), "r"(rank));
  asm volatile("ld.shared::cluster.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(ra) : "memory");
  return v;
}

struct GnClusterCfg {
  int NC, PX, RL, threads;
  size_t smem_bytes;
  bool ok;
};
// smallest cluster whose per-CTA chunk fits: prefer <= 100 KB (two CTAs per SM), else <= 190 KB; NC = 16 needs the
// non-portable cluster-size opt-in.
// Measured alternatives that LOST at the B=128 shapes (profiles/r01_norm_microbench_B128.md): 16-CTA clusters with 3-4 CTAs
// per SM (61 -> 65..97 us at 131072 x 320), and a persistent one-CTA-per-SM variant with two chunk buffers that prefetches
// the next image under the current apply pass (61 -> 72 us: every iteration waits at the cluster barrier for the slowest
// of its 8 CTAs and no second CTA is resident to fill that wait).  The apply pass with SiLU is MUFU-bound at about the
// HBM time (2 MUFU per element at 16 / clk / SM), so the remaining lever is GN statistics from the producer's epilogue.
static GnClusterCfg gn_cluster_cfg(int HW, int C, int ntensors_staged) {
  GnClusterCfg c;
  c.ok = false;
  const int V = C / 8;
  if (V > 512 || V < 1) return c;
  // pass 0: ~256-thread CTAs with <= 110 KB of shared memory, so that two CTAs (of different images) share an SM and
  // one's loads overlap the other's second pass; pass 1: one 512-thread CTA per SM with up to 225 KB.
  // tuning overrides (tools/norm_bench.py sweeps them): SD2_GN_P0_KB = pass-0 shared-memory limit, SD2_GN_P0_THREADS
  static const int p0_kb = getenv("SD2_GN_P0_KB") ? atoi(getenv("SD2_GN_P0_KB")) : 110;
  static const int p0_threads = getenv("SD2_GN_P0_THREADS") ? atoi(getenv("SD2_GN_P0_THREADS")) : 256;
  static const int p0_maxnc = getenv("SD2_GN_P0_MAXNC") ? atoi(getenv("SD2_GN_P0_MAXNC")) : 8;
  for (int pass = 0; pass < 3 && !c.ok; ++pass) {  // pass 2: 16-CTA (non-portable) clusters, the last resort
    const int target = pass == 0 ? p0_threads : 512;
    c.RL = target / V;
    if (c.RL < 1) c.RL = 1;
    c.threads = ((V * c.RL + 31) / 32) * 32;
    if (c.threads > 512) continue;
    const size_t fixed = (size_t)c.RL * C * 2 * 4 + (size_t)C * 2 * 4 + 64 * 2 * 4 * 2 + GNC_NSUB * 8 + 256;
    const size_t limit = pass == 0 ? (size_t)p0_kb * 1024 : 225 * 1024;
    for (int nc = (pass == 2 ? 16 : 1); nc <= (pass == 2 ? 16 : (pass == 0 ? p0_maxnc : 8)); nc *= 2) {
      if (HW % nc != 0) break;
      const int px = HW / nc;
      const size_t chunk = (size_t)px * C * 2 * ntensors_staged;
      if (chunk + fixed <= limit && px % GNC_NSUB == 0) {
        c.NC = nc;
        c.PX = px;
        c.smem_bytes = chunk + fixed;
        c.ok = true;
        break;
      }
      if (px == 1) break;
    }
  }
  return c;
}

template <bool SILU, bool HAS_ADD>
__global__ void __launch_bounds__(512, 1) gn_bwd_cluster_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x,
                                                                 const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                 const float* __restrict__ stats, const bf16* __restrict__ dx_add,
                                                                 bf16* __restrict__ dx, float* __restrict__ ws,
                                                                 float* __restrict__ cs_ws, int HW, int C, int G, int PX,
                                                                 int RL) {
  // cs_ws != null: also emit this CTA's per-channel column sums of the dx it writes (cs_ws[b*NC+rank][C]) - the bias
  // gradient of the conv / linear that produced x, and per image the gradient of its time-embedding projection
  extern __shared__ __align__(128) uint8_t gsm[];
  const int NC = gridDim.x, rank = blockIdx.x, b = blockIdx.y;  // cluster = the NC blocks of one image
  const int V = C / 8, cpg = C / G;
  const size_t chunk_bytes = (size_t)PX * C * 2;
  bf16* xs = reinterpret_cast<bf16*>(gsm);
  bf16* dsm = reinterpret_cast<bf16*>(gsm + chunk_bytes);
  float* red = reinterpret_cast<float*>(gsm + 2 * chunk_bytes);  // [RL][C][2]
  float* chan = red + (size_t)RL * C * 2;                        // [C][2]
  float2* gpart = reinterpret_cast<float2*>(chan + (size_t)C * 2);  // [64] this CTA's per-group partial (read by peers)
  float2* gtot = gpart + 64;                                     // [64] whole-image (db, ds) / n
  uint64_t* full = reinterpret_cast<uint64_t*>(gtot + 64);       // [GNC_NSUB]
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int i = 0; i < GNC_NSUB; ++i) mbar_init(&full[i], 1);
    fence_mbar_init();
  }
  __syncthreads();
  pdl_grid_sync();
  const long long pix0 = (long long)b * HW + (long long)rank * PX;
  const int sub_px = PX / GNC_NSUB;
  const uint32_t sub_bytes = (uint32_t)sub_px * C * 2;
  if (tid == 0) {
    for (int i = 0; i < GNC_NSUB; ++i) {
      mbar_arrive_expect_tx(&full[i], 2 * sub_bytes);
      bulk_load_1d(reinterpret_cast<uint8_t*>(xs) + (size_t)i * sub_bytes, x + (pix0 + (long long)i * sub_px) * C, sub_bytes, &full[i]);
      bulk_load_1d(reinterpret_cast<uint8_t*>(dsm) + (size_t)i * sub_bytes, dy + (pix0 + (long long)i * sub_px) * C, sub_bytes, &full[i]);
    }
  }
  const int v = tid % V, rl = tid / V;
  const bool act = rl < RL;
  float rs[8], nm[8], ga[8], be[8], s1[8], s2[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    s1[e] = s2[e] = 0.f;
    rs[e] = nm[e] = ga[e] = be[e] = 0.f;
    if (act) {
      const int c = v * 8 + e, g = c / cpg;
      const float mean = stats[((long long)b * G + g) * 2];
      rs[e] = stats[((long long)b * G + g) * 2 + 1];
      nm[e] = -mean * rs[e];
      ga[e] = gamma[c];
      be[e] = beta[c];
    }
  }
  // ---- pass 1
  for (int i = 0; i < GNC_NSUB; ++i) {
    mbar_wait(&full[i], 0);
    if (act) {
#pragma unroll 2
      for (int r = i * sub_px + rl; r < (i + 1) * sub_px; r += RL) {
        float xf[8], df[8];
        load8(xs + (size_t)r * C + v * 8, xf);
        load8(dsm + (size_t)r * C + v * 8, df);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float xh = fmaf(xf[e], rs[e], nm[e]);
          if (SILU) df[e] *= silu_grad_fast(fmaf(xh, ga[e], be[e]));
          s1[e] += df[e];
          s2[e] = fmaf(df[e], xh, s2[e]);
        }
        if (SILU) store8(dsm + (size_t)r * C + v * 8, df);
      }
    }
  }
  // ---- block reduce over the RL row lanes -> chan[c] = (sum dyh, sum dyh*xhat)
  if (act) {
    float* o = red + ((size_t)rl * C + v * 8) * 2;
#pragma unroll
    for (int e = 0; e < 8; e += 2) *reinterpret_cast<float4*>(o + 2 * e) = make_float4(s1[e], s2[e], s1[e + 1], s2[e + 1]);
  }
  __syncthreads();
  {
    float* wo = ws + ((long long)b * NC + rank) * C * 2;
    for (int i = tid; i < 2 * C; i += blockDim.x) {
      float t = 0.f;
      for (int r = 0; r < RL; ++r) t += red[(size_t)r * 2 * C + i];
      chan[i] = t;
      wo[i] = t;
    }
  }
  __syncthreads();
  if (tid < G) {
    float dbv = 0.f, dsv = 0.f;
    for (int c = tid * cpg; c < (tid + 1) * cpg; ++c) {
      const float gm = gamma[c];
      dbv = fmaf(gm, chan[c * 2], dbv);
      dsv = fmaf(gm, chan[c * 2 + 1], dsv);
    }
    gpart[tid] = make_float2(dbv, dsv);
  }
  cluster_sync_all();  // every CTA's gpart is published
  if (tid < G) {
    float dbv = 0.f, dsv = 0.f;
    for (int r = 0; r < NC; ++r) {
      const float2 t = ld_dsmem_f2(&gpart[tid], (uint32_t)r);
      dbv += t.x;
      dsv += t.y;
    }
    const float inv_n = 1.f / ((float)cpg * (float)HW);
    gtot[tid] = make_float2(dbv * inv_n, dsv * inv_n);
  }
  __syncthreads();
  cluster_arrive_relaxed();  // this CTA's remote reads are done (see cluster_wait at the end)
  // ---- pass 2
  if (act) {
    float a[8], k1[8], k2[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int g = (v * 8 + e) / cpg;
      const float2 t = gtot[g];
      a[e] = rs[e] * ga[e];
      k1[e] = -rs[e] * t.x;
      k2[e] = -rs[e] * t.y;
    }
    const bf16* ab = HAS_ADD ? dx_add + pix0 * C + v * 8 : nullptr;
    bf16* ob = dx + pix0 * C + v * 8;
#pragma unroll
    for (int e = 0; e < 8; ++e) s1[e] = 0.f;
#pragma unroll 4
    for (int r = rl; r < PX; r += RL) {
      float xf[8], df[8], o[8];
      if (HAS_ADD) load8(ab + (size_t)r * C, o);
      load8(xs + (size_t)r * C + v * 8, xf);
      load8(dsm + (size_t)r * C + v * 8, df);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float xh = fmaf(xf[e], rs[e], nm[e]);
        float t = fmaf(df[e], a[e], k1[e]);
        t = fmaf(xh, k2[e], t);
        o[e] = HAS_ADD ? o[e] + t : t;
        s1[e] += o[e];
      }
      store8(ob + (size_t)r * C, o);
    }
  }
  if (cs_ws != nullptr) {  // column sums of this CTA's dx rows: row lanes -> red[RL][C] -> cs_ws slab
    if (act) {
      float* o = red + (size_t)rl * C + v * 8;
      *reinterpret_cast<float4*>(o) = make_float4(s1[0], s1[1], s1[2], s1[3]);
      *reinterpret_cast<float4*>(o + 4) = make_float4(s1[4], s1[5], s1[6], s1[7]);
    }
    __syncthreads();
    float* wo = cs_ws + ((long long)b * NC + rank) * C;
    for (int i = tid; i < C; i += blockDim.x) {
      float t = 0.f;
      for (int r = 0; r < RL; ++r) t += red[(size_t)r * C + i];
      wo[i] = t;
    }
  }
  cluster_wait();  // no CTA of the cluster exits while a peer may still read its gpart
}

// ---- single-pass GroupNorm forward on the same cluster skeleton: x crosses HBM once (staged in smem), statistics are
// combined over the cluster through distributed shared memory, y = [silu](x*scale + shift) is written from the smem copy.
template <bool SILU>
__global__ void __launch_bounds__(512, 1) gn_fwd_cluster_kernel(const bf16* __restrict__ x, const float* __restrict__ gamma,
                                                                 const float* __restrict__ beta, bf16* __restrict__ y,
                                                                 float* __restrict__ stats, int HW, int C, int G, int PX, int RL,
                                                                 float eps) {
  extern __shared__ __align__(128) uint8_t gsm[];
  const int NC = gridDim.x, rank = blockIdx.x, b = blockIdx.y;
  const int V = C / 8, cpg = C / G;
  const size_t chunk_bytes = (size_t)PX * C * 2;
  bf16* xs = reinterpret_cast<bf16*>(gsm);
  float* red = reinterpret_cast<float*>(gsm + chunk_bytes);  // [RL][C][2]
  float* chan = red + (size_t)RL * C * 2;                    // [C][2]
  float2* gpart = reinterpret_cast<float2*>(chan + (size_t)C * 2);
  float2* gtot = gpart + 64;                                 // [64] (mean, rstd)
  uint64_t* full = reinterpret_cast<uint64_t*>(gtot + 64);
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int i = 0; i < GNC_NSUB; ++i) mbar_init(&full[i], 1);
    fence_mbar_init();
  }
  __syncthreads();
  pdl_grid_sync();
  const long long pix0 = (long long)b * HW + (long long)rank * PX;
  const int sub_px = PX / GNC_NSUB;
  const uint32_t sub_bytes = (uint32_t)sub_px * C * 2;
  if (tid == 0) {
    for (int i = 0; i < GNC_NSUB; ++i) {
      mbar_arrive_expect_tx(&full[i], sub_bytes);
      bulk_load_1d(reinterpret_cast<uint8_t*>(xs) + (size_t)i * sub_bytes, x + (pix0 + (long long)i * sub_px) * C, sub_bytes, &full[i]);
    }
  }
  const int v = tid % V, rl = tid / V;
  const bool act = rl < RL;
  float s1[8], s2[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) s1[e] = s2[e] = 0.f;
  for (int i = 0; i < GNC_NSUB; ++i) {
    mbar_wait(&full[i], 0);
    if (act) {
#pragma unroll 4
      for (int r = i * sub_px + rl; r < (i + 1) * sub_px; r += RL) {
        float xf[8];
        load8(xs + (size_t)r * C + v * 8, xf);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          s1[e] += xf[e];
          s2[e] = fmaf(xf[e], xf[e], s2[e]);
        }
      }
    }
  }
  if (act) {
    float* o = red + ((size_t)rl * C + v * 8) * 2;
#pragma unroll
    for (int e = 0; e < 8; e += 2) *reinterpret_cast<float4*>(o + 2 * e) = make_float4(s1[e], s2[e], s1[e + 1], s2[e + 1]);
  }
  __syncthreads();
  for (int i = tid; i < 2 * C; i += blockDim.x) {
    float t = 0.f;
    for (int r = 0; r < RL; ++r) t += red[(size_t)r * 2 * C + i];
    chan[i] = t;
  }
  __syncthreads();
  if (tid < G) {
    float a = 0.f, q = 0.f;
    for (int c = tid * cpg; c < (tid + 1) * cpg; ++c) {
      a += chan[c * 2];
      q += chan[c * 2 + 1];
    }
    gpart[tid] = make_float2(a, q);
  }
  cluster_sync_all();
  if (tid < G) {
    float a = 0.f, q = 0.f;
    for (int r = 0; r < NC; ++r) {
      const float2 t = ld_dsmem_f2(&gpart[tid], (uint32_t)r);
      a += t.x;
      q += t.y;
    }
    const float n = (float)cpg * (float)HW;
    const float mean = a / n;
    const float var = fmaxf(q / n - mean * mean, 0.f);
    const float rstd = rsqrtf(var + eps);
    gtot[tid] = make_float2(mean, rstd);
    if (rank == 0) *reinterpret_cast<float2*>(stats + ((long long)b * G + tid) * 2) = make_float2(mean, rstd);
  }
  __syncthreads();
  cluster_arrive_relaxed();
  if (act) {
    float sc[8], sh[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int c = v * 8 + e;
      const float2 t = gtot[c / cpg];
      sc[e] = gamma[c] * t.y;
      sh[e] = beta[c] - t.x * sc[e];
    }
    bf16* ob = y + pix0 * C + v * 8;
#pragma unroll 4
    for (int r = rl; r < PX; r += RL) {
      float xf[8];
      load8(xs + (size_t)r * C + v * 8, xf);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float z = fmaf(xf[e], sc[e], sh[e]);
        xf[e] = SILU ? silu_f(z) : z;
      }
      store8(ob + (size_t)r * C, xf);
    }
  }
  cluster_wait();
}

template <typename... P, typename... A>
static cudaError_t launch_cluster(void (*kern)(P...), dim3 grid, dim3 block, size_t smem, int cluster_x, cudaStream_t stream,
                                  A&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cluster_x;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<P>(std::forward<A>(args))...);
}

template <bool SILU, bool HAS_ADD>
static cudaError_t launch_gn_bwd_cluster(const GnClusterCfg& c, int B, cudaStream_t stream, const bf16* dy, const bf16* x,
                                         const float* gamma, const float* beta, const float* stats, const bf16* dx_add,
                                         bf16* dx, float* ws, float* cs_ws, int HW, int C, int G) {
  auto kern = gn_bwd_cluster_kernel<SILU, HAS_ADD>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  return launch_cluster(kern, dim3(c.NC, B), dim3(c.threads), c.smem_bytes, c.NC, stream, dy, x, gamma, beta, stats, dx_add, dx,
                        ws, cs_ws, HW, C, G, c.PX, c.RL);
}

// dgamma[c] += sum_i ws[i][c][1], dbeta[c] += sum_i ws[i][c][0] over `n_part` partial slabs.
// grid (ceil(C/32), slab slices), block (32 channels, 8 partial lanes): coalesced float2 reads, smem tree over the 8
// lanes, one atomic per (channel, slice).
static inline int affine_slices(int n_part) {  // ~32 slabs per block: 8 partial lanes x 4 loads in flight
  const int s = (n_part + 31) / 32;
  return s < 1 ? 1 : (s > 16 ? 16 : s);
}
__global__ void __launch_bounds__(256) affine_grad_reduce_kernel(const float* __restrict__ ws, int n_part, int C,
                                                                 float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                                 const float* __restrict__ ws_cs = nullptr,
                                                                 float* __restrict__ dcs = nullptr) {
  pdl_grid_sync();
  __shared__ float sm[8][32][2];
  const int cl = threadIdx.x & 31, pl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  float a1 = 0.f, a2 = 0.f;
  if (c < C) {
    // slab i is handled by (grid row i / 8 % gridDim.y, partial lane i % 8); four independent loads in flight per thread
    const int step = 8 * gridDim.y;
    int i = blockIdx.y * 8 + pl;
    for (; i + 3 * step < n_part; i += 4 * step) {
      float2 t[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) t[j] = *reinterpret_cast<const float2*>(ws + ((long long)(i + j * step) * C + c) * 2);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        a1 += t[j].x;
        a2 += t[j].y;
      }
    }
    for (; i < n_part; i += step) {
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
  if (ws_cs != nullptr) {  // third sum: column sums of dx -> bias gradient of the producing linear
    float a3 = 0.f;
    if (c < C)
      for (int i = blockIdx.y * 8 + pl; i < n_part; i += 8 * gridDim.y) a3 += ws_cs[(long long)i * C + c];
    __syncthreads();
    sm[pl][cl][0] = a3;
    __syncthreads();
    if (pl == 0 && c < C) {
#pragma unroll
      for (int k = 1; k < 8; ++k) a3 += sm[k][cl][0];
      atomicAdd(&dcs[c], a3);
    }
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

// ---- LayerNorm backward: persistent CTAs over row tiles staged in shared memory by the bulk-copy engine
// A tile is R consecutive rows = one contiguous byte range of x, dy (and dx_add), so each operand of a tile is ONE
// cp.async.bulk (no tensor map); LN_STAGES tiles are in flight per SM, which is what keeps HBM busy - the previous
// warp-per-row kernel had one row per warp in flight and ran at 1.6 TB/s.  Per tile:
//   phase 1 (warp per row)           : c1 = mean(dy*gamma), c2 = mean(dy*gamma*xhat) -> smem
//   phase 2 (thread per 8 channels)  : dx = rstd*(dy*gamma - c1 - xhat*c2) (+dx_add) -> global; dgamma/dbeta partials in
//                                      16 registers per thread (each thread owns 8 fixed channels, RL row lanes)
// then one block-level reduction of the partials -> ws[block][C][2] (summed by affine_grad_reduce_kernel).
static constexpr int LN_STAGES = 3;


struct LnTileCfg {
  int R, RL, threads, ntens;
  size_t tile_bytes, stage_bytes, body_bytes, smem_bytes;
};
// R = 2 * RL rows per tile: every thread owns 8 channels of exactly two rows, which it keeps unpacked in registers
// between the two phases.
static LnTileCfg ln_tile_cfg(int C, bool has_add) {
  LnTileCfg c;
  const int V = C / 8;
  c.RL = 512 / V;
  if (c.RL > 32) c.RL = 32;
  if (c.RL < 1) c.RL = 1;
  c.R = 2 * c.RL;
  c.threads = ((V * c.RL + 31) / 32) * 32;
  c.ntens = has_add ? 3 : 2;
  c.tile_bytes = (size_t)c.R * C * 2;
  c.stage_bytes = c.tile_bytes * c.ntens + (size_t)c.R * 8;
  const size_t red_bytes = (size_t)c.RL * C * 2 * sizeof(float);
  c.body_bytes = c.stage_bytes * LN_STAGES;
  if (c.body_bytes < red_bytes) c.body_bytes = red_bytes;
  c.body_bytes = (c.body_bytes + 127) & ~(size_t)127;
  // + part[R][V] float2, rowk[R] float2, barriers
  c.smem_bytes = c.body_bytes + (size_t)c.R * V * 8 + (size_t)c.R * 8 + 64;
  return c;
}

// CS: also emit per-block column sums of the dx written (ws_cs[block][C]) - dx (+dx_add) is the total gradient of the
// residual-stream node in front of this norm, i.e. the output gradient of the linear that produced it, so its column sums
// are that linear's bias gradient and the separate column-sum pass over the same tensor disappears.
template <bool HAS_ADD, bool CS>
__global__ void __launch_bounds__(512, 1) ln_bwd_tile_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x,
                                                              const float* __restrict__ gamma, const float* __restrict__ stats,
                                                              const bf16* __restrict__ dx_add, bf16* __restrict__ dx,
                                                              float* __restrict__ ws, float* __restrict__ ws_cs, long long rows,
                                                              int C, int RL, unsigned body_bytes) {
  extern __shared__ __align__(128) uint8_t ln_smem[];
  constexpr int NT = HAS_ADD ? 3 : 2;
  const int V = C / 8, R = 2 * RL;
  const unsigned tile_bytes = (unsigned)R * C * 2;
  const unsigned stage_bytes = tile_bytes * NT + (unsigned)R * 8;
  float2* part = reinterpret_cast<float2*>(ln_smem + body_bytes);        // [R][V] per-thread partial (sum g, sum g*(x-mean))
  float2* rowk = part + (size_t)R * V;                                   // [R] (rstd*c1, rstd*c2)
  uint64_t* full = reinterpret_cast<uint64_t*>(rowk + R);                // [LN_STAGES]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const long long tiles = (rows + R - 1) / R;
  const float invC = 1.f / (float)C;

  if (tid == 0) {
    for (int s = 0; s < LN_STAGES; ++s) mbar_init(&full[s], 1);
    fence_mbar_init();
  }
  __syncthreads();
  pdl_grid_sync();

  auto issue = [&](long long tile, int s) {  // thread 0 only
    const long long row0 = tile * R;
    const int rt = (int)((rows - row0) < R ? (rows - row0) : R);
    const uint32_t tb = (uint32_t)rt * C * 2, sb = (uint32_t)((rt & ~1) * 8);
    uint8_t* st = ln_smem + (size_t)s * stage_bytes;
    mbar_arrive_expect_tx(&full[s], tb * NT + sb);
    bulk_load_1d(st, x + row0 * C, tb, &full[s]);
    bulk_load_1d(st + tile_bytes, dy + row0 * C, tb, &full[s]);
    if (HAS_ADD) bulk_load_1d(st + 2 * tile_bytes, dx_add + row0 * C, tb, &full[s]);
    if (sb) bulk_load_1d(st + NT * tile_bytes, stats + row0 * 2, sb, &full[s]);
  };
  if (tid == 0) {
    for (int s = 0; s < LN_STAGES; ++s) {
      const long long t = blockIdx.x + (long long)s * gridDim.x;
      if (t < tiles) issue(t, s);
    }
  }

  const int v = tid % V, rl = tid / V;
  const bool p2 = rl < RL;
  float gv[8], ag[8], ab[8];
  {
    float4 g0 = make_float4(0.f, 0.f, 0.f, 0.f), g1 = g0;
    if (p2) {
      g0 = *reinterpret_cast<const float4*>(gamma + v * 8);
      g1 = *reinterpret_cast<const float4*>(gamma + v * 8 + 4);
    }
    gv[0] = g0.x; gv[1] = g0.y; gv[2] = g0.z; gv[3] = g0.w; gv[4] = g1.x; gv[5] = g1.y; gv[6] = g1.z; gv[7] = g1.w;
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) ag[e] = ab[e] = 0.f;
  float ac[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) ac[e] = 0.f;

  int s = 0;
  uint32_t ph = 0;
  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const long long row0 = tile * R;
    const int rt = (int)((rows - row0) < R ? (rows - row0) : R);
    const uint8_t* st = ln_smem + (size_t)s * stage_bytes;
    const float2* sts = reinterpret_cast<const float2*>(st + NT * tile_bytes);
    mbar_wait(&full[s], ph);
    // ---- A: every thread unpacks its 8 channels of its two rows once and publishes the row-sum partials
    float xf[2][8], df[2][8];
    float2 mr[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int r = rl + j * RL;
      if (p2 && r < rt) {
        if ((rt & 1) && r == rt - 1) mr[j] = *reinterpret_cast<const float2*>(stats + (row0 + r) * 2);  // odd tail row
        else mr[j] = sts[r];
        load8(reinterpret_cast<const bf16*>(st) + (size_t)r * C + v * 8, xf[j]);
        load8(reinterpret_cast<const bf16*>(st + tile_bytes) + (size_t)r * C + v * 8, df[j]);
        float p1 = 0.f, pq = 0.f;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float g = df[j][e] * gv[e];
          p1 += g;
          pq += g * (xf[j][e] - mr[j].x);
        }
        part[(size_t)r * V + v] = make_float2(p1, pq);
      }
    }
    __syncthreads();
    // ---- B: one warp per row sums the V partials
    for (int r = warp; r < rt; r += nwarps) {
      float a1 = 0.f, a2 = 0.f;
      for (int vv = lane; vv < V; vv += 32) {
        const float2 t = part[(size_t)r * V + vv];
        a1 += t.x;
        a2 += t.y;
      }
      a1 = warp_sum(a1);
      a2 = warp_sum(a2);
      if (lane == 0) {
        float rstd;
        if ((rt & 1) && r == rt - 1) rstd = stats[(row0 + r) * 2 + 1];
        else rstd = sts[r].y;
        rowk[r] = make_float2(rstd * a1 * invC, rstd * rstd * a2 * invC);
      }
    }
    __syncthreads();
    // ---- C: dx = rstd*(dy*gamma) - rstd*c1 - xhat*(rstd*c2) (+dx_add); per-channel partial sums
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int r = rl + j * RL;
      if (p2 && r < rt) {
        const float2 k = rowk[r];
        const float rstd = mr[j].y, nm = -mr[j].x * rstd;
        float o[8];
        if (HAS_ADD) load8(reinterpret_cast<const bf16*>(st + 2 * tile_bytes) + (size_t)r * C + v * 8, o);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float xh = fmaf(xf[j][e], rstd, nm);
          float t = fmaf(df[j][e], gv[e] * rstd, -k.x);
          t = fmaf(xh, -k.y, t);
          o[e] = HAS_ADD ? o[e] + t : t;
          ab[e] += df[j][e];
          ag[e] = fmaf(df[j][e], xh, ag[e]);
          if (CS) ac[e] += o[e];
        }
        store8(dx + (row0 + r) * C + v * 8, o);
      }
    }
    __syncthreads();  // every thread is done with stage s, part and rowk: refill the stage with the tile LN_STAGES ahead
    if (tid == 0) {
      const long long nt = tile + (long long)LN_STAGES * gridDim.x;
      if (nt < tiles) issue(nt, s);
    }
    if (++s == LN_STAGES) {
      s = 0;
      ph ^= 1u;
    }
  }
  // ---- block-level reduction of the partials over the RL row lanes (no copies are in flight any more)
  float* red = reinterpret_cast<float*>(ln_smem);  // [RL][C][2]
  if (p2) {
    float* o = red + ((size_t)rl * C + v * 8) * 2;
#pragma unroll
    for (int e = 0; e < 8; e += 2) *reinterpret_cast<float4*>(o + 2 * e) = make_float4(ab[e], ag[e], ab[e + 1], ag[e + 1]);
  }
  __syncthreads();
  float* o = ws + (long long)blockIdx.x * C * 2;
  for (int i = tid; i < 2 * C; i += blockDim.x) {
    float t = 0.f;
    for (int r = 0; r < RL; ++r) t += red[(size_t)r * 2 * C + i];
    o[i] = t;
  }
  if (CS) {
    __syncthreads();
    if (p2) {
      float* q = red + (size_t)rl * C + v * 8;
      *reinterpret_cast<float4*>(q) = make_float4(ac[0], ac[1], ac[2], ac[3]);
      *reinterpret_cast<float4*>(q + 4) = make_float4(ac[4], ac[5], ac[6], ac[7]);
    }
    __syncthreads();
    float* oc = ws_cs + (long long)blockIdx.x * C;
    for (int i = tid; i < C; i += blockDim.x) {
      float t = 0.f;
      for (int r = 0; r < RL; ++r) t += red[(size_t)r * C + i];
      oc[i] = t;
    }
  }
}

}  // namespace sd2

using namespace sd2;

extern "C" {

// [B][GN_MAXP][C][2] channel partials | [B][64][2] group sums | [B][GN_MAXP][C] column-sum partials (bias gradients)
long long sd2_groupnorm_ws_floats(int B, int C) { return (long long)B * GN_MAXP * C * 3 + (long long)B * 64 * 2; }

int sd2_groupnorm_fwd(sd2_ctx* ctx, const void* x, long long ldx, const float* gamma, const float* beta, void* y,
                      long long ldy, float* stats, float* ws, int B, int HW, int C, int G, float eps, int silu,
                      sd2_stream stream_) {
  if (!ctx) return 1;
  if (C % 8 != 0 || C % G != 0 || G > 64 || C / 8 > GN_THREADS) return fail(ctx, "sd2_groupnorm_fwd: unsupported C/G");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (ldx == C && ldy == C) {
    const GnClusterCfg cc = gn_cluster_cfg(HW, C, 1);
    if (cc.ok) {
      static bool attr_set = false;
      if (!attr_set) {
        if (cudaFuncSetAttribute(gn_fwd_cluster_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess ||
            cudaFuncSetAttribute(gn_fwd_cluster_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess ||
            cudaFuncSetAttribute(gn_fwd_cluster_kernel<true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess ||
            cudaFuncSetAttribute(gn_fwd_cluster_kernel<false>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess)
          return fail(ctx, "sd2_groupnorm_fwd: cannot configure the cluster kernel");
        attr_set = true;
      }
      const cudaError_t e =
          silu ? launch_cluster(gn_fwd_cluster_kernel<true>, dim3(cc.NC, B), dim3(cc.threads), cc.smem_bytes, cc.NC, stream,
                                reinterpret_cast<const bf16*>(x), gamma, beta, reinterpret_cast<bf16*>(y), stats, HW, C, G, cc.PX,
                                cc.RL, eps)
               : launch_cluster(gn_fwd_cluster_kernel<false>, dim3(cc.NC, B), dim3(cc.threads), cc.smem_bytes, cc.NC, stream,
                                reinterpret_cast<const bf16*>(x), gamma, beta, reinterpret_cast<bf16*>(y), stats, HW, C, G, cc.PX,
                                cc.RL, eps);
      if (e != cudaSuccess) return fail(ctx, std::string("sd2_groupnorm_fwd (cluster): ") + cudaGetErrorString(e));
      return check_launch(ctx, "groupnorm_fwd", 1);
    }
  }
  const dim3 grid(gn_chunks(HW, B, C, ctx->num_sms), B);
  const size_t red_smem = ((size_t)(GN_THREADS / (C / 8)) * C * 2 + (size_t)C * 2) * sizeof(float);
  launch_k(gn_stats_kernel, dim3(grid), dim3(GN_THREADS), red_smem, stream, reinterpret_cast<const bf16*>(x), ldx, ws, HW, C, G);
  launch_k(gn_apply_kernel, dim3(grid), dim3(GN_THREADS), 0, stream, reinterpret_cast<const bf16*>(x), ldx, gamma, beta,
                                                   reinterpret_cast<bf16*>(y), ldy, stats, ws, HW, C, G, eps, silu);
  return check_launch(ctx, "groupnorm_fwd", 2);
}

int sd2_concat_stats(sd2_ctx* ctx, const void* a, long long lda, const void* b, long long ldb, void* out, float* part, int B,
                     int HW, int Ca, int Cb, int P, sd2_stream stream_) {
  if (!ctx) return 1;
  const int C = Ca + Cb;
  if (Ca % 8 || Cb % 8 || lda % 8 || ldb % 8 || C / 8 > GN_THREADS) return fail(ctx, "sd2_concat_stats: channel counts / strides % 8");
  if (P < 1 || P > GN_MAXP || HW % P != 0) return fail(ctx, "sd2_concat_stats: P must divide HW (1..64)");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const size_t smem = ((size_t)(GN_THREADS / (C / 8)) * C * 2 + (size_t)C * 2) * sizeof(float);
  if (smem > 48 * 1024) {
    static bool opted = false;
    if (!opted) {
      if (cudaFuncSetAttribute(concat_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024) != cudaSuccess)
        return fail(ctx, "sd2_concat_stats: cannot raise the shared-memory limit");
      opted = true;
    }
  }
  launch_k(concat_stats_kernel, dim3(P, B), dim3(GN_THREADS), smem, stream, reinterpret_cast<const bf16*>(a), lda,
           reinterpret_cast<const bf16*>(b), ldb, reinterpret_cast<bf16*>(out), part, HW, Ca, Cb);
  return check_launch(ctx, "concat_stats");
}

int sd2_groupnorm_fwd_fused(sd2_ctx* ctx, const void* x, const float* gn_partial, int slab, const float* gamma,
                            const float* beta, void* y, float* stats, float* ws, int B, int HW, int C, int G, float eps,
                            int silu, sd2_stream stream_) {
  if (!ctx) return 1;
  (void)ws;
  if (C % 8 != 0 || C % G != 0 || G > 64 || C / 8 > GN_THREADS) return fail(ctx, "sd2_groupnorm_fwd_fused: unsupported C/G");
  if (slab < 1 || HW % slab != 0) return fail(ctx, "sd2_groupnorm_fwd_fused: slab must divide HW");
  if (!gn_partial) return fail(ctx, "sd2_groupnorm_fwd_fused: no partial statistics");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  launch_k(gn_part_finalize_kernel, dim3(B, (G + 7) / 8), dim3(256), 0, stream, gn_partial, stats, HW / slab, HW, C, G, eps);
  // enough blocks for ~4 per SM; every thread keeps >= 4 rows in flight
  const int R = GN_THREADS / (C / 8);
  int P = (4 * ctx->num_sms + B - 1) / B;
  const int pmax = HW / (4 * (R > 0 ? R : 1));
  if (P > pmax) P = pmax;
  if (P > GN_MAXP) P = GN_MAXP;
  if (P < 1) P = 1;
  launch_k(gn_apply_stats_kernel, dim3(P, B), dim3(GN_THREADS), 0, stream, reinterpret_cast<const bf16*>(x), gamma, beta,
           reinterpret_cast<bf16*>(y), (const float*)stats, HW, C, G, silu);
  return check_launch(ctx, "groupnorm_fwd_fused", 2);
}

int sd2_groupnorm_bwd(sd2_ctx* ctx, const void* dy, long long lddy, const void* x, long long ldx, const float* gamma,
                      const float* beta, const float* stats, const void* dx_add, long long ldadd, void* dx,
                      long long lddx, float* dgamma, float* dbeta, float* ws, int B, int HW, int C, int G, int silu,
                      float* drowsum, float* dcolsum1, float* dcolsum2, sd2_stream stream_) {
  if (!ctx) return 1;
  if (C % 8 != 0 || C % G != 0 || G > 64 || C / 8 > GN_THREADS) return fail(ctx, "sd2_groupnorm_bwd: unsupported C/G");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const bool emit = drowsum != nullptr || dcolsum1 != nullptr || dcolsum2 != nullptr;
  float* cs_ws = emit ? ws + (long long)B * GN_MAXP * C * 2 + (long long)B * 64 * 2 : nullptr;
  const dim3 fin_grid((C + 31) / 32, B >= 64 ? 8 : (B >= 8 ? 2 : 1));
  static const int force_two_pass = getenv("SD2_GN_BWD_TWOPASS") ? atoi(getenv("SD2_GN_BWD_TWOPASS")) : 0;  // A/B switch
  if (!force_two_pass && ldx == C && lddy == C && lddx == C && (dx_add == nullptr || ldadd == C)) {
    const GnClusterCfg cc = gn_cluster_cfg(HW, C, 2);
    if (cc.ok) {  // single-pass cluster kernel + the dgamma/dbeta reduction over the B * NC per-CTA slabs
      const bf16* dyp = reinterpret_cast<const bf16*>(dy);
      const bf16* xp = reinterpret_cast<const bf16*>(x);
      const bf16* ap = reinterpret_cast<const bf16*>(dx_add);
      bf16* dxp = reinterpret_cast<bf16*>(dx);
      cudaError_t e;
      if (silu && ap) e = launch_gn_bwd_cluster<true, true>(cc, B, stream, dyp, xp, gamma, beta, stats, ap, dxp, ws, cs_ws, HW, C, G);
      else if (silu) e = launch_gn_bwd_cluster<true, false>(cc, B, stream, dyp, xp, gamma, beta, stats, ap, dxp, ws, cs_ws, HW, C, G);
      else if (ap) e = launch_gn_bwd_cluster<false, true>(cc, B, stream, dyp, xp, gamma, beta, stats, ap, dxp, ws, cs_ws, HW, C, G);
      else e = launch_gn_bwd_cluster<false, false>(cc, B, stream, dyp, xp, gamma, beta, stats, ap, dxp, ws, cs_ws, HW, C, G);
      if (e != cudaSuccess) return fail(ctx, std::string("sd2_groupnorm_bwd (cluster): ") + cudaGetErrorString(e));
      if (emit)  // affine gradients + the column sums of dx, one launch
        launch_k(gn_bwd_finalize_kernel, fin_grid, dim3(256), 0, stream, (const float*)ws, (const float*)cs_ws, B, cc.NC, C, dgamma,
                 dbeta, drowsum, dcolsum1, dcolsum2);
      else
        launch_k(affine_grad_reduce_kernel, dim3((C + 31) / 32, affine_slices(B * cc.NC)), dim3(256), 0, stream, ws, B * cc.NC, C, dgamma, dbeta, nullptr, nullptr);
      return check_launch(ctx, "groupnorm_bwd", 2);
    }
  }
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
  const size_t cs_smem = emit ? (size_t)(GN_THREADS / (C / 8)) * C * sizeof(float) : 0;
  if (cs_smem > 48 * 1024) {
    static bool opted_apply = false;
    if (!opted_apply) {
      if (cudaFuncSetAttribute(gn_bwd_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024) != cudaSuccess)
        return fail(ctx, "sd2_groupnorm_bwd: cannot raise the shared-memory limit of the apply pass");
      opted_apply = true;
    }
  }
  launch_k(gn_bwd_apply_kernel, dim3(grid), dim3(GN_THREADS), cs_smem, stream, reinterpret_cast<const bf16*>(dy), lddy,
                                                       reinterpret_cast<const bf16*>(x), ldx, gamma, beta, stats, gstat,
                                                       reinterpret_cast<const bf16*>(dx_add), ldadd,
                                                       reinterpret_cast<bf16*>(dx), lddx, cs_ws, HW, C, G, silu);
  if (emit)
    launch_k(gn_colsum_only_kernel, fin_grid, dim3(256), 0, stream, (const float*)cs_ws, B, P, C, drowsum, dcolsum1, dcolsum2);
  return check_launch(ctx, "groupnorm_bwd", emit ? 4 : 3);
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

long long sd2_layernorm_ws_floats(long long rows, int C) { return (long long)148 * 4 * C * 2; }  // >= one [C][2] slab per CTA

int sd2_layernorm_bwd(sd2_ctx* ctx, const void* dy, const void* x, const float* gamma, const float* stats,
                      const void* dx_add, void* dx, float* dgamma, float* dbeta, float* dcolsum, float* ws, long long rows,
                      int C, sd2_stream stream_) {
  if (!ctx) return 1;
  if (C % 8 != 0 || C < 8) return fail(ctx, "sd2_layernorm_bwd: unsupported C");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (C / 8 > 512) return fail(ctx, "sd2_layernorm_bwd: C > 4096 unsupported");
  const LnTileCfg cfg = ln_tile_cfg(C, dx_add != nullptr);
  if (cfg.smem_bytes > 227 * 1024) return fail(ctx, "sd2_layernorm_bwd: tile does not fit in shared memory");
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(ln_bwd_tile_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess ||
        cudaFuncSetAttribute(ln_bwd_tile_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess ||
        cudaFuncSetAttribute(ln_bwd_tile_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess ||
        cudaFuncSetAttribute(ln_bwd_tile_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
      return fail(ctx, "sd2_layernorm_bwd: cannot raise the shared-memory limit");
    attr_set = true;
  }
  const long long tiles = (rows + cfg.R - 1) / cfg.R;
  const int blocks = (int)(tiles < ctx->num_sms ? tiles : ctx->num_sms);
  float* ws_cs = ws + (long long)blocks * C * 2;  // the workspace holds 4 x 148 slabs of [C][2]: room for the [C] sums behind them
#define LN_BWD(ADD, CSUM)                                                                                                  \
  launch_k(ln_bwd_tile_kernel<ADD, CSUM>, dim3(blocks), dim3(cfg.threads), cfg.smem_bytes, stream,                          \
           reinterpret_cast<const bf16*>(dy), reinterpret_cast<const bf16*>(x), gamma, stats,                              \
           reinterpret_cast<const bf16*>(dx_add), reinterpret_cast<bf16*>(dx), ws, ws_cs, rows, C, cfg.RL, (unsigned)cfg.body_bytes)
  if (dx_add && dcolsum) LN_BWD(true, true);
  else if (dx_add) LN_BWD(true, false);
  else if (dcolsum) LN_BWD(false, true);
  else LN_BWD(false, false);
#undef LN_BWD
  launch_k(affine_grad_reduce_kernel, dim3((C + 31) / 32, affine_slices(blocks)), dim3(256), 0, stream, ws, blocks, C, dgamma, dbeta,
           dcolsum ? ws_cs : nullptr, dcolsum);
  return check_launch(ctx, "layernorm_bwd", 2);
}

}  // extern "C"

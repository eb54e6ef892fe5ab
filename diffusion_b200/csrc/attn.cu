// Fused flash-style attention, forward and backward, head_dim 64, non-causal, for sm_100a (tcgen05 / TMEM / TMA).
//
// Replaces the xformers / SDPA memory-efficient attention diffusers' `Attention` calls inside every
// BasicTransformerBlock (self-attention over HW tokens and cross-attention over the 77 x 1024 CLIP context;
// reference diffusion/models/models.py:109-111 enables it, the call site is the UNet forward at
// diffusion/models/stable_diffusion.py:183) and its backward.  Scores never touch HBM.
//
// Layout: q/k/v/o/do are column slices of NHWC-flattened activations: token row (b*N + i), head h at columns
// h*64 .. h*64+63, row stride ld (elements).  lse is the base-2 log-sum-exp of the scaled scores, [B*heads][Nq] fp32.
//
// Forward  (CTA = one 128-query tile of one (b, h); 2 CTAs per SM):
//   warp 0 TMA: Q once, K/V tiles (128 keys) through a 2-stage ring
//   warp 1 MMA: S = Q K^T (M128 N128 K64) into TMEM; O_j = P_j V_j (M128 N64 K128, V read MN-major) into TMEM
//   warps 2-5 : one thread per query row: two passes over S in TMEM (row max, then exp2), P (bf16) -> swizzled smem
//               as the A operand of the second GEMM; running (m, l) and the fp32 output row stay in registers
// Backward (CTA = one 128-key tile of one (b, h), loops over the query tiles; 1 CTA per SM):
//   S^T = K Q^T, dP^T = V dO^T (TMEM) -> P^T = exp2(S^T c - lse), dS^T = P^T (dP^T - D) scale (bf16, smem)
//   dV += P^T dO, dK += dS^T Q (accumulate in TMEM over the loop), dQ_i = dS K (TMEM -> smem -> TMA reduce-add fp32)
//   The same smem image of a [128 x 64] tile serves as K-major operand of one GEMM and MN-major operand of another.
#include <algorithm>

#include "common.cuh"
#include "host.h"

namespace sd2 {

static constexpr int AT_TILE = 128 * 64 * 2;  // one [128 rows][64 bf16] tile, SWIZZLE_128B: 16 KB

struct AttnParams {
  int B, heads, Nq, Nk;
  float c2;     // scale * log2(e)
  float scale;
  bf16* o;
  long long ldo;
  float* lse;         // [B*heads][Nq]
  const float* stats;  // backward: [B*heads][query tiles][lse 128 | rowsum(dO * O) 128], see attn_bwd_prep_kernel
  bf16* dq;            // backward, single key tile: dQ is written directly
  long long lddq;
  float* dq32;         // backward, several key tiles: fp32 dQ accumulator [B*Nq][heads*64]
  int unordered;       // measurement only (SD2_ATTN_UNORDERED=1): skip the waits that fix the dQ summation order
  long long* dq64;     // backward, FX variant: 64-bit fixed-point dQ accumulator [B*Nq][heads*64]
  bf16* dk;
  long long lddk;
  bf16* dv;
  long long lddv;
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ================================================================================================ forward
static constexpr int AT_FWD_THREADS = 320;  // warp 0 TMA, warp 1 MMA, warps 2..9 softmax (two threads per query row)

// exp2(S c2 - m c2) of one 32-column chunk -> bf16 pairs -> 16 columns of the P operand in tensor memory (row = this
// thread's TMEM lane); returns the fp32 row sum of the chunk
template <bool MASKED>
__device__ __forceinline__ float attn_exp_chunk(const uint32_t* rs, float c2, float mc, int col0, int nvalid, uint32_t tP) {
  float sum = 0.f;
  uint32_t pk[16];
#pragma unroll
  for (int e = 0; e < 32; e += 2) {
    float x0 = ex2(fmaf(__uint_as_float(rs[e]), c2, -mc));
    float x1 = ex2(fmaf(__uint_as_float(rs[e + 1]), c2, -mc));
    if (MASKED && col0 + e >= nvalid) x0 = 0.f;
    if (MASKED && col0 + e + 1 >= nvalid) x1 = 0.f;
    sum += x0 + x1;
    pk[e >> 1] = pack_bf16x2(x0, x1);
  }
  tmem_st_32x32b_x16(tP, pk);
  return sum;
}

__global__ void __launch_bounds__(AT_FWD_THREADS, 2)
    attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + AT_TILE;      // [2] K tiles (ring)
  uint8_t* sV = sK + 2 * AT_TILE;  // one V tile (consumed late in a step, so a single buffer suffices)
  float* xch = reinterpret_cast<float*>(sV + AT_TILE);  // [2 parities][2 halves][128 rows] row-max / row-sum exchange
  uint64_t* bars = reinterpret_cast<uint64_t*>(xch + 512);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;   // [2]
  uint64_t* k_empty = bars + 3;  // [2]
  uint64_t* v_full = bars + 5;
  uint64_t* v_empty = bars + 6;
  uint64_t* s_full = bars + 7;
  uint64_t* s_empty = bars + 8;
  uint64_t* p_full = bars + 9;    // [2]
  uint64_t* p_empty = bars + 11;  // [2]
  uint64_t* o_full = bars + 13;
  uint64_t* o_empty = bars + 14;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 15);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int nkt = (p.Nk + 127) / 128;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == 1) {
    if (lane == 0) {
      mbar_init(q_full, 1);
      for (int s = 0; s < 2; ++s) {
        mbar_init(&k_full[s], 1);
        mbar_init(&k_empty[s], 1);
        mbar_init(&p_full[s], 4);
        mbar_init(&p_empty[s], 1);
      }
      mbar_init(v_full, 1);
      mbar_init(v_empty, 1);
      mbar_init(s_full, 1);
      mbar_init(s_empty, 8);
      mbar_init(o_full, 1);
      mbar_init(o_empty, 8);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // S: 128 fp32 columns; O: 64; P (bf16, two keys per column, A operand of the second GEMM straight from tensor memory): 64
  const uint32_t tS = tmem_base, tO = tmem_base + 128, tP = tmem_base + 192;
  pdl_grid_sync();

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer
    if (elect_one()) {
      mbar_arrive_expect_tx(q_full, AT_TILE);
      tma_load_4d(sQ, &tmQ, q_full, 0, qt * 128, h, b);
    }
    __syncwarp();
    for (int j = 0; j < nkt; ++j) {
      const int s = j & 1;
      mbar_wait(&k_empty[s], (uint32_t)((j >> 1) & 1) ^ 1u);
      if (elect_one()) {
        mbar_arrive_expect_tx(&k_full[s], AT_TILE);
        tma_load_4d(sK + s * AT_TILE, &tmK, &k_full[s], 0, j * 128, h, b);
      }
      __syncwarp();
      mbar_wait(v_empty, (uint32_t)(j & 1) ^ 1u);
      if (elect_one()) {
        mbar_arrive_expect_tx(v_full, AT_TILE);
        tma_load_4d(sV, &tmV, v_full, 0, j * 128, h, b);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);
    constexpr uint32_t idesc_o = umma_idesc_bf16(128, 64, 0, 1);
    const uint64_t dQ0 = umma_desc_sw128(smem_u32(sQ), 16, 1024);
    const uint64_t dK0 = umma_desc_sw128(smem_u32(sK), 16, 1024);   // K-major
    const uint64_t dV0 = umma_desc_sw128(smem_u32(sV), 8192, 1024); // MN-major
    constexpr uint64_t TS = AT_TILE >> 4;
    mbar_wait(q_full, 0);
    mbar_wait(&k_full[0], 0);
    tc_fence_after();
    if (elect_one()) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) tc_mma_bf16(tS, dQ0 + 2 * ks, dK0 + 2 * ks, idesc_s, ks != 0 ? 1u : 0u);
      tc_commit(s_full);
      tc_commit(&k_empty[0]);
    }
    __syncwarp();
    for (int j = 0; j < nkt; ++j) {
      if (j + 1 < nkt) {  // next score tile as soon as S_j has been read: runs under the exp work of tile j
        mbar_wait(s_empty, (uint32_t)(j & 1));
        mbar_wait(&k_full[(j + 1) & 1], (uint32_t)(((j + 1) >> 1) & 1));
        tc_fence_after();
        if (elect_one()) {
          const uint64_t dKs = dK0 + (uint64_t)((j + 1) & 1) * TS;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) tc_mma_bf16(tS, dQ0 + 2 * ks, dKs + 2 * ks, idesc_s, ks != 0 ? 1u : 0u);
          tc_commit(s_full);
          tc_commit(&k_empty[(j + 1) & 1]);
        }
        __syncwarp();
      }
      mbar_wait(v_full, (uint32_t)(j & 1));
      if (j > 0) mbar_wait(o_empty, (uint32_t)((j - 1) & 1));  // softmax warps have read O_{j-1}
#pragma unroll 1
      for (int hlf = 0; hlf < 2; ++hlf) {
        mbar_wait(&p_full[hlf], (uint32_t)(j & 1));
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            tc_mma_bf16_ts(tO, tP + (uint32_t)((hlf * 4 + ks) * 8), dV0 + (uint64_t)(hlf * 4 + ks) * 128, idesc_o,
                           (hlf | ks) != 0 ? 1u : 0u);
          tc_commit(&p_empty[hlf]);
          if (hlf == 1) {
            tc_commit(o_full);
            tc_commit(v_empty);
          }
        }
        __syncwarp();
      }
    }
  } else {
    // ---------------------------------------------------------------- softmax + output
    // Two threads per query row: warp group hh = 0 (warps 2-5) owns S columns 0-63 and O columns 0-31,
    // hh = 1 (warps 6-9) the other halves; the row maximum (per tile) and the row sum (once) are exchanged via smem.
    const int q = warp & 3;
    const int hh = (warp - 2) >> 2;
    const int r = q * 32 + lane;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const uint32_t tS_h = tS + lane_off + (uint32_t)(hh * 64);
    const uint32_t tP_h = tP + lane_off + (uint32_t)(hh * 32);
    const uint32_t xch_s = smem_u32(xch);
    float m = -INFINITY, l = 0.f, a_pend = 0.f;
    float o[32];
#pragma unroll
    for (int e = 0; e < 32; ++e) o[e] = 0.f;
    for (int j = 0; j < nkt; ++j) {
      const int nvalid = p.Nk - j * 128;  // columns >= nvalid are padding
      const bool masked = nvalid < 128;
      mbar_wait(s_full, (uint32_t)(j & 1));
      tc_fence_after();
      // pass 1: maximum over this thread's 64 columns
      float mx = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        uint32_t rs[32];
        tmem_ld_32x32b_x32(tS_h + (uint32_t)(c * 32), rs);
        tmem_wait_ld();
        if (!masked) {
#pragma unroll
          for (int e = 0; e < 32; ++e) mx = fmaxf(mx, __uint_as_float(rs[e]));
        } else {
#pragma unroll
          for (int e = 0; e < 32; ++e)
            if (hh * 64 + c * 32 + e < nvalid) mx = fmaxf(mx, __uint_as_float(rs[e]));
        }
      }
      const uint32_t xb = xch_s + (uint32_t)(j & 1) * 1024;
      sts32f(xb + (hh * 128 + r) * 4, mx);
      named_bar_sync(1, 256);
      const float m_new = fmaxf(m, fmaxf(mx, lds32f(xb + ((hh ^ 1) * 128 + r) * 4)));
      if (j > 0) {  // O += P_{j-1} V_{j-1}: deferred to here so that the tensor pipe's latency is hidden behind pass 1
        mbar_wait(o_full, (uint32_t)((j - 1) & 1));
        tc_fence_after();
        uint32_t ro[32];
        tmem_ld_32x32b_x32(tO + lane_off + (uint32_t)(hh * 32), ro);
        tmem_wait_ld();
#pragma unroll
        for (int e = 0; e < 32; ++e) o[e] = fmaf(o[e], a_pend, __uint_as_float(ro[e]));
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(o_empty);
      }
      const float a = ex2((m - m_new) * p.c2);
      const float mc = m_new * p.c2;
      m = m_new;
      float rowsum = 0.f;
      // pass 2: P = exp2(S c2 - m c2) -> bf16 -> this group's half of the P operand
      mbar_wait(&p_empty[hh], (uint32_t)(j & 1) ^ 1u);  // half consumed by the previous tile's P V
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        uint32_t rs[32];
        tmem_ld_32x32b_x32(tS_h + (uint32_t)(c * 32), rs);
        tmem_wait_ld();
        if (c == 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(s_empty);
        }
        if (!masked)
          rowsum += attn_exp_chunk<false>(rs, p.c2, mc, 0, 0, tP_h + (uint32_t)(c * 16));
        else
          rowsum += attn_exp_chunk<true>(rs, p.c2, mc, hh * 64 + c * 32, nvalid, tP_h + (uint32_t)(c * 16));
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[hh]);
      l = l * a + rowsum;
      a_pend = a;
    }
    if (nkt > 0) {  // last tile's P V
      mbar_wait(o_full, (uint32_t)((nkt - 1) & 1));
      tc_fence_after();
      uint32_t ro[32];
      tmem_ld_32x32b_x32(tO + lane_off + (uint32_t)(hh * 32), ro);
      tmem_wait_ld();
#pragma unroll
      for (int e = 0; e < 32; ++e) o[e] = fmaf(o[e], a_pend, __uint_as_float(ro[e]));
    }
    // total row sum = both halves
    const uint32_t xb = xch_s + (uint32_t)(nkt & 1) * 1024;
    sts32f(xb + (hh * 128 + r) * 4, l);
    named_bar_sync(1, 256);
    l += lds32f(xb + ((hh ^ 1) * 128 + r) * 4);
    const int row = qt * 128 + r;
    if (row < p.Nq) {
      const float inv = 1.f / l;
      bf16* dst = p.o + ((long long)b * p.Nq + row) * p.ldo + h * 64 + hh * 32;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint4 u;
        u.x = pack_bf16x2(o[g * 8 + 0] * inv, o[g * 8 + 1] * inv); u.y = pack_bf16x2(o[g * 8 + 2] * inv, o[g * 8 + 3] * inv);
        u.z = pack_bf16x2(o[g * 8 + 4] * inv, o[g * 8 + 5] * inv); u.w = pack_bf16x2(o[g * 8 + 6] * inv, o[g * 8 + 7] * inv);
        *reinterpret_cast<uint4*>(dst + g * 8) = u;
      }
      if (hh == 0) p.lse[((long long)b * p.heads + h) * p.Nq + row] = m * p.c2 + __log2f(l);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 256);
}

// ================================================================================================ backward
// Pre-pass of the backward: D[bh][i] = sum_d O[i][d] dO[i][d] and a copy of the forward's log-sum-exp, both into a
// buffer padded to whole 128-query tiles - st[bh][tile][0][128] = lse (+inf for rows >= Nq, which makes P vanish there),
// st[bh][tile][1][128] = D (0 for padding) - so that the main kernel fetches one contiguous 1 KB block per query tile
// with a bulk copy.  8 lanes per (row, head): 16-byte loads, shuffle reduce.
__global__ void attn_bwd_prep_kernel(const bf16* __restrict__ o, long long ldo, const bf16* __restrict__ d_o, long long lddo,
                                     const float* __restrict__ lse, float* __restrict__ st, int B, int heads, int Nq, int nqt) {
  pdl_grid_sync();
  const int Np = nqt * 128;
  const long long total = (long long)B * Np * heads * 8;
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  float acc = 0.f;
  const long long rowh = i >> 3;
  const int sub = (int)(i & 7);
  const bool ok = i < total;
  long long rowp = 0;
  int h = 0, qi = 0;
  long long bb = 0;
  bool real = false;
  if (ok) {
    rowp = rowh / heads;
    h = (int)(rowh % heads);
    bb = rowp / Np;
    qi = (int)(rowp % Np);
    real = qi < Nq;
    if (real) {
      const long long row = bb * Nq + qi;
      const uint4 a = *reinterpret_cast<const uint4*>(o + row * ldo + h * 64 + sub * 8);
      const uint4 g = *reinterpret_cast<const uint4*>(d_o + row * lddo + h * 64 + sub * 8);
      const float2 a0 = unpack_bf16x2(a.x), a1 = unpack_bf16x2(a.y), a2 = unpack_bf16x2(a.z), a3 = unpack_bf16x2(a.w);
      const float2 g0 = unpack_bf16x2(g.x), g1 = unpack_bf16x2(g.y), g2 = unpack_bf16x2(g.z), g3 = unpack_bf16x2(g.w);
      acc = a0.x * g0.x + a0.y * g0.y + a1.x * g1.x + a1.y * g1.y + a2.x * g2.x + a2.y * g2.y + a3.x * g3.x + a3.y * g3.y;
    }
  }
  acc += __shfl_xor_sync(0xffffffffu, acc, 1);
  acc += __shfl_xor_sync(0xffffffffu, acc, 2);
  acc += __shfl_xor_sync(0xffffffffu, acc, 4);
  if (ok && sub == 0) {
    const long long bh = bb * heads + h;
    float* dst = st + ((bh * nqt + (qi >> 7)) * 2) * 128 + (qi & 127);
    dst[0] = real ? lse[bh * Nq + qi] : INFINITY;
    dst[128] = real ? acc : 0.f;
  }
}

// dst[r][0..cols) bf16 (row stride ldd) = sum over `nparts` fp32 partials (dense [rows][cols] each, `part_stride` floats
// apart, added in index order)
__global__ void cast2d_f32_bf16_kernel(const float* __restrict__ src, bf16* __restrict__ dst, long long ldd, long long rows,
                                       int cols, int nparts, long long part_stride) {
  pdl_grid_sync();
  const int V = cols / 8;
  const long long n = rows * V;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / V;
    const int v = (int)(i % V);
    float4 a = *reinterpret_cast<const float4*>(src + r * cols + v * 8);
    float4 b = *reinterpret_cast<const float4*>(src + r * cols + v * 8 + 4);
    for (int pi = 1; pi < nparts; ++pi) {
      const float4 a2 = *reinterpret_cast<const float4*>(src + pi * part_stride + r * cols + v * 8);
      const float4 b2 = *reinterpret_cast<const float4*>(src + pi * part_stride + r * cols + v * 8 + 4);
      a.x += a2.x; a.y += a2.y; a.z += a2.z; a.w += a2.w;
      b.x += b2.x; b.y += b2.y; b.z += b2.z; b.w += b2.w;
    }
    uint4 u;
    u.x = pack_bf16x2(a.x, a.y); u.y = pack_bf16x2(a.z, a.w); u.z = pack_bf16x2(b.x, b.y); u.w = pack_bf16x2(b.z, b.w);
    *reinterpret_cast<uint4*>(dst + r * ldd + v * 8) = u;
  }
}

// dQ accumulates in 64-bit fixed point (2^-40 units): integer sums do not depend on the order in which the key tiles' CTAs
// add their contributions, so the result is bit-reproducible although the reduce-adds land in arrival order.
static constexpr float AT_FX_SCALE = 1099511627776.f;       // 2^40
static constexpr float AT_FX_INV = 1.f / 1099511627776.f;   // |dQ| < 2^23, resolution 9e-13

// dst[r][0..cols) bf16 (row stride ldd) = src[r][0..cols) fixed point (dense)
__global__ void cast2d_fx64_bf16_kernel(const long long* __restrict__ src, bf16* __restrict__ dst, long long ldd, long long rows,
                                        int cols) {
  pdl_grid_sync();
  const int V = cols / 8;
  const long long n = rows * V;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / V;
    const int v = (int)(i % V);
    const longlong2* p = reinterpret_cast<const longlong2*>(src + r * cols + v * 8);
    const longlong2 a = p[0], b = p[1], c = p[2], d = p[3];
    uint4 u;
    u.x = pack_bf16x2((float)a.x * AT_FX_INV, (float)a.y * AT_FX_INV);
    u.y = pack_bf16x2((float)b.x * AT_FX_INV, (float)b.y * AT_FX_INV);
    u.z = pack_bf16x2((float)c.x * AT_FX_INV, (float)c.y * AT_FX_INV);
    u.w = pack_bf16x2((float)d.x * AT_FX_INV, (float)d.y * AT_FX_INV);
    *reinterpret_cast<uint4*>(dst + r * ldd + v * 8) = u;
  }
}

static constexpr int AT_BWD_CWARPS = 16;                       // compute warps: four threads per key row
static constexpr int AT_BWD_THREADS = 64 + AT_BWD_CWARPS * 32;  // warp 0 TMA, warp 1 MMA, warps 2..17 compute

// dQ tile of one query block, this thread's query row x 16 columns.  Several key tiles (DIRECT = false): TMEM -> fp32
// staging (one [32 rows][16 fp32] box per warp, SWIZZLE_64B) -> TMA reduce-add into the fp32 accumulator: the copy engine
// sends whole 64-byte row segments to the L2 reduction units.  A single key tile (cross-attention over 77 tokens,
// self-attention over <= 128 tokens): the tile is complete, it goes to the bf16 gradient directly (no accumulator, no
// memset, no cast pass).
template <bool DIRECT, bool FX>
__device__ __forceinline__ void attn_bwd_drain_dq(uint32_t taddr, uint32_t my_stg, const void* my_stg_g, int lane, float scale,
                                                  uint64_t* dq_empty, const CUtensorMap* tmDQ, int col0, int row0, int Nq, int h,
                                                  int b, bf16* dq, long long lddq, bool one_q_tile, bool first_kt, int flags) {
  uint32_t rq[16];
  tmem_ld_32x32b_x16(taddr, rq);
  tmem_wait_ld();
  tc_fence_before();
  __syncwarp();
  if (DIRECT) {
    if (lane == 0) mbar_arrive(dq_empty);
    const int row = row0 + lane;
    if (row < Nq) {
      bf16* dst = dq + ((long long)b * Nq + row) * lddq + h * 64 + col0;
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        uint4 u;
        u.x = pack_bf16x2(__uint_as_float(rq[g * 8 + 0]) * scale, __uint_as_float(rq[g * 8 + 1]) * scale);
        u.y = pack_bf16x2(__uint_as_float(rq[g * 8 + 2]) * scale, __uint_as_float(rq[g * 8 + 3]) * scale);
        u.z = pack_bf16x2(__uint_as_float(rq[g * 8 + 4]) * scale, __uint_as_float(rq[g * 8 + 5]) * scale);
        u.w = pack_bf16x2(__uint_as_float(rq[g * 8 + 6]) * scale, __uint_as_float(rq[g * 8 + 7]) * scale);
        *reinterpret_cast<uint4*>(dst + g * 8) = u;
      }
    }
  } else if (FX) {
    // Shared accumulator in 64-bit fixed point: staging = one [32 rows][16 x 8 B] box per warp (128-byte rows, SWIZZLE_128B),
    // integer TMA reduce-add; no ordering needed (integer sums are order-independent)
    if (lane == 0) {
      mbar_arrive(dq_empty);
      bulk_wait_read<0>();  // the staging buffer was handed to the copy engine one iteration ago
    }
    __syncwarp();
    const float fx = scale * AT_FX_SCALE;
    const uint32_t bufp = my_stg + lane * 128;
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      const long long a = __float2ll_rn(__uint_as_float(rq[2 * g]) * fx), c = __float2ll_rn(__uint_as_float(rq[2 * g + 1]) * fx);
      uint4 u;
      u.x = (uint32_t)(unsigned long long)a; u.y = (uint32_t)((unsigned long long)a >> 32);
      u.z = (uint32_t)(unsigned long long)c; u.w = (uint32_t)((unsigned long long)c >> 32);
      sts128(bufp + (uint32_t)((g ^ (lane & 7)) << 4), u);
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      if (row0 < Nq) tma_reduce_add_4d(tmDQ, my_stg_g, col0, row0, h, b);
      bulk_commit();
    }
  } else {
    if (lane == 0) {
      mbar_arrive(dq_empty);
      bulk_wait_read<0>();  // the staging buffer was handed to the copy engine one iteration ago
      // Fixed summation order: the previous contribution to THESE dQ rows (same query tile, previous key tile = nqt tiles
      // ago) must have landed before this one is issued.  With several query tiles only the most recent reduce-add (other
      // rows) may still be in flight.
      if (!(flags & 1)) {
        if (one_q_tile) bulk_wait<0>();
        else bulk_wait<1>();
      }
    }
    __syncwarp();
    const uint32_t bufp = my_stg + lane * 64;  // 64 B rows, SWIZZLE_64B: 16-byte chunk index ^= (row >> 1) & 3
#pragma unroll
    for (int g = 0; g < 4; ++g)
      sts128f(bufp + (uint32_t)((g ^ ((lane >> 1) & 3)) << 4), __uint_as_float(rq[g * 4]) * scale,
              __uint_as_float(rq[g * 4 + 1]) * scale, __uint_as_float(rq[g * 4 + 2]) * scale, __uint_as_float(rq[g * 4 + 3]) * scale);
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      if (row0 < Nq) {  // the first key tile of this CTA's range initialises its partial (no memset), the others add to it
        if (first_kt) tma_store_4d(tmDQ, my_stg_g, col0, row0, h, b);
        else tma_reduce_add_4d(tmDQ, my_stg_g, col0, row0, h, b);
      }
      bulk_commit();
    }
  }
}

// One CTA owns the key tiles [kt0, kt1) of one (image, head) and walks them sequentially (outer loop), each against all
// query tiles (inner loop): dK / dV of a key tile accumulate in tensor memory over the inner loop, and every dQ tile
// receives this CTA's contributions in a FIXED order (same thread, same address, one reduce-add after the other) - the
// gradient is bit-reproducible run to run.  gridDim.x > 1 splits the key range over several CTAs when (images x heads) alone
// cannot fill the GPU; each split accumulates into its own fp32 partial and the cast pass adds the partials in order.
// FX = true (long sequences): one key tile per CTA (gridDim.x = key tiles), all key-tile CTAs of an (image, head) add into ONE
// dQ accumulator in 64-bit fixed point - integer sums are order-independent, so the result is still bit-reproducible, and
// the CTAs of an (image, head) are adjacent in launch order, so their reduce-adds meet in L2 (the sequential walk re-touches a
// 1 MB fp32 dQ slice per key tile from each of 148 CTAs: 148 MB live at N = 4096, past the 126 MB L2).
template <bool DIRECT, bool FX>
__global__ void __launch_bounds__(AT_BWD_THREADS, 1)
    attn_bwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO,
                    const __grid_constant__ CUtensorMap tmDQ, const AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];  // SWIZZLE_128B tiles need 1024-byte alignment (checked below)
  constexpr int KVB = FX ? 1 : 2;              // K / V buffers (FX: one key tile per CTA, nothing to prefetch)
  constexpr int STG_W = FX ? 4096 : 2048;      // dQ staging bytes per compute warp (FX: 16 x int64 per row)
  uint8_t* sKV = smem;                         // [KVB buffers][K tile | V tile]
  uint8_t* sQ = sKV + 2 * KVB * AT_TILE;       // [2]
  uint8_t* sdO = sQ + 2 * AT_TILE;   // [2]
  uint8_t* sdS = sdO + 2 * AT_TILE;  // dS^T : [2 buffers] x 2 query chunks x [128 key rows][128 B]
  uint8_t* stg = sdS + 4 * AT_TILE;  // per compute warp: staging for the dQ reduce-add (DIRECT: unused)
  float* sStat = reinterpret_cast<float*>(stg + AT_BWD_CWARPS * STG_W);  // [2 stages][lse 128 | D 128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sStat + 512);
  uint64_t* kv_full = bars;        // [2]
  uint64_t* kv_empty = bars + 2;   // [2]
  uint64_t* qdo_full = bars + 4;   // [2]
  uint64_t* qdo_empty = bars + 6;  // [2]
  uint64_t* s_full = bars + 8;
  uint64_t* s_empty = bars + 9;
  uint64_t* dp_full = bars + 10;
  uint64_t* dp_empty = bars + 11;
  uint64_t* pds_full = bars + 12;
  uint64_t* pt_empty = bars + 13;   // P^T (tensor memory) consumed by dV
  uint64_t* dq_full = bars + 14;
  uint64_t* dq_empty = bars + 15;
  uint64_t* dkv_full = bars + 16;   // dV / dK of a key tile complete
  uint64_t* dkv_empty = bars + 17;  // ... and drained from tensor memory
  uint64_t* ds_empty = bars + 18;   // [2] dS^T smem buffer consumed by dK and dQ
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.y, b = blockIdx.z;
  const int nqt = (p.Nq + 127) / 128, nkt = (p.Nk + 127) / 128;
  const int kt0 = (int)(((long long)blockIdx.x * nkt) / gridDim.x), kt1 = (int)(((long long)(blockIdx.x + 1) * nkt) / gridDim.x);
  const int nk = kt1 - kt0;
  const int ntiles = nk * nqt;

  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) {
    printf("sd2: attn_bwd dynamic shared memory is not 1024-byte aligned\n");
    __trap();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmdO);
    if (!DIRECT) tma_prefetch_desc(&tmDQ);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < 2; ++s) {
        mbar_init(&kv_full[s], 1);
        mbar_init(&kv_empty[s], 1);
        mbar_init(&qdo_full[s], 1);
        mbar_init(&qdo_empty[s], 1);
        mbar_init(&ds_empty[s], 1);
      }
      mbar_init(s_full, 1);
      mbar_init(s_empty, AT_BWD_CWARPS);
      mbar_init(dp_full, 1);
      mbar_init(dp_empty, AT_BWD_CWARPS);
      mbar_init(pds_full, AT_BWD_CWARPS);
      mbar_init(pt_empty, 1);
      mbar_init(dq_full, 1);
      mbar_init(dq_empty, AT_BWD_CWARPS);
      mbar_init(dkv_full, 1);
      mbar_init(dkv_empty, AT_BWD_CWARPS);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // S^T, dP^T: 128 fp32 columns each; dV, dK, dQ accumulators: 64 each; P^T as bf16 pairs (A operand of dV): 64
  const uint32_t tST = tmem_base, tdPT = tmem_base + 128, tdV = tmem_base + 256, tdK = tmem_base + 320,
                 tdQ = tmem_base + 384, tPT = tmem_base + 448;
  pdl_grid_sync();

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer
    const float* st_bh = p.stats + ((long long)b * p.heads + h) * nqt * 256;
    int t = 0;
    for (int kk = 0; kk < nk; ++kk) {
      const int sk = KVB == 2 ? (kk & 1) : 0;
      mbar_wait(&kv_empty[sk], (uint32_t)((KVB == 2 ? (kk >> 1) : kk) & 1) ^ 1u);  // the MMAs that used this buffer are done
      if (elect_one()) {
        mbar_arrive_expect_tx(&kv_full[sk], 2 * AT_TILE);
        tma_load_4d(sKV + sk * 2 * AT_TILE, &tmK, &kv_full[sk], 0, (kt0 + kk) * 128, h, b);
        tma_load_4d(sKV + sk * 2 * AT_TILE + AT_TILE, &tmV, &kv_full[sk], 0, (kt0 + kk) * 128, h, b);
      }
      __syncwarp();
      for (int i = 0; i < nqt; ++i, ++t) {
        const int s = t & 1;
        mbar_wait(&qdo_empty[s], (uint32_t)((t >> 1) & 1) ^ 1u);
        if (elect_one()) {
          mbar_arrive_expect_tx(&qdo_full[s], 2 * AT_TILE + 1024);
          tma_load_4d(sQ + s * AT_TILE, &tmQ, &qdo_full[s], 0, i * 128, h, b);
          tma_load_4d(sdO + s * AT_TILE, &tmdO, &qdo_full[s], 0, i * 128, h, b);
          bulk_load_1d(sStat + s * 256, st_bh + (long long)i * 256, 1024, &qdo_full[s]);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer
    // Tensor-pipe order per tile t:  [dV, dK, dQ of tile t-1]  S^T(t+1)  dP^T(t+1)  - the score tile of the next iteration
    // (possibly the first query tile of the NEXT key tile, whose K / V were prefetched into the other buffer) is issued as
    // soon as the compute warps hold S^T(t) in registers, the dP^T one when dP^T(t) has been read.
    constexpr uint32_t id_s = umma_idesc_bf16(128, 128, 0, 0);   // S^T, dP^T : A K-major, B K-major
    constexpr uint32_t id_kn = umma_idesc_bf16(128, 64, 0, 1);   // dV, dK    : A K-major (TMEM P^T / smem dS^T), B MN-major
    constexpr uint32_t id_mn = umma_idesc_bf16(128, 64, 1, 1);   // dQ        : A MN-major (dS), B MN-major (K)
    constexpr uint64_t TS = AT_TILE >> 4;
    const uint64_t aK0 = umma_desc_sw128(smem_u32(sKV), 16, 1024), aV0 = aK0 + TS;
    const uint64_t bQ0 = umma_desc_sw128(smem_u32(sQ), 16, 1024), bdO0 = umma_desc_sw128(smem_u32(sdO), 16, 1024);
    const uint64_t adS = umma_desc_sw128(smem_u32(sdS), 16, 1024);
    const uint64_t bQ0mn = umma_desc_sw128(smem_u32(sQ), 8192, 1024), bdO0mn = umma_desc_sw128(smem_u32(sdO), 8192, 1024);
    const uint64_t adSmn = umma_desc_sw128(smem_u32(sdS), 16384, 1024);  // M chunks (64 queries) 16 KB apart
    const uint64_t bKmn0 = umma_desc_sw128(smem_u32(sKV), 8192, 1024);
    mbar_wait(&kv_full[0], 0);
    mbar_wait(&qdo_full[0], 0);
    tc_fence_after();
    if (elect_one()) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) tc_mma_bf16(tST, aK0 + 2 * ks, bQ0 + 2 * ks, id_s, ks != 0 ? 1u : 0u);
      tc_commit(s_full);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) tc_mma_bf16(tdPT, aV0 + 2 * ks, bdO0 + 2 * ks, id_s, ks != 0 ? 1u : 0u);
      tc_commit(dp_full);
    }
    __syncwarp();
    int t = 0;
    for (int kk = 0; kk < nk; ++kk) {
      const uint64_t kvb = KVB == 2 ? (uint64_t)(kk & 1) * 2 * TS : 0;  // K / V buffer of this key tile
      for (int i = 0; i < nqt; ++i, ++t) {
        const int s = t & 1;
        if (t + 1 < ntiles) {
          const int s1 = (t + 1) & 1;
          const uint64_t bQ = bQ0 + (uint64_t)s1 * TS, bdO = bdO0 + (uint64_t)s1 * TS;
          uint64_t kvn = kvb;
          if (i == nqt - 1) {  // the next tile belongs to the next key tile: its K / V arrive in the other buffer
            kvn = KVB == 2 ? (uint64_t)((kk + 1) & 1) * 2 * TS : 0;
            mbar_wait(&kv_full[KVB == 2 ? ((kk + 1) & 1) : 0], (uint32_t)((KVB == 2 ? ((kk + 1) >> 1) : (kk + 1)) & 1));
          }
          mbar_wait(s_empty, (uint32_t)(t & 1));
          mbar_wait(&qdo_full[s1], (uint32_t)(((t + 1) >> 1) & 1));
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) tc_mma_bf16(tST, aK0 + kvn + 2 * ks, bQ + 2 * ks, id_s, ks != 0 ? 1u : 0u);
            tc_commit(s_full);
          }
          __syncwarp();
          mbar_wait(dp_empty, (uint32_t)(t & 1));
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) tc_mma_bf16(tdPT, aV0 + kvn + 2 * ks, bdO + 2 * ks, id_s, ks != 0 ? 1u : 0u);
            tc_commit(dp_full);
          }
          __syncwarp();
        }
        mbar_wait(pds_full, (uint32_t)(t & 1));
        if (i == 0 && kk > 0) mbar_wait(dkv_empty, (uint32_t)((kk - 1) & 1));  // dV / dK of the previous key tile drained
        tc_fence_after();
        const uint64_t bQ = bQ0mn + (uint64_t)s * TS, bdO = bdO0mn + (uint64_t)s * TS;
        const uint64_t dsb = (uint64_t)s * 2 * TS;  // dS^T buffer of this tile
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)  // K = 128 queries, 16 per step = 8 TMEM columns of P^T
            tc_mma_bf16_ts(tdV, tPT + (uint32_t)(ks * 8), bdO + (uint64_t)ks * 128, id_kn, (i | ks) != 0 ? 1u : 0u);
          tc_commit(pt_empty);
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {  // dS^T from smem: chunk ks>>2 (16 KB apart), 32 B per K step inside
            const uint64_t a = adS + dsb + (uint64_t)(ks >> 2) * TS + 2 * (ks & 3);
            tc_mma_bf16(tdK, a, bQ + (uint64_t)ks * 128, id_kn, (i | ks) != 0 ? 1u : 0u);
          }
        }
        __syncwarp();
        if (t > 0) mbar_wait(dq_empty, (uint32_t)((t - 1) & 1));  // dQ of the previous tile has left tensor memory
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)  // K = 128 keys: 16 key rows (2048 B) per step in both operands
            tc_mma_bf16(tdQ, adSmn + dsb + (uint64_t)ks * 128, bKmn0 + kvb + (uint64_t)ks * 128, id_mn, ks != 0 ? 1u : 0u);
          tc_commit(&ds_empty[s]);
          tc_commit(dq_full);
          tc_commit(&qdo_empty[s]);
          if (i == nqt - 1) {
            tc_commit(dkv_full);
            tc_commit(&kv_empty[KVB == 2 ? (kk & 1) : 0]);
          }
        }
        __syncwarp();
      }
    }
  } else {
    // ---------------------------------------------------------------- compute warps
    // Four threads per key row: thread (r, hq) owns query columns hq*32 .. +31 of S^T / dP^T (16 bf16-pair columns of
    // P^T, four 16-byte groups of a dS^T row) and output columns hq*16 .. +15 of dQ, dV and dK.  Four warps per SM
    // sub-partition keep the MUFU / FMA pipes busy across each other's TMEM-load and barrier latencies.
    // Padding key rows of the last key tile need no masking: their K rows are zero-filled by TMA, so they add nothing to
    // dQ = dS K, and their dV / dK rows are never stored.
    const int q = warp & 3;
    const int hq = (warp - 2) >> 2;
    const int r = q * 32 + lane;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const uint32_t ds_row0 = smem_u32(sdS) + (uint32_t)(hq >> 1) * AT_TILE + r * 128;
    const int dsg0 = (hq & 1) * 4;  // first 16-byte group of this thread inside the 128-byte dS^T row
    const uint32_t stat0 = smem_u32(sStat) + hq * 128;  // this thread's 32 queries of the lse block; D block 512 B further
    const uint8_t* my_stg_g = stg + (size_t)(warp - 2) * STG_W;
    const uint32_t my_stg = smem_u32(my_stg_g);
    const float c2 = p.c2;
    const int bq = (DIRECT || FX) ? b : (int)blockIdx.x * p.B + b;  // image index inside the (per key-range split) dQ partial

    auto drain_dkv = [&](int kt, int kk) {  // dV, dK of key tile kt (this thread's 16 columns of each) -> bf16, global
      mbar_wait(dkv_full, (uint32_t)(kk & 1));
      tc_fence_after();
      const int key = kt * 128 + r;
#pragma unroll 1
      for (int tt = 0; tt < 2; ++tt) {
        bf16* dst = (tt == 0 ? p.dv : p.dk) + ((long long)b * p.Nk + key) * (tt == 0 ? p.lddv : p.lddk) + h * 64 + hq * 16;
        const float mul = tt == 0 ? 1.f : p.scale;
        uint32_t rr[16];
        tmem_ld_32x32b_x16((tt == 0 ? tdV : tdK) + lane_off + (uint32_t)(hq * 16), rr);
        tmem_wait_ld();
        if (key < p.Nk) {
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            uint4 u;
            u.x = pack_bf16x2(__uint_as_float(rr[g * 8 + 0]) * mul, __uint_as_float(rr[g * 8 + 1]) * mul);
            u.y = pack_bf16x2(__uint_as_float(rr[g * 8 + 2]) * mul, __uint_as_float(rr[g * 8 + 3]) * mul);
            u.z = pack_bf16x2(__uint_as_float(rr[g * 8 + 4]) * mul, __uint_as_float(rr[g * 8 + 5]) * mul);
            u.w = pack_bf16x2(__uint_as_float(rr[g * 8 + 6]) * mul, __uint_as_float(rr[g * 8 + 7]) * mul);
            *reinterpret_cast<uint4*>(dst + g * 8) = u;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(dkv_empty);
    };

    int t = 0;
    for (int kk = 0; kk < nk; ++kk) {
      for (int i = 0; i < nqt; ++i, ++t) {
        const uint32_t stat = stat0 + (uint32_t)(t & 1) * 1024;
        const uint32_t ds_row = ds_row0 + (uint32_t)(t & 1) * 2 * AT_TILE;
        mbar_wait(&qdo_full[t & 1], (uint32_t)((t >> 1) & 1));  // the statistics block of this query tile has landed
        mbar_wait(s_full, (uint32_t)(t & 1));
        tc_fence_after();
        uint32_t rs[32];
        tmem_ld_32x32b_x32(tST + lane_off + (uint32_t)(hq * 32), rs);
        tmem_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(s_empty);  // S^T is in registers: the tensor pipe may overwrite it with the next tile
        // P^T = exp2(S^T c2 - lse), kept in fp32 for dS and packed to bf16 pairs for the dV operand
        uint32_t pk[16];
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const float4 l = lds128f(stat + g * 16);
          const float e0 = ex2(fmaf(__uint_as_float(rs[g * 4 + 0]), c2, -l.x));
          const float e1 = ex2(fmaf(__uint_as_float(rs[g * 4 + 1]), c2, -l.y));
          const float e2 = ex2(fmaf(__uint_as_float(rs[g * 4 + 2]), c2, -l.z));
          const float e3 = ex2(fmaf(__uint_as_float(rs[g * 4 + 3]), c2, -l.w));
          rs[g * 4 + 0] = __float_as_uint(e0); rs[g * 4 + 1] = __float_as_uint(e1);
          rs[g * 4 + 2] = __float_as_uint(e2); rs[g * 4 + 3] = __float_as_uint(e3);
          pk[g * 2] = pack_bf16x2(e0, e1);
          pk[g * 2 + 1] = pack_bf16x2(e2, e3);
        }
        // dV / dK of the previous key tile: complete since its last tile's MMAs, drained here (one exp phase later)
        if (i == 0 && kk > 0) drain_dkv(kt0 + kk - 1, kk - 1);
        mbar_wait(pt_empty, (uint32_t)(t & 1) ^ 1u);  // P^T of the previous tile consumed by its dV MMAs (first in their order)
        tc_fence_after();
        tmem_st_32x32b_x16(tPT + lane_off + (uint32_t)(hq * 16), pk);
        // dS^T = P^T (dP^T - D)   (the 1/sqrt(d) factor is applied when dQ / dK are drained)
        mbar_wait(dp_full, (uint32_t)(t & 1));
        tc_fence_after();
        uint32_t rd[32];
        tmem_ld_32x32b_x32(tdPT + lane_off + (uint32_t)(hq * 32), rd);
        tmem_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(dp_empty);
        mbar_wait(&ds_empty[t & 1], (uint32_t)((t >> 1) & 1) ^ 1u);  // this dS^T buffer was read by the MMAs of tile t-2
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const float4 d0 = lds128f(stat + 512 + g * 32), d1 = lds128f(stat + 512 + g * 32 + 16);
          const int e = g * 8;
          uint4 u;
          u.x = pack_bf16x2(__uint_as_float(rs[e + 0]) * (__uint_as_float(rd[e + 0]) - d0.x),
                            __uint_as_float(rs[e + 1]) * (__uint_as_float(rd[e + 1]) - d0.y));
          u.y = pack_bf16x2(__uint_as_float(rs[e + 2]) * (__uint_as_float(rd[e + 2]) - d0.z),
                            __uint_as_float(rs[e + 3]) * (__uint_as_float(rd[e + 3]) - d0.w));
          u.z = pack_bf16x2(__uint_as_float(rs[e + 4]) * (__uint_as_float(rd[e + 4]) - d1.x),
                            __uint_as_float(rs[e + 5]) * (__uint_as_float(rd[e + 5]) - d1.y));
          u.w = pack_bf16x2(__uint_as_float(rs[e + 6]) * (__uint_as_float(rd[e + 6]) - d1.z),
                            __uint_as_float(rs[e + 7]) * (__uint_as_float(rd[e + 7]) - d1.w));
          sts128(ds_row + (uint32_t)(((dsg0 + g) ^ (r & 7)) << 4), u);
        }
        tmem_wait_st();
        tc_fence_before();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(pds_full);
        if (t > 0) {  // dQ of the previous tile (rows = queries): its MMAs were issued a whole tile of math ago
          const int ip = i > 0 ? i - 1 : nqt - 1;
          mbar_wait(dq_full, (uint32_t)((t - 1) & 1));
          tc_fence_after();
          attn_bwd_drain_dq<DIRECT, FX>(tdQ + lane_off + (uint32_t)(hq * 16), my_stg, my_stg_g, lane, p.scale, dq_empty, &tmDQ,
                                    hq * 16, ip * 128 + q * 32, p.Nq, h, bq, p.dq, p.lddq, nqt == 1, t - 1 < nqt, p.unordered);
        }
      }
    }
    if (ntiles > 0) {  // last tile's dQ, last key tile's dV / dK
      mbar_wait(dq_full, (uint32_t)((ntiles - 1) & 1));
      tc_fence_after();
      attn_bwd_drain_dq<DIRECT, FX>(tdQ + lane_off + (uint32_t)(hq * 16), my_stg, my_stg_g, lane, p.scale, dq_empty, &tmDQ, hq * 16,
                                (nqt - 1) * 128 + q * 32, p.Nq, h, bq, p.dq, p.lddq, nqt == 1, ntiles - 1 < nqt, p.unordered);
      drain_dkv(kt1 - 1, nk - 1);
    }
    if (!DIRECT) {
      if (lane == 0) bulk_wait_read<0>();
      __syncwarp();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

static bool head_tmap(CUtensorMap* tm, const void* ptr, long long ld, int N, int heads, int B, std::string* err) {
  sd2_operand o;
  o.ptr = ptr;
  o.mn_major = 0;
  o.cols = 64;
  o.rows = N;
  o.ld = ld;
  o.nb0 = heads;
  o.nb1 = B;
  o.bs0 = 64;
  o.bs1 = (long long)N * ld;
  return plain_tmap(tm, o, 128, err);
}

}  // namespace sd2

using namespace sd2;

extern "C" {

int sd2_attn_fwd(sd2_ctx* ctx, const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv,
                 void* o, long long ldo, float* lse, int B, int heads, int Nq, int Nk, int head_dim, float scale,
                 sd2_stream stream_) {
  if (!ctx) return 1;
  if (head_dim != 64) return fail(ctx, "sd2_attn_fwd: head_dim must be 64");
  if (B < 1 || heads < 1 || Nq < 1 || Nk < 1) return fail(ctx, "sd2_attn_fwd: empty problem");
  if ((ldq | ldk | ldv | ldo) % 8) return fail(ctx, "sd2_attn_fwd: row strides must be multiples of 8");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CUtensorMap tmQ, tmK, tmV;
  std::string err;
  if (!head_tmap(&tmQ, q, ldq, Nq, heads, B, &err) || !head_tmap(&tmK, k, ldk, Nk, heads, B, &err) ||
      !head_tmap(&tmV, v, ldv, Nk, heads, B, &err))
    return fail(ctx, "sd2_attn_fwd: " + err);
  AttnParams p;
  memset(&p, 0, sizeof(p));
  p.B = B; p.heads = heads; p.Nq = Nq; p.Nk = Nk;
  p.scale = scale;
  p.c2 = scale * 1.4426950408889634f;
  p.o = reinterpret_cast<bf16*>(o);
  p.ldo = ldo;
  p.lse = lse;
  const size_t smem = 4 * AT_TILE + 512 * 4 + 16 * 8 + 16 + 1024;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(ctx, std::string("sd2_attn_fwd attr: ") + cudaGetErrorString(e));
    attr = true;
  }
  cudaError_t le = launch_k(attn_fwd_kernel, dim3((Nq + 127) / 128, heads, B), dim3(AT_FWD_THREADS), smem, stream, tmQ, tmK, tmV, p);
  if (le != cudaSuccess) return fail(ctx, std::string("sd2_attn_fwd launch: ") + cudaGetErrorString(le));
  return check_launch(ctx, "attn_fwd");
}

// Key-range splits per (image, head).  One CTA per SM: the grid runs in waves of 148, so few (image, head) pairs leave
// the last wave (or the whole GPU) partly idle; splitting the key range over 2 / 4 / 8 CTAs fills it at the price of one
// fp32 dQ partial per split (written by the kernel, summed in order by the cast pass).  Small cost model: waves x tiles per
// CTA x ~2 us per tile + partial traffic at ~6 TB/s.
static int attn_bwd_ksplit(int B, int heads, int Nq, int nkt) {
  const double bh = (double)B * heads;
  const int nqt = (Nq + 127) / 128;
  const double part_bytes = bh * Nq * 64.0 * 4.0;
  int best = 1;
  double best_cost = 1e30;
  for (int ks = 1; ks <= 8 && ks <= nkt; ks *= 2) {
    const double cost = ceil(bh * ks / 148.0) * ceil((double)nkt / ks) * nqt * 2.0e-6 + ks * part_bytes / 6.0e12;
    if (cost < best_cost * 0.97) {
      best_cost = cost;
      best = ks;
    }
  }
  return best;
}

// Which dQ accumulation scheme: the sequential walk (ordered fp32 partials) while its live dQ working set - one fp32 slice
// of Nq x 64 per resident CTA - fits well inside the L2; beyond that one key tile per CTA on the shared fixed-point accumulator.
// SD2_ATTN_FX=0/1 forces one of them (A/B measurements).
static bool attn_bwd_use_fx(int Nq, int nkt) {
  static const int forced = getenv("SD2_ATTN_FX") ? atoi(getenv("SD2_ATTN_FX")) : -1;
  if (nkt <= 1) return false;
  if (forced >= 0) return forced != 0;
  return 148.0 * Nq * 64.0 * 4.0 > 64.0e6;  // N >= 2048
}

long long sd2_attn_bwd_ws_bytes(int B, int heads, int Nq) {
  const long long nqt = (Nq + 127) / 128;
  const long long elems = (long long)B * Nq * heads * 64;
  const long long acc = std::max(elems * 8, (long long)attn_bwd_ksplit(B, heads, Nq, (int)nqt) * elems * 4);  // either scheme
  return acc + (long long)B * heads * nqt * 256 * 4;
}

int sd2_attn_bwd(sd2_ctx* ctx, const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv,
                 const void* o, long long ldo, const void* d_o, long long lddo, const float* lse, void* dq, long long lddq,
                 void* dk, long long lddk, void* dv, long long lddv, void* ws, int B, int heads, int Nq, int Nk,
                 int head_dim, float scale, sd2_stream stream_) {
  if (!ctx) return 1;
  if (head_dim != 64) return fail(ctx, "sd2_attn_bwd: head_dim must be 64");
  if (B < 1 || heads < 1 || Nq < 1 || Nk < 1) return fail(ctx, "sd2_attn_bwd: empty problem");
  if ((ldq | ldk | ldv | ldo | lddo | lddq | lddk | lddv) % 8) return fail(ctx, "sd2_attn_bwd: row strides must be multiples of 8");
  if (!ws) return fail(ctx, "sd2_attn_bwd: workspace required (sd2_attn_bwd_ws_bytes)");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const int C = heads * 64;
  const int nqt = (Nq + 127) / 128, nkt = (Nk + 127) / 128;
  const bool direct = nkt == 1;  // one key tile: every dQ tile has a single contribution and is written as bf16 directly
  const bool fx = !direct && attn_bwd_use_fx(Nq, nkt);
  // the workspace (sd2_attn_bwd_ws_bytes knows B, heads, Nq only) holds attn_bwd_ksplit(B, heads, nqt) partials
  const int ks_ws = attn_bwd_ksplit(B, heads, Nq, nqt);
  int ksplit = direct ? 1 : attn_bwd_ksplit(B, heads, Nq, nkt);
  if (ksplit > ks_ws) ksplit = ks_ws;
  const long long part = (long long)B * Nq * C;
  const long long acc_bytes = std::max(part * 8, (long long)ks_ws * part * 4);
  float* dq32 = reinterpret_cast<float*>(ws);                     // ordered scheme: [ksplit][B*Nq][C] fp32 dQ partials
  long long* dq64 = reinterpret_cast<long long*>(ws);             // fixed-point scheme: [B*Nq][C] int64
  float* stats = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + acc_bytes);  // [B*heads][nqt][lse 128 | D 128]
  CUtensorMap tmQ, tmK, tmV, tmdO, tmDQ;
  std::string err;
  if (!head_tmap(&tmQ, q, ldq, Nq, heads, B, &err) || !head_tmap(&tmK, k, ldk, Nk, heads, B, &err) ||
      !head_tmap(&tmV, v, ldv, Nk, heads, B, &err) || !head_tmap(&tmdO, d_o, lddo, Nq, heads, B, &err))
    return fail(ctx, "sd2_attn_bwd: " + err);
  if (fx) {  // accumulator as a [B][heads][Nq][64] view of 64-bit integers, box = 32 rows x 16 columns (128-byte rows, SWIZZLE_128B)
    const uint64_t dims[4] = {64, (uint64_t)Nq, (uint64_t)heads, (uint64_t)B};
    const uint64_t strides[3] = {(uint64_t)C * 8, 64 * 8, (uint64_t)Nq * C * 8};
    const uint32_t box[4] = {16, 32, 1, 1};
    if (!encode_tmap_4d(&tmDQ, SD2_DT_U64_INTERNAL, 128, dq64, dims, strides, box, &err))
      return fail(ctx, "sd2_attn_bwd dq map: " + err);
  } else {
    // fp32 dQ partials as a [ksplit * B][heads][Nq][64]-strided view, box = 32 rows x 16 columns (64-byte rows, SWIZZLE_64B)
    if (!out_tmap(&tmDQ, dq32, true, 16, 64, Nq, C, heads, (long long)B * ksplit, 64, (long long)Nq * C, &err))
      return fail(ctx, "sd2_attn_bwd dq map: " + err);
  }
  AttnParams p;
  memset(&p, 0, sizeof(p));
  p.B = B; p.heads = heads; p.Nq = Nq; p.Nk = Nk;
  p.scale = scale;
  p.c2 = scale * 1.4426950408889634f;
  p.stats = stats;
  p.dq = reinterpret_cast<bf16*>(dq);
  p.lddq = lddq;
  p.dq32 = dq32;
  p.dq64 = dq64;
  {
    static const int unordered = getenv("SD2_ATTN_UNORDERED") ? atoi(getenv("SD2_ATTN_UNORDERED")) : 0;
    p.unordered = unordered;
    static const int force_ks = getenv("SD2_ATTN_KSPLIT") ? atoi(getenv("SD2_ATTN_KSPLIT")) : 0;
    if (force_ks > 0 && !direct && force_ks <= ks_ws && force_ks <= nkt) ksplit = force_ks;
  }
  p.dk = reinterpret_cast<bf16*>(dk);
  p.lddk = lddk;
  p.dv = reinterpret_cast<bf16*>(dv);
  p.lddv = lddv;
  const long long nd = (long long)B * nqt * 128 * heads * 8;
  launch_k(attn_bwd_prep_kernel, dim3((unsigned)((nd + 255) / 256)), dim3(256), 0, stream, reinterpret_cast<const bf16*>(o), ldo,
           reinterpret_cast<const bf16*>(d_o), lddo, lse, stats, B, heads, Nq, nqt);
  cudaError_t e;
  // both layouts: 12 tiles + 32 KB staging (two K / V buffers) = 10 tiles + 64 KB staging (one K / V buffer, int64 staging)
  const size_t smem = 12 * AT_TILE + AT_BWD_CWARPS * 2048 + 512 * 4 + 20 * 8 + 16;
  static bool attr = false;
  if (!attr) {
    e = cudaFuncSetAttribute(attn_bwd_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_bwd_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_bwd_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(ctx, std::string("sd2_attn_bwd attr: ") + cudaGetErrorString(e));
    attr = true;
  }
  if (direct) {
    e = launch_k(attn_bwd_kernel<true, false>, dim3(1, heads, B), dim3(AT_BWD_THREADS), smem, stream, tmQ, tmK, tmV, tmdO, tmDQ, p);
    if (e != cudaSuccess) return fail(ctx, std::string("sd2_attn_bwd launch: ") + cudaGetErrorString(e));
    return check_launch(ctx, "attn_bwd", 2);
  }
  if (fx) {  // one key tile per CTA, shared fixed-point accumulator (cleared first: every CTA adds)
    e = cudaMemsetAsync(dq64, 0, (size_t)part * 8, stream);
    if (e != cudaSuccess) return fail(ctx, std::string("sd2_attn_bwd memset: ") + cudaGetErrorString(e));
    attn_bwd_kernel<false, true><<<dim3(nkt, heads, B), AT_BWD_THREADS, smem, stream>>>(tmQ, tmK, tmV, tmdO, tmDQ, p);
    launch_k(cast2d_fx64_bf16_kernel, dim3(grid_for((long long)B * Nq * (C / 8), 256, ctx->num_sms)), dim3(256), 0, stream,
             (const long long*)dq64, reinterpret_cast<bf16*>(dq), lddq, (long long)B * Nq, C);
    return check_launch(ctx, "attn_bwd", 3);
  }
  e = launch_k(attn_bwd_kernel<false, false>, dim3(ksplit, heads, B), dim3(AT_BWD_THREADS), smem, stream, tmQ, tmK, tmV, tmdO, tmDQ, p);
  if (e != cudaSuccess) return fail(ctx, std::string("sd2_attn_bwd launch: ") + cudaGetErrorString(e));
  launch_k(cast2d_f32_bf16_kernel, dim3(grid_for((long long)B * Nq * (C / 8), 256, ctx->num_sms)), dim3(256), 0, stream,
           (const float*)dq32, reinterpret_cast<bf16*>(dq), lddq, (long long)B * Nq, C, ksplit, part);
  return check_launch(ctx, "attn_bwd", 3);
}

}  // extern "C"

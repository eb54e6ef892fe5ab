// tcgen05 / TMEM / TMA GEMM and implicit-GEMM 3x3 convolution for sm_100a.
//
// Persistent kernel, one CTA per SM, each CTA walks work items (128 x BN output tiles, optionally one K-split of a
// tile) in a static round-robin order:
//   warp 0      : TMA producer  (cp.async.bulk.tensor 4-D boxes, SWIZZLE_128B, mbarrier complete_tx) into a 4..8 stage ring
//   warp 1      : TMEM allocator + single-thread tcgen05.mma issuer (kind::f16, bf16 x bf16 -> fp32) into one of TWO
//                 accumulator buffers in tensor memory, so the next tile's main loop overlaps this tile's epilogue
//   warps 2..9  : epilogue (two warps per TMEM lane quadrant, each owning half of the tile's columns, so every SM
//                 sub-partition has two epilogue warps to hide latencies):
//                 tcgen05.ld 32x32b -> registers -> alpha / bias / per-image bias / residual -> swizzled smem
//                 staging -> TMA store (bf16 / fp32) or TMA reduce-add (fp32 weight gradients, split-K sums); every
//                 global write is a full coalesced row segment issued by the copy engine, no per-thread stores or atomics
// CL = 2: the kernel runs as thread-block clusters of two CTAs that work on the two M tiles of a pair with the SAME B tile:
// each CTA fetches half of B and TMA-multicasts it into both CTAs' shared memory, so the operand traffic L2 -> SM per
// k-block drops from A + B to A + B/2 (the 128 x 256 tile is L2-fabric bound at ~2/3 of the tensor peak otherwise: 48 KB
// of operands per 512 tensor-pipe cycles on each of 148 SMs).  A stage is released by both CTAs' MMA warps (multicast commit).
// Operands can be K-major (forward, activations x weights) or MN-major (dgrad reads the weights transposed, wgrad
// contracts over pixels) - the same TMA boxes serve both, only the UMMA descriptors differ.
//
// Replaces: cuDNN implicit-GEMM conv + cuBLASLt linear behind diffusers' nn.Conv2d / nn.Linear, called from
// reference diffusion/models/stable_diffusion.py:183 (UNet forward) and the autograd backward of the same.
#include "gemm_tc.cuh"
#include "host.h"

namespace sd2 {

static constexpr int BM = 128;
static constexpr int BK = 64;
static constexpr int A_BYTES = BM * BK * 2;      // 16 KB
static constexpr int CHUNK_BYTES = 64 * BK * 2;  // one 64(mn) x 64(k) MN-major box = 8 KB
static constexpr int STG_BYTES = 4096;           // one epilogue staging buffer: 32 rows x 128 B
static constexpr int EPI_WARPS = 8;              // two warps per TMEM lane quadrant, each owning half of the tile's columns
static constexpr int STG_TOTAL = EPI_WARPS * STG_BYTES;
static constexpr int GEMM_THREADS = 64 + EPI_WARPS * 32;
static constexpr int SMEM_LIMIT = 227 * 1024;

template <int BN>
struct TileCfg {
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int ACC_STRIDE = BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : 256;  // TMEM columns per buffer
  static constexpr int TMEM_COLS = 2 * ACC_STRIDE;
  static constexpr int OUT_CH = (BN % 128 == 0) ? 64 : 32;  // bf16 columns per TMA store box (64 needs an even chunk split)
};

int gemm_stages(int BN) {
  const int stage_bytes = A_BYTES + BN * BK * 2;
  int s = (SMEM_LIMIT - 1024 - 256 - STG_TOTAL) / stage_bytes;
  return s < 2 ? 2 : (s > 8 ? 8 : s);
}
// cta_group::2 pair mode: a stage is the CTA's A tile and HALF of the B tile
int gemm_stages_pair(int BN) {
  const int stage_bytes = A_BYTES + BN * BK;
  int s = (SMEM_LIMIT - 1024 - 256 - STG_TOTAL) / stage_bytes;
  return s < 2 ? 2 : (s > 8 ? 8 : s);
}
int gemm_out_chunk(int BN) { return (BN % 128 == 0) ? 64 : 32; }

struct WorkItem {
  int m_tile, n_tile, split, batch, kb0, nkb;
};
// mt = number of M work units: M tiles, or pairs of M tiles when the kernel runs as 2-CTA clusters (m_tile then counts pairs)
__device__ __forceinline__ WorkItem decode_item(const GemmKParams& p, int item, int mt) {
  WorkItem w;
  int t;
  if (p.raster == 1) {
    w.n_tile = item % p.nt;
    t = item / p.nt;
    w.m_tile = t % mt;
    t /= mt;
    w.split = t % p.splits;
    w.batch = t / p.splits;
  } else {
    w.m_tile = item % mt;
    t = item / mt;
    w.n_tile = t % p.nt;
    t /= p.nt;
    if (p.raster == 2) {
      w.batch = t % p.batches;
      w.split = t / p.batches;
    } else {
      w.split = t % p.splits;
      w.batch = t / p.splits;
    }
  }
  w.kb0 = (int)(((long long)w.split * p.total_kb) / p.splits);
  w.nkb = (int)(((long long)(w.split + 1) * p.total_kb) / p.splits) - w.kb0;
  return w;
}

// ---------------------------------------------------------------------------------------------- epilogue pieces
struct EpiTile {  // per work item, per thread (thread = one output row of the 32-row slab of its warp)
  int row0, o2, o3, n_tile0, N, M, out_mode;
  const float* rb;  // per-image bias row of this thread (or null)
  const bf16* rs;   // residual row of this thread (or null)
  bool has_bias;
  float alpha;
  float bv[8];      // bias of tile columns lane*8 .. lane*8+7
  float* gn_part;   // GroupNorm partial sums of the output (or null), see GemmKParams
  int gn_slab;
  long long row;    // this thread's output row
  bool row_ok;
};
struct EpiPre {  // residual of one 32-column chunk, fetched one chunk ahead (the per-image bias rows are tiny and
  uint4 rs[4];    // L1-resident: they are read in place)
};
struct EpiState {
  int c0, c1;  // this warp's range of 32-column chunks
};

__device__ __forceinline__ void epi_prefetch(const EpiTile& t, int c, EpiPre& pre) {
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const int n = t.n_tile0 + c * 32 + g * 8;
    if (n < t.N && t.rs != nullptr) pre.rs[g] = *reinterpret_cast<const uint4*>(t.rs + n);
  }
}

// One 32-column chunk of the accumulator: TMEM -> registers -> (+bias +rowbias +residual) -> swizzled staging -> TMA.
// GroupNorm statistics of one staged bf16 tile (32 rows x OUT_CH columns, swizzled as the TMA store expects it): every lane
// owns one 32-bit word (two columns) and walks the rows: per column sum and sum of squares of the rounded values, kept
// apart for rows 0-15 / 16-31 (two 16-row slabs) or added (one 32-row slab).  With 32-column tiles the two half-warps split
// the rows.  Conflict-free: in every row the 32 lanes read 32 different banks.
template <int OUT_CH>
__device__ __forceinline__ void epi_gn_stats(const EpiTile& t, uint32_t buf_s, int col0, int lane) {
  float s[2][2] = {{0.f, 0.f}, {0.f, 0.f}}, q[2][2] = {{0.f, 0.f}, {0.f, 0.f}};  // [row half][column of the pair]
  const int rows_ok = t.M - t.row0;  // rows of this 32-row slab inside the matrix
  if (OUT_CH == 64) {
    const int c = lane >> 2, w = lane & 3;
#pragma unroll
    for (int hf = 0; hf < 2; ++hf)
#pragma unroll
      for (int k = hf * 16; k < hf * 16 + 16; ++k) {
        uint32_t u;
        asm volatile("ld.shared.b32 %0, [%1];" : "=r"(u) : "r"(buf_s + k * 128 + ((c ^ (k & 7)) << 4) + w * 4) : "memory");
        if (k < rows_ok) {
          const float2 x = unpack_bf16x2(u);
          s[hf][0] += x.x; q[hf][0] = fmaf(x.x, x.x, q[hf][0]);
          s[hf][1] += x.y; q[hf][1] = fmaf(x.y, x.y, q[hf][1]);
        }
      }
    const int col = col0 + 2 * lane;
    if (col < t.N) {
      if (t.gn_slab == 32) {
        float* o = t.gn_part + ((long long)(t.row0 >> 5) * t.N + col) * 2;
        *reinterpret_cast<float4*>(o) = make_float4(s[0][0] + s[1][0], q[0][0] + q[1][0], s[0][1] + s[1][1], q[0][1] + q[1][1]);
      } else {
        float* o = t.gn_part + ((long long)(t.row0 >> 4) * t.N + col) * 2;
        *reinterpret_cast<float4*>(o) = make_float4(s[0][0], q[0][0], s[0][1], q[0][1]);
        if (rows_ok > 16) *reinterpret_cast<float4*>(o + 2 * (long long)t.N) = make_float4(s[1][0], q[1][0], s[1][1], q[1][1]);
      }
    }
  } else {  // 32 columns: 16 words per row, half-warp hf = lane >> 4 takes rows hf*16 .. hf*16+15
    const int hf = lane >> 4, j = lane & 15, c = j >> 2, w = j & 3;
    float s0 = 0.f, q0 = 0.f, s1 = 0.f, q1 = 0.f;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const int row = hf * 16 + k;
      uint32_t u;
      asm volatile("ld.shared.b32 %0, [%1];" : "=r"(u) : "r"(buf_s + row * 64 + ((c ^ ((row >> 1) & 3)) << 4) + w * 4) : "memory");
      if (row < rows_ok) {
        const float2 x = unpack_bf16x2(u);
        s0 += x.x; q0 = fmaf(x.x, x.x, q0);
        s1 += x.y; q1 = fmaf(x.y, x.y, q1);
      }
    }
    const int col = col0 + 2 * j;
    if (t.gn_slab == 32) {  // one slab: add the two halves (lane j + lane j+16)
      s0 += __shfl_xor_sync(0xffffffffu, s0, 16); q0 += __shfl_xor_sync(0xffffffffu, q0, 16);
      s1 += __shfl_xor_sync(0xffffffffu, s1, 16); q1 += __shfl_xor_sync(0xffffffffu, q1, 16);
      if (hf == 0 && col < t.N)
        *reinterpret_cast<float4*>(t.gn_part + ((long long)(t.row0 >> 5) * t.N + col) * 2) = make_float4(s0, q0, s1, q1);
    } else if (col < t.N && hf * 16 < rows_ok) {
      *reinterpret_cast<float4*>(t.gn_part + ((long long)((t.row0 >> 4) + hf) * t.N + col) * 2) = make_float4(s0, q0, s1, q1);
    }
  }
}

__device__ __forceinline__ float mse_target_at(const void* tgt, int dtype, long long i) {
  if (dtype == SD2_DT_F32) return reinterpret_cast<const float*>(tgt)[i];
  if (dtype == SD2_DT_BF16) return __bfloat162float(reinterpret_cast<const bf16*>(tgt)[i]);
  return __half2float(reinterpret_cast<const __half*>(tgt)[i]);
}

template <int BN>
__device__ __forceinline__ void epi_chunk(const EpiTile& t, EpiState& st, int c, const EpiPre& cur, EpiPre& nxt,
                                          uint32_t t_addr, uint8_t* stg, int lane, const CUtensorMap* tmO,
                                          uint64_t* tmem_empty, bool empty_on_leader, const GemmKParams& p) {
  constexpr int OUT_CH = (BN % 128 == 0) ? 64 : 32;
  const bool f32_out = t.out_mode != OUT_BF16;
  uint32_t r[32];
  tmem_ld_32x32b_x32(t_addr + (uint32_t)(c * 32), r);
  tmem_wait_ld();
  if (c == st.c1 - 1) {  // this warp's columns are fully read: its share of handing the TMEM buffer back to the MMA warp
    tc_fence_before();
    __syncwarp();
    if (lane == 0) {
      if (empty_on_leader) mbar_arrive_cluster(tmem_empty, 0);  // cta_group::2: the even CTA's MMA warp serves both accumulators
      else mbar_arrive(tmem_empty);
    }
  } else {
    epi_prefetch(t, c + 1, nxt);
  }
  const int n_base = t.n_tile0 + c * 32;
  const int cl = c - st.c0;  // chunk index within this warp's column range
  const bool new_buf = f32_out || OUT_CH == 32 || (cl & 1) == 0;
  uint8_t* buf = stg;
  const uint32_t buf_s = smem_u32(stg);
  bool waited = false;
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const int n = n_base + g * 8;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(r[g * 8 + e]) * t.alpha;
    if (t.has_bias) {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] += __shfl_sync(0xffffffffu, t.bv[e], c * 4 + g);
    }
    if (n < t.N) {
      if (t.rb != nullptr) {
        const float4 b0 = *reinterpret_cast<const float4*>(t.rb + n), b1 = *reinterpret_cast<const float4*>(t.rb + n + 4);
        v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
        v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
      }
      if (t.rs != nullptr) {
        const uint4 rr = cur.rs[g];
        const float2 r0 = unpack_bf16x2(rr.x), r1 = unpack_bf16x2(rr.y), r2 = unpack_bf16x2(rr.z), r3 = unpack_bf16x2(rr.w);
        v[0] += r0.x; v[1] += r0.y; v[2] += r1.x; v[3] += r1.y;
        v[4] += r2.x; v[5] += r2.y; v[6] += r3.x; v[7] += r3.y;
      }
    }
    if (BN == 64 && p.mse_target != nullptr && g == 0 && n_base == 0) {
      // MSE head (conv_out): columns 0..3 of this row are the prediction of one pixel; compare the bf16-rounded values (the
      // stored tensor, what F.mse_loss would read) with the target noise
      float sq = 0.f;
      if (t.row_ok) {
        const long long b = t.row / p.mse_hw, hw = t.row % p.mse_hw;
        float d[4];
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          const float pr = __bfloat162float(__float2bfloat16_rn(v[ch]));
          const float diff = pr - mse_target_at(p.mse_target, p.mse_dtype, (b * 4 + ch) * p.mse_hw + hw);
          sq = fmaf(diff, diff, sq);
          d[ch] = p.mse_k * diff;
        }
        uint4 o;
        o.x = pack_bf16x2(d[0], d[1]);
        o.y = pack_bf16x2(d[2], d[3]);
        o.z = 0u;
        o.w = 0u;
        *reinterpret_cast<uint4*>(p.mse_dpred8 + t.row * 8) = o;
      }
      sq = warp_sum(sq);
      if (lane == 0) atomicAdd(p.mse_acc, sq);
    }
    if (new_buf && !waited) {  // the staging buffer was handed to the copy engine by the previous store of this warp
      if (lane == 0) bulk_wait_read<0>();
      __syncwarp();
      waited = true;
    }
    // explicit st.shared (the re-aligned dynamic smem pointer would otherwise compile to generic ST.E)
    if (f32_out) {  // staging tile: 32 rows x 32 fp32 (128 B rows), SWIZZLE_128B
      const uint32_t rowp = buf_s + lane * 128;
      sts128f(rowp + (((2 * g) ^ (lane & 7)) << 4), v[0], v[1], v[2], v[3]);
      sts128f(rowp + (((2 * g + 1) ^ (lane & 7)) << 4), v[4], v[5], v[6], v[7]);
    } else {
      uint4 o;
      o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]);
      o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
      if (OUT_CH == 64) {  // 32 rows x 64 bf16 (128 B rows), SWIZZLE_128B; this ld fills half a row
        const int j = (cl & 1) * 4 + g;
        sts128(buf_s + lane * 128 + ((j ^ (lane & 7)) << 4), o);
      } else {  // 32 rows x 32 bf16 (64 B rows), SWIZZLE_64B
        sts128(buf_s + lane * 64 + ((g ^ ((lane >> 1) & 3)) << 4), o);
      }
    }
  }
  const bool full = f32_out || OUT_CH == 32 || (cl & 1) == 1;
  if (full) {
    fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the async proxy (TMA)
    __syncwarp();
    const int col0 = (f32_out || OUT_CH == 32) ? n_base : n_base - 32;
    if (lane == 0 && t.row0 < t.M && col0 < t.N) {
      if (t.out_mode == OUT_F32_ACCUM)
        tma_reduce_add_4d(tmO, buf, col0, t.row0, t.o2, t.o3);
      else
        tma_store_4d(tmO, buf, col0, t.row0, t.o2, t.o3);
    }
    if (lane == 0) bulk_commit();
    // GroupNorm statistics of the tile just handed to the copy engine (it only reads the staging buffer, as we do)
    if (!f32_out && t.gn_part != nullptr && t.row0 < t.M && col0 < t.N) epi_gn_stats<OUT_CH>(t, buf_s, col0, lane);
  }
}

template <int BN, bool A_MN, bool B_MN, int CL>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
    gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   const __grid_constant__ CUtensorMap tmO, const GemmKParams p, const int stages) {
  using Cfg = TileCfg<BN>;
  static_assert(!B_MN || BN % 64 == 0, "MN-major B needs 64-wide chunks");
  static_assert(CL == 1 || (B_MN ? BN % 128 == 0 : BN % 16 == 0), "2-CTA clusters split the B tile in two halves");
  // CL = 2: each CTA runs its own 128 x BN MMAs on a B tile the two CTAs fetch half each and multicast.
  // CL = 3 (PAIR): cta_group::2 - the even CTA issues ONE 256 x BN MMA per K step for the pair; every CTA keeps its 128 rows of A,
  // its HALF of B and its 128 accumulator rows, so a stage is A + B/2 (more stages, half the shared-memory reads of B per SM).
  constexpr bool CLU = CL >= 2, PAIR = CL == 3;
  constexpr int STAGE_B = PAIR ? A_BYTES + Cfg::B_BYTES / 2 : Cfg::STAGE_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* stg_base = smem + (size_t)stages * STAGE_B;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(stg_base + STG_TOTAL);
  uint64_t* empty_bar = full_bar + stages;
  uint64_t* tmem_full_bar = empty_bar + stages;  // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // work units: one M tile (CL = 1) or a pair of M tiles, one per CTA of the cluster (CL = 2; an odd tile count leaves the
  // last pair's second CTA with an out-of-range tile: its loads are zero-filled, its stores clipped)
  const int crank = CLU ? (int)cluster_ctarank() : 0;
  const int mtd = CLU ? (p.mt + 1) / 2 : p.mt;
  const int n_items = mtd * p.nt * p.splits * p.batches;
  const int unit0 = CLU ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int unit_step = CLU ? (int)(gridDim.x >> 1) : (int)gridDim.x;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmO);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < stages; ++s) {
        mbar_init(&full_bar[s], 1);
        mbar_init(&empty_bar[s], CL == 2 ? 2 : 1);  // one MMA-warp commit per issuing CTA (PAIR: the even CTA's, multicast)
      }
      for (int a = 0; a < 2; ++a) {
        mbar_init(&tmem_full_bar[a], 1);
        mbar_init(&tmem_empty_bar[a], PAIR ? 2 * EPI_WARPS : EPI_WARPS);  // one arrive per epilogue warp (PAIR: of both CTAs)
      }
      fence_mbar_init();
    }
    __syncwarp();
    if (PAIR) {
      tmem_alloc_cg2(tmem_slot, Cfg::TMEM_COLS);
      tmem_relinquish_cg2();
    } else {
      tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if (CLU) cluster_sync_all();  // the peer's barriers are initialised before anything of ours can reach them
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_grid_sync();  // everything above is input-independent and overlaps the previous kernel's tail

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    // The whole warp runs the (warp-uniform) control flow; one elected lane issues the copies.  Per k-block the
    // loop is a barrier wait, one expect_tx and 2-5 TMA instructions: tap / channel-block / pixel-block counters are
    // advanced incrementally (no division in the loop) so the producer stays far ahead of the tensor pipe.
    int s = 0;
    uint32_t ph = 0;
    const int bpi = p.cnb == 1 ? (p.cH / p.cth) * p.cws : 1;  // pixel blocks (M tiles or wgrad k-blocks) per image
    constexpr int BH = BN / 2;                          // rows of B this CTA fetches for the cluster (K-major B)
    constexpr int NCH = BN / 64, NCH_CL = NCH / (CLU ? 2 : 1);     // 64-wide chunks of an MN-major B tile: all / per CTA
    const int j0 = crank * NCH_CL;
    bool first_tile = true;
    for (int item = unit0; item < n_items; item += unit_step) {
      WorkItem w = decode_item(p, item, mtd);
      if (CLU) w.m_tile = 2 * w.m_tile + crank;
      const int n_off = w.n_tile * BN, m_off = w.m_tile * BM;
      const int ab0 = p.a_batched ? w.batch % p.a_nb0 : 0, ab1 = p.a_batched ? w.batch / p.a_nb0 : 0;
      const int bb0 = p.b_batched ? w.batch % p.b_nb0 : 0, bb1 = p.b_batched ? w.batch / p.b_nb0 : 0;
      // conv: pixel-block origin of this M tile; (tap, channel block) of the first k-block
      int cn0 = 0, ch0 = 0, cw0 = 0, tap = 0, cb = 0;
      // wgrad: (image, row) origin of the first 64-pixel k-block
      int wn = 0, wh = 0;
      int dh = 0, dw = 0, dn = 0, tw = 0;
      if (p.kind == KIND_CONV) {
        if (p.cnb == 1) {
          const int t = w.m_tile % bpi;
          cn0 = w.m_tile / bpi;
          ch0 = (t / p.cws) * p.cth;
          cw0 = (t % p.cws) * p.cwseg;
        } else {
          cn0 = w.m_tile * p.cnb;
        }
        tap = w.kb0 / p.cblks;
        cb = w.kb0 % p.cblks;
      } else if (p.kind == KIND_CONV_WGRAD) {
        tap = w.batch;
        if (p.cnb == 1) {
          wn = w.kb0 / bpi;
          wh = (w.kb0 % bpi) * p.cth;
        } else {
          wn = w.kb0 * p.cnb;
        }
      }
      if (p.kind != KIND_PLAIN) {
        dh = p.tap_dh[tap];
        dw = p.tap_dw[tap];
        dn = p.tap_dn[tap];
        tw = p.tap_w[tap];
      }
      int kcol = w.kb0 * 64;
      for (int i = 0; i < w.nkb; ++i) {
        mbar_wait(&empty_bar[s], ph ^ 1u);
        if (elect_one()) {
          uint8_t* a_dst = smem + (size_t)s * STAGE_B;
          uint8_t* b_dst = a_dst + A_BYTES;
          uint64_t* fb = &full_bar[s];
          const bool load_b = CLU || !p.b_resident || first_tile;  // B-resident: the weight tile is already in this stage
          if (PAIR) {
            // both CTAs' bytes of this stage are counted on the even CTA's barrier, which its MMA warp waits on
            if (crank == 0) mbar_arrive_expect_tx(fb, 2 * STAGE_B);
            if (p.kind == KIND_CONV) tma_load_4d_cg2(a_dst, &tmA, fb, cb * 64, cw0 + dw, ch0 + dh, cn0 + dn);
            else if (!A_MN) tma_load_4d_cg2(a_dst, &tmA, fb, kcol, m_off, ab0, ab1);
            else {
#pragma unroll
              for (int j = 0; j < 2; ++j) tma_load_4d_cg2(a_dst + j * CHUNK_BYTES, &tmA, fb, m_off + 64 * j, kcol, ab0, ab1);
            }
            // this CTA's half of the B tile: rows [crank * BN/2, +BN/2) of a K-major tile, or its NCH/2 64-wide chunks
            if (!B_MN) {
              if (p.kind == KIND_CONV) tma_load_4d_cg2(b_dst, &tmB, fb, cb * 64, n_off + crank * BH, tw, 0);
              else tma_load_4d_cg2(b_dst, &tmB, fb, kcol, n_off + crank * BH, bb0, bb1);
            } else {
#pragma unroll
              for (int j = 0; j < NCH_CL; ++j) {
                if (p.kind == KIND_CONV) tma_load_4d_cg2(b_dst + j * CHUNK_BYTES, &tmB, fb, n_off + 64 * (j0 + j), cb * 64, tw, 0);
                else if (p.kind == KIND_PLAIN) tma_load_4d_cg2(b_dst + j * CHUNK_BYTES, &tmB, fb, n_off + 64 * (j0 + j), kcol, bb0, bb1);
                else tma_load_4d_cg2(b_dst + j * CHUNK_BYTES, &tmB, fb, n_off + 64 * (j0 + j), dw, wh + dh, wn + dn);
              }
            }
          } else {
          mbar_arrive_expect_tx(fb, load_b ? Cfg::STAGE_BYTES : A_BYTES);
          if (p.kind == KIND_CONV) {
            tma_load_4d(a_dst, &tmA, fb, cb * 64, cw0 + dw, ch0 + dh, cn0 + dn);
            if (CL == 2) {  // this CTA's half of the B tile, multicast to both CTAs of the cluster
              if (!B_MN) {
                tma_load_4d_mc(b_dst + crank * (BH * 128), &tmB, fb, cb * 64, n_off + crank * BH, tw, 0, 3);
              } else {
#pragma unroll
                for (int j = 0; j < NCH_CL; ++j)
                  tma_load_4d_mc(b_dst + (j0 + j) * CHUNK_BYTES, &tmB, fb, n_off + 64 * (j0 + j), cb * 64, tw, 0, 3);
              }
            } else if (!B_MN) {
              tma_load_4d(b_dst, &tmB, fb, cb * 64, n_off, tw, 0);
            } else {
#pragma unroll
              for (int j = 0; j < BN / 64; ++j) tma_load_4d(b_dst + j * CHUNK_BYTES, &tmB, fb, n_off + 64 * j, cb * 64, tw, 0);
            }
          } else {
            if (!A_MN) {
              tma_load_4d(a_dst, &tmA, fb, kcol, m_off, ab0, ab1);
            } else {
#pragma unroll
              for (int j = 0; j < 2; ++j) tma_load_4d(a_dst + j * CHUNK_BYTES, &tmA, fb, m_off + 64 * j, kcol, ab0, ab1);
            }
            if (p.kind == KIND_PLAIN && load_b) {
              if (CL == 2) {
                if (!B_MN) {
                  tma_load_4d_mc(b_dst + crank * (BH * 128), &tmB, fb, kcol, n_off + crank * BH, bb0, bb1, 3);
                } else {
#pragma unroll
                  for (int j = 0; j < NCH_CL; ++j)
                    tma_load_4d_mc(b_dst + (j0 + j) * CHUNK_BYTES, &tmB, fb, n_off + 64 * (j0 + j), kcol, bb0, bb1, 3);
                }
              } else if (!B_MN) {
                tma_load_4d(b_dst, &tmB, fb, kcol, n_off, bb0, bb1);
              } else {
#pragma unroll
                for (int j = 0; j < BN / 64; ++j) tma_load_4d(b_dst + j * CHUNK_BYTES, &tmB, fb, n_off + 64 * j, kcol, bb0, bb1);
              }
            } else if (p.kind != KIND_PLAIN && B_MN) {  // KIND_CONV_WGRAD: B = activations shifted by the tap, k-block = 64 pixels
              if (CL == 2) {
#pragma unroll
                for (int j = 0; j < NCH_CL; ++j)
                  tma_load_4d_mc(b_dst + (j0 + j) * CHUNK_BYTES, &tmB, fb, n_off + 64 * (j0 + j), dw, wh + dh, wn + dn, 3);
              } else {
#pragma unroll
                for (int j = 0; j < BN / 64; ++j)
                  tma_load_4d(b_dst + j * CHUNK_BYTES, &tmB, fb, n_off + 64 * j, dw, wh + dh, wn + dn);
              }
            }
          }
          }  // !PAIR
        }
        __syncwarp();
        // advance the k-block coordinates (warp-uniform)
        kcol += 64;
        if (p.kind == KIND_CONV) {
          if (++cb == p.cblks) {
            cb = 0;
            ++tap;
            if (i + 1 < w.nkb) {
              dh = p.tap_dh[tap];
              dw = p.tap_dw[tap];
              dn = p.tap_dn[tap];
              tw = p.tap_w[tap];
            }
          }
        } else if (p.kind == KIND_CONV_WGRAD) {
          if (p.cnb == 1) {
            wh += p.cth;
            if (wh == p.cH) {
              wh = 0;
              ++wn;
            }
          } else {
            wn += p.cnb;
          }
        }
        if (++s == stages) {
          s = 0;
          ph ^= 1u;
        }
      }
      first_tile = false;
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    // Warp-uniform loop; one elected lane (always the same one) issues the four K=16 MMAs of a stage and the commit.
    // The 64-bit smem descriptors are built once; per stage / per K step only their 14-bit address field moves
    // ((byte offset) >> 4 added to the low word), so a k-block costs a handful of instructions.
    constexpr uint32_t idesc = umma_idesc_bf16(PAIR ? 2 * BM : BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
    const uint32_t smem_a0 = smem_u32(smem);
    const uint64_t adesc0 = A_MN ? umma_desc_sw128(smem_a0, CHUNK_BYTES, 1024) : umma_desc_sw128(smem_a0, 16, 1024);
    const uint64_t bdesc0 = B_MN ? umma_desc_sw128(smem_a0 + A_BYTES, CHUNK_BYTES, 1024)
                                 : umma_desc_sw128(smem_a0 + A_BYTES, 16, 1024);
    constexpr uint64_t A_KSTEP = (A_MN ? 2048 : 32) >> 4, B_KSTEP = (B_MN ? 2048 : 32) >> 4;
    constexpr uint64_t STAGE_STEP = STAGE_B >> 4;
    int s = 0;
    uint32_t ph = 0;
    uint64_t ad = adesc0, bd = bdesc0;
    int it = 0;
    const int mma_items = (PAIR && crank != 0) ? 0 : n_items;  // cta_group::2: only the even CTA issues
    for (int item = unit0; item < mma_items; item += unit_step, ++it) {
      const WorkItem w = decode_item(p, item, mtd);
      const int acc = it & 1;
      const uint32_t acc_ph = (uint32_t)(it >> 1) & 1u;
      mbar_wait(&tmem_empty_bar[acc], acc_ph ^ 1u);  // epilogue has drained this accumulator buffer
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * Cfg::ACC_STRIDE);
      for (int i = 0; i < w.nkb; ++i) {
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        if (elect_one()) {
          if (PAIR) {
            tc_mma_bf16_cg2(d_tmem, ad, bd, idesc, i != 0 ? 1u : 0u);
            tc_mma_bf16_cg2(d_tmem, ad + A_KSTEP, bd + B_KSTEP, idesc, 1u);
            tc_mma_bf16_cg2(d_tmem, ad + 2 * A_KSTEP, bd + 2 * B_KSTEP, idesc, 1u);
            tc_mma_bf16_cg2(d_tmem, ad + 3 * A_KSTEP, bd + 3 * B_KSTEP, idesc, 1u);
            tc_commit_cg2_mc(&empty_bar[s], 3);                                // both CTAs' producers may refill the stage
            if (i == w.nkb - 1) tc_commit_cg2_mc(&tmem_full_bar[acc], 3);      // both CTAs' epilogues may read their rows
          } else {
          tc_mma_bf16(d_tmem, ad, bd, idesc, i != 0 ? 1u : 0u);
          tc_mma_bf16(d_tmem, ad + A_KSTEP, bd + B_KSTEP, idesc, 1u);
          tc_mma_bf16(d_tmem, ad + 2 * A_KSTEP, bd + 2 * B_KSTEP, idesc, 1u);
          tc_mma_bf16(d_tmem, ad + 3 * A_KSTEP, bd + 3 * B_KSTEP, idesc, 1u);
          if (CL == 2) tc_commit_mc(&empty_bar[s], 3);  // the stage is shared: both CTAs' producers wait for both MMA warps
          else tc_commit(&empty_bar[s]);  // frees this smem stage once the MMAs above have consumed it
          if (i == w.nkb - 1) tc_commit(&tmem_full_bar[acc]);  // accumulator complete
          }
        }
        __syncwarp();
        ad += STAGE_STEP;
        bd += STAGE_STEP;
        if (++s == stages) {
          s = 0;
          ph ^= 1u;
          ad = adesc0;
          bd = bdesc0;
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (4 warps = 4 TMEM lane quadrants)
    // Global operands of the epilogue never sit on the critical path: the tile's bias lives in registers (8 columns
    // per lane, broadcast by shuffles), per-image bias and residual of 32-column chunk c+1 are fetched while chunk c
    // is processed, and chunk 0's are issued before the wait on the accumulator.
    const int q = warp & 3;
    const int hh = (warp - 2) >> 2;
    uint8_t* stg = stg_base + (size_t)(warp - 2) * STG_BYTES;
    constexpr int NCH_ALL = BN / 32;
    constexpr int HALF = (Cfg::OUT_CH == 64) ? NCH_ALL / 2 : (NCH_ALL + 1) / 2;
    EpiState st;
    st.c0 = hh * HALF;
    st.c1 = hh == 0 ? HALF : NCH_ALL;
    int it = 0;
    const bool raw = p.out_mode == OUT_F32_PARTIAL;
    for (int item = unit0; item < n_items; item += unit_step, ++it) {
      WorkItem w = decode_item(p, item, mtd);
      if (CLU) w.m_tile = 2 * w.m_tile + crank;
      const int acc = it & 1;
      const uint32_t acc_ph = (uint32_t)(it >> 1) & 1u;
      EpiTile t;
      t.row0 = w.m_tile * BM + q * 32;
      const long long row = t.row0 + lane;
      const bool row_ok = row < p.M;
      long long res_off = 0;
      if (raw) {
        t.o2 = w.batch * p.splits + w.split;
        t.o3 = 0;
      } else {
        t.o2 = w.batch % p.out_nb0;
        t.o3 = w.batch / p.out_nb0;
        res_off = (long long)t.o2 * p.out_bs0 + (long long)t.o3 * p.out_bs1;
      }
      t.n_tile0 = w.n_tile * BN;
      t.rb = (!raw && p.rowbias != nullptr && row_ok) ? p.rowbias + (row / p.rows_per_group) * p.ld_rowbias : nullptr;
      t.rs = (!raw && p.residual != nullptr && row_ok) ? p.residual + res_off + row * p.ldr : nullptr;
      t.has_bias = !raw && p.bias != nullptr;
      t.alpha = raw ? 1.f : p.alpha;
      t.N = p.N;
      t.M = p.M;
      t.out_mode = p.out_mode;
      t.gn_part = p.gn_part;
      t.gn_slab = p.gn_slab;
      t.row = row;
      t.row_ok = row_ok;
#pragma unroll
      for (int e = 0; e < 8; ++e) t.bv[e] = 0.f;
      if (t.has_bias && lane * 8 < BN && t.n_tile0 + lane * 8 < p.N) {
        const float4 b0 = *reinterpret_cast<const float4*>(p.bias + t.n_tile0 + lane * 8);
        const float4 b1 = *reinterpret_cast<const float4*>(p.bias + t.n_tile0 + lane * 8 + 4);
        t.bv[0] = b0.x; t.bv[1] = b0.y; t.bv[2] = b0.z; t.bv[3] = b0.w;
        t.bv[4] = b1.x; t.bv[5] = b1.y; t.bv[6] = b1.z; t.bv[7] = b1.w;
      }
      EpiPre pa, pb;
      epi_prefetch(t, st.c0, pa);
      mbar_wait(&tmem_full_bar[acc], acc_ph);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * Cfg::ACC_STRIDE);
#pragma unroll 1
      for (int c = st.c0; c < st.c1; c += 2) {
        epi_chunk<BN>(t, st, c, pa, pb, t_addr, stg, lane, &tmO, &tmem_empty_bar[acc], PAIR, p);
        if (c + 1 < st.c1) epi_chunk<BN>(t, st, c + 1, pb, pa, t_addr, stg, lane, &tmO, &tmem_empty_bar[acc], PAIR, p);
      }
    }
    if (lane == 0) bulk_wait_read<0>();  // staging smem no longer read by the copy engine; the writes complete with the grid
    __syncwarp();
  }
  tc_fence_before();
  if (CLU) cluster_sync_all();  // no CTA leaves while its peer may still multicast into it or arrive on its barriers
  else __syncthreads();
  if (warp == 1) {
    if (PAIR) tmem_dealloc_cg2(tmem_base, Cfg::TMEM_COLS);
    else tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// out[m][n] = bf16/f32( alpha * sum_s ws[s][m][n] + bias[n] + rowbias[m/rpg][n] + residual[m][n] )
__global__ void splitk_finalize_kernel(const float* __restrict__ ws, int splits, long long M, int N, float alpha,
                                       const float* __restrict__ bias, const float* __restrict__ rowbias,
                                       int rows_per_group, long long ld_rowbias, const bf16* __restrict__ residual,
                                       long long ldr, void* __restrict__ out, long long ldo, int out_f32) {
  pdl_grid_sync();
  const long long nvec = M * (N / 8);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    const long long m = i / (N / 8);
    const int n = (int)(i % (N / 8)) * 8;
    float v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int s = 0; s < splits; ++s) {
      const float* src = ws + ((long long)s * M + m) * N + n;
      const float4 a = *reinterpret_cast<const float4*>(src), b = *reinterpret_cast<const float4*>(src + 4);
      v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w; v[4] += b.x; v[5] += b.y; v[6] += b.z; v[7] += b.w;
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] *= alpha;
    if (bias) {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] += bias[n + e];
    }
    if (rowbias) {
      const float* rb = rowbias + (m / rows_per_group) * ld_rowbias + n;
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] += rb[e];
    }
    if (residual) {
      const uint4 rr = *reinterpret_cast<const uint4*>(residual + m * ldr + n);
      const float2 r0 = unpack_bf16x2(rr.x), r1 = unpack_bf16x2(rr.y), r2 = unpack_bf16x2(rr.z), r3 = unpack_bf16x2(rr.w);
      v[0] += r0.x; v[1] += r0.y; v[2] += r1.x; v[3] += r1.y; v[4] += r2.x; v[5] += r2.y; v[6] += r3.x; v[7] += r3.y;
    }
    if (out_f32) {
      float* o = reinterpret_cast<float*>(out) + m * ldo + n;
      *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(o + 4) = make_float4(v[4], v[5], v[6], v[7]);
    } else {
      uint4 o;
      o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]); o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
      *reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(out) + m * ldo + n) = o;
    }
  }
}

// ---------------------------------------------------------------------------------------------- host side
template <int BN, bool A_MN, bool B_MN, int CL>
static cudaError_t launch_one(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmO, const GemmKParams& p,
                              int grid, int stages, cudaStream_t stream) {
  using Cfg = TileCfg<BN>;
  const size_t stage_bytes = CL == 3 ? A_BYTES + Cfg::B_BYTES / 2 : Cfg::STAGE_BYTES;
  const size_t smem = (size_t)stages * stage_bytes + STG_TOTAL + (2 * stages + 4) * 8 + 16 + 1024;
  auto kern = gemm_tc_kernel<BN, A_MN, B_MN, CL>;
  static bool attr_set = false;  // per instantiation
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  if (CL == 1) return launch_k(kern, dim3(grid), dim3(GEMM_THREADS), smem, stream, tmA, tmB, tmO, p, stages);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmO, p, stages);
}

// Should this launch run as 2-CTA clusters (B tile shared through TMA multicast)?  Measured per shape at the bench geometry
// (tools/gemm_shapes.py, SD2_GEMM_CLUSTER=0 vs 2, profiles/r02_gemm_cluster_ab.md): the implicit-GEMM convolutions (forward
// and dgrad) gain up to 17 %, the weight gradients lose 10-80 % (their M dimension is Cout: 3 or 5 tiles pair badly, and
// both operands are streamed), the linears are a wash.  Default policy: convolution forward / dgrad only, even M-tile
// count or many tiles.  SD2_GEMM_CLUSTER=0 disables clusters, =2 enables them wherever the kernel supports it, =3 runs the
// cta_group::2 pair mode (one 256-row MMA per SM pair) wherever it is instantiated.  Returns 0 / 2 / 3 (the kernel's CL).
int gemm_use_cluster(const GemmKParams& p, int BN, bool a_mn, bool b_mn, int num_sms) {
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("SD2_GEMM_CLUSTER");
    mode = e ? atoi(e) : 1;
  }
  if (mode == 0) return 0;
  if (b_mn && BN < 128) return 0;  // an MN-major B tile needs two 64-wide chunks to split
  if (p.mt < 2) return 0;
  const long long pair_items = (long long)((p.mt + 1) / 2) * p.nt * p.splits * p.batches;
  if (pair_items < num_sms / 2) return 0;
  const bool pair_ok = BN == 256 || (BN == 160 && !b_mn);  // the instantiated cta_group::2 kernels
  if (mode == 2) return 2;
  if (mode == 3) return pair_ok ? 3 : 0;  // cta_group::2 wherever it is instantiated (A/B runs)
  // cta_group::2 (profiles/r02_gemm_pair_ab.md): K-major linears with a long contraction (K >= 1024) and M >= 16384 gain 4-10 %;
  // MN-major operands (every weight gradient, the dgrads) and the K = 320 layers lose, the convolutions are +-3 % around the
  // multicast variant
  static const int pair_rule = getenv("SD2_GEMM_PAIR") ? atoi(getenv("SD2_GEMM_PAIR")) : 1;  // 0: A/B switch for this rule
  if (pair_rule && pair_ok && p.kind == KIND_PLAIN && !a_mn && !b_mn && p.total_kb >= 16 && p.mt >= 128 && p.splits == 1 && p.batches == 1) return 3;
  return (p.kind == KIND_CONV && (p.mt % 2 == 0 || p.mt >= 32)) ? 2 : 0;
}

cudaError_t launch_gemm_tc(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmO, const GemmKParams& p,
                           int BN, bool a_mn, bool b_mn, int cluster, int num_sms, cudaStream_t stream) {
  const int st = gemm_stages(BN);
  if (cluster == 3) {
    const long long items = (long long)((p.mt + 1) / 2) * p.nt * p.splits * p.batches;
    const int ncl = (int)(items < num_sms / 2 ? items : num_sms / 2);
    const int grid = 2 * ncl, stp = gemm_stages_pair(BN);
#define SD2_GEMM_CASE(bn, amn, bmn) \
  if (BN == bn && a_mn == amn && b_mn == bmn) return launch_one<bn, amn, bmn, 3>(tmA, tmB, tmO, p, grid, stp, stream);
    SD2_GEMM_CASE(256, false, false) SD2_GEMM_CASE(160, false, false)
    SD2_GEMM_CASE(256, false, true)
    SD2_GEMM_CASE(256, true, false) SD2_GEMM_CASE(160, true, false)
    SD2_GEMM_CASE(256, true, true)
#undef SD2_GEMM_CASE
    return cudaErrorInvalidValue;
  }
  if (cluster) {
    const long long items = (long long)((p.mt + 1) / 2) * p.nt * p.splits * p.batches;
    const int ncl = (int)(items < num_sms / 2 ? items : num_sms / 2);
    const int grid = 2 * ncl;
#define SD2_GEMM_CASE(bn, amn, bmn) \
  if (BN == bn && a_mn == amn && b_mn == bmn) return launch_one<bn, amn, bmn, 2>(tmA, tmB, tmO, p, grid, st, stream);
    SD2_GEMM_CASE(256, false, false) SD2_GEMM_CASE(160, false, false) SD2_GEMM_CASE(128, false, false)
    SD2_GEMM_CASE(64, false, false)
    SD2_GEMM_CASE(256, false, true) SD2_GEMM_CASE(128, false, true)
    SD2_GEMM_CASE(256, true, false) SD2_GEMM_CASE(160, true, false) SD2_GEMM_CASE(128, true, false)
    SD2_GEMM_CASE(64, true, false)
    SD2_GEMM_CASE(256, true, true) SD2_GEMM_CASE(128, true, true)
#undef SD2_GEMM_CASE
    return cudaErrorInvalidValue;
  }
  const long long items = (long long)p.mt * p.nt * p.splits * p.batches;
  int grid = (int)(items < num_sms ? items : num_sms);
  int st_run = st;
  GemmKParams pr = p;
  {  // B-resident mode, see GemmKParams::b_resident.  SD2_GEMM_BRES=0 disables it (A/B measurements).
    static int bres = -1;
    if (bres < 0) {
      const char* e = getenv("SD2_GEMM_BRES");
      bres = e ? atoi(e) : 1;
    }
    const int g_nt = p.nt > 0 ? (num_sms / p.nt) * p.nt : 0;  // a grid that is a multiple of nt keeps n_tile fixed per CTA
    if (bres && p.kind == KIND_PLAIN && p.total_kb <= st && p.total_kb >= 3 && p.splits == 1 && p.batches == 1 && p.raster == 1 &&
        g_nt > 0 && items >= 2LL * g_nt && g_nt * 100 >= num_sms * 94) {
      pr.b_resident = 1;
      grid = g_nt;
      st_run = p.total_kb;  // ring length = k-blocks per tile: stage s <-> k-block s
    }
  }
#define SD2_GEMM_CASE(bn, amn, bmn) \
  if (BN == bn && a_mn == amn && b_mn == bmn) return launch_one<bn, amn, bmn, 1>(tmA, tmB, tmO, pr, grid, st_run, stream);
  SD2_GEMM_CASE(256, false, false) SD2_GEMM_CASE(160, false, false) SD2_GEMM_CASE(128, false, false)
  SD2_GEMM_CASE(64, false, false)
  SD2_GEMM_CASE(256, false, true) SD2_GEMM_CASE(128, false, true) SD2_GEMM_CASE(64, false, true)
  SD2_GEMM_CASE(256, true, false) SD2_GEMM_CASE(160, true, false) SD2_GEMM_CASE(128, true, false)
  SD2_GEMM_CASE(64, true, false)
  SD2_GEMM_CASE(256, true, true) SD2_GEMM_CASE(128, true, true) SD2_GEMM_CASE(64, true, true)
#undef SD2_GEMM_CASE
  return cudaErrorInvalidValue;
}

cudaError_t launch_splitk_finalize(const float* ws, int splits, long long M, int N, float alpha, const float* bias,
                                   const float* rowbias, int rows_per_group, long long ld_rowbias, const bf16* residual,
                                   long long ldr, void* out, long long ldo, int out_f32, cudaStream_t stream) {
  const long long nvec = M * (N / 8);
  int blocks = (int)((nvec + 255) / 256);
  if (blocks > 148 * 24) blocks = 148 * 24;
  if (blocks < 1) blocks = 1;
  return launch_k(splitk_finalize_kernel, dim3(blocks), dim3(256), 0, stream, ws, splits, M, N, alpha, bias, rowbias,
                  rows_per_group, ld_rowbias, residual, ldr, out, ldo, out_f32);
}

}  // namespace sd2

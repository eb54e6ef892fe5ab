// C-ABI entry points: context lifecycle, TMA tensor-map construction and the sd2_gemm host-side planner
// (tile shape, split-K, tensor maps) in front of gemm_tc.cu.  See include/sd2b200.h for the contract.
#include <algorithm>
#include <mutex>

#include "gemm_tc.cuh"
#include "host.h"

namespace sd2 {

EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

bool encode_tmap_4d(CUtensorMap* out, int dtype, int swizzle_bytes, const void* ptr, const uint64_t dims[4],
                    const uint64_t strides_bytes[3], const uint32_t box[4], std::string* err) {
  EncodeTiledFn fn = get_encode_tiled();
  if (!fn) {
    if (err) *err = "cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)";
    return false;
  }
  cuuint64_t gdim[4], gstr[3];
  cuuint32_t bx[4], es[4] = {1, 1, 1, 1};
  for (int i = 0; i < 4; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
  }
  for (int i = 0; i < 3; ++i) gstr[i] = strides_bytes[i];
  const CUtensorMapDataType dt = dtype == SD2_DT_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                 : dtype == SD2_DT_U64_INTERNAL ? CU_TENSOR_MAP_DATA_TYPE_UINT64 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  const CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = fn(out, dt, 4, const_cast<void*>(ptr), gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    if (err) {
      char buf[256];
      snprintf(buf, sizeof(buf),
               "cuTensorMapEncodeTiled failed (%d): ptr=%p dims=(%llu,%llu,%llu,%llu) strides=(%llu,%llu,%llu) box=(%u,%u,%u,%u)",
               (int)r, ptr, (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2],
               (unsigned long long)dims[3], (unsigned long long)strides_bytes[0], (unsigned long long)strides_bytes[1],
               (unsigned long long)strides_bytes[2], box[0], box[1], box[2], box[3]);
      *err = buf;
    }
    return false;
  }
  return true;
}

bool encode_tmap_bf16_4d(CUtensorMap* out, const void* ptr, const uint64_t dims[4], const uint64_t strides_bytes[3],
                         const uint32_t box[4], std::string* err) {
  return encode_tmap_4d(out, SD2_DT_BF16, 128, ptr, dims, strides_bytes, box, err);
}

// output tensor map of the GEMM epilogue: dims (N, M, nb0, nb1), box (chunk columns, 32 rows); the inner box row is
// 128 B (SWIZZLE_128B: 64 bf16 / 32 fp32) or 64 B (SWIZZLE_64B: 32 bf16)
bool out_tmap(CUtensorMap* tm, void* ptr, bool f32, int chunk_cols, int N, int M, long long ldo, long long nb0,
                     long long nb1, long long bs0, long long bs1, std::string* err) {
  const uint64_t es = f32 ? 4 : 2;
  if (nb0 < 1) nb0 = 1;
  if (nb1 < 1) nb1 = 1;
  const uint64_t s1 = (uint64_t)ldo * es;
  const uint64_t s2 = nb0 > 1 ? (uint64_t)bs0 * es : s1 * (uint64_t)M;
  const uint64_t s3 = nb1 > 1 ? (uint64_t)bs1 * es : (nb0 > 1 ? s2 * (uint64_t)nb0 : s2);
  const uint64_t dims[4] = {(uint64_t)N, (uint64_t)M, (uint64_t)nb0, (uint64_t)nb1};
  const uint64_t strides[3] = {s1, s2, s3};
  const uint32_t box[4] = {(uint32_t)chunk_cols, 32, 1, 1};
  return encode_tmap_4d(tm, f32 ? SD2_DT_F32 : SD2_DT_BF16, (int)(chunk_cols * es), ptr, dims, strides, box, err);
}

// tensor map of a plain (possibly batched) operand; K-major box = (64, box_rows), MN-major box = (64, 64)
bool plain_tmap(CUtensorMap* tm, const sd2_operand& o, int box_rows, std::string* err) {
  const uint64_t nb0 = o.nb0 > 0 ? (uint64_t)o.nb0 : 1, nb1 = o.nb0 > 0 && o.nb1 > 0 ? (uint64_t)o.nb1 : 1;
  const uint64_t dims[4] = {(uint64_t)o.cols, (uint64_t)o.rows, nb0, nb1};
  // strides must be non-zero multiples of 16 B even for extent-1 dims
  const uint64_t s1 = (uint64_t)o.ld * 2;
  const uint64_t s2 = nb0 > 1 ? (uint64_t)o.bs0 * 2 : s1 * (uint64_t)o.rows;
  const uint64_t s3 = nb1 > 1 ? (uint64_t)o.bs1 * 2 : (nb0 > 1 ? s2 * nb0 : s2);
  const uint64_t strides[3] = {s1, s2, s3};
  const uint32_t box[4] = {64, (uint32_t)(o.mn_major ? 64 : box_rows), 1, 1};
  return encode_tmap_bf16_4d(tm, o.ptr, dims, strides, box, err);
}

// pixel-box geometry of a shifted NHWC operand: `pixels` (128 for an M tile, 64 for a wgrad k-block) = W * th * nb
static bool conv_box(const sd2_conv_geom& g, int pixels, int* th, int* nb, int* wseg = nullptr) {
  if (g.W <= 0 || g.H <= 0) return false;
  if (wseg) *wseg = g.W;
  if (wseg && g.W > pixels && g.W % pixels == 0) {  // wide image: the tile is a `pixels`-wide segment of one row
    *th = 1;
    *nb = 1;
    *wseg = pixels;
    return true;
  }
  if (pixels % g.W != 0) return false;
  int rows = pixels / g.W;
  if (rows <= g.H) {
    if (g.H % rows != 0) return false;
    *th = rows;
    *nb = 1;
  } else {
    if (rows % g.H != 0) return false;
    *th = g.H;
    *nb = rows / g.H;
  }
  return *th <= 256 && *nb <= 256 && g.W <= 256;
}

static bool conv_tmap(CUtensorMap* tm, const sd2_conv_geom& g, int th, int nb, std::string* err, int wseg = 0) {
  const uint64_t dims[4] = {(uint64_t)g.C, (uint64_t)g.W, (uint64_t)g.H, (uint64_t)g.n_planes};
  const uint64_t strides[3] = {(uint64_t)g.ldc * 2, (uint64_t)g.ldc * 2 * g.W, (uint64_t)g.ldc * 2 * g.W * g.H};
  const uint32_t box[4] = {64, (uint32_t)(wseg > 0 ? wseg : g.W), (uint32_t)th, (uint32_t)nb};
  return encode_tmap_bf16_4d(tm, g.ptr, dims, strides, box, err);
}

// Tile width and K-split of one GEMM launch, from a small cycle model of the persistent kernel:
//   cycles = waves(items over the SMs) x (k-blocks per item x c_kb(BN) + epilogue(BN)) [+ finalize pass]
// c_kb = the slower of the tensor pipe (2*BN cycles per 64-deep k-block of a 128 x BN tile) and the operand fetch of
// one stage.  Narrow tiles are chosen only when the problem cannot fill the 148 SMs with wide ones.
static void plan_gemm(const sd2_ctx* ctx, const sd2_gemm_desc* d, int N8, int total_kb, int batches, bool b_mn,
                      bool direct_store, int* BN_out, int* splits_out) {
  const int cands[4] = {256, 160, 128, 64};
  const long long mt = (d->M + 127) / 128;
  double best = 1e30;
  int best_bn = 64, best_s = 1;
  for (int ci = 0; ci < 4; ++ci) {
    const int bn = cands[ci];
    if (b_mn && bn % 64 != 0) continue;
    const long long nt = (N8 + bn - 1) / bn;
    const long long tiles = mt * nt * batches;
    const double c_kb = std::max(2.0 * bn, 180.0 + 1.3 * bn);
    const double epi = 600.0 + (bn / 32) * 200.0;
    long long smax = total_kb / 2;
    if (smax > 32) smax = 32;
    if (d->max_splits > 0 && smax > d->max_splits) smax = d->max_splits;
    if (direct_store) {
      if (batches != 1 || d->workspace == nullptr) smax = 1;
      else {
        const long long per_split = (long long)d->M * N8 * 4;
        if (per_split > 0 && d->workspace_bytes / per_split < smax) smax = d->workspace_bytes / per_split;
      }
    }
    if (smax < 1) smax = 1;
    for (long long s = 1; s <= smax; ++s) {
      const long long waves = (tiles * s + ctx->num_sms - 1) / ctx->num_sms;
      const long long kb = (total_kb + s - 1) / s;
      const double cost = (double)waves * ((double)kb * c_kb + epi) + ((s > 1 && direct_store) ? 7000.0 : 0.0);
      if (cost < best * 0.98) {  // candidates are visited wide-to-narrow, few-to-many splits: keep those on ties
        best = cost;
        best_bn = bn;
        best_s = (int)s;
      }
    }
  }
  *BN_out = best_bn;
  *splits_out = best_s;
}

}  // namespace sd2

using namespace sd2;

extern "C" {

int sd2_version(void) { return SD2_VERSION; }

int sd2_ctx_create(int device, sd2_ctx** out) {
  if (!out) return 1;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return 2;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return 3;
  if (prop.major != 10) return 4;  // sm_100a only: no fallback paths
  sd2_ctx* c = new sd2_ctx();
  c->device = device;
  c->num_sms = prop.multiProcessorCount;
  *out = c;
  return 0;
}

int sd2_ctx_destroy(sd2_ctx* ctx) {
  if (ctx) sd2_ddp_destroy(ctx);
  delete ctx;
  return 0;
}

const char* sd2_last_error(sd2_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
int sd2_num_sms(sd2_ctx* ctx) { return ctx ? ctx->num_sms : 0; }
long long sd2_launch_count(sd2_ctx* ctx) { return ctx ? ctx->launches : 0; }

int sd2_gemm(sd2_ctx* ctx, const sd2_gemm_desc* d, sd2_stream stream_) {
  if (!ctx || !d) return 1;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (d->M <= 0 || d->N <= 0 || d->K <= 0) return fail(ctx, "sd2_gemm: empty problem");

  GemmKParams p;
  memset(&p, 0, sizeof(p));
  p.kind = d->kind;
  p.M = d->M;
  p.N = (d->N + 7) & ~7;  // stores are 8-wide; the caller's ldo covers the round-up (columns beyond N get zeros)
  if (p.N > d->ldo) return fail(ctx, "sd2_gemm: ldo < round8(N)");
  p.alpha = d->alpha;
  p.bias = d->bias;
  p.rowbias = d->rowbias;
  p.rows_per_group = d->rows_per_group > 0 ? d->rows_per_group : 1;
  p.ld_rowbias = d->ld_rowbias;
  p.residual = reinterpret_cast<const bf16*>(d->residual);
  p.ldr = d->ldr;
  p.out_nb0 = d->out_nb0 > 0 ? d->out_nb0 : 1;
  p.out_bs0 = d->out_bs0;
  p.out_bs1 = d->out_bs1;
  const bool mse = d->mse_target != nullptr;
  if (mse) {
    if (d->kind != SD2_GEMM_CONV || d->out_mode != SD2_OUT_BF16 || d->N > 8 || d->N < 4)
      return fail(ctx, "sd2_gemm: the MSE head belongs to a bf16 convolution with 4..8 output columns (conv_out)");
    if (!d->mse_dpred8 || !d->mse_acc || d->mse_hw < 1 || d->M % d->mse_hw != 0)
      return fail(ctx, "sd2_gemm: MSE head needs mse_dpred8, mse_acc and mse_hw dividing M");
    if (d->mse_dtype != SD2_DT_F32 && d->mse_dtype != SD2_DT_BF16 && d->mse_dtype != SD2_DT_F16)
      return fail(ctx, "sd2_gemm: mse_dtype");
    p.mse_target = d->mse_target;
    p.mse_dpred8 = reinterpret_cast<bf16*>(d->mse_dpred8);
    p.mse_acc = d->mse_acc;
    p.mse_dtype = d->mse_dtype;
    p.mse_hw = d->mse_hw;
    p.mse_k = 2.f / ((float)d->M * 4.f);
  }
  const bool gn = d->gn_partial != nullptr;
  if (gn) {
    if (d->out_mode != SD2_OUT_BF16) return fail(ctx, "sd2_gemm: gn_partial needs a bf16 output");
    if (d->gn_slab != 16 && d->gn_slab != 32) return fail(ctx, "sd2_gemm: gn_slab must be 16 or 32");
    if (d->kind == SD2_GEMM_CONV_WGRAD || (d->kind == SD2_GEMM_PLAIN && d->batch > 1))
      return fail(ctx, "sd2_gemm: gn_partial is for single-batch forward GEMMs / convolutions");
    if (d->M % d->gn_slab != 0) return fail(ctx, "sd2_gemm: gn_slab must divide M");
    p.gn_part = d->gn_partial;
    p.gn_slab = d->gn_slab;
  }

  bool a_mn, b_mn;
  int batches = 1;
  CUtensorMap tmA, tmB;
  std::string err;
  int BN, splits = 1;
  const bool direct_store = d->out_mode == SD2_OUT_BF16 || d->out_mode == SD2_OUT_F32;
  {
    int kb, nbatch = 1;
    bool bmn;
    if (d->kind == SD2_GEMM_PLAIN) {
      kb = (d->K + 63) / 64;
      nbatch = d->batch > 0 ? d->batch : 1;
      bmn = d->B.mn_major != 0;
    } else if (d->kind == SD2_GEMM_CONV) {
      kb = d->conv.ntaps * ((d->conv.C + 63) / 64);
      bmn = d->B.mn_major != 0;
    } else {
      kb = (d->K + 63) / 64;
      nbatch = d->conv.ntaps;
      bmn = true;
    }
    if (kb < 1 || nbatch < 1) return fail(ctx, "sd2_gemm: empty contraction / batch");
    sd2_gemm_desc dd = *d;
    if (gn || mse) dd.max_splits = 1;  // the statistics / the loss come from the epilogue that writes the final values
    plan_gemm(ctx, &dd, p.N, kb, nbatch, bmn, direct_store, &BN, &splits);
    // measured plan (tools/autotune_gemm.py -> diffusion_b200/gemm_plans.json) overrides the cycle model
    if (d->force_bn == 256 || d->force_bn == 128 || d->force_bn == 64 || (d->force_bn == 160 && !bmn)) BN = d->force_bn;
    if (d->force_splits > 0) {
      long long smax = kb;
      if (direct_store) {
        if (nbatch != 1 || d->workspace == nullptr) smax = 1;
        else {
          const long long per_split = (long long)d->M * p.N * 4;
          if (per_split > 0 && d->workspace_bytes / per_split < smax) smax = d->workspace_bytes / per_split;
        }
      }
      if (smax < 1) smax = 1;
      splits = d->force_splits < smax ? d->force_splits : (int)smax;
    }
    if (gn || mse) splits = 1;
    if (mse) BN = 64;
  }

  // 2-CTA clusters (B tile shared through TMA multicast, or one cta_group::2 MMA per SM pair): decided before the tensor maps
  // are built, because a K-major B operand is then fetched as two half-height boxes (one per CTA of the cluster)
  int cluster;
  {
    const bool bmn = d->kind == SD2_GEMM_CONV_WGRAD ? true : d->B.mn_major != 0;
    GemmKParams q = p;
    q.mt = (int)((d->M + 127) / 128);
    q.nt = (p.N + BN - 1) / BN;
    q.splits = splits;
    q.batches = d->kind == SD2_GEMM_PLAIN ? (d->batch > 0 ? d->batch : 1) : (d->kind == SD2_GEMM_CONV_WGRAD ? d->conv.ntaps : 1);
    q.total_kb = d->kind == SD2_GEMM_CONV ? d->conv.ntaps * ((d->conv.C + 63) / 64) : (int)((d->K + 63) / 64);
    const bool amn = d->kind == SD2_GEMM_CONV_WGRAD ? true : (d->kind == SD2_GEMM_PLAIN && d->A.mn_major != 0);
    cluster = gemm_use_cluster(q, BN, amn, bmn, ctx->num_sms);
  }
  const int b_box_rows = cluster ? BN / 2 : BN;

  if (d->kind == SD2_GEMM_PLAIN) {
    a_mn = d->A.mn_major != 0;
    b_mn = d->B.mn_major != 0;
    batches = d->batch > 0 ? d->batch : 1;
    p.total_kb = (d->K + 63) / 64;
    p.a_batched = d->A.nb0 > 0;
    p.b_batched = d->B.nb0 > 0;
    p.a_nb0 = d->A.nb0 > 0 ? d->A.nb0 : 1;
    p.b_nb0 = d->B.nb0 > 0 ? d->B.nb0 : 1;
    if (!plain_tmap(&tmA, d->A, 128, &err) || !plain_tmap(&tmB, d->B, b_box_rows, &err)) return fail(ctx, "sd2_gemm plain: " + err);
  } else if (d->kind == SD2_GEMM_CONV) {
    const sd2_conv_geom& g = d->conv;
    a_mn = false;
    b_mn = d->B.mn_major != 0;
    int th, nb, wseg;
    if (!conv_box(g, 128, &th, &nb, &wseg)) return fail(ctx, "sd2_gemm conv: unsupported spatial geometry for a 128-pixel tile");
    p.cH = g.H;
    p.cth = th;
    p.cnb = nb;
    p.cws = g.W / wseg;
    p.cwseg = wseg;
    p.cblks = (g.C + 63) / 64;
    p.taps = g.ntaps;
    if (g.ntaps < 1 || g.ntaps > 9) return fail(ctx, "sd2_gemm conv: ntaps out of range");
    for (int t = 0; t < g.ntaps; ++t) {
      p.tap_dh[t] = (signed char)g.dh[t];
      p.tap_dw[t] = (signed char)g.dw[t];
      p.tap_dn[t] = g.dn[t];
      p.tap_w[t] = (signed char)g.wtap[t];
    }
    p.total_kb = g.ntaps * p.cblks;
    if (!conv_tmap(&tmA, g, th, nb, &err, wseg)) return fail(ctx, "sd2_gemm conv A: " + err);
    // weights [tap][rows][cols]: dims (cols, rows, 9, 1)
    sd2_operand w = d->B;
    w.nb0 = 9;
    w.nb1 = 1;
    if (!plain_tmap(&tmB, w, b_box_rows, &err)) return fail(ctx, "sd2_gemm conv B: " + err);
  } else if (d->kind == SD2_GEMM_CONV_WGRAD) {
    const sd2_conv_geom& g = d->conv;
    a_mn = true;
    b_mn = true;
    if (!d->A.mn_major) return fail(ctx, "sd2_gemm wgrad: A (dy) must be MN-major");
    int th, nb;
    if (!conv_box(g, 64, &th, &nb)) return fail(ctx, "sd2_gemm wgrad: unsupported spatial geometry for a 64-pixel k-block");
    p.cH = g.H;
    p.cth = th;
    p.cnb = nb;
    p.cws = 1;
    p.cwseg = g.W;
    p.taps = g.ntaps;
    if (g.ntaps < 1 || g.ntaps > 9) return fail(ctx, "sd2_gemm wgrad: ntaps out of range");
    for (int t = 0; t < g.ntaps; ++t) {
      p.tap_dh[t] = (signed char)g.dh[t];
      p.tap_dw[t] = (signed char)g.dw[t];
      p.tap_dn[t] = g.dn[t];
      p.tap_w[t] = (signed char)g.wtap[t];
    }
    batches = g.ntaps;
    p.total_kb = (d->K + 63) / 64;
    p.a_batched = 0;
    p.a_nb0 = 1;
    p.b_nb0 = 1;
    if (!plain_tmap(&tmA, d->A, 128, &err)) return fail(ctx, "sd2_gemm wgrad A: " + err);
    if (!conv_tmap(&tmB, g, th, nb, &err)) return fail(ctx, "sd2_gemm wgrad B: " + err);
  } else {
    return fail(ctx, "sd2_gemm: unknown kind");
  }

  p.mt = (d->M + 127) / 128;
  p.nt = (p.N + BN - 1) / BN;
  p.batches = batches;
  p.splits = splits;
  {
    // Item order = L2 sharing pattern of the concurrently running CTAs.  Weights always fit in the 126 MB L2, so
    // forward / dgrad walk n fastest (every activation tile crosses HBM once instead of once per n tile); the conv
    // weight gradient runs the 9 taps of one K range together (dy and x cross HBM once instead of 9 times).
    static int forced = -2;
    if (forced == -2) {
      const char* e = getenv("SD2_RASTER");
      forced = e ? atoi(e) : -1;
    }
    if (forced >= 0) p.raster = forced;
    else if (d->kind == SD2_GEMM_CONV_WGRAD) p.raster = 2;
    else if (d->kind == SD2_GEMM_CONV) p.raster = 1;
    else p.raster = (batches == 1 && !(a_mn && b_mn)) ? 1 : 0;
  }

  CUtensorMap tmO;
  cudaError_t e;
  if (splits > 1 && direct_store) {
    GemmKParams pp = p;
    pp.out_mode = OUT_F32_PARTIAL;
    pp.alpha = 1.f;
    pp.bias = nullptr;
    pp.rowbias = nullptr;
    pp.residual = nullptr;
    if (!out_tmap(&tmO, d->workspace, true, 32, p.N, d->M, p.N, splits, 1, (long long)d->M * p.N, 0, &err))
      return fail(ctx, "sd2_gemm split-K workspace map: " + err);
    e = launch_gemm_tc(tmA, tmB, tmO, pp, BN, a_mn, b_mn, cluster, ctx->num_sms, stream);
    if (e != cudaSuccess) return fail(ctx, std::string("sd2_gemm launch: ") + cudaGetErrorString(e));
    e = launch_splitk_finalize(reinterpret_cast<const float*>(d->workspace), splits, d->M, p.N, d->alpha, d->bias,
                               d->rowbias, p.rows_per_group, d->ld_rowbias, p.residual, d->ldr, d->out, d->ldo,
                               d->out_mode == SD2_OUT_F32, stream);
    if (e != cudaSuccess) return fail(ctx, std::string("sd2_gemm finalize: ") + cudaGetErrorString(e));
    ctx->launches += 2;
    return 0;
  }
  if (d->out_mode == SD2_OUT_BF16)
    p.out_mode = OUT_BF16;
  else if (d->out_mode == SD2_OUT_F32)
    p.out_mode = OUT_F32;
  else
    p.out_mode = OUT_F32_ACCUM;
  {
    const bool f32 = p.out_mode != OUT_BF16;
    if (f32 && d->ldo % 4 != 0) return fail(ctx, "sd2_gemm: fp32 ldo must be a multiple of 4");
    if (!f32 && d->ldo % 8 != 0) return fail(ctx, "sd2_gemm: bf16 ldo must be a multiple of 8");
    const long long nb0 = batches > 1 ? (p.out_nb0 < batches ? p.out_nb0 : batches) : 1;
    const long long nb1 = batches > 1 ? (batches + p.out_nb0 - 1) / p.out_nb0 : 1;
    if (!out_tmap(&tmO, d->out, f32, f32 ? 32 : gemm_out_chunk(BN), p.N, d->M, d->ldo, nb0, nb1, d->out_bs0, d->out_bs1, &err))
      return fail(ctx, "sd2_gemm output map: " + err);
  }
  e = launch_gemm_tc(tmA, tmB, tmO, p, BN, a_mn, b_mn, cluster, ctx->num_sms, stream);
  if (e != cudaSuccess) return fail(ctx, std::string("sd2_gemm launch: ") + cudaGetErrorString(e));
  ctx->launches += 1;
  return 0;
}

}  // extern "C"

// Device-side parameter block of the tcgen05 GEMM / implicit-GEMM convolution kernel (gemm_tc.cu).
#pragma once
#include "common.cuh"

namespace sd2 {

enum GemmKind {
  KIND_PLAIN = 0,       // D[b] = A[b] * B[b]^T            (operands K-major or MN-major, batched through tensor-map dims 2,3)
  KIND_CONV = 1,        // A = NHWC activations gathered by shifted TMA boxes (3x3, pad 1, stride 1), B = weights [tap][Cout][Cin]
  KIND_CONV_WGRAD = 2,  // A = dy^T (MN-major, [pixels][Cout]), B = shifted NHWC activations (MN-major), batch index = tap
};
enum GemmOut {
  OUT_BF16 = 0,         // bf16(alpha*acc + bias + rowbias + residual), TMA store
  OUT_F32 = 1,          // fp32 store of alpha*acc (+bias +rowbias), TMA store
  OUT_F32_ACCUM = 2,    // fp32 out += alpha*acc, TMA reduce-add (weight gradients; also sums split-K partials)
  OUT_F32_PARTIAL = 4,  // fp32 raw accumulators into workspace[batch*splits + split][M][N] (split-K, finalized by
                        // splitk_finalize), TMA store
};

struct GemmKParams {
  int M, N;            // output extent (per batch) used for load masking (stores are clipped by the output tensor map)
  int total_kb;        // number of 64-deep K blocks
  int splits;          // split-K factor
  int mt, nt, batches; // work decomposition: items = mt * nt * splits * batches
  int raster;          // item order (what the ~148 concurrently running CTAs share through L2):
                       //   0: m fastest, then n, split, batch      (one B slice shared, A streamed nt times)
                       //   1: n fastest, then m, split, batch      (each A tile read once by its nt CTAs; B stays in L2)
                       //   2: m, n, batch fastest, split slowest   (conv wgrad: the 9 taps of a K range run together)
  int kind;
  int a_batched, b_batched;  // plain operands: does the batch index move this operand?
  int a_nb0, b_nb0;          // batch -> (batch % nb0, batch / nb0) = tensor-map coords 2,3
  // conv geometry of the shifted operand: tensor map dims (C, W, H, Nimg), box (64, W, th, nb)
  // wide images (W > 128, forward conv only): a tile is one 128-pixel segment of a row: box (64, cwseg, 1, 1), cws
  // segments per row; cws = 1 everywhere else.
  int cH, cth, cnb, cblks, taps, cws, cwseg;
  // per-tap table: spatial shift (dh, dw), image-index offset (stride-2 phase planes) and weight tap index
  signed char tap_dh[9], tap_dw[9];
  int tap_dn[9];
  signed char tap_w[9];
  // epilogue
  int out_mode;
  int out_nb0;            // output batch -> tensor-map coords (batch % out_nb0, batch / out_nb0)
  long long out_bs0, out_bs1;  // element strides of those coords (used for the residual address only)
  const bf16* residual;
  long long ldr;
  const float* bias;      // [N] or null
  const float* rowbias;   // [M / rows_per_group][ld_rowbias] or null
  int rows_per_group;
  long long ld_rowbias;
  float alpha;
  // GroupNorm statistics of the OUTPUT, taken in the epilogue (bf16 outputs only): per slab of gn_slab (16 or 32) consecutive
  // rows and per column the sum and the sum of squares of the stored (bf16-rounded) values -> gn_part[row / gn_slab][N][2].
  // The consuming GroupNorm combines them per (image, group) and is then a pure streaming apply pass.
  float* gn_part;
  int gn_slab;
  // B-resident mode (set by launch_gemm_tc): the K loop has exactly as many k-blocks as the ring has stages and every work
  // item of a CTA has the same n tile, so stage s always holds k-block s of the SAME weight tile: the producer fetches B with
  // the CTA's first tile only and streams A alone afterwards (K = 320 linears: 100 KB of weights per 128 x 160 tile no longer
  // re-read from L2 for each of the ~28 tiles a CTA computes).
  int b_resident;
  // MSE head in the epilogue of the final conv (conv_out: 4 prediction channels in an 8-wide bf16 tensor): against the target
  // noise (NCHW, mse_dtype = SD2_DT_*), the epilogue thread of a pixel adds sum((pred - noise)^2) over the 4 channels to
  // mse_acc[0] (one atomic per warp) and writes dL/dpred = 2 (pred - noise) / count as the bf16 NHWC8 row of mse_dpred8 -
  // the loss and the first gradient of the backward pass without a pass over the prediction.
  const void* mse_target;
  bf16* mse_dpred8;
  float* mse_acc;
  int mse_dtype, mse_hw;
  float mse_k;  // 2 / (pixels * 4)
};

}  // namespace sd2

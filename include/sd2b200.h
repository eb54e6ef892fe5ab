/* sd2b200 - C ABI of the B200-native (sm_100a) SD-2 UNet training hot path.
 *
 * Drop-in boundary: this is what a maintainer of fanzhongyi/diffusion binds (ctypes) underneath
 * the unchanged Python surface `StableDiffusion.forward()/loss()` (reference diffusion/models/stable_diffusion.py:154-187)
 * and `stable_diffusion_2(...)` (reference diffusion/models/models.py:28-112).  The reference itself has no native
 * code; each entry point names the third-party library call it replaces.
 *
 * Conventions:
 *   - every function returns 0 on success, non-zero on error (message via sd2_last_error); never throws/exits;
 *   - the CALLER owns every buffer; the library allocates no tensor memory and keeps no pointers past a call;
 *   - all work is enqueued on the caller-supplied stream (a cudaStream_t passed as void*), no synchronisation;
 *   - activations are NHWC bf16 ([pixels][channels]); fp32 for statistics, gradients and master parameters;
 *   - one sd2_ctx per host thread at a time.
 */
#ifndef SD2B200_H
#define SD2B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SD2_VERSION 100

typedef struct sd2_ctx sd2_ctx;
typedef void* sd2_stream; /* cudaStream_t */

enum { SD2_DT_F32 = 0, SD2_DT_BF16 = 1, SD2_DT_F16 = 2 };

/* ---- lifecycle ------------------------------------------------------------------------------------------- */
int sd2_version(void);
int sd2_ctx_create(int device, sd2_ctx** out);
int sd2_ctx_destroy(sd2_ctx* ctx);
const char* sd2_last_error(sd2_ctx* ctx);
int sd2_num_sms(sd2_ctx* ctx);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
long long sd2_launch_count(sd2_ctx* ctx);

/* ---- K1: timesteps + noise + DDPM add_noise + sinusoidal timestep embedding ------------------------------------
 * Replaces torch.randint / torch.randn_like / DDPMScheduler.add_noise / diffusers Timesteps
 * (reference stable_diffusion.py:177,179,180 and the first op of :183).  Bit-exact with torch's CUDA Philox stream:
 * `seed`/`philox_offset` are the torch CUDA generator's state; *offset_used receives how far to advance it.
 * latents: [B,4,H,W] NCHW in lat_dtype.  Outputs: timesteps int64 [B]; noise [B,4,H,W] lat_dtype;
 * noised_nchw (optional, lat_dtype); noised_nhwc8 bf16 [B,H,W,8] (channels 4..7 zero) feeding conv_in;
 * temb bf16 [B, temb_dim] = [cos | sin] embedding rounded through lat_dtype like the reference. */
int sd2_noise_sched_fwd(sd2_ctx* ctx, uint64_t seed, uint64_t philox_offset, const void* latents, int lat_dtype, int B,
                        int H, int W, const float* alphas_cumprod, int num_train_timesteps, int64_t* out_timesteps,
                        void* out_noise, void* out_noised_nchw, void* out_noised_nhwc8, void* out_temb, int temb_dim,
                        uint64_t* offset_used, sd2_stream stream);

/* Stand-alone pieces of K1 for callers that bring their own timesteps / samples (diffusers-style
 * `unet(sample, timestep, encoder_hidden_states)`): sinusoidal embedding and NCHW(4) <-> NHWC(8) bf16 conversion. */
int sd2_timestep_embedding(sd2_ctx* ctx, const int64_t* timesteps, int B, void* out_temb, int temb_dim, int round_dtype,
                           sd2_stream stream);
int sd2_nchw4_to_nhwc8(sd2_ctx* ctx, const void* src, int src_dtype, void* dst_nhwc8, int B, int H, int W,
                       sd2_stream stream);
int sd2_nhwc8_to_nchw4(sd2_ctx* ctx, const void* src_nhwc8, void* dst, int dst_dtype, int B, int H, int W,
                       sd2_stream stream);
/* x (bf16, n elements) *= *scalar (fp32 on device) */
int sd2_scale_by_scalar(sd2_ctx* ctx, void* x, long long n, const float* scalar, sd2_stream stream);
int sd2_fill_f32(sd2_ctx* ctx, float* x, long long n, float value, sd2_stream stream);

/* ---- tcgen05 GEMM / implicit-GEMM convolution -----------------------------------------------------------------
 * Replaces cuBLASLt (nn.Linear), cuDNN (nn.Conv2d) and their autograd backward (dgrad, wgrad). */
enum { SD2_GEMM_PLAIN = 0, SD2_GEMM_CONV = 1, SD2_GEMM_CONV_WGRAD = 2 };
enum { SD2_OUT_BF16 = 0, SD2_OUT_F32 = 1, SD2_OUT_F32_ACCUM = 2 };

typedef struct sd2_operand {
  const void* ptr; /* bf16 */
  int mn_major;    /* 0: contraction dim contiguous ([rows][K]); 1: M/N dim contiguous ([K][rows]) */
  int cols, rows;  /* extents: cols = contiguous dim, rows = strided dim */
  long long ld;    /* elements between consecutive rows (multiple of 8) */
  int nb0, nb1;    /* batch b -> (b % nb0, b / nb0); nb0 = 0: operand shared by all batches */
  long long bs0, bs1;
} sd2_operand;

typedef struct sd2_conv_geom {
  const void* ptr; /* bf16 NHWC activations: [n_planes][H][W] pixels of C channels, pixel stride ldc */
  int n_planes, H, W, C;
  long long ldc;
  int ntaps;       /* 9 for 3x3; subsets for stride-2 phase dgrad */
  int dh[9], dw[9], dn[9], wtap[9]; /* per tap: spatial shift, plane offset, weight tap index */
} sd2_conv_geom;

typedef struct sd2_gemm_desc {
  int kind;
  int M, N, K;     /* per-batch extents; K = contraction length (PLAIN / CONV_WGRAD: pixels) */
  int batch;       /* PLAIN: number of batches; CONV_WGRAD: taps (set by library) */
  sd2_operand A, B;   /* CONV: B = weights [tap][rows][cols] with bs0 = tap stride; A ignored (conv used) */
  sd2_conv_geom conv; /* CONV: the A operand; CONV_WGRAD: the B operand (A = dy^T, mn_major) */
  int out_mode;
  void* out;
  long long ldo;
  int out_nb0;
  long long out_bs0, out_bs1;
  const void* residual; /* bf16 [M][ldr] or null */
  long long ldr;
  const float* bias;    /* [N] or null */
  const float* rowbias; /* [M / rows_per_group][ld_rowbias] or null (per-image time-embedding bias) */
  int rows_per_group;
  long long ld_rowbias;
  float alpha;
  void* workspace;      /* split-K scratch (fp32), may be null */
  long long workspace_bytes;
  int max_splits;       /* 0 = auto */
  int force_bn;         /* 0 = planner decides; else the tile width (256/160/128/64) from a measured plan table */
  int force_splits;     /* 0 = planner decides; else the K-split count (clamped to what the workspace / K allow) */
  /* GroupNorm statistics of the output, taken in the epilogue (SD2_OUT_BF16, single batch; disables split-K): per slab of
   * gn_slab (16 or 32; must divide the pixels per image) consecutive rows and per column the sum and sum of squares of the
   * stored values -> gn_partial[M / gn_slab][round8(N)][2] fp32.  Consumed by sd2_groupnorm_fwd_fused.  Null = off. */
  float* gn_partial;
  int gn_slab;
  /* MSE noise-prediction loss + its backward in the epilogue of the final conv (SD2_GEMM_CONV, bf16 output of <= 8 columns of
   * which the first 4 are the prediction): mse_target = the noise [B][4][H][W] in mse_dtype (SD2_DT_*), mse_hw = H*W;
   * mse_acc[0] += sum((pred - noise)^2) over the 4 channels of every pixel; mse_dpred8 = bf16 [M][8] rows
   * 2 (pred - noise) / (M * 4) (columns 4..7 zero).  Replaces F.mse_loss + its backward, reference stable_diffusion.py:185-187.
   * Null = off. */
  const void* mse_target;
  void* mse_dpred8;
  float* mse_acc;
  int mse_dtype, mse_hw;
} sd2_gemm_desc;

int sd2_gemm(sd2_ctx* ctx, const sd2_gemm_desc* d, sd2_stream stream);

/* ---- normalisation (replaces ATen group_norm / layer_norm fwd+bwd under composer LPGroupNorm/LPLayerNorm,
 *      reference diffusion/train.py:91-108) ---------------------------------------------------------------------- */
/* x,y: bf16 [B][HW][C] (pixel stride ldx / ldy); stats: fp32 [B][G][2] = (mean, rstd); ws: fp32 scratch
 * of sd2_groupnorm_ws_floats(B, C) floats. silu != 0 fuses y = silu(gn(x)). */
long long sd2_groupnorm_ws_floats(int B, int C);
int sd2_groupnorm_fwd(sd2_ctx* ctx, const void* x, long long ldx, const float* gamma, const float* beta, void* y,
                      long long ldy, float* stats, float* ws, int B, int HW, int C, int G, float eps, int silu,
                      sd2_stream stream);
/* GroupNorm(+SiLU) of a tensor whose statistics were taken by the epilogue of the GEMM / conv that produced it
 * (sd2_gemm_desc.gn_partial: [B*HW / slab][C][2]; slab = any divisor of HW) or by sd2_concat_stats: combines the partials per (image, group), writes stats [B][G][2] and
 * applies y = [silu]((x - mean) * rstd * gamma + beta) in one streaming pass - x crosses HBM once, no statistics pass. */
/* Skip concatenation out[B*HW][Ca+Cb] = [a | b] (diffusers' torch.cat in the up blocks) that also takes the GroupNorm
 * partial statistics of what it writes: part[(image * P + chunk)][Ca+Cb][2] over P pixel chunks per image (P divides HW);
 * feed them to sd2_groupnorm_fwd_fused with slab = HW / P. */
int sd2_concat_stats(sd2_ctx* ctx, const void* a, long long lda, const void* b, long long ldb, void* out, float* part, int B,
                     int HW, int Ca, int Cb, int P, sd2_stream stream);
int sd2_groupnorm_fwd_fused(sd2_ctx* ctx, const void* x, const float* gn_partial, int slab, const float* gamma,
                            const float* beta, void* y, float* stats, float* ws, int B, int HW, int C, int G, float eps,
                            int silu, sd2_stream stream);
/* dx = d(loss)/dx (+ dx_add if non-null); dgamma/dbeta accumulated (+=) in fp32.
 * Optional column sums of the dx this call writes (it is then the complete output gradient of the conv / linear that
 * produced x, so these are that layer's bias gradients - no separate pass over dx): drowsum fp32 [B][C] overwritten with
 * the per-image sums (gradient of a per-image bias: the ResNet time-embedding projection), dcolsum1 / dcolsum2 fp32 [C]
 * accumulated (+=) with the sums over the whole batch.  Any of the three may be null. */
int sd2_groupnorm_bwd(sd2_ctx* ctx, const void* dy, long long lddy, const void* x, long long ldx, const float* gamma,
                      const float* beta, const float* stats, const void* dx_add, long long ldadd, void* dx,
                      long long lddx, float* dgamma, float* dbeta, float* ws, int B, int HW, int C, int G, int silu,
                      float* drowsum, float* dcolsum1, float* dcolsum2, sd2_stream stream);
/* LayerNorm over the last dim of [rows][C]; stats fp32 [rows][2] */
int sd2_layernorm_fwd(sd2_ctx* ctx, const void* x, const float* gamma, const float* beta, void* y, float* stats,
                      long long rows, int C, float eps, sd2_stream stream);
long long sd2_layernorm_ws_floats(long long rows, int C);
/* dcolsum (optional, fp32 [C]): += column sums of the dx written (dx + dx_add = the total gradient of the normalised
 * tensor, i.e. the bias gradient of the linear that produced it) */
int sd2_layernorm_bwd(sd2_ctx* ctx, const void* dy, const void* x, const float* gamma, const float* stats,
                      const void* dx_add, void* dx, float* dgamma, float* dbeta, float* dcolsum, float* ws, long long rows,
                      int C, sd2_stream stream);

/* ---- attention pieces (materialised-score path; replaces xformers / SDPA, reference models.py:109-111) ------- */
int sd2_softmax_fwd(sd2_ctx* ctx, const float* S, long long lds, void* P, long long ldp, long long rows, int cols,
                    sd2_stream stream);
/* dS = P * (dP - rowsum(dP*P)) * scale  -> bf16 */
int sd2_softmax_bwd(sd2_ctx* ctx, const void* P, long long ldp, const float* dP, long long lddp, void* dS,
                    long long ldds, long long rows, int cols, float scale, sd2_stream stream);

/* ---- fused flash-style attention (scores never leave the SM; replaces xformers / SDPA, reference models.py:109-111).
 * q/k/v/o/d_o: bf16 column slices of [tokens][channels] activations: token row b*N + i, head h at columns h*64..h*64+63
 * relative to the given pointer, row stride ld* (elements, multiples of 8).  head_dim must be 64.
 * lse: fp32 [B*heads][Nq], base-2 log-sum-exp of the scaled scores (written by fwd, read by bwd).
 * bwd: ws = scratch of sd2_attn_bwd_ws_bytes(B, heads, Nq) bytes (fp32 dQ accumulation + rowsum(dO*O)). */
int sd2_attn_fwd(sd2_ctx* ctx, const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv,
                 void* o, long long ldo, float* lse, int B, int heads, int Nq, int Nk, int head_dim, float scale,
                 sd2_stream stream);
long long sd2_attn_bwd_ws_bytes(int B, int heads, int Nq);
int sd2_attn_bwd(sd2_ctx* ctx, const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv,
                 const void* o, long long ldo, const void* d_o, long long lddo, const float* lse, void* dq, long long lddq,
                 void* dk, long long lddk, void* dv, long long lddv, void* ws, int B, int heads, int Nq, int Nk,
                 int head_dim, float scale, sd2_stream stream);

/* ---- pointwise / layout kernels -------------------------------------------------------------------------------- */
/* GEGLU: h = [a | g] ([rows][2*C]) -> y = a * gelu_erf(g) ([rows][C]) */
int sd2_geglu_fwd(sd2_ctx* ctx, const void* h, void* y, long long rows, int C, sd2_stream stream);
/* dbias (optional, fp32 [2C]): += column sums of dh = the bias gradient of the projection that produced h */
int sd2_geglu_bwd(sd2_ctx* ctx, const void* h, const void* dy, void* dh, float* dbias, long long rows, int C,
                  sd2_stream stream);
int sd2_silu_fwd(sd2_ctx* ctx, const void* x, void* y, long long n, sd2_stream stream);
int sd2_silu_bwd(sd2_ctx* ctx, const void* x, const void* dy, void* dx, long long n, sd2_stream stream);
/* out = alpha*a + beta*b (b may be null); bf16 */
int sd2_axpby(sd2_ctx* ctx, const void* a, float alpha, const void* b, float beta, void* out, long long n,
              sd2_stream stream);
/* strided 2-D copy: dst[r][0..cols) = src[r][0..cols), row strides in elements (cols % 8 == 0); optional add */
int sd2_copy2d(sd2_ctx* ctx, const void* src, long long lds, void* dst, long long ldd, long long rows, int cols,
               int accumulate, sd2_stream stream);
int sd2_upsample2x_fwd(sd2_ctx* ctx, const void* x, void* y, int B, int H, int W, int C, sd2_stream stream);
int sd2_upsample2x_bwd(sd2_ctx* ctx, const void* dy, void* dx, int B, int H, int W, int C, sd2_stream stream);
/* Upsample2D = nearest x2 + 3x3 conv (diffusers resnet.py::Upsample2D, reached from models.py:74-78) without the 4x tensor:
 * every output phase (py, px) is a 4-tap convolution of the low-resolution input with summed weights.
 * weff16: bf16 [16 = phase*4 + tap][Cout][Cin] built from the fp32 taps w9 [9][Cout][Cin]; n = Cout*Cin.
 * scatter: dw9[t] += sum of the dweff16 (fp32) entries whose group contains tap t - the exact adjoint of the build. */
int sd2_upconv_weff_build(sd2_ctx* ctx, const float* w9, void* weff16, long long n, sd2_stream stream);
int sd2_upconv_wgrad_scatter(sd2_ctx* ctx, const float* dweff16, float* dw9, long long n, sd2_stream stream);
/* stride-2 phase split: x [B][H][W][C] -> planes [4][B][H/2][W/2][C], plane = (h%2)*2 + (w%2); and inverse */
int sd2_phase_split(sd2_ctx* ctx, const void* x, void* planes, int B, int H, int W, int C, sd2_stream stream);
int sd2_phase_merge(sd2_ctx* ctx, const void* planes, void* x, int B, int H, int W, int C, sd2_stream stream);
/* out[g][n] (+)= sum over rows r in group g of x[r][n]; x bf16 [groups*rows_per_group][ldx]; out fp32 */
int sd2_colsum(sd2_ctx* ctx, const void* x, long long ldx, float* out, long long ldo, int groups,
               long long rows_per_group, int N, int accumulate, sd2_stream stream);
int sd2_cast_f32_to_bf16(sd2_ctx* ctx, const float* src, void* dst, long long n, sd2_stream stream);
/* dst[r][0..cols_dst) bf16 = src[r][0..cols_src) fp32 zero-padded; and the transposed accumulate-back for grads */
int sd2_pad_cast_rows(sd2_ctx* ctx, const float* src, int cols_src, void* dst, int cols_dst, long long rows,
                      sd2_stream stream);
int sd2_unpad_accum_rows(sd2_ctx* ctx, const float* src, int cols_src, float* dst, int cols_dst, long long rows,
                         int accumulate, sd2_stream stream);

/* ---- loss head: replaces F.mse_loss + MeanSquaredError.update + their backward (reference
 *      stable_diffusion.py:76,101,185-187,241-242) ---------------------------------------------------------------
 * pred_nhwc8: bf16 [B][H][W][8] (conv_out output, 4 valid channels); noise: [B,4,H,W] in noise_dtype.
 * Writes loss_acc[0] += sum((pred-noise)^2), loss_acc[1] += count (fp32, caller zeroes), pred_nchw in noise_dtype,
 * and dpred_nhwc8 = gscale * 2 (pred - noise) / numel (bf16, channels 4..7 zero). */
int sd2_mse_head(sd2_ctx* ctx, const void* pred_nhwc8, const void* noise, int noise_dtype, void* pred_nchw,
                 void* dpred_nhwc8, float* loss_acc, float gscale, int B, int H, int W, sd2_stream stream);

/* ---- optimizer: replaces torch.optim.AdamW.step() (reference train.py:33, yaml SD-2-base-256.yaml:55-58) over one flat
 *      fp32 range: p *= 1 - lr*wd; m,v moments; p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps).  grad is multiplied by
 *      grad_scale first; param_bf16 (optional) receives the bf16 shadow copy; zero_grad != 0 clears grad afterwards. */
int sd2_adamw_step(sd2_ctx* ctx, float* param, float* grad, float* exp_avg, float* exp_avg_sq, void* param_bf16,
                   long long n, float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                   float grad_scale, int zero_grad, sd2_stream stream);

/* ---- rows f2-f4: the callers either side of the training step --------------------------------------------------------
 * sd2_cfg_ddim_step: one sampling step of StableDiffusion.generate() (reference stable_diffusion.py:353-371): classifier-
 *   free-guidance combine of the UNet output (pred_nhwc8 bf16 [(guidance ? 2 : 1) * B][H][W][8], unconditional half first),
 *   DDIMScheduler.step with eta = 0 on the fp32 NCHW latents [B,4,H,W] (updated in place) and the bf16 NHWC8 copy of the
 *   new latents for BOTH halves of the next UNet call (next_nhwc8).  Scalars: sqrt(1-a_t), sqrt(a_t), sqrt(a_prev),
 *   sqrt(1-a_prev), computed by the caller in fp32 like the scheduler does. */
int sd2_cfg_ddim_step(sd2_ctx* ctx, const void* pred_nhwc8, float* latents, void* next_nhwc8, int B, int H, int W,
                      int guidance, float guidance_scale, float sqrt_beta_t, float sqrt_alpha_t, float sqrt_alpha_prev,
                      float dir_coef, sd2_stream stream);
/* ema[i] = ema[i] * smoothing + param[i] * one_minus_smoothing (reference diffusion/algorithms/ema.py:62-63) */
int sd2_ema_update(sd2_ctx* ctx, float* ema, const float* param, long long n, float smoothing, float one_minus_smoothing,
                   sd2_stream stream);
/* dst bf16[n] = src[n] (src_dtype SD2_DT_F32 / F16 / BF16): the fp16 wire format of precomputed latents
 * (reference diffusion/datasets/laion/laion.py:103-111) -> the engine's bf16 context buffer */
int sd2_cast_to_bf16(sd2_ctx* ctx, const void* src, int src_dtype, void* dst, long long n, sd2_stream stream);
/* host only: gather n sample buffers of bytes_each bytes into one contiguous (pinned) batch buffer; 0 = ok */
int sd2_wire_gather(const void* const* src, int n, long long bytes_each, void* dst);

/* ---- row f1: glue kernels of the in-loop VAE encoder / CLIP text encoder (reference stable_diffusion.py:160-174); the
 *      contractions of both networks go through sd2_gemm, the norms through sd2_groupnorm_fwd / sd2_layernorm_fwd ------
 * images [B,C,H,W] (C <= 8, src_dtype) -> bf16 [B*H*W][8] rows, zero padded; and back with an affine + clamp
 * (dst = clamp(src * scale + shift, lo, hi): the `(image / 2 + 0.5).clamp(0, 1)` of generate(), :380). */
int sd2_nchw_to_nhwc8(sd2_ctx* ctx, const void* src, int src_dtype, void* dst_nhwc8, int B, int C, int H, int W,
                      sd2_stream stream);
int sd2_nhwc8_to_nchw(sd2_ctx* ctx, const void* src_nhwc8, void* dst, int dst_dtype, int B, int C, int H, int W, float scale,
                      float shift, float lo, float hi, sd2_stream stream);
/* quant_conv (1x1, quant_w fp32 [8][8], quant_b fp32 [8]) on the encoder moments (bf16 [B*H*W][8]) +
 * DiagonalGaussianDistribution.sample() with the caller's noise [B,4,H,W] + `latents *= scale`; latents / mean_out
 * (optional) are [B,4,H,W] in `dtype`, every tensor op of the reference rounded in that dtype. */
int sd2_vae_sample(sd2_ctx* ctx, const void* moments_nhwc8, const float* quant_w, const float* quant_b, const void* noise,
                   void* latents, void* mean_out, int dtype, int B, int H, int W, float scale, sd2_stream stream);
/* x bf16 [rows][D] = token_embedding[ids[r]] + position_embedding[r % L] (fp32 tables) */
int sd2_embed_tokens(sd2_ctx* ctx, const int64_t* ids, const float* token_embedding, const float* position_embedding, void* x,
                     long long rows, int L, int D, int vocab, sd2_stream stream);
/* P bf16 [rows][ldp] = softmax of S fp32 [rows][lds] over columns <= (row % period) (causal), zeros elsewhere */
int sd2_softmax_causal_fwd(sd2_ctx* ctx, const float* S, long long lds, void* P, long long ldp, long long rows, int cols,
                           int period, sd2_stream stream);
/* 1x1 convolution over <= 8 channels on NHWC8 rows: out[r][o] = b[o] + sum_k w[o][k] in[r][k] (w fp32 [n_out][n_in]);
 * the VAE decoder's post_quant_conv */
int sd2_pixel_linear8(sd2_ctx* ctx, const void* in_nhwc8, const float* w, const float* b, void* out_nhwc8, long long n, int n_in,
                      int n_out, sd2_stream stream);
/* y = gelu(x) (erf form), bf16, n % 8 == 0 */
int sd2_gelu_fwd(sd2_ctx* ctx, const void* x, void* y, long long n, sd2_stream stream);

/* ---- data parallel: gradient buckets over NCCL -------------------------------------------------------------------
 * Replaces the DistributedDataParallel bucket all-reduce Composer's Trainer wraps the model in (reference
 * diffusion/train.py:40).  The batch shards by rows, so this is the only collective of the path.  libnccl is resolved
 * at run time (the copy torch already loaded, else the system one).  Rendezvous stays with the host framework: rank 0
 * calls sd2_ddp_unique_id and ships the 128 bytes to the other ranks (torch.distributed broadcast, MPI, a file ...), every
 * rank then calls sd2_ddp_init.  sd2_ddp_allreduce_bucket reduces `count` elements in place on the caller's stream
 * (average != 0: mean over ranks, else sum) and returns without synchronising; call it as each bucket of the flat
 * gradient arena becomes final to overlap it with the rest of backward. */
int sd2_ddp_unique_id(sd2_ctx* ctx, void* out128);
int sd2_ddp_init(sd2_ctx* ctx, const void* unique_id128, int rank, int world);
int sd2_ddp_world(sd2_ctx* ctx);
int sd2_ddp_allreduce_bucket(sd2_ctx* ctx, void* ptr, long long count, int dtype, int average, sd2_stream stream);
int sd2_ddp_destroy(sd2_ctx* ctx);

/* Split-K scratch that leaves the GEMM / conv planner unconstrained for an [M, N] output cut into `splits` K ranges
 * (fp32 partial tiles).  Any smaller workspace (or none) is valid: the planner splits less. */
long long sd2_workspace_bytes(long long M, long long N, int splits);

#ifdef __cplusplus
}
#endif
#endif /* SD2B200_H */

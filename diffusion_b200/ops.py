"""Thin torch-tensor wrappers over the C ABI (include/sd2b200.h).  torch is plumbing here: device memory,
streams, pointers.  Every function enqueues on torch's current CUDA stream and returns immediately.

Layouts: activations are bf16 [pixels, channels] (NHWC flattened); 3x3 conv weights are [tap=kh*3+kw][Cout][Cin].
"""
import ctypes as C

import torch

from diffusion_b200 import _lib as L

_ctx_cache = {}


class Ctx:
    """Per-device sd2_ctx handle."""

    def __init__(self, device_index, dry=False):
        self.dry = dry
        if dry:  # CPU host-logic tests only (SD2_DRY_RUN=1): calls are type-checked and counted, nothing runs
            self.lib, self.h, self.device, self.num_sms = L.DryLib(), C.c_void_p(1), torch.device('cpu'), 148
            return
        self.lib = L.load()
        h = C.c_void_p()
        rc = self.lib.sd2_ctx_create(int(device_index), C.byref(h))
        if rc != 0:
            raise RuntimeError(f'sd2_ctx_create(device={device_index}) failed with code {rc} '
                               '(needs an sm_100 GPU; there is no fallback path)')
        self.h = h
        self.device = torch.device('cuda', device_index)
        self.num_sms = self.lib.sd2_num_sms(h)

    def check(self, rc):
        if rc != 0:
            raise RuntimeError('sd2b200: ' + self.lib.sd2_last_error(self.h).decode())

    @property
    def launches(self):
        return self.lib.sd2_launch_count(self.h)


def dry_run():
    import os
    return os.environ.get('SD2_DRY_RUN') == '1'


def get_ctx(device=None) -> Ctx:
    if dry_run():
        if 'dry' not in _ctx_cache:
            _ctx_cache['dry'] = Ctx(0, dry=True)
        return _ctx_cache['dry']
    idx = torch.cuda.current_device() if device is None else torch.device(device).index
    if idx is None:
        idx = torch.cuda.current_device()
    if idx not in _ctx_cache:
        _ctx_cache[idx] = Ctx(idx)
    return _ctx_cache[idx]


def _p(t):
    return None if t is None else t.data_ptr()


def _s():
    if dry_run():
        return None
    return torch.cuda.current_stream().cuda_stream


_DT = {torch.float32: L.DT_F32, torch.bfloat16: L.DT_BF16, torch.float16: L.DT_F16}


# ------------------------------------------------------------------------------------------------- K1
def noise_sched_fwd(ctx, latents, alphas_cumprod, seed, offset, temb_dim, want_noised_nchw=False, out_nhwc8=None,
                    out_temb=None, out_noise=None):
    """Returns (timesteps i64[B], noise like latents, noised_nhwc8 bf16 [B,H,W,8], temb bf16 [B,temb_dim],
    noised_nchw or None, philox offset consumed).  out_noise: a buffer like `latents` to receive the noise (the engine's
    static target buffer, read by the MSE head in conv_out's epilogue)."""
    B, Cc, H, W = latents.shape
    assert Cc == 4 and latents.is_contiguous()
    dev = latents.device
    ts = torch.empty(B, dtype=torch.int64, device=dev)
    noise = out_noise if out_noise is not None else torch.empty_like(latents)
    assert noise.shape == latents.shape and noise.dtype == latents.dtype and noise.is_contiguous()
    nhwc8 = out_nhwc8 if out_nhwc8 is not None else torch.empty(B, H, W, 8, dtype=torch.bfloat16, device=dev)
    temb = out_temb if out_temb is not None else torch.empty(B, temb_dim, dtype=torch.bfloat16, device=dev)
    nchw = torch.empty_like(latents) if want_noised_nchw else None
    used = C.c_uint64(0)
    ctx.check(
        ctx.lib.sd2_noise_sched_fwd(ctx.h, int(seed), int(offset), _p(latents), _DT[latents.dtype], B, H, W,
                                    _p(alphas_cumprod), alphas_cumprod.numel(), _p(ts), _p(noise), _p(nchw), _p(nhwc8),
                                    _p(temb), temb_dim, C.byref(used), _s()))
    return ts, noise, nhwc8, temb, nchw, used.value


def timestep_embedding(ctx, timesteps, out_temb, round_dtype):
    ctx.check(
        ctx.lib.sd2_timestep_embedding(ctx.h, _p(timesteps), timesteps.numel(), _p(out_temb), out_temb.shape[1],
                                       _DT[round_dtype], _s()))


def nchw4_to_nhwc8(ctx, src, dst8, B, H, W):
    ctx.check(ctx.lib.sd2_nchw4_to_nhwc8(ctx.h, _p(src), _DT[src.dtype], _p(dst8), B, H, W, _s()))


def nhwc8_to_nchw4(ctx, src8, dst, B, H, W):
    ctx.check(ctx.lib.sd2_nhwc8_to_nchw4(ctx.h, _p(src8), _p(dst), _DT[dst.dtype], B, H, W, _s()))


def scale_by_scalar(ctx, x, scalar):
    ctx.check(ctx.lib.sd2_scale_by_scalar(ctx.h, _p(x), x.numel(), _p(scalar), _s()))


def fill_f32(ctx, x, value):
    ctx.check(ctx.lib.sd2_fill_f32(ctx.h, _p(x), x.numel(), float(value), _s()))


# ------------------------------------------------------------------------------------------------- GEMM
def _operand(t, mn_major, cols, rows, ld, nb0=0, nb1=0, bs0=0, bs1=0):
    o = L.Operand()
    o.ptr, o.mn_major, o.cols, o.rows, o.ld = t if isinstance(t, int) else t.data_ptr(), int(mn_major), cols, rows, ld
    o.nb0, o.nb1, o.bs0, o.bs1 = nb0, nb1, bs0, bs1
    return o


def _epilogue(d, out, ldo, out_mode, bias=None, rowbias=None, rows_per_group=1, residual=None, ldr=0, alpha=1.0,
              workspace=None):
    d.out_mode, d.out, d.ldo = out_mode, out.data_ptr(), ldo
    d.out_nb0, d.out_bs0, d.out_bs1 = 1, 0, 0
    d.bias = _p(bias)
    d.rowbias = _p(rowbias)
    d.rows_per_group = rows_per_group
    d.ld_rowbias = rowbias.stride(0) if rowbias is not None else 0
    d.residual = _p(residual)
    d.ldr = ldr
    d.alpha = alpha
    if workspace is not None:
        d.workspace, d.workspace_bytes = workspace.data_ptr(), workspace.numel() * workspace.element_size()
    d.max_splits = 0


def run_gemm(ctx, d, plan=None):
    """plan = (tile width BN, K splits) from the measured plan table, or None for the library's cycle model."""
    if plan is not None:
        d.force_bn, d.force_splits = int(plan[0]), int(plan[1])
    ctx.check(ctx.lib.sd2_gemm(ctx.h, C.byref(d), _s()))


# ---- measured GEMM plans -----------------------------------------------------------------------------------------
# tools/autotune_gemm.py times every distinct GEMM / conv of the static schedule under each (tile width, K split) the
# kernel supports, on a B200, and stores the winners in gemm_plans.json next to this file.  The engine attaches the
# plan of a recorded op (if its key is in the table) as `plan=`; unknown shapes use the library's cycle model.
GEMM_OPS = ('linear_fwd', 'linear_dgrad', 'linear_wgrad', 'conv3x3_fwd', 'conv3x3_dgrad', 'conv3x3_wgrad')
_PLANS = None


def gemm_key(name, a, k):
    """Shape signature of a recorded GEMM op: a = positional arguments after ctx, k = keyword arguments."""
    r = int(k.get('residual') is not None)
    if name == 'linear_fwd':
        return f'linear_fwd|{a[0].shape[0]}|{a[1].shape[0]}|{a[0].shape[1]}|r{r}|f{int(bool(k.get("out_f32")))}'
    if name == 'linear_dgrad':
        return f'linear_dgrad|{a[0].shape[0]}|{a[1].shape[1]}|{a[0].shape[1]}|r{r}'
    if name == 'linear_wgrad':
        return f'linear_wgrad|{a[0].shape[1]}|{a[1].shape[1]}|{a[0].shape[0]}'
    nt = len(k.get('taps') or range(9))
    if name in ('conv3x3_fwd', 'conv3x3_dgrad'):  # (x, B, H, W, w9, out)
        return f'{name}|{a[1]}|{a[2]}|{a[3]}|{a[4].shape[1]}|{a[4].shape[2]}|{a[0].shape[1]}|t{nt}|r{r}'
    if name == 'conv3x3_wgrad':  # (dy, x, B, H, W, dw9)
        return f'conv3x3_wgrad|{a[2]}|{a[3]}|{a[4]}|{a[0].shape[1]}|{a[1].shape[1]}|t{nt}'
    return None


def gemm_plans():
    global _PLANS
    if _PLANS is None:
        import json
        import os
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'gemm_plans.json')
        _PLANS = {}
        if os.path.exists(path) and os.environ.get('SD2_NO_PLANS') != '1':
            with open(path) as f:
                _PLANS = {key: (v[0], v[1]) for key, v in json.load(f).items()}
    return _PLANS


def _gn_stats(d, gn_partial, gn_slab):
    if gn_partial is not None:
        d.gn_partial, d.gn_slab = gn_partial.data_ptr(), int(gn_slab)


def linear_fwd(ctx, x, w, out, bias=None, residual=None, rowbias=None, rows_per_group=1, alpha=1.0, out_f32=False,
               workspace=None, plan=None, gn_partial=None, gn_slab=32):
    """out[M,N] = alpha * x[M,K] @ w[N,K]^T (+bias +rowbias +residual). x, w bf16 with unit inner stride.
    gn_partial: fp32 [M / gn_slab, N, 2] receiving the GroupNorm partial statistics of `out` from the epilogue."""
    M, K = x.shape
    N = w.shape[0]
    d = L.GemmDesc()
    d.kind, d.M, d.N, d.K, d.batch = L.GEMM_PLAIN, M, N, K, 1
    d.A = _operand(x, 0, K, M, x.stride(0))
    d.B = _operand(w, 0, K, N, w.stride(0))
    _epilogue(d, out, out.stride(0), L.OUT_F32 if out_f32 else L.OUT_BF16, bias, rowbias, rows_per_group, residual,
              residual.stride(0) if residual is not None else 0, alpha, workspace)
    _gn_stats(d, gn_partial, gn_slab)
    if gn_partial is not None and plan is not None:
        plan = (plan[0], 1)
    run_gemm(ctx, d, plan)


def linear_dgrad(ctx, dy, w, dx, residual=None, workspace=None, plan=None):
    """dx[M,K] = dy[M,N] @ w[N,K]  (w read MN-major: no transposed weight copy) (+residual)."""
    M, N = dy.shape
    K = w.shape[1]
    d = L.GemmDesc()
    d.kind, d.M, d.N, d.K, d.batch = L.GEMM_PLAIN, M, K, N, 1
    d.A = _operand(dy, 0, N, M, dy.stride(0))
    d.B = _operand(w, 1, K, N, w.stride(0))
    _epilogue(d, dx, dx.stride(0), L.OUT_BF16, residual=residual, ldr=residual.stride(0) if residual is not None else 0,
              workspace=workspace)
    run_gemm(ctx, d, plan)


def linear_wgrad(ctx, dy, x, dw, plan=None):
    """dw[N,K] (fp32) += dy[M,N]^T @ x[M,K]; both operands read MN-major (contraction over rows)."""
    M, N = dy.shape
    K = x.shape[1]
    d = L.GemmDesc()
    d.kind, d.M, d.N, d.K, d.batch = L.GEMM_PLAIN, N, K, M, 1
    d.A = _operand(dy, 1, N, M, dy.stride(0))
    d.B = _operand(x, 1, K, M, x.stride(0))
    _epilogue(d, dw, dw.stride(0), L.OUT_F32_ACCUM)
    run_gemm(ctx, d, plan)


def _conv_geom(x_ptr, n_planes, H, W, Cc, ldc, taps):
    g = L.ConvGeom()
    g.ptr, g.n_planes, g.H, g.W, g.C, g.ldc = x_ptr, n_planes, H, W, Cc, ldc
    g.ntaps = len(taps)
    for i, (dh, dw, dn, wt) in enumerate(taps):
        g.dh[i], g.dw[i], g.dn[i], g.wtap[i] = dh, dw, dn, wt
    return g


TAPS_FWD = [(kh - 1, kw - 1, 0, kh * 3 + kw) for kh in range(3) for kw in range(3)]
# dgrad of a stride-1 3x3: dx[p] = sum_t dy[p - d_t] W_t^T  ->  shift by +d_t', weight tap 8 - t'
TAPS_DGRAD = [(kh - 1, kw - 1, 0, 8 - (kh * 3 + kw)) for kh in range(3) for kw in range(3)]


def taps_stride2(B):
    """Forward taps of a stride-2 pad-1 3x3 conv over the 4 phase planes [(h%2)*2+(w%2)][B][H/2][W/2]:
    input row 2*ho + kh - 1 = 2*(ho + dh) + ph with (kh=0: ph=1, dh=-1), (kh=1: ph=0, dh=0), (kh=2: ph=1, dh=0)."""
    m = {0: (1, -1), 1: (0, 0), 2: (1, 0)}
    return [(m[kh][1], m[kw][1], (m[kh][0] * 2 + m[kw][0]) * B, kh * 3 + kw) for kh in range(3) for kw in range(3)]


def taps_stride2_dgrad():
    """dgrad of the stride-2 conv, one tap subset per input phase plane: plane -> [(dh, dw, 0, weight tap)].
    dx_plane(p,q)[h2,w2] = sum_{taps t of that phase} dy[h2 - dh_t, w2 - dw_t] W_t^T."""
    m = {0: (1, -1), 1: (0, 0), 2: (1, 0)}
    out = {0: [], 1: [], 2: [], 3: []}
    for kh in range(3):
        for kw in range(3):
            out[m[kh][0] * 2 + m[kw][0]].append((-m[kh][1], -m[kw][1], 0, kh * 3 + kw))
    return out


def conv3x3_fwd(ctx, x, B, H, W, w9, out, bias=None, rowbias=None, residual=None, taps=None, n_planes=None,
                workspace=None, plan=None, gn_partial=None, gn_slab=32, mse_target=None, mse_dpred8=None, mse_acc=None):
    """x: bf16 [n_planes*H*W, Cin] NHWC; w9: bf16 [9, Cout, Cin]; out: bf16 [B*H*W, Cout].
    gn_partial: fp32 [B*H*W / gn_slab, Cout, 2] receiving the GroupNorm partial statistics of `out` from the epilogue.
    mse_target (noise [B,4,H,W]) + mse_dpred8 (bf16 [B*H*W, 8]) + mse_acc (fp32 [>=1]): the MSE head in the epilogue of
    conv_out - loss sum added to mse_acc[0], dL/dpred written to mse_dpred8 (see sd2b200.h)."""
    Cin, Cout = x.shape[1], w9.shape[1]
    d = L.GemmDesc()
    d.kind, d.M, d.N, d.K, d.batch = L.GEMM_CONV, B * H * W, Cout, 9 * Cin, 1
    d.conv = _conv_geom(x.data_ptr(), n_planes or B, H, W, Cin, x.stride(0), taps or TAPS_FWD)
    d.B = _operand(w9, 0, Cin, Cout, w9.stride(1), bs0=w9.stride(0))
    _epilogue(d, out, out.stride(0), L.OUT_BF16, bias, rowbias, H * W, residual,
              residual.stride(0) if residual is not None else 0, 1.0, workspace)
    _gn_stats(d, gn_partial, gn_slab)
    if gn_partial is not None and plan is not None:
        plan = (plan[0], 1)
    if mse_target is not None:
        assert mse_target.is_contiguous() and tuple(mse_target.shape) == (B, 4, H, W)
        d.mse_target, d.mse_dpred8, d.mse_acc = mse_target.data_ptr(), mse_dpred8.data_ptr(), mse_acc.data_ptr()
        d.mse_dtype, d.mse_hw = _DT[mse_target.dtype], H * W
        plan = None
    run_gemm(ctx, d, plan)


def conv3x3_dgrad(ctx, dy, B, H, W, w9, dx, residual=None, taps=None, n_planes=None, workspace=None, plan=None):
    """dx[B*H*W, Cin] = sum_taps shift(dy)[.., Cout] @ w9[tap'][Cout][Cin]  (weights read MN-major).
    dy may carry zero-padded channels beyond w9's Cout (conv_out: 4 real of 8)."""
    Cout, Cin = w9.shape[1], w9.shape[2]
    d = L.GemmDesc()
    d.kind, d.M, d.N, d.K, d.batch = L.GEMM_CONV, B * H * W, Cin, 9 * Cout, 1
    d.conv = _conv_geom(dy.data_ptr(), n_planes or B, H, W, Cout, dy.stride(0), taps or TAPS_DGRAD)
    d.B = _operand(w9, 1, Cin, Cout, w9.stride(1), bs0=w9.stride(0))
    _epilogue(d, dx, dx.stride(0), L.OUT_BF16, residual=residual, ldr=residual.stride(0) if residual is not None else 0,
              workspace=workspace)
    run_gemm(ctx, d, plan)


def conv3x3_wgrad(ctx, dy, x, B, H, W, dw9, taps=None, n_planes=None, plan=None):
    """dw9[tap][Cout][Cin] (fp32) += dy[B*H*W, Cout]^T @ shift_tap(x)[.., Cin]."""
    Cout, Cin = dy.shape[1], x.shape[1]
    M = B * H * W
    d = L.GemmDesc()
    d.kind, d.M, d.N, d.K, d.batch = L.GEMM_CONV_WGRAD, Cout, Cin, M, 9
    d.A = _operand(dy, 1, Cout, M, dy.stride(0))
    d.conv = _conv_geom(x.data_ptr(), n_planes or B, H, W, Cin, x.stride(0), taps or TAPS_FWD)
    _epilogue(d, dw9, dw9.stride(1), L.OUT_F32_ACCUM)
    d.out_nb0, d.out_bs0 = 16, dw9.stride(0)
    run_gemm(ctx, d, plan)


def bmm(ctx, A, a_mn, a_dims, B_, b_mn, b_dims, out, out_dims, M, N, K, batch, nb0, alpha=1.0, out_f32=False):
    """Batched GEMM over batch index b -> (b % nb0, b / nb0).  *_dims = (cols, rows, ld, bs0, bs1) in elements;
    out_dims = (ldo, bs0, bs1)."""
    d = L.GemmDesc()
    d.kind, d.M, d.N, d.K, d.batch = L.GEMM_PLAIN, M, N, K, batch
    nb1 = (batch + nb0 - 1) // nb0
    d.A = _operand(A, a_mn, a_dims[0], a_dims[1], a_dims[2], nb0, nb1, a_dims[3], a_dims[4])
    d.B = _operand(B_, b_mn, b_dims[0], b_dims[1], b_dims[2], nb0, nb1, b_dims[3], b_dims[4])
    _epilogue(d, out, out_dims[0], L.OUT_F32 if out_f32 else L.OUT_BF16, alpha=alpha)
    d.out_nb0, d.out_bs0, d.out_bs1 = nb0, out_dims[1], out_dims[2]
    run_gemm(ctx, d)


# ------------------------------------------------------------------------------------------------- norms
def groupnorm_ws(ctx, B, Cc, device):
    return torch.empty(ctx.lib.sd2_groupnorm_ws_floats(B, Cc), dtype=torch.float32, device=device)


def groupnorm_fwd(ctx, x, gamma, beta, y, stats, ws, B, HW, G, eps, silu):
    Cc = x.shape[1]
    ctx.check(
        ctx.lib.sd2_groupnorm_fwd(ctx.h, _p(x), x.stride(0), _p(gamma), _p(beta), _p(y), y.stride(0), _p(stats), _p(ws), B,
                                  HW, Cc, G, float(eps), int(silu), _s()))


def concat_stats(ctx, a, b, out, part, B, HW, P):
    """out = [a | b] along channels + GroupNorm partial statistics of out over P pixel chunks per image (see sd2b200.h)."""
    ctx.check(ctx.lib.sd2_concat_stats(ctx.h, _p(a), a.stride(0), _p(b), b.stride(0), _p(out), _p(part), B, HW, a.shape[1],
                                       b.shape[1], int(P), _s()))


def groupnorm_fwd_fused(ctx, x, gn_partial, gn_slab, gamma, beta, y, stats, ws, B, HW, G, eps, silu):
    """GroupNorm(+SiLU) of a tensor whose partial statistics the producing GEMM's epilogue wrote into gn_partial."""
    Cc = x.shape[1]
    assert x.stride(0) == Cc and y.stride(0) == Cc
    ctx.check(
        ctx.lib.sd2_groupnorm_fwd_fused(ctx.h, _p(x), _p(gn_partial), int(gn_slab), _p(gamma), _p(beta), _p(y), _p(stats),
                                        _p(ws), B, HW, Cc, G, float(eps), int(silu), _s()))


def groupnorm_bwd(ctx, dy, x, gamma, beta, stats, dx, dgamma, dbeta, ws, B, HW, G, silu, dx_add=None, drowsum=None,
                  dcolsum=None, dcolsum2=None):
    """drowsum [B, C] (overwritten) / dcolsum, dcolsum2 [C] (accumulated): column sums of the dx written, see sd2b200.h."""
    Cc = x.shape[1]
    ctx.check(
        ctx.lib.sd2_groupnorm_bwd(ctx.h, _p(dy), dy.stride(0), _p(x), x.stride(0), _p(gamma), _p(beta), _p(stats),
                                  _p(dx_add), dx_add.stride(0) if dx_add is not None else 0, _p(dx), dx.stride(0),
                                  _p(dgamma), _p(dbeta), _p(ws), B, HW, Cc, G, int(silu), _p(drowsum), _p(dcolsum),
                                  _p(dcolsum2), _s()))


def layernorm_fwd(ctx, x, gamma, beta, y, stats, eps=1e-5):
    rows, Cc = x.shape
    ctx.check(ctx.lib.sd2_layernorm_fwd(ctx.h, _p(x), _p(gamma), _p(beta), _p(y), _p(stats), rows, Cc, float(eps), _s()))


def layernorm_ws(ctx, rows, Cc, device):
    return torch.empty(ctx.lib.sd2_layernorm_ws_floats(rows, Cc), dtype=torch.float32, device=device)


def layernorm_bwd(ctx, dy, x, gamma, stats, dx, dgamma, dbeta, ws, dx_add=None, dcolsum=None):
    """dcolsum (fp32 [C], optional) += column sums of the dx written (= bias gradient of the linear that produced x)."""
    rows, Cc = x.shape
    ctx.check(
        ctx.lib.sd2_layernorm_bwd(ctx.h, _p(dy), _p(x), _p(gamma), _p(stats), _p(dx_add), _p(dx), _p(dgamma), _p(dbeta),
                                  _p(dcolsum), _p(ws), rows, Cc, _s()))


# ------------------------------------------------------------------------------------------------- pointwise
def softmax_fwd(ctx, S, P, rows, cols):
    ctx.check(ctx.lib.sd2_softmax_fwd(ctx.h, _p(S), S.stride(-2), _p(P), P.stride(-2), rows, cols, _s()))


def softmax_bwd(ctx, P, dP, dS, rows, cols, scale):
    ctx.check(
        ctx.lib.sd2_softmax_bwd(ctx.h, _p(P), P.stride(-2), _p(dP), dP.stride(-2), _p(dS), dS.stride(-2), rows, cols,
                                float(scale), _s()))


def geglu_fwd(ctx, h, y):
    ctx.check(ctx.lib.sd2_geglu_fwd(ctx.h, _p(h), _p(y), h.shape[0], y.shape[1], _s()))


def geglu_bwd(ctx, h, dy, dh, dbias=None):
    """dbias (fp32 [2C], optional) += column sums of dh: the bias gradient of the projection in front of the GEGLU."""
    ctx.check(ctx.lib.sd2_geglu_bwd(ctx.h, _p(h), _p(dy), _p(dh), _p(dbias), h.shape[0], dy.shape[1], _s()))


def silu_fwd(ctx, x, y):
    ctx.check(ctx.lib.sd2_silu_fwd(ctx.h, _p(x), _p(y), x.numel(), _s()))


def silu_bwd(ctx, x, dy, dx):
    ctx.check(ctx.lib.sd2_silu_bwd(ctx.h, _p(x), _p(dy), _p(dx), x.numel(), _s()))


def axpby(ctx, a, alpha, b, beta, out):
    ctx.check(ctx.lib.sd2_axpby(ctx.h, _p(a), float(alpha), _p(b), float(beta), _p(out), a.numel(), _s()))


def copy2d(ctx, src, dst, rows, cols, accumulate=False):
    ctx.check(ctx.lib.sd2_copy2d(ctx.h, _p(src), src.stride(0), _p(dst), dst.stride(0), rows, cols, int(accumulate), _s()))


def upsample2x_fwd(ctx, x, y, B, H, W):
    ctx.check(ctx.lib.sd2_upsample2x_fwd(ctx.h, _p(x), _p(y), B, H, W, x.shape[1], _s()))


def upsample2x_bwd(ctx, dy, dx, B, H, W):
    ctx.check(ctx.lib.sd2_upsample2x_bwd(ctx.h, _p(dy), _p(dx), B, H, W, dy.shape[1], _s()))


def _upconv_groups(parity):
    """Row (or column) groups of the 3 kernel rows for output parity `parity`: [(shift, [kernel indices])]."""
    return [(-1, [0]), (0, [1, 2])] if parity == 0 else [(0, [0, 1]), (1, [2])]


def taps_upconv(phase):
    """Taps of output phase (py, px) = (phase // 2, phase % 2) of conv3x3(nearest_up2(x)) over the LOW-resolution x:
    [(dh, dw, 0, tap)] with tap = a * 2 + b indexing weff[phase] (sd2_upconv_weff_build)."""
    py, px = phase >> 1, phase & 1
    return [(dh, dw, 0, a * 2 + b) for a, (dh, _) in enumerate(_upconv_groups(py)) for b, (dw, _) in enumerate(_upconv_groups(px))]


def taps_upconv_dgrad(phase):
    """dx[p] += sum_t dy_phase[p - d_t] weff[phase][t]^T."""
    return [(-dh, -dw, 0, t) for dh, dw, _, t in taps_upconv(phase)]


def upconv_weff_build(ctx, w9, weff):
    """w9: fp32 [9, Cout, Cin] (storage layout of the master weights) -> weff: bf16 [16, Cout, Cin]."""
    ctx.check(ctx.lib.sd2_upconv_weff_build(ctx.h, _p(w9), _p(weff), w9.shape[1] * w9.shape[2], _s()))


def upconv_wgrad_scatter(ctx, dweff, dw9):
    """dw9 (fp32 [9, Cout, Cin]) += adjoint of upconv_weff_build applied to dweff (fp32 [16, Cout, Cin])."""
    ctx.check(ctx.lib.sd2_upconv_wgrad_scatter(ctx.h, _p(dweff), _p(dw9), dw9.shape[1] * dw9.shape[2], _s()))


def phase_split(ctx, x, planes, B, H, W):
    ctx.check(ctx.lib.sd2_phase_split(ctx.h, _p(x), _p(planes), B, H, W, x.shape[1], _s()))


def phase_merge(ctx, planes, x, B, H, W):
    ctx.check(ctx.lib.sd2_phase_merge(ctx.h, _p(planes), _p(x), B, H, W, x.shape[1], _s()))


def colsum(ctx, x, out, groups, rows_per_group, accumulate):
    N = x.shape[1]
    ldo = out.stride(0) if out.dim() == 2 else 0
    ctx.check(ctx.lib.sd2_colsum(ctx.h, _p(x), x.stride(0), _p(out), ldo, groups, rows_per_group, N, int(accumulate), _s()))


def cast_f32_to_bf16(ctx, src, dst):
    ctx.check(ctx.lib.sd2_cast_f32_to_bf16(ctx.h, _p(src), _p(dst), src.numel(), _s()))


def pad_cast_rows(ctx, src, cols_src, dst, cols_dst, rows):
    ctx.check(ctx.lib.sd2_pad_cast_rows(ctx.h, _p(src), cols_src, _p(dst), cols_dst, rows, _s()))


def unpad_accum_rows(ctx, src, cols_src, dst, cols_dst, rows, accumulate=True):
    ctx.check(ctx.lib.sd2_unpad_accum_rows(ctx.h, _p(src), cols_src, _p(dst), cols_dst, rows, int(accumulate), _s()))


def mse_head(ctx, pred8, noise, pred_nchw, dpred8, loss_acc, gscale, B, H, W):
    ctx.check(
        ctx.lib.sd2_mse_head(ctx.h, _p(pred8), _p(noise), _DT[noise.dtype], _p(pred_nchw), _p(dpred8), _p(loss_acc),
                             float(gscale), B, H, W, _s()))


def adamw_step(ctx, p, g, m, v, p16, lr, beta1, beta2, eps, weight_decay, step, grad_scale=1.0, zero_grad=False):
    """In-place fused AdamW over flat fp32 tensors (p16: optional bf16 shadow of p)."""
    ctx.check(
        ctx.lib.sd2_adamw_step(ctx.h, _p(p), _p(g), _p(m), _p(v), _p(p16), p.numel(), float(lr), float(beta1), float(beta2),
                               float(eps), float(weight_decay), int(step), float(grad_scale), int(zero_grad), _s()))


# ------------------------------------------------------------------------------------------------- fused attention
def attn_bwd_ws(ctx, B, heads, Nq, device):
    return torch.empty(ctx.lib.sd2_attn_bwd_ws_bytes(B, heads, Nq), dtype=torch.uint8, device=device)


def attn_fwd(ctx, q, k, v, o, lse, B, heads, Nq, Nk, scale):
    """q/k/v/o: bf16 2-D column-slice views [B*N, heads*64] (any row stride); lse fp32 [B*heads, Nq]."""
    ctx.check(
        ctx.lib.sd2_attn_fwd(ctx.h, _p(q), q.stride(0), _p(k), k.stride(0), _p(v), v.stride(0), _p(o), o.stride(0), _p(lse), B,
                             heads, Nq, Nk, 64, float(scale), _s()))


def attn_bwd(ctx, q, k, v, o, do, lse, dq, dk, dv, ws, B, heads, Nq, Nk, scale):
    ctx.check(
        ctx.lib.sd2_attn_bwd(ctx.h, _p(q), q.stride(0), _p(k), k.stride(0), _p(v), v.stride(0), _p(o), o.stride(0), _p(do),
                             do.stride(0), _p(lse), _p(dq), dq.stride(0), _p(dk), dk.stride(0), _p(dv), dv.stride(0), _p(ws), B,
                             heads, Nq, Nk, 64, float(scale), _s()))


# ------------------------------------------------------------------------------------------------- rows f2-f4
def cfg_ddim_step(ctx, pred8, latents, next8, B, H, W, guidance, guidance_scale, sqrt_beta_t, sqrt_alpha_t, sqrt_alpha_prev,
                  dir_coef):
    """One sampling step: CFG combine of pred8 (bf16 [(2|1)*B*H*W, 8]) + DDIM step (eta=0) on the fp32 NCHW latents (in
    place) + the next UNet input (bf16 NHWC8, both halves) in next8."""
    assert latents.dtype == torch.float32 and latents.is_contiguous() and latents.shape == (B, 4, H, W)
    ctx.check(
        ctx.lib.sd2_cfg_ddim_step(ctx.h, _p(pred8), _p(latents), _p(next8), B, H, W, int(bool(guidance)), float(guidance_scale),
                                  float(sqrt_beta_t), float(sqrt_alpha_t), float(sqrt_alpha_prev), float(dir_coef), _s()))


def ema_update(ctx, ema, param, smoothing):
    """ema = ema * smoothing + param * (1 - smoothing) over flat fp32 tensors, in place."""
    assert ema.dtype == torch.float32 and param.dtype == torch.float32 and ema.numel() == param.numel()
    assert ema.is_contiguous() and param.is_contiguous()
    ctx.check(ctx.lib.sd2_ema_update(ctx.h, _p(ema), _p(param), ema.numel(), float(smoothing), float(1. - smoothing), _s()))


def cast_to_bf16(ctx, src, dst):
    """dst (bf16) = src (fp32 / fp16 / bf16), same number of elements, both contiguous."""
    assert dst.dtype == torch.bfloat16 and src.numel() == dst.numel() and src.is_contiguous() and dst.is_contiguous()
    ctx.check(ctx.lib.sd2_cast_to_bf16(ctx.h, _p(src), _DT[src.dtype], _p(dst), src.numel(), _s()))


# ------------------------------------------------------------------------------------------------- row f1 glue
def nchw_to_nhwc8(ctx, src, dst8, B, Cc, H, W):
    ctx.check(ctx.lib.sd2_nchw_to_nhwc8(ctx.h, _p(src), _DT[src.dtype], _p(dst8), B, Cc, H, W, _s()))


def nhwc8_to_nchw(ctx, src8, dst, B, Cc, H, W, scale=1.0, shift=0.0, lo=-3.0e38, hi=3.0e38):
    ctx.check(ctx.lib.sd2_nhwc8_to_nchw(ctx.h, _p(src8), _p(dst), _DT[dst.dtype], B, Cc, H, W, float(scale), float(shift),
                                        float(lo), float(hi), _s()))


def vae_sample(ctx, moments8, quant_w, quant_b, noise, latents, mean_out, B, H, W, scale):
    assert noise.dtype == latents.dtype and noise.is_contiguous() and latents.is_contiguous()
    ctx.check(ctx.lib.sd2_vae_sample(ctx.h, _p(moments8), _p(quant_w), _p(quant_b), _p(noise), _p(latents), _p(mean_out),
                                     _DT[latents.dtype], B, H, W, float(scale), _s()))


def embed_tokens(ctx, ids, tok, pos, x, L):
    assert ids.dtype == torch.int64 and ids.is_contiguous()
    ctx.check(ctx.lib.sd2_embed_tokens(ctx.h, _p(ids), _p(tok), _p(pos), _p(x), ids.numel(), L, tok.shape[1], tok.shape[0], _s()))


def softmax_causal_fwd(ctx, S, P, rows, cols, period):
    ctx.check(ctx.lib.sd2_softmax_causal_fwd(ctx.h, _p(S), S.stride(-2), _p(P), P.stride(-2), rows, cols, period, _s()))


def gelu_fwd(ctx, x, y):
    ctx.check(ctx.lib.sd2_gelu_fwd(ctx.h, _p(x), _p(y), x.numel(), _s()))


def pixel_linear8(ctx, in8, w, b, out8):
    ctx.check(ctx.lib.sd2_pixel_linear8(ctx.h, _p(in8), _p(w), _p(b), _p(out8), in8.shape[0], w.shape[1], w.shape[0], _s()))


def taps_stride2_vae(B):
    """Forward taps of the VAE's Downsample2D (F.pad(x, (0, 1, 0, 1)) + 3x3 stride-2 pad-0 conv) over the 4 phase planes:
    input row 2*ho + kh = 2*(ho + dh) + ph with (kh=0: ph=0, dh=0), (kh=1: ph=1, dh=0), (kh=2: ph=0, dh=+1); the bottom /
    right zero padding is the TMA out-of-bounds fill."""
    m = {0: (0, 0), 1: (1, 0), 2: (0, 1)}
    return [(m[kh][1], m[kw][1], (m[kh][0] * 2 + m[kw][0]) * B, kh * 3 + kw) for kh in range(3) for kw in range(3)]


# ---- data parallel (csrc/ddp.cu) ----------------------------------------------------------------------------------
def ddp_init(ctx, group=None):
    """Create this context's NCCL communicator over the ranks of `group` (idempotent).  torch.distributed is only the
    rendezvous: rank 0 of the group draws the unique id, a broadcast ships its 128 bytes."""
    import ctypes
    import torch.distributed as dist
    world = dist.get_world_size(group)
    if ctx.lib.sd2_ddp_world(ctx.h) == world:
        return
    rank = dist.get_rank(group)
    buf = ctypes.create_string_buffer(128)
    if rank == 0:
        ctx.check(ctx.lib.sd2_ddp_unique_id(ctx.h, ctypes.cast(buf, ctypes.c_void_p)))
    box = [bytes(buf.raw)]
    src = dist.get_global_rank(group, 0) if group is not None else 0
    dist.broadcast_object_list(box, src=src, group=group)
    idbuf = ctypes.create_string_buffer(box[0], 128)
    ctx.check(ctx.lib.sd2_ddp_init(ctx.h, ctypes.cast(idbuf, ctypes.c_void_p), rank, world))


def ddp_allreduce_bucket(ctx, flat, average, stream):
    """In-place all-reduce of a contiguous gradient range on `stream` (a torch.cuda.Stream)."""
    dt = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}[flat.dtype]
    ctx.check(ctx.lib.sd2_ddp_allreduce_bucket(ctx.h, _p(flat), flat.numel(), dt, 1 if average else 0, stream.cuda_stream))

"""Every GEMM / conv / weight-gradient shape of the measured plan table (diffusion_b200/gemm_plans.json: the shapes and
(tile width, K split) choices the bench geometries B=256@32^2, B=128@32^2 and B=64@64^2 actually run) against an fp32
torch restatement on the same bf16-rounded inputs.  Convolutions are checked through a generic tap-list reference
(sum over taps of shifted-input x weight-tap), so the stride-2 phase-plane tap subsets are covered as well."""
import json
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

_PLANS = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'diffusion_b200',
                                     'gemm_plans.json')))


@pytest.fixture(scope='module')
def ctx():
    from diffusion_b200 import ops
    return ops.get_ctx(torch.device('cuda', 0))


def bf(*shape, scale=1.0, seed=0):
    g = torch.Generator(device='cuda').manual_seed(seed)
    return (torch.randn(*shape, device='cuda', generator=g) * scale).to(torch.bfloat16)


def rel_err(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def shifted(x, B, H, W, dh, dw):
    """x: [B*H*W, C] NHWC -> y[b, h, w] = x[b, h + dh, w + dw], zero outside (fp32)."""
    C = x.shape[1]
    xn = x.float().view(B, H, W, C)
    y = torch.zeros_like(xn)
    h0, h1 = max(0, -dh), min(H, H - dh)
    w0, w1 = max(0, -dw), min(W, W - dw)
    if h1 > h0 and w1 > w0:
        y[:, h0:h1, w0:w1] = xn[:, h0 + dh:h1 + dh, w0 + dw:w1 + dw]
    return y.view(B * H * W, C)


def _taps(nt):
    from diffusion_b200 import ops
    if nt == 9:
        return None, ops.TAPS_FWD, ops.TAPS_DGRAD
    sub = [t for t in ops.taps_stride2_dgrad().values() if len(t) == nt][0]
    return sub, sub, sub


@pytest.mark.parametrize('key', sorted(_PLANS))
def test_planned_shape_matches_fp32(ctx, key):
    from diffusion_b200 import ops
    f = key.split('|')
    plan = (_PLANS[key][0], _PLANS[key][1])
    ws = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    if f[0] == 'linear_fwd':
        M, N, K, r, f32 = int(f[1]), int(f[2]), int(f[3]), f[4] == 'r1', f[5] == 'f1'
        x, w = bf(M, K, seed=1), bf(N, K, scale=K**-0.5, seed=2)
        bias = torch.randn(N, device='cuda')
        res = bf(M, N, seed=3) if r else None
        out = torch.empty(M, N, dtype=torch.float32 if f32 else torch.bfloat16, device='cuda')
        ops.linear_fwd(ctx, x, w, out, bias=bias, residual=res, out_f32=f32, workspace=ws, plan=plan)
        ref = x.float() @ w.float().t() + bias + (res.float() if r else 0)
        assert rel_err(out, ref) < (1e-4 if f32 else 4e-3), key
    elif f[0] == 'linear_dgrad':
        M, K, N, r = int(f[1]), int(f[2]), int(f[3]), f[4] == 'r1'
        dy, w = bf(M, N, seed=4), bf(N, K, scale=N**-0.5, seed=5)
        dx = bf(M, K, seed=6)
        res = dx.clone() if r else None
        ops.linear_dgrad(ctx, dy, w, dx, residual=dx if r else None, workspace=ws, plan=plan)
        ref = dy.float() @ w.float() + (res.float() if r else 0)
        assert rel_err(dx, ref) < 4e-3, key
    elif f[0] == 'linear_wgrad':
        N, K, M = int(f[1]), int(f[2]), int(f[3])
        dy, x = bf(M, N, seed=6), bf(M, K, seed=7)
        dw = torch.ones(N, K, dtype=torch.float32, device='cuda')
        ops.linear_wgrad(ctx, dy, x, dw, plan=plan)
        assert rel_err(dw, dy.double().t() @ x.double() + 1.0) < 3e-4, key  # fp32 tensor-core accumulation over M terms
    elif f[0] in ('conv3x3_fwd', 'conv3x3_dgrad'):
        B, H, W, w1, w2, xc, nt, r = int(f[1]), int(f[2]), int(f[3]), int(f[4]), int(f[5]), int(f[6]), int(f[7][1:]), f[8] == 'r1'
        sub, tf, td = _taps(nt)
        x = bf(B * H * W, xc, seed=1)
        w9 = bf(9, w1, w2, scale=(9 * w2)**-0.5, seed=2)
        if f[0] == 'conv3x3_fwd':
            assert xc == w2
            out = torch.empty(B * H * W, (w1 + 7) // 8 * 8, dtype=torch.bfloat16, device='cuda')
            res = bf(B * H * W, out.shape[1], seed=3) if r else None
            bias = torch.randn(out.shape[1], device='cuda')
            ops.conv3x3_fwd(ctx, x, B, H, W, w9, out, bias=bias, residual=res, taps=sub, workspace=ws, plan=plan)
            ref = sum(shifted(x, B, H, W, dh, dw) @ w9[wt].float().t() for dh, dw, _, wt in tf)
            ref = ref + bias[:w1] + (res[:, :w1].float() if r else 0)
            assert rel_err(out[:, :w1], ref) < 4e-3, key
        else:
            dx = bf(B * H * W, w2, seed=4)
            res = dx.clone() if r else None
            ops.conv3x3_dgrad(ctx, x, B, H, W, w9, dx, residual=dx if r else None, taps=sub, workspace=ws, plan=plan)
            ref = sum(shifted(x[:, :w1], B, H, W, dh, dw) @ w9[wt].float() for dh, dw, _, wt in td)
            ref = ref + (res.float() if r else 0)
            assert rel_err(dx, ref) < 4e-3, key
    elif f[0] == 'conv3x3_wgrad':
        B, H, W, dyc, xc, nt = int(f[1]), int(f[2]), int(f[3]), int(f[4]), int(f[5]), int(f[6][1:])
        assert nt == 9
        dy, x = bf(B * H * W, dyc, seed=5), bf(B * H * W, xc, seed=6)
        dw9 = torch.ones(9, dyc, xc, dtype=torch.float32, device='cuda')
        ops.conv3x3_wgrad(ctx, dy, x, B, H, W, dw9, plan=plan)
        ref = torch.stack([dy.double().t() @ shifted(x, B, H, W, dh, dw).double() for dh, dw, _, wt in ops.TAPS_FWD]) + 1.0
        assert rel_err(dw9, ref) < 3e-4, key  # fp32 tensor-core accumulation over B*H*W terms
    else:
        raise AssertionError(f'unknown plan key {key}')

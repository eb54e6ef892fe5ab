"""Dependent-launch latency floor: chains of tiny kernels inside one CUDA graph (each launch depends on the previous)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diffusion_b200 import ops  # noqa: E402

ctx = ops.get_ctx(torch.device('cuda', 0))
BF = torch.bfloat16


def chain(name, fn, n=100):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    print(f'| {name} | {e0.elapsed_time(e1) * 1e3 / n:.2f} us per launch |', flush=True)


x = torch.randn(128, 64, device='cuda').to(BF)
w = torch.randn(64, 64, device='cuda').to(BF)
y = torch.empty(128, 64, dtype=BF, device='cuda')
chain('gemm 128x64x64 (1 CTA, 1 k-block)', lambda: ops.linear_fwd(ctx, x, w, y))
x2 = torch.randn(128, 1280, device='cuda').to(BF)
w2 = torch.randn(64, 1280, device='cuda').to(BF)
chain('gemm 128x64x1280 (1 CTA, 20 k-blocks)', lambda: ops.linear_fwd(ctx, x2, w2, y))
x3 = torch.randn(16384, 320, device='cuda').to(BF)
w3 = torch.randn(320, 320, device='cuda').to(BF)
y3 = torch.empty(16384, 320, dtype=BF, device='cuda')
chain('gemm 16384x320x320 (256 tiles)', lambda: ops.linear_fwd(ctx, x3, w3, y3))
x4 = torch.randn(1024, 1280, device='cuda').to(BF)
w4 = torch.randn(1280, 1280, device='cuda').to(BF)
y4 = torch.empty(1024, 1280, dtype=BF, device='cuda')
chain('gemm 1024x1280x1280', lambda: ops.linear_fwd(ctx, x4, w4, y4))
a = torch.randn(64, 64, device='cuda').to(BF)
b = torch.empty_like(a)
chain('silu 4096 elements', lambda: ops.silu_fwd(ctx, a, b))
st = torch.empty(64, 2, device='cuda')
gam, bet = torch.ones(64, device='cuda'), torch.zeros(64, device='cuda')
chain('layernorm fwd 64x64', lambda: ops.layernorm_fwd(ctx, a, gam, bet, b, st))
q = torch.randn(16 * 64, 1280, device='cuda').to(BF)
o = torch.empty_like(q)
lse = torch.empty(16 * 20, 64, device='cuda')
chain('attn fwd B16 h20 N64', lambda: ops.attn_fwd(ctx, q, q, q, o, lse, 16, 20, 64, 64, 0.125))

"""In-graph timing of the fused attention kernels at the UNet's shapes.  Usage: python tools/attn_bench.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diffusion_b200 import ops  # noqa: E402


def bench(ctx, B, h, Nq, Nk):
    C = h * 64
    q = (torch.randn(B * Nq, C, device='cuda') * 0.5).bfloat16()
    k = (torch.randn(B * Nk, C, device='cuda') * 0.5).bfloat16()
    v = (torch.randn(B * Nk, C, device='cuda') * 0.5).bfloat16()
    o, lse, do = torch.empty_like(q), torch.empty(B * h, Nq, device='cuda'), torch.randn_like(q)
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    ws = ops.attn_bwd_ws(ctx, B, h, Nq, q.device)
    for name, fn, fl in (('fwd', lambda: ops.attn_fwd(ctx, q, k, v, o, lse, B, h, Nq, Nk, 0.125), 4),
                         ('bwd', lambda: ops.attn_bwd(ctx, q, k, v, o, do, lse, dq, dk, dv, ws, B, h, Nq, Nk, 0.125), 10)):
        fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(5):
                fn()
        g.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 200
        print(f'| attn {name} B{B} h{h} Nq{Nq} Nk{Nk} | {us:.1f} us | {fl * B * h * Nq * Nk * 64 / us / 1e6:.0f} TFLOP/s |', flush=True)


def profile_once(ctx, B, h, Nq, Nk):
    """Eager launches (no graph) of one shape, for `ncu -k regex:attn_...`."""
    C = h * 64
    q = (torch.randn(B * Nq, C, device='cuda') * 0.5).bfloat16()
    k = (torch.randn(B * Nk, C, device='cuda') * 0.5).bfloat16()
    v = (torch.randn(B * Nk, C, device='cuda') * 0.5).bfloat16()
    o, lse, do = torch.empty_like(q), torch.empty(B * h, Nq, device='cuda'), torch.randn_like(q)
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    ws = ops.attn_bwd_ws(ctx, B, h, Nq, q.device)
    for _ in range(3):
        ops.attn_fwd(ctx, q, k, v, o, lse, B, h, Nq, Nk, 0.125)
        ops.attn_bwd(ctx, q, k, v, o, do, lse, dq, dk, dv, ws, B, h, Nq, Nk, 0.125)
    torch.cuda.synchronize()


if __name__ == '__main__':
    c = ops.get_ctx(torch.device('cuda', 0))
    if len(sys.argv) > 1 and sys.argv[1] == '--one':
        profile_once(c, *[int(x) for x in sys.argv[2:6]])
        sys.exit(0)
    for cfg in [(16, 5, 4096, 4096), (16, 5, 1024, 1024), (16, 10, 256, 256), (16, 20, 64, 64), (16, 5, 4096, 77), (16, 5, 1024, 77)]:
        bench(c, *cfg)

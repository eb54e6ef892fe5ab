"""In-graph timing of the HBM-bound kernels (GroupNorm / LayerNorm fwd+bwd, GEGLU, colsum) at the UNet's shapes, with
achieved GB/s against the algorithmic bytes.  Usage: python tools/norm_bench.py [filter] [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diffusion_b200 import ops  # noqa: E402

BF = torch.bfloat16


def bf(*shape):
    return (torch.randn(*shape, device='cuda') * 0.5).to(BF)


def cases(ctx):
    out = []
    big = os.environ.get('NORM_BENCH_BIG') == '1'  # the default bench microbatch (B=128): tensors larger than L2
    gn_shapes = [(128, 1024, 320), (128, 1024, 640), (128, 256, 640), (128, 256, 1280), (128, 64, 1280), (128, 16, 1280)] if big else \
        [(16, 1024, 320), (64, 1024, 320), (16, 4096, 320), (64, 1024, 640), (64, 256, 640), (64, 64, 1280), (64, 16, 1280),
         (64, 256, 1920), (16, 16, 2560)]
    for B, HW, C in gn_shapes:
        x, y, dy, dx = bf(B * HW, C), bf(B * HW, C), bf(B * HW, C), bf(B * HW, C)
        gamma, beta = torch.ones(C, device='cuda'), torch.zeros(C, device='cuda')
        dg, db = torch.zeros(C, device='cuda'), torch.zeros(C, device='cuda')
        stats = torch.empty(B, 32, 2, device='cuda')
        ws = ops.groupnorm_ws(ctx, B, C, x.device)
        n = B * HW * C * 2
        out.append((f'gn_fwd  B{B} HW{HW} C{C}', 3 * n,
                    lambda x=x, y=y, st=stats, ws=ws, g=gamma, b=beta, B=B, HW=HW: ops.groupnorm_fwd(ctx, x, g, b, y, st, ws, B, HW, 32, 1e-5, 1)))
        out.append((f'gn_bwd  B{B} HW{HW} C{C}', 6 * n,
                    lambda x=x, dy=dy, dx=dx, st=stats, ws=ws, g=gamma, b=beta, dg=dg, db=db, B=B, HW=HW:
                    ops.groupnorm_bwd(ctx, dy, x, g, b, st, dx, dg, db, ws, B, HW, 32, 1, dx_add=dx)))
    for rows, C in ([(131072, 320), (32768, 640), (8192, 1280)] if big else
                    [(16384, 320), (65536, 320), (4096, 640), (16384, 640), (1024, 1280), (4096, 1280)]):
        x, y, dy, dx = bf(rows, C), bf(rows, C), bf(rows, C), bf(rows, C)
        gamma, beta = torch.ones(C, device='cuda'), torch.zeros(C, device='cuda')
        dg, db = torch.zeros(C, device='cuda'), torch.zeros(C, device='cuda')
        stats = torch.empty(rows, 2, device='cuda')
        ws = ops.layernorm_ws(ctx, rows, C, x.device)
        n = rows * C * 2
        out.append((f'ln_fwd  rows{rows} C{C}', 2 * n, lambda x=x, y=y, st=stats, g=gamma, b=beta: ops.layernorm_fwd(ctx, x, g, b, y, st)))
        out.append((f'ln_bwd  rows{rows} C{C}', 4 * n,
                    lambda x=x, dy=dy, dx=dx, st=stats, g=gamma, dg=dg, db=db, ws=ws: ops.layernorm_bwd(ctx, dy, x, g, st, dx, dg, db, ws, dx_add=dx)))
    for rows, C in ([(131072, 1280), (32768, 2560)] if big else [(16384, 1280), (65536, 1280)]):
        h, y, dy, dh = bf(rows, 2 * C), bf(rows, C), bf(rows, C), bf(rows, 2 * C)
        n = rows * C * 2
        out.append((f'geglu_fwd rows{rows} C{C}', 3 * n, lambda h=h, y=y: ops.geglu_fwd(ctx, h, y)))
        out.append((f'geglu_bwd rows{rows} C{C}', 5 * n, lambda h=h, dy=dy, dh=dh: ops.geglu_bwd(ctx, h, dy, dh)))
    for rows, C in ([(131072, 320), (131072, 2560), (32768, 640), (8192, 1280)] if big else [(16384, 320), (65536, 320), (16384, 2560)]):
        x = bf(rows, C)
        o = torch.zeros(C, device='cuda')
        out.append((f'colsum rows{rows} C{C}', rows * C * 2, lambda x=x, o=o, rows=rows: ops.colsum(ctx, x, o, 1, rows, True)))
    return out


def main(filt='', reps=5):
    ctx = ops.get_ctx(torch.device('cuda', 0))
    for name, nbytes, fn in cases(ctx):
        if filt and filt not in name:
            continue
        fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(reps):
                fn()
        g.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / reps
        print(f'| {name} | {us:.1f} us | {nbytes / us / 1e3:.0f} GB/s | {100 * nbytes / us / 1e3 / 6549:.0f}% of measured HBM |', flush=True)


if __name__ == '__main__':
    a = sys.argv[1:]
    main(a[0] if a else '', int(a[1]) if len(a) > 1 else 5)

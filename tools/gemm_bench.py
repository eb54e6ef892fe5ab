"""In-graph timing of representative GEMM / conv shapes of the SD-2 UNet step (B=16, 256^2) through the C ABI.
Each op is captured `reps` times into one CUDA graph (no host launch cost in the timed region) on rotating buffers
larger than L2 where the shape allows.  Usage: python tools/gemm_bench.py [filter-substring] [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diffusion_b200 import ops  # noqa: E402

BF = torch.bfloat16


def bf(*shape):
    return (torch.randn(*shape, device='cuda') * 0.5).to(BF)


def cases():
    out = []
    ws = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')

    def lin(M, N, K):
        x, w, y = bf(M, K), bf(N, K), torch.empty(M, N, dtype=BF, device='cuda')
        bias = torch.randn(N, device='cuda')
        out.append((f'linear_fwd   M{M} N{N} K{K}', 2 * M * N * K, lambda c: ops.linear_fwd(c, x, w, y, bias=bias, workspace=ws)))
        dx = torch.empty(M, K, dtype=BF, device='cuda')
        out.append((f'linear_dgrad M{M} N{N} K{K}', 2 * M * N * K, lambda c: ops.linear_dgrad(c, y, w, dx, workspace=ws)))
        dw = torch.zeros(N, K, device='cuda')
        out.append((f'linear_wgrad M{M} N{N} K{K}', 2 * M * N * K, lambda c: ops.linear_wgrad(c, y, x, dw)))

    def conv(B, H, Cin, Cout):
        M = B * H * H
        x, w9, y = bf(M, Cin), bf(9, Cout, Cin), bf(M, Cout)
        bias = torch.randn(Cout, device='cuda')
        fl = 2 * M * Cin * Cout * 9
        out.append((f'conv_fwd   B{B} {H}x{H} {Cin}->{Cout}', fl, lambda c: ops.conv3x3_fwd(c, x, B, H, H, w9, y, bias=bias, workspace=ws)))
        dx = torch.empty(M, Cin, dtype=BF, device='cuda')
        out.append((f'conv_dgrad B{B} {H}x{H} {Cin}->{Cout}', fl, lambda c: ops.conv3x3_dgrad(c, y, B, H, H, w9, dx, workspace=ws)))
        dw = torch.zeros(9, Cout, Cin, device='cuda')
        out.append((f'conv_wgrad B{B} {H}x{H} {Cin}->{Cout}', fl, lambda c: ops.conv3x3_wgrad(c, y, x, B, H, H, dw)))

    for M, N, K in [(16384, 320, 320), (16384, 2560, 320), (16384, 320, 1280), (4096, 640, 640), (4096, 5120, 640),
                    (4096, 640, 2560), (1024, 1280, 1280), (1024, 10240, 1280), (1024, 1280, 5120), (1232, 2560, 1024),
                    (16, 1280, 1280), (8192, 8192, 8192)]:
        lin(M, N, K)
    for B, H, Cin, Cout in [(16, 32, 320, 320), (16, 32, 640, 640), (16, 32, 640, 320), (16, 16, 640, 640),
                            (16, 16, 1280, 1280), (16, 8, 1280, 1280), (16, 8, 2560, 1280), (16, 4, 1280, 1280)]:
        conv(B, H, Cin, Cout)
    return out


def main(filt='', reps=10):
    ctx = ops.get_ctx(torch.device('cuda', 0))
    peak = 1631.2
    rows = []
    for name, flops, fn in cases():
        if filt and filt not in name:
            continue
        fn(ctx)
        torch.cuda.synchronize()
        s = torch.cuda.Stream()
        g = torch.cuda.CUDAGraph()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            with torch.cuda.graph(g, stream=s):
                for _ in range(reps):
                    fn(ctx)
        torch.cuda.synchronize()
        g.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / reps
        tf = flops / us / 1e6
        rows.append(f'| {name} | {us:.1f} | {tf:.0f} | {100 * tf / peak:.0f}% |')
        print(rows[-1], flush=True)
    return rows


if __name__ == '__main__':
    a = sys.argv[1:]
    main(a[0] if a else '', int(a[1]) if len(a) > 1 else 10)

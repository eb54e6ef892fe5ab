"""How does HBM bandwidth depend on the contiguous segment length of a strided access?  copy2d of the first `cols` columns
of a [rows, 320] bf16 matrix (640-byte row pitch): 64 / 128 / 256 / 640-byte segments per row.  The K = 320 GEMMs read their
A operand in 128-byte segments per k-block and store 64- or 128-byte row segments."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diffusion_b200 import ops  # noqa: E402

ctx = ops.get_ctx(torch.device('cuda', 0))
rows = 262144 * 4
src = torch.randn(rows, 320, device='cuda').bfloat16()
for cols in (32, 64, 128, 320):
    for mode in ('strided read', 'strided write'):
        if mode == 'strided read':
            a, b = src[:, :cols], torch.empty(rows, cols, dtype=torch.bfloat16, device='cuda')
        else:
            a, b = torch.randn(rows, cols, device='cuda').bfloat16(), src[:, :cols]
        for _ in range(3):
            ops.copy2d(ctx, a, b, rows, cols)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ops.copy2d(ctx, a, b, rows, cols)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 100
        print(f'| {cols * 2:4d}-byte segments, {mode:13s} | {us:8.1f} us | {2 * rows * cols * 2 / us / 1e3:7.0f} GB/s useful (read + write) |', flush=True)

"""Two GEMMs for an `ncu --set full --profile-from-start off` capture of the cta_group::2 pair mode against the 1-CTA kernel:
a K-major linear forward (16384 x 10240 x 1280, the GEGLU up-projection of the 1280 level) and an MN-major weight gradient
(1280 x 1280 over 16384 rows).  Run once with SD2_GEMM_CLUSTER=1 SD2_GEMM_PAIR=0 and once with SD2_GEMM_CLUSTER=3."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diffusion_b200 import ops  # noqa: E402

BF = torch.bfloat16
ctx = ops.get_ctx(torch.device('cuda', 0))
g = torch.Generator(device='cuda').manual_seed(0)
bf = lambda *s: (torch.randn(*s, device='cuda', generator=g) * 0.5).to(BF)
ws = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
M, N, K = 16384, 10240, 1280
x, w, y = bf(M, K), bf(N, K), torch.empty(M, N, dtype=BF, device='cuda')
bias = torch.randn(N, device='cuda')
dy, x2 = bf(M, 1280), bf(M, 1280)
dw = torch.zeros(1280, 1280, device='cuda')
fns = [lambda: ops.linear_fwd(ctx, x, w, y, bias=bias, workspace=ws, plan=(256, 1)),
       lambda: ops.linear_wgrad(ctx, dy, x2, dw, plan=(256, 3))]
for f in fns:
    f()
torch.cuda.synchronize()
torch.cuda.profiler.start()
for f in fns:
    f()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print('ok')

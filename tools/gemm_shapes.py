"""Time every distinct GEMM / implicit-conv shape of the static schedule with the plan the engine uses (no search):
`python tools/gemm_shapes.py B latent out.json`.  Used for A/B comparisons of kernel variants (e.g. SD2_GEMM_CLUSTER=0/1):
`python tools/gemm_shapes.py --diff a.json b.json` prints the per-shape and total difference."""
import json
import os
import sys
from functools import partial

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def diff(a, b):
    A, B = json.load(open(a)), json.load(open(b))
    rows = []
    for k in A:
        if k in B:
            rows.append(((B[k][0] - A[k][0]) * A[k][1] / 1e3, k, A[k][1], A[k][0], B[k][0]))
    rows.sort()
    ta = sum(v[0] * v[1] for v in A.values()) / 1e3
    tb = sum(v[0] * v[1] for v in B.values()) / 1e3
    print(f'total {ta:.2f} ms -> {tb:.2f} ms')
    for d, k, c, x, y in rows[:25] + rows[-25:]:
        print(f'{d:+7.3f} ms  {k:56s} x{c:3d} {x:8.1f} -> {y:8.1f} us')


def main():
    if sys.argv[1] == '--diff':
        import signal
        signal.signal(signal.SIGPIPE, signal.SIG_DFL)  # `... --diff a b | head` ends quietly
        return diff(sys.argv[2], sys.argv[3])
    from diffusion_b200 import ops
    from diffusion_b200.model import stable_diffusion_2
    from tools.autotune_gemm import time_op
    B, R, out = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
    dev = torch.device('cuda', 0)
    torch.manual_seed(17)
    model = stable_diffusion_2(pretrained=False, precomputed_latents=True, build_encoders=False, fsdp=False)
    batch = {'image_latents': torch.randn(B, 4, R, R, device=dev).to(torch.bfloat16),
             'caption_latents': torch.randn(B, 77, 1024, device=dev).to(torch.bfloat16)}
    o = model(batch)
    model.loss(o, batch).backward()
    torch.cuda.synchronize()
    eng = model._last_engine
    seen = {}
    for op in list(eng.fwd) + list(eng.bwd):
        name = op.func.__name__
        if name not in ops.GEMM_OPS:
            continue
        key = ops.gemm_key(name, op.args[1:], op.keywords)
        if key is not None:
            seen.setdefault(key, [op, 0])[1] += 1
    table = {}
    for key, (op, count) in seen.items():
        table[key] = [round(time_op(partial(op.func, *op.args, **op.keywords), reps=6), 2), count]
    tot = sum(v[0] * v[1] for v in table.values()) / 1e3
    print(f'B={B} latent={R}: {len(table)} shapes, GEMM family isolated {tot:.2f} ms', flush=True)
    with open(out, 'w') as f:
        json.dump(table, f, indent=0, sort_keys=True)


if __name__ == '__main__':
    main()

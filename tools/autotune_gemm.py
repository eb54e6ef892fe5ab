"""Measured tile plans for the tensor-core GEMM family.

For every distinct GEMM / implicit-conv shape of the static schedule (forward + backward) at the given per-GPU batch
and latent size, time the kernel under each (tile width BN, K-split) it supports - in a CUDA graph, on the op's real
buffers - and keep the fastest.  Output: a JSON table {shape key: [BN, splits, best us, planner us]} that is merged
into diffusion_b200/gemm_plans.json (the engine reads it at build time; unknown shapes use the library's cycle model).

Usage (on a B200): SD2_NO_PLANS=1 [SD2_TUNE_ONLY=substr,substr] python tools/autotune_gemm.py B latent out.json [B latent ...]
"""
import json
import os
import sys
from functools import partial

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ['SD2_NO_PLANS'] = '1'
from diffusion_b200 import ops  # noqa: E402
from diffusion_b200.model import stable_diffusion_2  # noqa: E402

BNS = (256, 160, 128, 64)
SPLITS = (1, 2, 3, 4, 5, 6, 8, 10, 12, 16, 20, 24, 32)


def time_op(fn, reps=8):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    best = 1e30
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(2):
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3 / reps)
    return best


def tune(B, R, table):
    dev = torch.device('cuda', 0)
    torch.manual_seed(17)
    model = stable_diffusion_2(pretrained=False, precomputed_latents=True, build_encoders=False, fsdp=False)
    batch = {'image_latents': torch.randn(B, 4, R, R, device=dev).to(torch.bfloat16),
             'caption_latents': torch.randn(B, 77, 1024, device=dev).to(torch.bfloat16)}
    out = model(batch)  # one real step: the timed GEMMs then run on real activations / gradients (tensor-core power,
    model.loss(out, batch).backward()  # hence the clocks, depend on the data: zeros would flatter every plan)
    torch.cuda.synchronize()
    eng = model._last_engine
    seen = {}
    for op in list(eng.fwd) + list(eng.bwd):
        name = op.func.__name__
        if name not in ops.GEMM_OPS:
            continue
        key = ops.gemm_key(name, op.args[1:], op.keywords)
        if key is None:
            continue
        only = os.environ.get('SD2_TUNE_ONLY')  # comma-separated substrings: tune only the shape keys containing one of them
        if only and not any(x in key for x in only.split(',')):
            continue
        seen.setdefault(key, [op, 0])[1] += 1
    tot_def = tot_best = 0.0
    for key, (op, count) in seen.items():
        kw = {k: v for k, v in op.keywords.items() if k != 'plan'}
        base = time_op(partial(op.func, *op.args, **kw))
        best, best_plan = base, None
        for bn in BNS:
            prev = None
            for s in SPLITS:
                try:
                    t = time_op(partial(op.func, *op.args, **kw, plan=(bn, s)), reps=5)
                except RuntimeError:
                    break
                if prev is not None and t > 1.6 * best:
                    break  # more splits only get slower from here
                prev = t
                if t < best * 0.97:
                    best, best_plan = t, (bn, s)
        if best_plan is not None:
            best = time_op(partial(op.func, *op.args, **kw, plan=best_plan))  # confirm with the longer measurement
            if best < base * 0.97:
                table[key] = [best_plan[0], best_plan[1], round(best, 1), round(base, 1)]
        tot_def += base * count
        tot_best += min(best, base) * count
        print(f'{key:60s} x{count:3d}  planner {base:8.1f} us  best {min(best, base):8.1f} us  plan {table.get(key, ["-", "-"])[:2]}', flush=True)
    print(f'B={B} latent={R}: GEMM family (isolated, warm L2) planner {tot_def / 1e3:.2f} ms -> tuned {tot_best / 1e3:.2f} ms', flush=True)
    del eng, model
    torch.cuda.empty_cache()


def main():
    a = sys.argv[1:]
    out = [x for x in a if x.endswith('.json')][0]
    nums = [int(x) for x in a if not x.endswith('.json')]
    table = {}
    for i in range(0, len(nums), 2):
        tune(nums[i], nums[i + 1], table)
    with open(out, 'w') as f:
        json.dump(table, f, indent=0, sort_keys=True)
    print(f'{len(table)} plans -> {out}')


if __name__ == '__main__':
    main()

"""Per-op timing of the real static schedule (CUDA events, each op replayed in isolation on its own buffers).
Usage: python tools/op_profile.py [B] [latent] [out.md]    -> table grouped by (kernel wrapper, shape signature)."""
import os
import sys
from collections import defaultdict

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diffusion_b200 import ops  # noqa: E402
from diffusion_b200.model import stable_diffusion_2  # noqa: E402


def sig(p):
    fn, a = p.func, p.args[1:]
    name = fn.__name__
    def shp(t):
        return 'x'.join(map(str, t.shape)) if torch.is_tensor(t) else str(t)
    if name in ('linear_fwd', 'linear_dgrad', 'linear_wgrad'):
        s = f'{shp(a[0])} , {shp(a[1])}'
        if name == 'linear_fwd':
            fl = 2 * a[0].shape[0] * a[0].shape[1] * a[1].shape[0]
        elif name == 'linear_dgrad':
            fl = 2 * a[0].shape[0] * a[0].shape[1] * a[1].shape[1]
        else:
            fl = 2 * a[0].shape[0] * a[0].shape[1] * a[1].shape[1]
    elif name in ('conv3x3_fwd', 'conv3x3_dgrad'):
        s = f'B{a[1]} {a[2]}x{a[3]} w{shp(a[4])}'
        ntaps = len(p.keywords.get('taps') or range(9))
        fl = 2 * a[1] * a[2] * a[3] * a[4].shape[1] * a[4].shape[2] * ntaps
    elif name == 'conv3x3_wgrad':
        s = f'B{a[2]} {a[3]}x{a[4]} w{shp(a[5])}'
        fl = 2 * a[2] * a[3] * a[4] * a[5].shape[1] * a[5].shape[2] * 9
    elif name == 'bmm':
        s = f'M{a[8]} N{a[9]} K{a[10]} b{a[11]} amn{a[1]} bmn{a[4]}'
        fl = 2 * a[8] * a[9] * a[10] * a[11]
    elif name == 'attn_fwd':  # (q, k, v, o, lse, B, heads, Nq, Nk, scale)
        s = f'B{a[5]} h{a[6]} Nq{a[7]} Nk{a[8]}'
        fl = 4 * a[5] * a[6] * a[7] * a[8] * 64
    elif name == 'attn_bwd':  # (q, k, v, o, do, lse, dq, dk, dv, ws, B, heads, Nq, Nk, scale)
        s = f'B{a[10]} h{a[11]} Nq{a[12]} Nk{a[13]}'
        fl = 10 * a[10] * a[11] * a[12] * a[13] * 64
    else:
        s = ' '.join(shp(t) for t in a[:3] if torch.is_tensor(t))
        fl = 0
    return name, s, fl


def main(B=16, R=32, out=None):
    dev = torch.device('cuda', 0)
    torch.manual_seed(17)
    model = stable_diffusion_2(pretrained=False, precomputed_latents=True, build_encoders=False, fsdp=False)
    batch = {'image_latents': torch.randn(B, 4, R, R, device=dev).to(torch.bfloat16),
             'caption_latents': torch.randn(B, 77, 1024, device=dev).to(torch.bfloat16)}
    o = model(batch)
    model.loss(o, batch).backward()
    torch.cuda.synchronize()
    eng = model._last_engine
    agg = defaultdict(lambda: [0, 0.0, 0.0])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    total = 0.0
    for phase, lst in (('fwd', eng.fwd), ('bwd', eng.bwd)):
        for p in lst:
            p()
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for _ in range(reps):
                    p()
            g.replay()
            e0.record()
            g.replay()
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / reps
            name, s, fl = sig(p)
            k = (phase, name, s)
            agg[k][0] += 1
            agg[k][1] += us
            agg[k][2] += fl
            total += us
    lines = [f'B={B} latent={R}: sum of isolated op times {total / 1e3:.2f} ms (each op captured {reps}x in a CUDA graph and replayed: no host launch cost, warm L2)', '',
             '| phase | op | shape | n | total us | share | TFLOP/s |', '|---|---|---|---:|---:|---:|---:|']
    for (phase, name, s), (n, us, fl) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        tf = f'{fl / us / 1e6:.0f}' if fl else ''
        lines.append(f'| {phase} | {name} | {s} | {n} | {us:.0f} | {100 * us / total:.1f}% | {tf} |')
    byname = defaultdict(float)
    for (phase, name, s), (n, us, fl) in agg.items():
        byname[phase + ' ' + name] += us
    lines += ['', '| op | total us | share |', '|---|---:|---:|']
    for k, us in sorted(byname.items(), key=lambda kv: -kv[1]):
        lines.append(f'| {k} | {us:.0f} | {100 * us / total:.1f}% |')
    text = '\n'.join(lines)
    print(text)
    if out:
        with open(out, 'w') as f:
            f.write(text + '\n')


if __name__ == '__main__':
    a = sys.argv[1:]
    main(int(a[0]) if a else 16, int(a[1]) if len(a) > 1 else 32, a[2] if len(a) > 2 else None)

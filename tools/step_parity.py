"""End-to-end parity probe: product training step (sm_100a kernels) vs the torch oracle on the same seed/inputs, through
tests/parity.py (the same code the GPU tests gate on).  Writes gpurun_out/step_parity_<cfg>_<B>x<R>.json with the
losses, every per-parameter gradient cosine (product vs fp32 oracle, bf16-autocast oracle vs fp32 oracle), the gate
verdict, and a run-to-run determinism check of the product's gradients.

  python tools/step_parity.py tiny|sd2 B R [graphs]
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import parity  # noqa: E402
from oracle.unet import SD2_BASE_UNET_CONFIG, TINY_UNET_CONFIG  # noqa: E402


def main(cfg_name='tiny', B=2, R=32, graphs=0):
    cfg = TINY_UNET_CONFIG if cfg_name == 'tiny' else SD2_BASE_UNET_CONFIG
    res = parity.step_triplet(cfg, B, R)
    out = parity.summary(res)
    out['cfg'] = cfg_name
    model, batch, g1 = res['model'], res['batch'], res['grads_product']
    eng = model._last_engine
    out['fwd_ops'], out['bwd_ops'], out['act_GB'] = len(eng.fwd), len(eng.bwd), eng.act_bytes / 2**30
    # run-to-run determinism of the product (same inputs, same RNG state)
    _, _, g2 = parity.product_step(model, batch)
    rr = {n: parity._cos(g1[n], g2[n]) for n in g1 if g1[n].norm().item() > 0}
    out['rerun_cos_min'] = min(rr.values())
    out['rerun_bit_identical'] = sum(bool(torch.equal(g1[n], g2[n])) for n in g1)
    out['rerun_tensors'] = len(g1)
    if graphs:
        eng.capture_graphs()
        _, _, g3 = parity.product_step(model, batch)
        out['graph_vs_eager_cos_min'] = min(parity._cos(g1[n], g3[n]) for n in g1 if g1[n].norm().item() > 0)
    out['cos_product_fp32'] = res['cos_product_fp32']
    out['cos_oracle16_fp32'] = res['cos_oracle16_fp32']
    os.makedirs('gpurun_out', exist_ok=True)
    with open(f'gpurun_out/step_parity_{cfg_name}_{B}x{R}.json', 'w') as f:
        json.dump(out, f, indent=1)
    print(json.dumps({k: v for k, v in out.items() if not k.startswith('cos_product') and not k.startswith('cos_oracle16')},
                     indent=1))


if __name__ == '__main__':
    a = sys.argv[1:]
    main(a[0] if a else 'tiny', int(a[1]) if len(a) > 1 else 2, int(a[2]) if len(a) > 2 else 32, int(a[3]) if len(a) > 3 else 0)

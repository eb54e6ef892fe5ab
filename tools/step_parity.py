"""End-to-end parity probe: product training step (sm_100a kernels) vs the torch oracle on the same seed/inputs.
Writes gpurun_out/step_parity_<cfg>.json with loss and per-parameter gradient cosines."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.stable_diffusion import StableDiffusionOracle, train_step  # noqa: E402
from oracle.unet import SD2_BASE_UNET_CONFIG, TINY_UNET_CONFIG  # noqa: E402


def main(cfg_name='tiny', B=2, HW=32, graphs=0):
    from diffusion_b200.model import stable_diffusion_2
    dev = torch.device('cuda', 0)
    cfg = TINY_UNET_CONFIG if cfg_name == 'tiny' else SD2_BASE_UNET_CONFIG
    torch.manual_seed(17)
    oracle = StableDiffusionOracle(cfg).to(dev)
    model = stable_diffusion_2(pretrained=False, precomputed_latents=True, unet_config=cfg, fsdp=False)
    model.unet.load_state_dict(oracle.unet.state_dict())
    g = torch.Generator(device=dev).manual_seed(5)
    batch = {
        'image_latents': torch.randn(B, 4, HW, HW, device=dev, generator=g).to(torch.bfloat16),
        'caption_latents': torch.randn(B, 77, 1024, device=dev, generator=g).to(torch.bfloat16),
    }
    res = {'cfg': cfg_name, 'B': B, 'HW': HW}
    # ---- product
    torch.manual_seed(123)
    t0 = time.time()
    out = model(batch)
    loss = model.loss(out, batch)
    loss.backward()
    torch.cuda.synchronize()
    res['product_first_step_s'] = time.time() - t0
    res['loss_product'] = loss.item()
    eng = model._last_engine
    res['fwd_ops'], res['bwd_ops'], res['act_GB'] = len(eng.fwd), len(eng.bwd), eng.act_bytes / 2**30
    grads_p = {n: p.grad.detach().float().clone() for n, p in model.unet.named_parameters()}
    # ---- oracle, bf16 autocast (the reference's amp_bf16 path) and fp32
    for tag, dt in (('bf16', torch.bfloat16), ('fp32', None)):
        oracle.zero_grad(set_to_none=True)
        torch.manual_seed(123)
        b = batch if dt is not None else {k: v.float() for k, v in batch.items()}
        lo, oo = train_step(oracle, b, autocast_dtype=dt, timesteps=None if dt is not None else out[2],
                            noise=None if dt is not None else out[1].float())
        res[f'loss_oracle_{tag}'] = lo.item()
        if dt is not None:
            res['timesteps_equal'] = bool(torch.equal(oo[2], out[2]))
            res['noise_bit_exact'] = bool(torch.equal(oo[1].view(torch.int16), out[1].view(torch.int16)))
            res['pred_max_abs_diff_vs_bf16'] = (oo[0].float() - out[0].float()).abs().max().item()
            res['pred_max_abs'] = oo[0].float().abs().max().item()
        cos = {}
        for n, p in oracle.unet.named_parameters():
            a, b_ = grads_p[n].flatten(), p.grad.detach().float().flatten()
            cos[n] = torch.nn.functional.cosine_similarity(a, b_, dim=0).item()
        worst = sorted(cos.items(), key=lambda kv: kv[1])[:12]
        res[f'cos_min_{tag}'] = worst[0][1]
        res[f'cos_worst_{tag}'] = worst
        res[f'cos_below_0.999_{tag}'] = sum(1 for v in cos.values() if v < 0.999)
        res[f'cos_mean_{tag}'] = sum(cos.values()) / len(cos)
        if tag == 'bf16':
            grads_o16 = {n: p.grad.detach().float().clone() for n, p in oracle.unet.named_parameters()}
        else:
            c2 = {n: torch.nn.functional.cosine_similarity(grads_o16[n].flatten(), p.grad.float().flatten(), dim=0).item()
                  for n, p in oracle.unet.named_parameters()}
            res['cos_min_oracle_bf16_vs_fp32'] = min(c2.values())
            res['cos_below_0.999_oracle_bf16_vs_fp32'] = sum(1 for v in c2.values() if v < 0.999)
    # ---- second step + optional CUDA graphs
    if graphs:
        model.unet.zero_grad(set_to_none=True)
        eng.capture_graphs()
        torch.manual_seed(123)
        out2 = model(batch)
        loss2 = model.loss(out2, batch)
        loss2.backward()
        torch.cuda.synchronize()
        res['loss_product_graph'] = loss2.item()
        gmax = max((model.unet.get_parameter(n).grad.float() - grads_p[n]).abs().max().item() for n in grads_p)
        res['graph_vs_eager_grad_maxdiff'] = gmax
    os.makedirs('gpurun_out', exist_ok=True)
    with open(f'gpurun_out/step_parity_{cfg_name}_{B}x{HW}.json', 'w') as f:
        json.dump(res, f, indent=1)
    print(json.dumps({k: v for k, v in res.items() if 'worst' not in k}, indent=1))
    for tag in ('bf16', 'fp32'):
        print(tag, res[f'cos_worst_{tag}'][:6])


if __name__ == '__main__':
    a = sys.argv[1:]
    main(a[0] if a else 'tiny', int(a[1]) if len(a) > 1 else 2, int(a[2]) if len(a) > 2 else 32, int(a[3]) if len(a) > 3 else 0)

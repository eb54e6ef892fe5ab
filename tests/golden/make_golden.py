"""Generates tests/golden/tiny_step.pt: seeded inputs -> outputs of the ORACLE for the tiny UNet config (config 1 of
BASELINE.json: CPU, fp32, batch 2, 32x32x4 latents).  The reference's own tests hold no numeric vectors
(tests/test_model.py:27-28,46 assert shapes only) and its stack (diffusers/composer) is not installable here, so
these goldens pin the oracle against regressions, not against diffusers ("parity unpinned").
Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle.philox import randint_cuda  # noqa: E402
from oracle.stable_diffusion import StableDiffusionOracle, train_step  # noqa: E402
from oracle.unet import TINY_UNET_CONFIG  # noqa: E402


def compute():
    torch.manual_seed(17)
    torch.set_num_threads(1)  # deterministic reduction order
    model = StableDiffusionOracle(TINY_UNET_CONFIG)
    g = torch.Generator().manual_seed(5)
    batch = {'image_latents': torch.randn(2, 4, 32, 32, generator=g), 'caption_latents': torch.randn(2, 77, 1024, generator=g)}
    ts = torch.tensor([37, 912])
    noise = torch.randn(2, 4, 32, 32, generator=g)
    loss, out = train_step(model, batch, timesteps=ts, noise=noise)
    names = ['conv_in.weight', 'time_embedding.linear_2.bias', 'down_blocks.1.attentions.0.transformer_blocks.0.attn2.to_k.weight',
             'mid_block.resnets.1.conv2.weight', 'up_blocks.3.resnets.2.conv_shortcut.weight', 'conv_out.bias']
    grads = {n: model.unet.get_parameter(n).grad.detach().clone() for n in names}
    return {
        'loss': loss.clone(), 'pred_sum': out[0].detach().double().sum().float(), 'pred_slice': out[0].detach()[0, :, 0, :4].clone(),
        'grad_norms': {n: g_.norm().clone() for n, g_ in grads.items()}, 'grad_conv_out_bias': grads['conv_out.bias'],
        'alphas_cumprod_0_499_999': model.noise_scheduler.alphas_cumprod[[0, 499, 999]].clone(),
        'randint_seed17_off0_B16': torch.from_numpy(randint_cuda(17, 0, 16, 1000)[0]),
        'randint_seed123_off8_B300': torch.from_numpy(randint_cuda(123, 8, 300, 1000)[0]),
    }


def compute_next_rows():
    """tests/golden/next_rows.pt: DDIM scheduler steps (f4) and the tiny VAE oracle (f1) on seeded inputs."""
    from oracle.ddim import DDIMSchedulerOracle
    from oracle.vae import TINY_VAE_CONFIG, AutoencoderKLOracle
    torch.set_num_threads(1)
    g = torch.Generator().manual_seed(9)
    x, e = torch.randn(2, 4, 8, 8, generator=g), torch.randn(2, 4, 8, 8, generator=g)
    sch = DDIMSchedulerOracle()
    sch.set_timesteps(50)
    steps = {t: sch.step(e, t, x).clone() for t in (981, 501, 1)}
    sch.set_timesteps(4)
    ts4 = sch.timesteps.clone()
    torch.manual_seed(11)
    vae = AutoencoderKLOracle(**TINY_VAE_CONFIG)
    img = torch.rand(1, 3, 32, 32, generator=g) * 2 - 1
    with torch.no_grad():
        mom = vae.moments(img)
        dec = vae.decode(torch.randn(1, 4, 4, 4, generator=g))
    return {'ddim_prev_981': steps[981], 'ddim_prev_501': steps[501], 'ddim_prev_1': steps[1], 'ddim_timesteps_4': ts4,
            'vae_moments': mom.clone(), 'vae_decode_slice': dec[0, :, :2, :6].clone(), 'vae_decode_sum': dec.double().sum().float()}


if __name__ == '__main__':
    here = os.path.dirname(os.path.abspath(__file__))
    torch.save(compute(), os.path.join(here, 'tiny_step.pt'))
    torch.save(compute_next_rows(), os.path.join(here, 'next_rows.pt'))
    print('wrote', here)

"""ORACLE (test infrastructure): restatement of the reference training step
`StableDiffusion.forward` (`diffusion/models/stable_diffusion.py:154-183`, precomputed-latents branch :157-158)
and `.loss` (:185-187), plus the Composer microbatch semantics of SURVEY.md B7 that change numerics
(bf16 autocast, low-precision GroupNorm/LayerNorm surgery from `diffusion/train.py:91-108`).
PARITY UNPINNED (see oracle/unet.py header).
"""
import contextlib

import torch
import torch.nn.functional as F

from oracle.ddpm import DDPMScheduler
from oracle.unet import UNet2DConditionModel, apply_low_precision_norms


class StableDiffusionOracle(torch.nn.Module):

    def __init__(self, unet_config, low_precision_norms=True):
        super().__init__()
        self.unet = UNet2DConditionModel(**unet_config)
        if low_precision_norms:
            apply_low_precision_norms(self.unet)
        self.noise_scheduler = DDPMScheduler()
        self.image_latents_key, self.text_latents_key = 'image_latents', 'caption_latents'

    def forward(self, batch, timesteps=None, noise=None):
        latents, conditioning = batch[self.image_latents_key], batch[self.text_latents_key]
        # order of RNG consumption is part of the contract: randint first, randn_like second (:177,:179)
        if timesteps is None:
            timesteps = torch.randint(0, len(self.noise_scheduler), (latents.shape[0],), device=latents.device)
        if noise is None:
            noise = torch.randn_like(latents)
        noised_latents = self.noise_scheduler.add_noise(latents, noise, timesteps)
        return self.unet(noised_latents, timesteps, conditioning)['sample'], noise, timesteps

    def loss(self, outputs, batch):
        return F.mse_loss(outputs[0], outputs[1])


def train_step(model: StableDiffusionOracle, batch, autocast_dtype=None, timesteps=None, noise=None):
    """fwd + loss + bwd of one microbatch; returns (loss, outputs). Grads land in .grad."""
    dev = batch['image_latents'].device.type
    ctx = torch.autocast(dev, dtype=autocast_dtype) if autocast_dtype is not None else contextlib.nullcontext()
    with ctx:
        outputs = model(batch, timesteps=timesteps, noise=noise)
        loss = model.loss(outputs, batch)
    loss.backward()
    return loss.detach(), outputs

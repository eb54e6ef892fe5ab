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

    def attach_encoders(self, vae, text_encoder, encode_dtype=None):
        """In-loop encoding branch of the reference forward (:159-174): `vae` = oracle.vae.AutoencoderKLOracle,
        `text_encoder` = a module whose call returns (last_hidden_state, ...) (transformers CLIPTextModel).
        encode_dtype: dtype the images are cast to before the VAE (the reference's `inputs.half()`)."""
        self.vae, self.text_encoder, self.encode_dtype = vae, text_encoder, encode_dtype

    def forward(self, batch, timesteps=None, noise=None):
        if self.image_latents_key in batch and self.text_latents_key in batch:
            latents, conditioning = batch[self.image_latents_key], batch[self.text_latents_key]
        else:
            inputs, conditioning = batch['image'], batch['captions']
            conditioning = conditioning.view(-1, conditioning.shape[-1])
            with torch.autocast(inputs.device.type, enabled=False), torch.no_grad():
                x = inputs.to(self.encode_dtype) if self.encode_dtype is not None else inputs
                # RNG order: the VAE posterior noise is drawn first (then randint, then randn_like below)
                noise_vae = torch.randn((x.shape[0], 4, x.shape[2] // 8, x.shape[3] // 8), device=x.device, dtype=x.dtype)
                mean, logvar = torch.chunk(self.vae.moments(x.float()), 2, dim=1)
                std = torch.exp(0.5 * torch.clamp(logvar, -30.0, 20.0))
                latents = (mean + std * noise_vae.float()).to(x.dtype)
                conditioning = self.text_encoder(conditioning)[0]
            latents = latents * 0.18215
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
    dev = next(iter(batch.values())).device.type
    ctx = torch.autocast(dev, dtype=autocast_dtype) if autocast_dtype is not None else contextlib.nullcontext()
    with ctx:
        outputs = model(batch, timesteps=timesteps, noise=noise)
        loss = model.loss(outputs, batch)
    loss.backward()
    return loss.detach(), outputs

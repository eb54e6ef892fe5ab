"""ORACLE (test infrastructure): restatement of diffusers `DDIMScheduler` as the reference builds it from the SD-2-base
scheduler config (`diffusion/models/models.py:89`: scaled_linear betas 0.00085..0.012, 1000 train steps, epsilon
prediction, clip_sample False, set_alpha_to_one False, steps_offset 1, 'leading' spacing) and calls it in
`StableDiffusion.generate` (`diffusion/models/stable_diffusion.py:348-371`: set_timesteps, init_noise_sigma,
scale_model_input, step), plus that sampling loop itself up to (not including) the VAE decode.
PARITY UNPINNED (diffusers absent): anchored on the published DDIM update (Song et al. 2020, eq. 12 with sigma = 0).
"""
import numpy as np
import torch


class DDIMSchedulerOracle:
    init_noise_sigma = 1.0

    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, set_alpha_to_one=False, steps_offset=1):
        self.num_train_timesteps, self.steps_offset = num_train_timesteps, steps_offset
        self.betas = torch.linspace(beta_start**0.5, beta_end**0.5, num_train_timesteps, dtype=torch.float32)**2
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy().astype(np.int64))

    def set_timesteps(self, num_inference_steps):
        self.num_inference_steps = num_inference_steps
        step_ratio = self.num_train_timesteps // self.num_inference_steps
        timesteps = (np.arange(0, num_inference_steps) * step_ratio).round()[::-1].copy().astype(np.int64)
        self.timesteps = torch.from_numpy(timesteps) + self.steps_offset

    def scale_model_input(self, sample, timestep=None):
        return sample

    def step(self, model_output, timestep, sample, eta=0.0):
        prev_timestep = timestep - self.num_train_timesteps // self.num_inference_steps
        alpha_prod_t = self.alphas_cumprod[timestep]
        alpha_prod_t_prev = self.alphas_cumprod[prev_timestep] if prev_timestep >= 0 else self.final_alpha_cumprod
        beta_prod_t = 1 - alpha_prod_t
        pred_original_sample = (sample - beta_prod_t**0.5 * model_output) / alpha_prod_t**0.5
        pred_epsilon = model_output
        variance = (1 - alpha_prod_t_prev) / (1 - alpha_prod_t) * (1 - alpha_prod_t / alpha_prod_t_prev)
        std_dev_t = eta * variance**0.5
        pred_sample_direction = (1 - alpha_prod_t_prev - std_dev_t**2)**0.5 * pred_epsilon
        return alpha_prod_t_prev**0.5 * pred_original_sample + pred_sample_direction


def generate_latents(unet, scheduler, prompt_embeds, negative_prompt_embeds, height, width, num_inference_steps=50,
                     guidance_scale=3.0, seed=None, autocast_dtype=torch.bfloat16, trace=None):
    """The sampling loop of reference stable_diffusion.py:323-371 on pre-embedded prompts; returns the final latents
    (before `1 / 0.18215 * latents` and the VAE decode).  trace: optional list receiving (noise_pred, latents) per step."""
    device = prompt_embeds.device
    rng = torch.Generator(device=device)
    if seed:
        rng = rng.manual_seed(seed)
    do_cfg = guidance_scale > 1.0
    text_embeddings = prompt_embeds
    batch_size = len(text_embeddings)
    if do_cfg:
        text_embeddings = torch.cat([negative_prompt_embeds, text_embeddings])
    latents = torch.randn((batch_size, 4, height // 8, width // 8), device=device, generator=rng)
    scheduler.set_timesteps(num_inference_steps)
    latents = latents * scheduler.init_noise_sigma
    ctx = torch.autocast(device.type, dtype=autocast_dtype) if autocast_dtype is not None else torch.autocast(device.type, enabled=False)
    with torch.no_grad(), ctx:
        for t in scheduler.timesteps:
            latent_model_input = torch.cat([latents] * 2) if do_cfg else latents
            latent_model_input = scheduler.scale_model_input(latent_model_input, t)
            noise_pred = unet(latent_model_input, t, text_embeddings)['sample']
            if do_cfg:
                noise_pred_uncond, noise_pred_text = noise_pred.chunk(2)
                noise_pred = noise_pred_uncond + guidance_scale * (noise_pred_text - noise_pred_uncond)
            latents = scheduler.step(noise_pred, t, latents)
            if trace is not None:
                trace.append((noise_pred.clone(), latents.clone()))
    return latents

"""ORACLE (test infrastructure): restatement of diffusers `DDPMScheduler.{__init__,add_noise,__len__}` as the
reference builds it (`diffusion/models/models.py:88`, SD-2-base scheduler_config.json) and calls it
(`diffusion/models/stable_diffusion.py:177,180`).  Spec: SURVEY.md Appendix B2.  PARITY UNPINNED (diffusers absent).
"""
import torch


class DDPMScheduler:

    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule='scaled_linear',
                 prediction_type='epsilon'):
        assert beta_schedule == 'scaled_linear'
        self.num_train_timesteps = num_train_timesteps
        self.prediction_type = prediction_type
        self.betas = torch.linspace(beta_start**0.5, beta_end**0.5, num_train_timesteps, dtype=torch.float32)**2
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)

    def __len__(self):
        return self.num_train_timesteps

    def add_noise(self, original_samples, noise, timesteps):
        # alphas_cumprod is cast to the *sample dtype* before the sqrt (low-precision sqrt for fp16/bf16 latents)
        ac = self.alphas_cumprod.to(device=original_samples.device, dtype=original_samples.dtype)
        timesteps = timesteps.to(original_samples.device)
        a = ac[timesteps]**0.5
        a = a.flatten()
        while a.dim() < original_samples.dim():
            a = a.unsqueeze(-1)
        s = (1 - ac[timesteps])**0.5
        s = s.flatten()
        while s.dim() < original_samples.dim():
            s = s.unsqueeze(-1)
        return a * original_samples + s * noise

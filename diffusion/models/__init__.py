from diffusion_b200.model import StableDiffusion, stable_diffusion_2

__all__ = ['StableDiffusion', 'stable_diffusion_2']

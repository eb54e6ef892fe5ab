"""Drop-in for reference diffusion/models/stable_diffusion.py:15."""
from diffusion_b200.model import StableDiffusion

__all__ = ['StableDiffusion']

"""Drop-in for reference diffusion/models/models.py:28-112 (the yaml `_target_`)."""
from diffusion_b200.model import stable_diffusion_2

__all__ = ['stable_diffusion_2']

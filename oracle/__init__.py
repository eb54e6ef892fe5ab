"""ORACLE — test infrastructure only.

CPU restatement of the reference hot path (fanzhongyi/diffusion: StableDiffusion.forward/loss on
precomputed latents).  PARITY UNPINNED at the diffusers/composer boundary (see oracle/unet.py header and
DESIGN.md).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this package; the product (diffusion_b200/) never does.
"""

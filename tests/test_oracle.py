"""CPU tests of the oracle (the CPU restatement of the reference step) against every anchor available:
Random123 known-answer vectors for Philox4x32-10, the published SD-2 UNet parameter inventory, an independent
sinusoidal-embedding restatement, the SD scheduler constants and the committed golden fixture."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import philox
from oracle.ddpm import DDPMScheduler
from oracle.unet import SD2_BASE_UNET_CONFIG, TINY_UNET_CONFIG, UNet2DConditionModel, get_timestep_embedding

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'tiny_step.pt')


def test_philox_random123_kat():
    # Random123 kat_vectors, philox4x32 R=10
    kat = [
        ([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
        ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
        ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0], [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
    ]
    for ctr, key, want in kat:
        assert [int(x) for x in philox.philox4x32_10(ctr, key)] == want


def test_aten_execution_policy():
    # ATen calc_execution_policy on a 148-SM, 2048-threads/SM device
    assert philox.aten_grid(16) == (1, 4)
    assert philox.aten_grid(16 * 4 * 32 * 32) == (256, 4)
    assert philox.aten_grid(16 * 4 * 64 * 64) == (1024, 4)
    assert philox.aten_grid(148 * 8 * 256 * 4 + 1) == (1184, 8)


def test_randint_stream_properties():
    a, off = philox.randint_cuda(17, 0, 16, 1000)
    assert off == 4 and a.min() >= 0 and a.max() < 1000
    b, _ = philox.randint_cuda(17, 0, 300, 1000)  # 2 blocks: thread ids continue
    assert (b[:16] == a).all()
    c, _ = philox.randint_cuda(18, 0, 16, 1000)
    assert (c != a).any()


def test_unet_parameter_inventory():
    with torch.device('meta'):
        m = UNet2DConditionModel(**SD2_BASE_UNET_CONFIG)
    ps = list(m.named_parameters())
    assert sum(p.numel() for _, p in ps) == 865_910_724  # published SD-2 UNet size
    assert len(ps) == 686
    names = {n for n, _ in ps}
    for n in ('time_embedding.linear_1.weight', 'down_blocks.0.attentions.1.transformer_blocks.0.attn2.to_k.weight',
              'down_blocks.2.downsamplers.0.conv.bias', 'mid_block.attentions.0.proj_out.weight',
              'up_blocks.1.resnets.2.conv_shortcut.weight', 'up_blocks.2.upsamplers.0.conv.weight',
              'up_blocks.3.attentions.2.transformer_blocks.0.ff.net.0.proj.weight', 'conv_norm_out.weight', 'conv_out.bias'):
        assert n in names, n
    shapes = dict((n, tuple(p.shape)) for n, p in ps)
    assert shapes['up_blocks.1.resnets.2.conv1.weight'] == (1280, 1920, 3, 3)
    assert shapes['down_blocks.0.attentions.0.transformer_blocks.0.ff.net.0.proj.weight'] == (2560, 320)
    assert shapes['down_blocks.1.attentions.0.transformer_blocks.0.attn2.to_v.weight'] == (640, 1024)


def test_tiny_parameter_inventory():
    m = UNet2DConditionModel(**TINY_UNET_CONFIG)
    assert sum(p.numel() for p in m.parameters()) == 38_800_580 and len(list(m.parameters())) == 686


def test_timestep_embedding_independent():
    t = torch.tensor([0, 1, 37, 999])
    emb = get_timestep_embedding(t, 320, True, 0)
    half = 160
    freqs = np.exp(-math.log(10000.0) * np.arange(half) / half)
    arg = t.numpy()[:, None].astype(np.float64) * freqs[None]
    want = np.concatenate([np.cos(arg), np.sin(arg)], axis=1)  # flip_sin_to_cos -> [cos | sin]
    assert np.abs(emb.numpy() - want).max() < 2e-4
    assert emb.shape == (4, 320) and float(emb[0, 0]) == 1.0 and float(emb[0, 160]) == 0.0


def test_ddpm_scheduler_constants():
    s = DDPMScheduler()
    assert len(s) == 1000
    betas = np.linspace(0.00085**0.5, 0.012**0.5, 1000, dtype=np.float64)**2
    ac = np.cumprod(1 - betas)
    assert np.abs(s.alphas_cumprod.numpy() - ac).max() < 1e-5
    assert abs(float(s.alphas_cumprod[0]) - 0.99915) < 1e-6 and abs(float(s.alphas_cumprod[999]) - 0.00466) < 1e-4
    x0, eps, t = torch.randn(3, 4, 8, 8), torch.randn(3, 4, 8, 8), torch.tensor([0, 500, 999])
    out = s.add_noise(x0, eps, t)
    want = ac[t.numpy()][:, None, None, None]**0.5 * x0.numpy() + (1 - ac[t.numpy()][:, None, None, None])**0.5 * eps.numpy()
    assert np.abs(out.numpy() - want).max() < 1e-5
    # low-precision semantics: the table is cast to the sample dtype before the sqrt
    o16 = s.add_noise(x0.bfloat16(), eps.bfloat16(), t)
    assert o16.dtype == torch.bfloat16 and (o16.float() - out).abs().max() < 0.2  # bf16(0.99915) == 1.0: the low-precision table drops the noise term at small t


def test_golden_fixture():
    import sys
    sys.path.insert(0, os.path.dirname(GOLDEN))
    from make_golden import compute
    want = torch.load(GOLDEN)
    got = compute()
    assert torch.allclose(got['loss'], want['loss'], rtol=1e-5)
    assert torch.allclose(got['pred_slice'], want['pred_slice'], rtol=1e-4, atol=1e-5)
    assert torch.allclose(got['pred_sum'], want['pred_sum'], rtol=1e-4, atol=1e-3)
    for n, v in want['grad_norms'].items():
        assert torch.allclose(got['grad_norms'][n], v, rtol=1e-4), n
    assert torch.allclose(got['grad_conv_out_bias'], want['grad_conv_out_bias'], rtol=1e-4, atol=1e-7)
    assert torch.equal(got['randint_seed17_off0_B16'], want['randint_seed17_off0_B16'])
    assert torch.equal(got['randint_seed123_off8_B300'], want['randint_seed123_off8_B300'])
    assert torch.allclose(got['alphas_cumprod_0_499_999'], want['alphas_cumprod_0_499_999'])


def test_golden_fixture_next_rows():
    """DDIM steps and the tiny VAE oracle against tests/golden/next_rows.pt (generator: tests/golden/make_golden.py)."""
    import sys
    sys.path.insert(0, os.path.dirname(GOLDEN))
    from make_golden import compute_next_rows
    want = torch.load(os.path.join(os.path.dirname(GOLDEN), 'next_rows.pt'))
    got = compute_next_rows()
    assert set(got) == set(want)
    assert torch.equal(got['ddim_timesteps_4'], want['ddim_timesteps_4']) and want['ddim_timesteps_4'].tolist() == [751, 501, 251, 1]
    for k in ('ddim_prev_981', 'ddim_prev_501', 'ddim_prev_1'):
        assert torch.allclose(got[k], want[k], rtol=1e-6, atol=1e-6), k
    assert torch.allclose(got['vae_moments'], want['vae_moments'], rtol=1e-4, atol=1e-5)
    assert torch.allclose(got['vae_decode_slice'], want['vae_decode_slice'], rtol=1e-4, atol=1e-5)
    assert torch.allclose(got['vae_decode_sum'], want['vae_decode_sum'], rtol=1e-4, atol=1e-2)


def test_vae_oracle_inventory_and_downsample_semantics():
    """83,653,863 parameters / 248 tensors for the SD-2 VAE (the published size of AutoencoderKL) with the product
    skeleton's names; Downsample2D pads bottom / right only."""
    from diffusion_b200.encoders import AutoencoderKL
    from oracle.vae import AutoencoderKLOracle
    o = AutoencoderKLOracle()
    shapes = {n: tuple(p.shape) for n, p in o.named_parameters()}
    assert sum(p.numel() for p in o.parameters()) == 83653863 and len(shapes) == 248
    assert shapes == {n: tuple(p.shape) for n, p in AutoencoderKL().named_parameters()}
    x = torch.zeros(1, 128, 4, 4)
    x[0, 0, 3, 3] = 1.0  # bottom-right pixel: seen by the last output position through the (0,1,0,1) padding
    conv = o.encoder.down_blocks[0].downsamplers[0].conv
    y = conv(torch.nn.functional.pad(x, (0, 1, 0, 1)))
    assert y.shape == (1, 128, 2, 2)
    assert torch.allclose(y[0, :, 1, 1] - conv.bias, conv.weight[:, 0, 1, 1], atol=1e-6)


def test_text_skeleton_matches_transformers_names():
    from transformers import CLIPTextConfig
    from transformers import CLIPTextModel as HFText
    from diffusion_b200.encoders import SD2_TEXT_CONFIG, CLIPTextModel
    cfg = dict(SD2_TEXT_CONFIG)
    cfg['num_hidden_layers'] = 1
    a = {n: tuple(p.shape) for n, p in HFText(CLIPTextConfig(**cfg, projection_dim=512)).named_parameters()}
    b = {n: tuple(p.shape) for n, p in CLIPTextModel(**cfg).named_parameters()}
    assert a == b

"""GPU parity of the in-loop encoders through the C ABI: wide-image implicit conv tiles, the VAE
encoder / decoder against the torch restatement (oracle/vae.py), the text tower against transformers' own
CLIPTextModel (the reference's actual dependency, installed in this image), and the in-loop training step
(RNG order: VAE noise -> timesteps -> training noise).  The product computes both networks in bf16 with fp32
accumulation; the oracles run in fp32, so the tolerances below are bf16 round-off through ~30 layers."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = 'cuda'


@pytest.fixture(scope='module')
def ctx():
    from diffusion_b200 import ops
    return ops.get_ctx(torch.device('cuda', 0))


def _rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm()).item()


def _cos(a, b):
    return F.cosine_similarity(a.float().flatten(), b.float().flatten(), dim=0).item()


@pytest.mark.parametrize('B,H,W,Cin,Cout', [(1, 4, 256, 64, 64), (2, 8, 512, 64, 128), (1, 256, 256, 128, 64)])
def test_conv3x3_wide_rows(ctx, B, H, W, Cin, Cout):
    """W > 128: the 128-pixel tile is a segment of one image row (VAE resolutions)."""
    from diffusion_b200 import ops
    g = torch.Generator(device=DEV).manual_seed(W + Cin)
    x = (torch.randn(B, Cin, H, W, device=DEV, generator=g) * 0.5).to(torch.bfloat16)
    w = (torch.randn(Cout, Cin, 3, 3, device=DEV, generator=g) * 0.05).to(torch.bfloat16)
    bias = torch.randn(Cout, device=DEV, generator=g)
    ref = F.conv2d(x.float(), w.float(), bias, padding=1)
    xn = x.permute(0, 2, 3, 1).reshape(B * H * W, Cin).contiguous()
    w9 = w.permute(2, 3, 0, 1).reshape(9, Cout, Cin).contiguous()
    out = torch.empty(B * H * W, Cout, dtype=torch.bfloat16, device=DEV)
    ws = torch.empty(64 << 20, dtype=torch.uint8, device=DEV)
    ops.conv3x3_fwd(ctx, xn, B, H, W, w9, out, bias=bias, workspace=ws)
    got = out.view(B, H, W, Cout).permute(0, 3, 1, 2).float()
    assert _rel(got, ref) < 6e-3, _rel(got, ref)
    # borders are the cases the segment offsets could get wrong
    assert torch.allclose(got[..., :, :2], ref[..., :, :2], atol=5e-2, rtol=2e-2)
    assert torch.allclose(got[..., :, -2:], ref[..., :, -2:], atol=5e-2, rtol=2e-2)
    assert torch.allclose(got[..., :, 126:130], ref[..., :, 126:130], atol=5e-2, rtol=2e-2)


@pytest.mark.parametrize('B,H,W,C', [(2, 16, 16, 64), (1, 64, 512, 64)])
def test_vae_downsample_taps(ctx, B, H, W, C):
    """F.pad(x, (0,1,0,1)) + stride-2 pad-0 conv == phase split + shifted taps with TMA zero fill."""
    from diffusion_b200 import ops
    g = torch.Generator(device=DEV).manual_seed(H)
    x = (torch.randn(B, C, H, W, device=DEV, generator=g) * 0.5).to(torch.bfloat16)
    w = (torch.randn(C, C, 3, 3, device=DEV, generator=g) * 0.05).to(torch.bfloat16)
    bias = torch.randn(C, device=DEV, generator=g)
    ref = F.conv2d(F.pad(x.float(), (0, 1, 0, 1)), w.float(), bias, stride=2)
    xn = x.permute(0, 2, 3, 1).reshape(B * H * W, C).contiguous()
    Ho, Wo = H // 2, W // 2
    planes = torch.empty(4 * B * Ho * Wo, C, dtype=torch.bfloat16, device=DEV)
    ops.phase_split(ctx, xn, planes, B, H, W)
    w9 = w.permute(2, 3, 0, 1).reshape(9, C, C).contiguous()
    out = torch.empty(B * Ho * Wo, C, dtype=torch.bfloat16, device=DEV)
    ws = torch.empty(64 << 20, dtype=torch.uint8, device=DEV)
    ops.conv3x3_fwd(ctx, planes, B, Ho, Wo, w9, out, bias=bias, taps=ops.taps_stride2_vae(B), n_planes=4 * B, workspace=ws)
    got = out.view(B, Ho, Wo, C).permute(0, 3, 1, 2).float()
    assert got.shape == ref.shape and _rel(got, ref) < 6e-3, _rel(got, ref)


def test_glue_kernels(ctx):
    from diffusion_b200 import ops
    g = torch.Generator(device=DEV).manual_seed(0)
    # images <-> NHWC8
    img = torch.randn(2, 3, 16, 24, device=DEV, generator=g)
    x8 = torch.empty(2 * 16 * 24, 8, dtype=torch.bfloat16, device=DEV)
    ops.nchw_to_nhwc8(ctx, img.half(), x8, 2, 3, 16, 24)
    want = torch.zeros(2, 16, 24, 8, device=DEV)
    want[..., :3] = img.half().float().permute(0, 2, 3, 1)
    assert torch.equal(x8.view(2, 16, 24, 8), want.to(torch.bfloat16))
    back = torch.empty(2, 3, 16, 24, device=DEV)
    ops.nhwc8_to_nchw(ctx, x8, back, 2, 3, 16, 24, scale=0.5, shift=0.5, lo=0.0, hi=1.0)
    assert torch.allclose(back, (img.half().to(torch.bfloat16).float() / 2 + 0.5).clamp(0, 1), atol=1e-6)
    # token + position embedding
    tok, pos = torch.randn(100, 64, device=DEV, generator=g), torch.randn(77, 64, device=DEV, generator=g)
    ids = torch.randint(0, 100, (3, 77), device=DEV, generator=g)
    x = torch.empty(3 * 77, 64, dtype=torch.bfloat16, device=DEV)
    ops.embed_tokens(ctx, ids.reshape(-1), tok, pos, x, 77)
    assert torch.equal(x.view(3, 77, 64), (tok[ids] + pos[None]).to(torch.bfloat16))
    # causal softmax
    S = torch.randn(4, 77, 80, device=DEV, generator=g)
    P = torch.empty(4, 77, 80, dtype=torch.bfloat16, device=DEV)
    ops.softmax_causal_fwd(ctx, S, P, 4 * 77, 77, 77)
    mask = torch.full((77, 77), float('-inf'), device=DEV).triu(1)
    ref = torch.softmax(S[..., :77] + mask, dim=-1)
    assert torch.allclose(P[..., :77].float(), ref, atol=4e-3) and P[..., 77:].abs().max().item() == 0
    # gelu, per-pixel 1x1
    a = torch.randn(8, 4096, device=DEV, generator=g).to(torch.bfloat16)
    y = torch.empty_like(a)
    ops.gelu_fwd(ctx, a, y)
    assert torch.allclose(y.float(), F.gelu(a.float()), atol=8e-3, rtol=8e-3)
    w, b = torch.randn(4, 4, device=DEV, generator=g), torch.randn(4, device=DEV, generator=g)
    z8 = torch.zeros(50, 8, dtype=torch.bfloat16, device=DEV)
    z8[:, :4] = torch.randn(50, 4, device=DEV, generator=g).to(torch.bfloat16)
    o8 = torch.empty_like(z8)
    ops.pixel_linear8(ctx, z8, w, b, o8)
    assert torch.allclose(o8[:, :4].float(), z8[:, :4].float() @ w.t() + b, atol=2e-2, rtol=1e-2) and o8[:, 4:].abs().max().item() == 0


def _vae_pair(cfg):
    from diffusion_b200.encoders import AutoencoderKL
    from oracle.vae import AutoencoderKLOracle
    torch.manual_seed(11)
    oracle = AutoencoderKLOracle(**cfg).to(DEV)
    vae = AutoencoderKL(**cfg).to(DEV)
    vae.load_state_dict(oracle.state_dict())
    return oracle, vae


@pytest.mark.parametrize('B,R', [(2, 64), (1, 256)])
def test_vae_encode_matches_oracle(B, R):
    from oracle.vae import TINY_VAE_CONFIG
    oracle, vae = _vae_pair(TINY_VAE_CONFIG)
    g = torch.Generator(device=DEV).manual_seed(R)
    x = torch.rand(B, 3, R, R, device=DEV, generator=g) * 2 - 1
    with torch.no_grad():
        mean_ref = oracle.moments(x)[:, :4]
        torch.manual_seed(77)
        z_ref = oracle.encode_sample(x)
    gen = torch.cuda.default_generators[0]
    off_ref = gen.get_offset()
    torch.manual_seed(77)
    dist = vae.encode(x)['latent_dist']
    z = dist.sample()
    assert gen.get_offset() == off_ref, 'latent_dist.sample() must consume the generator like torch.randn(mean.shape)'
    assert z.shape == z_ref.shape == (B, 4, R // 8, R // 8) and z.dtype == x.dtype
    mean = dist.mode()
    assert _cos(mean, mean_ref) > 0.9995 and _rel(mean, mean_ref) < 3e-2, (_cos(mean, mean_ref), _rel(mean, mean_ref))
    assert _cos(z, z_ref) > 0.9995 and _rel(z, z_ref) < 3e-2, (_cos(z, z_ref), _rel(z, z_ref))
    # fp16 images (the reference's inputs.half()) give fp16 latents drawn from the same stream
    torch.manual_seed(77)
    zh = vae.encode(x.half())['latent_dist'].sample()
    assert zh.dtype == torch.float16 and _rel(zh, z_ref) < 3e-2


def test_vae_decode_matches_oracle():
    from oracle.vae import TINY_VAE_CONFIG
    oracle, vae = _vae_pair(TINY_VAE_CONFIG)
    g = torch.Generator(device=DEV).manual_seed(4)
    z = torch.randn(2, 4, 16, 16, device=DEV, generator=g)
    with torch.no_grad():
        ref = oracle.decode(z)
    img = vae.decode(z).sample
    assert img.shape == ref.shape == (2, 3, 128, 128) and img.dtype == z.dtype
    assert _cos(img, ref) > 0.9995 and _rel(img, ref) < 3e-2, (_cos(img, ref), _rel(img, ref))


@pytest.mark.parametrize('layers,B', [(2, 3), (4, 1)])
def test_text_encoder_matches_transformers(layers, B):
    from transformers import CLIPTextConfig
    from transformers import CLIPTextModel as HFText
    from diffusion_b200.encoders import SD2_TEXT_CONFIG, CLIPTextModel
    cfg = dict(SD2_TEXT_CONFIG)
    cfg['num_hidden_layers'] = layers
    torch.manual_seed(5)
    hf = HFText(CLIPTextConfig(**cfg, projection_dim=512)).to(DEV).eval()
    mine = CLIPTextModel(**cfg).to(DEV)
    missing = mine.load_state_dict(hf.state_dict(), strict=False)
    assert not missing.missing_keys and all('position_ids' in k for k in missing.unexpected_keys)
    g = torch.Generator(device=DEV).manual_seed(B)
    ids = torch.randint(0, cfg['vocab_size'], (B, 77), device=DEV, generator=g)
    with torch.no_grad():
        ref = hf(ids)[0]
    out = mine(ids)[0].clone()  # the engine's output buffer is reused by the next call
    assert out.shape == ref.shape == (B, 77, 1024) and out.dtype == torch.bfloat16
    assert _cos(out, ref) > 0.9995 and _rel(out, ref) < 3e-2, (_cos(out, ref), _rel(out, ref))
    # causal: changing a later token must not change earlier positions
    ids2 = ids.clone()
    ids2[:, 40:] = (ids2[:, 40:] + 1) % cfg['vocab_size']
    out2 = mine(ids2)[0]
    assert torch.equal(out2[:, :40], out[:, :40]) and not torch.equal(out2[:, 40:], out[:, 40:])


def test_in_loop_training_step_matches_oracle():
    """precomputed_latents=False: images + token ids -> VAE + text tower -> K1 -> UNet -> loss, against the oracle's
    in-loop branch fed the same weights; the three RNG consumers must stay aligned with the torch generator."""
    from transformers import CLIPTextConfig
    from transformers import CLIPTextModel as HFText
    from diffusion_b200.encoders import SD2_TEXT_CONFIG
    from diffusion_b200.model import stable_diffusion_2
    from oracle.stable_diffusion import StableDiffusionOracle, train_step
    from oracle.unet import TINY_UNET_CONFIG
    from oracle.vae import TINY_VAE_CONFIG, AutoencoderKLOracle
    dev = torch.device('cuda', 0)
    tcfg = dict(SD2_TEXT_CONFIG)
    tcfg['num_hidden_layers'] = 2
    torch.manual_seed(17)
    oracle = StableDiffusionOracle(TINY_UNET_CONFIG).to(dev)
    ovae = AutoencoderKLOracle(**TINY_VAE_CONFIG).to(dev)
    hf = HFText(CLIPTextConfig(**tcfg, projection_dim=512)).to(dev).eval()
    oracle.attach_encoders(ovae, hf, encode_dtype=torch.float16)
    model = stable_diffusion_2(pretrained=False, precomputed_latents=False, encode_latents_in_fp16=True, fsdp=False,
                               unet_config=TINY_UNET_CONFIG, vae_config=TINY_VAE_CONFIG, text_encoder_config=tcfg)
    model.unet.load_state_dict(oracle.unet.state_dict())
    model.vae.load_state_dict(ovae.state_dict())
    model.text_encoder.load_state_dict(hf.state_dict(), strict=False)
    assert all(not p.requires_grad for p in model.vae.parameters()) and all(not p.requires_grad for p in model.text_encoder.parameters())
    g = torch.Generator(device=dev).manual_seed(2)
    batch = {'image': torch.rand(2, 3, 128, 128, device=dev, generator=g) * 2 - 1,
             'captions': torch.randint(0, tcfg['vocab_size'], (2, 1, 77), device=dev, generator=g)}
    gen = torch.cuda.default_generators[0]
    torch.manual_seed(123)
    out = model(batch)
    loss = model.loss(out, batch)
    loss.backward()
    off = gen.get_offset()
    torch.manual_seed(123)
    lo, oo = train_step(oracle, batch, autocast_dtype=torch.bfloat16)
    assert gen.get_offset() == off, 'VAE noise + randint + randn_like must advance the generator like the reference'
    assert torch.equal(out[2], oo[2]), 'timesteps'
    assert out[1].dtype == torch.float16 and torch.equal(out[1].view(torch.int16), oo[1].view(torch.int16)), 'noise'
    assert out[0].shape == (2, 4, 16, 16)
    assert abs(loss.item() - lo.item()) <= 2e-2 * abs(lo.item()), (loss.item(), lo.item())
    cos = [F.cosine_similarity(model.unet.get_parameter(n).grad.flatten(), p.grad.float().flatten(), dim=0).item()
           for n, p in oracle.unet.named_parameters()]
    assert min(cos) > 0.98 and sum(cos) / len(cos) > 0.995, (min(cos), sum(cos) / len(cos))


def test_vae_chunked_passes_keep_one_noise_draw():
    """Batches above `max_chunk` run the encoder chunk by chunk but draw the posterior noise once, like the reference."""
    from oracle.vae import TINY_VAE_CONFIG
    _, vae = _vae_pair(TINY_VAE_CONFIG)
    g = torch.Generator(device=DEV).manual_seed(8)
    x = torch.rand(5, 3, 64, 64, device=DEV, generator=g) * 2 - 1
    torch.manual_seed(3)
    z_one = vae.encode(x)['latent_dist'].sample()
    off = torch.cuda.default_generators[0].get_offset()
    vae.max_chunk = 2  # 2 + 2 + 1
    torch.manual_seed(3)
    z_chunks = vae.encode(x)['latent_dist'].sample()
    assert torch.cuda.default_generators[0].get_offset() == off
    assert _rel(z_chunks, z_one) < 1e-2 and _cos(z_chunks, z_one) > 0.9999
    img_chunks = vae.decode(z_one.float()).sample
    vae.max_chunk = 16
    img_one = vae.decode(z_one.float()).sample
    assert img_one.shape == (5, 3, 64, 64) and _rel(img_chunks, img_one) < 1e-2


def test_generate_images_end_to_end():
    """generate(): token ids -> text tower -> DDIM/CFG loop -> VAE decode -> [0, 1] images, against the oracle pipeline."""
    from transformers import CLIPTextConfig
    from transformers import CLIPTextModel as HFText
    from diffusion_b200.encoders import SD2_TEXT_CONFIG
    from diffusion_b200.model import stable_diffusion_2
    from oracle.ddim import DDIMSchedulerOracle, generate_latents
    from oracle.stable_diffusion import StableDiffusionOracle
    from oracle.unet import TINY_UNET_CONFIG
    from oracle.vae import TINY_VAE_CONFIG, AutoencoderKLOracle
    dev = torch.device('cuda', 0)
    tcfg = dict(SD2_TEXT_CONFIG)
    tcfg['num_hidden_layers'] = 2
    torch.manual_seed(17)
    oracle = StableDiffusionOracle(TINY_UNET_CONFIG).to(dev)
    ovae = AutoencoderKLOracle(**TINY_VAE_CONFIG).to(dev)
    hf = HFText(CLIPTextConfig(**tcfg, projection_dim=512)).to(dev).eval()
    model = stable_diffusion_2(pretrained=False, precomputed_latents=False, fsdp=False, unet_config=TINY_UNET_CONFIG,
                               vae_config=TINY_VAE_CONFIG, text_encoder_config=tcfg)
    model.unet.load_state_dict(oracle.unet.state_dict())
    model.vae.load_state_dict(ovae.state_dict())
    model.text_encoder.load_state_dict(hf.state_dict(), strict=False)
    g = torch.Generator(device=dev).manual_seed(1)
    ids = torch.randint(0, tcfg['vocab_size'], (2, 77), device=dev, generator=g)
    neg = torch.randint(0, tcfg['vocab_size'], (2, 77), device=dev, generator=g)
    img = model.generate(tokenized_prompts=ids, tokenized_negative_prompts=neg, height=128, width=128, num_inference_steps=3,
                         guidance_scale=3.0, seed=7, progress_bar=False)
    assert img.shape == (2, 3, 128, 128) and img.dtype == torch.float32
    assert img.min().item() >= 0.0 and img.max().item() <= 1.0
    with torch.no_grad():
        lat = generate_latents(oracle.unet, DDIMSchedulerOracle(), hf(ids)[0], hf(neg)[0], 128, 128, 3, 3.0, seed=7)
        ref = (ovae.decode(1 / 0.18215 * lat) / 2 + 0.5).clamp(0, 1)
    assert _cos(img, ref) > 0.995 and (img - ref).abs().mean().item() < 2e-2, (_cos(img, ref), (img - ref).abs().mean().item())
    with pytest.raises(ValueError):  # no tokenizer files offline: string prompts need one
        model.generate(prompt=['a photo'], height=128, width=128, progress_bar=False)

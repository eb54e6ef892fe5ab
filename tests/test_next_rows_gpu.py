"""GPU parity of the rows either side of the training step through the C ABI: fused CFG + DDIM step and
`generate()` (f4), EMA update (f2), fp16 wire format -> training step (f3)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = 'cuda'


@pytest.fixture(scope='module')
def ctx():
    from diffusion_b200 import ops
    return ops.get_ctx(torch.device('cuda', 0))


# ------------------------------------------------------------------------------------------------ f4
@pytest.mark.parametrize('guidance', [True, False])
@pytest.mark.parametrize('B,H,W', [(2, 32, 32), (3, 8, 8), (1, 64, 64)])
def test_cfg_ddim_step_bit_exact(ctx, guidance, B, H, W):
    """Kernel == the reference loop body (stable_diffusion.py:364-371) evaluated by torch with its own type promotion."""
    from diffusion_b200 import ops
    from oracle.ddim import DDIMSchedulerOracle
    g = torch.Generator(device=DEV).manual_seed(B * 100 + H)
    nb = 2 if guidance else 1
    pred = torch.randn(nb * B, 4, H, W, device=DEV, generator=g).to(torch.bfloat16)
    latents = torch.randn(B, 4, H, W, device=DEV, generator=g)
    sch = DDIMSchedulerOracle()
    sch.set_timesteps(50)
    gs = 7.5
    for t in (981, 481, 1):
        noise_pred = pred
        if guidance:
            u, c = pred.chunk(2)
            noise_pred = u + gs * (c - u)
        want = sch.step(noise_pred, torch.tensor(t), latents)
        assert want.dtype == torch.float32
        pred8 = torch.zeros(nb * B * H * W, 8, dtype=torch.bfloat16, device=DEV)
        ops.nchw4_to_nhwc8(ctx, pred.contiguous(), pred8, nb * B, H, W)
        lat = latents.clone()
        next8 = torch.full((nb * B * H * W, 8), 7.0, dtype=torch.bfloat16, device=DEV)
        from diffusion_b200.model import DDIMScheduler
        ps = DDIMScheduler()
        ps.set_timesteps(50)
        sb, sa, sap, dc = (float(x) for x in ps.step_scalars(t))
        ops.cfg_ddim_step(ctx, pred8, lat, next8, B, H, W, guidance, gs, sb, sa, sap, dc)
        assert torch.equal(lat, want), (lat - want).abs().max().item()
        want8 = torch.zeros(B, H, W, 8, dtype=torch.bfloat16, device=DEV)
        want8[..., :4] = want.permute(0, 2, 3, 1).to(torch.bfloat16)
        for half in range(nb):
            assert torch.equal(next8.view(nb, B, H, W, 8)[half], want8)


@pytest.mark.parametrize('gs', [3.0, 1.0])
def test_generate_latents_match_oracle_loop(gs):
    from diffusion_b200.model import stable_diffusion_2
    from oracle.ddim import DDIMSchedulerOracle, generate_latents
    from oracle.stable_diffusion import StableDiffusionOracle
    from oracle.unet import TINY_UNET_CONFIG
    dev = torch.device('cuda', 0)
    torch.manual_seed(17)
    oracle = StableDiffusionOracle(TINY_UNET_CONFIG).to(dev)
    model = stable_diffusion_2(pretrained=False, precomputed_latents=True, build_encoders=False, unet_config=TINY_UNET_CONFIG, fsdp=False)
    model.unet.load_state_dict(oracle.unet.state_dict())
    g = torch.Generator(device=dev).manual_seed(9)
    emb = torch.randn(2, 77, 1024, device=dev, generator=g)
    neg = torch.randn(2, 77, 1024, device=dev, generator=g)
    steps = 4
    got = model.generate(prompt_embeds=emb, negative_prompt_embeds=neg, height=256, width=256, num_inference_steps=steps,
                         guidance_scale=gs, seed=1138, progress_bar=False, output_type='latent')
    trace = []
    want = generate_latents(oracle.unet, DDIMSchedulerOracle(), emb, neg, 256, 256, steps, gs, seed=1138, trace=trace)
    assert got.shape == want.shape == (2, 4, 32, 32) and got.dtype == torch.float32
    sampler = [e for k, e in model.unet._engines.items() if k[-1]]
    assert sampler and all(e.forward_only and len(e.bwd) == 0 and e.graph_fwd is not None for e in sampler)
    with pytest.raises(RuntimeError):
        sampler[0].run_backward()
    assert torch.isfinite(got).all()
    cos = torch.nn.functional.cosine_similarity(got.flatten(), want.flatten(), dim=0).item()
    rel = ((got - want).norm() / want.norm()).item()
    assert cos > 0.999 and rel < 3e-2, (cos, rel)
    # same seed -> same images; other seed -> other images (the torch generator is the reference's)
    again = model.generate(prompt_embeds=emb, negative_prompt_embeds=neg, height=256, width=256, num_inference_steps=steps,
                           guidance_scale=gs, seed=1138, progress_bar=False, output_type='latent')
    assert torch.equal(again, got)
    # num_images_per_prompt duplicates the embeddings like the reference
    many = model.generate(prompt_embeds=emb[:1], negative_prompt_embeds=neg[:1], height=128, width=128, num_inference_steps=2,
                          guidance_scale=gs, num_images_per_prompt=3, seed=5, progress_bar=False, output_type='latent')
    assert many.shape == (3, 4, 16, 16)


# ------------------------------------------------------------------------------------------------ f2
@pytest.mark.parametrize('n', [1, 3, 4, 1000, 1 << 20, (1 << 20) + 3])
def test_ema_update_bit_exact(ctx, n):
    from diffusion_b200 import ops
    g = torch.Generator(device=DEV).manual_seed(n)
    ema, p = torch.randn(n, device=DEV, generator=g), torch.randn(n, device=DEV, generator=g)
    s = 0.9999
    want = ema * s + p * (1. - s)  # the reference expression (ema.py:63)
    ops.ema_update(ctx, ema, p, s)
    assert torch.equal(ema, want)


def test_ema_of_the_unet_arena_matches_reference_loop():
    from diffusion_b200.ema import EMA, EMAParameters, compute_ema
    from diffusion_b200.model import stable_diffusion_2
    from diffusion_b200.optim import FusedAdamW
    from oracle.unet import TINY_UNET_CONFIG
    dev = torch.device('cuda', 0)
    torch.manual_seed(3)
    model = stable_diffusion_2(pretrained=False, precomputed_latents=True, build_encoders=False, unet_config=TINY_UNET_CONFIG, fsdp=False)
    g = torch.Generator(device=dev).manual_seed(5)
    batch = {'image_latents': torch.randn(2, 4, 16, 16, device=dev, generator=g).to(torch.bfloat16),
             'caption_latents': torch.randn(2, 77, 1024, device=dev, generator=g).to(torch.bfloat16)}
    opt = FusedAdamW(model.parameters(), lr=1e-3)
    model.loss(model(batch), batch).backward()  # builds the engine arena the parameters live in
    ema = EMA(half_life=None, smoothing=0.75)
    ref = {n: p.detach().clone() for n, p in model.named_parameters()}
    assert ema.update(model, batch=0) is True  # starts (ema_start 0.0dur) and averages once: ema == params
    assert len(ema.ema_model._flat) == 1, 'the UNet arena must be averaged by one flat launch'
    for n, p in model.named_parameters():
        ref[n] = ref[n] * 0.75 + p.detach() * (1. - 0.75)
    for step in range(1, 3):
        opt.step()
        opt.zero_grad()
        model.loss(model(batch), batch).backward()
        ema.update(model, batch=step)
        for n, p in model.named_parameters():
            ref[n] = ref[n] * 0.75 + p.detach() * (1. - 0.75)
    got = dict(ema.ema_model.named_parameters())
    assert set(got) == set(ref)
    for n in ref:
        assert torch.equal(got[n], ref[n]), n
    # swap in / out like the EVAL_START / EVAL_END events
    before = {n: p.detach().clone() for n, p in model.named_parameters()}
    ema.ema_model.swap_params(model)
    for n, p in model.named_parameters():
        assert torch.equal(p.detach(), ref[n])
    ema.ema_model.swap_params(model)
    for n, p in model.named_parameters():
        assert torch.equal(p.detach(), before[n])
    # module-to-module form of compute_ema
    a, b = torch.nn.Linear(8, 8).to(dev), torch.nn.Linear(8, 8).to(dev)
    want = {k: v * 0.5 + dict(a.named_parameters())[k].detach() * 0.5 for k, v in b.state_dict().items()}
    compute_ema(a, b, 0.5)
    for k, v in b.state_dict().items():
        assert torch.equal(v, want[k])
    assert isinstance(EMAParameters(None).named_parameters_dict, dict)


# ------------------------------------------------------------------------------------------------ f3
def test_cast_to_bf16(ctx):
    from diffusion_b200 import ops
    g = torch.Generator(device=DEV).manual_seed(1)
    for dt in (torch.float16, torch.float32, torch.bfloat16):
        for n in (1, 7, 77 * 1024 * 3):
            src = torch.randn(n, device=DEV, generator=g).to(dt)
            dst = torch.empty(n, dtype=torch.bfloat16, device=DEV)
            ops.cast_to_bf16(ctx, src, dst)
            assert torch.equal(dst, src.to(torch.bfloat16))


def test_fp16_wire_batch_trains_like_the_oracle():
    """Dataset bytes -> LatentBatcher -> model.forward/loss/backward, against the oracle fed the same fp16 tensors."""
    from diffusion_b200.model import stable_diffusion_2
    from diffusion_b200.wire import LatentBatcher
    from oracle.stable_diffusion import StableDiffusionOracle
    from oracle.unet import TINY_UNET_CONFIG
    dev = torch.device('cuda', 0)
    torch.manual_seed(17)
    oracle = StableDiffusionOracle(TINY_UNET_CONFIG).to(dev)
    model = stable_diffusion_2(pretrained=False, precomputed_latents=True, unet_config=TINY_UNET_CONFIG, fsdp=False,
                               build_encoders=False)
    model.unet.load_state_dict(oracle.unet.state_dict())
    rng = np.random.default_rng(0)
    samples = [{'latents_256': (rng.standard_normal((4, 32, 32)) * 0.8).astype(np.float16).tobytes(),
                'caption_latents': rng.standard_normal((77, 1024)).astype(np.float16).tobytes()} for _ in range(3)]
    bt = LatentBatcher(4, image_size=256, device=dev)
    batch = bt.to_device(bt.collate(samples))
    assert batch['image_latents'].dtype == torch.float16 and batch['image_latents'].shape == (3, 4, 32, 32)
    assert batch['caption_latents'].is_cuda
    import parity
    res = parity.step_triplet(TINY_UNET_CONFIG, 3, 32, pair=(oracle, model, batch))
    assert res['out'][1].dtype == torch.float16
    assert not parity.gate_failures(res), parity.gate_failures(res)[:8]

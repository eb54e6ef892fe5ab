"""The reference's own test file (reference tests/test_model.py:13-48) against the B200 implementation: same factory
call, same batch keys, same assertions (shapes only - that is all the reference asserts), on the smallest geometries
the tiled kernels support (the reference uses 8x8 images = 1x1 latents, which needs diffusers' odd-size up-sampling
path; here 64x64 images = 8x8 latents, whose lowest UNet level is also 1x1).  Differences forced by the environment:
prompts are token ids (no tokenizer vocabulary on disk) and the model lives on the GPU."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def model():
    from diffusion.models.models import stable_diffusion_2  # the reference import path
    torch.manual_seed(0)
    return stable_diffusion_2(pretrained=False, fsdp=False, encode_latents_in_fp16=False, precomputed_latents=False)


def test_model_forward(model):
    batch_size, H, W = 1, 64, 64
    dev = next(model.parameters()).device
    image = torch.randn(batch_size, 3, H, W, device=dev)
    latent = torch.randn(batch_size, 4, H // 8, W // 8)
    caption = torch.randint(low=0, high=128, size=(batch_size, 77), dtype=torch.long, device=dev)
    batch = {'image': image, 'captions': caption}
    output, target, _ = model(batch)  # model.forward generates the unet output noise or v_pred target.
    assert output.shape == latent.shape
    assert target.shape == latent.shape
    assert torch.isfinite(output.float()).all()
    loss = model.loss((output, target, _), batch)
    loss.backward()
    assert torch.isfinite(loss) and all(p.grad is not None and torch.isfinite(p.grad).all() for p in model.unet.parameters())


@pytest.mark.parametrize('guidance_scale', [0.0, 3.0])
@pytest.mark.parametrize('negative_prompt', [None, 'so cool'])
def test_model_generate(model, guidance_scale, negative_prompt):
    dev = next(model.parameters()).device
    g = torch.Generator(device=dev).manual_seed(1)
    prompt = torch.randint(0, 128, (1, 77), device=dev, generator=g)  # 'a cool doge', tokenized
    neg = torch.randint(0, 128, (1, 77), device=dev, generator=g) if (negative_prompt is not None or guidance_scale > 1.0) else None
    output = model.generate(
        tokenized_prompts=prompt,
        tokenized_negative_prompts=neg,
        num_inference_steps=1,
        num_images_per_prompt=1,
        height=64,
        width=64,
        guidance_scale=guidance_scale,
        progress_bar=False,
    )
    assert output.shape == (1, 3, 64, 64)
    assert torch.isfinite(output).all() and output.min() >= 0 and output.max() <= 1

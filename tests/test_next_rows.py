"""CPU tests of the rows either side of the training step: wire format collate (f3), EMA host logic
(f2), DDIM scheduler + generate() argument handling (f4).  No kernel runs here."""
import math

import numpy as np
import pytest
import torch


# ------------------------------------------------------------------------------------------------ f3: wire format
def _samples(n, size=256, seed=0):
    rng = np.random.default_rng(seed)
    shape = (4, 32, 32) if size == 256 else (4, 64, 64)
    out = []
    for _ in range(n):
        lat = rng.standard_normal(shape).astype(np.float16)
        cap = rng.standard_normal((77, 1024)).astype(np.float16)
        out.append({f'latents_{size}': lat.tobytes(), 'caption_latents': cap.tobytes(), 'jpg': b'', 'caption': 'x'})
    return out


@pytest.mark.parametrize('size', [256, 512])
def test_wire_collate_matches_reference_reader(size):
    """Same tensors as the reference reader (laion.py:103-111: np.frombuffer(...).copy().reshape) + default collate."""
    from diffusion_b200.wire import LatentBatcher
    samples = _samples(5, size)
    bt = LatentBatcher(8, image_size=size, pin=False)
    batch = bt.collate(samples)  # ragged last batch: 5 of 8
    shape = (4, 32, 32) if size == 256 else (4, 64, 64)
    ref_lat = torch.stack([torch.from_numpy(np.frombuffer(s[f'latents_{size}'], dtype=np.float16).copy()).reshape(shape) for s in samples])
    ref_cap = torch.stack([torch.from_numpy(np.frombuffer(s['caption_latents'], dtype=np.float16).copy()).reshape(77, 1024) for s in samples])
    assert batch['image_latents'].dtype == torch.float16 and batch['image_latents'].shape == ref_lat.shape
    assert torch.equal(batch['image_latents'], ref_lat) and torch.equal(batch['caption_latents'], ref_cap)
    # samples already decoded by the reference __getitem__ (tensors) are accepted too, slots rotate
    dec = [{'image_latents': ref_lat[i], 'caption_latents': ref_cap[i]} for i in range(5)]
    b2 = bt.collate(dec)
    assert b2['image_latents'].data_ptr() != batch['image_latents'].data_ptr()
    assert torch.equal(b2['image_latents'], ref_lat) and torch.equal(b2['caption_latents'], ref_cap)


def test_wire_collate_rejects_bad_input():
    from diffusion_b200.wire import LatentBatcher
    bt = LatentBatcher(2, image_size=256, pin=False)
    with pytest.raises(ValueError):
        bt.collate([])
    with pytest.raises(ValueError):
        bt.collate(_samples(3))
    bad = _samples(1)
    bad[0]['latents_256'] = bad[0]['latents_256'][:-2]
    with pytest.raises(ValueError):
        bt.collate(bad)
    with pytest.raises(KeyError):
        bt.collate([{'caption_latents': _samples(1)[0]['caption_latents']}])
    with pytest.raises(ValueError):
        bt.collate([{'image_latents': torch.zeros(4, 32, 32), 'caption_latents': torch.zeros(77, 1024)}])  # fp32 is not the wire dtype
    with pytest.raises(ValueError):
        LatentBatcher(2, image_size=384)
    with pytest.raises(RuntimeError):
        bt.to_device(bt.collate(_samples(2)), 'cpu')


# ------------------------------------------------------------------------------------------------ f2: EMA host logic
def test_ema_constructor_contract():
    """Validation and smoothing formula of the reference algorithm (ema.py:133-185)."""
    from diffusion_b200.ema import EMA
    assert EMA().smoothing == 2**(-(1 / 1000))
    assert EMA(half_life='100ba', update_interval='10ba').smoothing == 2**(-(10 / 100))
    e = EMA(half_life=None, smoothing=0.9999)
    assert e.smoothing == 0.9999 and e.update_interval == (1, 'ba') and e.update_event == 'BATCH_END'
    assert EMA(half_life='2ep').update_event == 'EPOCH_END'
    for kw in (dict(half_life=None), dict(half_life='10ba', smoothing=0.5), dict(half_life='10ba', update_interval='1ep'),
               dict(half_life=None, smoothing=0.5, update_interval='0.5dur'), dict(half_life='abc'),
               dict(half_life=None, smoothing=0.5, update_interval=3)):
        with pytest.raises(ValueError):
            EMA(**kw)


def test_ema_has_no_cpu_path():
    from diffusion_b200.ema import EMA, EMAParameters, compute_ema
    m = torch.nn.Linear(4, 4)
    with pytest.raises(RuntimeError):
        compute_ema(m, EMAParameters(m), 0.5)
    with pytest.raises(ValueError):
        compute_ema(m, object(), 0.5)
    e = EMA(half_life=None, smoothing=0.5, ema_start='3ba')
    assert e.update(m, batch=1) is False and e.ema_model is None  # not started yet


def test_ema_matches_checkpoint_events_when_a_saver_is_about_to_write():
    """reference ema.py:221-227, 276-279: on BATCH_CHECKPOINT / EPOCH_CHECKPOINT the algorithm matches iff EMA has started
    and a checkpoint saver's save_interval(state, event) is True; apply() then swaps the EMA weights into the model."""
    from diffusion_b200.ema import EMA

    class Saver:
        def __init__(self, due):
            self.due = due

        def save_interval(self, state, event):
            return self.due

    class State:
        callbacks = []
        model = None

    class Params:
        swaps = 0

        def swap_params(self, model):
            Params.swaps += 1

    e = EMA(half_life=None, smoothing=0.5)
    st = State()
    st.callbacks = [object(), Saver(True)]
    assert e.match('BATCH_CHECKPOINT', st) is False  # not started
    e.ema_started, e.ema_model = True, Params()
    assert e.match('BATCH_CHECKPOINT', st) is True and e.match('EPOCH_CHECKPOINT', st) is True
    st.callbacks = [Saver(False)]
    assert e.match('BATCH_CHECKPOINT', st) is False
    st.callbacks = []
    assert e.match('EPOCH_CHECKPOINT', st) is False
    e.apply('BATCH_CHECKPOINT', st)
    assert Params.swaps == 1 and e.ema_weights_active is True
    e.apply('BATCH_START', st)  # training resumes on the training weights
    assert Params.swaps == 2 and e.ema_weights_active is False


# ------------------------------------------------------------------------------------------------ f4: DDIM / generate
def test_ddim_scheduler_matches_oracle_and_known_answers():
    from diffusion_b200.model import DDIMScheduler
    from oracle.ddim import DDIMSchedulerOracle
    a, b = DDIMScheduler(), DDIMSchedulerOracle()
    for n in (50, 20, 4, 1000):
        a.set_timesteps(n)
        b.set_timesteps(n)
        assert torch.equal(a.timesteps, b.timesteps)
    a.set_timesteps(50)
    b.set_timesteps(50)
    assert a.timesteps[0].item() == 981 and a.timesteps[-1].item() == 1 and len(a.timesteps) == 50
    assert a.init_noise_sigma == 1.0
    # SD-2-base schedule constants (scaled_linear 0.00085..0.012): alphas_cumprod[0] = 1 - 0.00085
    assert abs(a.alphas_cumprod[0].item() - (1 - 0.00085)) < 1e-7 and abs(a.alphas_cumprod[-1].item() - 0.0046604) < 1e-6
    g = torch.Generator().manual_seed(3)
    x, e = torch.randn(2, 4, 8, 8, generator=g), torch.randn(2, 4, 8, 8, generator=g)
    for dt in (torch.float32, torch.bfloat16):
        for t in a.timesteps[[0, 17, 49]]:
            pa, pb = a.step(e.to(dt), t, x).prev_sample, b.step(e.to(dt), t, x)
            assert pa.dtype == torch.float32 and torch.equal(pa, pb)  # bf16 model output, fp32 latents -> fp32
    with pytest.raises(ValueError):
        DDIMScheduler().step(e, 981, x)  # set_timesteps not called
    with pytest.raises(ValueError):
        a.step(e, 981, x, eta=0.5)


def test_ddim_step_is_the_deterministic_ddim_update():
    """Song et al. eq. 12 with sigma = 0: if eps is the true noise of x_t, the step lands on the same x_0 / eps pair at t_prev."""
    from oracle.ddim import DDIMSchedulerOracle
    s = DDIMSchedulerOracle()
    s.set_timesteps(50)
    g = torch.Generator().manual_seed(0)
    x0, eps = torch.randn(3, 4, 8, 8, generator=g, dtype=torch.float64), torch.randn(3, 4, 8, 8, generator=g, dtype=torch.float64)
    s.alphas_cumprod = s.alphas_cumprod.double()
    s.final_alpha_cumprod = s.final_alpha_cumprod.double()
    for t in (981, 501, 21, 1):
        a_t = s.alphas_cumprod[t]
        prev = t - 20
        a_p = s.alphas_cumprod[prev] if prev >= 0 else s.final_alpha_cumprod
        x_t = a_t.sqrt() * x0 + (1 - a_t).sqrt() * eps
        want = a_p.sqrt() * x0 + (1 - a_p).sqrt() * eps
        assert torch.allclose(s.step(eps, t, x_t), want, atol=1e-12)


def test_generate_argument_contract():
    from diffusion_b200.model import DDIMScheduler, StableDiffusion, _require_some_prompt, _require_matching_negatives
    from diffusion_b200.unet import UNet2DConditionModel, UNetOutput
    from oracle.unet import TINY_UNET_CONFIG
    with pytest.raises(ValueError):
        _require_some_prompt(None, None, None)
    with pytest.raises(ValueError):
        _require_matching_negatives(['a', 'b'], ['c'])
    _require_matching_negatives(['a'], None)
    m = StableDiffusion(UNet2DConditionModel(**TINY_UNET_CONFIG), None, None, None, None, DDIMScheduler(), precomputed_latents=True)
    emb = torch.zeros(1, 77, 1024)
    with pytest.raises(ValueError):
        m.generate()  # no prompt of any kind
    with pytest.raises(ValueError):
        m.generate(prompt_embeds=emb)  # no VAE -> images cannot be decoded
    with pytest.raises(ValueError):
        m.generate(prompt=['a cat'], output_type='latent')  # no text encoder
    with pytest.raises(RuntimeError):
        m.generate(prompt_embeds=emb, guidance_scale=1.0, output_type='latent')  # CPU model: no fallback
    o = UNetOutput(sample=emb)
    assert o.sample is emb and o['sample'] is emb


def test_abi_wire_gather_host_function():
    import ctypes as C
    from diffusion_b200 import _lib
    lib = _lib.load()
    a, b = np.arange(16, dtype=np.uint8), np.arange(16, 32, dtype=np.uint8)
    dst = np.zeros(32, dtype=np.uint8)
    ptrs = (C.c_void_p * 2)(a.ctypes.data, b.ctypes.data)
    assert lib.sd2_wire_gather(C.cast(ptrs, C.c_void_p), 2, 16, dst.ctypes.data) == 0
    assert (dst == np.arange(32)).all()
    ptrs[1] = None
    assert lib.sd2_wire_gather(C.cast(ptrs, C.c_void_p), 2, 16, dst.ctypes.data) == 2
    assert lib.sd2_wire_gather(None, -1, 16, None) == 1
    assert math.isfinite(lib.sd2_version())

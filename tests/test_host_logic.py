"""Host-side logic on CPU (no kernels run): parameter skeleton == oracle inventory, flat arenas with the
[kh][kw][Cout][Cin] conv storage, the static schedule builds for both configs and every parameter receives a
gradient-producing op, gradient buckets tile the arena.  Uses the type-checking DryLib test double (SD2_DRY_RUN=1)."""
import os

import pytest
import torch

from oracle.unet import SD2_BASE_UNET_CONFIG, TINY_UNET_CONFIG
from oracle.unet import UNet2DConditionModel as OracleUNet


@pytest.fixture()
def dry(monkeypatch):
    monkeypatch.setenv('SD2_DRY_RUN', '1')
    from diffusion_b200 import ops
    ops._ctx_cache.pop('dry', None)
    yield


def test_skeleton_matches_oracle_names_and_shapes():
    from diffusion_b200.unet import UNet2DConditionModel
    with torch.device('meta'):
        a = UNet2DConditionModel(**SD2_BASE_UNET_CONFIG)
        b = OracleUNet(**SD2_BASE_UNET_CONFIG)
    sa = {n: tuple(p.shape) for n, p in a.named_parameters()}
    sb = {n: tuple(p.shape) for n, p in b.named_parameters()}
    assert sa == sb and len(sa) == 686 and sum(p.numel() for p in a.parameters()) == 865_910_724


def test_reference_import_paths():
    from diffusion.models import StableDiffusion
    from diffusion.models.models import stable_diffusion_2
    from diffusion.models.stable_diffusion import StableDiffusion as SD2
    import inspect
    assert SD2 is StableDiffusion
    sig = inspect.signature(stable_diffusion_2)
    for k in ('model_name', 'pretrained', 'train_metrics', 'val_metrics', 'val_guidance_scales', 'val_seed', 'loss_bins',
              'precomputed_latents', 'encode_latents_in_fp16', 'fsdp'):  # reference models.py:28-39
        assert k in sig.parameters
    for m in ('forward', 'loss', 'eval_forward', 'get_metrics', 'update_metric'):  # ComposerModel protocol
        assert callable(getattr(StableDiffusion, m))
    with pytest.raises(ValueError):
        stable_diffusion_2(pretrained=True)


def _instantiate(block):
    """Hydra's `instantiate` for the subset the recipes use: `_target_` = dotted path, nested blocks / lists recursed.
    `torchmetrics.*` resolves to the stand-in the product ships when torchmetrics is not installed (as in this image)."""
    import importlib
    if isinstance(block, list):
        return [_instantiate(b) for b in block]
    if not isinstance(block, dict):
        return block
    kwargs = {k: _instantiate(v) for k, v in block.items() if k != '_target_'}
    if '_target_' not in block:
        return kwargs
    mod, _, attr = block['_target_'].rpartition('.')
    if mod == 'torchmetrics':
        try:
            importlib.import_module(mod)
        except ImportError:
            mod = 'diffusion_b200.model'
    return getattr(importlib.import_module(mod), attr)(**kwargs)


@pytest.mark.parametrize('recipe', ['SD-2-base-256.yaml', 'SD-2-base-512.yaml'])
def test_factory_accepts_the_reference_yaml_model_block(recipe):
    """Drop-in at the recipe level: the `model:` block of the reference yamls (`_target_:
    diffusion.models.models.stable_diffusion_2` + kwargs, fixture tests/golden/reference_model_blocks.json) instantiates
    the B200-native model through the reference import path with exactly those keyword arguments."""
    import json
    import os
    here = os.path.dirname(os.path.abspath(__file__))
    with open(os.path.join(here, 'golden', 'reference_model_blocks.json')) as f:
        block = json.load(f)[recipe]
    ref = f'/root/reference/yamls/hydra-yamls/{recipe}'
    if os.path.exists(ref):  # in the build container: the committed fixture must be the reference's current block
        import yaml
        with open(ref) as f:
            assert yaml.safe_load(f)['model'] == block
    assert block['_target_'] == 'diffusion.models.models.stable_diffusion_2'
    assert block['precomputed_latents'] is True and block['fsdp'] is True and block['pretrained'] is False
    model = _instantiate(block)
    from diffusion.models.stable_diffusion import StableDiffusion
    assert isinstance(model, StableDiffusion)
    assert sum(p.numel() for p in model.unet.parameters()) == 865910724
    assert len(list(model.unet.named_parameters())) == 686
    # like the reference (models.py:80-85) the recipe's model carries frozen VAE + text encoder for the image/caption eval set
    assert model.vae is not None and model.text_encoder is not None
    assert not any(p.requires_grad for p in model.vae.parameters())
    assert not any(p.requires_grad for p in model.text_encoder.parameters())
    assert model.precomputed_latents and model.encode_latents_in_fp16
    # val_guidance_scales: [] / loss_bins: [] in the recipe -> only the full-loss MSE metric remains (reference :111-133)
    assert list(model.get_metrics(is_train=False)) == ['MeanSquaredError']
    assert list(model.get_metrics(is_train=True)) == ['MeanSquaredError']
    assert all(getattr(m, '_fsdp_wrap', None) is False for m in (model.unet, model.vae, model.text_encoder))


def test_metric_plumbing_mirrors_the_reference():
    """reference stable_diffusion.py:111-133 (per-guidance-scale / per-bin metric copies) and :228-257 (update_metric)."""
    from diffusion_b200.model import MeanSquaredError, Metric, StableDiffusion, DDPMScheduler

    class FrechetInceptionDistance(Metric):
        def __init__(self, feature=64, **_kw):
            super().__init__()
            self.feature, self.calls = feature, []

        def update(self, imgs, real):
            self.calls.append((tuple(imgs.shape), real))

    class InceptionScore(Metric):
        def __init__(self, **_kw):
            super().__init__()
            self.calls = []

        def update(self, imgs):
            self.calls.append(tuple(imgs.shape))

    class Other(Metric):
        def update(self, a, b):
            self.seen = (a, b)

    unet = torch.nn.Linear(1, 1)
    m = StableDiffusion(unet, None, None, None, DDPMScheduler(), None,
                        val_metrics=[MeanSquaredError(), FrechetInceptionDistance(feature=192), InceptionScore(), Other()],
                        val_guidance_scales=[1.0, 7.5], loss_bins=[(0, 0.5), (0.5, 1)])
    names = list(m.get_metrics(is_train=False))
    assert names == ['MeanSquaredError-bin-0-to-0p5', 'MeanSquaredError-bin-0p5-to-1', 'FrechetInceptionDistance-scale-1p0',
                     'FrechetInceptionDistance-scale-7p5', 'InceptionScore-scale-1p0', 'InceptionScore-scale-7p5', 'Other',
                     'MeanSquaredError']
    vm = m.get_metrics(is_train=False)
    assert vm['FrechetInceptionDistance-scale-7p5'].guidance_scale == 7.5 and vm['FrechetInceptionDistance-scale-7p5'].feature == 192
    assert vm['FrechetInceptionDistance-scale-1p0'] is not vm['FrechetInceptionDistance-scale-7p5']
    pred, noise, ts = torch.ones(4, 2), torch.zeros(4, 2), torch.tensor([10, 400, 600, 999])
    imgs = {1.0: torch.rand(4, 3, 8, 8), 7.5: torch.rand(4, 3, 8, 8)}
    batch = {'image': torch.rand(4, 3, 8, 8), 'captions': torch.zeros(4, 77, dtype=torch.long)}
    out = (pred, noise, ts, imgs)
    m.update_metric(batch, out, vm['MeanSquaredError-bin-0-to-0p5'])
    assert float(vm['MeanSquaredError-bin-0-to-0p5'].total) == 4.0  # two of four samples fall in [0, 500)
    m.update_metric(batch, out, vm['MeanSquaredError'])
    assert float(vm['MeanSquaredError'].compute()) == 1.0
    m.update_metric(batch, out, vm['FrechetInceptionDistance-scale-7p5'])
    assert vm['FrechetInceptionDistance-scale-7p5'].calls == [((4, 3, 8, 8), True), ((4, 3, 8, 8), False)]
    m.update_metric(batch, out, vm['InceptionScore-scale-1p0'])
    assert vm['InceptionScore-scale-1p0'].calls == [(4, 3, 8, 8)]
    m.update_metric(batch, out, vm['Other'])
    assert vm['Other'].seen[0] is pred
    with pytest.raises(ValueError):  # image metrics without generated images fail loudly instead of mis-updating
        m.update_metric(batch, (pred, noise, ts), vm['InceptionScore-scale-1p0'])
    m.val_metrics['bad'] = object()
    with pytest.raises(TypeError):
        m.get_metrics(is_train=False)


def test_cpu_model_fails_loudly():
    from diffusion_b200.model import stable_diffusion_2
    if torch.cuda.is_available():
        pytest.skip('CPU-only check')
    m = stable_diffusion_2(pretrained=False, precomputed_latents=True, build_encoders=False, unet_config=TINY_UNET_CONFIG, fsdp=False)
    batch = {'image_latents': torch.randn(2, 4, 32, 32), 'caption_latents': torch.randn(2, 77, 1024)}
    with pytest.raises(RuntimeError):
        m(batch)


def test_arena_storage_layout(dry):
    from diffusion_b200.unet import UNet2DConditionModel
    torch.manual_seed(0)
    u = UNet2DConditionModel(**TINY_UNET_CONFIG)
    before = {n: p.detach().clone() for n, p in u.named_parameters()}
    eng = u.engine(2, 32, 32, 77)
    ar = eng.arena
    assert ar.bound()
    for n, p in u.named_parameters():
        assert torch.equal(p.detach(), before[n]), n  # values preserved, shapes unchanged
    w = u.get_parameter('down_blocks.0.resnets.0.conv1.weight')
    assert w.shape == (64, 64, 3, 3) and w.stride() == (64, 1, 3 * 64 * 64, 64 * 64)
    st = ar.storage(ar.p32, 'down_blocks.0.resnets.0.conv1.weight')
    assert st.shape == (9, 64, 64) and torch.equal(st, w.detach().permute(2, 3, 0, 1).reshape(9, 64, 64))
    names = ['down_blocks.0.attentions.0.transformer_blocks.0.attn1.' + k + '.weight' for k in ('to_q', 'to_k', 'to_v')]
    f = ar.fused(ar.p32, names)
    assert f.shape == (192, 64) and torch.equal(f[64:128], u.get_parameter(names[1]).detach())
    # a state_dict round trip keeps working through the strided views
    sd = {k: v.clone() for k, v in u.state_dict().items()}
    u.load_state_dict(sd)
    assert ar.bound()


@pytest.mark.parametrize('cfg,B,R', [(TINY_UNET_CONFIG, 2, 32), (SD2_BASE_UNET_CONFIG, 1, 32), (TINY_UNET_CONFIG, 2, 64)])
def test_schedule_builds_and_covers_every_parameter(dry, cfg, B, R):
    from diffusion_b200.unet import UNet2DConditionModel
    u = UNet2DConditionModel(**cfg)
    eng = u.engine(B, R, R, 77)
    assert set(eng.grad_ready) == set(eng.arena.entries)
    assert len(eng.fwd) > 300 and len(eng.bwd) > 600
    eng.run_forward()
    eng.run_backward()
    calls = eng.ctx.lib.calls
    # 61 GroupNorms, none with a statistics pass of its own: 49 get them from the epilogue of the conv / linear that produced
    # their input, the 12 that read a skip concatenation (norm1 of the up-path ResNets) from the concatenation kernel
    assert calls['sd2_groupnorm_fwd_fused'] == 61 and calls['sd2_concat_stats'] == 12 and 'sd2_groupnorm_fwd' not in calls
    assert calls['sd2_layernorm_fwd'] == 48 and calls['sd2_attn_fwd'] == 32 and calls['sd2_attn_bwd'] == 32
    # buckets tile the arena exactly, in completion order
    lo = sorted(b[0] for b in eng.buckets)
    hi = sorted(b[1] for b in eng.buckets)
    assert lo[0] == 0 and hi[-1] == eng.arena.total and lo[1:] == hi[:-1]
    assert [b[2] for b in eng.buckets] == sorted(b[2] for b in eng.buckets) and eng.buckets[-1][2] == len(eng.bwd)
    # the optimizer may take over a bucket (FusedAdamW.arm) only at a segment end, never before its gradients are final, and no
    # backward op after that point touches the bucket's weights or gradients; in this schedule that is the bucket's own end
    pts = eng._update_points()
    assert all(p in eng.segments and p >= b[2] for p, b in zip(pts, eng.buckets))
    from diffusion_b200.engine import _op_io, _overlap
    a = eng.arena
    for p, (lo_, hi_, _) in zip(pts, eng.buckets):
        spans = [(t.data_ptr() + lo_ * t.element_size(), t.data_ptr() + hi_ * t.element_size()) for t in (a.p32, a.p16, a.g32)]
        assert not any(_overlap(_op_io(op)[0], spans) for op in eng.bwd[p:])
    assert pts == [b[2] for b in eng.buckets]
    if cfg is SD2_BASE_UNET_CONFIG:
        per_image = eng.gemm_flops / B / 1e12
        assert 0.50 < per_image < 0.56  # 0.543 TFLOP/image fwd+bwd at 32x32


def test_armed_optimizer_update_runs_bucket_by_bucket_inside_backward(dry):
    """FusedAdamW.arm(): one sd2_adamw_step per gradient bucket, issued right after the last backward op that touches the bucket,
    in completion order; step() then only counts.  Unarmed, step() is the single launch over the whole arena."""
    from diffusion_b200.optim import FusedAdamW
    from diffusion_b200.unet import UNet2DConditionModel
    u = UNet2DConditionModel(**TINY_UNET_CONFIG)
    eng = u.engine(2, 32, 32, 77)
    arena, lib = eng.arena, eng.ctx.lib
    opt = FusedAdamW(u.parameters(), lr=1e-3)
    assert opt.arm() is False  # gradients not bound to the arena yet
    for n, p in arena.params.items():
        p.grad = arena.grad_view(n)
    pts = eng._update_points()
    log = []
    eng.bwd = [(lambda i=i, op=op: (log.append(i), op())) for i, op in enumerate(eng.bwd)]
    assert opt.arm() is True
    hook = arena.armed_update
    arena.armed_update = lambda lo, hi: (log.append(('update', lo, hi)), hook(lo, hi))
    eng.run_forward()
    before = lib.calls.get('sd2_adamw_step', 0)
    eng.run_backward(allow_update=True)
    assert lib.calls['sd2_adamw_step'] - before == len(eng.buckets)
    ups = [(i, e) for i, e in enumerate(log) if isinstance(e, tuple)]
    assert [e[1:] for _, e in ups] == [(lo, hi) for lo, hi, _ in eng.buckets]
    for (pos, _), pt in zip(ups, pts):
        prev = [e for e in log[:pos] if not isinstance(e, tuple)]
        assert prev[-1] == pt - 1 and len(prev) == pt  # every backward op up to the bucket's update point ran before it
    assert arena.update_applied and arena.armed_update is None
    opt.step()
    assert lib.calls['sd2_adamw_step'] - before == len(eng.buckets) and opt.state['arena_group0']['step'] == 1
    assert not arena.update_applied
    # unarmed: the ordinary one-launch step; an armed step whose backward never ran falls back to it as well
    log.clear()
    eng.run_backward(allow_update=True)
    assert not any(isinstance(e, tuple) for e in log)
    opt.step()
    assert lib.calls['sd2_adamw_step'] - before == len(eng.buckets) + 1 and opt.state['arena_group0']['step'] == 2
    assert opt.arm() is True
    opt.step()
    assert lib.calls['sd2_adamw_step'] - before == len(eng.buckets) + 2 and arena.armed_update is None


def test_upsample_fold_tap_tables_reproduce_conv_of_the_upsampled_tensor():
    """The host-side tap tables of the folded Upsample2D (ops.taps_upconv / taps_upconv_dgrad / _upconv_groups) restated on the
    CPU: four 4-tap phase convolutions of the low-resolution input with group-summed weights equal conv3x3(nearest x2 (x)), and
    the dgrad taps are their adjoint.  (The kernels are checked against torch on the GPU in tests/test_kernels_gpu.py.)"""
    import torch.nn.functional as F
    from diffusion_b200 import ops
    torch.manual_seed(3)
    B, C, Co, H, W = 2, 5, 4, 6, 7
    x = torch.randn(B, C, H, W, dtype=torch.float64, requires_grad=True)
    w = torch.randn(Co, C, 3, 3, dtype=torch.float64)
    ref = F.conv2d(F.interpolate(x, scale_factor=2, mode='nearest'), w, padding=1)

    def shifted(t, dh, dw):  # y[h, w] = t[h + dh, w + dw], zero outside
        p = F.pad(t, (1, 1, 1, 1))
        return p[:, :, 1 + dh:1 + dh + t.shape[2], 1 + dw:1 + dw + t.shape[3]]

    def weff(ph, tap):
        a, b = tap >> 1, tap & 1
        kys, kxs = ops._upconv_groups(ph >> 1)[a][1], ops._upconv_groups(ph & 1)[b][1]
        return sum(w[:, :, ky, kx] for ky in kys for kx in kxs)

    out = torch.zeros_like(ref)
    for ph in range(4):
        taps = ops.taps_upconv(ph)
        assert len(taps) == 4 and sorted(t[3] for t in taps) == [0, 1, 2, 3] and all(t[2] == 0 for t in taps)
        plane = sum(torch.einsum('bchw,oc->bohw', shifted(x, dh, dw), weff(ph, t)) for dh, dw, _, t in taps)
        out[:, :, (ph >> 1)::2, (ph & 1)::2] = plane
    assert torch.allclose(out, ref, atol=1e-12)
    # dgrad: dx = sum over phases and taps of the phase gradient shifted by (-dh, -dw) times weff^T
    g = torch.randn_like(ref)
    ref.backward(g)
    dx = torch.zeros_like(x)
    for ph in range(4):
        gp = g[:, :, (ph >> 1)::2, (ph & 1)::2]
        fwd = {t: (dh, dw) for dh, dw, _, t in ops.taps_upconv(ph)}
        for dh, dw, _, t in ops.taps_upconv_dgrad(ph):
            assert (dh, dw) == (-fwd[t][0], -fwd[t][1])
            dx = dx + torch.einsum('bohw,oc->bchw', shifted(gp, dh, dw), weff(ph, t))
    assert torch.allclose(dx, x.grad, atol=1e-12)


def test_low_precision_norm_surgery_finds_nothing_to_replace():
    """reference train.py:91-108 runs composer's module surgery over `model.unet`, replacing every nn.GroupNorm / nn.LayerNorm
    instance by a new module with new parameters; the product's norm holders must not be such instances (their parameters
    live in the arena), while keeping the diffusers parameter names."""
    import torch
    from diffusion_b200.encoders import AutoencoderKL, CLIPTextModel
    from diffusion_b200.unet import NormParams, UNet2DConditionModel
    u = UNet2DConditionModel(**TINY_UNET_CONFIG)
    mods = [u, AutoencoderKL(block_out_channels=(64, 64, 64, 64), layers_per_block=1),
            CLIPTextModel(num_hidden_layers=1, vocab_size=64)]
    for m in mods:
        assert not any(isinstance(x, (torch.nn.GroupNorm, torch.nn.LayerNorm)) for x in m.modules())
    norms = [x for x in u.modules() if isinstance(x, NormParams)]
    assert len(norms) == 61 + 48  # 61 GroupNorms (resnets, transformers, conv_norm_out) + 3 LayerNorms in each of 16 blocks
    assert all(torch.equal(n.weight, torch.ones_like(n.weight)) and not n.bias.any() for n in norms)
    assert 'down_blocks.0.resnets.0.norm1.weight' in dict(u.named_parameters())


def test_documents_reference_existing_files():
    """Every profiles/ tools/ tests/ oracle/ diffusion_b200/ path named in the top-level documents exists."""
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    missing = []
    for doc in ('DESIGN.md', 'INTEGRATION.md', 'README.md', os.path.join('profiles', 'README.md')):
        text = open(os.path.join(root, doc)).read()
        for m in re.finditer(r'`((?:profiles|tools|tests|oracle|diffusion_b200|include)/[\w./{},\-*]+)`', text):
            path = m.group(1).split('::')[0].rstrip('.,')
            if '*' in path:
                continue
            variants = [path]
            b = re.search(r'\{([^}]*)\}', path)
            if b:  # brace lists: a{1,2}b -> a1b, a2b
                variants = [path[:b.start()] + alt + path[b.end():] for alt in b.group(1).split(',')]
            for v in variants:
                base = v if doc != os.path.join('profiles', 'README.md') or '/' in v else os.path.join('profiles', v)
                if not os.path.exists(os.path.join(root, base)):
                    missing.append((doc, v))
    assert not missing, missing

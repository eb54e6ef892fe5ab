"""B200-native `UNet2DConditionModel`: same constructor config, parameter names, shapes and call signature as the
diffusers class the reference builds at `diffusion/models/models.py:74-78` and calls at
`diffusion/models/stable_diffusion.py:183`, but the arithmetic is a static schedule of hand-written sm_100a kernels
(diffusion_b200/engine.py) behind the C ABI.

The module tree below only HOLDS parameters (torch-default init, diffusers names -> optimizers, checkpoints, EMA and
composer's low-precision-norm surgery keep working); none of the sub-modules' own forward() is ever called.
"""
from typing import Sequence

import torch
import torch.nn as nn

SD2_BASE_UNET_CONFIG = dict(
    in_channels=4,
    out_channels=4,
    block_out_channels=(320, 640, 1280, 1280),
    down_block_types=('CrossAttnDownBlock2D', 'CrossAttnDownBlock2D', 'CrossAttnDownBlock2D', 'DownBlock2D'),
    up_block_types=('UpBlock2D', 'CrossAttnUpBlock2D', 'CrossAttnUpBlock2D', 'CrossAttnUpBlock2D'),
    layers_per_block=2,
    attention_head_dim=(5, 10, 20, 20),
    cross_attention_dim=1024,
    norm_num_groups=32,
    norm_eps=1e-5,
    flip_sin_to_cos=True,
    freq_shift=0,
    use_linear_projection=True,
)


class UNetOutput(dict):
    """`{'sample': ...}` with attribute access: the reference reads both `unet(...)['sample']`
    (stable_diffusion.py:183) and `unet(...).sample` (:362)."""
    __getattr__ = dict.__getitem__


class NormParams(nn.Module):
    """weight / bias holder with nn.GroupNorm's / nn.LayerNorm's parameter names, shapes and initialisation (ones, zeros).
    Deliberately NOT a subclass of those classes: composer's low-precision-norm surgery (reference train.py:91-108,
    `apply_low_precision_groupnorm/layernorm`) replaces every instance of them by a new module with NEW parameters, which
    would orphan the arena-backed ones and silently stop the norm weights from training.  With this holder the surgery
    finds nothing to replace, and the kernels already implement its semantics (bf16 I/O, fp32 statistics)."""

    def __init__(self, channels, eps=1e-5, num_groups=None):
        super().__init__()
        self.num_channels, self.eps, self.num_groups = channels, eps, num_groups
        self.weight = nn.Parameter(torch.ones(channels))
        self.bias = nn.Parameter(torch.zeros(channels))

    def extra_repr(self):
        return f'{self.num_channels}, eps={self.eps}' + (f', groups={self.num_groups}' if self.num_groups else '')


def _holder(**children):
    m = nn.Module()
    for k, v in children.items():
        if v is not None:
            setattr(m, k, v)
    return m


def _resnet(cin, cout, temb, groups, eps):
    return _holder(norm1=NormParams(cin, eps, groups),
                   conv1=nn.Conv2d(cin, cout, 3, padding=1),
                   time_emb_proj=nn.Linear(temb, cout),
                   norm2=NormParams(cout, eps, groups),
                   conv2=nn.Conv2d(cout, cout, 3, padding=1),
                   conv_shortcut=nn.Conv2d(cin, cout, 1) if cin != cout else None)


def _attention(dim, cross_dim):
    return _holder(to_q=nn.Linear(dim, dim, bias=False),
                   to_k=nn.Linear(cross_dim or dim, dim, bias=False),
                   to_v=nn.Linear(cross_dim or dim, dim, bias=False),
                   to_out=nn.ModuleList([nn.Linear(dim, dim), nn.Dropout(0.0)]))


def _transformer(ch, cross_dim, groups):
    block = _holder(norm1=NormParams(ch),
                    attn1=_attention(ch, None),
                    norm2=NormParams(ch),
                    attn2=_attention(ch, cross_dim),
                    norm3=NormParams(ch),
                    ff=_holder(net=nn.ModuleList([_holder(proj=nn.Linear(ch, ch * 8)),
                                                  nn.Dropout(0.0),
                                                  nn.Linear(ch * 4, ch)])))
    return _holder(norm=NormParams(ch, 1e-6, groups),
                   proj_in=nn.Linear(ch, ch),
                   transformer_blocks=nn.ModuleList([block]),
                   proj_out=nn.Linear(ch, ch))


class UNet2DConditionModel(nn.Module):

    def __init__(self,
                 in_channels=4,
                 out_channels=4,
                 block_out_channels: Sequence[int] = (320, 640, 1280, 1280),
                 down_block_types: Sequence[str] = SD2_BASE_UNET_CONFIG['down_block_types'],
                 up_block_types: Sequence[str] = SD2_BASE_UNET_CONFIG['up_block_types'],
                 layers_per_block=2,
                 attention_head_dim=(5, 10, 20, 20),
                 cross_attention_dim=1024,
                 norm_num_groups=32,
                 norm_eps=1e-5,
                 flip_sin_to_cos=True,
                 freq_shift=0,
                 use_linear_projection=True,
                 **_ignored):
        super().__init__()
        if not (use_linear_projection and flip_sin_to_cos and freq_shift == 0 and in_channels == 4 and out_channels == 4):
            raise ValueError('only the SD-2 UNet family (linear projections, 4 latent channels) is implemented')
        boc = tuple(block_out_channels)
        heads = tuple(attention_head_dim) if not isinstance(attention_head_dim, int) else (attention_head_dim,) * len(boc)
        self.config = dict(in_channels=in_channels, out_channels=out_channels, block_out_channels=boc,
                           down_block_types=tuple(down_block_types), up_block_types=tuple(up_block_types),
                           layers_per_block=layers_per_block, attention_head_dim=heads,
                           cross_attention_dim=cross_attention_dim, norm_num_groups=norm_num_groups, norm_eps=norm_eps)
        g, eps, xd = norm_num_groups, norm_eps, cross_attention_dim
        temb = boc[0] * 4
        self.conv_in = nn.Conv2d(in_channels, boc[0], 3, padding=1)
        self.time_embedding = _holder(linear_1=nn.Linear(boc[0], temb), linear_2=nn.Linear(temb, temb))
        self.down_blocks = nn.ModuleList()
        out_ch = boc[0]
        for i, t in enumerate(down_block_types):
            in_ch, out_ch = out_ch, boc[i]
            attn = t == 'CrossAttnDownBlock2D'
            blk = _holder(
                attentions=nn.ModuleList([_transformer(out_ch, xd, g) for _ in range(layers_per_block)]) if attn else None,
                resnets=nn.ModuleList(
                    [_resnet(in_ch if j == 0 else out_ch, out_ch, temb, g, eps) for j in range(layers_per_block)]),
                downsamplers=nn.ModuleList([_holder(conv=nn.Conv2d(out_ch, out_ch, 3, stride=2, padding=1))])
                if i != len(boc) - 1 else None)
            self.down_blocks.append(blk)
        self.mid_block = _holder(attentions=nn.ModuleList([_transformer(boc[-1], xd, g)]),
                                 resnets=nn.ModuleList([_resnet(boc[-1], boc[-1], temb, g, eps) for _ in range(2)]))
        self.up_blocks = nn.ModuleList()
        rev = boc[::-1]
        out_ch = rev[0]
        for i, t in enumerate(up_block_types):
            prev, out_ch = out_ch, rev[i]
            in_ch = rev[min(i + 1, len(boc) - 1)]
            attn = t == 'CrossAttnUpBlock2D'
            n = layers_per_block + 1
            res = []
            for j in range(n):
                skip = in_ch if j == n - 1 else out_ch
                rin = prev if j == 0 else out_ch
                res.append(_resnet(rin + skip, out_ch, temb, g, eps))
            blk = _holder(attentions=nn.ModuleList([_transformer(out_ch, xd, g) for _ in range(n)]) if attn else None,
                          resnets=nn.ModuleList(res),
                          upsamplers=nn.ModuleList([_holder(conv=nn.Conv2d(out_ch, out_ch, 3, padding=1))])
                          if i != len(boc) - 1 else None)
            self.up_blocks.append(blk)
        self.conv_norm_out = NormParams(boc[0], eps, g)
        self.conv_out = nn.Conv2d(boc[0], out_channels, 3, padding=1)
        self._engines = {}
        self._arena = None
        # True under a torch DistributedDataParallel wrapper (or SD2_DDP_COMPAT=1): every backward hands autograd a gradient
        # for every parameter so that the wrapper's reduction hooks fire, also when the gradients are accumulated in place
        import os
        self.ddp_compat = os.environ.get('SD2_DDP_COMPAT') == '1'

    # ------------------------------------------------------------------------------------------------------------
    def bind_arena(self):
        """Move the parameters into the flat fp32 arena NOW (otherwise the first forward does it).  The factory calls this
        right after the model reaches the GPU: wrappers that record parameter strides at construction time (torch
        DistributedDataParallel) must see the final, arena-backed parameters - rebinding afterwards silently corrupts the
        gradients DDP hands back (reproduced with a plain permuted-view parameter on CPU / gloo)."""
        from diffusion_b200.engine import ParamArena
        dev = self.conv_in.weight.device
        arena = getattr(self, '_arena', None)
        if arena is None or not arena.bound():
            self._arena = ParamArena(self, dev)
        return self._arena

    def drop_engines(self):
        """Release every cached kernel schedule (activation buffers, CUDA graphs); the parameter arena stays."""
        self._engines.clear()

    def engine(self, B, H, W, ctx_len, forward_only=False):
        """Static kernel schedule for one input geometry (built lazily, cached)."""
        import os
        from diffusion_b200.engine import Engine
        dev = self.conv_in.weight.device
        from diffusion_b200.ops import dry_run
        if dev.type != 'cuda' and not dry_run():
            raise RuntimeError('diffusion_b200 runs on sm_100a GPUs only (no CPU fallback): move the model to CUDA')
        key = (B, H, W, ctx_len, dev.index, bool(forward_only))
        eng = self._engines.get(key)
        if eng is None or not eng.params_bound():
            prev = next(iter(self._engines.values()), None)
            eng = Engine(self, B, H, W, ctx_len, shared=prev, forward_only=forward_only)
            self._engines[key] = eng
            # data parallel by default: with torch.distributed initialised on more than one rank the engine averages its
            # gradient buckets itself (idempotent under an additional DDP wrapper; SD2_NO_AUTO_SYNC=1 or eng.sync_grads = False
            # turn it off)
            if not forward_only and not dry_run() and os.environ.get('SD2_NO_AUTO_SYNC') != '1':
                import torch.distributed as dist
                if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1 and not getattr(eng, 'sync_grads', False):
                    eng.enable_grad_sync()
        # Under a torch DistributedDataParallel wrapper (Composer's default; the wrapper marks itself active for the duration
        # of its forward) the wrapper owns the gradient reduction: its per-parameter hooks only fire when autograd receives a
        # gradient for every parameter (ddp_compat), and the engine's own bucket all-reduce is switched off so the arena is
        # reduced once, and DDP.no_sync() / Composer microbatching keep their meaning.
        wrapped = getattr(torch.nn.parallel.DistributedDataParallel, '_active_ddp_module', None) is not None
        if wrapped and not forward_only:
            eng.sync_grads = False
        eng.ddp_compat = self.ddp_compat or wrapped
        return eng

    def forward(self, sample, timestep, encoder_hidden_states, **_unused):
        """diffusers call signature; returns {'sample': eps_pred} (B,4,h,w) in sample.dtype.  Differentiable with
        respect to the parameters (custom autograd node that replays the static backward schedule)."""
        from diffusion_b200.engine import unet_apply
        return UNetOutput(sample=unet_apply(self, sample, timestep, encoder_hidden_states))

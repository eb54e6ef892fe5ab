"""ORACLE (test infrastructure, never the product path): torch-only restatement of the
third-party `diffusers.UNet2DConditionModel` that the reference constructs at
`diffusion/models/models.py:74-78` and calls at `diffusion/models/stable_diffusion.py:183`.

PARITY UNPINNED: diffusers is not vendored in /root/reference nor installable here, and the
reference's own tests assert shapes only (`tests/test_model.py:27-28,46`).  The anchors this
restatement is pinned against are (tests/test_oracle.py):
  * parameter inventory 865,910,724 parameters / 686 tensors for the SD-2-base config,
  * diffusers parameter names,
  * an independent sinusoidal-embedding restatement.
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline/reference arm may import this.

Behavioural spec followed: SURVEY.md Appendix B3-B5 (diffusers ~0.16-0.19 semantics).
"""
import math
from typing import List, Optional, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F

SD2_BASE_UNET_CONFIG = dict(
    in_channels=4,
    out_channels=4,
    block_out_channels=(320, 640, 1280, 1280),
    down_block_types=('CrossAttnDownBlock2D', 'CrossAttnDownBlock2D', 'CrossAttnDownBlock2D', 'DownBlock2D'),
    up_block_types=('UpBlock2D', 'CrossAttnUpBlock2D', 'CrossAttnUpBlock2D', 'CrossAttnUpBlock2D'),
    layers_per_block=2,
    attention_head_dim=(5, 10, 20, 20),  # really "number of heads" in this diffusers era
    cross_attention_dim=1024,
    norm_num_groups=32,
    norm_eps=1e-5,
    flip_sin_to_cos=True,
    freq_shift=0,
    use_linear_projection=True,
)

TINY_UNET_CONFIG = dict(SD2_BASE_UNET_CONFIG,
                        block_out_channels=(64, 128, 256, 256),
                        attention_head_dim=(1, 2, 4, 4))


def get_timestep_embedding(timesteps: torch.Tensor, embedding_dim: int, flip_sin_to_cos: bool,
                           downscale_freq_shift: float, max_period: int = 10000) -> torch.Tensor:
    """diffusers `models/embeddings.py::get_timestep_embedding`."""
    half = embedding_dim // 2
    exponent = -math.log(max_period) * torch.arange(0, half, dtype=torch.float32, device=timesteps.device)
    exponent = exponent / (half - downscale_freq_shift)
    emb = torch.exp(exponent)
    emb = timesteps[:, None].float() * emb[None, :]
    emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)
    if flip_sin_to_cos:
        emb = torch.cat([emb[:, half:], emb[:, :half]], dim=-1)
    return emb


class TimestepEmbedding(nn.Module):

    def __init__(self, in_channels, time_embed_dim):
        super().__init__()
        self.linear_1 = nn.Linear(in_channels, time_embed_dim)
        self.linear_2 = nn.Linear(time_embed_dim, time_embed_dim)

    def forward(self, x):
        return self.linear_2(F.silu(self.linear_1(x)))


class ResnetBlock2D(nn.Module):
    """diffusers `models/resnet.py::ResnetBlock2D`."""

    def __init__(self, cin, cout, temb_channels, groups, eps):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=eps, affine=True)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb_channels, cout)
        self.norm2 = nn.GroupNorm(groups, cout, eps=eps, affine=True)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x, temb):
        h = self.conv1(F.silu(self.norm1(x)))
        h = h + self.time_emb_proj(F.silu(temb))[:, :, None, None]
        h = self.conv2(F.silu(self.norm2(h)))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return x + h


class Attention(nn.Module):
    """diffusers `models/attention_processor.py::Attention`."""

    def __init__(self, query_dim, cross_dim, heads, dim_head):
        super().__init__()
        inner = heads * dim_head
        self.heads = heads
        self.to_q = nn.Linear(query_dim, inner, bias=False)
        self.to_k = nn.Linear(cross_dim or query_dim, inner, bias=False)
        self.to_v = nn.Linear(cross_dim or query_dim, inner, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(inner, query_dim), nn.Dropout(0.0)])

    def forward(self, x, ctx=None):
        src = x if ctx is None else ctx
        B, N, _ = x.shape
        q, k, v = self.to_q(x), self.to_k(src), self.to_v(src)

        def split(t):
            return t.view(B, t.shape[1], self.heads, -1).transpose(1, 2)

        o = F.scaled_dot_product_attention(split(q), split(k), split(v))
        o = o.transpose(1, 2).reshape(B, N, -1)
        return self.to_out[0](o)


class GEGLU(nn.Module):

    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)

    def forward(self, x):
        a, g = self.proj(x).chunk(2, dim=-1)
        return a * F.gelu(g)


class FeedForward(nn.Module):

    def __init__(self, dim):
        super().__init__()
        self.net = nn.ModuleList([GEGLU(dim, dim * 4), nn.Dropout(0.0), nn.Linear(dim * 4, dim)])

    def forward(self, x):
        for m in self.net:
            x = m(x)
        return x


class BasicTransformerBlock(nn.Module):

    def __init__(self, dim, heads, dim_head, cross_dim):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn1 = Attention(dim, None, heads, dim_head)
        self.norm2 = nn.LayerNorm(dim)
        self.attn2 = Attention(dim, cross_dim, heads, dim_head)
        self.norm3 = nn.LayerNorm(dim)
        self.ff = FeedForward(dim)

    def forward(self, h, ctx):
        h = self.attn1(self.norm1(h)) + h
        h = self.attn2(self.norm2(h), ctx) + h
        h = self.ff(self.norm3(h)) + h
        return h


class Transformer2DModel(nn.Module):
    """diffusers `models/transformer_2d.py::Transformer2DModel`, use_linear_projection=True."""

    def __init__(self, channels, heads, cross_dim, groups):
        super().__init__()
        dim_head = channels // heads
        self.norm = nn.GroupNorm(groups, channels, eps=1e-6, affine=True)
        self.proj_in = nn.Linear(channels, channels)
        self.transformer_blocks = nn.ModuleList([BasicTransformerBlock(channels, heads, dim_head, cross_dim)])
        self.proj_out = nn.Linear(channels, channels)

    def forward(self, x, ctx):
        B, C, H, W = x.shape
        r = x
        h = self.norm(x)
        h = h.permute(0, 2, 3, 1).reshape(B, H * W, C)
        h = self.proj_in(h)
        for blk in self.transformer_blocks:
            h = blk(h, ctx)
        h = self.proj_out(h)
        h = h.reshape(B, H, W, C).permute(0, 3, 1, 2).contiguous()
        return h + r


class Downsample2D(nn.Module):

    def __init__(self, channels):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels, 3, stride=2, padding=1)

    def forward(self, x):
        return self.conv(x)


class Upsample2D(nn.Module):

    def __init__(self, channels):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode='nearest'))


class DownBlock(nn.Module):

    def __init__(self, cin, cout, temb, layers, groups, eps, heads, cross_dim, has_attn, add_down):
        super().__init__()
        self.resnets = nn.ModuleList(
            [ResnetBlock2D(cin if i == 0 else cout, cout, temb, groups, eps) for i in range(layers)])
        if has_attn:
            self.attentions = nn.ModuleList([Transformer2DModel(cout, heads, cross_dim, groups) for _ in range(layers)])
        else:
            self.attentions = None
        self.downsamplers = nn.ModuleList([Downsample2D(cout)]) if add_down else None

    def forward(self, x, temb, ctx):
        outs = []
        for i, res in enumerate(self.resnets):
            x = res(x, temb)
            if self.attentions is not None:
                x = self.attentions[i](x, ctx)
            outs.append(x)
        if self.downsamplers is not None:
            x = self.downsamplers[0](x)
            outs.append(x)
        return x, outs


class MidBlock(nn.Module):

    def __init__(self, ch, temb, groups, eps, heads, cross_dim):
        super().__init__()
        # diffusers registers `attentions` before `resnets` in UNetMidBlock2DCrossAttn
        self.attentions = nn.ModuleList([Transformer2DModel(ch, heads, cross_dim, groups)])
        self.resnets = nn.ModuleList([ResnetBlock2D(ch, ch, temb, groups, eps) for _ in range(2)])

    def forward(self, x, temb, ctx):
        x = self.resnets[0](x, temb)
        x = self.attentions[0](x, ctx)
        return self.resnets[1](x, temb)


class UpBlock(nn.Module):

    def __init__(self, cin, cout, prev, temb, layers, groups, eps, heads, cross_dim, has_attn, add_up):
        super().__init__()
        res = []
        for i in range(layers):
            skip = cin if i == layers - 1 else cout
            rin = prev if i == 0 else cout
            res.append(ResnetBlock2D(rin + skip, cout, temb, groups, eps))
        self.resnets = nn.ModuleList(res)
        if has_attn:
            self.attentions = nn.ModuleList([Transformer2DModel(cout, heads, cross_dim, groups) for _ in range(layers)])
        else:
            self.attentions = None
        self.upsamplers = nn.ModuleList([Upsample2D(cout)]) if add_up else None

    def forward(self, x, skips: List[torch.Tensor], temb, ctx):
        for i, res in enumerate(self.resnets):
            x = torch.cat([x, skips.pop()], dim=1)
            x = res(x, temb)
            if self.attentions is not None:
                x = self.attentions[i](x, ctx)
        if self.upsamplers is not None:
            x = self.upsamplers[0](x)
        return x


class UNet2DConditionModel(nn.Module):
    """Restated `diffusers.UNet2DConditionModel` with the diffusers parameter names."""

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
        assert use_linear_projection
        boc = tuple(block_out_channels)
        heads = tuple(attention_head_dim) if not isinstance(attention_head_dim, int) else (attention_head_dim,) * len(boc)
        self.flip_sin_to_cos, self.freq_shift = flip_sin_to_cos, freq_shift
        temb = boc[0] * 4
        self.conv_in = nn.Conv2d(in_channels, boc[0], 3, padding=1)
        self.time_embedding = TimestepEmbedding(boc[0], temb)
        self.down_blocks = nn.ModuleList()
        out_ch = boc[0]
        for i, t in enumerate(down_block_types):
            in_ch, out_ch = out_ch, boc[i]
            self.down_blocks.append(
                DownBlock(in_ch, out_ch, temb, layers_per_block, norm_num_groups, norm_eps, heads[i],
                          cross_attention_dim, t == 'CrossAttnDownBlock2D', i != len(boc) - 1))
        self.mid_block = MidBlock(boc[-1], temb, norm_num_groups, norm_eps, heads[-1], cross_attention_dim)
        self.up_blocks = nn.ModuleList()
        rev, rheads = boc[::-1], heads[::-1]
        out_ch = rev[0]
        for i, t in enumerate(up_block_types):
            prev, out_ch = out_ch, rev[i]
            in_ch = rev[min(i + 1, len(boc) - 1)]
            self.up_blocks.append(
                UpBlock(in_ch, out_ch, prev, temb, layers_per_block + 1, norm_num_groups, norm_eps, rheads[i],
                        cross_attention_dim, t == 'CrossAttnUpBlock2D', i != len(boc) - 1))
        self.conv_norm_out = nn.GroupNorm(norm_num_groups, boc[0], eps=norm_eps)
        self.conv_out = nn.Conv2d(boc[0], out_channels, 3, padding=1)
        self.block0 = boc[0]

    def forward(self, sample, timestep, encoder_hidden_states):
        if not torch.is_tensor(timestep):
            timestep = torch.tensor([timestep], dtype=torch.long, device=sample.device)
        timestep = timestep.reshape(-1).to(sample.device).expand(sample.shape[0])  # 0-dim scheduler timesteps live on the CPU
        t_emb = get_timestep_embedding(timestep, self.block0, self.flip_sin_to_cos, self.freq_shift)
        t_emb = t_emb.to(dtype=sample.dtype)
        emb = self.time_embedding(t_emb)
        x = self.conv_in(sample)
        skips = [x]
        for blk in self.down_blocks:
            x, outs = blk(x, emb, encoder_hidden_states)
            skips.extend(outs)
        x = self.mid_block(x, emb, encoder_hidden_states)
        for blk in self.up_blocks:
            x = blk(x, skips, emb, encoder_hidden_states)
        x = self.conv_out(F.silu(self.conv_norm_out(x)))
        return {'sample': x}


# ---- composer low-precision norm surgery restated (reference diffusion/train.py:91-108) ----------------


class LPGroupNorm(nn.GroupNorm):
    """composer `algorithms/low_precision_groupnorm`: GN on autocast-dtype operands, autocast disabled."""

    def forward(self, x):
        if not torch.is_autocast_enabled(x.device.type):
            return super().forward(x)
        dt = torch.get_autocast_dtype(x.device.type)
        with torch.autocast(enabled=False, device_type=x.device.type):
            return F.group_norm(x.to(dt), self.num_groups, self.weight.to(dt), self.bias.to(dt), self.eps)


class LPLayerNorm(nn.LayerNorm):

    def forward(self, x):
        if not torch.is_autocast_enabled(x.device.type):
            return super().forward(x)
        dt = torch.get_autocast_dtype(x.device.type)
        with torch.autocast(enabled=False, device_type=x.device.type):
            return F.layer_norm(x.to(dt), self.normalized_shape, self.weight.to(dt), self.bias.to(dt), self.eps)


def apply_low_precision_norms(module: nn.Module) -> nn.Module:
    """In-place class swap, keeps parameters (same effect as composer's module surgery)."""
    for m in module.modules():
        if type(m) is nn.GroupNorm:
            m.__class__ = LPGroupNorm
        elif type(m) is nn.LayerNorm:
            m.__class__ = LPLayerNorm
    return module

"""ORACLE (test infrastructure): torch restatement of diffusers `AutoencoderKL` as the reference loads it for SD-2
(`diffusion/models/models.py:80-85`, subfolder 'vae') and calls it: `vae.encode(x)['latent_dist'].sample()`
(`diffusion/models/stable_diffusion.py:167,170`) and `vae.decode(latents).sample` (:376).  diffusers modules restated:
`models/vae.py::{Encoder, Decoder, DiagonalGaussianDistribution}`, `models/unet_2d_blocks.py::{DownEncoderBlock2D,
UpDecoderBlock2D, UNetMidBlock2D}`, `models/resnet.py::{ResnetBlock2D (temb=None), Downsample2D (padding 0 ->
F.pad (0,1,0,1)), Upsample2D (nearest 2x + conv)}`, `models/attention_processor.py::Attention` (1 head, group norm,
residual connection).  SD-2 VAE config: block_out_channels (128, 256, 512, 512), layers_per_block 2, latent_channels 4,
norm_num_groups 32, eps 1e-6, act silu, scaling_factor 0.18215.  Parameter names equal diffusers' (and therefore
diffusion_b200.encoders.AutoencoderKL's), 83,653,863 parameters for the SD-2 config.
PARITY UNPINNED (diffusers absent): anchored on the parameter count / names and torch's own op definitions.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


class ResnetBlock(nn.Module):

    def __init__(self, cin, cout, groups=32):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=1e-6)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.norm2 = nn.GroupNorm(groups, cout, eps=1e-6)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        if cin != cout:
            self.conv_shortcut = nn.Conv2d(cin, cout, 1)

    def forward(self, x):
        h = self.conv1(F.silu(self.norm1(x)))
        h = self.conv2(F.silu(self.norm2(h)))
        if hasattr(self, 'conv_shortcut'):
            x = self.conv_shortcut(x)
        return x + h


class VaeAttention(nn.Module):

    def __init__(self, ch, groups=32):
        super().__init__()
        self.group_norm = nn.GroupNorm(groups, ch, eps=1e-6)
        self.to_q, self.to_k, self.to_v = nn.Linear(ch, ch), nn.Linear(ch, ch), nn.Linear(ch, ch)
        self.to_out = nn.ModuleList([nn.Linear(ch, ch), nn.Dropout(0.0)])

    def forward(self, x):
        B, C, H, W = x.shape
        h = self.group_norm(x).view(B, C, H * W).transpose(1, 2)
        q, k, v = self.to_q(h), self.to_k(h), self.to_v(h)
        p = torch.softmax(torch.baddbmm(torch.zeros((), dtype=q.dtype, device=q.device), q, k.transpose(1, 2), beta=0,
                                        alpha=C**-0.5).float(), dim=-1).to(q.dtype)
        o = self.to_out[0](torch.bmm(p, v))
        return x + o.transpose(1, 2).reshape(B, C, H, W)


class MidBlock(nn.Module):

    def __init__(self, ch, groups=32):
        super().__init__()
        self.attentions = nn.ModuleList([VaeAttention(ch, groups)])
        self.resnets = nn.ModuleList([ResnetBlock(ch, ch, groups), ResnetBlock(ch, ch, groups)])

    def forward(self, x):
        return self.resnets[1](self.attentions[0](self.resnets[0](x)))


class _Conv(nn.Module):

    def __init__(self, conv):
        super().__init__()
        self.conv = conv


class Encoder(nn.Module):

    def __init__(self, in_channels, boc, layers, latent, groups):
        super().__init__()
        self.conv_in = nn.Conv2d(in_channels, boc[0], 3, padding=1)
        self.down_blocks = nn.ModuleList()
        ch = boc[0]
        for i, out_ch in enumerate(boc):
            blk = nn.Module()
            blk.resnets = nn.ModuleList([ResnetBlock(ch if j == 0 else out_ch, out_ch, groups) for j in range(layers)])
            if i != len(boc) - 1:
                blk.downsamplers = nn.ModuleList([_Conv(nn.Conv2d(out_ch, out_ch, 3, stride=2, padding=0))])
            self.down_blocks.append(blk)
            ch = out_ch
        self.mid_block = MidBlock(boc[-1], groups)
        self.conv_norm_out = nn.GroupNorm(groups, boc[-1], eps=1e-6)
        self.conv_out = nn.Conv2d(boc[-1], 2 * latent, 3, padding=1)

    def forward(self, x):
        x = self.conv_in(x)
        for blk in self.down_blocks:
            for r in blk.resnets:
                x = r(x)
            if hasattr(blk, 'downsamplers'):
                x = blk.downsamplers[0].conv(F.pad(x, (0, 1, 0, 1), mode='constant', value=0))
        x = self.mid_block(x)
        return self.conv_out(F.silu(self.conv_norm_out(x)))


class Decoder(nn.Module):

    def __init__(self, out_channels, boc, layers, latent, groups):
        super().__init__()
        rev = boc[::-1]
        self.conv_in = nn.Conv2d(latent, rev[0], 3, padding=1)
        self.mid_block = MidBlock(rev[0], groups)
        self.up_blocks = nn.ModuleList()
        ch = rev[0]
        for i, out_ch in enumerate(rev):
            blk = nn.Module()
            blk.resnets = nn.ModuleList([ResnetBlock(ch if j == 0 else out_ch, out_ch, groups) for j in range(layers + 1)])
            if i != len(rev) - 1:
                blk.upsamplers = nn.ModuleList([_Conv(nn.Conv2d(out_ch, out_ch, 3, padding=1))])
            self.up_blocks.append(blk)
            ch = out_ch
        self.conv_norm_out = nn.GroupNorm(groups, boc[0], eps=1e-6)
        self.conv_out = nn.Conv2d(boc[0], out_channels, 3, padding=1)

    def forward(self, z):
        x = self.mid_block(self.conv_in(z))
        for blk in self.up_blocks:
            for r in blk.resnets:
                x = r(x)
            if hasattr(blk, 'upsamplers'):
                x = blk.upsamplers[0].conv(F.interpolate(x, scale_factor=2.0, mode='nearest'))
        return self.conv_out(F.silu(self.conv_norm_out(x)))


class AutoencoderKLOracle(nn.Module):

    def __init__(self, in_channels=3, out_channels=3, block_out_channels=(128, 256, 512, 512), layers_per_block=2,
                 latent_channels=4, norm_num_groups=32, scaling_factor=0.18215):
        super().__init__()
        boc = tuple(block_out_channels)
        self.encoder = Encoder(in_channels, boc, layers_per_block, latent_channels, norm_num_groups)
        self.decoder = Decoder(out_channels, boc, layers_per_block, latent_channels, norm_num_groups)
        self.quant_conv = nn.Conv2d(2 * latent_channels, 2 * latent_channels, 1)
        self.post_quant_conv = nn.Conv2d(latent_channels, latent_channels, 1)

    def moments(self, x):
        return self.quant_conv(self.encoder(x))

    def encode_sample(self, x, generator=None):
        """`vae.encode(x)['latent_dist'].sample()`: DiagonalGaussianDistribution with logvar clamped to [-30, 20]."""
        mean, logvar = torch.chunk(self.moments(x), 2, dim=1)
        std = torch.exp(0.5 * torch.clamp(logvar, -30.0, 20.0))
        noise = torch.randn(mean.shape, generator=generator, device=mean.device, dtype=mean.dtype)
        return mean + std * noise

    def decode(self, z):
        return self.decoder(self.post_quant_conv(z))


TINY_VAE_CONFIG = dict(in_channels=3, out_channels=3, block_out_channels=(64, 128, 128, 128), layers_per_block=1,
                       latent_channels=4, norm_num_groups=32)

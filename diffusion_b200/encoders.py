"""In-loop latent encoders on the sm_100a kernels: the frozen VAE and CLIP text encoder the reference
runs when `precomputed_latents` is false (diffusion/models/stable_diffusion.py:160-174, built at models.py:80-85), plus the
VAE decoder `generate()` ends with (:375-379).

  * `AutoencoderKL`  - parameter skeleton with diffusers' names/shapes for the SD-2 VAE (encoder, quant_conv,
                       post_quant_conv, decoder); `encode(x)['latent_dist'].sample()` and `decode(z).sample` like diffusers
  * `CLIPTextModel`  - parameter skeleton with transformers' names/shapes for the SD-2 (OpenCLIP ViT-H) text tower;
                       `model(input_ids)[0]` is the final-layer-norm hidden state like transformers
  * `FrozenEngine`   - forward-only static schedule over such a module: flat fp32 arena + bf16 shadow (no gradient
                       arena), launch list over preallocated activations, replayed as ONE CUDA graph per geometry

Every contraction (3x3 / 1x1 convolutions, linears, Q K^T, P V) runs on the tcgen05 GEMM kernel, the norms on the norm
kernels, the rest on the glue kernels of csrc/encoders.cu.  Compute dtype: bf16 operands, fp32 accumulation and
statistics (the reference runs these two networks in fp16; the latents / conditioning they hand to the UNet are fp16 /
bf16 tensors either way - tolerances in tests/test_encoders_gpu.py).  No torch op runs inside either network; the
Gaussian noise of `latent_dist.sample()` is drawn by `torch.randn` from the default CUDA generator exactly where the
reference draws it, so the RNG stream (VAE noise, then timesteps, then training noise) stays bit-identical.
"""
from functools import partial

import torch
import torch.nn as nn

from diffusion_b200 import ops
from diffusion_b200.engine import ParamArena
from diffusion_b200.unet import NormParams

BF16 = torch.bfloat16

SD2_VAE_CONFIG = dict(in_channels=3, out_channels=3, block_out_channels=(128, 256, 512, 512), layers_per_block=2,
                      latent_channels=4, norm_num_groups=32, scaling_factor=0.18215)
SD2_TEXT_CONFIG = dict(vocab_size=49408, hidden_size=1024, intermediate_size=4096, num_hidden_layers=23,
                       num_attention_heads=16, max_position_embeddings=77, layer_norm_eps=1e-5, hidden_act='gelu')


class _Out(dict):
    """dict with attribute access (diffusers BaseOutput / transformers ModelOutput style); integer index 0 = first value."""
    __getattr__ = dict.__getitem__

    def __getitem__(self, k):
        if isinstance(k, int):
            return list(self.values())[k]
        return dict.__getitem__(self, k)


def _holder(**children):
    m = nn.Module()
    for k, v in children.items():
        if v is not None:
            setattr(m, k, v)
    return m


def _resnet(cin, cout, groups):
    return _holder(norm1=NormParams(cin, 1e-6, groups), conv1=nn.Conv2d(cin, cout, 3, padding=1),
                   norm2=NormParams(cout, 1e-6, groups), conv2=nn.Conv2d(cout, cout, 3, padding=1),
                   conv_shortcut=nn.Conv2d(cin, cout, 1) if cin != cout else None)


def _vae_attention(ch, groups):
    return _holder(group_norm=NormParams(ch, 1e-6, groups), to_q=nn.Linear(ch, ch), to_k=nn.Linear(ch, ch),
                   to_v=nn.Linear(ch, ch), to_out=nn.ModuleList([nn.Linear(ch, ch), nn.Dropout(0.0)]))


def _mid(ch, groups):
    return _holder(attentions=nn.ModuleList([_vae_attention(ch, groups)]),
                   resnets=nn.ModuleList([_resnet(ch, ch, groups), _resnet(ch, ch, groups)]))


# ==================================================================================================== engine
class FrozenEngine:
    """Forward-only launch list over a frozen module's parameters (see module docstring)."""
    _scratch = {}

    def __init__(self, module, dev):
        self.ctx, self.dev = ops.get_ctx(dev), dev
        if getattr(module, '_arena', None) is None or not module._arena.bound():
            module._arena = ParamArena(module, dev, with_grads=False)
        self.arena = module._arena
        self.fwd, self.act_bytes, self.graph = [], 0, None
        key = (dev.index, 'ws')
        if key not in FrozenEngine._scratch:
            FrozenEngine._scratch[key] = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        self.ws = FrozenEngine._scratch[key]
        self._gn_ws = None

    # ---- recording helpers
    def buf(self, *shape, dtype=BF16, zero=False):
        t = (torch.zeros if zero else torch.empty)(*shape, dtype=dtype, device=self.dev)
        self.act_bytes += t.numel() * t.element_size()
        return t

    def f(self, fn, *a, **k):
        self.fwd.append(partial(fn, self.ctx, *a, **k))

    def w16(self, name):
        return self.arena.storage(self.arena.p16, name)

    def p32(self, name):
        return self.arena.storage(self.arena.p32, name)

    def has(self, name):
        return name in self.arena.entries

    def lin(self, x, prefix, residual=None, bias=True):
        w = self.w16(prefix + '.weight')
        out = self.buf(x.shape[0], w.shape[0])
        self.f(ops.linear_fwd, x, w, out, bias=self.p32(prefix + '.bias') if bias else None, residual=residual, workspace=self.ws)
        return out

    def conv(self, x, B, H, W, prefix=None, residual=None, w=None, bias=None, taps=None, n_planes=None, out_cols=None):
        w = self.w16(prefix + '.weight') if w is None else w
        bias = self.p32(prefix + '.bias') if bias is None else bias
        out = self.buf(B * H * W, out_cols or w.shape[1])
        self.f(ops.conv3x3_fwd, x, B, H, W, w, out, bias=bias, residual=residual, taps=taps, n_planes=n_planes, workspace=self.ws)
        return out

    def gn(self, x, prefix, B, HW, G, eps, silu):
        need = self.ctx.lib.sd2_groupnorm_ws_floats(B, x.shape[1])
        if self._gn_ws is None or self._gn_ws.numel() < need:
            self._gn_ws = torch.empty(need, dtype=torch.float32, device=self.dev)
        y = self.buf(*x.shape)
        stats = self.buf(B, G, 2, dtype=torch.float32)
        self.f(ops.groupnorm_fwd, x, self.p32(prefix + '.weight'), self.p32(prefix + '.bias'), y, stats, self._gn_ws, B, HW, G,
               eps, int(silu))
        return y

    def ln(self, x, prefix, eps=1e-5):
        y = self.buf(*x.shape)
        stats = self.buf(x.shape[0], 2, dtype=torch.float32)
        self.f(ops.layernorm_fwd, x, self.p32(prefix + '.weight'), self.p32(prefix + '.bias'), y, stats, eps)
        return y

    def attention(self, q, k, v, B, heads, N, d, causal):
        """softmax(q k^T / sqrt(d)) v per (batch, head) with materialised scores (these two networks have a single
        512-wide head / 77-token causal heads, which the UNet's fused head_dim-64 kernel does not cover)."""
        ld, C = q.stride(0), heads * d
        ldp = (N + 7) // 8 * 8
        nbh = B * heads
        S = self.buf(nbh, N, ldp, dtype=torch.float32)
        P = self.buf(nbh, N, ldp)
        out = self.buf(B * N, C)
        qd = (d, N, ld, d, N * ld)
        sd = (ldp, N * ldp, heads * N * ldp)
        pd = (N, N, ldp, N * ldp, heads * N * ldp)
        od = (C, d, N * C)
        self.f(ops.bmm, q, 0, qd, k, 0, qd, S, sd, N, N, d, nbh, heads, alpha=float(d)**-0.5, out_f32=True)
        if causal:
            self.f(ops.softmax_causal_fwd, S, P, nbh * N, N, N)
        else:
            self.f(ops.softmax_fwd, S, P, nbh * N, N)
        self.f(ops.bmm, P, 0, pd, v, 1, qd, out, od, N, d, N, nbh, heads)
        return out

    # ---- execution
    def run(self):
        self.arena.refresh_shadow(self.ctx)
        if self.graph is not None:
            self.graph.replay()
        else:
            for op in self.fwd:
                op()

    def capture(self):
        """One CUDA graph over the whole launch list (call after one eager run)."""
        torch.cuda.synchronize(self.dev)
        s = torch.cuda.Stream(self.dev)
        s.wait_stream(torch.cuda.current_stream(self.dev))
        g = torch.cuda.CUDAGraph()
        with torch.cuda.stream(s):
            with torch.cuda.graph(g, stream=s):
                for op in self.fwd:
                    op()
        torch.cuda.current_stream(self.dev).wait_stream(s)
        torch.cuda.synchronize(self.dev)
        self.graph = g


def _device_of(module):
    dev = next(module.parameters()).device
    if dev.type != 'cuda' and not ops.dry_run():
        raise RuntimeError('diffusion_b200 runs on sm_100a GPUs only (no CPU fallback): move the model to CUDA')
    return dev


# ==================================================================================================== VAE
class DiagonalGaussianDistribution:
    """`latent_dist` of `AutoencoderKL.encode`: sample() draws torch.randn from the default generator (shape / dtype /
    device of the mean, like diffusers' randn_tensor: ONE draw for the whole batch) and applies it in the fused
    quant_conv + sampling kernel."""

    def __init__(self, eng, moments8, shape, dtype):
        self._eng, self._moments8, self._shape, self._dtype = eng, moments8, shape, dtype

    def _apply(self, noise, scale, want_mean):
        eng = self._eng
        B, _, h, w = self._shape
        z = torch.empty_like(noise)
        mean = torch.empty_like(noise) if want_mean else None
        ops.vae_sample(eng.ctx, self._moments8, eng.p32('quant_conv.weight'), eng.p32('quant_conv.bias'), noise, z, mean, B, h, w, scale)
        return mean if want_mean else z

    def sample(self, generator=None, scale=1.0):
        noise = torch.randn(self._shape, generator=generator, device=self._eng.dev, dtype=self._dtype)
        return self._apply(noise, scale, False)

    def mode(self):
        return self._apply(torch.zeros(self._shape, device=self._eng.dev, dtype=self._dtype), 1.0, True)


class AutoencoderKL(nn.Module):
    _sd2_native = True  # computes in bf16 from its own shadow weights: StableDiffusion.__init__ must not .half() it

    def __init__(self, in_channels=3, out_channels=3, block_out_channels=(128, 256, 512, 512), layers_per_block=2,
                 latent_channels=4, norm_num_groups=32, scaling_factor=0.18215, **_ignored):
        super().__init__()
        boc, g = tuple(block_out_channels), norm_num_groups
        if in_channels > 8 or out_channels > 8 or 2 * latent_channels != 8 or any(c % 64 for c in boc):
            raise ValueError('only the SD VAE family (<= 8 image channels, 4 latent channels, widths % 64) is implemented')
        self.config = dict(in_channels=in_channels, out_channels=out_channels, block_out_channels=boc,
                           layers_per_block=layers_per_block, latent_channels=latent_channels, norm_num_groups=g,
                           scaling_factor=scaling_factor)
        enc = _holder(conv_in=nn.Conv2d(in_channels, boc[0], 3, padding=1), down_blocks=nn.ModuleList())
        ch = boc[0]
        for i, out_ch in enumerate(boc):
            enc.down_blocks.append(_holder(
                resnets=nn.ModuleList([_resnet(ch if j == 0 else out_ch, out_ch, g) for j in range(layers_per_block)]),
                downsamplers=nn.ModuleList([_holder(conv=nn.Conv2d(out_ch, out_ch, 3, stride=2, padding=0))])
                if i != len(boc) - 1 else None))
            ch = out_ch
        enc.mid_block = _mid(boc[-1], g)
        enc.conv_norm_out = NormParams(boc[-1], 1e-6, g)
        enc.conv_out = nn.Conv2d(boc[-1], 2 * latent_channels, 3, padding=1)
        self.encoder = enc
        dec = _holder(conv_in=nn.Conv2d(latent_channels, boc[-1], 3, padding=1), mid_block=_mid(boc[-1], g),
                      up_blocks=nn.ModuleList())
        rev = boc[::-1]
        ch = rev[0]
        for i, out_ch in enumerate(rev):
            dec.up_blocks.append(_holder(
                resnets=nn.ModuleList([_resnet(ch if j == 0 else out_ch, out_ch, g) for j in range(layers_per_block + 1)]),
                upsamplers=nn.ModuleList([_holder(conv=nn.Conv2d(out_ch, out_ch, 3, padding=1))]) if i != len(boc) - 1 else None))
            ch = out_ch
        dec.conv_norm_out = NormParams(boc[0], 1e-6, g)
        dec.conv_out = nn.Conv2d(boc[0], out_channels, 3, padding=1)
        self.decoder = dec
        self.quant_conv = nn.Conv2d(2 * latent_channels, 2 * latent_channels, 1)
        self.post_quant_conv = nn.Conv2d(latent_channels, latent_channels, 1)
        self.requires_grad_(False)
        self._arena, self._engines = None, {}
        self.max_chunk = 16  # images per encoder / decoder pass (the reference's microbatch)

    @property
    def device(self):
        return next(self.parameters()).device

    # ---- shared blocks
    @staticmethod
    def _resnet(e, x, prefix, B, H, W, G):
        a1 = e.gn(x, prefix + '.norm1', B, H * W, G, 1e-6, True)
        h1 = e.conv(a1, B, H, W, prefix + '.conv1')
        a2 = e.gn(h1, prefix + '.norm2', B, H * W, G, 1e-6, True)
        sc = e.lin(x, prefix + '.conv_shortcut') if e.has(prefix + '.conv_shortcut.weight') else x
        return e.conv(a2, B, H, W, prefix + '.conv2', residual=sc)

    @staticmethod
    def _mid(e, x, prefix, B, H, W, G):
        x = AutoencoderKL._resnet(e, x, prefix + '.resnets.0', B, H, W, G)
        a = prefix + '.attentions.0'
        n = e.gn(x, a + '.group_norm', B, H * W, G, 1e-6, False)
        q, k, v = e.lin(n, a + '.to_q'), e.lin(n, a + '.to_k'), e.lin(n, a + '.to_v')
        o = e.attention(q, k, v, B, 1, H * W, x.shape[1], causal=False)
        x = e.lin(o, a + '.to_out.0', residual=x)
        return AutoencoderKL._resnet(e, x, prefix + '.resnets.1', B, H, W, G)

    # ---- encoder
    def _encoder_engine(self, B, H, W):
        key = ('enc', B, H, W)
        e = self._engines.get(key)
        if e is not None and e.arena.bound():
            return e
        cfg = self.config
        boc, G = cfg['block_out_channels'], cfg['norm_num_groups']
        if H % (2**(len(boc) - 1)) or W % (2**(len(boc) - 1)):
            raise ValueError(f'image size {(H, W)} must be divisible by {2**(len(boc) - 1)}')
        e = FrozenEngine(self, _device_of(self))
        e.in_x8 = e.buf(B * H * W, 8)
        w_in = e.buf(9, boc[0], 8)
        e.f(ops.pad_cast_rows, e.p32('encoder.conv_in.weight').reshape(-1), cfg['in_channels'], w_in, 8, 9 * boc[0])
        x = e.conv(e.in_x8, B, H, W, w=w_in, bias=e.p32('encoder.conv_in.bias'))
        Hc, Wc = H, W
        for i in range(len(boc)):
            for j in range(cfg['layers_per_block']):
                x = self._resnet(e, x, f'encoder.down_blocks.{i}.resnets.{j}', B, Hc, Wc, G)
            if i != len(boc) - 1:
                C, Ho, Wo = x.shape[1], Hc // 2, Wc // 2
                planes = e.buf(4 * B * Ho * Wo, C)
                e.f(ops.phase_split, x, planes, B, Hc, Wc)
                x = e.conv(planes, B, Ho, Wo, f'encoder.down_blocks.{i}.downsamplers.0.conv', taps=ops.taps_stride2_vae(B),
                           n_planes=4 * B)
                Hc, Wc = Ho, Wo
        x = self._mid(e, x, 'encoder.mid_block', B, Hc, Wc, G)
        n = e.gn(x, 'encoder.conv_norm_out', B, Hc * Wc, G, 1e-6, True)
        e.moments8 = e.conv(n, B, Hc, Wc, 'encoder.conv_out')
        e.out_hw = (Hc, Wc)
        self._engines[key] = e
        return e

    def encode(self, x, return_dict=True):
        """x: images (B, C, H, W), any float dtype.  Returns {'latent_dist': DiagonalGaussianDistribution}.  Batches larger
        than `max_chunk` images go through the encoder in chunks (the activations of a 512^2 image are ~1.3 GB); only the
        8-channel moments of every chunk are kept."""
        B, Cc, H, W = x.shape
        if Cc != self.config['in_channels']:
            raise ValueError(f'expected {self.config["in_channels"]} image channels, got {Cc}')
        if x.device.type != 'cuda':
            raise RuntimeError('diffusion_b200 needs CUDA (sm_100a) tensors: there is no CPU fallback')
        x = x.contiguous()
        moments, e, b0 = None, None, 0
        while b0 < B:
            n = min(self.max_chunk, B - b0)
            e = self._encoder_engine(n, H, W)
            ops.nchw_to_nhwc8(e.ctx, x[b0:b0 + n], e.in_x8, n, Cc, H, W)
            e.run()
            if n == B:
                moments = e.moments8
                break
            h, w = e.out_hw
            if moments is None:
                moments = torch.empty(B * h * w, 8, dtype=BF16, device=x.device)
            ops.copy2d(e.ctx, e.moments8, moments[b0 * h * w:(b0 + n) * h * w], n * h * w, 8)
            b0 += n
        h, w = e.out_hw
        return _Out(latent_dist=DiagonalGaussianDistribution(e, moments, (B, self.config['latent_channels'], h, w), x.dtype))

    # ---- decoder
    def _decoder_engine(self, B, h, w):
        key = ('dec', B, h, w)
        e = self._engines.get(key)
        if e is not None and e.arena.bound():
            return e
        cfg = self.config
        boc, G, L = cfg['block_out_channels'], cfg['norm_num_groups'], cfg['latent_channels']
        rev = boc[::-1]
        e = FrozenEngine(self, _device_of(self))
        e.in_z8 = e.buf(B * h * w, 8)
        # post_quant_conv (1x1, 4 -> 4) as a per-pixel map, then conv_in with its 4 input channels padded to 8
        z8 = e.buf(B * h * w, 8)
        e.f(ops.pixel_linear8, e.in_z8, e.p32('post_quant_conv.weight'), e.p32('post_quant_conv.bias'), z8)
        w_in = e.buf(9, rev[0], 8)
        e.f(ops.pad_cast_rows, e.p32('decoder.conv_in.weight').reshape(-1), L, w_in, 8, 9 * rev[0])
        x = e.conv(z8, B, h, w, w=w_in, bias=e.p32('decoder.conv_in.bias'))
        x = self._mid(e, x, 'decoder.mid_block', B, h, w, G)
        Hc, Wc = h, w
        for i in range(len(rev)):
            for j in range(cfg['layers_per_block'] + 1):
                x = self._resnet(e, x, f'decoder.up_blocks.{i}.resnets.{j}', B, Hc, Wc, G)
            if i != len(rev) - 1:
                up = e.buf(4 * x.shape[0], x.shape[1])
                e.f(ops.upsample2x_fwd, x, up, B, Hc, Wc)
                Hc, Wc = 2 * Hc, 2 * Wc
                x = e.conv(up, B, Hc, Wc, f'decoder.up_blocks.{i}.upsamplers.0.conv')
        n = e.gn(x, 'decoder.conv_norm_out', B, Hc * Wc, G, 1e-6, True)
        b_out = e.buf(8, dtype=torch.float32, zero=True)
        e.f(ops.unpad_accum_rows, e.p32('decoder.conv_out.bias'), cfg['out_channels'], b_out, cfg['out_channels'], 1, False)
        e.image8 = e.conv(n, B, Hc, Wc, 'decoder.conv_out', bias=b_out, out_cols=8)
        e.out_hw = (Hc, Wc)
        self._engines[key] = e
        return e

    def decode(self, z, return_dict=True):
        """z: latents (B, 4, h, w).  Returns {'sample': images (B, 3, 8h, 8w)} in z.dtype."""
        B, L, h, w = z.shape
        if L != self.config['latent_channels']:
            raise ValueError(f'expected {self.config["latent_channels"]} latent channels, got {L}')
        if z.device.type != 'cuda':
            raise RuntimeError('diffusion_b200 needs CUDA (sm_100a) tensors: there is no CPU fallback')
        z = z.contiguous()
        img, b0 = None, 0
        while b0 < B:
            n = min(self.max_chunk, B - b0)
            e = self._decoder_engine(n, h, w)
            ops.nchw_to_nhwc8(e.ctx, z[b0:b0 + n], e.in_z8, n, L, h, w)
            e.run()
            if img is None:
                img = torch.empty(B, self.config['out_channels'], e.out_hw[0], e.out_hw[1], dtype=z.dtype, device=z.device)
            ops.nhwc8_to_nchw(e.ctx, e.image8, img[b0:b0 + n], n, self.config['out_channels'], e.out_hw[0], e.out_hw[1])
            b0 += n
        return _Out(sample=img)


# ==================================================================================================== CLIP text
class CLIPTextModel(nn.Module):
    _sd2_native = True

    def __init__(self, vocab_size=49408, hidden_size=1024, intermediate_size=4096, num_hidden_layers=23,
                 num_attention_heads=16, max_position_embeddings=77, layer_norm_eps=1e-5, hidden_act='gelu', **_ignored):
        super().__init__()
        if hidden_act != 'gelu' or hidden_size % num_attention_heads or hidden_size % 64 or hidden_size > 1280:
            raise ValueError('only the SD-2 text tower family (erf GELU, width % 64, width <= 1280) is implemented')
        self.config = dict(vocab_size=vocab_size, hidden_size=hidden_size, intermediate_size=intermediate_size,
                           num_hidden_layers=num_hidden_layers, num_attention_heads=num_attention_heads,
                           max_position_embeddings=max_position_embeddings, layer_norm_eps=layer_norm_eps)
        D = hidden_size

        def layer():
            return _holder(self_attn=_holder(k_proj=nn.Linear(D, D), v_proj=nn.Linear(D, D), q_proj=nn.Linear(D, D),
                                             out_proj=nn.Linear(D, D)),
                           layer_norm1=NormParams(D, layer_norm_eps),
                           mlp=_holder(fc1=nn.Linear(D, intermediate_size), fc2=nn.Linear(intermediate_size, D)),
                           layer_norm2=NormParams(D, layer_norm_eps))

        self.text_model = _holder(
            embeddings=_holder(token_embedding=nn.Embedding(vocab_size, D),
                               position_embedding=nn.Embedding(max_position_embeddings, D)),
            encoder=_holder(layers=nn.ModuleList([layer() for _ in range(num_hidden_layers)])),
            final_layer_norm=NormParams(D, layer_norm_eps))
        self.requires_grad_(False)
        self._arena, self._engines = None, {}

    @property
    def device(self):
        return next(self.parameters()).device

    def _engine(self, B, L):
        e = self._engines.get((B, L))
        if e is not None and e.arena.bound():
            return e
        cfg = self.config
        D, H, eps = cfg['hidden_size'], cfg['num_attention_heads'], cfg['layer_norm_eps']
        if L > cfg['max_position_embeddings']:
            raise ValueError(f'sequence length {L} exceeds max_position_embeddings {cfg["max_position_embeddings"]}')
        e = FrozenEngine(self, _device_of(self))
        e.ids = torch.zeros(B * L, dtype=torch.int64, device=e.dev)
        x = e.buf(B * L, D)
        e.f(ops.embed_tokens, e.ids, e.p32('text_model.embeddings.token_embedding.weight'),
            e.p32('text_model.embeddings.position_embedding.weight'), x, L)
        for i in range(cfg['num_hidden_layers']):
            p = f'text_model.encoder.layers.{i}'
            h = e.ln(x, p + '.layer_norm1', eps)
            q, k, v = e.lin(h, p + '.self_attn.q_proj'), e.lin(h, p + '.self_attn.k_proj'), e.lin(h, p + '.self_attn.v_proj')
            o = e.attention(q, k, v, B, H, L, D // H, causal=True)
            x = e.lin(o, p + '.self_attn.out_proj', residual=x)
            h = e.ln(x, p + '.layer_norm2', eps)
            m = e.lin(h, p + '.mlp.fc1')
            g = e.buf(*m.shape)
            e.f(ops.gelu_fwd, m, g)
            x = e.lin(g, p + '.mlp.fc2', residual=x)
        e.out = e.ln(x, 'text_model.final_layer_norm', eps)
        self._engines[(B, L)] = e
        return e

    def forward(self, input_ids, **_unused):
        """Token ids (B, L) -> (last_hidden_state (B, L, D) bf16,) like `transformers.CLIPTextModel(...)[0]`."""
        if input_ids.device.type != 'cuda':
            raise RuntimeError('diffusion_b200 needs CUDA (sm_100a) tensors: there is no CPU fallback')
        B, L = input_ids.shape
        e = self._engine(B, L)
        e.ids.copy_(input_ids.reshape(-1))
        e.run()
        return _Out(last_hidden_state=e.out.view(B, L, -1).clone())  # the engine's output buffer is reused by the next call

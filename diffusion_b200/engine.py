"""Static kernel schedule of the SD-2 UNet training step (forward + backward) over the sm_100a kernels.

The engine walks the parameter skeleton (diffusion_b200/unet.py) once per input geometry and records two flat lists
of kernel launches (`fwd`, `bwd`) over pre-allocated NHWC bf16 activation buffers - no per-step Python graph, no
autograd tape inside the UNet; after warm-up the two lists are replayed as CUDA graphs.

Parameters live in three flat arenas with identical offsets:
  p32 - fp32 master weights (the nn.Parameters are views into it; 3x3 conv weights are stored [kh][kw][Cout][Cin] and
        exposed with PyTorch's (Cout,Cin,kh,kw) shape through strides, so the implicit-GEMM kernels read/write them
        coalesced and no per-step transposition exists),
  p16 - bf16 shadow used by the tensor-core kernels (one cast kernel per step),
  g32 - fp32 gradients (param.grad are views into it; one contiguous buffer for the DDP all-reduce).

Reference semantics restated: diffusers UNet2DConditionModel.forward, called from
reference diffusion/models/stable_diffusion.py:183; bf16 autocast + fp32 accumulation/statistics as composer runs it
(reference diffusion/train.py:91-108).
"""
from functools import partial

import torch

from diffusion_b200 import ops

BF16 = torch.bfloat16


_GEMM_FUNCS = (ops.linear_fwd, ops.linear_dgrad, ops.linear_wgrad, ops.conv3x3_fwd, ops.conv3x3_dgrad, ops.conv3x3_wgrad,
               ops.bmm, ops.attn_fwd, ops.attn_bwd)


# Output operands of every recorded op: (positional indices after ctx, keyword names).  Every other tensor argument is
# an input.  The two-stream scheduler derives read / write address ranges from this table.
_OP_WRITES = {
    'linear_fwd': ((2,), ('workspace', 'gn_partial')), 'linear_dgrad': ((2,), ('workspace',)), 'linear_wgrad': ((2,), ()),
    'conv3x3_fwd': ((5,), ('workspace', 'gn_partial', 'mse_dpred8', 'mse_acc')), 'conv3x3_dgrad': ((5,), ('workspace',)), 'conv3x3_wgrad': ((5,), ()),
    'attn_fwd': ((3, 4), ()), 'attn_bwd': ((6, 7, 8, 9), ()),
    'groupnorm_fwd': ((3, 4, 5), ()), 'groupnorm_fwd_fused': ((5, 6, 7), ()), 'groupnorm_bwd': ((5, 6, 7, 8), ('drowsum', 'dcolsum', 'dcolsum2')),
    'layernorm_fwd': ((3, 4), ()), 'layernorm_bwd': ((4, 5, 6, 7), ('dcolsum',)),
    'geglu_fwd': ((1,), ()), 'geglu_bwd': ((2,), ('dbias',)), 'silu_fwd': ((1,), ()), 'silu_bwd': ((2,), ()),
    'axpby': ((4,), ()), 'copy2d': ((1,), ()), 'concat_stats': ((2, 3), ()), 'upsample2x_fwd': ((1,), ()), 'upsample2x_bwd': ((1,), ()),
    'phase_split': ((1,), ()), 'phase_merge': ((1,), ()), 'colsum': ((1,), ()), 'cast_f32_to_bf16': ((1,), ()),
    'pad_cast_rows': ((2,), ()), 'unpad_accum_rows': ((2,), ()), 'fill_f32': ((0,), ()),
    'upconv_weff_build': ((1,), ()), 'upconv_wgrad_scatter': ((1,), ()),
}


def _span(t):
    """Address range [lo, hi) touched by a (possibly strided) tensor view."""
    if t.numel() == 0:
        return (t.data_ptr(), t.data_ptr())
    last = sum((n - 1) * st for n, st in zip(t.shape, t.stride()))
    return (t.data_ptr(), t.data_ptr() + (last + 1) * t.element_size())


def _op_io(op):
    """(read ranges, write ranges) of a recorded op (functools.partial(fn, ctx, *args, **kwargs))."""
    name = op.func.__name__
    pos, kws = _OP_WRITES[name]
    args = op.args[1:]
    reads, writes = [], []
    for i, a in enumerate(args):
        if torch.is_tensor(a):
            (writes if i in pos else reads).append(_span(a))
    for k, a in op.keywords.items():
        if torch.is_tensor(a):
            (writes if k in kws else reads).append(_span(a))
    # accumulating ops also read their destination; treating every output as read+write is always safe
    return reads + writes, writes


def _overlap(xs, ys):
    for lo, hi in xs:
        for lo2, hi2 in ys:
            if lo < hi2 and lo2 < hi:
                return True
    return False


class Node:
    """An activation [M, C] (bf16) with its lazily allocated gradient buffer."""
    __slots__ = ('data', 'grad', 'gw', 'M', 'C', 'writes', 'last_writer', 'writes_at_norm', 'reads', 'bias_names', 'rowsum',
                 'claimed', 'producer', 'gn_part')

    def __init__(self, data):
        self.data, self.grad, self.gw = data, None, False
        self.M, self.C = data.shape
        self.writes, self.last_writer, self.writes_at_norm = 0, None, -1  # gradient-write bookkeeping (bias grads from norms)
        # Bias gradients taken from the GroupNorm backward (no separate column-sum pass over the output gradient):
        #   reads      - forward ops recorded so far that consume this node,
        #   bias_names - parameters whose gradient is the column sum of this node's TOTAL gradient (set by the producer),
        #   rowsum     - [B, C] fp32 buffer that wants the per-image sums (time-embedding projection of a ResNet conv1),
        #   claimed    - a GroupNorm is the node's FIRST forward consumer, hence the last writer of its gradient: its
        #                backward kernel emits those sums and the producer skips its own pass.
        self.reads, self.bias_names, self.rowsum, self.claimed = 0, [], None, False
        # the recorded forward GEMM / conv op (a functools.partial) that writes this node as a whole bf16 tensor: a GroupNorm
        # that consumes the node asks that op's epilogue for the statistics (Engine.groupnorm)
        self.producer = None
        self.gn_part = None  # (partials, slab) when the op that wrote this node already took the GroupNorm statistics


def _align(n, a=64):
    return (n + a - 1) // a * a


class ParamArena:
    _live = []  # weak references to every arena (FusedAdamW finds the arena of its parameters here)

    @classmethod
    def live(cls):
        out = [r() for r in cls._live]
        cls._live = [r for r, a in zip(cls._live, out) if a is not None]
        return [a for a in out if a is not None]

    def __init__(self, unet, device, with_grads=True):
        """with_grads=False: frozen modules (VAE, text encoder) get the fp32 arena and the bf16 shadow only."""
        import weakref
        if with_grads:
            ParamArena._live.append(weakref.ref(self))
        self.shadow_version = None  # parameter-version stamp the bf16 shadow was last refreshed at
        # optimizer update applied bucket by bucket DURING the next backward (FusedAdamW.arm): hook(lo, hi) enqueues the update of
        # arena elements [lo, hi) on the current stream; `update_applied` tells the optimizer's step() that it has run
        self.armed_update, self.update_applied = None, False
        self.entries = {}
        off = 0
        plist = list(unet.named_parameters())
        for name, p in plist:
            if p.device != device:
                raise RuntimeError('all UNet parameters must live on the engine device')
            self.entries[name] = (off, p.numel(), tuple(p.shape))
            off = _align(off + p.numel())
        self.total = off
        self.p32 = torch.zeros(off, dtype=torch.float32, device=device)
        self.g32 = torch.zeros(off if with_grads else 0, dtype=torch.float32, device=device)
        self.p16 = torch.zeros(off, dtype=BF16, device=device)
        self.params = {}
        with torch.no_grad():
            for name, p in plist:
                view = self._view(self.p32, name)
                view.copy_(p.data.float())
                p.data = view
                p.grad = None
                self.params[name] = p

    def _view(self, arena, name):
        off, n, shape = self.entries[name]
        flat = arena[off:off + n]
        if len(shape) == 4 and shape[-1] == 3:  # 3x3 conv: storage [kh][kw][Cout][Cin]
            return flat.view(3, 3, shape[0], shape[1]).permute(2, 3, 0, 1)
        return flat.view(shape)

    def storage(self, arena, name):
        """Storage-layout view: [9, Cout, Cin] for 3x3 convs, [Cout, Cin] for 1x1 convs, natural shape otherwise."""
        off, n, shape = self.entries[name]
        flat = arena[off:off + n]
        if len(shape) == 4 and shape[-1] == 3:
            return flat.view(9, shape[0], shape[1])
        if len(shape) == 4:
            return flat.view(shape[0], shape[1])
        return flat.view(shape)

    def fused(self, arena, names):
        """One [sum(out), in] view over adjacent linear weights (to_q|to_k|to_v)."""
        off0, _, shape0 = self.entries[names[0]]
        rows, off = 0, off0
        for nm in names:
            o, n, shape = self.entries[nm]
            if o != off or shape[1] != shape0[1]:
                raise RuntimeError(f'parameters {names} are not adjacent in the arena')
            off, rows = o + n, rows + shape[0]
        return arena[off0:off].view(rows, shape0[1])

    def grad_view(self, name):
        return self._view(self.g32, name)

    def bound(self):
        return all(p.data_ptr() == self.p32.data_ptr() + 4 * self.entries[n][0] for n, p in self.params.items())

    def grads_bound(self):
        return all(p.grad is not None and p.grad.data_ptr() == self.g32.data_ptr() + 4 * self.entries[n][0]
                   for n, p in self.params.items())

    def param_set(self):
        return {id(p) for p in self.params.values()}

    def version(self):
        """Changes whenever any parameter is modified in place through torch (optimizer step, load_state_dict, EMA
        swap): decides whether the bf16 shadow must be re-cast before a forward."""
        return sum(p._version for p in self.params.values())

    def mark_shadow_fresh(self):
        self.shadow_version = self.version()

    def refresh_shadow(self, ctx):
        v = self.version()
        if v != self.shadow_version:
            ops.cast_f32_to_bf16(ctx, self.p32, self.p16)
            self.shadow_version = v


class Engine:

    def __init__(self, unet, B, H, W, L, shared=None, forward_only=False):
        """shared: an engine of the same module built earlier (other geometry): its parameter arena and split-K scratch
        are reused."""
        dev = unet.conv_in.weight.device
        self.ctx = ops.get_ctx(dev)
        self.dev = dev
        self.cfg = unet.config
        if shared is not None and shared.arena.bound():
            self.arena = shared.arena
        elif hasattr(unet, 'bind_arena'):
            self.arena = unet.bind_arena()  # the module's own arena (bound early by the factory, see UNet2DConditionModel.bind_arena)
        else:
            self.arena = ParamArena(unet, dev)
        self.B, self.H, self.W, self.L = B, H, W, L
        self.fwd, self.bwd, self._bwd_builders = [], [], []
        self._touched, self.grad_ready = set(), {}
        self._norm_emitted = set()  # bias parameters whose gradient a GroupNorm backward produced
        import os
        self.fuse_gn_stats = os.environ.get('SD2_NO_GN_FUSION') != '1'  # A/B switch of the epilogue GroupNorm statistics
        self.fold_upsample = os.environ.get('SD2_NO_UPCONV_FOLD') != '1'  # A/B switch: Upsample2D as four 4-tap phase convolutions
        self.fold_upsample_min_rows = int(os.environ.get('SD2_UPCONV_FOLD_MIN_ROWS', '4096'))
        self.gn_part = None  # scratch of those statistics: written by a producer's epilogue, read by the norm right behind it
        reuse = shared is not None and getattr(shared, 'ws', None) is not None
        self.ws = shared.ws if reuse else torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        self.G = self.cfg['norm_num_groups']
        cmax = max(self.cfg['block_out_channels']) * 2
        self.gn_ws = ops.groupnorm_ws(self.ctx, B, cmax, dev)
        self.ln_ws = ops.layernorm_ws(self.ctx, B * H * W, max(self.cfg['block_out_channels']), dev)
        self.act_bytes = 0
        self.attn_ws = None  # scratch of the fused attention backward (fp32 dQ accumulation), shared by all layers
        self.gemm_flops = 0  # algorithmic 2*M*N*K of every tensor-core GEMM recorded (fwd + bwd)
        self.attn_flops = 0  # the fused-attention share of it
        self.fwd_is_gemm, self.bwd_is_gemm = [], []
        self.fwd_side, self.bwd_side = [], []
        self.ws_side = shared.ws_side if reuse else torch.empty(64 << 20, dtype=torch.uint8, device=dev)
        # static inputs (written by K1 / the prep kernels)
        c0 = self.cfg['block_out_channels'][0]
        self.in_x8 = torch.zeros(B * H * W, 8, dtype=BF16, device=dev)
        self.in_temb = torch.zeros(B, c0, dtype=BF16, device=dev)
        self.in_ctx = torch.zeros(B * L, self.cfg['cross_attention_dim'], dtype=BF16, device=dev)
        self.pred8 = None
        self.dpred8 = torch.zeros(B * H * W, 8, dtype=BF16, device=dev)
        self.loss_acc = torch.zeros(2, dtype=torch.float32, device=dev)
        # MSE head in conv_out's epilogue (training engines): the target noise lives in a static buffer that K1 writes and the
        # epilogue reads; its dtype is fixed by the first training forward (before graph capture).  mse_generation = the
        # forward whose loss sum / dL/dpred the epilogue produced.
        self.noise_target = None
        self.mse_generation = -1
        self._conv_out_op = None
        self.forward_only = forward_only  # sampling / eval: no backward schedule, no gradient buffers
        self._build()
        self.graph_fwd = self.graph_bwd = None
        self.generation = 0  # bumped by every run_forward: an autograd node may only replay the backward of ITS forward

    def params_bound(self):
        return self.arena.bound()

    def fused_mse_target(self, like):
        """The static noise buffer the conv_out epilogue compares the prediction with, or None when the MSE head cannot be in
        the epilogue for this call (forward-only engine, another latent dtype than the one the schedule was set up for, graphs
        already captured without it).  The first training forward decides the dtype and arms the recorded conv_out launch."""
        if self.forward_only or self._conv_out_op is None:
            return None
        if self.noise_target is None:
            if self.graph_fwd is not None:
                return None
            self.noise_target = torch.zeros(self.B, 4, self.H, self.W, dtype=like.dtype, device=self.dev)
            self.loss_acc[1] = float(self.B * 4 * self.H * self.W)
            self._conv_out_op.keywords.update(mse_target=self.noise_target, mse_dpred8=self.dpred8, mse_acc=self.loss_acc)
            # the loss sum is cleared at the start of every forward (first recorded op)
            self.fwd.insert(0, partial(ops.fill_f32, self.ctx, self.loss_acc[0:1], 0.0))
            self.fwd_is_gemm.insert(0, False)
            self.fwd_side.insert(0, False)
        if like.dtype != self.noise_target.dtype or tuple(like.shape) != tuple(self.noise_target.shape):
            return None
        return self.noise_target

    def prepare_inputs(self, sample, timestep, enc):
        """Inputs of a plain `unet(sample, timestep, encoder_hidden_states)` call -> the static input buffers."""
        B = self.B
        if not torch.is_tensor(timestep):
            timestep = torch.tensor([timestep], dtype=torch.int64, device=self.dev)
        timestep = timestep.to(device=self.dev, dtype=torch.int64).expand(B).contiguous()
        ops.timestep_embedding(self.ctx, timestep, self.in_temb, sample.dtype)
        ops.nchw4_to_nhwc8(self.ctx, sample.contiguous(), self.in_x8, B, self.H, self.W)
        self.set_context(enc)

    def set_context(self, enc):
        """Text conditioning (B, L, D) in its wire dtype (fp16 from the precomputed-latent dataset, bf16, fp32) -> the
        engine's bf16 context buffer, by the library's cast kernel."""
        if enc.numel() != self.in_ctx.numel():
            raise ValueError(f'conditioning has {tuple(enc.shape)}, engine was built for ({self.B}, {self.L}, {self.in_ctx.shape[1]})')
        ops.cast_to_bf16(self.ctx, enc.contiguous(), self.in_ctx)

    # ---------------------------------------------------------------------------------------------- helpers
    def node(self, M, C, dtype=BF16):
        t = torch.empty(M, C, dtype=dtype, device=self.dev)
        self.act_bytes += t.numel() * t.element_size()
        return Node(t)

    def buf(self, *shape, dtype=BF16, zero=False):
        t = (torch.zeros if zero else torch.empty)(*shape, dtype=dtype, device=self.dev)
        self.act_bytes += t.numel() * t.element_size()
        return t

    def w16(self, name):
        return self.arena.storage(self.arena.p16, name)

    def p32(self, name):
        return self.arena.storage(self.arena.p32, name)

    def g32(self, name):
        self._touched.add(name)  # called from backward builders only: records which builder completes this gradient
        return self.arena.storage(self.arena.g32, name)

    def f(self, fn, *a, side=False, **k):
        """Record a forward op.  side=True: the op is off the critical path (it only feeds later ops) and may run on the
        side stream, concurrently with the main chain; the scheduler inserts the dependencies."""
        self._attach_plan(fn, a, k)
        op = partial(fn, self.ctx, *a, **k)
        self.fwd.append(op)
        self.fwd_is_gemm.append(fn in _GEMM_FUNCS)
        self.fwd_side.append(bool(side))
        self._count_flops(fn, a, k)
        return op

    def b(self, fn, *a, side=None, **k):
        """Record a backward op.  Weight-gradient GEMMs and bias-gradient column sums write only parameter gradients,
        which nothing reads before the end of backward: they go to the side stream by default."""
        self._attach_plan(fn, a, k)
        op = partial(fn, self.ctx, *a, **k)
        if side is None:
            side = fn in (ops.linear_wgrad, ops.conv3x3_wgrad) or (fn is ops.colsum and self._is_param_grad(a[1]))
        self.bwd.append(op)
        self.bwd_is_gemm.append(fn in _GEMM_FUNCS)
        self.bwd_side.append(bool(side))
        self._count_flops(fn, a, k)

    @staticmethod
    def _attach_plan(fn, a, k):
        """Measured (tile width, K split) of this GEMM shape, if tools/autotune_gemm.py has one (ops.gemm_plans)."""
        if fn.__name__ in ops.GEMM_OPS and 'plan' not in k:
            plan = ops.gemm_plans().get(ops.gemm_key(fn.__name__, a, k))
            if plan is not None:
                k['plan'] = plan

    def _is_param_grad(self, t):
        g = self.arena.g32
        return g.data_ptr() <= t.data_ptr() < g.data_ptr() + g.numel() * 4

    def _count_flops(self, fn, a, k=None):
        ntaps = len(k['taps']) if k and k.get('taps') else 9
        if fn in (ops.linear_fwd, ops.linear_dgrad):
            self.gemm_flops += 2 * a[0].shape[0] * a[0].shape[1] * a[1].shape[0 if fn is ops.linear_fwd else 1]
        elif fn is ops.linear_wgrad:
            self.gemm_flops += 2 * a[0].shape[0] * a[0].shape[1] * a[1].shape[1]
        elif fn in (ops.conv3x3_fwd, ops.conv3x3_dgrad):  # (x, B, H, W, w9, ...)
            self.gemm_flops += 2 * a[1] * a[2] * a[3] * a[4].shape[1] * a[4].shape[2] * ntaps
        elif fn is ops.conv3x3_wgrad:  # (dy, x, B, H, W, dw9)
            self.gemm_flops += 2 * a[2] * a[3] * a[4] * a[5].shape[1] * min(a[5].shape[2], a[1].shape[1]) * ntaps
        elif fn is ops.bmm:  # (..., M, N, K, batch, nb0) at positions 9..12
            self.gemm_flops += 2 * a[9] * a[10] * a[11] * a[12]

    def _gout(self, node):
        """Gradient buffer of `node` for a backward op to write: (buffer, accumulate?)."""
        if node.grad is None:
            node.grad = torch.empty_like(node.data)
            self.act_bytes += node.grad.numel() * 2
        acc, node.gw = node.gw, True
        node.writes += 1
        return node.grad, acc

    def _pass(self, g, node):
        """d(node) += g for an identity edge: alias the buffer when node has no gradient yet."""
        node.writes += 1
        if not node.gw:
            node.grad, node.gw = g, True
        else:
            self.b(ops.axpby, node.grad, 1.0, g, 1.0, node.grad)

    # ---------------------------------------------------------------------------------------------- records
    def linear(self, x, wname, bname=None, residual=None, w16=None, gw=None, gnames=(), side=False, bias_grad_elsewhere=False):
        w = self.w16(wname) if w16 is None else w16
        gwv = self.arena.storage(self.arena.g32, wname) if gw is None else gw
        N = w.shape[0]
        out = self.node(x.M, N)
        bias = self.p32(bname) if bname else None
        self._use(x, residual)
        if bname and not bias_grad_elsewhere:
            out.bias_names = [bname]
        op = self.f(ops.linear_fwd, x.data, w, out.data, bias=bias, residual=residual.data if residual else None,
                    workspace=self.ws_side if side else self.ws, side=side)
        if not side:
            out.producer = op

        def bwd():
            g = out.grad
            assert out.gw, f'no gradient reaches output of {wname}'
            if residual is not None:
                self._pass(g, residual)
            if x is not None and x.data is not self.in_ctx and x.data is not self.in_temb:
                gx, acc = self._gout(x)
                self.b(ops.linear_dgrad, g, w, gx, residual=gx if acc else None, workspace=self.ws)
            self._touched.update(gnames or (wname,))
            self.b(ops.linear_wgrad, g, x.data, gwv)
            if bname and not bias_grad_elsewhere:
                if self._bias_from_norm(out, bname):
                    self._touched.add(bname)
                else:
                    self.b(ops.colsum, g, self.g32(bname), 1, x.M, True)
            elif bname and bias_grad_elsewhere == 'deferred':
                # conv_shortcut: its output is only the residual of conv2, so its output gradient IS conv2's; when a GroupNorm
                # emitted conv2's bias gradient it emitted this one with it (Node.bias_names of conv2's output)
                if bname in getattr(self, '_norm_emitted', ()):
                    self._touched.add(bname)
                else:
                    self.b(ops.colsum, g, self.g32(bname), 1, x.M, True)
            elif bname and bias_grad_elsewhere == 'norm':
                # the norm that emitted the column sums must have been the LAST op to write this node's gradient
                assert getattr(out, 'last_writer', None) is not None and out.writes == out.writes_at_norm, \
                    f'{bname}: the bias gradient was taken from a norm backward that is not the last writer of the node'

        self._bwd_builders.append(bwd)
        return out

    def silu(self, x):
        out = self.node(x.M, x.C)
        self._use(x)
        self.f(ops.silu_fwd, x.data, out.data)

        def bwd():
            gx, acc = self._gout(x)
            assert not acc
            self.b(ops.silu_bwd, x.data, out.grad, gx)

        self._bwd_builders.append(bwd)
        return out

    def _gn_part_floats(self):
        # largest [rows / 16][C][2] any level can ask for: rows * C is largest at the top level (C0 channels, or 2*C0 never
        # fused: concatenations have no producing GEMM)
        return (self.B * self.H * self.W // 16) * self.cfg['block_out_channels'][0] * 2

    def _gn_scratch(self, need):
        if self.gn_part is None or self.gn_part.numel() < need:
            self.gn_part = torch.empty(max(need, self._gn_part_floats()), dtype=torch.float32, device=self.dev)
        return self.gn_part[:need]

    def _use(self, *nodes):
        for n in nodes:
            if n is not None:
                n.reads += 1

    def _bias_from_norm(self, out, bname):
        """True if a GroupNorm backward already produced the gradient of `bname` (it claimed the node at forward time);
        checks that the norm really was the last writer of the node's gradient."""
        if not out.claimed or bname not in out.bias_names:
            return False
        assert out.last_writer is not None and out.writes == out.writes_at_norm, \
            f'{bname}: bias gradient was taken from a GroupNorm backward that is not the last writer of the node'
        return True

    def groupnorm(self, x, prefix, eps, silu, HW):
        y = self.node(x.M, x.C)
        stats = self.buf(self.B, self.G, 2, dtype=torch.float32)
        gamma, beta = self.p32(prefix + '.weight'), self.p32(prefix + '.bias')
        # first forward consumer of a conv / linear output = last writer of its gradient: this norm's backward emits the
        # producer's bias gradient(s) (and the per-image sums a ResNet conv1 needs) as column sums of the dx it writes
        if x.reads == 0 and (x.bias_names or x.rowsum is not None) and len(x.bias_names) <= 2 and not self.forward_only:
            x.claimed = True
        self._use(x)
        # GroupNorm fused with the producing conv / linear: that op's epilogue takes the per-slab column sums of the tensor it
        # writes, this norm only combines them and streams the apply(+SiLU) pass (north_star: GroupNorm in the conv epilogue)
        slab = 32 if HW % 32 == 0 else (16 if HW % 16 == 0 else 0)
        fused_args = None
        if x.gn_part is not None and self.fwd and self.fwd[-1].func is ops.concat_stats:
            fused_args = x.gn_part  # the skip concatenation recorded right before this norm took the statistics
        elif x.producer is not None and slab and 'gn_partial' not in x.producer.keywords and self.fuse_gn_stats:
            # the statistics scratch is shared: nothing recorded since the producer may write it
            at = next(i for i in range(len(self.fwd) - 1, -1, -1) if self.fwd[i] is x.producer)
            if not any('gn_partial' in op.keywords or op.func is ops.concat_stats for op in self.fwd[at + 1:]):
                need = (x.M // slab) * x.C * 2
                part = self._gn_scratch(need).view(x.M // slab, x.C, 2)
                x.producer.keywords['gn_partial'] = part
                x.producer.keywords['gn_slab'] = slab
                fused_args = (part, slab)
        if fused_args is not None:
            self.f(ops.groupnorm_fwd_fused, x.data, fused_args[0], fused_args[1], gamma, beta, y.data, stats, self.gn_ws, self.B,
                   HW, self.G, eps, silu)
        else:
            self.f(ops.groupnorm_fwd, x.data, gamma, beta, y.data, stats, self.gn_ws, self.B, HW, self.G, eps, silu)

        def bwd():
            assert y.gw
            gx, acc = self._gout(x)
            extra = {}
            if x.claimed:
                names = list(x.bias_names)
                if x.rowsum is not None:
                    extra['drowsum'] = x.rowsum
                if len(names) > 0:
                    extra['dcolsum'] = self.arena.storage(self.arena.g32, names[0])
                if len(names) > 1:
                    extra['dcolsum2'] = self.arena.storage(self.arena.g32, names[1])
                self._norm_emitted.update(names)
            self.b(ops.groupnorm_bwd, y.grad, x.data, gamma, beta, stats, gx, self.g32(prefix + '.weight'),
                   self.g32(prefix + '.bias'), self.gn_ws, self.B, HW, self.G, silu, dx_add=gx if acc else None, **extra)
            x.last_writer, x.writes_at_norm = prefix, x.writes

        self._bwd_builders.append(bwd)
        return y

    def layernorm(self, x, prefix, colsum_into=None):
        """colsum_into: bias name of the linear that produced x.  This norm is the first forward consumer of x, hence the
        last contributor to x.grad in backward: the dx (+ dx_add) it writes is the total output gradient of that linear,
        and its column sums - accumulated by the same kernel - are the bias gradient (that linear skips its own pass)."""
        y = self.node(x.M, x.C)
        stats = self.buf(x.M, 2, dtype=torch.float32)
        gamma, beta = self.p32(prefix + '.weight'), self.p32(prefix + '.bias')
        self._use(x)
        self.f(ops.layernorm_fwd, x.data, gamma, beta, y.data, stats)

        def bwd():
            assert y.gw
            gx, acc = self._gout(x)
            self.b(ops.layernorm_bwd, y.grad, x.data, gamma, stats, gx, self.g32(prefix + '.weight'),
                   self.g32(prefix + '.bias'), self.ln_ws, dx_add=gx if acc else None,
                   dcolsum=self.g32(colsum_into) if colsum_into else None)
            x.last_writer, x.writes_at_norm = prefix, x.writes  # checked by the producing linear (bias_grad_elsewhere='norm')

        self._bwd_builders.append(bwd)
        return y

    def conv3(self, x, Hc, Wc, prefix, rowbias=None, residual=None, rowbias_bwd=None, rowsum=None, extra_bias=()):
        """3x3 stride-1 conv (+bias +per-image bias +residual).  rowbias_bwd(g, have_rowsum) is called with the output
        gradient; rowsum = [B, Cout] fp32 buffer for the per-image sums of that gradient (filled by the GroupNorm backward
        behind this conv when it can, have_rowsum tells); extra_bias: further parameters whose gradient is the column sum
        of this conv's output gradient (the bias of the conv_shortcut that feeds its residual)."""
        w, bias = self.w16(prefix + '.weight'), self.p32(prefix + '.bias')
        out = self.node(x.M, w.shape[1])
        out.bias_names = [prefix + '.bias'] + list(extra_bias)
        out.rowsum = rowsum
        B = self.B
        self._use(x, residual)
        out.producer = self.f(ops.conv3x3_fwd, x.data, B, Hc, Wc, w, out.data, bias=bias, rowbias=rowbias,
                              residual=residual.data if residual else None, workspace=self.ws)

        def bwd():
            g = out.grad
            assert out.gw
            from_norm = self._bias_from_norm(out, prefix + '.bias')
            if residual is not None:
                self._pass(g, residual)
            if rowbias_bwd is not None:
                rowbias_bwd(g, from_norm)
            gx, acc = self._gout(x)
            self.b(ops.conv3x3_dgrad, g, B, Hc, Wc, w, gx, residual=gx if acc else None, workspace=self.ws)
            self.b(ops.conv3x3_wgrad, g, x.data, B, Hc, Wc, self.g32(prefix + '.weight'))
            if from_norm:
                self._touched.add(prefix + '.bias')
            else:
                self.b(ops.colsum, g, self.g32(prefix + '.bias'), 1, x.M, True)

        self._bwd_builders.append(bwd)
        return out

    def downsample(self, x, Hc, Wc, prefix):
        B, C = self.B, x.C
        Ho, Wo = Hc // 2, Wc // 2
        Mo = B * Ho * Wo
        w, bias = self.w16(prefix + '.weight'), self.p32(prefix + '.bias')
        planes = self.buf(4 * Mo, C)
        out = self.node(Mo, C)
        out.bias_names = [prefix + '.bias']
        taps = ops.taps_stride2(B)
        self._use(x)
        self.f(ops.phase_split, x.data, planes, B, Hc, Wc)
        out.producer = self.f(ops.conv3x3_fwd, planes, B, Ho, Wo, w, out.data, bias=bias, taps=taps, n_planes=4 * B,
                              workspace=self.ws)

        def bwd():
            g = out.grad
            assert out.gw
            dplanes = self.buf(4 * Mo, C)
            for plane, sub in ops.taps_stride2_dgrad().items():
                self.b(ops.conv3x3_dgrad, g, B, Ho, Wo, w, dplanes[plane * Mo:(plane + 1) * Mo], taps=sub, workspace=self.ws)
            gx, acc = self._gout(x)
            if acc:
                tmp = self.buf(x.M, C)
                self.b(ops.phase_merge, dplanes, tmp, B, Hc, Wc)
                self.b(ops.copy2d, tmp, gx, x.M, C, True)
            else:
                self.b(ops.phase_merge, dplanes, gx, B, Hc, Wc)
            self.b(ops.conv3x3_wgrad, g, planes, B, Ho, Wo, self.g32(prefix + '.weight'), taps=taps, n_planes=4 * B)
            if self._bias_from_norm(out, prefix + '.bias'):
                self._touched.add(prefix + '.bias')
            else:
                self.b(ops.colsum, g, self.g32(prefix + '.bias'), 1, Mo, True)

        self._bwd_builders.append(bwd)
        return out

    def upsample(self, x, Hc, Wc, prefix):
        """Upsample2D (nearest x2, then 3x3 conv) without the 4x tensor: the four output phases are 4-tap convolutions of the
        low-resolution input with summed weights (ops.upconv_weff_build) - 16 instead of 36 tap-products per input pixel, in
        forward, dgrad and wgrad alike; the phase planes are interleaved by the stride-2 merge kernel."""
        # below ~4096 low-resolution pixels the twelve small launches cost more than the FLOPs they save (microbatch 16: 822 -> 811
        # img/s with everything folded)
        if not self.fold_upsample or x.M < self.fold_upsample_min_rows:
            return self._upsample_materialised(x, Hc, Wc, prefix)
        B, C, M = self.B, x.C, x.M
        w32, bias = self.p32(prefix + '.weight'), self.p32(prefix + '.bias')  # fp32 master taps [9, Cout, Cin]
        Cout = w32.shape[1]
        weff = self.buf(16, Cout, C)
        planes = self.buf(4 * M, Cout)
        out = self.node(4 * M, Cout)
        self._use(x)
        self.f(ops.upconv_weff_build, w32, weff)
        for ph in range(4):
            self.f(ops.conv3x3_fwd, x.data, B, Hc, Wc, weff[4 * ph:4 * ph + 4], planes[ph * M:(ph + 1) * M], bias=bias,
                   taps=ops.taps_upconv(ph), workspace=self.ws)
        self.f(ops.phase_merge, planes, out.data, B, 2 * Hc, 2 * Wc)

        def bwd():
            g = out.grad
            assert out.gw
            gplanes = self.buf(4 * M, Cout)
            dweff = self.buf(16, Cout, C, dtype=torch.float32)
            self.b(ops.phase_split, g, gplanes, B, 2 * Hc, 2 * Wc)
            gx, acc = self._gout(x)
            for ph in range(4):
                self.b(ops.conv3x3_dgrad, gplanes[ph * M:(ph + 1) * M], B, Hc, Wc, weff[4 * ph:4 * ph + 4], gx,
                       residual=gx if (acc or ph > 0) else None, taps=ops.taps_upconv_dgrad(ph), workspace=self.ws)
            self.b(ops.fill_f32, dweff.view(-1), 0.0)
            for ph in range(4):
                self.b(ops.conv3x3_wgrad, gplanes[ph * M:(ph + 1) * M], x.data, B, Hc, Wc, dweff[4 * ph:4 * ph + 4],
                       taps=ops.taps_upconv(ph))
            self.b(ops.upconv_wgrad_scatter, dweff, self.g32(prefix + '.weight'))
            self.b(ops.colsum, g, self.g32(prefix + '.bias'), 1, 4 * M, True)

        self._bwd_builders.append(bwd)
        return out

    def _upsample_materialised(self, x, Hc, Wc, prefix):
        up = self.node(4 * x.M, x.C)
        self._use(x)
        self.f(ops.upsample2x_fwd, x.data, up.data, self.B, Hc, Wc)

        def bwd():
            assert up.gw
            gx, acc = self._gout(x)
            assert not acc
            self.b(ops.upsample2x_bwd, up.grad, gx, self.B, Hc, Wc)

        self._bwd_builders.append(bwd)
        return self.conv3(up, 2 * Hc, 2 * Wc, prefix)

    def concat(self, a, b_):
        out = self.node(a.M, a.C + b_.C)
        self._use(a, b_)
        HW = a.M // self.B
        if self.fuse_gn_stats and (a.C + b_.C) // 8 <= 512:
            # one kernel copies both halves and takes the GroupNorm statistics of the result (its consumer is a ResNet norm1)
            P = 1
            while self.B * P < 4 * self.ctx.num_sms and P * 2 <= max(1, HW // 16) and P < 64:
                P *= 2
            part = self._gn_scratch(self.B * P * out.C * 2).view(self.B * P, out.C, 2)
            self.f(ops.concat_stats, a.data, b_.data, out.data, part, self.B, HW, P)
            out.gn_part = (part, HW // P)
        else:
            self.f(ops.copy2d, a.data, out.data[:, :a.C], a.M, a.C)
            self.f(ops.copy2d, b_.data, out.data[:, a.C:], a.M, b_.C)

        def bwd():
            assert out.gw
            ga, acc_a = self._gout(a)
            self.b(ops.copy2d, out.grad[:, :a.C], ga, a.M, a.C, acc_a)
            gb, acc_b = self._gout(b_)
            self.b(ops.copy2d, out.grad[:, a.C:], gb, a.M, b_.C, acc_b)

        self._bwd_builders.append(bwd)
        return out

    def geglu(self, h, bias_name=None):
        """bias_name: bias of the projection that produced h; its gradient (the column sums of dh) is then accumulated by
        the GEGLU backward kernel itself and that linear skips its own column-sum pass."""
        y = self.node(h.M, h.C // 2)
        self._use(h)
        self.f(ops.geglu_fwd, h.data, y.data)

        def bwd():
            gh, acc = self._gout(h)
            assert not acc
            self.b(ops.geglu_bwd, h.data, y.grad, gh, dbias=self.g32(bias_name) if bias_name else None)

        self._bwd_builders.append(bwd)
        return y

    def attention(self, q, k, v, qn, kvn, q_col, k_col, v_col, Nq, Nk, heads, C):
        """softmax(q k^T / sqrt(d)) v per (image, head), fused flash-style kernel (csrc/attn.cu): scores and
        probabilities never leave the SM; only the output and the per-row log-sum-exp are saved for backward.
        q/k/v are column slices (offset *_col) of nodes qn / kvn."""
        B, d = self.B, C // heads
        scale = float(d)**-0.5
        out = self.node(B * Nq, C)
        lse = self.buf(B * heads, Nq, dtype=torch.float32)
        self._use(qn, kvn)
        self.f(ops.attn_fwd, q, k, v, out.data, lse, B, heads, Nq, Nk, scale)
        self.gemm_flops += 4 * B * heads * Nq * Nk * d
        self.attn_flops += 4 * B * heads * Nq * Nk * d
        need = self.ctx.lib.sd2_attn_bwd_ws_bytes(B, heads, Nq)
        if self.attn_ws is None or self.attn_ws.numel() < need:
            self.attn_ws = torch.empty(need, dtype=torch.uint8, device=self.dev)

        def bwd():
            assert out.gw
            dO = out.grad
            if qn.grad is None:
                qn.grad = torch.empty_like(qn.data)
            if kvn.grad is None:
                kvn.grad = torch.empty_like(kvn.data)
            dq = qn.grad[:, q_col:q_col + C]
            dk = kvn.grad[:, k_col:k_col + C]
            dv = kvn.grad[:, v_col:v_col + C]
            self.b(ops.attn_bwd, q, k, v, out.data, dO, lse, dq, dk, dv, self.attn_ws, B, heads, Nq, Nk, scale)
            self.gemm_flops += 10 * B * heads * Nq * Nk * d
            self.attn_flops += 10 * B * heads * Nq * Nk * d
            qn.gw = kvn.gw = True

        self._bwd_builders.append(bwd)
        return out

    # ---------------------------------------------------------------------------------------------- blocks
    def resnet(self, x, prefix, Hc, Wc, semb):
        arena = self.arena
        B, HW = self.B, Hc * Wc
        eps = self.cfg['norm_eps']
        cout = arena.entries[prefix + '.conv1.weight'][2][0]
        has_sc = (prefix + '.conv_shortcut.weight') in arena.entries
        a1 = self.groupnorm(x, prefix + '.norm1', eps, 1, HW)
        # time-embedding projection -> per-image bias (fp32) added in conv1's epilogue
        wt, bt = self.w16(prefix + '.time_emb_proj.weight'), self.p32(prefix + '.time_emb_proj.bias')
        rb = self.buf(B, cout, dtype=torch.float32)
        self.f(ops.linear_fwd, semb.data, wt, rb, bias=bt, out_f32=True, side=True)
        d_tp32 = self.buf(B, cout, dtype=torch.float32)
        d_tp16 = self.buf(B, cout)

        def tproj_bwd(g, have_rowsum):
            # d(time_emb_proj out)[b] = sum over the image's pixels of d(h1): emitted by norm2's backward when it could
            # (have_rowsum), else one column-sum pass per image
            if not have_rowsum:
                self.b(ops.colsum, g, d_tp32, B, HW, False, side=True)
            self.b(ops.cast_f32_to_bf16, d_tp32.view(-1), d_tp16.view(-1), side=True)
            gs, acc = self._gout(semb)
            self.b(ops.linear_dgrad, d_tp16, wt, gs, residual=gs if acc else None, side=True)
            self.b(ops.linear_wgrad, d_tp16, semb.data, self.g32(prefix + '.time_emb_proj.weight'))
            self.b(ops.colsum, d_tp16, self.g32(prefix + '.time_emb_proj.bias'), 1, B, True)

        h1 = self.conv3(a1, Hc, Wc, prefix + '.conv1', rowbias=rb, rowbias_bwd=tproj_bwd, rowsum=d_tp32)
        a2 = self.groupnorm(h1, prefix + '.norm2', eps, 1, HW)
        if has_sc:
            sc = self.linear(x, prefix + '.conv_shortcut.weight', prefix + '.conv_shortcut.bias', bias_grad_elsewhere='deferred')
            return self.conv3(a2, Hc, Wc, prefix + '.conv2', residual=sc, extra_bias=(prefix + '.conv_shortcut.bias',))
        return self.conv3(a2, Hc, Wc, prefix + '.conv2', residual=x)

    def transformer(self, x, prefix, Hc, Wc, heads):
        arena = self.arena
        C, B, L = x.C, self.B, self.L
        HW = Hc * Wc
        n = self.groupnorm(x, prefix + '.norm', 1e-6, 0, HW)
        nb = 'norm'  # bias gradients of proj_in / attn1.to_out / attn2.to_out come from the LayerNorm backward behind them
        h0 = self.linear(n, prefix + '.proj_in.weight', prefix + '.proj_in.bias', bias_grad_elsewhere=nb)
        tb = prefix + '.transformer_blocks.0'
        # --- self attention
        l1 = self.layernorm(h0, tb + '.norm1', colsum_into=prefix + '.proj_in.bias' if nb else None)
        names = [tb + '.attn1.to_q.weight', tb + '.attn1.to_k.weight', tb + '.attn1.to_v.weight']
        qkv = self.linear(l1, names[0], w16=arena.fused(arena.p16, names), gw=arena.fused(arena.g32, names), gnames=names)
        o1 = self.attention(qkv.data[:, :C], qkv.data[:, C:2 * C], qkv.data[:, 2 * C:], qkv, qkv, 0, C, 2 * C, HW, HW, heads, C)
        h1 = self.linear(o1, tb + '.attn1.to_out.0.weight', tb + '.attn1.to_out.0.bias', residual=h0, bias_grad_elsewhere=nb)
        # --- cross attention over the text context
        l2 = self.layernorm(h1, tb + '.norm2', colsum_into=tb + '.attn1.to_out.0.bias' if nb else None)
        q2 = self.linear(l2, tb + '.attn2.to_q.weight')
        names = [tb + '.attn2.to_k.weight', tb + '.attn2.to_v.weight']
        kv = self.linear(self.ctx_node, names[0], w16=arena.fused(arena.p16, names), gw=arena.fused(arena.g32, names),
                         gnames=names, side=True)  # depends only on the text context: off the critical path
        o2 = self.attention(q2.data, kv.data[:, :C], kv.data[:, C:], q2, kv, 0, 0, C, HW, L, heads, C)
        h2 = self.linear(o2, tb + '.attn2.to_out.0.weight', tb + '.attn2.to_out.0.bias', residual=h1, bias_grad_elsewhere=nb)
        # --- GEGLU feed-forward
        l3 = self.layernorm(h2, tb + '.norm3', colsum_into=tb + '.attn2.to_out.0.bias' if nb else None)
        fuse = True  # bias gradient of ff.net.0.proj is accumulated by geglu_bwd
        ff1 = self.linear(l3, tb + '.ff.net.0.proj.weight', tb + '.ff.net.0.proj.bias', bias_grad_elsewhere=fuse)
        gg = self.geglu(ff1, bias_name=tb + '.ff.net.0.proj.bias' if fuse else None)
        h3 = self.linear(gg, tb + '.ff.net.2.weight', tb + '.ff.net.2.bias', residual=h2)
        return self.linear(h3, prefix + '.proj_out.weight', prefix + '.proj_out.bias', residual=x)

    # ---------------------------------------------------------------------------------------------- whole network
    def _build(self):
        cfg, arena, ctx = self.cfg, self.arena, self.ctx
        B, H, W = self.B, self.H, self.W
        boc, heads = cfg['block_out_channels'], cfg['attention_head_dim']
        M = B * H * W
        # ---- per-step weight refresh: fp32 master -> bf16 shadow (+ the two padded special cases)
        cin_w = self.p32('conv_in.weight')  # storage [9, C0, 4]
        self.w_in16 = self.buf(9, boc[0], 8)
        self.f(ops.pad_cast_rows, cin_w.reshape(-1), 4, self.w_in16, 8, 9 * boc[0])
        self.b_out8 = self.buf(8, dtype=torch.float32, zero=True)
        b_out = self.p32('conv_out.bias')
        self.f(ops.unpad_accum_rows, b_out, 4, self.b_out8, 4, 1, False)
        # ---- time embedding
        self.ctx_node = Node(self.in_ctx)
        temb = Node(self.in_temb)
        e1 = self.linear(temb, 'time_embedding.linear_1.weight', 'time_embedding.linear_1.bias')
        emb = self.linear(self.silu(e1), 'time_embedding.linear_2.weight', 'time_embedding.linear_2.bias')
        semb = self.silu(emb)
        # ---- conv_in (4 latent channels zero-padded to 8 so that the pixel stride is 16 bytes)
        x8 = Node(self.in_x8)
        x = self.node(M, boc[0])
        x.bias_names = ['conv_in.bias']
        x.producer = self.f(ops.conv3x3_fwd, x8.data, B, H, W, self.w_in16, x.data, bias=self.p32('conv_in.bias'),
                            workspace=self.ws)
        gw_in = self.buf(9, boc[0], 8, dtype=torch.float32)
        x0 = x  # `x` is rebound below; the closure must keep conv_in's own output node

        def conv_in_bwd():
            assert x0.gw
            self.b(ops.fill_f32, gw_in, 0.0)
            self.b(ops.conv3x3_wgrad, x0.grad, x8.data, B, H, W, gw_in)
            self.b(ops.unpad_accum_rows, gw_in.view(-1), 8, self.g32('conv_in.weight').reshape(-1), 4, 9 * boc[0], True)
            if self._bias_from_norm(x0, 'conv_in.bias'):
                self._touched.add('conv_in.bias')
            else:
                self.b(ops.colsum, x0.grad, self.g32('conv_in.bias'), 1, M, True)

        self._bwd_builders.append(conv_in_bwd)
        # ---- down path
        skips = [(x, H, W)]
        Hc, Wc = H, W
        for i, t in enumerate(cfg['down_block_types']):
            for j in range(cfg['layers_per_block']):
                x = self.resnet(x, f'down_blocks.{i}.resnets.{j}', Hc, Wc, semb)
                if t == 'CrossAttnDownBlock2D':
                    x = self.transformer(x, f'down_blocks.{i}.attentions.{j}', Hc, Wc, heads[i])
                skips.append((x, Hc, Wc))
            if i != len(boc) - 1:
                x = self.downsample(x, Hc, Wc, f'down_blocks.{i}.downsamplers.0.conv')
                Hc, Wc = Hc // 2, Wc // 2
                skips.append((x, Hc, Wc))
        # ---- mid
        x = self.resnet(x, 'mid_block.resnets.0', Hc, Wc, semb)
        x = self.transformer(x, 'mid_block.attentions.0', Hc, Wc, heads[-1])
        x = self.resnet(x, 'mid_block.resnets.1', Hc, Wc, semb)
        # ---- up path
        rheads = heads[::-1]
        for i, t in enumerate(cfg['up_block_types']):
            for j in range(cfg['layers_per_block'] + 1):
                s, sh, sw = skips.pop()
                assert (sh, sw) == (Hc, Wc)
                x = self.resnet(self.concat(x, s), f'up_blocks.{i}.resnets.{j}', Hc, Wc, semb)
                if t == 'CrossAttnUpBlock2D':
                    x = self.transformer(x, f'up_blocks.{i}.attentions.{j}', Hc, Wc, rheads[i])
            if i != len(boc) - 1:
                x = self.upsample(x, Hc, Wc, f'up_blocks.{i}.upsamplers.0.conv')
                Hc, Wc = Hc * 2, Wc * 2
        # ---- head: GN + SiLU + conv_out (4 output channels padded to 8)
        n = self.groupnorm(x, 'conv_norm_out', cfg['norm_eps'], 1, H * W)
        w_out = self.w16('conv_out.weight')  # [9, 4, C0]
        if self.pred8 is None:
            self.pred8 = self.buf(M, 8)
        self._conv_out_op = self.f(ops.conv3x3_fwd, n.data, B, H, W, w_out, self.pred8, bias=self.b_out8, workspace=self.ws)
        gb8 = self.buf(8, dtype=torch.float32)

        def head_bwd():
            g = self.dpred8
            gn_, acc = self._gout(n)
            self.b(ops.conv3x3_dgrad, g, B, H, W, w_out, gn_, workspace=self.ws)
            self.b(ops.conv3x3_wgrad, g[:, :4], n.data, B, H, W, self.g32('conv_out.weight'))
            self.b(ops.colsum, g, gb8, 1, M, False)
            self.b(ops.unpad_accum_rows, gb8, 8, self.g32('conv_out.bias'), 4, 1, True)

        self._bwd_builders.append(head_bwd)
        if self.forward_only:
            self._bwd_builders = None
            self.buckets, self.segments = [], []
            self._hoist_forward_side_ops()
            return
        for builder in reversed(self._bwd_builders):
            self._touched = set()
            builder()
            for name in self._touched:  # gradient of `name` is final once bwd[:len(self.bwd)] has run
                self.grad_ready[name] = len(self.bwd)
        self._bwd_builders = None
        missing = [n for n in arena.entries if n not in self.grad_ready]
        assert not missing, f'parameters without a gradient-producing op: {missing[:5]}'
        self._make_buckets()
        self._hoist_forward_side_ops()

    def _hoist_forward_side_ops(self):
        """Move every forward side op (time-embedding projections, context K/V projections) to the earliest list
        position its inputs allow, so that the side stream starts them at the beginning of the step instead of right
        before their consumer."""
        ios = [_op_io(op) for op in self.fwd]
        order = list(range(len(self.fwd)))
        for i in range(len(self.fwd)):
            if not self.fwd_side[i]:
                continue
            pos = order.index(i)
            j = pos
            while j > 0:
                prev = order[j - 1]
                pr, pw = ios[prev]
                r, w = ios[i]
                if _overlap(r, pw) or _overlap(w, pr):  # true dependency (or shared scratch): stop here
                    break
                j -= 1
            if j != pos:
                order.pop(pos)
                order.insert(j, i)
        self.fwd = [self.fwd[i] for i in order]
        self.fwd_side = [self.fwd_side[i] for i in order]
        self.fwd_is_gemm = [self.fwd_is_gemm[i] for i in order]

    # ---------------------------------------------------------------------------------------------- data parallel (K5)
    def _make_buckets(self, n_buckets=8):
        """Contiguous arena ranges, in the order their gradients complete during backward (end of the arena first).
        bucket = (elem_begin, elem_end, bwd_op_index_after_which_it_is_complete)."""
        arena = self.arena
        names = list(arena.entries)  # arena (= forward) order
        target = arena.total / n_buckets
        buckets, hi, ready = [], arena.total, 0
        for name in reversed(names):
            off = arena.entries[name][0]
            ready = max(ready, self.grad_ready[name])
            if hi - off >= target or off == 0:
                buckets.append([off, hi, ready])
                hi, ready = off, 0
        # a bucket can only be sent once every earlier-sent bucket boundary is also respected: make `ready` monotone
        for i in range(1, len(buckets)):
            buckets[i][2] = max(buckets[i][2], buckets[i - 1][2])
        buckets[-1][2] = len(self.bwd)
        self.buckets = [tuple(b) for b in buckets]
        self.segments = sorted(set(b[2] for b in self.buckets))

    def enable_grad_sync(self, group=None):
        """Average gradients over the data-parallel group, bucket by bucket, overlapped with the rest of backward.
        On CUDA the all-reduce is the library's own entry point (sd2_ddp_allreduce_bucket: NCCL on the communication stream,
        communicator created here from a unique id that torch.distributed merely ships between the ranks); on CPU (gloo,
        host-logic tests) it is torch.distributed's."""
        import torch.distributed as dist
        self.dp_group = group
        self.dp_world = dist.get_world_size(group)
        self.comm_stream = torch.cuda.Stream(self.dev) if self.dev.type == 'cuda' else None
        if self.comm_stream is not None and not ops.dry_run():
            ops.ddp_init(self.ctx, group)
        self.sync_grads = True

    def no_sync(self):
        """Context manager: backward passes inside accumulate locally (the analogue of DistributedDataParallel.no_sync for
        all but the last microbatch of an optimizer step)."""
        import contextlib

        @contextlib.contextmanager
        def cm():
            prev = getattr(self, 'sync_grads', False)
            self.sync_grads = False
            try:
                yield
            finally:
                self.sync_grads = prev

        return cm()

    def _update_points(self):
        """For every gradient bucket: the end of the backward segment after which NO later backward op touches the bucket's
        parameters (fp32 master, bf16 shadow) or gradients - from there on the optimizer may rewrite them while the rest of
        backward is still running (FusedAdamW.arm).  Derived from the recorded ops' address ranges, not from layer order."""
        if getattr(self, '_update_at', None) is None:
            a = self.arena
            ios = [_op_io(op)[0] for op in self.bwd]  # reads + writes of every backward op
            pts = []
            for lo, hi, ready in self.buckets:
                spans = [(t.data_ptr() + lo * t.element_size(), t.data_ptr() + hi * t.element_size()) for t in (a.p32, a.p16, a.g32)]
                last = ready
                for i in range(len(self.bwd) - 1, ready - 1, -1):
                    if _overlap(ios[i], spans):
                        last = i + 1
                        break
                pts.append(min(e for e in self.segments if e >= last))
            self._update_at = pts
        return self._update_at

    def _side_stream(self):
        if getattr(self, 'comm_stream', None) is None:
            self.comm_stream = torch.cuda.Stream(self.dev)
        return self.comm_stream

    def _allreduce_bucket(self, lo, hi):
        import torch.distributed as dist
        flat = self.arena.g32[lo:hi]
        if self.comm_stream is None:  # CPU (gloo) path used by the host-logic tests
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.dp_group)
            flat.div_(self.dp_world)
            return
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.dev))
        self.comm_stream.wait_event(ev)
        ops.ddp_allreduce_bucket(self.ctx, flat, True, self.comm_stream)

    # ---------------------------------------------------------------------------------------------- execution
    def run_forward(self):
        self.generation += 1
        self.arena.refresh_shadow(self.ctx)  # fp32 master -> bf16 shadow, only if the weights changed outside FusedAdamW
        if self.graph_fwd is not None:
            self.graph_fwd.replay()
        else:
            for op in self.fwd:
                op()

    def run_backward(self, allow_update=False):
        if self.forward_only:
            raise RuntimeError('this engine was built forward-only (sampling / eval): it has no backward schedule')
        sync = getattr(self, 'sync_grads', False) and getattr(self, 'dp_world', 1) > 1
        # an armed optimizer update (FusedAdamW.arm) runs bucket by bucket on the side stream, behind the bucket's all-reduce
        update = self.arena.armed_update if allow_update and (sync or getattr(self, 'dp_world', 1) == 1) else None
        if not sync and update is None:
            if self.graph_bwd is not None:
                for g in self.graph_bwd:
                    g.replay()
            else:
                for op in self.bwd:
                    op()
            return
        # segment k ends where bucket(s) with ready == segments[k] become final -> launch their all-reduce on the
        # communication stream while the next segment computes
        start = 0
        for k, end in enumerate(self.segments):
            if self.graph_bwd is not None:
                self.graph_bwd[k].replay()
            else:
                for op in self.bwd[start:end]:
                    op()
            done = None
            for bi, (lo, hi, ready) in enumerate(self.buckets):
                if sync and ready == end:
                    self._allreduce_bucket(lo, hi)
                if update is not None and self._update_points()[bi] == end:
                    if self.dev.type != 'cuda':  # host-logic tests (dry run): no streams, same order
                        update(lo, hi)
                        continue
                    side = self._side_stream()
                    if done is None:
                        done = torch.cuda.Event()
                        done.record(torch.cuda.current_stream(self.dev))
                    side.wait_event(done)
                    with torch.cuda.stream(side):
                        update(lo, hi)
            start = end
        if getattr(self, 'comm_stream', None) is not None:
            torch.cuda.current_stream(self.dev).wait_stream(self.comm_stream)
        if update is not None:
            self.arena.armed_update, self.arena.update_applied = None, True

    def _run_two_streams(self, op_list, side_flags, main, side):
        """Issue `op_list` on two streams: ops flagged `side` go to `side`, everything else to `main` (the current
        stream).  Dependencies are derived from the ops' read / write address ranges:
          * a side op starts after everything issued on main before it (event recorded on main),
          * a main op that touches data a still-pending side op writes (or overwrites data it reads) first waits for
            that side op's completion event,
          * at the end main joins the side stream.
        Under stream capture the events become graph edges, so the replayed graph runs both chains concurrently."""
        pending = []  # (read ranges, write ranges, completion event) of side ops main has not synchronised with yet
        main_dirty = True
        fork = None
        for op, is_side in zip(op_list, side_flags):
            reads, writes = _op_io(op)
            if is_side:
                if main_dirty or fork is None:
                    fork = torch.cuda.Event()
                    fork.record(main)
                    main_dirty = False
                side.wait_event(fork)
                with torch.cuda.stream(side):
                    op()
                    done = torch.cuda.Event()
                    done.record(side)
                pending.append((reads, writes, done))
            else:
                last = -1
                for i, (pr, pw, _) in enumerate(pending):
                    if _overlap(writes, pr) or _overlap(reads, pw):
                        last = i
                if last >= 0:  # the side stream is in order: waiting for entry `last` covers all earlier ones
                    main.wait_event(pending[last][2])
                    pending = pending[last + 1:]
                op()
                main_dirty = True
        if pending:
            main.wait_event(pending[-1][2])

    def capture_graphs(self):
        """Capture the two static schedules as CUDA graphs (call after at least one eager warm-up step).  Inside the
        graphs the off-critical-path ops (weight-gradient GEMMs, bias-gradient sums, time-embedding and context K/V
        projections) run on a second stream, concurrently with the main chain."""
        torch.cuda.synchronize(self.dev)
        s = torch.cuda.Stream(self.dev)
        s2 = torch.cuda.Stream(self.dev)
        s.wait_stream(torch.cuda.current_stream(self.dev))
        gf, gbs = torch.cuda.CUDAGraph(), []
        with torch.cuda.stream(s):
            with torch.cuda.graph(gf, stream=s):
                self._run_two_streams(self.fwd, self.fwd_side, s, s2)
            start = 0
            for end in self.segments:  # one graph per gradient-bucket segment (all-reduces are issued in between)
                gb = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gb, stream=s):
                    self._run_two_streams(self.bwd[start:end], self.bwd_side[start:end], s, s2)
                gbs.append(gb)
                start = end
        torch.cuda.current_stream(self.dev).wait_stream(s)
        torch.cuda.synchronize(self.dev)
        self.graph_fwd, self.graph_bwd = gf, gbs

    def capture_gemm_only(self, only=None):
        """A CUDA graph holding only the tensor-core GEMM launches of one step (forward + backward), replayed by
        bench.py to time the dominant kernel family on its own stream with CUDA events (roofline.achieved).
        only: restrict to ops with these names (e.g. ('attn_fwd', 'attn_bwd'))."""
        torch.cuda.synchronize(self.dev)
        s = torch.cuda.Stream(self.dev)
        s.wait_stream(torch.cuda.current_stream(self.dev))
        g = torch.cuda.CUDAGraph()
        n = 0
        with torch.cuda.stream(s):
            with torch.cuda.graph(g, stream=s):
                for op, is_gemm in list(zip(self.fwd, self.fwd_is_gemm)) + list(zip(self.bwd, self.bwd_is_gemm)):
                    if is_gemm and (only is None or op.func.__name__ in only):
                        op()
                        n += 1
        torch.cuda.current_stream(self.dev).wait_stream(s)
        torch.cuda.synchronize(self.dev)
        return g, n


# ==================================================================================================== autograd glue
class _UNetFn(torch.autograd.Function):
    """pred = UNet(sample, t, ctx) on the static schedule; backward replays the backward schedule and hands the
    parameter gradients (views of the g32 arena) to autograd."""

    @staticmethod
    def forward(ctx_, eng, prepared, sample, timestep, enc, *params):
        if not prepared:
            eng.prepare_inputs(sample, timestep, enc)
        eng.run_forward()
        pred = torch.empty(eng.B, 4, eng.H, eng.W, dtype=BF16, device=eng.dev)
        ops.nhwc8_to_nchw4(eng.ctx, eng.pred8, pred, eng.B, eng.H, eng.W)
        ctx_.eng = eng
        ctx_.generation = eng.generation
        ctx_.nparams = len(params)
        return pred

    @staticmethod
    def backward(ctx_, gpred):
        eng = ctx_.eng
        if eng.generation != ctx_.generation:
            raise RuntimeError('the UNet engine ran another forward on this input geometry after the forward this backward belongs '
                               'to: its saved activations are gone.  Call backward() before the next forward of the same shape '
                               '(no-grad / eval calls use a separate forward-only engine and are safe).')
        arena = eng.arena
        fused = getattr(eng, '_fused_loss_scale', None)
        if fused is not None and gpred.stride() == (0, 0, 0, 0):
            # gradient comes from the fused MSE head: dpred8 already holds 2(pred-noise)/N, scale by dL/dloss
            ops.scale_by_scalar(eng.ctx, eng.dpred8, fused)
            eng._fused_loss_scale = None
        else:
            ops.nchw4_to_nhwc8(eng.ctx, gpred.contiguous(), eng.dpred8, eng.B, eng.H, eng.W)
        plist = list(arena.params.items())
        accumulate = all(p.grad is not None and p.grad.data_ptr() == arena.g32.data_ptr() + 4 * arena.entries[n][0]
                         for n, p in plist)
        if not accumulate:
            arena.g32.zero_()
        # (the armed optimizer update clears the gradient arena as it goes: only when the gradients stay in place and no wrapper
        # reduces them afterwards)
        eng.run_backward(allow_update=accumulate and not getattr(eng, 'ddp_compat', False))
        if accumulate:  # gradients were accumulated in place into the arena that param.grad already views
            if getattr(eng, 'ddp_compat', False):
                # a DistributedDataParallel wrapper reduces `param.grad` from per-parameter autograd hooks, which only fire
                # if autograd receives a gradient: hand it zero-stride zeros (`grad += 0`), the real gradient is in place
                z = torch.zeros((), dtype=torch.float32, device=eng.dev)
                return (None, None, None, None, None) + tuple(z.expand(p.shape) for _, p in plist)
            return (None,) * (5 + ctx_.nparams)
        grads = tuple(arena.grad_view(n) for n, _ in plist)
        return (None, None, None, None, None) + grads


def unet_apply(unet, sample, timestep, enc, prepared=False):
    B, C, H, W = sample.shape
    if not torch.is_grad_enabled() and not prepared:
        # eval / no-grad call: a forward-only engine, so the activations a pending training backward needs stay intact
        eng = unet.engine(B, H, W, enc.shape[1], forward_only=True)
        eng.prepare_inputs(sample, timestep, enc)
        eng.run_forward()
        pred = torch.empty(B, 4, H, W, dtype=BF16, device=eng.dev)
        ops.nhwc8_to_nchw4(eng.ctx, eng.pred8, pred, B, H, W)
        return pred
    eng = unet.engine(B, H, W, enc.shape[1])
    params = [p for _, p in eng.arena.params.items()]
    return _UNetFn.apply(eng, prepared, sample, timestep, enc, *params)

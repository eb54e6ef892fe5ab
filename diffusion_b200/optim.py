"""Fused AdamW for the B200 UNet: same constructor and update rule as `torch.optim.AdamW`
(the reference instantiates it from yaml, `diffusion/train.py:33`, `yamls/hydra-yamls/SD-2-base-256.yaml:55-58`), but
one kernel launch per step over the flat fp32 parameter arena of the engine.  The same launch refreshes the bf16
shadow weights the tensor-core kernels read and clears the gradient arena for the next accumulation, so the separate
cast kernel, the memset and the 686-tensor multi-tensor-apply launches disappear from the step.

Parameters that do not live in an engine arena (or before the first forward has built it) are updated tensor by
tensor through the same kernel.
"""
import torch

from diffusion_b200 import ops


class FusedAdamW(torch.optim.Optimizer):

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, amsgrad=False,
                 keep_grads_bound=True, **_ignored):
        if amsgrad:
            raise ValueError('FusedAdamW: amsgrad is not implemented')
        if lr < 0 or eps < 0 or not 0 <= betas[0] < 1 or not 0 <= betas[1] < 1 or weight_decay < 0:
            raise ValueError('FusedAdamW: invalid hyper-parameter')
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay))
        self._grads_cleared = False
        # True (default): zero_grad() leaves `.grad` viewing the cleared gradient arena, the next backward accumulates in place
        # and hands autograd no gradient tensors.  A torch DistributedDataParallel wrapper (Composer's default for world > 1)
        # needs its per-parameter hooks to fire: pass keep_grads_bound=False there (or skip the wrapper and use
        # Engine.enable_grad_sync(), which overlaps the bucket all-reduces with backward itself).
        self.keep_grads_bound = keep_grads_bound

    def _arena_of(self, group):
        """The engine arena holding every parameter (and gradient) of this group, if there is one."""
        from diffusion_b200.engine import ParamArena
        ps = [p for p in group['params'] if p.requires_grad]
        if not ps:
            return None
        for arena in ParamArena.live():
            mine = arena.param_set()
            if all(id(p) in mine for p in ps) and len(ps) == len(mine) and arena.bound() and arena.grads_bound():
                return arena
        return None

    def _arena_state(self, gi, arena):
        # keyed by the parameter-group index: stable across processes, so `state_dict()` / `load_state_dict()`
        # (Composer checkpoints, reference train.py resume) carry the flat moments
        st = self.state.setdefault('arena_group%d' % gi, {})
        if not st:
            st['step'] = 0
            st['exp_avg'] = torch.zeros_like(arena.p32)
            st['exp_avg_sq'] = torch.zeros_like(arena.p32)
        if st['exp_avg'].numel() != arena.p32.numel():
            raise RuntimeError('FusedAdamW: the loaded optimizer state does not match this model\'s parameter arena')
        for k in ('exp_avg', 'exp_avg_sq'):  # a checkpoint loaded with map_location='cpu'
            if st[k].device != arena.p32.device or st[k].dtype != torch.float32:
                st[k] = st[k].to(device=arena.p32.device, dtype=torch.float32).contiguous()
        st['step'] = int(st['step'])
        return st

    @torch.no_grad()
    def arm(self):
        """Opt in, for ONE optimizer step: apply the update DURING the coming backward, gradient bucket by gradient bucket, as
        soon as a bucket's gradients are final (and averaged over the data-parallel group) and nothing later in backward reads
        its weights.  The HBM-bound update then overlaps the tensor-bound rest of backward and only the last bucket's
        all-reduce + update stay exposed; `step()` afterwards merely counts the step.  Same arithmetic as the one-launch path.
        Call it right before the LAST backward of the step (gradient accumulation: before the last microbatch), with the
        learning rate of this step already set, and only when nothing between backward and `step()` looks at the gradients
        (no clipping, no GradScaler inf check) - they are consumed and cleared bucket by bucket.
        Returns False (and changes nothing) when the parameters are not in a bound engine arena yet (first step)."""
        armed = False
        for gi, group in enumerate(self.param_groups):
            arena = self._arena_of(group)
            if arena is None:
                continue
            st = self._arena_state(gi, arena)
            lr, (b1, b2), eps, wd = group['lr'], group['betas'], group['eps'], group['weight_decay']
            step = st['step'] + 1
            m, v = st['exp_avg'], st['exp_avg_sq']

            def update(lo, hi, arena=arena, m=m, v=v, lr=lr, b1=b1, b2=b2, eps=eps, wd=wd, step=step):
                ops.adamw_step(ops.get_ctx(arena.p32.device), arena.p32[lo:hi], arena.g32[lo:hi], m[lo:hi], v[lo:hi],
                               arena.p16[lo:hi], lr, b1, b2, eps, wd, step, zero_grad=True)

            arena.armed_update, arena.update_applied = update, False
            armed = True
        return armed

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        cleared = True
        for gi, group in enumerate(self.param_groups):
            lr, (b1, b2), eps, wd = group['lr'], group['betas'], group['eps'], group['weight_decay']
            arena = self._arena_of(group)
            if arena is not None:
                st = self._arena_state(gi, arena)
                arena.armed_update = None  # armed, but that backward never ran: the one-launch path below
                if arena.update_applied:  # arm(): backward has already applied this step's update, bucket by bucket
                    arena.update_applied = False
                    st['step'] += 1
                    arena.mark_shadow_fresh()
                    continue
                st['step'] += 1
                ctx = ops.get_ctx(arena.p32.device)
                ops.adamw_step(ctx, arena.p32, arena.g32, st['exp_avg'], st['exp_avg_sq'], arena.p16, lr, b1, b2, eps, wd,
                               st['step'], zero_grad=True)
                arena.mark_shadow_fresh()
                continue
            cleared = False
            for p in group['params']:
                if p.grad is None:
                    continue
                if p.device.type != 'cuda' or p.dtype != torch.float32 or not p.is_contiguous() or not p.grad.is_contiguous():
                    raise RuntimeError('FusedAdamW needs contiguous fp32 CUDA parameters (no CPU fallback)')
                st = self.state[p]
                if not st:
                    st['step'] = 0
                    st['exp_avg'] = torch.zeros_like(p)
                    st['exp_avg_sq'] = torch.zeros_like(p)
                st['step'] += 1
                ops.adamw_step(ops.get_ctx(p.device), p, p.grad, st['exp_avg'], st['exp_avg_sq'], None, lr, b1, b2, eps, wd,
                               st['step'])
        self._grads_cleared = cleared
        return loss

    def zero_grad(self, set_to_none=True):
        """The arena path has already cleared the gradients inside step(); keep the `.grad` views bound so that the
        next backward accumulates in place (no memset, no 686 gradient hand-offs through autograd)."""
        if self._grads_cleared and self.keep_grads_bound:
            self._grads_cleared = False
            return
        self._grads_cleared = False
        super().zero_grad(set_to_none=set_to_none)

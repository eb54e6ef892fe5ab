"""2-GPU check of the two gradient-averaging routes: (a) Engine.enable_grad_sync(), (b) a torch
DistributedDataParallel wrapper around the model with torch.optim.AdamW / FusedAdamW(keep_grads_bound=False).  Both must
leave identical, rank-averaged gradients on every rank.  Run: torchrun --nproc-per-node 2 tools/ddp_wrapper_check.py"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    rank, local = int(os.environ['RANK']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    from diffusion_b200.model import stable_diffusion_2
    from diffusion_b200.optim import FusedAdamW
    from oracle.unet import TINY_UNET_CONFIG

    def make():
        torch.manual_seed(17)
        return stable_diffusion_2(pretrained=False, precomputed_latents=True, build_encoders=False, unet_config=TINY_UNET_CONFIG, fsdp=True)

    g = torch.Generator(device=dev).manual_seed(100 + rank)  # different data per rank
    batch = {'image_latents': torch.randn(2, 4, 16, 16, device=dev, generator=g).to(torch.bfloat16),
             'caption_latents': torch.randn(2, 77, 1024, device=dev, generator=g).to(torch.bfloat16)}

    def grads_after(model, wrapped, opt, steps=2):
        for s in range(steps):
            torch.manual_seed(1000 + s)  # same noise / timesteps on both ranks: only the data differs
            out = wrapped(batch)
            loss = model.loss(out, batch)
            loss.backward()
            if s < steps - 1:
                opt.step()
                opt.zero_grad(set_to_none=True)
        torch.cuda.synchronize()
        return torch.cat([p.grad.float().flatten() for p in model.unet.parameters()])

    # reference for ONE step: local gradients without any synchronisation, averaged explicitly
    os.environ['SD2_NO_AUTO_SYNC'] = '1'
    m_r = make()
    g_loc = grads_after(m_r, m_r, None, steps=1)
    del os.environ['SD2_NO_AUTO_SYNC']
    g_ref = g_loc.clone()
    dist.all_reduce(g_ref)
    g_ref /= dist.get_world_size()
    for name, mk in (('engine', 'a'), ('engine (automatic: torch.distributed world > 1, no wrapper, yaml unchanged)', 'auto'), ('ddp', 'b')):
        m1 = make()
        if mk == 'a':
            m1.unet.engine(2, 16, 16, 77).enable_grad_sync()
            g1 = grads_after(m1, m1, None, steps=1)
        elif mk == 'auto':
            assert m1.unet._fsdp_wrap is False
            g1 = grads_after(m1, m1, None, steps=1)
        else:
            g1 = grads_after(m1, torch.nn.parallel.DistributedDataParallel(m1, device_ids=[local]), None, steps=1)
        if rank == 1:
            print(f'one step, {name}: cosine vs explicit average %.6f, vs local %.6f' % (
                torch.nn.functional.cosine_similarity(g1, g_ref, dim=0).item(),
                torch.nn.functional.cosine_similarity(g1, g_loc, dim=0).item()))
    # (a) engine route
    m_a = make()
    m_a.unet.engine(2, 16, 16, 77).enable_grad_sync()
    ga = grads_after(m_a, m_a, FusedAdamW(m_a.parameters(), lr=1e-3))
    # (b) DDP wrapper, torch AdamW
    m_b = make()
    ddp_b = torch.nn.parallel.DistributedDataParallel(m_b, device_ids=[local])
    gb = grads_after(m_b, ddp_b, torch.optim.AdamW(m_b.parameters(), lr=1e-3))
    # (b') DDP wrapper, FusedAdamW with unbound gradients
    m_c = make()
    ddp_c = torch.nn.parallel.DistributedDataParallel(m_c, device_ids=[local])
    gc = grads_after(m_c, ddp_c, FusedAdamW(m_c.parameters(), lr=1e-3, keep_grads_bound=False))
    # (c) Composer-style microbatching under the wrapper: two microbatches per step, no_sync() on all but the last, FusedAdamW
    #     with bound gradients (in-place accumulation) -> needs unet.ddp_compat so the last microbatch fires the hooks
    def accumulate_two(model, wrapped, sync_ctx):
        model.zero_grad(set_to_none=True)
        for mb in range(2):
            torch.manual_seed(2000 + mb)
            b = {k: v.roll(mb, 0) for k, v in batch.items()}
            ctx = sync_ctx() if mb == 0 else contextlib.nullcontext()
            with ctx:
                (model.loss(wrapped(b), b) / 2).backward()
        torch.cuda.synchronize()
        return torch.cat([p.grad.float().flatten() for p in model.unet.parameters()])

    import contextlib
    os.environ['SD2_NO_AUTO_SYNC'] = '1'
    m_d = make()
    g_local = accumulate_two(m_d, m_d, contextlib.nullcontext)
    del os.environ['SD2_NO_AUTO_SYNC']
    g_want = g_local.clone()
    dist.all_reduce(g_want)
    g_want /= dist.get_world_size()
    m_e = make()
    m_e.unet.ddp_compat = True
    ddp_e = torch.nn.parallel.DistributedDataParallel(m_e, device_ids=[local])
    opt_e = FusedAdamW(m_e.parameters(), lr=0.0)  # lr 0: a first step only binds the gradients (zero_grad keeps them bound)
    (m_e.loss(ddp_e(batch), batch)).backward()
    opt_e.step()
    opt_e.zero_grad()
    assert m_e._last_engine.arena.grads_bound()
    # accumulate on the bound, cleared arena (zero_grad(set_to_none=True) would unbind the gradients)
    for mb in range(2):
        torch.manual_seed(2000 + mb)
        b = {k: v.roll(mb, 0) for k, v in batch.items()}
        ctx = ddp_e.no_sync() if mb == 0 else contextlib.nullcontext()
        with ctx:
            (m_e.loss(ddp_e(b), b) / 2).backward()
    torch.cuda.synchronize()
    g_ddp = torch.cat([p.grad.float().flatten() for p in m_e.unet.parameters()])
    cos_mb = torch.nn.functional.cosine_similarity(g_ddp, g_want, dim=0).item()
    # (d) engine route with the optimizer update applied per bucket inside backward (FusedAdamW.arm), against the one-launch step
    def train(armed, steps=3):
        m = make()
        m.unet.engine(2, 16, 16, 77).enable_grad_sync()
        opt = FusedAdamW(m.parameters(), lr=1e-3)
        for s in range(steps):
            torch.manual_seed(3000 + s)
            loss = m.loss(m(batch), batch)
            if armed:
                assert opt.arm() == (s > 0)
            loss.backward()
            opt.step()
            opt.zero_grad(set_to_none=True)
        torch.cuda.synchronize()
        return torch.cat([p.detach().float().flatten() for p in m.unet.parameters()])

    p_plain, p_armed = train(False), train(True)
    d_armed = (p_plain - p_armed).abs()
    res = {}
    for name, gx in (('engine', ga), ('ddp+AdamW', gb), ('ddp+FusedAdamW', gc), ('ddp+microbatches', g_ddp),
                     ('weights after 3 armed steps', p_armed)):
        other = gx.clone()
        dist.broadcast(other, src=0)
        res[name] = (gx - other).abs().max().item()  # ranks agree?
    cos_ab = torch.nn.functional.cosine_similarity(ga, gb, dim=0).item()
    cos_ac = torch.nn.functional.cosine_similarity(ga, gc, dim=0).item()
    if rank == 1:
        print('max |grad(rank1) - grad(rank0)| per route:', res)
        print('cosine engine vs ddp+AdamW: %.6f, engine vs ddp+FusedAdamW: %.6f' % (cos_ab, cos_ac))
        print('microbatch accumulation under DDP no_sync + ddp_compat: cosine vs explicit average %.6f' % cos_mb)
        print('AdamW per bucket inside backward vs one launch, 3 steps: max |dp| %.2e, mean |dp| %.2e' % (d_armed.max().item(), d_armed.mean().item()))
        ok = all(v < 1e-6 for v in res.values()) and cos_ab > 0.995 and cos_ac > 0.995 and cos_mb > 0.9999 \
            and d_armed.max().item() <= 6.1e-3 and d_armed.mean().item() < 2e-5
        print('DDP ROUTES OK' if ok else 'DDP ROUTES MISMATCH', flush=True)
        if not ok:
            os._exit(1)
    dist.destroy_process_group()


if __name__ == '__main__':
    main()

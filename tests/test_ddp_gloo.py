"""World-size-2 data-parallel gradient averaging (K5 host logic) on CPU over gloo: the bucketed, segment-by-segment
all-reduce of the flat gradient arena must leave rank-averaged gradients on both ranks."""
import os
import sys

import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), SD2_DRY_RUN='1')
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from diffusion_b200.unet import UNet2DConditionModel
    from oracle.unet import TINY_UNET_CONFIG
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.manual_seed(0)
    u = UNet2DConditionModel(**TINY_UNET_CONFIG)
    eng = u.engine(2, 32, 32, 77)
    eng.enable_grad_sync()
    g = torch.Generator().manual_seed(100 + rank)
    eng.arena.g32.copy_(torch.randn(eng.arena.total, generator=g))
    mine = eng.arena.g32.clone()
    eng.run_backward()  # dry kernels; the bucket all-reduces are real
    other = torch.randn(eng.arena.total, generator=torch.Generator().manual_seed(100 + (1 - rank)))
    ok = torch.allclose(eng.arena.g32, (mine + other) / 2, atol=1e-6)
    # no_sync: gradients stay local, synchronisation resumes afterwards
    with eng.no_sync():
        eng.arena.g32.copy_(mine)
        eng.run_backward()
        ok = ok and torch.equal(eng.arena.g32, mine)
    eng.run_backward()
    ok = ok and torch.allclose(eng.arena.g32, (mine + other) / 2, atol=1e-6)
    q.put((rank, bool(ok), len(eng.buckets)))
    dist.destroy_process_group()


def test_bucketed_allreduce_world2():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
    assert res[0][2] >= 4


def _ddp_wrapper_worker(rank, world, port, q):
    """A torch DistributedDataParallel wrapper records parameter strides at construction: the arena must already be bound
    (UNet2DConditionModel.bind_arena, called by the factory), otherwise the wrapper hands back corrupted gradients for the
    permuted-stride 3x3 weights.  Gradients are returned as arena views by a stand-in autograd function (no kernels)."""
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), SD2_DRY_RUN='1')
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from diffusion_b200.unet import UNet2DConditionModel
    from oracle.unet import TINY_UNET_CONFIG
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.manual_seed(0)
    u = UNet2DConditionModel(**TINY_UNET_CONFIG)
    arena = u.bind_arena()
    assert arena.bound() and u.bind_arena() is arena
    names = list(arena.params)
    ddp = torch.nn.parallel.DistributedDataParallel(u)

    class Fake(torch.autograd.Function):

        @staticmethod
        def forward(ctx, x, *params):
            return x.sum() + 0 * sum(p.flatten()[0] for p in params)

        @staticmethod
        def backward(ctx, g):
            gen = torch.Generator().manual_seed(100 + rank)
            arena.g32.copy_(torch.randn(arena.total, generator=gen))
            return (None,) + tuple(arena.grad_view(n) for n in names)

    u.forward = lambda x: Fake.apply(x, *[arena.params[n] for n in names])
    ddp(torch.ones(1)).backward()
    mine = torch.randn(arena.total, generator=torch.Generator().manual_seed(100 + rank))
    other = torch.randn(arena.total, generator=torch.Generator().manual_seed(100 + (1 - rank)))
    want = (mine + other) / 2
    ok = True
    for n in names:
        off, num, _ = arena.entries[n]
        got = arena.params[n].grad
        ref = arena._view(want, n)
        ok = ok and got.stride() == arena.params[n].stride() and torch.allclose(got, ref, atol=1e-6)
    q.put((rank, bool(ok), len(names)))
    dist.destroy_process_group()


def test_ddp_wrapper_sees_arena_backed_parameters_world2():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_ddp_wrapper_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
    assert res[0][2] == 686

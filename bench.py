"""Benchmark of the SD-2 UNet noise-prediction training step (BASELINE.json metric) on B200.

  python bench.py --gpus N --steps K --warmup W            # our arm (sm_100a kernels behind the C ABI)
  python bench.py --impl reference --gpus N --steps K ...  # reference arm: the torch oracle of the reference step
                                                           # on the box's host cores (the reference itself has no
                                                           # installable stack here: diffusers/composer are absent)
A step = StableDiffusion.forward (K1 + UNet) + loss + backward + gradient all-reduce (N>1) + AdamW step on one
microbatch of synthetic latents per GPU (SD-2-base-256: 32x32x4 latents, 77x1024 context, bf16, random init).
Per-GPU microbatch: the reference recipe trains with a device batch of 256 (global 2048 / 8 GPUs,
yamls/hydra-yamls/SD-2-base-256.yaml:2,87) cut into microbatches of 16 to fit 40/80 GB parts; microbatching does not
change the optimizer step (GroupNorm is per sample, the loss is a mean), so on 180 GB B200s the default here is the
whole device batch, 256, in one pass (--batch 16 reproduces the yaml's microbatch; measured on one B200: 630 img/s at
16, 983 at 64, 1156 at 128, 1199 at 256).  The optimizer runs every step.
One JSON line is printed by rank 0.  See DESIGN.md "Measurement" for the roofline arithmetic.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# algorithmic training FLOPs per image = 3 (fwd+dgrad+wgrad) * 2 * forward MACs
TFLOP_PER_IMAGE = {32: 3 * 2 * 90.55e9 / 1e12, 64: 3 * 2 * 402.13e9 / 1e12}
METRIC = 'sd2_unet_train_images_per_sec'
CPU_SAMPLE_BATCH = 2  # images per CPU step in BOTH CPU legs (--impl reference and cpu_baseline): one baseline, not two


def measured_traffic(batch, latent):
    """DRAM bytes per step of the tensor-core family from the committed ncu pass of this command
    (profiles/traffic.json, written by tools/ncu_traffic.py), or None if that configuration was not captured."""
    path = os.path.join(ROOT, 'profiles', 'traffic.json')
    if not os.path.exists(path):
        return None
    with open(path) as f:
        for row in json.load(f):
            if row.get('per_gpu_microbatch') == batch and row.get('latent') == latent:
                return row.get('tensor_family_dram_bytes_per_step')
    return None


def read_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return p.get('bf16_tflops_sustained', 1384.6), p.get('bf16_tflops', 1631.2), p.get('hbm_gbs', 6549.1), 'measured'
    return 1400.0, 1590.0, 6650.0, 'fallback'


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-lms', '100'],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[3:7]):
                    if v.lower().startswith('active'):
                        reasons.add(name)
            except Exception:
                pass
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': mx, 'reasons': sorted(reasons),
                'samples': len(sm)}


def oracle_cpu_step_rate(res, batch, steps, warmup, threads, device_batch):
    """The reference training step restated in torch (oracle/), fp32 on the host cores.  The timed sample is `steps` microbatches
    of `batch` images (fwd + loss + bwd) plus ONE torch.optim.AdamW step (reference yaml :55-58); the rate is that of a whole
    device batch run as microbatches of `batch` with one optimizer step, the schedule of the reference trainer:
    images/s = device_batch / (device_batch / batch * t_microbatch + t_adamw).  Returns (images/s, s per microbatch, s per AdamW)."""
    from oracle.stable_diffusion import StableDiffusionOracle, train_step
    from oracle.unet import SD2_BASE_UNET_CONFIG
    torch.set_num_threads(threads)
    torch.manual_seed(17)
    model = StableDiffusionOracle(SD2_BASE_UNET_CONFIG)
    opt = torch.optim.AdamW(model.unet.parameters(), lr=1.0e-4, weight_decay=0.01)
    b = {'image_latents': torch.randn(batch, 4, res, res), 'caption_latents': torch.randn(batch, 77, 1024)}
    for _ in range(warmup):
        train_step(model, b)
    opt.step()  # warm-up: allocates the moments
    model.zero_grad(set_to_none=True)
    t0 = time.perf_counter()
    for _ in range(steps):
        train_step(model, b)
    t_mb = (time.perf_counter() - t0) / steps
    t0 = time.perf_counter()
    opt.step()
    t_opt = time.perf_counter() - t0
    n_mb = max(1, device_batch // batch)
    return n_mb * batch / (n_mb * t_mb + t_opt), t_mb, t_opt


def gpu_library_step_rate(res, batch, steps, warmup, dev):
    """SURVEY.md 8(d) comparator: the oracle restatement of the reference step under bf16 autocast, eager torch on the SAME
    GPU (cuDNN implicit-GEMM convs, cuBLASLt linears, SDPA attention - what reference models.py:74-78,109-111 would dispatch
    to on this box) + torch's fused AdamW.  Informational: it never touches the product path."""
    from oracle.stable_diffusion import StableDiffusionOracle, train_step
    from oracle.unet import SD2_BASE_UNET_CONFIG
    torch.manual_seed(17)
    model = StableDiffusionOracle(SD2_BASE_UNET_CONFIG).to(dev)
    opt = torch.optim.AdamW(model.parameters(), lr=1.0e-4, weight_decay=0.01, fused=True)
    b = {'image_latents': torch.randn(batch, 4, res, res, device=dev).bfloat16(),
         'caption_latents': torch.randn(batch, 77, 1024, device=dev).bfloat16()}

    def one():
        train_step(model, b, autocast_dtype=torch.bfloat16)
        opt.step()
        opt.zero_grad(set_to_none=True)

    for _ in range(warmup):
        one()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        one()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    del model, opt
    torch.cuda.empty_cache()
    return batch / (ms * 1e-3), ms


def time_graph(g, reps=5):
    for _ in range(2):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    batch = CPU_SAMPLE_BATCH  # bounded sample of the microbatch (the same one the cpu_baseline leg of our arm times)
    val, s_per_step, s_opt = oracle_cpu_step_rate(args.latent, batch, args.steps, max(1, min(args.warmup, 1)), cores, args.batch)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': 'images/s', 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': s_per_step * 1e3, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'fp32', 'data': 'synthetic',
        'config': {'workload': f'SD-2-base-{args.latent * 8} UNet train step: K1 + UNet fwd + MSE + bwd + AdamW, precomputed latents, random '
                               'init; oracle restatement of the reference step on the host cores',
                   'per_gpu_microbatch': args.batch, 'latent': [4, args.latent, args.latent], 'context': [77, 1024], 'params': 865910724},
        'cpu_baseline': {'value': val, 'unit': 'images/s', 'cores': cores, 'kind': 'port',
                         'sample': f'{args.steps} timed microbatches of {batch} images ({s_per_step:.2f} s each) + one torch AdamW step '
                                   f'({s_opt:.2f} s), rate of a {args.batch}-image device batch run that way; fp32, torch {torch.__version__}'},
        'e2e': {'value': val, 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line), flush=True)


def measure_microbatch(model, opt, mb, device_batch, R, dev, steps, warmup, graphs, overlap=False):
    """The reference yaml's own schedule on one GPU: the device batch is cut into microbatches of `mb` images
    (SD-2-base-256.yaml:87 device_train_microbatch_size: 16; device batch 256 = global 2048 / 8 GPUs, :2), gradients accumulate
    over the microbatches (each loss scaled by its share, as Composer does) and the optimizer steps once per device batch.
    Own engine over the shared parameter arena.  Returns images/s over whole optimizer steps."""
    lat = torch.randn(mb, 4, R, R, device=dev).to(torch.bfloat16)
    ctx = torch.randn(mb, 77, 1024, device=dev).to(torch.bfloat16)
    batch = {'image_latents': lat, 'caption_latents': ctx}
    eng = model.unet.engine(mb, R, R, 77)
    n_mb = max(1, device_batch // mb)

    def opt_step():
        for i in range(n_mb):
            loss = model.loss(model(batch), batch) * (1.0 / n_mb)
            if overlap and i == n_mb - 1:
                opt.arm()  # the update rides on the last microbatch's backward
            loss.backward()
        opt.step()
        opt.zero_grad(set_to_none=True)

    opt_step()
    l0 = eng.ctx.launches
    opt_step()
    torch.cuda.synchronize()
    launches = eng.ctx.launches - l0
    if graphs:
        eng.capture_graphs()
    for _ in range(max(1, warmup // 2)):
        opt_step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        opt_step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {'per_gpu_microbatch': mb, 'microbatches_per_optimizer_step': n_mb, 'device_batch': n_mb * mb,
            'value': n_mb * mb / (ms * 1e-3), 'unit': 'images/s', 'ms_per_optimizer_step': ms, 'ms_per_microbatch': ms / n_mb,
            'gpu_launches_per_optimizer_step': launches,
            'note': 'reference yaml schedule: device_train_microbatch_size 16 (SD-2-base-256.yaml:87), gradient accumulation over '
                    'the device batch, one optimizer step per device batch'}


def run_ours(args):
    import torch.distributed as dist
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if not torch.cuda.is_available():
        raise RuntimeError('bench.py needs a B200: the product path has no CPU fallback')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)
    from diffusion_b200.model import stable_diffusion_2
    B, R = args.batch, args.latent
    torch.manual_seed(17)  # identical init on every rank (reference train.py:29) ...
    model = stable_diffusion_2(pretrained=False, precomputed_latents=not args.in_loop, fsdp=False)
    from diffusion_b200.optim import FusedAdamW
    opt = FusedAdamW(model.parameters(), lr=1.0e-4, weight_decay=0.01)  # reference yaml :55-58 (torch AdamW defaults)
    torch.manual_seed(17 + rank)  # ... then per-rank noise / timestep / data streams (composer reseeds seed + rank)
    if args.in_loop:  # BASELINE config 5: VAE encoder + CLIP text encoder inside the step (images ~U(-1,1), random token ids)
        lat_h = (torch.rand(B, 3, R * 8, R * 8) * 2 - 1).pin_memory()
        ctx_h = torch.randint(0, 49408, (B, 77)).pin_memory()
        lat_d, ctx_d = lat_h.to(dev), ctx_h.to(dev)
        batch = {'image': lat_d, 'captions': ctx_d}
    else:
        lat_h = torch.randn(B, 4, R, R).to(torch.bfloat16).pin_memory()
        ctx_h = torch.randn(B, 77, 1024).to(torch.bfloat16).pin_memory()
        lat_d, ctx_d = lat_h.to(dev), ctx_h.to(dev)
        batch = {'image_latents': lat_d, 'caption_latents': ctx_d}
    eng = model.unet.engine(B, R, R, 77)
    if world > 1:
        eng.enable_grad_sync()

    def step(e2e):
        if e2e:
            lat_d.copy_(lat_h, non_blocking=True)
            ctx_d.copy_(ctx_h, non_blocking=True)
        out = model(batch)
        loss = model.loss(out, batch)
        if args.optimizer_overlap:
            opt.arm()  # AdamW per gradient bucket inside backward (FusedAdamW.arm); opt.step() then only counts the step
        loss.backward()
        opt.step()
        opt.zero_grad(set_to_none=True)
        return loss.item() if e2e else loss

    # eager warm-up (also counts our kernel launches per step), then CUDA-graph capture of the static schedules
    step(False)  # the first step binds the gradient arena
    l0 = eng.ctx.launches
    step(False)
    torch.cuda.synchronize()
    launches_per_step = eng.ctx.launches - l0
    if not args.no_graphs:
        eng.capture_graphs()
        if args.in_loop:
            for m in (model.vae, model.text_encoder):
                for fe in m._engines.values():
                    fe.capture()
    for _ in range(args.warmup):
        step(False)

    if args.profile_step:
        # ncu --profile-from-start off: exactly one warmed-up step (graph replays) inside the profiler range, no timing
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        step(False)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        if world > 1:
            dist.destroy_process_group()
        return

    def timed(e2e):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        last = None
        for _ in range(args.steps):
            last = step(e2e)
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item(), float(last.detach()) if torch.is_tensor(last) else float(last)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_dev, loss_dev = timed(False)
    ms_e2e, loss_e2e = timed(True)
    clocks = sampler.stop() if rank == 0 else None

    # ---- roofline: the tensor-core GEMM family (gemm_tc_kernel) timed alone on its stream with CUDA events
    roof = None
    if rank == 0:
        sustained, burst, hbm, how = read_peaks()
        g, n_gemm = eng.capture_gemm_only()
        gemm_ms = time_graph(g)
        achieved = TFLOP_PER_IMAGE[R] * B / (gemm_ms * 1e-3)
        step_ms = ms_dev / args.steps
        step_tflops = TFLOP_PER_IMAGE[R] * B / (step_ms * 1e-3)
        # fused attention alone (attn_fwd / attn_bwd launches of the step): its own algorithmic FLOPs (4 N_q N_k d forward,
        # 10 N_q N_k d backward per head) over its own time
        ga, n_attn = eng.capture_gemm_only(only=('attn_fwd', 'attn_bwd'))
        attn_ms = time_graph(ga)
        attn_tf = eng.attn_flops / 1e12
        roof = {'bound': 'tensor', 'achieved': achieved, 'peak': sustained, 'unit': 'TFLOP/s', 'frac': achieved / sustained,
                'traffic': measured_traffic(B, R), 'kernel': 'tensor-core family of one step replayed alone: gemm_tc_kernel<BN,A_MN,B_MN> (every GEMM/conv) + attn_fwd/bwd_kernel',
                'launches_per_step': n_gemm, 'ms_per_step_in_kernel': gemm_ms,
                'share_of_step': gemm_ms / step_ms, 'peak_source': f'{how} bf16_tflops_sustained',
                'plan_tflop_per_step': eng.gemm_flops / 1e12,
                'executed': {'achieved': eng.gemm_flops / 1e12 / (gemm_ms * 1e-3),
                             'frac': eng.gemm_flops / 1e12 / (gemm_ms * 1e-3) / sustained,
                             'note': 'FLOPs the family actually executes (plan_tflop_per_step, attention included): below the algorithmic count '
                                     'because Upsample2D runs as four 4-tap phase convolutions (16 instead of 36 tap-products per pixel)'},
                'frac_of_burst_peak': achieved / burst,
                'step': {'achieved': step_tflops, 'frac': step_tflops / sustained,
                         'note': 'algorithmic TFLOP of the step / whole step time (every kernel, optimizer included)'},
                'attention': {'achieved': attn_tf / (attn_ms * 1e-3), 'frac': attn_tf / (attn_ms * 1e-3) / sustained,
                              'ms_per_step_in_kernel': attn_ms, 'launches_per_step': n_attn, 'tflop_per_step': attn_tf,
                              'kernel': 'attn_fwd_kernel + attn_bwd_kernel (+ its pre / cast passes) replayed alone'}}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # ---- CPU baseline (oracle port) on this box's host cores, N=1 only
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        n_cpu = 8 if R <= 32 else 2  # ~10-20 s of host work
        v, s_per, s_opt = oracle_cpu_step_rate(R, CPU_SAMPLE_BATCH, n_cpu, 1, cores, B)
        cpu = {'value': v, 'unit': 'images/s', 'cores': cores, 'kind': 'port',
               'sample': f'{n_cpu} timed microbatches of {CPU_SAMPLE_BATCH} images ({s_per:.2f} s each) + one torch AdamW step ({s_opt:.2f} s), '
                         f'rate of the {B}-image device batch run that way; fp32 oracle after 1 warm-up'}
    # ---- the yaml's own microbatch (SD-2-base-256.yaml:87 device_train_microbatch_size: 16): same step, N=1 only
    yaml_mb = None
    if world == 1 and not args.in_loop and B != 16 and not args.no_secondary:
        yaml_mb = measure_microbatch(model, opt, 16, B, R, dev, max(2, args.steps // 4), args.warmup, not args.no_graphs,
                                     overlap=bool(args.optimizer_overlap))
    # ---- same-GPU library comparator (eager torch bf16 autocast of the oracle), N=1 only, after our arm is done
    lib = None
    if world == 1 and not args.in_loop and not args.no_secondary:
        del eng
        model.unet.drop_engines()
        torch.cuda.empty_cache()
        lib = {'unit': 'images/s', 'impl': 'oracle restatement of the reference step, torch eager, bf16 autocast, cuDNN / cuBLASLt / '
                                           'SDPA kernels + torch fused AdamW, same GPU',
               'torch': torch.__version__}
        for lb in ((16, 64) if R <= 32 else (4, 16)):
            try:
                v, ms = gpu_library_step_rate(R, lb, 5, 2, dev)
                lib[f'batch_{lb}'] = {'value': v, 'ms_per_step': ms}
            except Exception as e:  # informational leg: never fail the bench line
                lib[f'batch_{lb}'] = {'error': f'{type(e).__name__}: {str(e)[:120]}'}
                torch.cuda.empty_cache()
    imgs = B * world * args.steps
    line = {
        'metric': METRIC, 'value': imgs / (ms_dev * 1e-3), 'unit': 'images/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms_dev / args.steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'bf16', 'data': 'synthetic',
        'config': {'workload': f'SD-2-base-{R * 8} UNet train step: K1 + UNet fwd + MSE + bwd + grad all-reduce + AdamW, '
                               + ('in-loop VAE encoder + CLIP text encoder (precomputed_latents=false), ' if args.in_loop else 'precomputed latents, ')
                               + 'random init', 'per_gpu_microbatch': B, 'global_batch': B * world,
                   'latent': [4, R, R], 'context': [77, 1024], 'params': 865910724, 'parallelism': f'dp{world}',
                   'optimizer': 'fused AdamW, ' + ('applied per gradient bucket during backward (FusedAdamW.arm)' if args.optimizer_overlap
                                                   else 'one launch in optimizer.step()'),
                   'cuda_graphs': not args.no_graphs,
                   'l2': 'no explicit flush: every step streams 3.5 GB of fp32 parameters + optimizer state and tens of GB of '
                         'activations, far beyond the 126 MB L2'},
        'e2e': {'value': imgs / (ms_e2e * 1e-3), 'unit': 'images/s', 'ms_per_step': ms_e2e / args.steps,
                'h2d_bytes_per_step': lat_h.numel() * lat_h.element_size() + ctx_h.numel() * ctx_h.element_size(),
                'd2h_bytes_per_step': 4},
        'gpu_launches': launches_per_step * args.steps, 'gpu_launches_per_step': launches_per_step,
        'clocks': clocks, 'roofline': roof, 'cpu_baseline': cpu, 'loss': loss_e2e,
        'yaml_microbatch': yaml_mb, 'gpu_library_baseline': lib,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--latent', type=int, default=32, help='latent side: 32 = SD-2-base-256 (default), 64 = SD-2-base-512')
    ap.add_argument('--batch', type=int, default=None,
                    help='per-GPU microbatch (default 256 at 256^2, 64 at 512^2; the reference yaml uses 16 on 40/80 GB GPUs)')
    ap.add_argument('--in-loop', action='store_true',
                    help='BASELINE config 5: precomputed_latents=false, VAE encoder + CLIP text encoder run inside the step')
    ap.add_argument('--no-graphs', action='store_true')
    ap.add_argument('--optimizer-overlap', type=int, default=0, choices=[0, 1],
                    help='1: FusedAdamW.arm() - the update runs bucket by bucket during backward; 0 (default): one launch in opt.step(). '
                         'Measured neutral on a power-capped B200 (DESIGN.md section 4)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-secondary', action='store_true',
                    help='skip the informational legs (yaml microbatch 16 line, same-GPU torch library baseline)')
    ap.add_argument('--profile-step', action='store_true',
                    help='run one warmed-up step inside cudaProfilerStart/Stop and exit (for ncu --profile-from-start off)')
    args = ap.parse_args()
    if args.batch is None:
        args.batch = 256 if args.latent <= 32 else 64
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()

"""Host-side mirror of the reference model API for the hot path (same names, arguments and error behaviour):

  * `StableDiffusion`           <- reference diffusion/models/stable_diffusion.py:15  (forward :154-183, loss :185-187,
                                   eval_forward :189-208 early-return, get_metrics :210-226, update_metric :228-257)
  * `stable_diffusion_2(...)`   <- reference diffusion/models/models.py:28-112
  * `DDPMScheduler`             <- diffusers DDPMScheduler as configured by SD-2-base

forward() on precomputed latents runs K1 (timesteps + noise + add_noise + timestep embedding, bit-exact with the torch
CUDA RNG stream) and the UNet engine; loss() uses the fused MSE head.  Everything executes in the sm_100a kernels
behind the C ABI; there is no CPU or eager-PyTorch fallback (a CPU model raises).
"""
from typing import List, Optional

import torch
import torch.nn.functional as F

from diffusion_b200 import ops
from diffusion_b200.unet import SD2_BASE_UNET_CONFIG, UNet2DConditionModel

try:  # the reference's base class / metric, when their packages exist
    from composer.models import ComposerModel  # type: ignore
except Exception:  # pragma: no cover - composer is absent in this image
    ComposerModel = torch.nn.Module
try:
    from torchmetrics import MeanSquaredError, Metric  # type: ignore
except Exception:  # pragma: no cover - torchmetrics is absent in this image

    class Metric(torch.nn.Module):
        pass

    class MeanSquaredError(Metric):
        """Minimal stand-in with torchmetrics' update/compute/reset contract (sum_squared_error / total)."""

        def __init__(self, **_kw):
            super().__init__()
            self.register_buffer('sum_squared_error', torch.zeros((), dtype=torch.float32))
            self.register_buffer('total', torch.zeros((), dtype=torch.float32))

        def update(self, preds, target):
            d = preds.float() - target.float()
            self.sum_squared_error = self.sum_squared_error.to(d.device) + (d * d).sum()
            self.total = self.total.to(d.device) + d.numel()

        def compute(self):
            return self.sum_squared_error / self.total

        def reset(self):
            self.sum_squared_error.zero_()
            self.total.zero_()


class DDPMScheduler:
    """SD-2-base training scheduler: scaled_linear betas, 1000 steps, epsilon prediction."""

    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule='scaled_linear',
                 prediction_type='epsilon', **_ignored):
        if beta_schedule != 'scaled_linear':
            raise ValueError('only the scaled_linear schedule of SD-2-base is implemented')
        self.num_train_timesteps = num_train_timesteps
        self.prediction_type = prediction_type
        self.betas = torch.linspace(beta_start**0.5, beta_end**0.5, num_train_timesteps, dtype=torch.float32)**2
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self._ac_dev = {}

    def __len__(self):
        return self.num_train_timesteps

    def alphas_cumprod_on(self, device):
        if device not in self._ac_dev:
            self._ac_dev[device] = self.alphas_cumprod.to(device).contiguous()
        return self._ac_dev[device]

    def add_noise(self, original_samples, noise, timesteps):
        """Stand-alone add_noise (same rounding as diffusers: alphas_cumprod cast to the sample dtype first)."""
        ac = self.alphas_cumprod.to(device=original_samples.device, dtype=original_samples.dtype)
        a = (ac[timesteps]**0.5).flatten()
        s = ((1 - ac[timesteps])**0.5).flatten()
        while a.dim() < original_samples.dim():
            a, s = a.unsqueeze(-1), s.unsqueeze(-1)
        return a * original_samples + s * noise


class _StepOutput(dict):
    """Dict with attribute access, like diffusers' BaseOutput (`scheduler.step(...).prev_sample`, `unet(...).sample`)."""
    __getattr__ = dict.__getitem__


class DDIMScheduler:
    """Inference scheduler of `generate()`: diffusers' DDIMScheduler as `from_pretrained(SD-2-base scheduler config)`
    builds it at reference models.py:89 - scaled_linear betas, epsilon prediction, clip_sample False,
    set_alpha_to_one False, steps_offset 1, 'leading' timestep spacing, eta 0.  Scalars stay fp32 CPU tensors like the
    original, so `step()` on torch tensors promotes types exactly as the reference loop does; `generate()` itself runs
    the fused `sd2_cfg_ddim_step` kernel with the same scalars."""
    init_noise_sigma = 1.0

    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule='scaled_linear',
                 prediction_type='epsilon', clip_sample=False, set_alpha_to_one=False, steps_offset=1, **_ignored):
        if beta_schedule != 'scaled_linear' or prediction_type != 'epsilon' or clip_sample:
            raise ValueError('only the SD-2-base inference scheduler configuration is implemented')
        self.num_train_timesteps, self.steps_offset = num_train_timesteps, steps_offset
        self.betas = torch.linspace(beta_start**0.5, beta_end**0.5, num_train_timesteps, dtype=torch.float32)**2
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.num_inference_steps = None
        self.timesteps = torch.arange(num_train_timesteps - 1, -1, -1, dtype=torch.int64)

    def __len__(self):
        return self.num_train_timesteps

    def set_timesteps(self, num_inference_steps, device=None):
        if num_inference_steps > self.num_train_timesteps:
            raise ValueError(f'`num_inference_steps`: {num_inference_steps} cannot be larger than '
                             f'`num_train_timesteps`: {self.num_train_timesteps}')
        self.num_inference_steps = num_inference_steps
        ratio = self.num_train_timesteps // num_inference_steps
        ts = (torch.arange(0, num_inference_steps, dtype=torch.int64) * ratio).flip(0) + self.steps_offset
        self.timesteps = ts.to(device) if device is not None else ts

    def scale_model_input(self, sample, timestep=None):
        return sample

    def step_scalars(self, timestep):
        """fp32 (sqrt(1 - a_t), sqrt(a_t), sqrt(a_prev), sqrt(1 - a_prev)) of one eta = 0 step."""
        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None', you need to run 'set_timesteps' after creating the scheduler")
        t = int(timestep)
        prev = t - self.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev] if prev >= 0 else self.final_alpha_cumprod
        return (1 - a_t)**0.5, a_t**0.5, a_prev**0.5, (1 - a_prev)**0.5

    def step(self, model_output, timestep, sample, eta=0.0, generator=None, **_unused):
        if eta != 0.0:
            raise ValueError('only eta = 0 (deterministic DDIM) is implemented')
        sb, sa, sap, dc = self.step_scalars(timestep)
        pred_original_sample = (sample - sb * model_output) / sa
        prev_sample = sap * pred_original_sample + dc * model_output
        return _StepOutput(prev_sample=prev_sample, pred_original_sample=pred_original_sample)


class _FusedMSE(torch.autograd.Function):
    """loss = mean((pred - noise)^2).  Training forwards produce the sum of squares and dL/dpred in the epilogue of conv_out
    itself (the target noise sits in the engine's static buffer, csrc/gemm_tc.cu MSE head): this node then only divides;
    otherwise the stand-alone head kernel computes both from the engine's pred8 buffer."""

    @staticmethod
    def forward(ctx_, pred, noise, eng):
        in_epilogue = eng.mse_generation == eng.generation and noise is eng.noise_target
        if not in_epilogue:  # the conv_out epilogue did not see this target (other dtype / tensor): stand-alone head kernel
            ops.fill_f32(eng.ctx, eng.loss_acc, 0.0)
            ops.mse_head(eng.ctx, eng.pred8, noise.contiguous(), None, eng.dpred8, eng.loss_acc, 1.0, eng.B, eng.H, eng.W)
        ctx_.eng = eng
        ctx_.shape = pred.shape
        return (eng.loss_acc[0] / eng.loss_acc[1]).to(torch.float32)

    @staticmethod
    def backward(ctx_, g):
        eng = ctx_.eng
        eng._fused_loss_scale = g.detach().to(torch.float32).reshape(1).contiguous()
        return g.to(torch.bfloat16).expand(ctx_.shape), None, None


_IMAGE_METRICS = ('FrechetInceptionDistance', 'InceptionScore', 'CLIPScore')


def _clone_metric(metric):
    """A fresh metric of the same class and constructor state.  The reference does `type(m)(**vars(m))` (:117,:125), which
    relies on torchmetrics accepting its own attribute dict; fall back to a plain re-construction / deep copy."""
    import copy
    try:
        return type(metric)(**{k: v for k, v in vars(metric).items() if not k.startswith('_')})
    except Exception:
        try:
            return copy.deepcopy(metric)
        except Exception:
            return type(metric)()


class StableDiffusion(ComposerModel):
    """Stable Diffusion ComposerModel on the B200-native UNet (see module docstring for the reference mapping)."""

    def __init__(self,
                 unet,
                 vae,
                 text_encoder,
                 tokenizer,
                 noise_scheduler,
                 inference_noise_scheduler,
                 loss_fn=F.mse_loss,
                 train_metrics: Optional[List] = None,
                 val_metrics: Optional[List] = None,
                 val_seed: int = 1138,
                 val_guidance_scales: Optional[List] = None,
                 loss_bins: Optional[List] = None,
                 image_key: str = 'image',
                 text_key: str = 'captions',
                 image_latents_key: str = 'image_latents',
                 text_latents_key: str = 'caption_latents',
                 precomputed_latents: bool = False,
                 encode_latents_in_fp16: bool = False,
                 fsdp: bool = False):
        super().__init__()
        self.unet = unet
        self.vae = vae
        self.noise_scheduler = noise_scheduler
        self.loss_fn = loss_fn
        self.val_seed = val_seed
        self.image_key = image_key
        self.image_latents_key = image_latents_key
        self.precomputed_latents = precomputed_latents
        self.train_metrics = [MeanSquaredError()] if train_metrics is None else train_metrics
        if val_metrics is None:
            val_metrics = [MeanSquaredError()]
        if val_guidance_scales is None:
            val_guidance_scales = [0.0]
        if loss_bins is None:
            loss_bins = [(0, 1)]
        self.val_guidance_scales = val_guidance_scales
        self.val_metrics = {}
        for metric in val_metrics:
            kind = metric.__class__.__name__
            if kind in _IMAGE_METRICS:  # one copy per guidance scale, fed with the images generated at that scale (ref :114-122)
                for scale in val_guidance_scales:
                    new_metric = _clone_metric(metric)
                    new_metric.guidance_scale = scale
                    self.val_metrics[f'{kind}-scale-' + str(scale).replace('.', 'p')] = new_metric
            elif isinstance(metric, MeanSquaredError):  # one copy per timestep bin (ref :123-129)
                for bin_ in loss_bins:
                    new_metric = _clone_metric(metric)
                    new_metric.loss_bin = bin_
                    self.val_metrics[f'{kind}-bin-{bin_[0]}-to-{bin_[1]}'.replace('.', 'p')] = new_metric
            else:
                self.val_metrics[kind] = metric
        self.val_metrics['MeanSquaredError'] = MeanSquaredError()
        self.text_encoder = text_encoder
        self.tokenizer = tokenizer
        self.inference_scheduler = inference_noise_scheduler
        self.text_key = text_key
        self.text_latents_key = text_latents_key
        self.encode_latents_in_fp16 = encode_latents_in_fp16
        for frozen in (self.text_encoder, self.vae):
            if frozen is not None:
                frozen.requires_grad_(False)
                if self.encode_latents_in_fp16 and not getattr(frozen, '_sd2_native', False):
                    frozen.half()  # the native encoders keep fp32 masters and compute from their own bf16 shadow
        if fsdp:
            # reference :90-97 marks the UNet for FSDP wrapping (SHARD_GRAD_OP).  Here the parameters are replicated (15.6 GB of
            # state on a 180 GB part) and live in the engine's flat arenas, which an FSDP FlatParameter would replace - so every
            # sub-module is marked "do not wrap" (Composer's prepare_fsdp_module skips children with _fsdp_wrap == False) and
            # the engine averages the gradients itself whenever torch.distributed has more than one rank.
            for m in (self.text_encoder, self.vae, self.unet):
                if m is not None:
                    m._fsdp_wrap = False
        self._last_engine = None

    # -- reference stable_diffusion.py:154-183 -----------------------------------------------------------------
    def forward(self, batch):
        if self.precomputed_latents and self.image_latents_key in batch and self.text_latents_key in batch:
            latents, conditioning = batch[self.image_latents_key], batch[self.text_latents_key]
        else:
            if self.vae is None or self.text_encoder is None:
                raise ValueError('batch has no precomputed latents and the model was built without VAE / text encoder '
                                 '(in-loop encoding is SURVEY.md row f1, not part of this hot path)')
            inputs, conditioning = batch[self.image_key], batch[self.text_key]
            conditioning = conditioning.view(-1, conditioning.shape[-1])
            if self.encode_latents_in_fp16:
                inputs = inputs.half()
            with torch.no_grad():
                latents = self.vae.encode(inputs)['latent_dist'].sample().data
                conditioning = self.text_encoder(conditioning)[0]
            latents = latents * 0.18215
        if latents.device.type != 'cuda':
            raise RuntimeError('diffusion_b200 needs CUDA (sm_100a) tensors: there is no CPU fallback')
        latents = latents.contiguous()
        B, _, H, W = latents.shape
        eng = self.unet.engine(B, H, W, conditioning.shape[1])
        self._last_engine = eng
        # K1: timesteps (randint) -> noise (randn_like) -> add_noise -> sinusoidal embedding, one launch, consuming the
        # torch CUDA generator exactly like the reference's two torch calls
        gen = torch.cuda.default_generators[latents.device.index]
        seed, offset = gen.initial_seed(), gen.get_offset()
        ac = self.noise_scheduler.alphas_cumprod_on(latents.device)
        target = eng.fused_mse_target(latents) if (torch.is_grad_enabled() and self.loss_fn is F.mse_loss) else None
        timesteps, noise, _, _, _, used = ops.noise_sched_fwd(eng.ctx, latents, ac, seed, offset, eng.in_temb.shape[1],
                                                              out_nhwc8=eng.in_x8, out_temb=eng.in_temb, out_noise=target)
        gen.set_offset(offset + used)
        eng.set_context(conditioning)
        from diffusion_b200.engine import unet_apply
        pred = unet_apply(self.unet, latents, timesteps, conditioning, prepared=True)
        if target is not None:
            eng.mse_generation = eng.generation  # conv_out's epilogue compared this forward's prediction with `noise`
        return pred, noise, timesteps

    # -- reference stable_diffusion.py:185-187 -----------------------------------------------------------------
    def loss(self, outputs, batch):
        pred, target = outputs[0], outputs[1]
        eng = self._last_engine
        if self.loss_fn is F.mse_loss and eng is not None and pred.grad_fn is not None and \
                getattr(pred.grad_fn, 'eng', None) is eng:
            return _FusedMSE.apply(pred, target, eng)
        return self.loss_fn(pred.float(), target.float())

    # -- reference stable_diffusion.py:189-208 (training path: early return) -----------------------------------
    def eval_forward(self, batch, outputs=None):
        if outputs is not None:
            return outputs
        unet_out, noise, timesteps = self.forward(batch)
        generated_images = {}
        if self.text_encoder is not None and self.vae is not None and self.text_key in batch and self.image_key in batch:
            prompts = batch[self.text_key]
            height, width = batch[self.image_key].shape[-2], batch[self.image_key].shape[-1]
            for guidance_scale in self.val_guidance_scales:
                generated_images[guidance_scale] = self.generate(tokenized_prompts=prompts, height=height, width=width,
                                                                 guidance_scale=guidance_scale, seed=self.val_seed,
                                                                 progress_bar=False)
        return unet_out, noise, timesteps, generated_images

    def get_metrics(self, is_train: bool = False):
        metrics = self.train_metrics if is_train else self.val_metrics
        if isinstance(metrics, Metric):
            return {metrics.__class__.__name__: metrics}
        if isinstance(metrics, list):
            return {m.__class__.__name__: m for m in metrics}
        for name, metric in metrics.items():
            if not isinstance(metric, Metric):
                raise TypeError(f'val metric {name!r} is a {type(metric).__name__}, not a torchmetrics Metric')
        return dict(metrics)

    def update_metric(self, batch, outputs, metric):
        """reference :228-257.  MSE metrics see (prediction, noise), per timestep bin when the metric carries one; image
        metrics see the images `eval_forward` generated at the metric's guidance scale."""
        kind = metric.__class__.__name__
        if isinstance(metric, MeanSquaredError) and hasattr(metric, 'loss_bin'):
            lo, hi = metric.loss_bin
            T_max = self.noise_scheduler.num_train_timesteps
            idx = torch.where((outputs[2] >= lo * T_max) & (outputs[2] < hi * T_max))
            metric.update(outputs[0][idx], outputs[1][idx])
        elif isinstance(metric, MeanSquaredError):
            metric.update(outputs[0], outputs[1])
        elif kind in _IMAGE_METRICS:
            scale = getattr(metric, 'guidance_scale', None)
            if len(outputs) < 4 or scale not in outputs[3]:
                raise ValueError(f'{kind} needs the images eval_forward generates at guidance scale {scale}: the batch must carry '
                                 f'{self.image_key!r} and {self.text_key!r} and the model its VAE and text encoder')
            images = outputs[3][scale]
            if kind == 'FrechetInceptionDistance':
                metric.update(batch[self.image_key], real=True)
                metric.update(images, real=False)
            elif kind == 'InceptionScore':
                metric.update(images)
            else:  # CLIPScore scores the images against the caption strings
                if self.tokenizer is None:
                    raise ValueError('CLIPScore needs the tokenizer to turn the token ids back into captions; this model was '
                                     'built without one (no vocabulary files on disk)')
                captions = [self.tokenizer.decode(c, skip_special_tokens=True) for c in batch[self.text_key]]
                metric.update((images * 255).to(torch.uint8), captions)
        else:
            metric.update(outputs[0], outputs[1])

    # -- reference stable_diffusion.py:259-382 -----------------------------------------------------------------
    @torch.no_grad()
    def generate(self,
                 prompt: Optional[list] = None,
                 negative_prompt: Optional[list] = None,
                 tokenized_prompts: Optional[torch.Tensor] = None,
                 tokenized_negative_prompts: Optional[torch.Tensor] = None,
                 prompt_embeds: Optional[torch.Tensor] = None,
                 negative_prompt_embeds: Optional[torch.Tensor] = None,
                 height: Optional[int] = None,
                 width: Optional[int] = None,
                 num_inference_steps: Optional[int] = 50,
                 guidance_scale: Optional[float] = 3.0,
                 num_images_per_prompt: Optional[int] = 1,
                 seed: Optional[int] = None,
                 progress_bar: Optional[bool] = True,
                 output_type: str = 'image'):
        """Backward diffusion from noise: same arguments and defaults as the reference plus `output_type` ('image' decodes
        with `self.vae` like the reference; 'latent' returns the final fp32 latents, which is all a model built without a
        VAE can return).  Per step: one timestep-embedding launch, the UNet forward graph on the (2x) batch, and ONE fused
        launch for classifier-free guidance + the DDIM update + the next UNet input (`sd2_cfg_ddim_step`)."""
        _require_some_prompt(prompt, tokenized_prompts, prompt_embeds)
        _require_matching_negatives(prompt, negative_prompt)
        _require_matching_negatives(tokenized_prompts, tokenized_negative_prompts)
        _require_matching_negatives(prompt_embeds, negative_prompt_embeds)
        if output_type not in ('image', 'latent'):
            raise ValueError("output_type must be 'image' or 'latent'")
        if output_type == 'image' and self.vae is None:
            raise ValueError("this model was built without a VAE: call generate(..., output_type='latent')")
        if self.inference_scheduler is None:
            raise ValueError('the model has no inference scheduler')
        device = next(self.unet.parameters()).device
        vae_scale_factor = 8
        sample_size = self.unet.config.get('sample_size', 64)
        height = height or sample_size * vae_scale_factor
        width = width or sample_size * vae_scale_factor
        do_cfg = guidance_scale > 1.0
        text_embeddings = self._prepare_text_embeddings(prompt, tokenized_prompts, prompt_embeds, num_images_per_prompt)
        batch_size = len(text_embeddings)
        if do_cfg:
            if negative_prompt is None and tokenized_negative_prompts is None and negative_prompt_embeds is None:
                negative_prompt = [''] * (batch_size // num_images_per_prompt)
            uncond = self._prepare_text_embeddings(negative_prompt, tokenized_negative_prompts, negative_prompt_embeds,
                                                   num_images_per_prompt)
            text_embeddings = torch.cat([uncond.to(text_embeddings.dtype), text_embeddings])
        if device.type != 'cuda':
            raise RuntimeError('diffusion_b200 needs CUDA (sm_100a) tensors: there is no CPU fallback')
        rng_generator = torch.Generator(device=device)
        if seed:
            rng_generator = rng_generator.manual_seed(seed)
        h, w = height // vae_scale_factor, width // vae_scale_factor
        latents = torch.randn((batch_size, self.unet.config['in_channels'], h, w), device=device, generator=rng_generator)
        self.inference_scheduler.set_timesteps(num_inference_steps)
        latents = (latents * self.inference_scheduler.init_noise_sigma).contiguous()
        nb = 2 if do_cfg else 1
        eng = self.unet.engine(nb * batch_size, h, w, text_embeddings.shape[1], forward_only=True)
        eng.set_context(text_embeddings.to(device))
        for half in range(nb):  # UNet input of the first step: bf16 NHWC8 copy of the initial latents, per CFG half
            ops.nchw4_to_nhwc8(eng.ctx, latents, eng.in_x8[half * batch_size * h * w:(half + 1) * batch_size * h * w],
                               batch_size, h, w)
        steps = self.inference_scheduler.timesteps
        if progress_bar:
            try:
                from tqdm.auto import tqdm
                steps = tqdm(steps)
            except Exception:  # pragma: no cover
                pass
        tvec = torch.empty(nb * batch_size, dtype=torch.int64, device=device)
        for t in steps:
            tvec.fill_(int(t))
            ops.timestep_embedding(eng.ctx, tvec, eng.in_temb, torch.float32)
            eng.run_forward()
            sb, sa, sap, dc = (float(x) for x in self.inference_scheduler.step_scalars(t))
            ops.cfg_ddim_step(eng.ctx, eng.pred8, latents, eng.in_x8, batch_size, h, w, do_cfg, guidance_scale, sb, sa, sap, dc)
            if eng.graph_fwd is None and len(steps) > 2:
                eng.capture_graphs()  # the remaining steps replay the forward schedule as one CUDA graph
        if output_type == 'latent':
            return latents.detach()
        latents = 1 / 0.18215 * latents
        image = self.vae.decode(latents).sample
        image = (image / 2 + 0.5).clamp(0, 1)
        return image.detach()

    def _prepare_text_embeddings(self, prompt, tokenized_prompts, prompt_embeds, num_images_per_prompt):
        """Tokenizes and embeds prompts if needed, then duplicates embeddings per generated image
        (reference stable_diffusion.py:384-405)."""
        if prompt_embeds is None:
            if self.text_encoder is None:
                raise ValueError('this model was built without a text encoder: pass prompt_embeds / negative_prompt_embeds')
            device = next(self.text_encoder.parameters()).device
            if tokenized_prompts is None:
                if self.tokenizer is None:
                    raise ValueError('this model was built without a tokenizer: pass tokenized_prompts or prompt_embeds')
                tokenized_prompts = self.tokenizer(prompt, padding='max_length', max_length=self.tokenizer.model_max_length,
                                                   truncation=True, return_tensors='pt').input_ids
            text_embeddings = self.text_encoder(tokenized_prompts.to(device))[0]
        else:
            text_embeddings = prompt_embeds
        bs_embed, seq_len, _ = text_embeddings.shape
        text_embeddings = text_embeddings.repeat(1, num_images_per_prompt, 1)
        return text_embeddings.view(bs_embed * num_images_per_prompt, seq_len, -1)


def _count(x):
    """Number of prompts in a str / list of str / tensor argument."""
    return 1 if isinstance(x, str) else len(x)


def _require_matching_negatives(positive, negative):
    """generate() accepts prompts, token ids or embeddings, each with an optional negative counterpart: when the negative is
    given it needs one entry per positive one (same error type and wording as reference stable_diffusion.py:405-414)."""
    if positive is None and negative is None:
        return
    if negative is None or len(negative) == 0:
        return
    if _count(negative) != _count(positive):
        raise ValueError('len(prompts) and len(negative_prompts) must be the same. '
                         'A negative prompt must be provided for each given prompt.')


def _require_some_prompt(*candidates):
    """reference stable_diffusion.py:416-418: at least one way of giving the prompt."""
    if all(c is None for c in candidates):
        raise ValueError('Must provide one of `prompt`, `tokenized_prompts`, or `prompt_embeds`')


def stable_diffusion_2(
    model_name: str = 'stabilityai/stable-diffusion-2-base',
    pretrained: bool = True,
    train_metrics: Optional[List] = None,
    val_metrics: Optional[List] = None,
    val_guidance_scales: Optional[List] = None,
    val_seed: int = 1138,
    loss_bins: Optional[List] = None,
    precomputed_latents: bool = False,
    encode_latents_in_fp16: bool = True,
    fsdp: bool = True,
    unet_config: Optional[dict] = None,
    build_encoders: Optional[bool] = None,
    vae_config: Optional[dict] = None,
    text_encoder_config: Optional[dict] = None,
):
    """Same signature as reference diffusion/models/models.py:28-39 (+ `unet_config` to override the SD-2-base UNet
    config for small test models).  The UNet is random-initialised from the SD-2-base config (`pretrained=False`,
    the yaml default); the HF hub is not reachable here, so `pretrained=True` raises, and VAE / CLIP / tokenizer are
    ALWAYS attached like the reference does at models.py:80-85 (native modules of diffusion_b200/encoders.py, random
    init, frozen, their engines built lazily on first use) - the SD-2-base yamls train on precomputed latents but evaluate
    on image / caption batches; `build_encoders=False` skips them explicitly (small test models).  The tokenizer needs
    vocabulary files that are not on disk, so batches must carry token ids (the reference dataset tokenizes in
    `__getitem__`)."""
    if pretrained:
        raise ValueError('pretrained=True needs the HF hub checkpoint of the UNet; load a state_dict into '
                         'model.unet instead (parameter names follow diffusers)')
    if train_metrics is None:
        train_metrics = [MeanSquaredError()]
    if val_metrics is None:
        val_metrics = [MeanSquaredError()]
    if val_guidance_scales is None:
        val_guidance_scales = [1.0, 3.0, 7.0]
    if loss_bins is None:
        loss_bins = [(0, 1)]
    unet = UNet2DConditionModel(**(unet_config or SD2_BASE_UNET_CONFIG))
    vae = text_encoder = None
    if build_encoders is None:
        build_encoders = True
    if build_encoders:  # random-init native VAE / text tower (diffusers / transformers parameter names: load_state_dict works)
        from diffusion_b200.encoders import SD2_TEXT_CONFIG, SD2_VAE_CONFIG, AutoencoderKL, CLIPTextModel
        vae = AutoencoderKL(**(vae_config or SD2_VAE_CONFIG))
        text_encoder = CLIPTextModel(**(text_encoder_config or SD2_TEXT_CONFIG))
    model = StableDiffusion(unet=unet, vae=vae, text_encoder=text_encoder, tokenizer=None, noise_scheduler=DDPMScheduler(),
                            inference_noise_scheduler=DDIMScheduler(), train_metrics=train_metrics, val_metrics=val_metrics,
                            val_guidance_scales=val_guidance_scales, val_seed=val_seed, loss_bins=loss_bins,
                            precomputed_latents=precomputed_latents, encode_latents_in_fp16=encode_latents_in_fp16, fsdp=fsdp)
    if torch.cuda.is_available():
        model = model.to(torch.device('cuda', torch.cuda.current_device()))
        model.unet.bind_arena()  # final parameter storage before any wrapper (DDP) looks at the parameters
    return model

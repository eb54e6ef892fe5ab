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


class _FusedMSE(torch.autograd.Function):
    """loss = mean((pred - noise)^2) computed by the fused head kernel from the engine's pred8 buffer; the same
    launch writes dL/dpred so that backward starts without an elementwise pass."""

    @staticmethod
    def forward(ctx_, pred, noise, eng):
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
            if isinstance(metric, MeanSquaredError):
                for bin_ in loss_bins:
                    new_metric = type(metric)()
                    new_metric.loss_bin = bin_
                    self.val_metrics[f'{metric.__class__.__name__}-bin-{bin_[0]}-to-{bin_[1]}'.replace('.', 'p')] = new_metric
            else:
                self.val_metrics[metric.__class__.__name__] = metric
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
                if self.encode_latents_in_fp16:
                    frozen.half()
        if fsdp:
            for m, flag in ((self.text_encoder, False), (self.vae, False), (self.unet, True)):
                if m is not None:
                    m._fsdp_wrap = flag
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
        timesteps, noise, _, _, _, used = ops.noise_sched_fwd(eng.ctx, latents, ac, seed, offset, eng.in_temb.shape[1],
                                                              out_nhwc8=eng.in_x8, out_temb=eng.in_temb)
        gen.set_offset(offset + used)
        eng.set_context(conditioning)
        from diffusion_b200.engine import unet_apply
        pred = unet_apply(self.unet, latents, timesteps, conditioning, prepared=True)
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
        return self.forward(batch)

    def get_metrics(self, is_train: bool = False):
        metrics = self.train_metrics if is_train else self.val_metrics
        if isinstance(metrics, Metric):
            return {metrics.__class__.__name__: metrics}
        if isinstance(metrics, list):
            return {m.__class__.__name__: m for m in metrics}
        return dict(metrics)

    def update_metric(self, batch, outputs, metric):
        if isinstance(metric, MeanSquaredError) and hasattr(metric, 'loss_bin'):
            lo, hi = metric.loss_bin
            T_max = self.noise_scheduler.num_train_timesteps
            idx = torch.where((outputs[2] >= lo * T_max) & (outputs[2] < hi * T_max))
            metric.update(outputs[0][idx], outputs[1][idx])
        else:
            metric.update(outputs[0], outputs[1])


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
):
    """Same signature as reference diffusion/models/models.py:28-39 (+ `unet_config` to override the SD-2-base UNet
    config for small test models).  The UNet is random-initialised from the SD-2-base config (`pretrained=False`,
    the yaml default); the HF hub is not reachable here, so `pretrained=True` raises, and VAE / CLIP / tokenizer are
    only attached when precomputed latents are not used and their weights can be loaded."""
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
    model = StableDiffusion(unet=unet, vae=None, text_encoder=None, tokenizer=None, noise_scheduler=DDPMScheduler(),
                            inference_noise_scheduler=None, train_metrics=train_metrics, val_metrics=val_metrics,
                            val_guidance_scales=val_guidance_scales, val_seed=val_seed, loss_bins=loss_bins,
                            precomputed_latents=precomputed_latents, encode_latents_in_fp16=encode_latents_in_fp16, fsdp=fsdp)
    if torch.cuda.is_available():
        model = model.to(torch.device('cuda', torch.cuda.current_device()))
    return model

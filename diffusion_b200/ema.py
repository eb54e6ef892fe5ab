"""Exponential moving average of the training weights - host-side mirror of the reference
`diffusion/algorithms/ema.py` (`compute_ema` :26-76, `EMA` :90-367, `EMAParameters` :370-447) with the update itself done
by the library kernel `sd2_ema_update`:

  * parameters that live in an engine `ParamArena` are averaged by ONE launch over the flat fp32 arena
    (12 B per parameter: read ema + param, write ema);
  * any other contiguous fp32 CUDA tensor goes through the same kernel tensor by tensor;
  * there is no CPU path: a CPU model raises.

The arithmetic is the reference's expression `ema * smoothing + param * (1 - smoothing)` evaluated with the same three
roundings, so the averaged weights are bit-identical to `compute_ema` run with torch ops on the same inputs.

Composer itself is not part of this repository; `EMA.match/apply` accept Composer's `Event`/`State` objects by duck typing
(event name, `state.model`, `state.timestamp.batch/epoch`), and `EMA.update(model, batch=..)` is the Composer-free entry.
"""
import itertools
import re

import torch

from diffusion_b200 import ops

__all__ = ['EMA', 'EMAParameters', 'compute_ema']


def _arenas_of(model):
    """Engine arenas whose parameters all belong to `model`: [(arena, {qualified name: arena name})]."""
    from diffusion_b200.engine import ParamArena
    by_id = {id(p): n for n, p in model.named_parameters()}
    out = []
    for arena in ParamArena.live():
        if not arena.bound():
            continue
        names = {}
        for an, p in arena.params.items():
            qn = by_id.get(id(p))
            if qn is None:
                names = None
                break
            names[qn] = an
        if names:
            out.append((arena, names))
    return out


class EMAParameters:
    """Stores the parameters and buffers of a model needed for averaging (reference ema.py:370-447).  Parameters of an
    engine arena are stored as views of one flat fp32 clone of that arena."""

    def __init__(self, model):
        self.named_parameters_dict = {}
        self.named_buffers_dict = {}
        self._flat = []  # (arena, flat fp32 clone of arena.p32)
        if model is None:
            return
        covered = set()
        with torch.no_grad():
            for arena, names in _arenas_of(model):
                flat = arena.p32.clone()
                self._flat.append((arena, flat))
                for qn, an in names.items():
                    if arena.params[an].requires_grad:
                        self.named_parameters_dict[qn] = arena._view(flat, an)
                    covered.add(qn)
            for name, p in model.named_parameters():
                if name not in covered and p.requires_grad:
                    self.named_parameters_dict[name] = p.data.clone()
            self.named_buffers_dict = {name: b.data.clone() for name, b in model.named_buffers()}

    def named_parameters(self):
        return self.named_parameters_dict.items()

    def named_buffers(self):
        return self.named_buffers_dict.items()

    def swap_params(self, model):
        """Swaps the parameters and buffers of a model with the ema parameters."""
        with torch.no_grad():
            for name, t in itertools.chain(model.named_parameters(), model.named_buffers()):
                store = self.named_parameters_dict if name in self.named_parameters_dict else self.named_buffers_dict
                if name in store:
                    tmp = t.detach().clone()
                    t.copy_(store[name])
                    store[name].copy_(tmp)

    def transfer_ema_params(self, model):
        """Transfers the parameters and buffers from the ema model to the supplied model."""
        with torch.no_grad():
            for name, t in itertools.chain(model.named_parameters(), model.named_buffers()):
                store = self.named_parameters_dict if name in self.named_parameters_dict else self.named_buffers_dict
                if name in store:
                    t.copy_(store[name])

    def move_params_to_device(self, destination_model):
        """Moves the ema parameters and buffers to the device of a destination model."""
        by_name = dict(itertools.chain(destination_model.named_parameters(), destination_model.named_buffers()))
        for store in (self.named_parameters_dict, self.named_buffers_dict):
            for name in list(store):
                if name in by_name and store[name].device != by_name[name].device:
                    store[name] = store[name].to(by_name[name].device)
                    self._flat = []  # views of the flat clones are gone with the move


def _ema_tensor(ema, param, smoothing):
    if ema.device.type != 'cuda' or param.device.type != 'cuda':
        raise RuntimeError('diffusion_b200.ema needs CUDA tensors (sm_100a): there is no CPU fallback')
    if ema.dtype == torch.float32 and param.dtype == torch.float32 and ema.is_contiguous() and param.is_contiguous():
        ops.ema_update(ops.get_ctx(ema.device), ema, param, smoothing)
    else:  # non-fp32 buffers or strided tensors that are not part of a flat arena clone: the reference expression
        ema.copy_(ema * smoothing + param * (1. - smoothing))


def compute_ema(model, ema_model, smoothing: float = 0.99) -> None:
    """In-place `W_ema = smoothing * W_ema + (1 - smoothing) * W_model` (reference ema.py:26-76)."""
    with torch.no_grad():
        if isinstance(ema_model, torch.nn.Module):
            ema_params = ema_model.state_dict()
            for name, p in itertools.chain(model.named_parameters(), model.named_buffers()):
                if name in ema_params:
                    _ema_tensor(ema_params[name], p.data, smoothing)
            return
        if not isinstance(ema_model, EMAParameters):
            raise ValueError('ema_model must be a torch.nn.Module or EMAParameters')
        done = set()
        for arena, flat in ema_model._flat:
            names = dict(_arenas_of(model)).get(arena)
            if names is None or not arena.bound():
                continue
            ops.ema_update(arena_ctx(arena), flat, arena.p32, smoothing)  # one launch over the whole arena
            done.update(names)
        for name, p in itertools.chain(model.named_parameters(), model.named_buffers()):
            if name in done:
                continue
            if name in ema_model.named_parameters_dict:
                _ema_tensor(ema_model.named_parameters_dict[name], p.data, smoothing)
            if name in ema_model.named_buffers_dict:
                _ema_tensor(ema_model.named_buffers_dict[name], p.data, smoothing)


def arena_ctx(arena):
    return ops.get_ctx(arena.p32.device)


_TIME = re.compile(r'^\s*([0-9]*\.?[0-9]+(?:[eE][-+]?[0-9]+)?)\s*(ba|ep|dur)\s*$')


def _parse_time(s):
    m = _TIME.match(s) if isinstance(s, str) else None
    if m is None:
        raise ValueError(f'cannot parse time string {s!r} (supported units: ba, ep, dur)')
    v = float(m.group(1))
    return (int(v) if m.group(2) != 'dur' else v), m.group(2)


def _event_name(event):
    return str(getattr(event, 'name', event)).upper()


class EMA:
    """Same constructor arguments, validation and smoothing formula as the reference algorithm (ema.py:90-190)."""

    def __init__(self, half_life='1000ba', smoothing=None, ema_start='0.0dur', update_interval=None):
        self.ema_model = None
        self.ema_weights_active = False
        self.ema_started = False
        self.serialized_attributes = ['ema_model', 'ema_weights_active', 'ema_started']
        if half_life is None and smoothing is None:
            raise ValueError('Either half_life or smoothing must be specified')
        if half_life is not None and smoothing is not None:
            raise ValueError('Only one of  half_life or smoothing can be specified')
        self.half_life = _parse_time(half_life) if half_life is not None else None
        self.ema_start = _parse_time(ema_start)
        if update_interval is None:
            self.update_interval = (1, self.half_life[1]) if self.half_life else (1, 'ba')
        elif isinstance(update_interval, str):
            self.update_interval = _parse_time(update_interval)
        else:
            raise ValueError('update_interval must be None or a time string.')
        if self.half_life is not None and self.half_life[1] != self.update_interval[1]:
            raise ValueError('Units of half_life and update_interval must match.')
        if self.update_interval[1] not in ('ba', 'ep'):
            raise ValueError(f'Invalid time unit for parameter update_interval: {self.update_interval[1]}')
        if smoothing is None and self.half_life:
            self.smoothing = 2**(-(self.update_interval[0] / self.half_life[0]))
        else:
            self.smoothing = smoothing
        self.update_event = 'BATCH_END' if self.update_interval[1] == 'ba' else 'EPOCH_END'

    # ---- Composer-free entry point ---------------------------------------------------------------------------
    def update(self, model, batch=None, epoch=None, elapsed_duration=None):
        """Call once per finished batch (or epoch).  Starts the average when `ema_start` is reached and applies
        `compute_ema` every `update_interval`.  Returns True if the average was updated."""
        count = batch if self.update_interval[1] == 'ba' else epoch
        if not self.ema_started:
            v, unit = self.ema_start
            now = {'ba': batch, 'ep': epoch, 'dur': elapsed_duration}[unit]
            if now is None:
                now = 0 if v == 0 else None
            if now is None or now < v:
                return False
            self.ema_model = EMAParameters(model)
            self.ema_started = True
        if count is not None and count % self.update_interval[0] != 0:
            return False
        compute_ema(model, self.ema_model, smoothing=self.smoothing)
        return True

    # ---- Composer Algorithm protocol (duck-typed) ------------------------------------------------------------
    def _count(self, state, unit):
        ts = getattr(state, 'timestamp', None)
        v = getattr(ts, {'ba': 'batch', 'ep': 'epoch'}[unit], 0)
        return int(getattr(v, 'value', v))

    def match(self, event, state) -> bool:
        ev = _event_name(event)
        if ev == 'INIT':
            return True
        if ev == self.update_event and not self.ema_started:
            v, unit = self.ema_start
            if unit == 'dur':
                d = state.get_elapsed_duration() if hasattr(state, 'get_elapsed_duration') else None
                start = d is not None and v <= float(getattr(d, 'value', d))
            else:
                start = v <= self._count(state, unit)
            if start:
                self.ema_model = EMAParameters(state.model)
                self.ema_started = True
        if ev in ('BATCH_START', 'EVAL_START', 'EVAL_END', 'FIT_START', 'PREDICT_START') and self.ema_started:
            return True
        if ev == self.update_event and self.ema_started:
            return self._count(state, self.update_interval[1]) % self.update_interval[0] == 0
        # reference ema.py:221-227: on checkpoint events, match when a checkpoint saver is about to write, so that apply()
        # swaps the EMA weights into the model and the checkpoint carries them with ema_weights_active=True.  Savers are found
        # by duck type (Composer is not importable here): a callback with a callable `save_interval(state, event)`.
        if ev in ('BATCH_CHECKPOINT', 'EPOCH_CHECKPOINT') and self.ema_started:
            for cb in getattr(state, 'callbacks', None) or ():
                interval = getattr(cb, 'save_interval', None)
                if callable(interval) and interval(state, event) is True:
                    return True
        return False

    def apply(self, event, state, logger=None) -> None:
        ev = _event_name(event)
        if ev == 'INIT':
            self.ema_model = EMAParameters(state.model)
        assert self.ema_model is not None
        if ev in ('FIT_START', 'PREDICT_START'):
            self.ema_model.move_params_to_device(destination_model=state.model)
        if ev == 'BATCH_START' and self.ema_weights_active:
            self._ensure_training_weights_active(state)
        if ev in ('BATCH_END', 'EPOCH_END'):
            compute_ema(state.model, self.ema_model, smoothing=self.smoothing)
        if ev == 'EVAL_START':
            self.ema_model.move_params_to_device(destination_model=state.model)
            self._ensure_ema_weights_active(state)
        if ev == 'EVAL_END':
            self._ensure_training_weights_active(state)
        if ev in ('BATCH_CHECKPOINT', 'EPOCH_CHECKPOINT'):
            self._ensure_ema_weights_active(state)

    def _ensure_training_weights_active(self, state):
        if self.ema_weights_active and self.ema_model is not None:
            self.ema_model.swap_params(model=state.model)
            self.ema_weights_active = False

    def _ensure_ema_weights_active(self, state):
        if not self.ema_weights_active and self.ema_model is not None:
            self.ema_model.swap_params(model=state.model)
            self.ema_weights_active = True

    def get_ema_model(self, model):
        assert self.ema_model is not None
        if self.ema_weights_active:
            raise ValueError('The ema weight are currently contained in the composer model.')
        self.ema_model.transfer_ema_params(model=model)
        return model

    def get_training_model(self, model):
        assert self.ema_model is not None
        if not self.ema_weights_active:
            raise ValueError('The training weights are currently contained in the composer model.')
        self.ema_model.transfer_ema_params(model=model)
        return model

    def state_dict(self):
        out = {}
        for name in self.serialized_attributes:
            if name == 'ema_model':
                out[name] = {'named_parameters_dict': self.ema_model.named_parameters_dict if self.ema_model else {},
                             'named_buffers_dict': self.ema_model.named_buffers_dict if self.ema_model else {}}
            else:
                out[name] = getattr(self, name)
        return out

    def load_state_dict(self, state, strict=False):
        for name, value in state.items():
            if name == 'ema_model':
                self.ema_model = EMAParameters(None)
                self.ema_model.named_parameters_dict = value['named_parameters_dict']
                self.ema_model.named_buffers_dict = value['named_buffers_dict']
            elif name != 'repr':
                setattr(self, name, value)

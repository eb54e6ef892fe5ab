"""Precomputed-latent wire format -> device batch.

The reference dataset stores, per sample, raw little-endian fp16 bytes: `caption_latents` (77 x 1024), `latents_256`
(4 x 32 x 32) and `latents_512` (4 x 64 x 64) (writer: reference scripts/precompute_latents.py:252-272,324-326; reader:
diffusion/datasets/laion/laion.py:103-111, which copies every field through numpy into a fresh tensor and leaves the
stacking to torch's default collate).  Here the bytes of a whole batch are gathered ONCE, by the library's host function
`sd2_wire_gather`, straight into a pinned staging buffer, and shipped with one asynchronous copy per field on a
dedicated copy stream; the consumer waits on an event, so the upload of batch k+1 overlaps the step of batch k.  The
device batch stays fp16 - K1 reads fp16 latents natively and the engine casts the conditioning to bf16 with
`sd2_cast_to_bf16` - so no torch op touches the data between the dataset bytes and the kernels.

`LatentBatcher.collate` also accepts the tensors the reference `__getitem__` produces (`image_latents`,
`caption_latents`), so it can be passed as `collate_fn` to an unchanged reference dataset.
"""
import ctypes as C

import numpy as np
import torch

from diffusion_b200 import _lib

CAPTION_SHAPE = (77, 1024)
LATENT_SHAPES = {256: (4, 32, 32), 512: (4, 64, 64)}


def _buffer_address(obj):
    """(address, nbytes, keepalive) of a bytes-like object or a CPU tensor, without copying."""
    if torch.is_tensor(obj):
        if obj.device.type != 'cpu' or not obj.is_contiguous():
            raise ValueError('wire samples must be contiguous CPU tensors or bytes-like objects')
        return obj.data_ptr(), obj.numel() * obj.element_size(), obj
    arr = np.frombuffer(obj, dtype=np.uint8)
    return arr.ctypes.data, arr.nbytes, arr


class LatentBatcher:
    """Pinned, double-buffered collate + upload of precomputed-latent batches.

    batch_size: samples per batch; image_size: 256 or 512 (selects `latents_256` / `latents_512`); depth: number of pinned
    staging slots (2 = one being filled while one is in flight)."""

    def __init__(self, batch_size, image_size=256, device=None, depth=2, image_latents_key='image_latents',
                 text_latents_key='caption_latents', pin=None):
        if image_size not in LATENT_SHAPES:
            raise ValueError(f'image_size must be one of {sorted(LATENT_SHAPES)}')
        self.B, self.image_size = int(batch_size), image_size
        self.lat_shape, self.cap_shape = LATENT_SHAPES[image_size], CAPTION_SHAPE
        self.lat_bytes = int(np.prod(self.lat_shape)) * 2
        self.cap_bytes = int(np.prod(self.cap_shape)) * 2
        self.image_latents_key, self.text_latents_key = image_latents_key, text_latents_key
        self.device = torch.device(device) if device is not None else None
        pin = torch.cuda.is_available() if pin is None else pin
        self.lib = _lib.load()
        self._slots = []
        for _ in range(max(1, depth)):
            lat = torch.empty((self.B,) + self.lat_shape, dtype=torch.float16)
            cap = torch.empty((self.B,) + self.cap_shape, dtype=torch.float16)
            if pin:
                lat, cap = lat.pin_memory(), cap.pin_memory()
            self._slots.append({'lat': lat, 'cap': cap, 'free': None})
        self._next = 0
        self._stream = None

    # ---- host side -------------------------------------------------------------------------------------------
    def _field(self, sample, wire_key, tensor_key, nbytes):
        obj = sample.get(wire_key) if wire_key in sample else sample.get(tensor_key)
        if obj is None:
            raise KeyError(f'sample has neither {wire_key!r} nor {tensor_key!r}')
        addr, n, keep = _buffer_address(obj)
        if torch.is_tensor(obj) and obj.dtype != torch.float16:
            raise ValueError(f'{tensor_key}: the wire format is fp16, got {obj.dtype}')
        if n != nbytes:
            raise ValueError(f'{wire_key}/{tensor_key}: expected {nbytes} bytes per sample, got {n}')
        return addr, keep

    def _gather(self, samples, wire_key, tensor_key, nbytes, dst):
        n = len(samples)
        ptrs, keep = (C.c_void_p * n)(), []
        for i, s in enumerate(samples):
            ptrs[i], k = self._field(s, wire_key, tensor_key, nbytes)
            keep.append(k)
        rc = self.lib.sd2_wire_gather(C.cast(ptrs, C.c_void_p), n, nbytes, dst.data_ptr())
        if rc != 0:
            raise RuntimeError(f'sd2_wire_gather failed with code {rc}')

    def collate(self, samples):
        """List of dataset samples -> {'image_latents': fp16 (n,4,h,w), 'caption_latents': fp16 (n,77,1024)} in the next
        pinned staging slot (n <= batch_size).  Usable as a DataLoader `collate_fn`."""
        n = len(samples)
        if n == 0 or n > self.B:
            raise ValueError(f'batch of {n} samples (batcher was built for 1..{self.B})')
        slot = self._slots[self._next]
        self._next = (self._next + 1) % len(self._slots)
        if slot['free'] is not None:  # the previous upload from this slot must have left the host buffer
            slot['free'].synchronize()
            slot['free'] = None
        self._gather(samples, f'latents_{self.image_size}', self.image_latents_key, self.lat_bytes, slot['lat'])
        self._gather(samples, 'caption_latents', self.text_latents_key, self.cap_bytes, slot['cap'])
        out = {self.image_latents_key: slot['lat'][:n], self.text_latents_key: slot['cap'][:n]}
        out['_slot'] = slot
        return out

    # ---- device side -----------------------------------------------------------------------------------------
    def to_device(self, batch, device=None):
        """Asynchronous upload of a collated batch on the copy stream; the current stream is made to wait for it."""
        device = torch.device(device) if device is not None else self.device
        if device is None or device.type != 'cuda':
            raise RuntimeError('LatentBatcher.to_device needs a CUDA device (no CPU path)')
        if self._stream is None:
            self._stream = torch.cuda.Stream(device)
        slot = batch.get('_slot')
        out = {}
        with torch.cuda.stream(self._stream):
            for k, v in batch.items():
                if k != '_slot':
                    out[k] = v.to(device, non_blocking=True)
            done = torch.cuda.Event()
            done.record(self._stream)
        if slot is not None:
            slot['free'] = done
        cur = torch.cuda.current_stream(device)
        cur.wait_event(done)
        for v in out.values():
            v.record_stream(cur)
        return out

    def __call__(self, samples):
        return self.collate(samples)

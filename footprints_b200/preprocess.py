"""Input edge of the inference callers: uint8 RGB frames -> the network's float32 ``[B,3,H,W]`` input, on the GPU.

Drop-in for the two transforms the reference applies to every frame on the host
(``predict_simple.py:41-43,51-60``, ``datasets/inference_dataset.py:25-27,45-48``)::

    self.resizer = transforms.Resize((H, W), interpolation=Image.ANTIALIAS)
    self.totensor = transforms.ToTensor()
    image = self.totensor(self.resizer(image))

``ResizeToTensor((H, W))(image)`` returns the same tensor, bit for bit, computed by ``fp_resize_u8_horizontal`` +
``fp_resize_u8_vertical_to_f32`` (csrc/input_edge.cu): Pillow's two-pass fixed-point Lanczos resampling with its uint8
intermediate, fused with the ``/255`` conversion.  The host only builds the per-geometry tap tables (Pillow's
``precompute_coeffs`` + ``normalize_coeffs_8bpc``: a few thousand numbers per (input size, output size), cached and kept
resident on the device) and copies the raw uint8 frame -- 1/4 of the bytes of the float tensor the reference uploads.
There is no CPU fallback.
"""
import math

import numpy as np
import torch

from . import _lib

_PRECISION_BITS = 32 - 8 - 2


def _lanczos(x):
    if not (-3.0 <= x < 3.0):
        return 0.0
    if x == 0.0:
        return 1.0
    a = x * math.pi
    b = (x / 3.0) * math.pi
    return (math.sin(a) / a) * (math.sin(b) / b)


def tap_table(in_size, out_size):
    """Pillow's tap table for resampling ``in_size`` -> ``out_size`` samples with the Lanczos (a=3) filter over the whole
    axis: ``(ksize, bounds int32 [out,2] = (first tap, count), taps int32 [ksize,out])``.  Weights are evaluated with
    ``math.sin`` (the C library Pillow itself calls), normalised in double precision and rounded to 22-bit fixed point
    half away from zero, exactly like ``Resample.c``."""
    scale = float(in_size) / out_size
    filterscale = scale if scale > 1.0 else 1.0
    support = 3.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    inv = 1.0 / filterscale
    one = float(1 << _PRECISION_BITS)
    bounds = np.zeros((out_size, 2), np.int32)
    taps = np.zeros((ksize, out_size), np.int32)
    for o in range(out_size):
        centre = (o + 0.5) * scale
        first = max(int(centre - support + 0.5), 0)
        count = min(int(centre + support + 0.5), in_size) - first
        w = [_lanczos((t + first - centre + 0.5) * inv) for t in range(count)]
        total = 0.0
        for v in w:
            total += v
        for t, v in enumerate(w):
            if total != 0.0:
                v = v / total
            taps[t, o] = int(v * one - 0.5) if v < 0 else int(v * one + 0.5)
        bounds[o] = (first, count)
    return ksize, bounds, taps


class ResizeToTensor:
    """``transforms.Resize(size, interpolation=Image.ANTIALIAS)`` followed by ``transforms.ToTensor()``, on the GPU.

    Accepts a PIL RGB image, a uint8 ``[H,W,3]`` / ``[B,H,W,3]`` numpy array or torch tensor (host or device).  Returns a
    float32 CUDA tensor ``[3,h,w]`` for a single frame (what ``totensor(resizer(image))`` returns) or ``[B,3,h,w]`` for a
    batch of equally sized frames.
    """

    def __init__(self, size, device=None):
        self.size = (int(size[0]), int(size[1]))
        self.device = torch.device('cuda' if device is None else device)
        if self.device.type != 'cuda':
            raise RuntimeError('footprints_b200.preprocess runs on a B200 only (no CPU fallback)')
        self._lib = _lib.load()
        self._tables = {}

    def _table(self, hin, win):
        key = (hin, win)
        t = self._tables.get(key)
        if t is None:
            ho, wo = self.size
            kh, bh, th = tap_table(win, wo)
            kv, bv, tv = tap_table(hin, ho)
            y0 = int(bv[0, 0])
            y1 = int(bv[-1, 0] + bv[-1, 1])
            bv = bv.copy()
            bv[:, 0] -= y0                                        # relative to the first source row the vertical pass reads
            dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(self.device)   # noqa: E731
            t = dict(y0=y0, ht=y1 - y0, kh=kh, kv=kv, bh=dev(bh), th=dev(th), bv=dev(bv), tv=dev(tv))
            self._tables[key] = t
        return t

    def __call__(self, frames):
        single = False
        if not isinstance(frames, (np.ndarray, torch.Tensor)):     # PIL image
            if getattr(frames, 'mode', 'RGB') != 'RGB':
                frames = frames.convert('RGB')                     # pil_loader does the same (utils.py:24-29)
            frames = np.array(frames)                              # (a copy: PIL hands out a read-only view)
        if isinstance(frames, np.ndarray):
            frames = torch.from_numpy(np.ascontiguousarray(frames))
        if frames.dtype != torch.uint8 or frames.dim() not in (3, 4) or frames.shape[-1] != 3:
            raise RuntimeError(f'ResizeToTensor expects uint8 [H,W,3] or [B,H,W,3] frames, got {frames.dtype} {tuple(frames.shape)}')
        if frames.dim() == 3:
            single, frames = True, frames[None]
        frames = frames.to(self.device, non_blocking=True).contiguous()
        b, hin, win, _ = frames.shape
        ho, wo = self.size
        t = self._table(hin, win)
        with torch.cuda.device(self.device):
            _lib.require_device()
            stream = torch.cuda.current_stream().cuda_stream
            tmp = torch.empty((b, t['ht'], wo, 4), dtype=torch.uint8, device=self.device)
            out = torch.empty((b, 3, ho, wo), dtype=torch.float32, device=self.device)
            _lib.check(self._lib.fp_resize_u8_horizontal(frames.data_ptr(), b, hin, win, t['y0'], t['ht'], wo, t['bh'].data_ptr(),
                                                         t['th'].data_ptr(), t['kh'], tmp.data_ptr(), stream))
            _lib.check(self._lib.fp_resize_u8_vertical_to_f32(tmp.data_ptr(), b, t['ht'], wo, ho, t['bv'].data_ptr(),
                                                              t['tv'].data_ptr(), t['kv'], out.data_ptr(), stream))
        return out[0] if single else out

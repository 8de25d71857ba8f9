"""PIL-compatible bicubic resize of uint8 images on the device.

The reflection loop feeds each parent image back as a 512x512 condition:
`Image.open(path).resize((condition_size, condition_size))` (tts/tts_reflectionflow.py:276-277), i.e.
Pillow's default BICUBIC filter with antialiasing.  Pillow's uint8 path is deterministic integer
arithmetic (ImagingResample: coefficients in double -> fixed point with 22 fractional bits, a
horizontal pass rounded to uint8, then a vertical pass rounded to uint8), restated here so that the
device result is bit-identical to PIL (tests/test_resize.py pins the host restatement against PIL
itself; the -m gpu test pins the kernel)."""
from __future__ import annotations

import ctypes
import math
from functools import lru_cache

import numpy as np
import torch

PRECISION_BITS = 32 - 8 - 2


def _bicubic(x: float) -> float:
    a = -0.5
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


@lru_cache(maxsize=16)
def precompute_coeffs(in_size: int, out_size: int):
    """Pillow's precompute_coeffs() + normalize_coeffs_8bpc() for the BICUBIC filter (support 2).
    Returns (bounds int32 [out, 2] = (xmin, count), coeffs int32 [out, ksize])."""
    scale = filterscale = in_size / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.float64)
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        ss = 1.0 / filterscale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        ww = 0.0
        for x in range(xmax):
            w = _bicubic((x + xmin - center + 0.5) * ss)
            kk[xx, x] = w
            ww += w
        if ww != 0.0:
            kk[xx, :xmax] /= ww
        bounds[xx] = (xmin, xmax)
    ki = np.where(kk < 0, np.trunc(-0.5 + kk * (1 << PRECISION_BITS)),
                  np.trunc(0.5 + kk * (1 << PRECISION_BITS))).astype(np.int32)
    return bounds, ki


def resize_u8_reference(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """numpy restatement of ImagingResample for uint8 HWC images (horizontal pass, then vertical)."""
    h, w, c = img.shape
    half = 1 << (PRECISION_BITS - 1)
    x = img.astype(np.int64)
    if out_w != w:
        b, k = precompute_coeffs(w, out_w)
        out = np.empty((h, out_w, c), dtype=np.int64)
        for xx in range(out_w):
            x0, n = int(b[xx, 0]), int(b[xx, 1])
            acc = half + (x[:, x0:x0 + n, :] * k[xx, :n][None, :, None].astype(np.int64)).sum(axis=1)
            out[:, xx, :] = np.clip(acc >> PRECISION_BITS, 0, 255)
        x = out
    if out_h != h:
        b, k = precompute_coeffs(h, out_h)
        out = np.empty((out_h, x.shape[1], c), dtype=np.int64)
        for yy in range(out_h):
            y0, n = int(b[yy, 0]), int(b[yy, 1])
            acc = half + (x[y0:y0 + n, :, :] * k[yy, :n][:, None, None].astype(np.int64)).sum(axis=0)
            out[yy] = np.clip(acc >> PRECISION_BITS, 0, 255)
        x = out
    return x.astype(np.uint8)


_TABLES = {}


def resize_u8(img: torch.Tensor, out_h: int, out_w: int) -> torch.Tensor:
    """uint8 [B, H, W, 3] (device) -> uint8 [B, out_h, out_w, 3], bit-identical to
    PIL.Image.resize((out_w, out_h)) (BICUBIC).  Runs rf_op_resize_u8 (two integer passes)."""
    from . import _lib as L
    assert img.dtype == torch.uint8 and img.dim() == 4 and img.shape[-1] == 3 and img.is_cuda
    B, H, W, _ = img.shape
    dev = img.device
    key = (H, W, out_h, out_w, str(dev))
    if key not in _TABLES:
        bx, kx = precompute_coeffs(W, out_w)
        by, ky = precompute_coeffs(H, out_h)
        _TABLES[key] = tuple(torch.from_numpy(np.ascontiguousarray(t)).to(dev) for t in (bx, kx, by, ky))
    bx, kx, by, ky = _TABLES[key]
    img = img.contiguous()
    tmp = torch.empty((B, H, out_w, 3), dtype=torch.uint8, device=dev)
    out = torch.empty((B, out_h, out_w, 3), dtype=torch.uint8, device=dev)
    lib = L.load()
    with torch.cuda.device(dev):
        for b in range(B):
            L.check(lib.rf_op_resize_u8(L.ptr(img[b]), H, W, L.ptr(tmp[b]), L.ptr(out[b]), out_h, out_w,
                                        L.ptr(bx), L.ptr(kx), kx.shape[1], L.ptr(by), L.ptr(ky),
                                        ky.shape[1], L.cur_stream()), "rf_op_resize_u8")
    return out

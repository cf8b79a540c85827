"""NumPy restatement of the reference's image transform for the recording pass (oracle; test infrastructure only).

The reference builds every dataset item with ``transforms.Resize(img_size)``, ``CenterCrop(img_size)``, ``ToTensor``,
``Normalize(0.5, 0.5)`` (diagan-pkg/diagan/datasets/transform.py:3-41) applied to a PIL image
(e.g. color_mnist.py:90-100).  The arithmetic lives in two third-party dependencies that are NOT vendored in
/root/reference: torchvision (``environment.yml:44``: ``torchvision=0.8.2``; geometry of Resize(int) / CenterCrop) and Pillow
(``environment.yml:34``: ``pillow=8.1.0``; ``Image.resize(..., BILINEAR)`` = libImaging/Resample.c, 8-bit fixed-point path).  Restated here from
their published algorithms:

* Resize(int s) on a (w, h) image: the shorter side becomes s, the longer ``int(s * long / short)``
  (torchvision.transforms.functional.resize);  CenterCrop(s): top = int(round((h - s) / 2.0)), left likewise.
* Pillow ``ImagingResample`` for 8-bit channels: separable, HORIZONTAL pass first into a uint8 intermediate, then vertical.
  Per output coordinate xx: scale = in/out, filterscale = max(scale, 1), support = 1.0 * filterscale (bilinear/triangle),
  center = (xx + 0.5) * scale, xmin = int(center - support + 0.5) clamped at 0, xmax = int(center + support + 0.5) clamped at
  the input size, weights w = triangle((x + xmin - center + 0.5) / filterscale) normalised to sum 1, then converted to
  22-bit fixed point k = int(0.5 + w * 2**22) (``normalize_coeffs_8bpc``); a pixel is
  clip8((2**21 + sum_x in[x + xmin] * k[x]) >> 22).  Same-size resizes are a copy.

PINNED: ``tests/test_oracle_golden.py`` checks this restatement bit-for-bit against Pillow itself (12.2, the version installed in
this image) on every shape used by the reference's datasets plus odd shapes, and against ``tests/golden/resize_*.npz`` (outputs
of the reference's own ``get_transform`` run here).  The reference pins pillow 8.1.0, which cannot be installed here; its 8-bit
``ImagingResample`` path (PRECISION_BITS = 22, round-half-up, horizontal pass first) is the same algorithm as restated above
[UNVERIFIED against an 8.1.0 install].
"""
from __future__ import annotations

import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def resized_size(w: int, h: int, size: int):
    """torchvision Resize(int): (new_w, new_h)."""
    if w <= h:
        return size, int(size * h / w)
    return int(size * w / h), size


def crop_offsets(w: int, h: int, size: int):
    """torchvision CenterCrop(size) on a (w, h) image that is at least size x size: (left, top)."""
    return int(round((w - size) / 2.0)), int(round((h - size) / 2.0))


def coeffs(in_size: int, out_size: int):
    """Pillow precompute_coeffs + normalize_coeffs_8bpc for the bilinear filter over the whole axis.
    -> (bounds int32 [out,2] = (xmin, count), k int32 [out, ksize])."""
    scale = float(in_size) / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = np.zeros(ksize, np.float64)
        for x in range(xmax):
            t = abs((x + xmin - center + 0.5) * ss)
            w[x] = 1.0 - t if t < 1.0 else 0.0
        # Pillow accumulates ww in a plain left-to-right double loop
        ww = 0.0
        for x in range(xmax):
            ww += w[x]
        if ww != 0.0:
            w[:xmax] = w[:xmax] / ww
        for x in range(ksize):
            v = w[x] * (1 << PRECISION_BITS)
            kk[xx, x] = int(-0.5 + v) if w[x] < 0 else int(0.5 + v)
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _resample_axis(img: np.ndarray, axis: int, out_size: int) -> np.ndarray:
    """img uint8 [..., H, W, C]; resample `axis` (-3 = vertical, -2 = horizontal)."""
    in_size = img.shape[axis]
    bounds, kk = coeffs(in_size, out_size)
    src = np.moveaxis(img, axis, 0).astype(np.int64)
    out = np.empty((out_size,) + src.shape[1:], np.uint8)
    for xx in range(out_size):
        xmin, cnt = bounds[xx]
        acc = np.full(src.shape[1:], 1 << (PRECISION_BITS - 1), np.int64)
        for x in range(cnt):
            acc += src[xmin + x] * int(kk[xx, x])
        out[xx] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def resize_bilinear_u8(img: np.ndarray, out_w: int, out_h: int) -> np.ndarray:
    """Pillow Image.resize((out_w, out_h), BILINEAR) on uint8 [..., H, W, C]."""
    h, w = img.shape[-3], img.shape[-2]
    if (w, h) == (out_w, out_h):
        return img.copy()
    x = img
    if w != out_w:
        x = _resample_axis(x, -2, out_w)           # horizontal pass first, rounded to uint8
    if h != out_h:
        x = _resample_axis(x, -3, out_h)
    return x


def resize_center_crop_u8(img: np.ndarray, size: int) -> np.ndarray:
    """Resize(size) + CenterCrop(size) (transform.py:3-41) on uint8 [..., H, W, C] -> [..., size, size, C]."""
    h, w = img.shape[-3], img.shape[-2]
    nw, nh = resized_size(w, h, size)
    x = resize_bilinear_u8(img, nw, nh)
    left, top = crop_offsets(nw, nh, size)
    return np.ascontiguousarray(x[..., top:top + size, left:left + size, :])

"""CPU oracle of the recognizer's image preprocessing — TEST INFRASTRUCTURE, never the product path.

Restates, in plain numpy, the arithmetic the reference runs per image before the recognizer forward:

  * ``doc2tex/utils/predict_utils.py::resize`` (lines 14-115, the ``imgH is None``, no-resizer branch): optional
    ``cv2.resize(INTER_AREA)`` down-sampling, ``pad()`` crop-to-ink when ``opt['pad']``, ``minmax_size``, the
    albumentations test transform (``transform/math_transform.py:42-50``: ToGray, Normalize(mean, std), ToTensorV2) and the
    channel-0 slice;
  * ``doc2tex/utils/data_utils.py::pad`` (10-47) and ``minmax_size`` / ``get_divisible_size`` (50-82);
  * Pillow's ``Image.resize(..., LANCZOS)`` on 8-bit images (third-party: Pillow ``src/libImaging/Resample.c`` —
    ``precompute_coeffs``, ``normalize_coeffs_8bpc``, two 8-bit passes with 22 fractional bits), which ``minmax_size`` calls.

Pinning: ``oracle/make_golden.py::preprocess_case`` runs the reference's own ``pad`` / ``minmax_size`` (they import in the
build container: PIL + cv2 are present) and Pillow's / OpenCV's resamplers on seeded synthetic crops, asserts this
restatement equals them bit for bit, and stores the reference outputs in tests/golden/preprocess.npz.  The albumentations
normalisation cannot be imported (package absent); its published arithmetic ``(v - mean*255) * (1 / (std*255))`` in float32
is restated and pinned by the stored values only.

Divergence on purpose (documented in DESIGN.md): ``get_divisible_size`` leaves ``new_h`` / ``new_w`` unassigned when the
scaled size is already a multiple of 32 (data_utils.py:50-59) — ``minmax_size`` then dies with UnboundLocalError, which is
what happens to most images that need resizing.  The oracle implements the evident intent (a size that is already
divisible stays) and the golden cases use sizes where the reference survives.
"""
from __future__ import annotations

import math
from typing import Optional, Sequence, Tuple

import numpy as np

PRECISION_BITS = 32 - 8 - 2   # Pillow Resample.c


# ---------------------------------------------------------------------------------------------------------------------
# cv2.resize(..., INTER_AREA) with an integer scale (predict_utils.py:33-44): rounded box mean
# ---------------------------------------------------------------------------------------------------------------------
def area_downsample(img: np.ndarray, ratio: int) -> np.ndarray:
    H, W = img.shape
    assert H % ratio == 0 and W % ratio == 0, "integer-scale INTER_AREA needs sizes divisible by the ratio"
    box = img.reshape(H // ratio, ratio, W // ratio, ratio).astype(np.int64).sum(axis=(1, 3))
    n = ratio * ratio
    return ((box * 2 + n) // (2 * n)).astype(np.uint8)       # round half up of the box mean


# ---------------------------------------------------------------------------------------------------------------------
# data_utils.py::pad (10-47): min/max stretch, polarity, ink bounding box, crop, pad to /32 with BLACK (Image.new("L", dims))
# ---------------------------------------------------------------------------------------------------------------------
def ink_box(img: np.ndarray) -> Tuple[int, int, int, int, bool, int]:
    """(x, y, w, h, inverted, vmin) of data_utils.py:21-33.  The LA conversion adds a constant alpha of 255, so
    ``data.max()`` is 255 and ``data.min()`` the smallest grey value."""
    vmin = int(img.min())
    if vmin == 255:
        raise ValueError("blank image: no ink to crop to")
    data0 = (img - np.uint8(vmin)) / np.uint8(255 - vmin) * 255          # float64, the reference's expression order
    inverted = not (data0.mean() > 128)
    gray = (data0 > 128) if inverted else (data0 < 128)
    ys, xs = np.nonzero(gray)
    if ys.size == 0:
        raise ValueError("no pixel crosses the ink threshold")
    x0, x1, y0, y1 = int(xs.min()), int(xs.max()), int(ys.min()), int(ys.max())
    return x0, y0, x1 - x0 + 1, y1 - y0 + 1, inverted, vmin


def pad_to_ink(img: np.ndarray, divable: int = 32) -> Tuple[np.ndarray, Tuple[int, int, int, int]]:
    x, y, w, h, inverted, vmin = ink_box(img)
    data0 = (img - np.uint8(vmin)) / np.uint8(255 - vmin) * 255
    if inverted:
        data0 = 255 - data0
    crop = data0[y:y + h, x:x + w].astype(np.uint8)                       # truncation, like ndarray.astype
    nz_rows, nz_cols = np.nonzero(crop.any(axis=1))[0], np.nonzero(crop.any(axis=0))[0]
    if nz_rows.size == 0 or (nz_rows[0], nz_rows[-1], nz_cols[0], nz_cols[-1]) != (0, h - 1, 0, w - 1):
        # padded.paste(im, im.getbbox()) with a box smaller than the image raises in the reference (a full border row of zeros)
        raise ValueError("the cropped image has an all-zero border row or column")
    W2, H2 = (divable * ((v + divable - 1) // divable) for v in (w, h))
    out = np.zeros((H2, W2), dtype=np.uint8)
    out[:h, :w] = crop
    return out, (x, y, w, h)


# ---------------------------------------------------------------------------------------------------------------------
# Pillow 8-bit LANCZOS resize
# ---------------------------------------------------------------------------------------------------------------------
def _lanczos(x: float) -> float:
    if -3.0 <= x < 3.0:
        def sinc(v):
            if v == 0.0:
                return 1.0
            v = v * math.pi
            return math.sin(v) / v
        return sinc(x) * sinc(x / 3.0)
    return 0.0


def lanczos_coeffs(in_size: int, out_size: int):
    """Per output sample: (xmin, n taps, fixed-point weights) — Resample.c precompute_coeffs + normalize_coeffs_8bpc."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 3.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    ss = 1.0 / filterscale
    xmins = np.zeros(out_size, dtype=np.int32)
    counts = np.zeros(out_size, dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = [_lanczos((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            v = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        xmins[xx], counts[xx] = xmin, xmax
    return xmins, counts, kk


def _resample_axis(img: np.ndarray, out_size: int) -> np.ndarray:
    """One 8-bit pass along the last axis."""
    xmins, counts, kk = lanczos_coeffs(img.shape[1], out_size)
    out = np.empty((img.shape[0], out_size), dtype=np.uint8)
    src = img.astype(np.int64)
    for xx in range(out_size):
        n = counts[xx]
        acc = (1 << (PRECISION_BITS - 1)) + (src[:, xmins[xx]:xmins[xx] + n] * kk[xx, :n].astype(np.int64)).sum(axis=1)
        out[:, xx] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return out


def lanczos_resize(img: np.ndarray, new_h: int, new_w: int) -> np.ndarray:
    """Image.resize((new_w, new_h), Image.LANCZOS) for mode "L": horizontal pass, then vertical pass, 8-bit in between."""
    out = img
    if new_w != img.shape[1]:
        out = _resample_axis(out, new_w)
    if new_h != img.shape[0]:
        out = _resample_axis(np.ascontiguousarray(out.T), new_h).T
    return np.ascontiguousarray(out)


# ---------------------------------------------------------------------------------------------------------------------
# data_utils.py::minmax_size (62-82) with the intended get_divisible_size (50-59)
# ---------------------------------------------------------------------------------------------------------------------
def divisible_size(h: float, w: float, max_dimension: Sequence[int], factor: int = 32) -> Tuple[int, int]:
    out = []
    for v, cap in ((h, max_dimension[0]), (w, max_dimension[1])):
        n = v
        if v % factor:
            n = math.ceil(v / factor) * factor
            if n > cap:
                n = math.floor(v / factor) * factor
        out.append(int(n))
    return out[0], out[1]


def minmax_plan(h: int, w: int, max_dimensions: Optional[Sequence[int]], min_dimensions: Optional[Sequence[int]]):
    """The size decisions of minmax_size: ((resize_h, resize_w) or None, (canvas_h, canvas_w) or None)."""
    resize = canvas = None
    if max_dimensions is not None:
        ratios = [h / max_dimensions[0], w / max_dimensions[1]]
        if any(r > 1 for r in ratios):
            size_w, size_h = np.array([w, h]) / max(ratios)
            resize = divisible_size(size_h, size_w, max_dimensions)
            h, w = resize
    if min_dimensions is not None:
        ratios = [h / min_dimensions[0], w / min_dimensions[1]]
        if any(r < 1 for r in ratios):
            canvas = divisible_size(h / min(ratios), w / min(ratios), max_dimensions)
    return resize, canvas


def minmax_size(img: np.ndarray, max_dimensions, min_dimensions) -> np.ndarray:
    resize, canvas = minmax_plan(img.shape[0], img.shape[1], max_dimensions, min_dimensions)
    if resize is not None:
        img = lanczos_resize(img, resize[0], resize[1])
    if canvas is not None:
        # Image.new("L", size, 255) + paste(img, img.getbbox()): the image at its own non-zero bounding box = top-left when its
        # first row / column hold a non-zero pixel (a blank first row or column shifts nothing: getbbox then starts later and the
        # reference's paste raises on the size mismatch)
        out = np.full(canvas, 255, dtype=np.uint8)
        hh, ww = min(canvas[0], img.shape[0]), min(canvas[1], img.shape[1])
        out[:hh, :ww] = img[:hh, :ww]
        img = out
    return img


def normalize(img: np.ndarray, mean: float = 0.5, std: float = 0.5) -> np.ndarray:
    """albumentations Normalize (max_pixel_value 255) in float32: (v - mean*255) * (1 / (std*255))."""
    x = img.astype(np.float32)
    x -= np.float32(mean * 255.0)
    x *= np.float32(1.0 / (std * 255.0))
    return x


def preprocess(img: np.ndarray, opt: dict) -> np.ndarray:
    """predict_utils.py::resize for one grey image (imgH None, no resizer): float32 (1, 1, H, W)."""
    ratio = opt.get("downsample")
    if ratio is not None:
        h, w = img.shape
        if h / ratio >= opt["min_dimension"][0] and w / ratio >= opt["min_dimension"][1]:
            img = area_downsample(img, int(ratio))
    if opt.get("pad"):
        img, _ = pad_to_ink(img)
    img = minmax_size(img, opt["max_dimension"], opt["min_dimension"])
    return normalize(img, opt["mean"], opt["std"])[None, None]

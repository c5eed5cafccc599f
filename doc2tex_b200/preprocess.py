"""Image preprocessing of the recognizer on the B200 — host side.

Mirrors ``doc2tex/utils/predict_utils.py::resize`` (lines 14-115, the ``imgH is None`` / no-resizer branch the shipped
configs take) together with ``utils/data_utils.py::pad`` (10-47) and ``minmax_size`` (62-82), but for a LIST of grey crops
at once — the shape of the work ``demo/app.py:170-180`` produces (a page → many differently sized formula crops):

    prep = Preprocessor(engine, opt)                 # opt: the reference's keys (max_dimension, min_dimension, mean, std,
    buckets = prep(images)                           #      pad, downsample, rgb)
    for (H, W), (batch, index) in buckets.items():   # batch (B, 1, H, W) fp32 on the device, index = positions in `images`
        ...

The pixel work (down-sampling, min-max stretch, ink box, crop, Pillow-exact LANCZOS shrink, canvas, normalisation) runs in
hand-written kernels behind the C ABI (``d2t_prep_measure`` / ``d2t_prep_render``); the host only plans: it reads back the
ink boxes (8 ints per image), decides every image's output size exactly as ``minmax_size`` does, groups images of equal
(H, W) into buckets — the recognizer's batches are always same-sized images (torch_dataset.py:46-66, collate_fn.py:32) —
and builds the resampling tables.  There is no CPU fallback for the pixel work.

Two defects of the reference on this path are NOT reproduced: ``get_divisible_size`` leaves its result
unassigned when a scaled side is already a multiple of 32 (data_utils.py:50-59), so ``minmax_size`` raises UnboundLocalError
for most images that need resizing — here such a side simply stays; and an ink crop whose border row or column is entirely
zero makes ``padded.paste(im, im.getbbox())`` raise — here it raises a ``PreprocessError`` naming the image.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from .engine import Engine, EngineError

PRECISION_BITS = 32 - 8 - 2    # Pillow's 8-bit resampler keeps 22 fractional bits


class PreprocessError(EngineError):
    pass


def _divisible(v: float, cap: int, factor: int = 32) -> int:
    """One side of data_utils.py::get_divisible_size (50-59): up to the next multiple of 32, down if that exceeds the cap;
    a side that already is a multiple stays (the reference leaves it unassigned and crashes)."""
    if v % factor:
        n = math.ceil(v / factor) * factor
        return int(math.floor(v / factor) * factor if n > cap else n)
    return int(v)


def plan_sizes(h: int, w: int, max_dim: Optional[Sequence[int]], min_dim: Optional[Sequence[int]]):
    """minmax_size's decisions (data_utils.py:62-82): (shrink to (h, w) or None, canvas (h, w) or None)."""
    shrink = canvas = None
    if max_dim is not None:
        ratios = [h / max_dim[0], w / max_dim[1]]
        if any(r > 1 for r in ratios):
            size = np.array([w, h]) / max(ratios)
            shrink = (_divisible(size[1], max_dim[0]), _divisible(size[0], max_dim[1]))
            h, w = shrink
    if min_dim is not None:
        ratios = [h / min_dim[0], w / min_dim[1]]
        if any(r < 1 for r in ratios):
            canvas = (_divisible(h / min(ratios), max_dim[0]), _divisible(w / min(ratios), max_dim[1]))
    return shrink, canvas


def _lanczos(x: float) -> float:
    if -3.0 <= x < 3.0:
        if x == 0.0:
            return 1.0
        a, b = x * math.pi, x / 3.0 * math.pi
        return (math.sin(a) / a) * (math.sin(b) / b)
    return 0.0


def lanczos_table(in_size: int, out_size: int) -> Tuple[np.ndarray, int]:
    """Fixed-point taps of Pillow's LANCZOS filter for one axis (Resample.c precompute_coeffs + normalize_coeffs_8bpc):
    int32 rows [first source index, tap count, k[0..ksize)].  Built on the host in float64 like Pillow does."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 3.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    ss = 1.0 / filterscale
    tab = np.zeros((out_size, ksize + 2), dtype=np.int32)
    one = float(1 << PRECISION_BITS)
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        first = max(int(center - support + 0.5), 0)
        count = min(int(center + support + 0.5), in_size) - first
        w = [_lanczos((x + first - center + 0.5) * ss) for x in range(count)]
        total = 0.0
        for v in w:
            total += v
        tab[xx, 0], tab[xx, 1] = first, count
        for x, v in enumerate(w):
            if total != 0.0:
                v = v / total
            tab[xx, 2 + x] = int(-0.5 + v * one) if v < 0 else int(0.5 + v * one)
    return tab, ksize


class Preprocessor:
    def __init__(self, engine: Engine, opt: dict):
        if opt.get("rgb", False):
            raise PreprocessError("the recognizer is grey-scale (rgb: False in every shipped config)")
        if opt.get("imgH") is not None:
            raise PreprocessError("fixed-height preprocessing (imgH set) is the other branch of predict_utils.resize; "
                                  "the shipped configs leave imgH None")
        self.eng = engine
        self.lib = engine.lib
        self.max_dim = list(opt["max_dimension"]) if opt.get("max_dimension") is not None else None
        self.min_dim = list(opt["min_dimension"]) if opt.get("min_dimension") is not None else None
        self.pad = bool(opt.get("pad", False))
        ds = opt.get("downsample")
        self.downsample = None if ds is None else int(ds)
        if ds is not None and self.downsample != ds:
            raise PreprocessError("only integer down-sampling ratios (cv2.INTER_AREA with an integer scale) are supported")
        mean, std = float(opt.get("mean", 0.5)), float(opt.get("std", 0.5))
        self.sub = float(np.float32(mean * 255.0))
        self.mul = float(np.float32(1.0 / (std * 255.0)))

    # ------------------------------------------------------------------------------------------------------------
    def __call__(self, images: List[np.ndarray]) -> Dict[Tuple[int, int], Tuple[torch.Tensor, List[int]]]:
        dev = self.eng.device
        n = len(images)
        if n == 0:
            return {}
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        # ---- pack the crops into one pinned buffer: ONE host-to-device copy for the whole list ----
        imgs = (_lib.PrepImage * n)()
        off = 0
        for i, a in enumerate(images):
            if a.dtype != np.uint8 or a.ndim != 2:
                raise PreprocessError(f"image {i}: expected a 2-D uint8 (grey) array, got {a.dtype} {a.shape}")
            h, w = a.shape
            ds = 1
            if self.downsample is not None and self.downsample > 1:
                r = self.downsample     # predict_utils.py:33-38: only when the result stays above min_dimension
                if self.min_dim is None or (h / r >= self.min_dim[0] and w / r >= self.min_dim[1]):
                    if h % r or w % r:
                        raise PreprocessError(f"image {i}: {h}x{w} is not divisible by the down-sampling ratio {r} "
                                              "(fractional INTER_AREA is not implemented)")
                    ds = r
            imgs[i].src_off, imgs[i].h0, imgs[i].w0, imgs[i].ds = off, h, w, ds
            off += h * w
        packed = torch.empty(off, dtype=torch.uint8).pin_memory()
        pn = packed.numpy()
        for i, a in enumerate(images):
            pn[imgs[i].src_off: imgs[i].src_off + a.size] = a.reshape(-1)
        packed_dev = packed.to(dev, non_blocking=True)
        imgs_dev = self._upload(bytes(imgs), dev)
        # ---- measure: min, polarity, ink box per image (needed only for the crop-to-ink) ----
        sizes = [(imgs[i].h0 // imgs[i].ds, imgs[i].w0 // imgs[i].ds) for i in range(n)]
        stats = None
        if self.pad:
            stats_dev = torch.empty(n, 8, dtype=torch.int32, device=dev)
            self.eng._check(self.lib.d2t_prep_measure(self.eng.h, packed_dev.data_ptr(), imgs_dev.data_ptr(), n,
                                                      stats_dev.data_ptr(), stream), "d2t_prep_measure")
            stats = stats_dev.cpu().numpy()    # 32 bytes per image; the sizes of the outputs depend on it
        # ---- plan: every image's stages and output size, buckets by exact (H, W) ----
        plans = (_lib.PrepPlan * n)()
        tables: List[np.ndarray] = []
        table_cache: Dict[Tuple[int, int], Tuple[int, int]] = {}
        coef_len = 0
        scratch = 0
        any_resize = False
        buckets: Dict[Tuple[int, int], List[int]] = {}
        for i in range(n):
            p = plans[i]
            h, w = sizes[i]
            if self.pad:
                x, y, cw, ch, inv, vmin, status, _ = (int(v) for v in stats[i])
                if status == 1:
                    raise PreprocessError(f"image {i}: blank image, no ink to crop to")
                if status == 2:
                    raise PreprocessError(f"image {i}: no pixel crosses the ink threshold")
                p.use_crop, p.crop_x, p.crop_y, p.crop_w, p.crop_h, p.inverted, p.vmin = 1, x, y, cw, ch, inv, vmin
                h, w = 32 * ((ch + 31) // 32), 32 * ((cw + 31) // 32)     # data_utils.py:38-43
            p.hb, p.wb = h, w
            p.off_b = scratch
            scratch += h * w
            shrink, canvas = plan_sizes(h, w, self.max_dim, self.min_dim)
            if shrink is not None:
                any_resize = True
                p.do_resize, p.rh, p.rw = 1, shrink[0], shrink[1]
                for axis, (src_n, dst_n) in enumerate(((w, shrink[1]), (h, shrink[0]))):
                    key = (src_n, dst_n)
                    if key not in table_cache:
                        tab, ksize = lanczos_table(src_n, dst_n)
                        table_cache[key] = (coef_len, ksize)
                        tables.append(tab.reshape(-1))
                        coef_len += tab.size
                    o, ks = table_cache[key]
                    if axis == 0:
                        p.kx_off, p.kx_ksize = o, ks
                    else:
                        p.ky_off, p.ky_ksize = o, ks
                p.off_t = scratch
                scratch += h * shrink[1]
                p.off_r = scratch
                scratch += shrink[0] * shrink[1]
                h, w = shrink
            if canvas is not None:
                h, w = canvas
            if h % 32 or w % 32:
                raise PreprocessError(f"image {i}: preprocessed size {h}x{w} is not a multiple of 32 — the recognizer needs /32 "
                                      "sizes (data_utils.py:10-47); enable `pad` or feed /32 crops")
            p.out_h, p.out_w = h, w
            buckets.setdefault((h, w), []).append(i)
        out: Dict[Tuple[int, int], Tuple[torch.Tensor, List[int]]] = {}
        for (h, w), idx in buckets.items():
            t = torch.empty(len(idx), 1, h, w, dtype=torch.float32, device=dev)
            for slot, i in enumerate(idx):
                plans[i].dst = t.data_ptr() + slot * h * w * 4
            out[(h, w)] = (t, idx)
        plans_dev = self._upload(bytes(plans), dev)
        coefs_dev = self._upload(np.concatenate(tables).tobytes(), dev) if tables else None
        scratch_dev = torch.empty(max(scratch, 1), dtype=torch.uint8, device=dev)
        self.eng._check(self.lib.d2t_prep_render(self.eng.h, packed_dev.data_ptr(), imgs_dev.data_ptr(), plans_dev.data_ptr(), n,
                                                 coefs_dev.data_ptr() if coefs_dev is not None else None, scratch_dev.data_ptr(),
                                                 1 if any_resize else 0, self.sub, self.mul, stream), "d2t_prep_render")
        # the kernels are stream-ordered after the uploads; keep the temporaries alive until the stream is past them
        for t in (packed_dev, imgs_dev, plans_dev, scratch_dev) + ((coefs_dev,) if coefs_dev is not None else ()):
            t.record_stream(torch.cuda.current_stream(dev))
        self.last_boxes = None if stats is None else stats[:, :4].copy()
        return out

    @staticmethod
    def _upload(raw: bytes, dev) -> torch.Tensor:
        host = torch.frombuffer(bytearray(raw), dtype=torch.uint8).pin_memory()
        return host.to(dev, non_blocking=True)

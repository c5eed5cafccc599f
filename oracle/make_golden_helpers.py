"""Seeded synthetic inputs shared by the golden generator and the tests (test infrastructure)."""
import numpy as np


def synth_crop(h, w, seed, dark_ink=True, margin=(5, 9, 7, 11)):
    """A formula-crop-shaped grey image: noisy paper, sparse ink strokes, blank margins (top, bottom, left, right)."""
    r = np.random.default_rng(seed)
    img = np.full((h, w), 250 if dark_ink else 12, dtype=np.uint8) + r.integers(0, 5, (h, w)).astype(np.uint8)
    t, b, l, rr = margin
    ink = r.random((h, w)) < 0.06
    ink[:t] = False; ink[h - b:] = False; ink[:, :l] = False; ink[:, w - rr:] = False
    vals = r.integers(0, 60, (h, w)).astype(np.uint8) if dark_ink else r.integers(200, 255, (h, w)).astype(np.uint8)
    img[ink] = vals[ink]
    return img

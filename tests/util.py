"""Shared helpers for the parity tests."""
import os

import numpy as np
import torch

from doc2tex_b200 import synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# fp32 mode gate (BASELINE.json north_star): logits / ctx within 1e-3 relative to the reference's max-abs.
REL_TOL_FP32 = 1e-3


def load_golden(name):
    return dict(np.load(os.path.join(GOLD, name + ".npz")))


def end_bias_of(g):
    v = float(g["end_bias"])
    return None if np.isnan(v) else v


def rel_err(a: torch.Tensor, ref: torch.Tensor) -> float:
    return ((a.double() - ref.double()).abs().max() / ref.double().abs().max().clamp_min(1e-30)).item()


_SD_CACHE = {}


def state_dict_for(head, end_bias, sharpen=1.0):
    key = (head, end_bias, sharpen)
    if key not in _SD_CACHE:
        cfg = synth.make_config(head)
        _SD_CACHE[key] = (cfg, synth.make_state_dict(cfg, seed=1111, end_bias=end_bias, sharpen=sharpen))
    return _SD_CACHE[key]


def sharpen_of(g):
    return float(g["sharpen"]) if "sharpen" in g else 1.0


def golden_images(g, n, H=64, W=256):
    """The images a fixture was minted on: per-image seeds when the fixture names them, else the seed-2024 batch."""
    if "img_seeds" in g:
        return torch.cat([synth.make_images(1, H, W, seed=int(s)) for s in g["img_seeds"][:n]], 0)
    return synth.make_images(n, H, W, seed=2024)


# A beam decision whose candidates are separated by fewer than this many fp32 ulps of the cumulative score is a NEAR-TIE:
# any other fp32 summation order may legitimately flip it.  Decisions with a wider margin must be identical.
NEAR_TIE_ULPS = 32.0


def decision_margins(trace, trace_score, runner_up):
    """Margin of every beam decision: the smallest gap between neighbours among the k selected candidates and the best one
    left out.  trace (B,T,k,2) int (-1 = unused), trace_score (B,T,k) fp32 in top-k order, runner_up (B,T) fp32 = best
    candidate NOT selected (-inf when there is none).  Returns (gap, ulp): two (B,T) float64 arrays — the absolute gap
    (+inf where no decision was taken: image finished / a single candidate) and the fp32 spacing at the scores' magnitude."""
    tr = np.asarray(trace); ts = np.asarray(trace_score, dtype=np.float64); ru = np.asarray(runner_up, dtype=np.float64)
    B, T, K = ts.shape
    gap = np.full((B, T), np.inf)
    ulp = np.full((B, T), 1.0)
    for b in range(B):
        for t in range(T):
            k = int((tr[b, t, :, 0] >= 0).sum())
            if k == 0:
                continue
            vals = list(ts[b, t, :k])
            if np.isfinite(ru[b, t]):
                vals.append(ru[b, t])
            if len(vals) < 2:
                continue
            v = np.array(vals)
            ulp[b, t] = float(np.spacing(np.float32(np.abs(v).max())))
            gap[b, t] = float((v[:-1] - v[1:]).min())
    return gap, ulp

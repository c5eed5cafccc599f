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


def state_dict_for(head, end_bias):
    key = (head, end_bias)
    if key not in _SD_CACHE:
        cfg = synth.make_config(head)
        _SD_CACHE[key] = (cfg, synth.make_state_dict(cfg, seed=1111, end_bias=end_bias))
    return _SD_CACHE[key]

"""Small driver for ncu: one encode + a short decode of the bench workload (no CUDA graph so every launch is listed).

    python tools/profile_path.py --batch 256 --steps 4 --mode greedy --precision bf16x3
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from doc2tex_b200 import synth  # noqa: E402
from doc2tex_b200.engine import Engine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--steps", type=int, default=4)
ap.add_argument("--mode", default="greedy")
ap.add_argument("--beam", type=int, default=5)
ap.add_argument("--precision", default="bf16x3")
ap.add_argument("--height", type=int, default=64)
ap.add_argument("--width", type=int, default=256)
ap.add_argument("--warm", type=int, default=1)
ap.add_argument("--graphs", action="store_true")
ap.add_argument("--head", default="TFM", choices=["TFM", "Attnv2"])
ap.add_argument("--prep", type=int, default=0, help="also run the GPU preprocessing on this many synthetic crops")
ap.add_argument("--opt", action="append", default=[], help="engine option key=value (repeatable)")
ap.add_argument("--images", type=int, default=0, help="decode this many images (ctx repeated) instead of --batch")
a = ap.parse_args()

cfg = synth.make_config(a.head)
sd = synth.make_state_dict(cfg, seed=1111, suppress_end=True)
eng = Engine(cfg, "cuda:0", precision=a.precision, use_graphs=a.graphs)
eng.load_state_dict(sd)
for kv in a.opt:
    k, v = kv.split("=")
    eng.set_option(k, int(v))
img = synth.make_images(a.batch, a.height, a.width, seed=2024).cuda()
for i in range(a.warm + 1):
    l0 = eng.launch_count()
    ctx, _, _ = eng.encode(img)
    if a.images > 0:
        ctx = ctx.repeat((a.images + a.batch - 1) // a.batch, 1, 1)[:a.images].contiguous()
    l1 = eng.launch_count()
    if a.mode == "greedy":
        eng.decode_greedy(ctx, a.steps, is_test=True, return_logits=False)
    else:
        eng.decode_beam(ctx, a.beam, a.steps)
    torch.cuda.synchronize()
    if a.prep:
        import numpy as np
        from doc2tex_b200.preprocess import Preprocessor
        rng = np.random.default_rng(0)
        crops = []
        for k in range(a.prep):
            h, w = int(rng.integers(40, 200)), int(rng.integers(150, 900))
            im = np.full((h, w), 250, dtype=np.uint8)
            im[8:h - 8, 10:w - 10][rng.random((h - 16, w - 20)) < 0.06] = 20
            crops.append(im)
        opt = {"max_dimension": [192, 896], "min_dimension": [32, 32], "mean": 0.5, "std": 0.5, "rgb": False, "imgH": None,
               "pad": True, "downsample": None}
        out = Preprocessor(eng, opt)(crops)
        torch.cuda.synchronize()
        print(f"preprocessed {a.prep} crops into {len(out)} buckets")
    print(f"pass {i}: encode launches {l1 - l0}, decode launches {eng.launch_count() - l1}")

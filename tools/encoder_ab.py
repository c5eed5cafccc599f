"""Encoder time under engine options (CUDA events, batch 256, 64x256).   python tools/encoder_ab.py key=v0,v1 [precision ...]"""
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from doc2tex_b200 import synth  # noqa: E402
from doc2tex_b200.engine import Engine  # noqa: E402

key, vals = sys.argv[1].split("=")
vals = [int(v) for v in vals.split(",")]
precs = sys.argv[2:] or ["bf16x3", "bf16"]
cfg = synth.make_config("TFM")
sd = synth.make_state_dict(cfg, seed=1111, suppress_end=True)
img = synth.make_images(256, 64, 256, seed=2024).cuda()
for prec in precs:
    eng = Engine(cfg, "cuda:0", precision=prec)
    eng.load_state_dict(sd)
    res = {v: [] for v in vals}
    ref = None
    worst = 0.0
    for rnd in range(4):
        for v in vals:
            eng.set_option(key, v)
            ctx, _, _ = eng.encode(img)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                ctx, _, _ = eng.encode(img)
            e1.record()
            torch.cuda.synchronize()
            res[v].append(e0.elapsed_time(e1) / 3)
            if ref is None:
                ref = ctx.clone()
            else:
                d = float((ctx - ref).abs().max() / ref.abs().max())
                # options that only move data are bit-equal; options that change a summation order (vit_planes) stay far inside
                # the fp32-parity tolerance of the mode (1e-3)
                assert d < 2e-4, f"{key}={v}: encoder output differs from {key}={vals[0]} by {d:.2e}"
                worst = max(worst, d)
    print(f"{prec}: " + ", ".join(f"{key}={v}: {statistics.median(res[v]):.2f} ms" for v in vals) + f"  (max rel diff {worst:.1e})", flush=True)
    eng.close()

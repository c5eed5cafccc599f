"""Decode duration under one engine option (round robin, medians).   python tools/option_ab.py key=v0,v1 [precision]"""
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from doc2tex_b200 import synth  # noqa: E402
from doc2tex_b200.engine import Engine  # noqa: E402

key, vals = sys.argv[1].split("=")
vals = [int(v) for v in vals.split(",")]
prec = sys.argv[2] if len(sys.argv) > 2 else "bf16x3"
cfg = synth.make_config("TFM")
sd = synth.make_state_dict(cfg, seed=1111, suppress_end=True)
eng = Engine(cfg, "cuda:0", precision=prec)
eng.load_state_dict(sd)
ctx, _, _ = eng.encode(synth.make_images(256, 64, 256, seed=2024).cuda())
WORK = [("beam", 32), ("beam", 256), ("beam", 1280), ("greedy", 256), ("greedy", 2560)]
res = {}
for rnd in range(3):
    for v in vals:
        eng.set_option(key, v)
        for mode, n in WORK:
            c = ctx.repeat((n + 255) // 256, 1, 1)[:n].contiguous()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            if mode == "greedy":
                eng.decode_greedy(c, 151, is_test=True, return_logits=False)
            else:
                eng.decode_beam(c, 5, 151)
            e1.record()
            torch.cuda.synchronize()
            res.setdefault((v, mode, n), []).append(e0.elapsed_time(e1))
for mode, n in WORK:
    print(f"{prec} {mode} {n} images: " + " | ".join(f"{key}={v} {1e3 * statistics.median(res[(v, mode, n)]) / 151:7.1f} us/step" for v in vals), flush=True)

"""Decode duration under option attn_kpi (keys in flight per quarter warp).   python tools/attn_kpi_ab.py [precision]"""
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from doc2tex_b200 import synth  # noqa: E402
from doc2tex_b200.engine import Engine  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
cfg = synth.make_config("TFM")
sd = synth.make_state_dict(cfg, seed=1111, suppress_end=True)
eng = Engine(cfg, "cuda:0", precision=prec)
eng.load_state_dict(sd)
ctx, _, _ = eng.encode(synth.make_images(256, 64, 256, seed=2024).cuda())
WORK = [("beam", 256), ("beam", 1024), ("greedy", 256), ("greedy", 2560)]
VALS = (4, 2)
res = {}
for rnd in range(3):
    for kpi in VALS:
        eng.set_option("attn_kpi", kpi)
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
            res.setdefault((kpi, mode, n), []).append(e0.elapsed_time(e1))
for mode, n in WORK:
    print(f"{prec} {mode} {n} images: " + " | ".join(f"attn_kpi {k} {1e3 * statistics.median(res[(k, mode, n)]) / 151:7.1f} us/step" for k in VALS), flush=True)

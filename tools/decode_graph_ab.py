"""Decode duration with and without CUDA graphs (Engine(use_graphs=...)) at the row counts the schedule runs.
    python tools/decode_graph_ab.py [precision]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from doc2tex_b200 import synth  # noqa: E402
from doc2tex_b200.engine import Engine  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
cfg = synth.make_config("TFM")
sd = synth.make_state_dict(cfg, seed=1111, suppress_end=True)
img = synth.make_images(256, 64, 256, seed=2024).cuda()
engs = {}
for g in (True, False):
    engs[g] = Engine(cfg, "cuda:0", precision=prec, use_graphs=g)
    engs[g].load_state_dict(sd)
ctx, _, _ = engs[True].encode(img)
for mode, n, steps in (("greedy", 256, 151), ("greedy", 2560, 151), ("greedy", 2560, 100), ("beam", 256, 151), ("beam", 1024, 151), ("beam", 1024, 100)):
    c = ctx.repeat((n + 255) // 256, 1, 1)[:n].contiguous()
    out = []
    for g in (True, False):
        eng = engs[g]
        best = 1e9
        for i in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            if mode == "greedy":
                eng.decode_greedy(c, steps, is_test=True, return_logits=False)
            else:
                eng.decode_beam(c, 5, steps)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        out.append(f"{'graphs' if g else 'eager '} {best:7.1f} ms = {1e3 * best / steps:7.1f} us/step")
    print(f"{mode} {n} images, {steps} steps: " + " | ".join(out), flush=True)

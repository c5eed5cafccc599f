"""Duration of a decode call against its number of steps: separates the per-call cost (allocation, cross K/V projection,
result copies) from the marginal cost of a step.    python tools/decode_steps_fit.py [precision]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from doc2tex_b200 import synth  # noqa: E402
from doc2tex_b200.engine import Engine  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
cfg = synth.make_config("TFM")
sd = synth.make_state_dict(cfg, seed=1111, suppress_end=True)
img = synth.make_images(256, 64, 256, seed=2024).cuda()
eng = Engine(cfg, "cuda:0", precision=prec)
eng.load_state_dict(sd)
ctx, _, _ = eng.encode(img)
for mode, n in (("greedy", 256), ("greedy", 2560), ("beam", 1024)):
    c = ctx.repeat((n + 255) // 256, 1, 1)[:n].contiguous()
    prev = None
    for steps in (1, 9, 17, 33, 65, 97, 129, 151):
        best, wall = 1e9, 1e9
        for i in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            e0.record()
            if mode == "greedy":
                eng.decode_greedy(c, steps, is_test=True, return_logits=False)
            else:
                eng.decode_beam(c, 5, steps)
            e1.record()
            torch.cuda.synchronize()
            wall = min(wall, 1e3 * (time.perf_counter() - t0))
            best = min(best, e0.elapsed_time(e1))
        marg = "" if prev is None else f"  marginal {1e3 * (best - prev[1]) / (steps - prev[0]):7.1f} us/step"
        print(f"{mode} {n} images, {steps:3d} steps: {best:7.2f} ms (wall {wall:7.2f}){marg}", flush=True)
        prev = (steps, best)

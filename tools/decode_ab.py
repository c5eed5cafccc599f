"""A/B timing of the decode step under engine options (us per step, CUDA-event timed decode calls of 151 steps).

    python tools/decode_ab.py [--precision bf16x3] [--timeline]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from doc2tex_b200 import synth  # noqa: E402
from doc2tex_b200.engine import Engine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--precision", default="bf16x3")
ap.add_argument("--timeline", action="store_true")
ap.add_argument("--sets", default="")
a = ap.parse_args()

cfg = synth.make_config("TFM")
sd = synth.make_state_dict(cfg, seed=1111, suppress_end=True)
eng = Engine(cfg, "cuda:0", precision=a.precision)
eng.load_state_dict(sd)
img = synth.make_images(256, 64, 256, seed=2024).cuda()
ctx, _, _ = eng.encode(img)

BASE = {"stack_mma": 1, "steps_per_graph": 8, "attn_split": 0, "attn_kpi": 4, "split_k": 1, "pdl": 1, "pdl_max_rows": 2048}
SETS = [
    ("default", {}),
    ("attn_kpi 2", {"attn_kpi": 2}),
    ("attn_kpi 8", {"attn_kpi": 8}),
    ("stack_mma 0", {"stack_mma": 0}),
    ("1 step per graph", {"steps_per_graph": 1}),
    ("split_k 0", {"split_k": 0}),
    ("pdl 0", {"pdl": 0}),
    ("pdl at any size", {"pdl_max_rows": 1 << 30}),
]
WORK = [("greedy", 32), ("greedy", 256), ("greedy", 1024), ("greedy", 2048), ("beam", 32), ("beam", 256), ("beam", 512), ("beam", 1024)]


def run(mode, n):
    c = ctx.repeat((n + 255) // 256, 1, 1)[:n].contiguous()
    best = 1e9
    for i in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if mode == "greedy":
            eng.decode_greedy(c, 151, is_test=True, return_logits=False)
        else:
            eng.decode_beam(c, 5, 151)
        e1.record()
        torch.cuda.synchronize()
        if i > 0 or True:
            best = min(best, e0.elapsed_time(e1))
    return best


import statistics

ROUNDS = 3
res = {name: {w: [] for w in WORK} for name, _ in SETS}
for rnd in range(ROUNDS):      # round robin over the option sets: clock / thermal drift hits every set alike; medians reported
    for name, opts in SETS:
        o = dict(BASE)
        o.update(opts)
        for k, v in o.items():
            eng.set_option(k, v)
        for w in WORK:
            res[name][w].append(run(*w))
for name, _ in SETS:
    print(f"{name:<34}" + " ".join(f"{m[0]}{n * (5 if m == 'beam' else 1)}r={1e3 * statistics.median(res[name][(m, n)]) / 151:7.1f}"
                                   for m, n in WORK), flush=True)

if a.timeline:
    for k, v in dict(BASE, steps_per_graph=1).items():
        eng.set_option(k, v)
    eng.set_option("dbg_timeline", 1)
    for mode, n in (("greedy", 256), ("greedy", 1024), ("beam", 256)):
        print(f"--- timeline {mode} {n} images", file=sys.stderr, flush=True)
        c = ctx.repeat((n + 255) // 256, 1, 1)[:n].contiguous()
        if mode == "greedy":
            eng.decode_greedy(c, 100, is_test=True, return_logits=False)
        else:
            eng.decode_beam(c, 5, 100)
    eng.set_option("dbg_timeline", 0)

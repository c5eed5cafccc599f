"""Per-launch timeline of one decode step (option dbg_timeline; printed by the engine on stderr) at the row counts the
pipelined schedule runs.     python tools/decode_timeline.py [precision] [mode:images ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from doc2tex_b200 import synth  # noqa: E402
from doc2tex_b200.engine import Engine  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
work = [w.split(":") for w in sys.argv[2:]] or [("greedy", "2560"), ("beam", "1024")]
cfg = synth.make_config("TFM")
sd = synth.make_state_dict(cfg, seed=1111, suppress_end=True)
eng = Engine(cfg, "cuda:0", precision=prec)
eng.load_state_dict(sd)
img = synth.make_images(256, 64, 256, seed=2024).cuda()
ctx, _, _ = eng.encode(img)
eng.set_option("steps_per_graph", 1)
eng.set_option("dbg_timeline", 1)
for mode, n in work:
    n = int(n)
    print(f"--- timeline {mode} {n} images ({prec})", file=sys.stderr, flush=True)
    c = ctx.repeat((n + 255) // 256, 1, 1)[:n].contiguous()
    if mode == "greedy":
        eng.decode_greedy(c, 100, is_test=True, return_logits=False)
    else:
        eng.decode_beam(c, 5, 100)
    torch.cuda.synchronize()

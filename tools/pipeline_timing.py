"""Per-batch encode / decode durations inside the pipelined schedule (interference check)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from doc2tex_b200 import synth
from doc2tex_b200.engine import Engine
from doc2tex_b200.pipeline import PipelinedRecognizer

sms = int(sys.argv[1]) if len(sys.argv) > 1 else 112
mode = sys.argv[2] if len(sys.argv) > 2 else "greedy"
cfg = synth.make_config("TFM")
sd = synth.make_state_dict(cfg, seed=1111, suppress_end=True)
eng = Engine(cfg, "cuda:0", precision="bf16x3")
eng.load_state_dict(sd)
img = synth.make_images(256, 64, 256, seed=2024).cuda()
pipe = PipelinedRecognizer(eng, mode, 5, 151, encoder_sms=sms)
list(pipe.run([img] * 3))
pipe.timing = []
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
list(pipe.run([img] * 8))
torch.cuda.synchronize()
print(f"sms={sms} {mode}: total {1e3 * (time.perf_counter() - t0) / 8:.1f} ms/batch")
for e, d in pipe.timing:
    print(f"  encode {e:6.1f} ms   decode {d:6.1f} ms")

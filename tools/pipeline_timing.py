"""Per-batch encode / decode durations inside the pipelined schedule (interference check).

    python tools/pipeline_timing.py SMS MODE MERGE [PRECISION]
    python tools/pipeline_timing.py rows            # decode duration vs rows of one call (no encoder running)
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from doc2tex_b200 import synth
from doc2tex_b200.engine import Engine
from doc2tex_b200.pipeline import PipelinedRecognizer

prec = sys.argv[4] if len(sys.argv) > 4 else "bf16x3"
cfg = synth.make_config("TFM")
sd = synth.make_state_dict(cfg, seed=1111, suppress_end=True)
eng = Engine(cfg, "cuda:0", precision=prec)
eng.load_state_dict(sd)
img = synth.make_images(256, 64, 256, seed=2024).cuda()

if sys.argv[1] == "rows":
    ctx, _, _ = eng.encode(img)
    for mode, sizes in (("greedy", (64, 128, 256, 512, 1024, 2048)), ("beam", (32, 64, 128, 256, 512))):
        for n in sizes:
            c = ctx.repeat((n + 255) // 256, 1, 1)[:n].contiguous()
            ts = []
            for i in range(3):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                if mode == "greedy":
                    eng.decode_greedy(c, 151, is_test=True, return_logits=False)
                else:
                    eng.decode_beam(c, 5, 151)
                torch.cuda.synchronize()
                ts.append(1e3 * (time.perf_counter() - t0))
            print(f"{mode} images={n:5d} rows={n * (5 if mode == 'beam' else 1):5d}: decode {min(ts[1:]):7.1f} ms  "
                  f"({1e3 * min(ts[1:]) / 151:6.1f} us/step, {min(ts[1:]) * 256 / n:6.1f} ms per 256 images)", flush=True)
    sys.exit(0)

sms = int(sys.argv[1]) if len(sys.argv) > 1 else 112
mode = sys.argv[2] if len(sys.argv) > 2 else "greedy"
merge = int(sys.argv[3]) if len(sys.argv) > 3 else 1
pipe = PipelinedRecognizer(eng, mode, 5, 151, encoder_sms=sms, decode_merge=merge)
list(pipe.run([img] * (2 * merge)))
pipe.timing = []
torch.cuda.synchronize()
t0 = time.perf_counter()
list(pipe.run([img] * (4 * merge)))
torch.cuda.synchronize()
print(f"sms={sms} {mode} merge={merge} {prec}: total {1e3 * (time.perf_counter() - t0) / (4 * merge):.1f} ms/batch")
for e, d in pipe.timing:
    print(f"  encode {e:6.1f} ms   decode {d:6.1f} ms per batch")

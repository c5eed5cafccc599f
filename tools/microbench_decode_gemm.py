import os, sys
sys.path.insert(0, "/root/repo")
from doc2tex_b200 import synth
from doc2tex_b200.engine import Engine
eng = Engine(synth.make_config("TFM"), "cuda:0", precision="fp32")
for prec in ("fp32", "bf16x3", "bf16"):
    for (M, N, K) in [(256, 256, 64), (256, 256, 128), (256, 256, 256), (256, 256, 512), (256, 256, 1024), (256, 1024, 256), (256, 64, 256), (1280, 256, 1024)]:
        t = eng.gemm_bench(M, N, K, prec, 200)
        print(f"{prec:7s} M={M:6d} N={N:5d} K={K:5d}: {t:9.2f} us/launch", flush=True)

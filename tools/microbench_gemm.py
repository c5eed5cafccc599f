"""Per-launch latency / throughput of the contraction kernels on decode- and encoder-shaped problems."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from doc2tex_b200 import synth  # noqa: E402
from doc2tex_b200.engine import Engine  # noqa: E402

eng = Engine(synth.make_config("TFM"), "cuda:0", precision="fp32")
shapes = [(256, 256, 256), (256, 768, 256), (256, 1024, 256), (256, 256, 1024), (1280, 256, 256), (1280, 768, 256),
          (1280, 256, 1024), (17152, 768, 256), (133120, 512, 4608)]
for prec in ("fp32", "bf16x3", "bf16"):
    for (M, N, K) in shapes:
        if prec == "fp32" and M > 20000:
            continue
        it = 20 if M > 20000 else 200
        t = eng.gemm_bench(M, N, K, prec, it)
        t2 = eng.gemm_bench(M, N, K, prec, it, interleave=True) if N <= 1024 and N % 128 == 0 else float("nan")
        print(f"{prec:7s} M={M:6d} N={N:5d} K={K:5d}: {t:9.2f} us/launch  {2 * M * N * K / t / 1e6:9.2f} TFLOP/s   "
              f"with LN interleaved: {t2:9.2f} us", flush=True)

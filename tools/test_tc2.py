"""cta_group::2 vs single-CTA contraction kernel: timing, and the pure MMA issue rate (D2T_DBG_ACT=64: no operand waits)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from doc2tex_b200 import synth
from doc2tex_b200.engine import Engine
eng = Engine(synth.make_config("TFM"), "cuda:0", precision="fp32")
for prec in ("bf16x3", "bf16"):
    for tc2 in (0, 1):
        eng.set_option("tc2", tc2)
        t = eng.gemm_bench(133120, 512, 4608, prec, 10)
        print(f"act={os.environ.get('D2T_DBG_ACT')} {prec} tc2={tc2} M=133120 N=512 K=4608: {t:.1f} us  {2*133120*512*4608/t/1e6:.1f} TFLOP/s algorithmic", flush=True)

#!/usr/bin/env python
"""bench.py — formulas/sec of the recognizer hot path (encode + autoregressive decode).

    python bench.py --gpus N --steps K --warmup W              # the B200 engine
    python bench.py --impl reference --gpus N --steps K ...    # the reference's CPU implementation on host cores

One "step" = one pass of the hot path over one batch of synthetic images:
ResNet stem -> patch-embed -> ViT encoder -> greedy (or beam-5) decode of exactly 151 steps
(END suppressed = the deterministic "full-length" regime) -> one all-gather of the token ids.

The JSON line (rank 0, ONE line on stdout) is the record of BASELINE.json configs[1] — HybridViT greedy, batch 256 per
GPU, fp32-parity mode `bf16x3` — and carries, under "records", the other configurations BASELINE.json names, each with
its own value / ms_per_step / e2e / roofline measured the same way (fewer steps):
    beam5          configs[2]'s decode (beam-5, batch 256 per GPU) in the fp32-parity mode
    bf16_greedy, bf16_beam5    the single-pass bf16 mode (configs[2]: "bf16")
    strong_greedy, strong_beam5   (N > 1) ONE global batch of 256 sharded 256/N per rank (configs[2] "batch-sharded")
    attnv2_b512    configs[3]: the config/train.yaml default stack, greedy, global batch 512 sharded over the ranks
    natural_greedy, natural_beam5   the "natural" decode-length regime (unmodified END logit, the reference's early exit)
    sweep_<H>x<W>_b<B>   configs[4]: points of the beam-5 image-size x batch sweep (B images per GPU)
The schedule (top-level "schedule" of every record) is the pipelined recognizer: encode(i+1) overlaps decode(i), and
`decode_merge` encoded batches go to ONE decode call (auto_merge: about 2 560 greedy / 6 400 beam-5 rows per call).  A
sub-record that fails carries {"error": ...} instead of taking the line with it.
`--records none` prints the main record only; `--mode/--precision/...` change what the MAIN record measures.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from doc2tex_b200 import synth  # noqa: E402

ENC_GFLOP = {(64, 256): 51.52, (96, 384): 115.58, (128, 512): 205.28, (160, 704): 352.79, (192, 896): 539.31}
REF_DIR = os.path.join(ROOT, "baseline", "_ref")   # copy of the pure-Python reference (oracle/install_reference.py); git-ignored
T_STEPS = 151


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="images per GPU")
    ap.add_argument("--mode", default="greedy", choices=["greedy", "beam"])
    ap.add_argument("--beam", type=int, default=5)
    ap.add_argument("--precision", default="bf16x3", choices=["fp32", "tf32x3", "bf16x3", "bf16"],
                    help="bf16x3 = the fp32-parity mode of record on tensor cores (error-compensated 3-pass split)")
    ap.add_argument("--height", type=int, default=64)
    ap.add_argument("--width", type=int, default=256)
    ap.add_argument("--ref-batch", type=int, default=8, help="images per step of the CPU reference sample")
    ap.add_argument("--cpu-sample", type=int, default=8, help="images of the cpu_baseline sample (0 = skip)")
    ap.add_argument("--no-graphs", action="store_true")
    ap.add_argument("--head", default="TFM", choices=["TFM", "Attnv2"],
                    help="TFM = HybridViT + transformer decoder (configs 1,2,3,5); Attnv2 = config/train.yaml default stack (config 4)")
    ap.add_argument("--encoder-sms", type=int, default=0,
                    help="SMs given to the encoder's persistent kernels; the rest run the overlapped decode of the previous batch")
    ap.add_argument("--decode-groups", type=int, default=0, help="concurrent row groups of a decode call (0 = engine default)")
    ap.add_argument("--decode-merge", type=int, default=0,
                    help="encoded batches handed to one decode call by the pipelined schedule (0 = default for the mode)")
    ap.add_argument("--natural", action="store_true",
                    help="natural decode regime: unmodified END logit, early exit as in the reference (tfm.py:138-140) "
                         "instead of the deterministic full-length regime (SURVEY 8d)")
    ap.add_argument("--no-overlap", action="store_true",
                    help="encode decode_merge batches back to back on all SMs, then decode them in one call (no stage overlap)")
    ap.add_argument("--sequential", action="store_true", help="no encode/decode overlap across batches")
    ap.add_argument("--records", default="all",
                    help="'all', 'none' or a comma list of beam5,bf16_greedy,bf16_beam5,strong_greedy,strong_beam5,attnv2_b512,natural_greedy,natural_beam5,"
                         "sweep_<H>x<W>_b<images per GPU> (beam-5, BASELINE configs[4])")
    ap.add_argument("--record-steps", type=int, default=0, help="timed steps of each sub-record (0 = min(steps, 4))")
    ap.add_argument("--opt", action="append", default=[], help="engine option key=value (d2t_set_option), repeatable")
    a = ap.parse_args()
    if a.encoder_sms <= 0:
        a.encoder_sms = 132  # measured sweet spot with merged decode (112 / 88 without)
    return a


MERGE_TARGET_IMAGES = {"greedy": 2560, "beam": 1280}   # images per decode call: 2 560 greedy rows / 6 400 beam-5 rows


def merge_target(mode: str, batch: int) -> int:
    return max(1, MERGE_TARGET_IMAGES[mode] // max(1, batch))


def auto_merge(mode: str, steps: int, batch: int = 256) -> int:
    """Encoded batches handed to one decode call.  The decode step is a chain of dependent launches whose duration grows far
    slower than its rows until the attention walks are HBM-bound (profiles/r02c_decode_vs_rows.txt: 256 rows 34.4 ms, 1 024
    rows 16.5 ms, 2 048 rows 12.2 ms per 256 images), so consecutive batches are merged up to about 2 560 greedy rows /
    6 400 beam-5 rows per call (10 / 5 batches of 256 images; more batches when a rank holds a smaller shard).  The count
    divides the number of timed steps whenever it can: a partial group costs a whole decode chain inside the bracket
    (profiles/r02c_schedule_sweep*.txt)."""
    tgt = merge_target(mode, batch)
    if steps % tgt == 0:
        return tgt
    best = max(d for d in range(1, min(steps, tgt * 5 // 4) + 1) if steps % d == 0)
    return best if 2 * best >= min(steps, tgt) else min(steps, tgt)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


def cpu_model() -> str:
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = sorted(float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 9 for n, v in zip(names, r[5:9]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation on the host cores.  When baseline/_ref holds the copy of the
# pure-Python reference (oracle/install_reference.py, made in the build container and shipped with the snapshot) it is the
# UNMODIFIED reference through its own Model API ("reference"); otherwise the oracle port ("port").
# ------------------------------------------------------------------------------------------------
class CpuReference:
    def __init__(self, head, beam, natural):
        self.head, self.beam = head, beam
        self.cfg = synth.make_config(head)
        self.sd = synth.make_state_dict(self.cfg, seed=1111, suppress_end=not natural)
        self.kind, self.model = "port", None
        if os.path.isdir(os.path.join(REF_DIR, "doc2tex")):
            try:
                import copy
                if REF_DIR not in sys.path:
                    sys.path.insert(0, REF_DIR)
                from doc2tex.modules.build_model import Model as RefModel   # the reference's own class
                m = RefModel(copy.deepcopy(self.cfg)).eval()
                m.load_state_dict(self.sd, strict=True)
                self.model, self.kind = m, "reference"
            except Exception as exc:   # missing dependency on this box: fall back to the port, say so
                print(f"[bench] reference import failed ({exc!r}); using the oracle port", file=sys.stderr)

    def step(self, img, mode):
        with torch.no_grad():
            if self.model is None:
                from oracle import oracle_model as om  # the reference arm is the one place bench.py may run oracle/
                if mode == "greedy":
                    return om.recognize_greedy(self.sd, img, self.head, 150, True)[-1]
                return om.recognize_beam(self.sd, img, self.beam, 150)
            B = img.shape[0]
            if self.head == "TFM":
                if mode == "greedy":   # validation_step's call (engine/inferencing.py:142-153) in eval mode
                    return self.model(img, torch.full((B, 1), 1, dtype=torch.long), is_train=False, is_test=True)[0]
                from doc2tex.tools.beam import Beam
                ctx, _, _ = self.model.forward_encoder(img)
                head, out = self.model.predicter.Prediction, []
                for i in range(B):   # the reference beam is batch-1 only (tfm.py:146-148); fresh Beam per image
                    head.beam = Beam(ignore_w=0, start_w=1, stop_w=2, max_len=150, device="cpu")
                    out.append(head.forward_beam(ctx[i:i + 1], self.beam))
                return out
            if mode == "greedy":
                return self.model(img, torch.zeros(B, 151, dtype=torch.long), is_train=False, is_test=True)[0]
            ctx, _, _ = self.model.forward_encoder(img)
            return [self.model.predicter.Prediction.forward_beam(ctx[i:i + 1], batch_max_length=150, beam_size=self.beam)
                    for i in range(B)]


def cpu_sample(args, head, mode, n_images, H, W, repeats=1):
    """(formulas/s, description) of the CPU reference on a bounded sample, warmed once on one image."""
    torch.set_num_threads(os.cpu_count())
    ref = CpuReference(head, args.beam, args.natural)
    n = n_images if mode == "greedy" else max(1, n_images // 4)
    img = synth.make_images(n, H, W, seed=2024)
    ref.step(img[:1], mode)   # warm-up: thread pool, allocator, oneDNN primitive caches
    t0 = time.perf_counter()
    for _ in range(repeats):
        ref.step(img, mode)
    dt = (time.perf_counter() - t0) / repeats
    what = ("the unmodified reference (baseline/_ref) through its own Model API" if ref.kind == "reference"
            else "oracle port of the reference algorithm")
    return n / dt, dt, n, {"value": n / dt, "unit": "formulas/s", "cores": torch.get_num_threads(), "kind": ref.kind,
                           "cpu_model": cpu_model(),
                           "sample": f"{n} images per pass, {repeats} pass(es) after one warm-up image, {mode} full-length "
                                     f"(151 steps, no KV cache), {H}x{W}; {what}"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    value, dt, n, cb = cpu_sample(args, args.head, args.mode, args.ref_batch, args.height, args.width, repeats=max(1, args.steps))
    line = {
        "impl": "reference", "metric": "formulas/sec", "value": value, "unit": "formulas/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        # the SAME config object as the engine arm's line; how each arm walks the workload is the top-level "schedule"
        "config": workload_config(args.head, args.mode, args.beam, args.precision, args.batch, args.height, args.width, args.natural),
        "schedule": "reference CPU implementation on the host cores (bounded sample of the workload)",
        "cpu_baseline": cb,
        "e2e": {"value": value, "unit": "formulas/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(head, mode, beam, precision, batch, H, W, natural):
    dec = "greedy" if mode == "greedy" else f"beam-{beam}"
    return {
        "workload": f"HybridViT (ResNet stem + 6-block ViT + "
                    f"{'4-layer TFM decoder' if head == 'TFM' else 'Attnv2 LSTM coverage-attention decoder'}) {dec} decode, batch {batch} per GPU, "
                    f"{H}x{W} grayscale, max_len 150 "
                    f"({'natural regime: early exit when every row has emitted END' if natural else '151 full-length steps, END suppressed'}), "
                    f"{precision} mode",
        "batch_per_gpu": batch, "image": [H, W], "decode": dec, "decode_steps": T_STEPS,
        "precision": precision, "vocab": 504 if head == "TFM" else 503, "head": head,
        "return_logits": False,
        "api": "PipelinedRecognizer.run (doc2tex_b200/pipeline.py) over Engine.encode / decode_*: token ids, lengths and scores "
               "are returned; the per-step logits tensor (B, l, V) that Model.forward also returns (78 MB per batch) is not "
               "materialised (return_logits=False)",
        "l2": "inputs larger than L2: activations + KV cache per step (>1 GB) exceed the 126 MB L2",
    }


class Bench:
    """One process per GPU; measures one configuration at a time (measure())."""

    def __init__(self, args):
        self.args = args
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)
        self.pk, self.pk_kind = peaks()
        self.sms = torch.cuda.get_device_properties(self.dev).multi_processor_count

    def barrier(self):
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def measure(self, head, mode, precision, batch, steps, warmup, main=False, encode_merge=1):
        """batch = images per rank.  Returns the record dict (same keys on every rank; timings are max over ranks)."""
        from doc2tex_b200 import dist as d2dist
        from doc2tex_b200.engine import Engine
        from doc2tex_b200.modules.build_model import Model
        from doc2tex_b200.pipeline import PipelinedRecognizer
        a, dev, world, rank = self.args, self.dev, self.world, self.rank
        H, W, T = a.height, a.width, T_STEPS
        cfg = synth.make_config(head, beam_size=(a.beam if mode == "beam" else 1))
        cfg["engine"] = {"precision": precision, "use_graphs": not a.no_graphs}
        sd = synth.make_state_dict(cfg, seed=1111, suppress_end=not a.natural)
        model = Model(cfg)
        model.load_state_dict(sd, strict=True)
        model = model.to(dev)
        eng: Engine = model.engine
        if a.decode_groups > 0:
            eng.set_option("decode_groups", a.decode_groups)
        for kv in a.opt:
            k, v = kv.split("=")
            eng.set_option(k, int(v))
        B = batch
        # rank r holds images [r*B, (r+1)*B) of the global batch (seeded per image)
        img_host = synth.make_images(B, H, W, seed=2024 + rank * B).pin_memory()
        img_dev = img_host.to(dev)
        merge = a.decode_merge if a.decode_merge > 0 else auto_merge(mode, steps, batch)
        pipe = PipelinedRecognizer(eng, mode, a.beam, T, encoder_sms=None if a.sequential else a.encoder_sms,
                                   decode_merge=1 if a.sequential else merge, overlap=not a.no_overlap,
                                   encode_merge=1 if a.sequential else encode_merge)
        if a.no_overlap and not a.sequential:
            eng.set_option("encoder_sms", self.sms)

        def gather(res):
            return d2dist.gather_results(res["ids"], res.get("lens"), res.get("scores"), n_total=B * world)

        def step_device():     # one batch, strictly sequential (latency view)
            ctx, _, _ = eng.encode(img_dev)
            return gather(pipe._decode(ctx))

        def run_steps(k, host):
            """k batches through the public pipelined API; host=True adds the H2D copy of every batch (pinned memory)
            and a D2H read of every result to the timed region."""
            src = img_host if host else img_dev
            if a.sequential:
                for _ in range(k):
                    ctx, _, _ = eng.encode(src.to(dev, non_blocking=True))
                    out = gather(pipe._decode(ctx))[0]
                    if host:
                        out.cpu()
                return
            for res in pipe.run([src] * k):
                out = gather(res)[0]
                if host:
                    out.cpu()

        for _ in range(max(warmup, 3)):
            step_device()
        # the pipelined schedule decodes `decode_merge` batches per call: warm THAT shape too (KV-cache allocation and the
        # step graph of the merged row count must not fall into the timed region)
        run_steps(2 * pipe.decode_merge, False)
        self.barrier()

        sampler = ClockSampler(self.local)
        if rank == 0 and main:
            sampler.start()
        # ---- timed region: exactly K steps (batches), one bracket; working set per step (activations + KV cache,
        # > 1 GB) exceeds the 126 MB L2, the flush buffer is written once before the bracket ----
        self.flush.fill_(1)
        self.barrier()
        eng.set_option("time_conv", 1)   # CUDA events around the dominant kernel's launches, on the launching stream
        eng.conv_time()
        l0 = eng.launch_count()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        run_steps(steps, False)
        ev1.record()
        self.barrier()
        launches = eng.launch_count() - l0
        t_ms = ev0.elapsed_time(ev1)
        conv_ms, conv_n, conv_flops = eng.conv_time()
        eng.set_option("time_conv", 0)
        # stage split (separate sequential pass, same workload): encoder vs decode, CUDA events on the launching stream
        enc_ms = dec_ms = 0.0
        eng.set_option("encoder_sms", self.sms)
        n_split = min(steps, 3)
        for _ in range(n_split):
            self.flush.fill_(1)
            torch.cuda.synchronize()
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            e0.record()
            ctx, _, _ = eng.encode(img_dev)
            e1.record()
            pipe._decode(ctx)
            e2.record()
            torch.cuda.synchronize()
            enc_ms += e0.elapsed_time(e1)
            dec_ms += e1.elapsed_time(e2)
        enc_ms /= n_split
        dec_ms /= n_split
        seq_ms = enc_ms + dec_ms
        # decode roofline: one eager decode call of the row count the schedule really runs (decode_merge batches), the
        # memory-bound launches of decoder layer 1 / the pick kernels bracketed by CUDA events on the launching stream
        roof_dec = None
        if head == "TFM":
            merged = ctx if (a.sequential or pipe.decode_merge == 1) else ctx.repeat(pipe.decode_merge, 1, 1)
            eng.set_option("time_decode", 1)
            for k in range(4):
                eng.decode_time(k)
            pipe._decode(merged)
            torch.cuda.synchronize()
            hbm = self.pk["hbm_gbs"]
            roof_dec = {"rows": int(merged.shape[0] * (a.beam if mode == "beam" else 1)), "peak": hbm, "unit": "GB/s",
                        "peak_kind": f"{self.pk_kind} HBM copy bandwidth", "bound": "hbm",
                        "how": "algorithmic bytes (SURVEY 8d) / CUDA-event time of every launch of decoder layer 1 over the 151 steps "
                               "of one eager decode call (events cannot be timed inside the step graph); self-attention counts every "
                               "row's K/V prefix, so under beam search (hypotheses share prefixes and the encoder memory through "
                               "L1 / L2) it may exceed the DRAM peak"}
            names = {0: "self_attention", 1: "cross_attention", 2: "beam_step", 3: "greedy_pick"}
            kern = {0: "decode_attention_image_kernel" if mode == "beam" else "decode_attention_kernel",
                    1: "decode_attention_image_kernel" if mode == "beam" else "decode_attention_kernel",
                    2: "beam_step_kernel", 3: "greedy_pick_kernel"}
            for k in range(4):
                ms, n, by = eng.decode_time(k)
                if n > 0 and ms > 0:
                    ach = by / (ms * 1e-3) / 1e9
                    roof_dec[names[k]] = {"kernel": kern[k], "achieved": ach, "frac": ach / hbm, "launches_timed": int(n),
                                          "avg_launch_us": 1e3 * ms / n, "algorithmic_mb_per_launch": by / n / 1e6}
            eng.set_option("time_decode", 0)
        if not a.sequential and not a.no_overlap:
            eng.set_option("encoder_sms", a.encoder_sms)
        # end to end through the public API with host buffers
        run_steps(pipe.decode_merge, True)
        self.barrier()
        t0 = time.perf_counter()
        run_steps(steps, True)
        self.barrier()
        e2e_s = time.perf_counter() - t0
        clocks = sampler.stop() if (rank == 0 and main) else None

        tt = torch.tensor([t_ms, e2e_s * 1e3], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_ms, e2e_ms = tt.tolist()
        value = B * world * steps / (t_ms / 1e3)
        e2e_value = B * world * steps / (e2e_ms / 1e3)
        # roofline of the dominant kernel: the tcgen05 implicit-GEMM contraction on a layer3/4 3x3 convolution
        # (512 -> 512 channels; 16 such launches are ~70 % of the encoder).  achieved = algorithmic FLOPs of
        # one launch (2*M*N*K) / average launch duration, CUDA events around the launch inside the timed region above.
        tf_peak = self.pk["bf16_tflops_sustained"]
        passes = {"fp32": 0, "tf32x3": 3, "bf16x3": 3, "bf16": 1}[precision]
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            if tj.get("precision") == precision and tj.get("batch") == B and tj.get("image") == [H, W]:
                traffic = tj.get("dram_bytes_per_launch")
        roof = None
        if conv_n > 0 and conv_ms > 0:
            ach = conv_flops / (conv_ms / conv_n * 1e-3) / 1e12
            gflop = ENC_GFLOP.get((H, W))
            roof = {"bound": "tensor", "achieved": ach, "peak": tf_peak, "unit": "TFLOP/s", "frac": ach / tf_peak,
                    "traffic": traffic,
                    "kernel": "conv_gemm_tc5_kernel (tcgen05 cta_group::2 implicit GEMM, M=256 per CTA pair, activations by TMA "
                              "im2col and weights by TMA from bf16 planes) on layer3.1.conv1 "
                              f"(M={int(conv_flops / (2 * 512 * 4608))}, N=512, K=4608)",
                    "flops_per_launch": conv_flops, "launch_ms": conv_ms / conv_n, "launches_timed": int(conv_n),
                    "mma_passes": passes, "executed_tflops": ach * max(passes, 1),
                    "note": "achieved counts ALGORITHMIC FLOPs; in bf16x3 (fp32-parity) mode the tensor pipe executes 3x that",
                    "peak_kind": f"{self.pk_kind} bf16 sustained (kernel timed inside a long step)",
                    "encoder_tflops": (gflop * B / enc_ms) if gflop else None, "encode_ms": enc_ms, "decode_ms": dec_ms}
        rec = {
            "metric": "formulas/sec", "value": value, "unit": "formulas/s", "n_gpus": world, "steps": steps,
            "warmup": max(warmup, 3), "ms_per_step": t_ms / steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": {"fp32": "f32", "tf32x3": "tf32x3", "bf16x3": "bf16x3", "bf16": "bf16"}[precision],
            "data": "synthetic", "config": workload_config(head, mode, a.beam, precision, B, H, W, a.natural),
            "schedule": (
                "sequential" if a.sequential else
                (f"grouped: {merge} batches encoded back to back on all SMs, then decoded in one call; one batch "
                 f"alone takes {seq_ms:.1f} ms") if a.no_overlap else
                f"pipelined: encode on {a.encoder_sms} SMs overlaps the decode of the previous batches, {merge} encoded "
                f"batch(es) per decode call" + (f", {encode_merge} per encode call" if encode_merge > 1 else "") +
                f"; one batch alone takes {seq_ms:.1f} ms (encode {enc_ms:.1f} + decode {dec_ms:.1f})"),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "formulas/s", "h2d_bytes_per_step": img_host.numel() * 4,
                    "d2h_bytes_per_step": B * world * (T + 2) * 8 if world > 1 else B * T * 8},
            "gpu_launches": int(launches), "roofline": roof, "roofline_decode": roof_dec,
        }
        if not main:
            del rec["clocks"]
        eng.close()
        model._engine = None
        del model, eng, pipe
        torch.cuda.empty_cache()
        return rec


def run_engine(args):
    b = Bench(args)
    world, rank = b.world, b.rank
    line = b.measure(args.head, args.mode, args.precision, args.batch, args.steps, args.warmup, main=True)
    want = args.records
    names = [] if want == "none" else (
        ["beam5", "bf16_greedy", "bf16_beam5", "strong_greedy", "strong_beam5", "attnv2_b512", "natural_greedy", "natural_beam5",
         "sweep_64x256_b32", "sweep_64x256_b1024", "sweep_96x384_b256", "sweep_128x512_b128", "sweep_160x704_b64",
         "sweep_192x896_b64"] if want == "all" else want.split(","))
    # sub-records: two full merged decode groups when the main record has that many steps (the second group's encodes overlap
    # the first group's decode, as in the main record), else what the main record runs
    rs = args.record_steps if args.record_steps > 0 else min(args.steps, 10)
    rs_g = args.record_steps if args.record_steps > 0 else min(args.steps, 20)
    records = {}

    def one(name):
        if name == "beam5":
            return b.measure("TFM", "beam", "bf16x3", args.batch, rs, 3)
        if name == "bf16_greedy":
            return b.measure("TFM", "greedy", "bf16", args.batch, rs_g, 3)
        if name == "bf16_beam5":
            return b.measure("TFM", "beam", "bf16", args.batch, rs, 3)
        if name in ("strong_greedy", "strong_beam5"):
            # BASELINE configs[2]: ONE batch of 256 images, batch-sharded 256 / N per rank -> value = 256 * steps / time
            # a rank's shard is small, so consecutive global batches are merged per decode call like the main record's are:
            # the timed steps are one full merge group (at least the usual record length)
            mode_s = "greedy" if name == "strong_greedy" else "beam"
            shard = max(1, 256 // world)
            # ... and consecutive shards are ENCODED together up to 256 images (the stem's tiles quantise over the SMs at 32)
            r = b.measure("TFM", mode_s, "bf16x3", shard, max(rs_g if mode_s == "greedy" else rs, merge_target(mode_s, shard)), 3,
                          encode_merge=max(1, 256 // shard))
            r["scaling"] = "strong"
            r["config"]["global_batch"] = max(1, 256 // world) * world
            return r
        if name == "attnv2_b512":
            # BASELINE configs[3]: config/train.yaml default stack (Attnv2), greedy, global batch 512 sharded over the ranks
            r = b.measure("Attnv2", "greedy", "bf16x3", max(1, 512 // world), min(rs, 4), 3)
            r["scaling"] = "strong"
            r["config"]["global_batch"] = max(1, 512 // world) * world
            return r
        if name in ("natural_greedy", "natural_beam5"):
            # SURVEY 8d's second decode-length regime: unmodified END logit, early exit as in the reference (tfm.py:138-140, 174)
            keep = args.natural
            args.natural = True
            try:
                return b.measure("TFM", "greedy" if name == "natural_greedy" else "beam", "bf16x3", args.batch,
                                 rs_g if name == "natural_greedy" else rs, 3)
            finally:
                args.natural = keep
        if name.startswith("sweep_"):
            # BASELINE configs[4]: beam-5 over image sizes / batches, e.g. sweep_192x896_b64 (images per GPU; weak scaling)
            hw, bs = name[len("sweep_"):].split("_b")
            h, w = (int(v) for v in hw.split("x"))
            keep = (args.height, args.width)
            args.height, args.width = h, w
            try:
                return b.measure("TFM", "beam", "bf16x3", int(bs), 2, 3)
            finally:
                args.height, args.width = keep
        raise SystemExit(f"unknown record {name!r}")

    for name in names:
        if name.startswith("strong") and world == 1:
            continue   # one rank: the strong-scaling shard IS the main / beam5 record
        try:
            records[name] = one(name)
        except SystemExit:
            raise
        except Exception as ex:   # a failing sub-record must not take the main record's line with it (same on every rank)
            records[name] = {"error": f"{type(ex).__name__}: {ex}"[:300]}
            torch.cuda.empty_cache()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    line["records"] = records
    if world == 1 and args.cpu_sample > 0:
        _, _, _, cb = cpu_sample(args, args.head, args.mode, args.cpu_sample, args.height, args.width)
        line["cpu_baseline"] = cb
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    # stdout carries exactly ONE JSON line: anything libraries print there (e.g. "NCCL version ...") goes to stderr
    _real_stdout = os.dup(1)
    sys.stdout.flush()
    os.dup2(2, 1)
    _print = print

    def print(*args, **kw):  # noqa: A001  (the two JSON prints below)
        if kw.get("file") is None:
            sys.stdout.flush()
            os.write(_real_stdout, (" ".join(str(x) for x in args) + "\n").encode())
        else:
            _print(*args, **kw)

    if a.impl == "reference":
        run_reference(a)
    else:
        run_engine(a)

#!/usr/bin/env python
"""bench.py — formulas/sec of the recognizer hot path (encode + autoregressive decode).

    python bench.py --gpus N --steps K --warmup W              # the B200 engine
    python bench.py --impl reference --gpus N --steps K ...    # the reference algorithm on host CPU cores

One "step" = one pass of the hot path over one batch of synthetic images:
ResNet stem -> patch-embed -> ViT encoder -> greedy (or beam-5) decode of exactly 151 steps
(END suppressed = the deterministic "full-length" regime) -> one all-gather of the token ids.
Per-GPU work is fixed (batch 256 per rank, weak scaling); ranks share nothing but the final
all-gather.  Prints ONE JSON line (rank 0).
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
    ap.add_argument("--decode-groups", type=int, default=0,
                    help="concurrent row groups of a decode call (0 = engine default)")
    ap.add_argument("--decode-merge", type=int, default=0,
                    help="encoded batches handed to one decode call by the pipelined schedule (0 = default for the mode)")
    ap.add_argument("--natural", action="store_true",
                    help="natural decode regime: unmodified END logit, early exit as in the reference (tfm.py:138-140) "
                         "instead of the deterministic full-length regime (SURVEY 8d)")
    ap.add_argument("--no-overlap", action="store_true",
                    help="encode decode_merge batches back to back on all SMs, then decode them in one call (no stage overlap)")
    ap.add_argument("--sequential", action="store_true", help="no encode/decode overlap across batches")
    a = ap.parse_args()
    if a.decode_merge <= 0:
        a.decode_merge = 4   # measured: profiles/r01_pipeline_sweep.txt (decode of 4 encoded batches costs ~1.6x one)
    if a.encoder_sms <= 0:
        a.encoder_sms = 132  # measured sweet spot with merged decode (112 / 88 without)
    return a


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


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
# reference arm: the reference's algorithm (oracle port — the reference is pure Python/PyTorch and
# /root/reference does not exist on the GPU box) on the host cores, O(T^2) decode and all.
# ------------------------------------------------------------------------------------------------
def cpu_reference_step(sd, img, mode, beam, head="TFM"):
    from oracle import oracle_model as om  # the reference arm is the one place bench.py may run oracle/
    if mode == "greedy":
        return om.recognize_greedy(sd, img, head, 150, True)[-1]
    return om.recognize_beam(sd, img, beam, 150)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count())
    cfg = synth.make_config(args.head)
    sd = synth.make_state_dict(cfg, seed=1111, suppress_end=not args.natural)
    n = args.ref_batch if args.mode == "greedy" else max(1, args.ref_batch // 4)
    img = synth.make_images(n, args.height, args.width, seed=2024)
    for _ in range(min(args.warmup, 1)):
        cpu_reference_step(sd, img[:1], args.mode, args.beam, args.head)
    times = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        cpu_reference_step(sd, img, args.mode, args.beam, args.head)
        times.append(time.perf_counter() - t0)
    total = sum(times)
    value = n * args.steps / total
    sample = f"{n} images/step x {args.steps} steps, {args.mode} full-length (151 steps), {args.height}x{args.width}"
    line = {
        "impl": "reference", "metric": "formulas/sec", "value": value, "unit": "formulas/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": dict(workload_config(args, args.batch), schedule="reference algorithm on host CPU cores (bounded sample)"),
        "cpu_baseline": {"value": value, "unit": "formulas/s", "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "formulas/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, batch):
    dec = "greedy" if args.mode == "greedy" else f"beam-{args.beam}"
    return {
        "workload": f"HybridViT (ResNet stem + 6-block ViT + "
                    f"{'4-layer TFM decoder' if args.head == 'TFM' else 'Attnv2 LSTM coverage-attention decoder'}) {dec} decode, batch {batch} per GPU, "
                    f"{args.height}x{args.width} grayscale, max_len 150 "
                    f"({'natural regime: early exit when every row has emitted END' if args.natural else '151 full-length steps, END suppressed'}), "
                    f"{args.precision} mode",
        "batch_per_gpu": batch, "image": [args.height, args.width], "decode": dec, "decode_steps": 151,
        "precision": args.precision, "vocab": 504 if args.head == "TFM" else 503, "head": args.head,
        "l2": "inputs larger than L2: activations + KV cache per step (>1 GB) exceed the 126 MB L2",
    }


def run_engine(args):
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from doc2tex_b200 import dist as d2dist
    from doc2tex_b200.engine import Engine
    from doc2tex_b200.modules.build_model import Model

    cfg = synth.make_config(args.head, beam_size=(args.beam if args.mode == "beam" else 1))
    cfg["engine"] = {"precision": args.precision, "use_graphs": not args.no_graphs}
    sd = synth.make_state_dict(cfg, seed=1111, suppress_end=not args.natural)
    model = Model(cfg)
    model.load_state_dict(sd, strict=True)
    model = model.to(dev)
    eng: Engine = model.engine
    if args.decode_groups > 0:
        eng.set_option("decode_groups", args.decode_groups)

    B, H, W = args.batch, args.height, args.width
    # rank r holds images [r*B, (r+1)*B) of the global batch (seeded per image)
    img_host = synth.make_images(B, H, W, seed=2024 + rank * B).pin_memory()
    img_dev = img_host.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    T = 151

    from doc2tex_b200.pipeline import PipelinedRecognizer
    pipe = PipelinedRecognizer(eng, args.mode, args.beam, T, encoder_sms=None if args.sequential else args.encoder_sms,
                               decode_merge=1 if args.sequential else args.decode_merge, overlap=not args.no_overlap)
    if args.no_overlap and not args.sequential:
        eng.set_option("encoder_sms", torch.cuda.get_device_properties(dev).multi_processor_count)

    def gather(res):
        return d2dist.gather_results(res["ids"], res.get("lens"), res.get("scores"), n_total=B * world)

    def step_device():     # one batch, strictly sequential (latency view)
        ctx, _, _ = eng.encode(img_dev)
        return gather(pipe._decode(ctx))

    def run_steps(k, host):
        """k batches through the public pipelined API; host=True adds the H2D copy of every batch (pinned memory)
        and a D2H read of every result to the timed region."""
        src = img_host if host else img_dev
        if args.sequential:
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

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step_device()
    # the pipelined schedule decodes `decode_merge` batches per call: warm THAT shape too (KV-cache allocation and the
    # step graph of the merged row count must not fall into the timed region)
    run_steps(2 * pipe.decode_merge, False)
    barrier()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # ---- timed region: exactly K steps (batches), one bracket; working set per step (activations + KV cache,
    # > 1 GB) exceeds the 126 MB L2, the flush buffer is written once before the bracket ----
    flush.fill_(1)
    barrier()
    eng.set_option("time_conv", 1)   # CUDA events around the dominant kernel's launches, on the launching stream
    eng.conv_time()
    l0 = eng.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    run_steps(args.steps, False)
    ev1.record()
    barrier()
    launches = eng.launch_count() - l0
    t_ms = ev0.elapsed_time(ev1)
    conv_ms, conv_n, conv_flops = eng.conv_time()
    eng.set_option("time_conv", 0)
    # stage split (separate sequential pass, same workload): encoder vs decode, CUDA events on the launching stream
    enc_ms = dec_ms = seq_ms = 0.0
    eng.set_option("encoder_sms", torch.cuda.get_device_properties(dev).multi_processor_count)
    for _ in range(args.steps):
        flush.fill_(1)
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
    enc_ms /= args.steps
    dec_ms /= args.steps
    seq_ms = enc_ms + dec_ms
    if not args.sequential and not args.no_overlap:
        eng.set_option("encoder_sms", args.encoder_sms)
    # end to end through the public API with host buffers
    run_steps(pipe.decode_merge, True)
    barrier()
    t0 = time.perf_counter()
    run_steps(args.steps, True)
    barrier()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None

    tt = torch.tensor([t_ms, e2e_s * 1e3], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_ms, e2e_ms = tt.tolist()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pk, pk_kind = peaks()
    value = B * world * args.steps / (t_ms / 1e3)
    e2e_value = B * world * args.steps / (e2e_ms / 1e3)
    # roofline of the dominant kernel: the tcgen05 implicit-GEMM contraction on a layer3/4 3x3 convolution
    # (512 -> 512 channels; 16 such launches are ~70 % of the encoder).  achieved = algorithmic FLOPs of
    # one launch (2*M*N*K) / average launch duration, CUDA events around the launch inside the timed region above.
    tf_peak = pk["bf16_tflops_sustained"]
    passes = {"fp32": 0, "tf32x3": 3, "bf16x3": 3, "bf16": 1}[args.precision]
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        if tj.get("precision") == args.precision and tj.get("batch") == B and tj.get("image") == [H, W]:
            traffic = tj.get("dram_bytes_per_launch")
    roof = None
    if conv_n > 0 and conv_ms > 0:
        ach = conv_flops / (conv_ms / conv_n * 1e-3) / 1e12
        gflop = ENC_GFLOP.get((H, W))
        roof = {"bound": "tensor", "achieved": ach, "peak": tf_peak, "unit": "TFLOP/s", "frac": ach / tf_peak,
                "traffic": traffic,
                "kernel": "conv_gemm_tc3_kernel (tcgen05 implicit GEMM, cp.async-fed bf16 planes) on layer3.1.conv1 "
                          f"(M={int(conv_flops / (2 * 512 * 4608))}, N=512, K=4608)",
                "flops_per_launch": conv_flops, "launch_ms": conv_ms / conv_n, "launches_timed": int(conv_n),
                "mma_passes": passes, "executed_tflops": ach * max(passes, 1),
                "note": "achieved counts ALGORITHMIC FLOPs; in bf16x3 (fp32-parity) mode the tensor pipe executes 3x that",
                "peak_kind": f"{pk_kind} bf16 sustained (kernel timed inside a long step)",
                "encoder_tflops": (gflop * B / enc_ms) if gflop else None, "encode_ms": enc_ms, "decode_ms": dec_ms}
    line = {
        "metric": "formulas/sec", "value": value, "unit": "formulas/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": t_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": {"fp32": "f32", "tf32x3": "tf32x3", "bf16x3": "bf16x3", "bf16": "bf16"}[args.precision],
        "data": "synthetic", "config": dict(workload_config(args, B), schedule=(
            "sequential" if args.sequential else
            (f"grouped: {args.decode_merge} batches encoded back to back on all SMs, then decoded in one call; one batch "
             f"alone takes {seq_ms:.1f} ms") if args.no_overlap else
            f"pipelined: encode on {args.encoder_sms} SMs overlaps the decode of the previous batches, {args.decode_merge} encoded "
            f"batch(es) per decode call; one batch alone takes {seq_ms:.1f} ms")),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "formulas/s", "h2d_bytes_per_step": img_host.numel() * 4,
                "d2h_bytes_per_step": B * world * (T + 2) * 8 if world > 1 else B * T * 8},
        "gpu_launches": int(launches), "roofline": roof,
    }
    if world == 1 and args.cpu_sample > 0:
        torch.set_num_threads(os.cpu_count())
        n = args.cpu_sample if args.mode == "greedy" else max(1, args.cpu_sample // 4)
        sub = synth.make_images(n, H, W, seed=2024)
        t0 = time.perf_counter()
        cpu_reference_step(sd, sub, args.mode, args.beam, args.head)
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": n / dt, "unit": "formulas/s", "cores": torch.get_num_threads(), "kind": "port",
                                "sample": f"{n} images, one pass, {args.mode} full-length (151 steps), {H}x{W}; "
                                          f"oracle port of the reference algorithm (no KV cache)"}
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

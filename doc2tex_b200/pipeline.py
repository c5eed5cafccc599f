"""Two-stage software pipeline over batches: encode(i+1) overlaps decode(i).

The two halves of the path stress different resources: the encoder is tensor-pipe bound and fills every SM
it is given, the autoregressive decode is a chain of ~45 small latency-bound kernels per step that occupy a few
dozen SMs.  Giving the encoder's persistent kernels ``encoder_sms`` SMs (d2t_set_option) and running it on a
side stream lets the decode of the previous batch proceed on the remaining SMs, so the steady-state cost per
batch is max(encode, decode) instead of their sum.  Results per batch are identical to the sequential calls.
"""
from __future__ import annotations

from typing import Iterable, Iterator, Optional

import torch

from .engine import Engine


class PipelinedRecognizer:
    def __init__(self, engine: Engine, mode: str = "greedy", beam: int = 5, max_steps: Optional[int] = None,
                 encoder_sms: Optional[int] = None, is_test: bool = True, return_logits: bool = False):
        self.eng, self.mode, self.beam, self.max_steps = engine, mode, beam, max_steps
        self.is_test, self.return_logits = is_test, return_logits
        self.enc_stream = torch.cuda.Stream(device=engine.device)
        self.timing = None   # set to [] to collect (encode_ms, decode_ms) per batch (CUDA events; adds two syncs per batch)
        if encoder_sms is not None:
            engine.set_option("encoder_sms", encoder_sms)

    def _decode(self, ctx):
        if self.mode == "greedy":
            ids, logits, steps = self.eng.decode_greedy(ctx, self.max_steps, is_test=self.is_test,
                                                        return_logits=self.return_logits)
            return {"ids": ids[:, :steps], "logits": None if logits is None else logits[:, :steps], "steps": steps}
        ids, lens, scores, steps, _, _ = self.eng.decode_beam(ctx, self.beam, self.max_steps)
        return {"ids": ids, "lens": lens, "scores": scores, "steps": steps}

    def run(self, batches: Iterable[torch.Tensor]) -> Iterator[dict]:
        """batches: (B,1,H,W) fp32 tensors, on the device or in (pinned) host memory.  Yields one result dict per
        batch, in order."""
        main = torch.cuda.current_stream(self.eng.device)
        pending = None
        for img in batches:
            self.enc_stream.wait_stream(main)
            with torch.cuda.stream(self.enc_stream):
                x = img.to(self.eng.device, non_blocking=True)
                t0 = torch.cuda.Event(enable_timing=True) if self.timing is not None else None
                if t0 is not None:
                    t0.record(self.enc_stream)
                ctx, _, _ = self.eng.encode(x)
                done = torch.cuda.Event(enable_timing=self.timing is not None)
                done.record(self.enc_stream)
            ctx.record_stream(main)
            x.record_stream(self.enc_stream)
            if pending is not None:
                yield self._finish(pending, main)
            pending = (ctx, done, t0)
        if pending is not None:
            yield self._finish(pending, main)

    def _finish(self, pending, main):
        ctx, done, t0 = pending
        main.wait_event(done)          # decode(i) starts when encode(i) is done; encode(i+1) is already enqueued
        if self.timing is None:
            return self._decode(ctx)
        d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        d0.record(main)
        out = self._decode(ctx)
        d1.record(main)
        d1.synchronize()
        self.timing.append((t0.elapsed_time(done), d0.elapsed_time(d1)))
        return out

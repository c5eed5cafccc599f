"""Two-stage software pipeline over batches: encode(i+1) overlaps decode(i).

The two halves of the path stress different resources: the encoder is tensor-pipe bound and fills every SM
it is given, the autoregressive decode is a chain of ~60 small latency-bound kernels per step that occupy a few
dozen SMs at a few hundred rows.  Giving the encoder's persistent kernels ``encoder_sms`` SMs (d2t_set_option) and running it on a
side stream lets the decode of the previous batch proceed on the remaining SMs, so the steady-state cost per
batch is max(encode, decode) instead of their sum.  Results per batch are identical to the sequential calls.

``decode_merge = M`` additionally hands the decode stage M encoded batches per call: the decode step is a chain of
dependent launches whose duration grows far slower than its rows until the attention walks are HBM-bound (151 steps:
256 rows 34 ms, 1 024 rows 66 ms, 2 560 rows 119 ms), so decoding M batches as one call costs far less than M calls and the
decode stage stops being the bottleneck.  Rows never interact across images, so the per-batch results (tokens, lengths, scores, and the
reference's early-exit step count, recomputed per batch from its own rows) are unchanged.
"""
from __future__ import annotations

from typing import Iterable, Iterator, Optional

import torch

from .engine import Engine


class PipelinedRecognizer:
    def __init__(self, engine: Engine, mode: str = "greedy", beam: int = 5, max_steps: Optional[int] = None,
                 encoder_sms: Optional[int] = None, is_test: bool = True, return_logits: bool = False,
                 decode_merge: int = 1, overlap: bool = True, encode_merge: int = 1):
        self.eng, self.mode, self.beam, self.max_steps = engine, mode, beam, max_steps
        self.decode_merge = max(1, int(decode_merge))
        # consecutive SMALL batches of one image size are also encoded as one call: the stem convolutions' tiles quantise
        # over the SMs (a 32-image batch runs at 131 us per image against 106 at 256: two waves of 130 CTA-pair tiles on 74
        # pairs); results still come back per input batch
        self.encode_merge = max(1, int(encode_merge))
        # overlap=False: encode the batches of a group back to back on all SMs, then decode them in one call — no
        # concurrency between the stages (they slow each other down by about what the overlap saves), only the
        # amortisation of the decode chain over more rows
        self.overlap = overlap
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
        M = self.decode_merge
        queue = []   # encoded (or being encoded) batches not yet decoded
        if not self.overlap:
            for img in batches:
                x = img.to(self.eng.device, non_blocking=True)
                ctx, _, _ = self.eng.encode(x)
                done = torch.cuda.Event()
                done.record(main)
                queue.append((ctx, done, None))
                if len(queue) >= M:
                    yield from self._finish(queue[:M], main)
                    del queue[:M]
            if queue:
                yield from self._finish(queue, main)
            return
        group = []   # input batches waiting to be encoded together (encode_merge)

        def encode_group():
            self.enc_stream.wait_stream(main)
            with torch.cuda.stream(self.enc_stream):
                xs = [g.to(self.eng.device, non_blocking=True) for g in group]
                x = xs[0] if len(xs) == 1 else torch.cat(xs, 0)
                t0 = torch.cuda.Event(enable_timing=True) if self.timing is not None else None
                if t0 is not None:
                    t0.record(self.enc_stream)
                ctx, _, _ = self.eng.encode(x)
                done = torch.cuda.Event(enable_timing=self.timing is not None)
                done.record(self.enc_stream)
            ctx.record_stream(main)
            for t in xs + [x]:
                t.record_stream(self.enc_stream)
            r0 = 0
            for g in group:
                queue.append((ctx[r0: r0 + g.shape[0]], done, t0))
                r0 += g.shape[0]
            group.clear()

        for img in batches:
            if group and img.shape[1:] != group[0].shape[1:]:
                encode_group()
            group.append(img)
            if len(group) >= self.encode_merge:
                encode_group()
            # decode blocks the host, so the encodes that should overlap it must be enqueued first:
            # M batches are decoded while the next M are already on the encoder stream
            while len(queue) >= 2 * M:
                yield from self._finish(queue[:M], main)
                del queue[:M]
        if group:
            encode_group()
        while queue:
            yield from self._finish(queue[:M], main)
            del queue[:M]

    def _finish(self, pending, main):
        # only batches of one token geometry can share a decode call
        if any(p[0].shape[1:] != pending[0][0].shape[1:] for p in pending):
            out = []
            for p in pending:
                out += self._finish([p], main)
            return out
        for _, done, _ in pending:
            main.wait_event(done)      # decode starts when its encodes are done; the next encodes are already enqueued
        ctxs = [p[0] for p in pending]
        ctx = ctxs[0] if len(ctxs) == 1 else torch.cat(ctxs, 0)
        d0 = d1 = None
        if self.timing is not None:
            d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            d0.record(main)
        res = self._decode(ctx)
        if d1 is not None:
            d1.record(main)
            d1.synchronize()
            for _, done, t0 in pending:
                self.timing.append((t0.elapsed_time(done), d0.elapsed_time(d1) / len(pending)))
        return self._split(res, [c.shape[0] for c in ctxs])

    def _split(self, res, sizes):
        """Per-batch views of a merged decode result."""
        if len(sizes) == 1:
            return [res]
        out, r0 = [], 0
        for n in sizes:
            sl = slice(r0, r0 + n)
            r0 += n
            if self.mode != "greedy":
                out.append({"ids": res["ids"][sl], "lens": res["lens"][sl], "scores": res["scores"][sl], "steps": res["steps"]})
                continue
            ids, steps = res["ids"][sl], res["steps"]
            if self.is_test and steps > 0:
                # the reference stops a batch at the first step where every one of ITS rows has emitted END
                # (tfm.py:138-140, seq2seq_v2.py:286-289)
                is_end = ids == self.eng.end_id
                if bool(is_end.any(1).all()):
                    steps = int((is_end.int().argmax(1) + 1).max())
            out.append({"ids": ids[:, :steps], "steps": steps,
                        "logits": None if res["logits"] is None else res["logits"][sl][:, :steps]})
        return out

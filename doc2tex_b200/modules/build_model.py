"""Drop-in ``Model`` for the recognizer hot path, backed by the B200 engine.

Mirrors ``doc2tex/modules/build_model.py::Model`` of the reference (lines 7-79): the same
constructor argument (the YAML/config dict), the same ``state_dict`` key schema (SURVEY.md
Appendix C, so reference checkpoints load with ``strict=True``), and the same call surface::

    Model(opt).forward(input, text, is_train=True, is_test=False, rtl_text=None)
        -> (prediction, logits, addition_outputs)                      # build_model.py:55-79
    Model.forward_encoder(input) -> (ctx, output_shape, feat_pad)      # :36-43
    Model.forward_decoder(ctx, text, is_train, is_test, rtl_text)
        -> (prediction, logits, decoder_attn, addition_outputs)        # :45-53

Differences, all on purpose:
  * inference only — ``train(True)`` raises (the reference trains through the same class).  ``is_train`` follows the
    reference: the TFM head never looks at it (PredictBuilder drops it, build_pred.py:46-49; tfm.py:188-195 branches on
    ``self.training``), so ``model(image, text)`` with the default ``is_train=True`` — the call of the batched evaluation
    loop, engine/inferencing.py:151-153 — decodes greedily in eval mode; the LSTM heads use it for teacher forcing only
    (seq2seq_v2.py:276), which is training and raises here;
  * the arithmetic runs in hand-written sm_100a kernels behind the C ABI; there is no PyTorch
    or CPU fallback (a missing library / non-CUDA device raises);
  * beam search accepts B > 1 for both heads (the reference asserts B == 1, tfm.py:146-148 and
    seq2seq_v2.py:18-19) and gives every image a fresh beam (the reference never resets the TFM one —
    SURVEY quirk Q6);
  * optional ``opt["engine"] = {"precision": "fp32"|"tf32x3"|"bf16x3"|"bf16", "use_graphs": bool}``.
"""
from __future__ import annotations

import copy
from typing import Optional

import torch
import torch.nn as nn

from .. import synth
from ..engine import Engine, EngineError

_BUFFER_LEAVES = ("running_mean", "running_var", "num_batches_tracked", "pe")


def _register_tree(root: nn.Module, sd) -> None:
    for key, t in sd.items():
        parts = key.split(".")
        mod = root
        for p in parts[:-1]:
            if not hasattr(mod, p):
                mod.add_module(p, nn.Module())
            mod = getattr(mod, p)
        if parts[-1] in _BUFFER_LEAVES:
            mod.register_buffer(parts[-1], t.clone())
        else:
            mod.register_parameter(parts[-1], nn.Parameter(t.clone(), requires_grad=False))


class Model(nn.Module):
    def __init__(self, opt: dict):
        super().__init__()
        self.opt = opt
        self.stages = {
            "Feat": opt["FeatureExtraction"]["name"],
            "Seq": opt["SequenceModeling"]["name"],
            "Pred": opt["Prediction"]["name"],
        }
        if "Vi" in self.stages["Seq"]:
            assert self.stages["Feat"] == "None"
        # PredictBuilder injects these into the params dict (build_pred.py:16-26); keep the side effect.
        opt["Prediction"].setdefault("params", {})
        opt["Prediction"]["params"]["num_classes"] = opt["num_class"]
        opt["Prediction"]["params"]["device"] = opt.get("device", "cuda")
        # random init with the reference's distributions, in the reference's state_dict schema
        _register_tree(self, synth.make_state_dict(opt, seed=int(opt.get("manualSeed", 1111) or 1111)))
        eng = opt.get("engine", {}) or {}
        self._precision = eng.get("precision", "fp32")
        self._use_graphs = bool(eng.get("use_graphs", True))
        self._engine: Optional[Engine] = None
        self._dirty = True
        self.eval()

    # ---- weight plumbing -------------------------------------------------------------------
    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        out = super().load_state_dict(state_dict, strict=strict, **kw)
        self._dirty = True
        return out

    def _apply(self, fn, *a, **kw):
        out = super()._apply(fn, *a, **kw)
        self._dirty = True
        return out

    def train(self, mode: bool = True):
        if mode:
            raise EngineError("doc2tex_b200.Model is an inference engine; training stays in the reference")
        return super().train(False)

    @property
    def engine(self) -> Engine:
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise EngineError("Model must be moved to a CUDA device (.to('cuda')): doc2tex_b200 has no CPU fallback")
        if self._engine is None or self._engine.device != dev:
            if self._engine is not None:
                self._engine.close()
            opt = copy.deepcopy({k: v for k, v in self.opt.items() if k != "engine"})
            self._engine = Engine(opt, dev, self._precision, self._use_graphs)
            self._dirty = True
        if self._dirty:
            self._engine.load_state_dict(self.state_dict())
            self._dirty = False
        return self._engine

    # ---- reference surface -----------------------------------------------------------------
    def forward_encoder(self, input, *args, **kwargs):
        ctx, grid, pad = self.engine.encode(input)
        return ctx, grid, pad

    def forward_decoder(self, contextual_feature, text, is_train=True, is_test=False, rtl_text=None):
        if is_train and self.stages["Pred"] != "TFM":
            # Attn / Attnv2: is_train=True = teacher forcing on the ground-truth text (seq2seq_v2.py:276) — a training path
            raise EngineError("is_train=True (teacher forcing) of the LSTM heads is training; the inference engine decodes "
                              "with is_train=False as engine/inferencing.py:73-76 does")
        eng = self.engine
        beam_size = self.opt.get("beam_size", 1)
        addition_outputs = {}
        ctx = contextual_feature.contiguous()
        if self.stages["Pred"] == "TFM":
            if beam_size > 1:
                ids, lens, score, steps, _, _ = eng.decode_beam(ctx, beam_size)
                ids, lens, score = ids.cpu(), lens.cpu(), score.cpu()
                if ctx.shape[0] == 1:  # the reference's return: (LongTensor (1, len) on CPU, python float)
                    return ids[:, : int(lens[0])], float(score[0]), None, addition_outputs
                addition_outputs["lengths"] = lens
                return ids[:, : int(lens.max())], score.tolist(), None, addition_outputs
            ids, logits, steps = eng.decode_greedy(ctx, is_test=is_test)
            return ids[:, :steps], logits[:, :steps], None, addition_outputs
        steps_total = int(self.opt["batch_max_length"]) + 1
        if beam_size > 1:
            # AttentionV2.forward_beam (seq2seq_v2.py:152-174): (LongTensor (1, len) on CPU, score tensor, alphas or None);
            # batched here (the reference asserts batch 1): ids padded to the longest result, per-image lengths on the side
            ids, lens, score, steps, _, _ = eng.decode_beam(ctx, beam_size, max_steps=steps_total)
            ids, lens, score = ids.cpu(), lens.cpu(), score.cpu()
            if ctx.shape[0] == 1:
                return ids[:, : int(lens[0])], score[0], None, addition_outputs
            addition_outputs["lengths"] = lens
            return ids[:, : int(lens.max())], score, None, addition_outputs
        ids, logits, steps = eng.decode_greedy(ctx, max_steps=steps_total, is_test=is_test)
        return ids, logits, None, addition_outputs

    def forward(self, input, text, is_train=True, is_test=False, rtl_text=None):
        ctx, output_shape, feat_pad = self.forward_encoder(input)
        prediction, logits, decoder_attn, addition_outputs = self.forward_decoder(
            ctx, text=text, is_train=is_train, is_test=is_test, rtl_text=rtl_text)
        return prediction, logits, addition_outputs

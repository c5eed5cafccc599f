"""ORACLE — test infrastructure, NOT product code.

CPU restatement (torch CPU ops, fp32) of the reference recognizer hot path:
ResNet stem -> HybridEmbed -> ViT encoder -> TFM greedy / TFM beam / Attnv2
greedy, computed straight from a reference-schema ``state_dict``.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module; the engine in
``doc2tex_b200/`` never does.

Where the arithmetic lives: in PyTorch itself (third-party; the reference pins
torch==2.7.0 in envs/requirements.txt, this image has 2.11.0).  The
restatement therefore calls the very same torch operators the reference's
modules dispatch to (``F.conv2d``, ``F.batch_norm``, ``F.layer_norm``,
``F.gelu``, ``nn.TransformerDecoder``, ``nn.LSTMCell`` ...), and keeps the
reference's *algorithm*, including its O(T^2) no-KV-cache decode
(tfm.py:125-136) and its Python-side beam bookkeeping (tools/beam.py:68-105).

Parity pinning: the reference has no tests or golden vectors for this path
(SURVEY.md §4), so the oracle is pinned against the live reference imported in
the build container: ``oracle/make_golden.py`` runs both on the same seeded
weights/images, asserts agreement, and commits small fixtures under
``tests/golden/`` which the CPU test-suite re-checks the oracle against.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

SEQ = "seqmodeler.SequenceModeling."
NET = SEQ + "patch_embed.backbone.ConvNet."
PRED = "predicter.Prediction."

SD = Dict[str, torch.Tensor]


# --------------------------------------------------------------------------------------
# encoder
# --------------------------------------------------------------------------------------
def _conv_bn(sd: SD, x, conv: str, bn: str, stride=1, padding=0, relu=True):
    """Conv2d(bias=False) + eval-mode BatchNorm2d (+ReLU)  (resnet.py:32-40, 206-243)."""
    x = F.conv2d(x, sd[NET + conv + ".weight"], None, stride, padding)
    x = F.batch_norm(x, sd[NET + bn + ".running_mean"], sd[NET + bn + ".running_var"],
                     sd[NET + bn + ".weight"], sd[NET + bn + ".bias"], False, 0.0, 1e-5)
    return F.relu(x) if relu else x


def _basic_block(sd: SD, x, name: str):
    """BasicBlock.forward (resnet.py:32-48)."""
    out = _conv_bn(sd, x, name + ".conv1", name + ".bn1", 1, 1, True)
    out = _conv_bn(sd, out, name + ".conv2", name + ".bn2", 1, 1, False)
    res = x
    if (NET + name + ".downsample.0.weight") in sd:
        res = _conv_bn(sd, x, name + ".downsample.0", name + ".downsample.1", 1, 0, False)
    return F.relu(out + res)


def resnet_stem(sd: SD, x: torch.Tensor, taps: Optional[dict] = None) -> torch.Tensor:
    """ResNet.forward (resnet.py:205-245), BasicBlock counts [1,2,5,3] (:262)."""
    def tap(n, t):
        if taps is not None:
            taps[n] = t
        return t
    x = tap("conv0_1", _conv_bn(sd, x, "conv0_1", "bn0_1", 1, 1))
    x = tap("conv0_2", _conv_bn(sd, x, "conv0_2", "bn0_2", 1, 1))
    x = F.max_pool2d(x, 2, 2, 0)
    x = tap("layer1", _basic_block(sd, x, "layer1.0"))
    x = tap("conv1", _conv_bn(sd, x, "conv1", "bn1", 1, 1))
    x = F.max_pool2d(x, 2, 2, 0)
    for b in range(2):
        x = _basic_block(sd, x, f"layer2.{b}")
    tap("layer2", x)
    x = tap("conv2", _conv_bn(sd, x, "conv2", "bn2", 1, 1))
    x = tap("pool3", F.max_pool2d(x, 2, (2, 1), (0, 1)))
    for b in range(5):
        x = _basic_block(sd, x, f"layer3.{b}")
    tap("layer3", x)
    x = tap("conv3", _conv_bn(sd, x, "conv3", "bn3", 1, 1))
    for b in range(3):
        x = _basic_block(sd, x, f"layer4.{b}")
    tap("layer4", x)
    x = tap("conv4_1", _conv_bn(sd, x, "conv4_1", "bn4_1", (2, 1), (0, 1)))
    x = tap("conv4_2", _conv_bn(sd, x, "conv4_2", "bn4_2", 1, 0))
    return x


def patch_embed(sd: SD, feat: torch.Tensor):
    """HybridEmbed.forward after the backbone (patchembed.py:117-141), patch 2x2."""
    fh, fw = feat.shape[2:]
    pad_h, pad_w = (-fh) % 2, (-fw) % 2
    x = F.pad(feat, (0, pad_w, 0, pad_h))
    tok = F.conv2d(x, sd[SEQ + "patch_embed.proj.weight"], sd[SEQ + "patch_embed.proj.bias"], 2)
    gh, gw = tok.shape[2:]
    return tok.flatten(2).transpose(1, 2), (gh, gw), (pad_w, pad_h)


def vit_block(sd: SD, x: torch.Tensor, i: int, heads: int) -> torch.Tensor:
    """Block.forward / Attention.forward / Mlp.forward (vision_transformer.py:26-32, 61-81, 119-122)."""
    p = f"{SEQ}blocks.{i}."
    B, N, C = x.shape
    h = F.layer_norm(x, (C,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], 1e-6)
    qkv = F.linear(h, sd[p + "attn.qkv.weight"], sd[p + "attn.qkv.bias"])
    qkv = qkv.reshape(B, N, 3, heads, C // heads).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    attn = (q @ k.transpose(-2, -1)) * ((C // heads) ** -0.5)
    attn = attn.softmax(dim=-1)
    h = (attn @ v).transpose(1, 2).reshape(B, N, C)
    x = x + F.linear(h, sd[p + "attn.proj.weight"], sd[p + "attn.proj.bias"])
    h = F.layer_norm(x, (C,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], 1e-6)
    h = F.gelu(F.linear(h, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"]))
    return x + F.linear(h, sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])


def interpolated_pos_embed(pos: torch.Tensor, max_grid: Tuple[int, int], height: int, width: int, patch: int = 2):
    """ViTEncoder.interpolating_pos_embedding (vit_encoder.py:58-95): bicubic resize of the max-grid table to the
    grid of this image; height/width = padded feature map size, scale factors (h0 + 0.1) / emb_height."""
    eh, ew = max_grid
    dim = pos.shape[-1]
    h0, w0 = height // patch + 0.1, width // patch + 0.1
    pp = F.interpolate(pos[:, 1:].reshape(1, eh, ew, dim).permute(0, 3, 1, 2), scale_factor=(h0 / eh, w0 / ew),
                       mode="bicubic", align_corners=False)
    assert int(h0) == pp.shape[-2] and int(w0) == pp.shape[-1]
    return torch.cat((pos[:, 0].unsqueeze(0), pp.permute(0, 2, 3, 1).reshape(1, -1, dim)), dim=1)


def encoder_forward(sd: SD, img: torch.Tensor, heads: int = 8, taps: Optional[dict] = None,
                    pos_mode: str = "prefix", max_grid: Optional[Tuple[int, int]] = None):
    """Model.forward_encoder for Seq='ViT' (build_model.py:36-43, build_seq.py:59-66).  pos_mode 'prefix' =
    ViTEncoderV3 / ViTEncoderV2 (vit_encoder.py:205-268: first N+1 rows of the table); 'interpolate' = ViTEncoder
    (:96-118: bicubic resize of the max-grid table unless the image has the max grid).
    Returns (ctx, (gh,gw), (pad_W,pad_H))."""
    with torch.no_grad():
        feat = resnet_stem(sd, img, taps)
        tok, grid, pad = patch_embed(sd, feat)
        if taps is not None:
            taps["patch_embed"] = tok
        B, N, C = tok.shape
        x = torch.cat((sd[SEQ + "cls_token"].expand(B, -1, -1), tok), dim=1)
        pos = sd[SEQ + "pos_embed"]
        if pos_mode == "interpolate" and tuple(grid) != tuple(max_grid):
            x = x + interpolated_pos_embed(pos, max_grid, 2 * grid[0], 2 * grid[1])
        else:
            x = x + pos[:, : N + 1]                          # PREFIX slice (quirk Q3); the full table at the max grid
        depth = 1 + max(int(k.split(".")[3]) for k in sd if k.startswith(SEQ + "blocks."))
        for i in range(depth):
            x = vit_block(sd, x, i, heads)
            if taps is not None:
                taps[f"block{i}"] = x
        C = x.shape[-1]
        x = F.layer_norm(x, (C,), sd[SEQ + "norm.weight"], sd[SEQ + "norm.bias"], 1e-6)
    return x, grid, pad


def topk_margin(cand: torch.Tensor, k: int) -> Tuple[float, float]:
    """Near-tie audit of one beam decision: the smallest gap between neighbours among the k selected
    candidates AND the best one left out (the top k+1 of ``cand``), and the fp32 spacing (ulp) at their magnitude.
    A decision whose gap is a few ulps can legitimately flip under any other fp32 summation order."""
    import numpy as np
    vals = torch.topk(cand.reshape(-1), min(k + 1, cand.numel())).values
    if vals.numel() < 2:
        return float("inf"), 0.0
    gap = float((vals[:-1] - vals[1:]).min())
    return gap, float(np.spacing(np.float32(float(vals.abs().max()))))


# --------------------------------------------------------------------------------------
# TFM head
# --------------------------------------------------------------------------------------
class TFMHead:
    """TransformerPrediction in eval mode (tfm.py:35-195), rebuilt on torch's own
    nn.TransformerDecoder (post-norm, ReLU, eps 1e-5 — tfm.py:19-26 leaves the defaults)."""

    PAD, GO, END = 0, 1, 2

    def __init__(self, sd: SD, nhead: int = 8, max_seq_len: int = 150):
        w1 = sd[PRED + "model.layers.0.linear1.weight"]
        d, ff = w1.shape[1], w1.shape[0]
        nl = 1 + max(int(k.split(".")[4]) for k in sd if k.startswith(PRED + "model.layers."))
        layer = nn.TransformerDecoderLayer(d_model=d, nhead=nhead, dim_feedforward=ff, dropout=0.1)
        self.model = nn.TransformerDecoder(layer, nl)
        self.model.load_state_dict({k[len(PRED + "model."):]: v for k, v in sd.items()
                                    if k.startswith(PRED + "model.")})
        self.model.eval()
        self.embed = sd[PRED + "word_embed.weight"]
        self.pe = sd[PRED + "pos_enc.pe"]
        self.proj_w, self.proj_b = sd[PRED + "proj.weight"], sd[PRED + "proj.bias"]
        self.d, self.max_seq_len = d, max_seq_len

    @staticmethod
    def causal_mask(l: int) -> torch.Tensor:
        """0 on/below the diagonal, -inf above (tfm.py:74-84)."""
        return torch.full((l, l), float("-inf")).triu(1)

    def _run(self, tokens: torch.Tensor, memory_tbd: torch.Tensor) -> torch.Tensor:
        """tokens (b,l) -> logits (b,l,V): embed*sqrt(d)+pe, full decoder pass, proj (tfm.py:86-94,125-133)."""
        l = tokens.shape[1]
        emb = self.embed[tokens] * math.sqrt(self.d) + self.pe[:l][None]
        out = self.model(tgt=emb.transpose(0, 1), memory=memory_tbd, tgt_mask=self.causal_mask(l))
        return F.linear(out.transpose(0, 1), self.proj_w, self.proj_b)

    def greedy(self, ctx: torch.Tensor, is_test: bool = True, max_steps: Optional[int] = None):
        """forward_greedy eval branch (tfm.py:119-143).  Returns the reference's pair
        (preds_index, logits of the LAST pass) plus the generated tokens (B, steps)."""
        with torch.no_grad():
            mem = ctx.transpose(0, 1)
            B = ctx.shape[0]
            tgt = torch.full((B, 1), self.GO, dtype=torch.long)
            end = torch.zeros(B, dtype=torch.bool)
            steps = self.max_seq_len + 1 if max_steps is None else max_steps
            out = None
            for _ in range(steps):
                out = self._run(tgt, mem)
                nxt = torch.argmax(F.softmax(out, dim=-1)[:, -1:, :], dim=-1)
                tgt = torch.cat([tgt, nxt], dim=-1)
                end = end | (nxt[:, 0] == self.END)
                if bool(end.all()) and is_test:
                    break
            return out.max(dim=2)[1], out, tgt[:, 1:]

    def beam(self, ctx1: torch.Tensor, beam_size: int, trace: Optional[list] = None, margins: Optional[list] = None):
        """forward_beam for ONE image with a FRESH beam (tfm.py:145-186, tools/beam.py:38-140;
        the library never resets its Beam — quirk Q6 — the demo does, which is the semantics kept).

        Returns (seq list[int], score float).  ``trace`` (optional list) receives per step
        ``(parents, words, scores)`` of the top-k candidates in top-k order; ``margins`` (optional list) the per-step
        ``topk_margin`` (gap, ulp) of the decision."""
        assert ctx1.shape[0] == 1
        with torch.no_grad():
            L = self.max_seq_len + 2
            hyps = torch.full((1, L), self.PAD, dtype=torch.long)
            hyps[:, 0] = self.GO
            scores = torch.zeros(1)
            completed: List[Tuple[List[int], float]] = []
            mem1 = ctx1[0]
            for step in range(self.max_seq_len + 1):
                n = hyps.shape[0]
                mem = mem1[:, None, :].expand(-1, n, -1)
                logits = self._run(hyps, mem)
                logp = F.log_softmax(logits[:, step, :], dim=-1)
                V = logp.shape[1]
                k = beam_size - len(completed)
                cand = (scores[:, None] + logp).reshape(-1)
                top_s, top_i = torch.topk(cand, k=k)
                parents, words = top_i // V, top_i % V
                if margins is not None:
                    margins.append(topk_margin(cand, k))
                if trace is not None:
                    trace.append((parents.tolist(), words.tolist(), top_s.tolist()))
                new_h, new_s = [], []
                for p, w, s in zip(parents.tolist(), words.tolist(), top_s.tolist()):
                    hyps[p, step + 1] = w
                    if w == self.END:
                        completed.append((hyps[p, 1:step + 2].tolist(), s))
                    else:
                        new_h.append(hyps[p].clone())
                        new_s.append(s)
                if len(completed) == beam_size:
                    break
                hyps = torch.stack(new_h, 0)
                scores = torch.tensor(new_s, dtype=torch.float)
            if not completed:
                completed.append((hyps[0, 1:].tolist(), float(scores[0])))
            best = max(completed, key=lambda h: h[1] / max(len(h[0]), 1))
            return best[0], best[1]


# --------------------------------------------------------------------------------------
# Attnv2 (LSTM + coverage location-aware attention) head, greedy
# --------------------------------------------------------------------------------------
class AttnV2Head:
    """AttentionV2.forward_greedy in eval mode (seq2seq_v2.py:176-293) with
    LocationAwareAttention / LocationAwareAttentionCell (attention1D.py:121-161, 205-242)."""

    GO, END = 0, 1

    def __init__(self, sd: SD, include_cls: bool = False):
        """include_cls=True restates the base ``Attention`` head (Prediction.name 'Attn', seq2seq.py:225-331 and :82-223):
        identical loops, but the decoder attends over every encoder token including cls (seq2seq.py:236-238)."""
        g = lambda n: sd[PRED + n]
        self.sd = sd
        self.tok0 = 0 if include_cls else 1
        self.embedding = g("embedding.weight")
        hs = g("proj_init_h.weight").shape[0]
        ins = g("attention_cell.rnn.weight_ih").shape[1] - self.embedding.shape[1]
        self.rnn = nn.LSTMCell(ins + self.embedding.shape[1], hs)
        self.rnn.load_state_dict({k.split("rnn.")[1]: v for k, v in sd.items() if ".rnn." in k})
        self.rnn.eval()
        self.V = g("attention_cell.generator.weight").shape[0]

    def greedy(self, ctx: torch.Tensor, batch_max_length: int = 150, is_test: bool = True):
        sd, P = self.sd, PRED
        a = P + "attention_cell.attn."
        with torch.no_grad():
            B = ctx.shape[0]
            steps = batch_max_length + 1
            H = ctx[:, self.tok0:, :]
            init = ctx[:, 0, :]
            h = F.linear(init, sd[P + "proj_init_h.weight"], sd[P + "proj_init_h.bias"])
            c = F.linear(init, sd[P + "proj_init_c.weight"], sd[P + "proj_init_c.bias"])
            targets = torch.zeros(B, dtype=torch.long)
            probs = torch.zeros(B, steps, self.V)
            alpha_cum = torch.zeros(B, H.shape[1], 1)
            mem = None
            pad = sd[a + "loc_conv.weight"].shape[2] // 2
            end = torch.zeros(B, dtype=torch.bool)
            for i in range(steps):
                emb = self.embedding[targets]
                kp = F.linear(H, sd[a + "key_proj.weight"], sd[a + "key_proj.bias"])
                qp = F.linear(h, sd[a + "query_proj.weight"], sd[a + "query_proj.bias"]).unsqueeze(1)
                last = mem if mem is not None else h.new_zeros(B, H.shape[1], 1)
                loc = F.conv1d(last.permute(0, 2, 1), sd[a + "loc_conv.weight"], sd[a + "loc_conv.bias"], padding=pad)
                loc = F.linear(loc.transpose(1, 2), sd[a + "loc_proj.weight"], sd[a + "loc_proj.bias"])
                e = F.linear(torch.tanh(kp + qp + loc), sd[a + "score.weight"], sd[a + "score.bias"])
                alpha = F.softmax(e / 1.0, dim=1)
                context = torch.bmm(alpha.permute(0, 2, 1), H).squeeze(1)
                h, c = self.rnn(torch.cat([context, emb], 1), (h, c))
                out = F.linear(h, sd[P + "attention_cell.generator.weight"], sd[P + "attention_cell.generator.bias"])
                alpha_cum = alpha_cum + alpha
                mem = alpha_cum
                probs[:, i, :] = out
                if i == steps - 1:
                    break
                targets = out.max(1)[1]
                if is_test:
                    end = end | (targets == self.END)
                    if bool(end.all()):
                        break
            return probs.max(2)[1], probs


    def _cell(self, h, c, H, emb, mem):
        """LocationAwareAttentionCell.forward (attention1D.py:223-242) for n rows; mem = coverage memory or None."""
        sd, P = self.sd, PRED
        a = P + "attention_cell.attn."
        n = h.shape[0]
        pad = sd[a + "loc_conv.weight"].shape[2] // 2
        kp = F.linear(H, sd[a + "key_proj.weight"], sd[a + "key_proj.bias"])
        qp = F.linear(h, sd[a + "query_proj.weight"], sd[a + "query_proj.bias"]).unsqueeze(1)
        last = mem if mem is not None else h.new_zeros(n, H.shape[1], 1)
        loc = F.conv1d(last.permute(0, 2, 1), sd[a + "loc_conv.weight"], sd[a + "loc_conv.bias"], padding=pad)
        loc = F.linear(loc.transpose(1, 2), sd[a + "loc_proj.weight"], sd[a + "loc_proj.bias"])
        e = F.linear(torch.tanh(kp + qp + loc), sd[a + "score.weight"], sd[a + "score.bias"])
        alpha = F.softmax(e / 1.0, dim=1)
        context = torch.bmm(alpha.permute(0, 2, 1), H).squeeze(1)
        h, c = self.rnn(torch.cat([context, emb], 1), (h, c))
        out = F.linear(h, sd[P + "attention_cell.generator.weight"], sd[P + "attention_cell.generator.bias"])
        return out, h, c, alpha

    def beam(self, ctx1: torch.Tensor, beam_size: int = 5, batch_max_length: int = 150, trace: Optional[list] = None,
             margins: Optional[list] = None):
        """AttentionV2.forward_beam (seq2seq_v2.py:12-174), batch 1, with its quirks:
        step 0 ranks row 0 only (:98-99); hidden follows the parent but the coverage memory is re-indexed by top-k
        POSITION (:137-147); if the LAST executed step completed nothing, live beam 0 wins over earlier completions
        (:152-160); else first max of fp32 score / len(seq incl. GO and END), returned score = max score (:162-168).
        Returns (token list without GO, score float)."""
        sd, P = self.sd, PRED
        with torch.no_grad():
            assert ctx1.shape[0] == 1
            num_steps = batch_max_length + 1
            bH = ctx1[0][None].expand(beam_size, -1, -1)
            H = bH[:, self.tok0:, :]
            init = bH[:, 0, :]
            h = F.linear(init, sd[P + "proj_init_h.weight"], sd[P + "proj_init_h.bias"])
            c = F.linear(init, sd[P + "proj_init_c.weight"], sd[P + "proj_init_c.bias"])
            alpha_cum = torch.zeros(beam_size, H.shape[1], 1)
            mem = None
            seqs = torch.full((beam_size, 1), self.GO, dtype=torch.long)
            targets = seqs.squeeze(-1)
            top_k_scores = torch.zeros(beam_size, 1)
            complete_seqs, complete_scores = [], []
            complete_inds = []
            for step in range(num_steps):
                emb = self.embedding[targets]
                out, h, c, alpha = self._cell(h, c, H, emb, mem)
                V = out.shape[1]
                scores = top_k_scores.expand_as(out) + F.log_softmax(out, dim=-1)
                if margins is not None:
                    margins.append(topk_margin(scores[0] if step == 0 else scores, beam_size))
                if step == 0:
                    top_k_scores, top_k_words = scores[0].topk(beam_size, 0, True, True)
                else:
                    top_k_scores, top_k_words = scores.view(-1).topk(beam_size, 0, True, True)
                prev = top_k_words // V
                nxt = top_k_words % V
                if trace is not None:
                    trace.append((prev.tolist(), nxt.tolist(), top_k_scores.tolist()))
                seqs = torch.cat([seqs[prev], nxt.unsqueeze(1)], dim=1)
                incomplete_inds = [i for i, w in enumerate(nxt.tolist()) if w != self.END]
                complete_inds = list(set(range(len(nxt))) - set(incomplete_inds))
                if complete_inds:
                    complete_seqs.extend(seqs[complete_inds].tolist())
                    complete_scores.extend(top_k_scores[complete_inds])
                beam_size -= len(complete_inds)
                if beam_size == 0:
                    break
                seqs = seqs[incomplete_inds]
                h, c = h[prev[incomplete_inds]], c[prev[incomplete_inds]]
                H = H[prev[incomplete_inds]]
                top_k_scores = top_k_scores[incomplete_inds].unsqueeze(1)
                targets = nxt[incomplete_inds]
                alpha_cum = (alpha_cum + alpha)[incomplete_inds]   # by position, not by parent (quirk Q10)
                mem = alpha_cum
            if len(complete_inds) == 0:
                return seqs[0][1:].tolist(), float(top_k_scores[0])
            combine = tuple(zip(complete_seqs, complete_scores))
            best = combine.index(max(combine, key=lambda x: x[1] / len(x[0])))
            return complete_seqs[best][1:], float(max(complete_scores))


# --------------------------------------------------------------------------------------
# whole-path helpers (what bench.py's reference arm times)
# --------------------------------------------------------------------------------------
def recognize_greedy(sd: SD, img: torch.Tensor, head: str = "TFM", max_len: int = 150, is_test: bool = True):
    ctx, _, _ = encoder_forward(sd, img)
    if head == "TFM":
        return TFMHead(sd, max_seq_len=max_len).greedy(ctx, is_test)
    return AttnV2Head(sd).greedy(ctx, max_len, is_test)


def recognize_beam(sd: SD, img: torch.Tensor, beam_size: int = 5, max_len: int = 150):
    """Reference beam is batch-1 only (tfm.py:146-148): loop images, fresh beam each."""
    ctx, _, _ = encoder_forward(sd, img)
    head = TFMHead(sd, max_seq_len=max_len)
    return [head.beam(ctx[i:i + 1], beam_size) for i in range(ctx.shape[0])]

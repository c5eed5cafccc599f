"""Batched evaluation loop on the B200 engine — the caller side of the hot path.

Mirrors ``doc2tex/engine/inferencing.py::validation_step`` (lines 12-247): same signature, same 11-tuple
``(all_loss, names, mean_loss, accuracy, bleu, norm_ED, word_ED, preds, labels, infer_time, n)``, same per-sample
string handling (cut at the first "[s]", optional whitespace post-processing, exact-match accuracy, ICDAR-2019
normalised edit distance, word-level NED, corpus BLEU on the token lists).  What changes is HOW the batches run:

* the recognizer forward goes through ``PipelinedRecognizer``: encode(i+1) overlaps decode(i), ``decode_merge`` encoded
  batches share one decode call, all on hand-written sm_100a kernels behind the C ABI;
* the reference's per-sample Python post-processing (``converter.decode`` builds every full-length string, then
  ``str.find("[s]")``; ``converter.detokenize`` walks every id; inferencing.py:110-140, 185-230) is vectorised: the
  first-END cut is ONE tensor op per batch on the device, ids cross to the host once, tokens are gathered with one numpy
  take and every row costs one ``str.join``;
* the confidence product of inferencing.py:98-105 is not materialised: it starts its reduction at ``torch.zeros`` and is
  therefore identically zero, and it is never returned.

The loss is the caller's ``criterion`` applied to the engine's per-step logits exactly as the reference applies it
(``preds.view(-1, V)`` against ``text_for_loss[:, 1:]``), so any criterion the reference accepts works unchanged.
"""
from __future__ import annotations

import re
import time
from typing import Iterable, List, Sequence

import numpy as np
import torch

from .pipeline import PipelinedRecognizer


# ---------------------------------------------------------------------------------------------------------------------
# vectorised converter.decode / detokenize
# ---------------------------------------------------------------------------------------------------------------------
def first_end_cut(ids: torch.Tensor, end_id: int) -> torch.Tensor:
    """Per row: index of the first END token, or the row length when there is none — one tensor op on the ids' device
    (what ``pred[:pred.find("[s]")]`` / the ``detokenize`` break compute per sample on the host, tfm_converter.py:71-82)."""
    is_end = ids == end_id
    T = ids.shape[1]
    pos = torch.where(is_end, torch.arange(T, device=ids.device).expand_as(ids), torch.full_like(ids, T))
    return pos.min(dim=1).values


def decode_cut(converter, ids: torch.Tensor, token_level: str = "word", cut: torch.Tensor | None = None):
    """(strings, token lists) of a batch of id rows, both cut at the first "[s]".

    strings[i] equals the reference's ``converter.decode(ids)[i]`` followed by ``s[:s.find("[s]")]`` (inferencing.py:88-92,
    119-121) INCLUDING its quirks: with word-level joining the cut string keeps the separator before "[s]" (a trailing
    blank), and a row without any "[s]" loses its last character (``find`` returns -1).  tokens[i] equals
    ``converter.detokenize(ids)[i]``.
    """
    end_id = converter.dict["[s]"]
    if cut is None:
        cut = first_end_cut(ids, end_id)
    ids_np = ids.detach().to("cpu").numpy()
    cut_np = cut.detach().to("cpu").numpy()
    table = getattr(converter, "_np_table", None)
    if table is None:
        table = np.asarray(converter.character, dtype=object)
        converter._np_table = table
        # a vocabulary token that contains the text "[s]" would move the reference's str.find(); none does in practice
        converter._plain_cut = not any("[s]" in tok for i, tok in enumerate(converter.character) if i != end_id)
    sep = " " if token_level == "word" else ""
    toks = table[ids_np]                       # one gather for the whole batch
    T = ids_np.shape[1]
    strings, tokens = [], []
    for row, c in zip(toks, cut_np):
        head = row[:c]
        tokens.append(head.tolist())
        if not converter._plain_cut:
            full = sep.join(row.tolist())
            strings.append(full[: full.find("[s]")])
        elif c < T:
            s = sep.join(head.tolist())
            strings.append(s + sep if c > 0 else s)
        else:
            strings.append(sep.join(head.tolist())[:-1])
    return strings, tokens


# ---------------------------------------------------------------------------------------------------------------------
# string scoring (the reference uses python-Levenshtein and a torchtext-style BLEU; restated here, numpy only)
# ---------------------------------------------------------------------------------------------------------------------
def edit_distance(a: Sequence, b: Sequence) -> int:
    """Levenshtein distance between two sequences (strings or token lists): row-wise DP, the insertion recurrence
    resolved with a running minimum so that every row is a handful of numpy ops."""
    if len(a) < len(b):
        a, b = b, a
    if len(b) == 0:
        return len(a)
    if isinstance(a, str) and isinstance(b, str):
        av = np.frombuffer(a.encode("utf-32-le"), dtype=np.uint32)
        bv = np.frombuffer(b.encode("utf-32-le"), dtype=np.uint32)
    else:
        vocab = {}
        av = np.array([vocab.setdefault(x, len(vocab)) for x in a], dtype=np.int64)
        bv = np.array([vocab.setdefault(x, len(vocab)) for x in b], dtype=np.int64)
    n = len(bv)
    idx = np.arange(n + 1)
    prev = idx.copy()
    for i in range(len(av)):
        cur = np.empty(n + 1, dtype=np.int64)
        cur[0] = i + 1
        cur[1:] = np.minimum(prev[1:] + 1, prev[:-1] + (bv != av[i]))
        cur = np.minimum.accumulate(cur - idx) + idx      # cur[j] = min_k<=j (cur[k] + j - k): insertions
        prev = cur
    return int(prev[n])


def single_ed(gt: str, pred: str) -> float:
    """ICDAR-2019 normalised edit distance of one pair (modules/metrics/ed.py:4-12)."""
    if len(gt) == 0 or len(pred) == 0:
        return 0
    return 1 - edit_distance(pred, gt) / max(len(gt), len(pred))


def word_ned(pred: str, gt: str) -> float:
    """Word-level normalised edit distance of one pair (modules/metrics/ed.py:15-40 with single strings)."""
    wg, wp = gt.split(), pred.split()
    if len(gt) == 0 or len(pred) == 0:
        return 0.0
    return 1 - edit_distance(wg, wp) / max(len(wg), len(wp))


def corpus_bleu(candidates: List[List[str]], references: List[List[List[str]]], max_n: int = 4) -> float:
    """Corpus BLEU with uniform weights and brevity penalty (modules/metrics/bleu.py:56-120: clipped n-gram counts summed over
    the corpus, closest reference length, 0 when any order has no match)."""
    import collections
    import math
    clipped = [0.0] * max_n
    total = [0.0] * max_n
    cand_len = ref_len = 0.0

    def grams(tokens):
        cnt = collections.Counter()
        for n in range(1, max_n + 1):
            for i in range(len(tokens) - n + 1):
                cnt[tuple(tokens[i:i + n])] += 1
        return cnt

    for cand, refs in zip(candidates, references):
        cand_len += len(cand)
        ref_len += min((float(len(r)) for r in refs), key=lambda x: abs(len(cand) - x))
        ref_cnt = collections.Counter()
        for r in refs:
            ref_cnt |= grams(r)
        for g, c in (grams(cand) & ref_cnt).items():
            clipped[len(g) - 1] += c
        for n in range(max_n):
            total[n] += max(len(cand) - n, 0)
    if min(clipped) == 0:
        return 0.0
    log_p = sum(math.log(c / t) for c, t in zip(clipped, total)) / max_n
    bp = math.exp(min(1 - ref_len / cand_len, 0))
    return bp * math.exp(log_p)


_TEXT_CMD = re.compile(r"(\\(operatorname|mathrm|mathbf|mathsf|mathit|mathfrak|mathnormal)\s?\*? {.*?})")
_LETTER, _NOLETTER = "[a-zA-Z]", r"[\W_^\d]"
_SQUEEZE = [re.compile(r"(?!\\ )(%s)\s+?(%s)" % (_NOLETTER, _NOLETTER)), re.compile(r"(?!\\ )(%s)\s+?(%s)" % (_NOLETTER, _LETTER)),
            re.compile(r"(%s)\s+?(%s)" % (_LETTER, _NOLETTER))]


def squeeze_latex_whitespace(s: str) -> str:
    """Blanks that do not separate two letters are dropped, text-mode commands are packed first, repeated to a fixed point
    (what ``Postprocessing.remove_unused_whitespace`` does, utils/data_utils.py:433-455 — like the reference, the value
    returned is the string of the last pass that still changed something's input)."""
    packed = [m[0].replace(" ", "") for m in _TEXT_CMD.findall(s)]
    s = _TEXT_CMD.sub(lambda _m: str(packed.pop(0)), s)
    nxt = s
    while True:
        s = nxt
        for rx in _SQUEEZE:
            nxt = rx.sub(r"\1\2", nxt)
        if nxt == s:
            return s


class _Mean:
    def __init__(self):
        self.total, self.count = 0.0, 0

    def add(self, v: torch.Tensor):
        self.total = self.total + v.detach().sum()
        self.count += v.numel()

    def val(self):
        return self.total / float(self.count) if self.count else 0


def validation_step(model, augment, criterion, evaluation_loader: Iterable, converter, config: dict, args, device,
                    decode_merge: int | None = None, encoder_sms: int | None = 132):
    """validation or evaluation (doc2tex/engine/inferencing.py:12-247) on the engine.

    ``model`` is ``doc2tex_b200.modules.build_model.Model``; ``evaluation_loader`` yields ``(image_tensors, labels,
    img_names)`` like the reference's loader (images already on ``device``, one (H, W) per batch).  Batches are pipelined:
    results come back in order, ``decode_merge`` batches per decode call (default: as many batches of the loader's batch size
    as make about 2 560 rows, at most 16 — the row count at which the decode's attention walks are HBM-bound).
    """
    eng = model.engine
    is_attn = "Attn" in config["Prediction"]["name"]
    T = int(config["batch_max_length"]) + 1
    level = config.get("token_level", "word")
    post = config.get("postprocess", True)
    n_correct, norm_ED, word_ED, length_of_data = 0, 0.0, 0.0, 0
    loss_avg = _Mean()
    all_loss: List[float] = []
    total_pred_tokens, total_truth_tokens = [], []
    total_names, total_labels, total_preds = [], [], []
    writer = fo = None
    if config.get("export_csv"):
        import csv
        import os
        eval_data = str(config.get("eval_data", "eval")).split("/")[-1]
        save_path = f"./result/{config['exp_name']}/{args.log_path[:-4]}_{eval_data}.csv"
        os.makedirs(os.path.dirname(save_path), exist_ok=True)
        fo = open(save_path, "wt")
        writer = csv.writer(fo)

    meta = []     # (labels, img_names) of the batches handed to the pipeline, in order

    def batches():
        for image_tensors, labels, img_names in evaluation_loader:
            if image_tensors is None and labels is None and img_names is None:
                break
            assert image_tensors.device.type == device
            if augment:
                image_tensors = torch.clamp(image_tensors, min=0.0, max=255.0).div(255.0)
                image_tensors = getattr(augment, "normalize")(image_tensors)
            meta.append((labels, img_names))
            yield image_tensors
            if config.get("sanity_check"):
                break

    # greedy, every one of the T steps (the reference calls model(image, text[, is_train=False]) with is_test=False:
    # no early exit, inferencing.py:73-76 / 151-153), logits kept for the caller's criterion
    stream = batches()
    first = next(stream, None)
    if decode_merge is None:
        decode_merge = 1 if first is None else max(1, min(16, 2560 // max(1, int(first.shape[0]))))
    pipe = PipelinedRecognizer(eng, "greedy", 1, T, encoder_sms=encoder_sms, is_test=False, return_logits=True,
                               decode_merge=decode_merge)
    start_time = time.time()
    import itertools
    for k, res in enumerate(pipe.run(itertools.chain([] if first is None else [first], stream))):
        labels, img_names = meta[k]
        preds_index, preds = res["ids"], res["logits"]
        batch_size = preds_index.shape[0]
        length_of_data += batch_size
        text_for_loss, _ = converter.encode(labels, batch_max_length=config["batch_max_length"])
        target = text_for_loss[:, 1:].to(preds.device)   # without [GO] Symbol
        costs = criterion(preds.contiguous().view(-1, preds.shape[-1]), target.contiguous().view(-1))
        costs = costs.view(batch_size, -1).mean(dim=1)
        loss_avg.add(costs)
        np_costs = costs.detach().cpu().numpy().tolist()
        all_loss += np_costs
        gts, truth_tokens = decode_cut(converter, target, level)
        preds_str, pred_tokens = decode_cut(converter, preds_index, level)
        for cost, img_name, gt, pred, pred_token, gt_token in zip(np_costs, img_names, gts, preds_str, pred_tokens, truth_tokens):
            if post:
                pred = squeeze_latex_whitespace(pred)
                gt = squeeze_latex_whitespace(gt)
            if pred == gt:
                n_correct += 1
            if writer is not None:
                writer.writerow((cost, img_name, pred, gt, 1 if pred == gt else 0))
            norm_ED += single_ed(gt, pred)
            word_ED += word_ned(pred, gt)
            total_names.append(img_name)
            total_labels.append(gt)
            total_preds.append(pred)
            total_pred_tokens.append(pred_token)
            total_truth_tokens.append(gt_token)
    torch.cuda.synchronize()
    infer_time = time.time() - start_time   # whole pipelined loop (the reference sums unsynchronised per-batch forward times)
    if fo is not None:
        fo.close()
    accuracy = n_correct / float(length_of_data)
    norm_ED = norm_ED / float(length_of_data)
    word_ED = word_ED / float(length_of_data)
    bleu_score = corpus_bleu(total_pred_tokens, [[s] for s in total_truth_tokens]) if level == "word" else None
    mean_loss = loss_avg.val()
    return (all_loss, total_names, mean_loss, accuracy, bleu_score, norm_ED, word_ED, total_preds, total_labels, infer_time,
            length_of_data)

#!/usr/bin/env python
"""Inference entry point with the reference's CLI surface (api/infer.py:358-415), on the B200 engine.

    python api/infer.py --config cfg.yaml --csv_dir labels.tsv --data_dir images/ --log_path run.log --batch_size 256

Same flags (--config --csv_dir --start_idx --data_dir --amp --resizer --log_path --batch_size --num_workers
--strong_log --console), same YAML keys, same printed / logged summary lines (infer.py:332-354) and the same CSV
row layout (:230-236).  Differences, on purpose:
  * ``--batch_size`` is honoured: images are bucketed by their exact (H, W) — the reference's sampler contract
    (torch_dataset.py:46-66) — and every bucket runs batched through ``Model`` (the reference iterates the Dataset
    itself, so its batch is always 1: api/infer.py:94);
  * the model is ``doc2tex_b200.modules.build_model.Model`` (no CPU fallback);
  * ``--synthetic N`` runs N seeded synthetic images when no dataset is at hand (GPU box smoke test);
  * metrics: exact match and the edit distances are computed in-process (doc2tex_b200/engine_inferencing.py; the
    reference needs nltk / Levenshtein, which are not part of the hot path); BLEU is omitted;
  * the image preprocessing of the reference (utils/predict_utils.py::resize, per image on the host with PIL / cv2 /
    albumentations: optional down-sampling, crop to ink when `pad`, minmax_size, normalise) runs on the GPU for the whole
    list at once (doc2tex_b200/preprocess.py) and hands back the same-(H, W) buckets directly; --resizer (the
    learned width predictor, predict_utils.py:59-83) is not part of the accelerated path and is refused.
"""
from __future__ import annotations

import argparse
import csv
import os
import random
import sys
import time
from collections import defaultdict

import numpy as np
import torch
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from doc2tex_b200 import synth  # noqa: E402
from doc2tex_b200.engine_inferencing import edit_distance  # noqa: E402
from doc2tex_b200.modules.build_model import Model  # noqa: E402
from doc2tex_b200.modules.converter import builder  # noqa: E402

DELIMITER = "\t"  # doc2tex/data/data_const.py:18


def load_grey(path: str) -> np.ndarray:
    from PIL import Image
    return np.asarray(Image.open(path).convert("L"), dtype=np.uint8)     # predict_utils.py:16


def read_rows(args, config):
    if args.synthetic:
        imgs = synth.make_images(args.synthetic, 64, 256, seed=int(config.get("manualSeed", 1111)))
        return [(f"synthetic_{i}", None, imgs[i]) for i in range(args.synthetic)]
    rows = []
    with open(args.csv_dir, newline="") as f:
        for i, r in enumerate(csv.DictReader(f, delimiter=DELIMITER)):
            if i < args.start_idx:
                continue
            label = r.get("label", "")
            toks = str(label).strip().split() if config.get("token_level", "word") == "word" else list(str(label))
            rows.append((r["id"], toks, None))
    return rows


def run_infer(model, rows, converter, config, args):
    is_attn = "Attn" in config["Prediction"]["name"]
    max_len = config["batch_max_length"]
    device = config["device"]
    buckets = defaultdict(list)
    pre_t = 0.0
    kept = [(name, label, tensor) for name, label, tensor in rows
            if not (config.get("data_filtering", True) and label is not None and len(label) > max_len)]
    todo = [(name, label) for name, label, tensor in kept if tensor is None]
    for name, label, tensor in kept:
        if tensor is not None:                       # --synthetic: already a normalised (1, H, W) tensor
            buckets[tuple(tensor.shape[-2:])].append((name, label, tensor.to(device)))
    if todo:
        from doc2tex_b200.preprocess import Preprocessor
        prep = Preprocessor(model.engine, config)
        for lo in range(0, len(todo), 1024):          # the whole chunk is packed, uploaded and preprocessed in one go
            part = todo[lo: lo + 1024]
            t0 = time.time()
            greys = [load_grey(os.path.join(config["eval_data"], name)) for name, _ in part]
            for (H, W), (batch, idx) in prep(greys).items():
                for slot, i in enumerate(idx):
                    buckets[(H, W)].append((part[i][0], part[i][1], batch[slot]))
            torch.cuda.synchronize()
            pre_t += time.time() - t0
    n = n_correct = 0
    norm_ed = word_ed = 0.0
    infer_time = post_time = 0.0
    writer = fo = None
    if config.get("export_csv"):
        os.makedirs(os.path.dirname(args.export_path) or ".", exist_ok=True)
        fo = open(args.export_path, "wt" if args.start_idx == 0 else "at", newline="")
        writer = csv.writer(fo)
    # The batches of a bucket go through the pipelined recognizer (doc2tex_b200/pipeline.py): encode(i+1) overlaps decode(i),
    # consecutive batches are decoded in one call (about 2 560 greedy rows / 1 280 beam images per call), and
    # the (B, l, V) logits tensor that model(image, text) also returns — and this loop never read — is not materialised.
    # Per-batch results are identical to model(image, text, is_train=False, is_test=True) (tests: merged == sequential).
    from doc2tex_b200.pipeline import PipelinedRecognizer
    eng = model.engine
    beam_size = int(config.get("beam_size", 1) or 1)
    mode = "beam" if beam_size > 1 else "greedy"
    bs = int(config["batch_size"])
    steps = (max_len + 1) if is_attn else None          # Model.forward_decoder: the LSTM heads decode batch_max_length + 1 steps
    merge = max(1, (1280 if mode == "beam" else 2560) // max(1, bs))
    for (H, W), items in buckets.items():
        chunks = [items[lo: lo + bs] for lo in range(0, len(items), bs)]
        pipe = PipelinedRecognizer(eng, mode, beam_size, steps, is_test=True, return_logits=False, decode_merge=merge)
        torch.cuda.synchronize()
        t0 = time.time()
        results = list(pipe.run(torch.stack([c[2] for c in chunk]) for chunk in chunks))
        torch.cuda.synchronize()
        dt_img = (time.time() - t0) / max(1, len(items))
        infer_time += dt_img * len(items)
        for chunk, res in zip(chunks, results):
            B = len(chunk)
            dt = dt_img * B
            t0 = time.time()
            ids = res["ids"].cpu()
            if mode == "beam":                           # padded to the longest hypothesis: cut every row at its own length
                lens = res["lens"].cpu().tolist()
                pred_tokens = converter.detokenize([row[:n_] for row, n_ in zip(ids.tolist(), lens)])
            else:
                pred_tokens = converter.detokenize(ids)
            post_time += time.time() - t0
            for (name, label, _), pred in zip(chunk, pred_tokens):
                n += 1
                if label is None:
                    if writer:
                        writer.writerow([name, " ".join(pred), "", "", dt / B])
                    continue
                ed = edit_distance(pred, label)
                n_correct += int(pred == label)
                word_ed += 1.0 - ed / max(len(pred), len(label), 1)
                ps, ls = " ".join(pred), " ".join(label)
                norm_ed += 1.0 - edit_distance(ps, ls) / max(len(ps), len(ls), 1) if len(ps) + len(ls) < 2000 else 0.0
                if writer:
                    writer.writerow([name, ps, ls, int(pred == label), dt / B])
    if fo:
        fo.close()
    n = max(n, 1)
    mem = torch.cuda.max_memory_allocated() / 2 ** 20
    return n_correct / n * 100, norm_ed / n, word_ed / n, mem, infer_time, post_time, pre_t, n


def infer(config, args):
    if args.synthetic and not config.get("vocab"):
        config["character"] = synth.make_vocab()
    converter = builder.create_converter(config, config["device"])
    config["num_class"] = len(converter.character)
    model = Model(config)
    if config.get("saved_model") and os.path.exists(config["saved_model"]):
        ckpt = torch.load(config["saved_model"], map_location="cpu")
        model.load_state_dict(ckpt.get("model", ckpt), strict=True)
    elif not args.synthetic:
        raise FileNotFoundError(f"saved_model {config.get('saved_model')!r} not found")
    model = model.to(config["device"]).eval()
    params_num = sum(int(np.prod(p.size())) for p in model.parameters())
    rows = read_rows(args, config)
    with torch.no_grad():
        acc, norm_ED, word_ED, mem, infer_time, post_time, pre_time, n = run_infer(model, rows, converter, config, args)
    lines_print = [
        f"Acc: {acc:0.3f}", f"Norm Edit Distance: {norm_ED:0.5f}", f"Symbol Match (Word Edit Distance): {word_ED:0.5f}",
        f"Infer time {infer_time} s", f"Avg infer time {infer_time / float(n)} s", f"Preprocess time: {pre_time} s",
        f"Avg pre time: {pre_time / float(n)}", f"Postprocess time: {post_time} s", f"Avg post time {post_time / float(n)} s",
        f"Memory used: {mem} MB\n"]
    print("\n".join(lines_print))
    if not args.console:
        os.makedirs(os.path.dirname(args.log_path) or ".", exist_ok=True)
        with open(args.log_path, "w") as log:
            log.write(f"Trainable params num: {params_num}\n")
            log.write(f"Acc: {acc:0.3f}\n")
            log.write(f"Norm Edit Distance: {norm_ED:0.5f}\n")
            log.write(f"Symbol Match (Word Edit Distance): {word_ED:0.5f}\n")
            log.write(f"Total Infer Time: {infer_time} s\n")
            log.write(f"Avg Infer Time: {infer_time / float(n)} s\n")
            log.write(f"Postprocess time: {post_time} s\n")
            log.write(f"Avg post time {post_time / float(n)} s\n")
            log.write(f"Memory used: {mem} MB\n")
    return acc, n


def main(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument("--config", required=True, help="Path to config yaml file")
    parser.add_argument("--csv_dir", default="", help="Path to csv file (tab separated: id, label)")
    parser.add_argument("--start_idx", type=int, default=0, help="Index to start from csv file")
    parser.add_argument("--data_dir", default="", help="Path to image folder to infer")
    parser.add_argument("--amp", type=bool, default=False)
    parser.add_argument("--resizer", action="store_true", default=False)
    parser.add_argument("--log_path", required=True, help="Path to save evaluation result")
    parser.add_argument("--batch_size", required=True, type=int, help="test on batch or with single sample")
    parser.add_argument("--num_workers", type=int, default=-1, help="number of workers")
    parser.add_argument("--strong_log", action="store_true", default=False)
    parser.add_argument("--console", default=False)
    parser.add_argument("--synthetic", type=int, default=0, help="run N seeded synthetic 64x256 images instead of a dataset")
    parser.add_argument("--precision", default=None, choices=[None, "fp32", "tf32x3", "bf16x3", "bf16"])
    args = parser.parse_args(argv)
    if not args.synthetic and not (args.csv_dir and args.data_dir):
        parser.error("--csv_dir and --data_dir are required unless --synthetic N is given")
    config = yaml.load(open(args.config), Loader=yaml.FullLoader)
    config["batch_size"] = args.batch_size
    config["workers"] = args.num_workers
    config["use_amp"] = bool(args.amp)
    if args.resizer:
        parser.error("--resizer (the learned width predictor of predict_utils.py:59-83) is not on the accelerated path")
    config["use_resizer"] = args.resizer
    config["eval_data"] = args.data_dir
    if args.precision or args.amp:
        config.setdefault("engine", {})["precision"] = args.precision or "bf16"   # --amp maps to the bf16 mode
    seed = int(config.get("manualSeed", 1111) or 1111)
    random.seed(seed); np.random.seed(seed); torch.manual_seed(seed)
    if not torch.cuda.is_available():
        raise RuntimeError("doc2tex_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    config["num_gpu"] = torch.cuda.device_count()
    config["device"] = "cuda"
    args.export_path = os.path.splitext(args.log_path)[0] + ".csv"
    return infer(config, args)


if __name__ == "__main__":
    main()

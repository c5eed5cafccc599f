"""Synthetic configs, weights and images for the recognizer hot path.

The reference ships no checkpoint, no vocabulary and no golden vectors
(SURVEY.md §4, §8d), so parity tests and ``bench.py`` run on seeded synthetic
state.  Everything here is deterministic torch-CPU RNG, so the same
``state_dict`` can be rebuilt on the GPU box (where ``/root/reference`` does
not exist) and fed to the oracle and to the engine alike.

The key schema and shapes follow ``Model.state_dict()`` of the reference
(doc2tex/modules/build_model.py:7-34; per-module shapes: resnet.py:51-156,
vision_transformer.py:56-58,119-122, patchembed.py:111-113, tfm.py:48-65,
seq2seq.py:30-35, attention1D.py:121-134).  ``oracle/make_golden.py`` asserts
the schema against the live reference.
"""
from __future__ import annotations

import copy
import math
from collections import OrderedDict

import numpy as np
import torch

SEQ = "seqmodeler.SequenceModeling."
NET = SEQ + "patch_embed.backbone.ConvNet."
PRED = "predicter.Prediction."

# TFMLabelConverter ids (tfm_converter.py:8) / AttnLabelConverter ids (attn_converter.py:8)
TFM_PAD, TFM_GO, TFM_END, TFM_UNK = 0, 1, 2, 3
ATTN_GO, ATTN_END, ATTN_UNK = 0, 1, 2

N_SYMBOLS = 500  # synthetic vocabulary size


def _base_config() -> dict:
    # Same top-level keys as config/test.yaml:1-51 of the reference.
    return {
        "imgH": None,
        "imgW": None,
        "max_dimension": [192, 896],
        "min_dimension": [32, 32],
        "batch_max_length": 150,
        "rgb": False,
        "pad": False,
        "beam_size": 1,
        "mean": 0.5,
        "std": 0.5,
        "FeatureExtraction": {"name": "None"},
        "SequenceModeling": {
            "name": "ViT",
            "params": {
                "backbone": {
                    "name": "resnet",
                    "input_channel": 1,
                    "output_channel": 512,
                    "gcb": False,
                },
                "fix_embed": True,
                "input_channel": 1,
                "patching_style": "2d",
                "patch_size": [2, 2],
                "depth": 6,
                "num_heads": 8,
                "hidden_size": 256,
            },
        },
        "manualSeed": 1111,
        "device": "cpu",
    }


def make_config(head: str = "TFM", **overrides) -> dict:
    """HybridViT-TFM (configs 1,2,3,5) or the config/train.yaml default stack (config 4)."""
    cfg = _base_config()
    if head == "TFM":
        cfg["Prediction"] = {
            "name": "TFM",
            "params": {
                "d_model": 256,
                "nhead": 8,
                "num_decoder_layers": 4,
                "dim_feedforward": 1024,
                "dropout": 0.1,
                "max_seq_len": 150,
                "padding_idx": 0,
            },
        }
        cfg["num_class"] = N_SYMBOLS + 4
    elif head in ("Attnv2", "Attn"):
        cfg["Prediction"] = {
            "name": head,
            "params": {
                "seqmodel": "TFM" if head == "Attnv2" else "ViT",
                "input_size": 256,
                "hidden_size": 256,
                "kernel_size": 2,
                "kernel_dim": 128,
                "embed_target": True,
                "enc_init": True,
                "attn_type": "coverage",
                "method": "concat",
                "teacher_forcing": 1.0,
                "droprate": 0.25,
            },
        }
        cfg["num_class"] = N_SYMBOLS + 3
    else:
        raise ValueError(f"unknown head {head!r}")
    cfg.update(copy.deepcopy(overrides))
    return cfg


def make_vocab(n: int = N_SYMBOLS) -> list:
    return [f"\\tok{i}" for i in range(n)]


def backbone_out_hw(H: int, W: int) -> tuple:
    """Feature-map size of the ResNet stem (resnet.py:205-245)."""
    return H // 16 - 1, W // 4 + 1


def grid_hw(H: int, W: int) -> tuple:
    fh, fw = backbone_out_hw(H, W)
    return (fh + 1) // 2, (fw + 1) // 2


def sincos_pos_embed(dim: int, gh: int, gw: int) -> torch.Tensor:
    """Fixed 2-D sin-cos table with a zero cls row (mae_posembed.py:20-70).

    Half of the channels encode the *w* coordinate first, then *h* ("w goes
    first", mae_posembed.py:27), each half as [sin | cos] over dim/4 frequencies.
    """
    def one_axis(d, pos):
        omega = np.arange(d // 2, dtype=np.float32)
        omega /= d / 2.0
        omega = 1.0 / 10000 ** omega
        out = np.einsum("m,d->md", pos.reshape(-1), omega)
        return np.concatenate([np.sin(out), np.cos(out)], axis=1)

    ys, xs = np.meshgrid(np.arange(gh, dtype=np.float32), np.arange(gw, dtype=np.float32), indexing="ij")
    emb = np.concatenate([one_axis(dim // 2, xs), one_axis(dim // 2, ys)], axis=1)
    emb = np.concatenate([np.zeros([1, dim]), emb], axis=0)
    return torch.from_numpy(emb).float().unsqueeze(0)


def word_pos_enc(d_model: int = 256, max_len: int = 500) -> torch.Tensor:
    """`WordPosEnc.pe` buffer (position_encoding.py:7-22)."""
    pe = torch.zeros(max_len, d_model)
    position = torch.arange(0, max_len, dtype=torch.float)
    dim_t = torch.arange(0, d_model, 2, dtype=torch.float)
    div_term = 1.0 / (10000.0 ** (dim_t / d_model))
    ang = position[:, None] * div_term[None, :]
    pe[:, 0::2] = ang.sin()
    pe[:, 1::2] = ang.cos()
    return pe


class _Init:
    def __init__(self, seed: int):
        self.g = torch.Generator().manual_seed(seed)

    def normal(self, shape, std=1.0, mean=0.0):
        return torch.randn(shape, generator=self.g) * std + mean

    def uniform(self, shape, lo, hi):
        return torch.rand(shape, generator=self.g) * (hi - lo) + lo

    def trunc_normal(self, shape, std):
        return self.normal(shape, std).clamp_(-2 * std, 2 * std)

    def xavier(self, shape):
        bound = math.sqrt(6.0 / (shape[0] + shape[1]))
        return self.uniform(shape, -bound, bound)

    def linear_default(self, out_f, in_f):
        b = 1.0 / math.sqrt(in_f)
        return self.uniform((out_f, in_f), -b, b), self.uniform((out_f,), -b, b)


def make_state_dict(cfg: dict, seed: int = 1111, suppress_end: bool = False,
                    end_bias: float | None = None, sharpen: float = 1.0) -> "OrderedDict[str, torch.Tensor]":
    """Seeded random weights in the reference's ``Model.state_dict()`` schema.

    Distributions follow the reference initialisers (kaiming fan_out convs
    resnet.py:164-173, trunc-normal .02 ViT vision_transformer.py:230-237,
    xavier decoder tfm.py:28-30) with BN statistics / affine and the biases
    randomised so that BN folding and every bias path is exercised
    (SURVEY.md §8d "Weights").  ``suppress_end`` sets the END logit bias to
    -1e4 (the "full-length" decode regime of §8d); ``end_bias`` sets it to an
    arbitrary value instead (a positive bias makes random-init beams actually
    complete, which is what exercises the END bookkeeping of tools/beam.py:92-100).
    ``sharpen`` multiplies the output projection (TFM ``proj`` / Attn ``generator`` weight and bias) by a constant: the
    random-init heads emit nearly uniform distributions, where beam candidates tie to within a few fp32 ulps of the
    cumulative score; a sharpened head is peaked like a trained one, so every top-k decision has a margin that no fp32
    summation order can flip and beam traces can be asserted identical for 100 % of the images.
    """
    if suppress_end:
        end_bias = -1e4
    rng = _Init(seed)
    sd = OrderedDict()
    sp = cfg["SequenceModeling"]["params"]
    D = sp["hidden_size"]
    depth = sp["depth"]
    C = sp["backbone"]["output_channel"]
    cin = sp["backbone"]["input_channel"]
    max_h, max_w = cfg["max_dimension"] if not cfg.get("imgH") else (cfg["imgH"], cfg["max_dimension"][1])
    gh, gw = grid_hw(max_h, max_w)

    sd[SEQ + "cls_token"] = rng.trunc_normal((1, 1, D), 0.02)
    if cfg["SequenceModeling"]["params"].get("fix_embed", False):
        sd[SEQ + "pos_embed"] = sincos_pos_embed(D, gh, gw)                 # ViTEncoderV3 (vit_encoder.py:229-247)
    else:
        sd[SEQ + "pos_embed"] = rng.trunc_normal((1, gh * gw + 1, D), 0.02)  # ViTEncoder / V2: learnable (:43-49)
    for i in range(depth):
        p = f"{SEQ}blocks.{i}."
        sd[p + "norm1.weight"] = rng.normal((D,), 0.1, 1.0)
        sd[p + "norm1.bias"] = rng.normal((D,), 0.1)
        sd[p + "attn.qkv.weight"] = rng.trunc_normal((3 * D, D), 0.02)
        sd[p + "attn.qkv.bias"] = rng.normal((3 * D,), 0.02)
        sd[p + "attn.proj.weight"] = rng.trunc_normal((D, D), 0.02)
        sd[p + "attn.proj.bias"] = rng.normal((D,), 0.02)
        sd[p + "norm2.weight"] = rng.normal((D,), 0.1, 1.0)
        sd[p + "norm2.bias"] = rng.normal((D,), 0.1)
        sd[p + "mlp.fc1.weight"] = rng.trunc_normal((4 * D, D), 0.02)
        sd[p + "mlp.fc1.bias"] = rng.normal((4 * D,), 0.02)
        sd[p + "mlp.fc2.weight"] = rng.trunc_normal((D, 4 * D), 0.02)
        sd[p + "mlp.fc2.bias"] = rng.normal((D,), 0.02)
    sd[SEQ + "norm.weight"] = rng.normal((D,), 0.1, 1.0)
    sd[SEQ + "norm.bias"] = rng.normal((D,), 0.1)

    def conv(name, co, ci, kh, kw):
        std = math.sqrt(2.0 / (co * kh * kw))
        sd[NET + name + ".weight"] = rng.normal((co, ci, kh, kw), std)

    def bn(name, c):
        sd[NET + name + ".weight"] = rng.uniform((c,), 0.5, 1.5)
        sd[NET + name + ".bias"] = rng.normal((c,), 0.1)
        sd[NET + name + ".running_mean"] = rng.normal((c,), 0.1)
        sd[NET + name + ".running_var"] = rng.uniform((c,), 0.5, 1.5)
        sd[NET + name + ".num_batches_tracked"] = torch.tensor(0, dtype=torch.int64)

    def layer(name, cin_, planes, blocks):
        for b in range(blocks):
            ci = cin_ if b == 0 else planes
            conv(f"{name}.{b}.conv1", planes, ci, 3, 3)
            bn(f"{name}.{b}.bn1", planes)
            conv(f"{name}.{b}.conv2", planes, planes, 3, 3)
            bn(f"{name}.{b}.bn2", planes)
            if b == 0 and ci != planes:
                conv(f"{name}.{b}.downsample.0", planes, ci, 1, 1)
                bn(f"{name}.{b}.downsample.1", planes)

    c16, c8, c4, c2 = C // 16, C // 8, C // 4, C // 2
    conv("conv0_1", c16, cin, 3, 3); bn("bn0_1", c16)
    conv("conv0_2", c8, c16, 3, 3); bn("bn0_2", c8)
    layer("layer1", c8, c4, 1)
    conv("conv1", c4, c4, 3, 3); bn("bn1", c4)
    layer("layer2", c4, c2, 2)
    conv("conv2", c2, c2, 3, 3); bn("bn2", c2)
    layer("layer3", c2, C, 5)
    conv("conv3", C, C, 3, 3); bn("bn3", C)
    layer("layer4", C, C, 3)
    conv("conv4_1", C, C, 2, 2); bn("bn4_1", C)
    conv("conv4_2", C, C, 2, 2); bn("bn4_2", C)
    std = math.sqrt(2.0 / (D * 4))
    sd[SEQ + "patch_embed.proj.weight"] = rng.normal((D, C, 2, 2), std)
    sd[SEQ + "patch_embed.proj.bias"] = rng.normal((D,), 0.02)

    V = cfg["num_class"]
    pp = cfg["Prediction"]["params"]
    if cfg["Prediction"]["name"] == "TFM":
        d, F_ = pp["d_model"], pp["dim_feedforward"]
        emb = rng.normal((V, d))
        emb[pp["padding_idx"]] = 0.0
        sd[PRED + "word_embed.weight"] = emb
        sd[PRED + "pos_enc.pe"] = word_pos_enc(d)
        for l in range(pp["num_decoder_layers"]):
            p = f"{PRED}model.layers.{l}."
            for att in ("self_attn", "multihead_attn"):
                sd[p + att + ".in_proj_weight"] = rng.xavier((3 * d, d))
                sd[p + att + ".in_proj_bias"] = rng.normal((3 * d,), 0.02)
                sd[p + att + ".out_proj.weight"] = rng.xavier((d, d))
                sd[p + att + ".out_proj.bias"] = rng.normal((d,), 0.02)
            sd[p + "linear1.weight"] = rng.xavier((F_, d))
            sd[p + "linear1.bias"] = rng.normal((F_,), 0.02)
            sd[p + "linear2.weight"] = rng.xavier((d, F_))
            sd[p + "linear2.bias"] = rng.normal((d,), 0.02)
            for n in ("norm1", "norm2", "norm3"):
                sd[p + n + ".weight"] = rng.normal((d,), 0.1, 1.0)
                sd[p + n + ".bias"] = rng.normal((d,), 0.1)
        w, b = rng.linear_default(V, d)
        w, b = w * sharpen, b * sharpen
        if end_bias is not None:
            b[TFM_END] = end_bias
        sd[PRED + "proj.weight"], sd[PRED + "proj.bias"] = w, b
    else:
        hs, ins, kd, ks = pp["hidden_size"], pp["input_size"], pp["kernel_dim"], pp["kernel_size"]
        emb = rng.normal((V, ins))
        emb[ATTN_GO] = 0.0  # padding_idx = [GO] (seq2seq.py:32-35)
        sd[PRED + "embedding.weight"] = emb
        a = PRED + "attention_cell.attn."
        bnd = 1.0 / math.sqrt(2 * ks + 1)
        sd[a + "loc_conv.weight"] = rng.uniform((kd, 1, 2 * ks + 1), -bnd, bnd)
        sd[a + "loc_conv.bias"] = rng.uniform((kd,), -bnd, bnd)
        for n, (o, i) in (("loc_proj", (hs, kd)), ("query_proj", (hs, hs)), ("key_proj", (hs, ins)), ("score", (1, hs))):
            sd[a + n + ".weight"], sd[a + n + ".bias"] = rng.linear_default(o, i)
        r = PRED + "attention_cell.rnn."
        bnd = 1.0 / math.sqrt(hs)
        sd[r + "weight_ih"] = rng.uniform((4 * hs, ins + ins), -bnd, bnd)
        sd[r + "weight_hh"] = rng.uniform((4 * hs, hs), -bnd, bnd)
        sd[r + "bias_ih"] = rng.uniform((4 * hs,), -bnd, bnd)
        sd[r + "bias_hh"] = rng.uniform((4 * hs,), -bnd, bnd)
        w, b = rng.linear_default(V, hs)
        w, b = w * sharpen, b * sharpen
        if end_bias is not None:
            b[ATTN_END] = end_bias
        sd[PRED + "attention_cell.generator.weight"] = w
        sd[PRED + "attention_cell.generator.bias"] = b
        for n in ("proj_init_h", "proj_init_c"):
            sd[PRED + n + ".weight"], sd[PRED + n + ".bias"] = rng.linear_default(hs, ins)
    return sd


def make_images(B: int, H: int = 64, W: int = 256, seed: int = 2024) -> torch.Tensor:
    """im2latex-shaped synthetic batch: white (+1) page, 3 % stroke pixels ~ U(-1,1).

    Matches what the reference feeds the model after ``Normalize(mean=.5, std=.5)``
    (SURVEY.md §8d "Inputs"): dense ``(B,1,H,W)`` fp32 in [-1, 1], H and W
    multiples of 32 (data_utils.py:10-47).  Image ``i`` uses seed ``seed + i`` so a
    shard of a batch equals the same rows of the full batch.
    """
    assert H % 32 == 0 and W % 32 == 0, "H, W must be multiples of 32"
    out = torch.empty(B, 1, H, W)
    for i in range(B):
        g = torch.Generator().manual_seed(seed + i)
        ink = torch.rand(H, W, generator=g) < 0.03
        val = torch.rand(H, W, generator=g) * 2.0 - 1.0
        out[i, 0] = torch.where(ink, val, torch.ones(()))
    return out

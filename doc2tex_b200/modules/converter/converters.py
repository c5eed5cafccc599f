"""Host-side label converters with the reference's surface and ids.

Mirrors doc2tex/modules/converter/{attn_converter.py:5-77, tfm_converter.py:5-82, builder.py:5-15}:
special-token ids ([GO]=0,[s]=1,[UNK]=2 for Attn*; [PAD]=0,[GO]=1,[s]=2,[UNK]=3 for TFM),
``encode`` (text -> padded id matrix + lengths), ``decode`` (ids -> joined string, NOT cut at
"[s]"; callers cut, api/infer.py:186-187) and ``detokenize`` (cut at "[s]").
"""
from __future__ import annotations

from typing import List, Sequence

import torch


class _Converter:
    list_token: List[str] = []
    pad_token: str = ""

    def __init__(self, character: Sequence[str], device):
        self.character = list(self.list_token) + list(character)
        self.device = device
        self.dict = {tok: i for i, tok in enumerate(self.character)}
        self.ignore_idx = self.dict[self.pad_token]

    @classmethod
    def _id(cls, tok: str) -> int:
        return cls.list_token.index(tok)

    @classmethod
    def START(cls) -> int:
        return cls._id("[GO]")

    @classmethod
    def END(cls) -> int:
        return cls._id("[s]")

    @classmethod
    def UNK(cls) -> int:
        return cls._id("[UNK]")

    def encode(self, text, batch_max_length: int = 25):
        """Rows: [GO], tokens..., [s], padding; width batch_max_length + 2."""
        width = batch_max_length + 2
        limit = batch_max_length + 1
        out = torch.full((len(text), width), self.ignore_idx, dtype=torch.long)
        lengths = []
        unk = self.dict["[UNK]"]
        for row, sample in enumerate(text):
            toks = list(sample)
            lengths.append(len(toks) + 1)
            if len(toks) > limit:
                toks = toks[: limit - 1]
            ids = [self.dict.get(t, unk) for t in toks] + [self.dict["[s]"]]
            out[row, 0] = self.dict["[GO]"]
            out[row, 1:1 + len(ids)] = torch.tensor(ids, dtype=torch.long)
        return out.to(self.device), torch.tensor(lengths, dtype=torch.int32).to(self.device)

    def decode(self, text_index, token_level: str = "word"):
        sep = " " if token_level == "word" else ""
        rows = text_index.tolist() if hasattr(text_index, "tolist") else text_index
        return [sep.join(self.character[i] for i in row) for row in rows]

    def detokenize(self, token_ids):
        rows = token_ids.tolist() if hasattr(token_ids, "tolist") else token_ids
        end = self.dict["[s]"]
        out = []
        for row in rows:
            cut = row.index(end) if end in row else len(row)
            out.append([self.character[i] for i in row[:cut]])
        return out


class AttnLabelConverter(_Converter):
    list_token = ["[GO]", "[s]", "[UNK]"]
    pad_token = "[GO]"


class TFMLabelConverter(_Converter):
    list_token = ["[PAD]", "[GO]", "[s]", "[UNK]"]
    pad_token = "[PAD]"

    @classmethod
    def PAD(cls) -> int:
        return cls._id("[PAD]")


def create_converter(config: dict, device):
    """builder.py:5-15: reads ``config['vocab']`` (one token per line) into ``config['character']``."""
    if config.get("vocab"):
        with open(config["vocab"], "r") as f:
            config["character"] = [line.strip() for line in f.readlines()]
    name = config["Prediction"]["name"]
    if "Attn" in name:
        return AttnLabelConverter(config["character"], device)
    if name in ("TFM", "MS_TFM"):
        return TFMLabelConverter(config["character"], device)
    raise ValueError(f"no converter for Prediction.name={name!r}")

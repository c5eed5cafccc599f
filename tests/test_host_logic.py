"""CPU: host-side logic — converters, synthetic weights, config translation, Model surface, sharding + gather (gloo)."""
import os
import subprocess
import sys
import textwrap

import pytest
import torch

from doc2tex_b200 import synth
from doc2tex_b200.modules.converter import AttnLabelConverter, TFMLabelConverter, create_converter

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_converter_ids_and_roundtrip(tmp_path):
    vocab = synth.make_vocab(10)
    t = TFMLabelConverter(vocab, "cpu")
    a = AttnLabelConverter(vocab, "cpu")
    assert (t.PAD(), t.START(), t.END(), t.UNK()) == (0, 1, 2, 3)       # tfm_converter.py:8
    assert (a.START(), a.END(), a.UNK()) == (0, 1, 2)                   # attn_converter.py:8
    ids, lens = t.encode([vocab[:3], ["nope"] + vocab[:1]], batch_max_length=6)
    assert ids.shape == (2, 8) and lens.tolist() == [4, 3]
    assert ids[0].tolist() == [1, 4, 5, 6, 2, 0, 0, 0] and ids[1, 1].item() == 3
    assert t.decode(ids[:, 1:5])[0] == f"{vocab[0]} {vocab[1]} {vocab[2]} [s]"
    assert t.detokenize(ids[:, 1:])[0] == vocab[:3]
    ids_a, _ = a.encode([vocab[:2]], batch_max_length=4)
    assert ids_a[0].tolist() == [0, 3, 4, 1, 0, 0]
    # over-long labels are truncated to batch_max_length tokens + [s]
    ids_l, lens_l = t.encode([vocab * 2], batch_max_length=5)
    assert ids_l.shape == (1, 7) and ids_l[0, -1].item() == 2 and lens_l.item() == 21
    p = tmp_path / "vocab.txt"
    p.write_text("\n".join(vocab) + "\n")
    cfg = {"vocab": str(p), "Prediction": {"name": "TFM"}}
    c = create_converter(cfg, "cpu")
    assert isinstance(c, TFMLabelConverter) and cfg["character"] == vocab
    cfg["Prediction"]["name"] = "Attnv2"
    assert isinstance(create_converter(cfg, "cpu"), AttnLabelConverter)


def test_synth_state_dict_is_deterministic_and_shaped():
    cfg = synth.make_config("TFM")
    a = synth.make_state_dict(cfg, seed=7)
    b = synth.make_state_dict(cfg, seed=7)
    assert len(a) == 346 and all(torch.equal(a[k], b[k]) for k in a)
    assert a[synth.SEQ + "pos_embed"].shape == (1, 679, 256)
    assert a[synth.PRED + "proj.weight"].shape == (504, 256)
    assert synth.make_state_dict(cfg, seed=7, suppress_end=True)[synth.PRED + "proj.bias"][2].item() == -1e4
    x = synth.make_images(3, 64, 256, seed=5)
    assert x.shape == (3, 1, 64, 256) and x.max() <= 1 and x.min() >= -1
    assert torch.equal(x[1:], synth.make_images(2, 64, 256, seed=6))    # a shard equals the rows of the full batch
    assert synth.grid_hw(64, 256) == (2, 33) and synth.grid_hw(192, 896) == (6, 113)


def test_config_translation_and_rejections():
    from doc2tex_b200.engine import EngineError, config_from_opt
    c = config_from_opt(synth.make_config("TFM"), "bf16x3")
    assert (c.hidden, c.depth, c.heads, c.max_tokens, c.head, c.vocab, c.dec_layers, c.dec_ff, c.precision) == \
           (256, 6, 8, 679, 1, 504, 4, 1024, 2)
    c = config_from_opt(synth.make_config("Attnv2"))
    assert (c.head, c.vocab, c.attn_hidden, c.attn_kernel_dim, c.attn_kernel_size, c.max_seq_len) == (2, 503, 256, 128, 2, 150)
    c = config_from_opt(synth.make_config("Attn"))          # base Attention head: same decoder, cls token attended
    assert (c.head, c.vocab, c.attn_hidden) == (3, 503, 256)
    ok = synth.make_config("TFM")       # ViTEncoder / ViTEncoderV2 (learnable pos_embed) are accepted
    ok["SequenceModeling"]["params"]["fix_embed"] = False
    assert config_from_opt(ok).max_tokens == 679
    bad = synth.make_config("TFM")
    bad["SequenceModeling"]["params"]["patching_style"] = "1d"
    with pytest.raises(EngineError):
        config_from_opt(bad)
    bad = synth.make_config("TFM")
    bad["Prediction"]["name"] = "CTC"
    with pytest.raises(EngineError):
        config_from_opt(bad)


def test_model_surface_without_gpu_fails_loudly():
    from doc2tex_b200.engine import EngineError
    from doc2tex_b200.modules.build_model import Model
    cfg = synth.make_config("TFM")
    m = Model(cfg)
    sd = synth.make_state_dict(cfg, seed=3)
    assert set(m.state_dict().keys()) == set(sd.keys())                # reference schema
    m.load_state_dict(sd, strict=True)
    assert cfg["Prediction"]["params"]["num_classes"] == 504            # build_pred.py:16-26 side effect kept
    with pytest.raises(EngineError):
        m.train()
    if not torch.cuda.is_available():
        with pytest.raises(EngineError):
            m.forward_encoder(torch.zeros(1, 1, 64, 256))               # no silent CPU fallback


def test_shard_range_partitions():
    from doc2tex_b200.dist import shard_range
    for n, w in [(256, 8), (10, 4), (3, 8), (1, 1)]:
        spans = [shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def test_gather_results_gloo_world2(tmp_path):
    """N>1 host path on CPU: two gloo ranks with ragged shards, one all-gather, batch order restored."""
    script = tmp_path / "w.py"
    script.write_text(textwrap.dedent(f"""
        import os, sys
        sys.path.insert(0, {ROOT!r})
        import torch, torch.distributed as dist
        from doc2tex_b200.dist import shard_range, gather_results
        dist.init_process_group("gloo")
        r, w = dist.get_rank(), dist.get_world_size()
        n, T = 5, 7
        full = torch.arange(n * T, dtype=torch.int64).reshape(n, T)
        lens = torch.arange(n, dtype=torch.int32) + 1
        sc = -torch.arange(n, dtype=torch.float32) - 0.25
        lo, hi = shard_range(n, r, w)
        ids, l, s = gather_results(full[lo:hi].clone(), lens[lo:hi].clone(), sc[lo:hi].clone(), n_total=n)
        assert torch.equal(ids, full) and torch.equal(l, lens) and torch.equal(s, sc), (ids, l, s)
        ids2, _, _ = gather_results(full[lo:hi].clone())
        assert torch.equal(ids2, full)
        dist.destroy_process_group()
        print("rank", r, "ok")
    """))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29571", str(script)],
                         capture_output=True, text=True, env=env, timeout=240)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("ok") == 2


def test_merged_decode_split_recovers_per_batch_steps():
    """PipelinedRecognizer._split: a merged greedy decode runs until EVERY row of EVERY merged batch has emitted END;
    each batch must still get the reference's own step count = first step at which all of ITS rows have ended
    (tfm.py:138-140), or the merged count when one of its rows never ended."""
    import torch
    from doc2tex_b200.pipeline import PipelinedRecognizer

    class FakeEngine:
        end_id = 2
        device = "cpu"

    pipe = PipelinedRecognizer.__new__(PipelinedRecognizer)
    pipe.eng, pipe.mode, pipe.is_test = FakeEngine(), "greedy", True
    END = 2
    ids = torch.full((5, 9), 7, dtype=torch.int64)
    ids[0, 2] = END; ids[1, 4] = END            # batch 0 (rows 0-1): done after 5 steps
    ids[2, 0] = END; ids[2, 3] = END            # batch 1 (row 2): done after 1 step (a later END does not matter)
    ids[3, 8] = END                              # batch 2 (rows 3-4): row 4 never ends -> merged count
    logits = torch.arange(5 * 9 * 3, dtype=torch.float32).reshape(5, 9, 3)
    out = pipe._split({"ids": ids, "logits": logits, "steps": 9}, [2, 1, 2])
    assert [o["steps"] for o in out] == [5, 1, 9]
    assert torch.equal(out[0]["ids"], ids[0:2, :5]) and torch.equal(out[1]["ids"], ids[2:3, :1])
    assert torch.equal(out[2]["ids"], ids[3:5]) and torch.equal(out[1]["logits"], logits[2:3, :1])
    pipe.mode = "beam"
    res = {"ids": ids, "lens": torch.arange(5), "scores": torch.arange(5.0), "steps": 9}
    outb = pipe._split(res, [2, 1, 2])
    assert torch.equal(outb[1]["lens"], torch.tensor([2])) and torch.equal(outb[2]["ids"], ids[3:5])


def test_infer_cli_keeps_the_reference_flags():
    """api/infer.py mirrors the reference CLI (api/infer.py:359-387 of the reference): every flag it defines must stay
    accepted; --precision and --synthetic are this repo's additions."""
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(root, "api", "infer.py")).read()
    ours = set(re.findall(r"[\"'](--[a-z_A-Z0-9]+)[\"']", src))
    reference = {"--amp", "--batch_size", "--config", "--console", "--csv_dir", "--data_dir", "--log_path", "--num_workers",
                 "--resizer", "--start_idx", "--strong_log"}
    assert reference <= ours
    assert ours - reference == {"--precision", "--synthetic"}


def test_converters_match_live_reference_fixture():
    """a12 / f2: encode / decode / detokenize and the vectorised decode_cut against outputs of the LIVE reference converters
    (tests/golden/converters.json, minted by oracle/make_golden.py::converter_case), including the callers' cut quirks:
    a blank kept before "[s]" with word-level joining, the last character lost when a row has no "[s]"."""
    import json
    from doc2tex_b200.engine_inferencing import decode_cut, first_end_cut
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "converters.json")))
    for name, cls in (("TFM", TFMLabelConverter), ("Attn", AttnLabelConverter)):
        rec = g["cases"][name]
        conv = cls(g["vocab"], "cpu")
        assert {k: conv.dict[k] for k in rec["special"]} == rec["special"]
        ids = torch.tensor(rec["ids"])
        enc, lens = conv.encode(rec["labels"], batch_max_length=12)
        assert enc.tolist() == rec["encode"] and lens.tolist() == rec["encode_len"]
        assert conv.detokenize(ids) == rec["detokenize"]
        cut = first_end_cut(ids, conv.dict["[s]"])
        for level in ("word", "char"):
            assert conv.decode(ids, level) == rec[f"decode_{level}"]
            strings, tokens = decode_cut(conv, ids, level, cut)
            assert strings == rec[f"cut_{level}"], (name, level)
            assert tokens == rec["detokenize"]
        # a vocabulary token that contains "[s]" switches decode_cut to the reference's string search
        odd = cls(["x[s]y"] + g["vocab"][1:], "cpu")
        first_tok = len(cls.list_token)
        row = torch.tensor([[first_tok, first_tok + 1, odd.dict["[s]"], first_tok + 2]])
        full = odd.decode(row)[0]
        assert decode_cut(odd, row)[0] == [full[: full.find("[s]")]]


def test_string_metrics_small_cases():
    from doc2tex_b200.engine_inferencing import corpus_bleu, edit_distance, single_ed, squeeze_latex_whitespace, word_ned
    assert edit_distance("kitten", "sitting") == 3 and edit_distance("", "abc") == 3 and edit_distance("abc", "abc") == 0
    assert edit_distance(["a", "b", "c"], ["a", "c"]) == 1
    assert single_ed("", "x") == 0 and abs(single_ed("abcd", "abed") - 0.75) < 1e-12
    assert abs(word_ned("a b c", "a x c") - (1 - 1 / 3)) < 1e-12 and word_ned("", "a") == 0.0
    assert corpus_bleu([["a", "b", "c", "d", "e"]], [[["a", "b", "c", "d", "e"]]]) == pytest.approx(1.0)
    assert corpus_bleu([["a", "b"]], [[["c", "d"]]]) == 0.0
    assert squeeze_latex_whitespace("x ^ { 2 } + \\mathrm { d } y") == "x^{2}+\\mathrm{d}y"
    assert squeeze_latex_whitespace("\\sin x") == "\\sin x"


def test_bench_auto_merge_divides_the_timed_steps():
    """bench.py hands `decode_merge` encoded batches to one decode call; the count must divide the timed steps (a partial
    group would cost a whole decode chain inside the bracket)."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    assert bench.auto_merge("greedy", 20) == 10 and bench.auto_merge("greedy", 16) == 8 and bench.auto_merge("greedy", 5) == 5
    assert bench.auto_merge("beam", 20) == 5 and bench.auto_merge("beam", 5) == 5 and bench.auto_merge("beam", 4) == 4
    # a rank's strong-scaling shard (256 / N images): more batches per call, same rows
    assert bench.auto_merge("greedy", 80, 32) == 80 and bench.auto_merge("beam", 40, 32) == 40 and bench.auto_merge("greedy", 20, 128) == 20
    for mode in ("greedy", "beam"):
        for batch in (32, 256, 1024):
            tgt = bench.merge_target(mode, batch)
            for k in range(1, 41):
                m = bench.auto_merge(mode, k, batch)
                assert 1 <= m <= k and m <= max(1, tgt * 5 // 4)
                assert k % m == 0 or m == min(k, tgt)

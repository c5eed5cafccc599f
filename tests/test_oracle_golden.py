"""CPU: the oracle restatement (oracle/oracle_model.py) against the golden vectors minted from the LIVE
reference by oracle/make_golden.py.  This is what pins the oracle (the reference has no tests)."""
import numpy as np
import pytest
import torch

from doc2tex_b200 import synth
from oracle import oracle_model as om
from tests.util import end_bias_of, load_golden, state_dict_for


@pytest.fixture(scope="module")
def threads():
    torch.set_num_threads(min(8, torch.get_num_threads()))


def _ctx(case, H=64, W=256, B=2):
    g = load_golden(case)
    cfg, sd = state_dict_for("TFM", end_bias_of(g))
    img = synth.make_images(B, H, W, seed=2024)
    ctx, grid, pad = om.encoder_forward(sd, img)
    return g, sd, ctx, grid, pad


def test_encoder_matches_reference_golden(threads):
    g, sd, ctx, grid, pad = _ctx("tfm_64x256_natural")
    assert tuple(grid) == tuple(g["grid"]) == (2, 33) and tuple(pad) == tuple(g["pad"]) == (1, 1)
    ref = torch.from_numpy(g["ctx"])
    assert (ctx - ref).abs().max().item() <= 2e-5 * ref.abs().max().item()


def test_encoder_other_size_matches_reference_golden(threads):
    g, sd, ctx, grid, pad = _ctx("tfm_96x384_full", 96, 384, 1)
    assert tuple(grid) == tuple(g["grid"]) and ctx.shape[1] == 148
    ref = torch.from_numpy(g["ctx"])
    assert (ctx - ref).abs().max().item() <= 2e-5 * ref.abs().max().item()


@pytest.mark.parametrize("case", ["tfm_64x256_full", "tfm_64x256_end15", "tfm_64x256_end20"])
def test_tfm_greedy_matches_reference_golden(threads, case):
    g, sd, ctx, _, _ = _ctx(case)
    ids, logits, gen = om.TFMHead(sd, max_seq_len=150).greedy(ctx, is_test=True)
    assert torch.equal(ids, torch.from_numpy(g["greedy_ids"]))          # the reference's preds_index
    assert torch.equal(gen, torch.from_numpy(g["greedy_gen"]))          # the generated tokens
    ref = torch.from_numpy(g["greedy_logits"])
    for j, s in enumerate(g["greedy_logit_steps"].tolist()):
        assert (logits[:, s] - ref[:, j]).abs().max().item() <= 1e-4


@pytest.mark.parametrize("case,img", [("tfm_64x256_end15", 0), ("tfm_64x256_end20", 1)])
def test_tfm_beam_matches_reference_golden(threads, case, img):
    g, sd, ctx, _, _ = _ctx(case)
    tr = []
    seq, score = om.TFMHead(sd, max_seq_len=150).beam(ctx[img:img + 1], 5, trace=tr)
    n = int(g["beam_len"][img])
    assert seq == g["beam_seq"][img, :n].tolist()
    assert abs(score - float(g["beam_score"][img])) <= 1e-3
    for t, (p, w, s) in enumerate(tr[:8]):
        k = len(p)
        assert p == g["beam_parents"][img, t, :k].tolist() and w == g["beam_words"][img, t, :k].tolist()


@pytest.mark.parametrize("case", ["attnv2_64x256_full", "attnv2_64x256_end"])
def test_attnv2_greedy_matches_reference_golden(threads, case):
    g = load_golden(case)
    cfg, sd = state_dict_for("Attnv2", end_bias_of(g))
    img = synth.make_images(2, 64, 256, seed=2024)
    ctx, _, _ = om.encoder_forward(sd, img)
    ids, probs = om.AttnV2Head(sd).greedy(ctx, 150, True)
    assert torch.equal(ids, torch.from_numpy(g["ids"]))
    ref = torch.from_numpy(g["logits"])
    for j, s in enumerate(g["logit_steps"].tolist()):
        assert (probs[:, s] - ref[:, j]).abs().max().item() <= 1e-4


@pytest.mark.parametrize("case,img", [("attnv2_beam_64x256_end04", 1), ("attnv2_beam_64x256_end05", 0),
                                      ("attnv2_beam_64x256_end30", 1)])
def test_attnv2_beam_matches_reference_golden(threads, case, img):
    """AttentionV2.forward_beam restatement (seq2seq_v2.py:12-174) against the live reference's outputs: a beam that
    shrinks over 124 steps and ends on the live hypothesis (Q11), one that completes everything by step 10, one that
    ends after two steps; top-k (parent, word) per step included."""
    g = load_golden(case)
    cfg, sd = state_dict_for("Attnv2", end_bias_of(g))
    x = synth.make_images(int(g["beam_len"].shape[0]), 64, 256, seed=2024)
    ctx, _, _ = om.encoder_forward(sd, x)
    tr = []
    seq, score = om.AttnV2Head(sd).beam(ctx[img:img + 1], 5, 150, trace=tr)
    n = int(g["beam_len"][img])
    assert seq == g["beam_seq"][img, :n].tolist()
    assert abs(score - float(g["beam_score"][img])) <= 1e-3 * max(1.0, abs(score))
    assert len(tr) == int(g["beam_steps"][img])
    for t, (p, w, _) in enumerate(tr):
        k = len(p)
        assert p == g["beam_parents"][img, t, :k].tolist() and w == g["beam_words"][img, t, :k].tolist()


@pytest.mark.parametrize("case,interp", [("vit_interp_posembed", True), ("vit_v2_posembed", False)])
def test_encoder_variants_match_reference_golden(threads, case, interp):
    """ViTEncoder (bicubic-interpolated learnable pos_embed, vit_encoder.py:58-118) and ViTEncoderV2 (prefix slice,
    :205-226) against the live reference's ctx."""
    g = load_golden(case)
    cfg = synth.make_config("TFM")
    cfg["SequenceModeling"]["params"].update(fix_embed=False, interpolate_embed=interp)
    sd = synth.make_state_dict(cfg, seed=1111, end_bias=None)
    max_grid = synth.grid_hw(*cfg["max_dimension"])
    for (H, W) in [(64, 256), (96, 384)]:
        ctx, _, _ = om.encoder_forward(sd, synth.make_images(1, H, W, seed=2024),
                                       pos_mode="interpolate" if interp else "prefix", max_grid=max_grid)
        assert (ctx - torch.from_numpy(g[f"ctx_{H}x{W}"])).abs().max().item() <= 1e-5


def test_attn_base_head_matches_reference_golden(threads):
    """Prediction.name 'Attn' (seq2seq.py): same loops as Attnv2 but the cls token is attended too."""
    g = load_golden("attn_64x256_full")
    cfg = synth.make_config("Attn")
    sd = synth.make_state_dict(cfg, seed=1111, end_bias=end_bias_of(g))
    ctx, _, _ = om.encoder_forward(sd, synth.make_images(2, 64, 256, seed=2024))
    ids, probs = om.AttnV2Head(sd, include_cls=True).greedy(ctx, 150, True)
    assert torch.equal(ids, torch.from_numpy(g["ids"]))
    gb = load_golden("attn_beam_64x256_end05")
    sd = synth.make_state_dict(cfg, seed=1111, end_bias=end_bias_of(gb))
    ctx, _, _ = om.encoder_forward(sd, synth.make_images(2, 64, 256, seed=2024))
    seq, score = om.AttnV2Head(sd, include_cls=True).beam(ctx[:1], 5, 150)
    assert seq == gb["beam_seq"][0, : int(gb["beam_len"][0])].tolist()
    assert abs(score - float(gb["beam_score"][0])) <= 1e-3 * abs(score)


def test_oracle_edge_cases():
    """Quirks the engine must share: causal mask values, prefix pos-embed slice, -inf pooling pad, tie rule."""
    m = om.TFMHead.causal_mask(4)
    assert m[2, 1] == 0 and m[1, 2] == float("-inf") and m[3, 3] == 0
    x = torch.zeros(1, 1, 2, 3)
    x[0, 0, :, 0] = -5.0
    y = torch.nn.functional.max_pool2d(x, 2, (2, 1), (0, 1))
    assert y.shape == (1, 1, 1, 4) and y[0, 0, 0, 0] == -5.0          # padding is -inf, not 0 (quirk Q1)
    assert int(torch.argmax(torch.tensor([1.0, 3.0, 3.0]))) == 1       # lowest index wins
    pe = synth.sincos_pos_embed(256, 6, 113)
    assert pe.shape == (1, 679, 256) and float(pe[0, 0].abs().max()) == 0.0


def test_preprocess_oracle_matches_reference_outputs():
    """f3: the numpy restatement of pad / minmax_size / INTER_AREA / Pillow LANCZOS against the REFERENCE's outputs stored in
    tests/golden/preprocess.npz (ink boxes and 8-bit images bit for bit)."""
    import numpy as np
    from oracle import preprocess_oracle as po
    from tests.util import load_golden
    g = load_golden("preprocess")
    maxd, mind = g["max_dimension"].tolist(), g["min_dimension"].tolist()
    shrunk = 0
    for i in range(int(g["a_count"])):
        a = g[f"a{i}_img"]
        padded, box = po.pad_to_ink(a)
        assert list(box) == g[f"a{i}_box"].tolist()
        out = po.minmax_size(padded, maxd, mind)
        assert np.array_equal(out, g[f"a{i}_u8"]), i
        shrunk += out.shape != padded.shape
    assert shrunk >= 1      # at least one case goes through the LANCZOS shrink
    for i in range(int(g["b_count"])):
        small = po.area_downsample(g[f"b{i}_img"], 2)
        assert np.array_equal(po.minmax_size(small, maxd, mind), g[f"b{i}_u8"]), i
    x = po.normalize(np.array([[0, 127, 128, 255]], dtype=np.uint8))
    assert np.allclose(x, [[-1.0, -1 / 255, 1 / 255, 1.0]], atol=1e-7)
    # sizes on which the reference itself dies (get_divisible_size): the oracle implements the evident intent
    assert po.minmax_plan(500, 1800, maxd, mind) == ((288, 960), None)      # 266.67 -> up to 288
    assert po.minmax_plan(20, 90, maxd, mind) == (None, (32, 160))

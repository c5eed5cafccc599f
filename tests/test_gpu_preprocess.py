"""f3: GPU preprocessing (d2t_prep_measure / d2t_prep_render behind doc2tex_b200.preprocess.Preprocessor) against the
outputs of the reference's own pad / minmax_size / cv2.INTER_AREA (tests/golden/preprocess.npz, minted by
oracle/make_golden.py::preprocess_case) and against the numpy oracle: ink boxes bit-exact, 8-bit images bit-exact,
normalised pixels within 1e-6."""
import numpy as np
import pytest
import torch

from doc2tex_b200 import synth
from tests.util import load_golden, state_dict_for

pytestmark = pytest.mark.gpu

OPT = {"max_dimension": [448, 960], "min_dimension": [32, 32], "mean": 0.5, "std": 0.5, "rgb": False, "imgH": None}


def _engine():
    from doc2tex_b200.engine import Engine
    cfg, _ = state_dict_for("TFM", None)
    return Engine(cfg, "cuda:0")


def _collect(buckets, n):
    out = [None] * n
    for (h, w), (batch, idx) in buckets.items():
        assert batch.shape == (len(idx), 1, h, w) and batch.dtype == torch.float32
        for slot, i in enumerate(idx):
            out[i] = batch[slot, 0].cpu().numpy()
    return out


def test_crop_to_ink_matches_reference_outputs(built_lib):
    from doc2tex_b200.preprocess import Preprocessor
    from oracle import preprocess_oracle as po
    g = load_golden("preprocess")
    n = int(g["a_count"])
    imgs = [g[f"a{i}_img"] for i in range(n)]
    e = _engine()
    prep = Preprocessor(e, dict(OPT, pad=True, downsample=None))
    got = _collect(prep(imgs), n)
    for i in range(n):
        assert prep.last_boxes[i].tolist() == g[f"a{i}_box"].tolist(), i                # integer crop box, bit-exact
        ref = po.normalize(g[f"a{i}_u8"], 0.5, 0.5)
        assert got[i].shape == ref.shape, (i, got[i].shape, ref.shape)
        assert np.abs(got[i] - ref).max() <= 1e-6, i
        u8 = np.rint(got[i] * 127.5 + 127.5).astype(np.uint8)
        assert np.array_equal(u8, g[f"a{i}_u8"]), i                                      # the 8-bit image, bit-exact
    # images of equal output size share a bucket
    sizes = {}
    for (h, w), (_, idx) in prep(imgs).items():
        sizes[(h, w)] = idx
    assert sum(len(v) for v in sizes.values()) == n
    e.close()


def test_downsample_path_matches_reference_outputs(built_lib):
    from doc2tex_b200.preprocess import Preprocessor
    from oracle import preprocess_oracle as po
    g = load_golden("preprocess")
    n = int(g["b_count"])
    imgs = [g[f"b{i}_img"] for i in range(n)]
    e = _engine()
    got = _collect(Preprocessor(e, dict(OPT, pad=False, downsample=2))(imgs), n)
    for i in range(n):
        ref = po.normalize(g[f"b{i}_u8"], 0.5, 0.5)
        assert got[i].shape == ref.shape and np.abs(got[i] - ref).max() <= 1e-6, i
    e.close()


def test_shrink_and_canvas_match_oracle(built_lib):
    """Sizes where the reference's minmax_size itself raises UnboundLocalError (get_divisible_size, data_utils.py:50-59; see
    oracle/make_golden.py::preprocess_case): the Pillow-exact LANCZOS shrink and the white canvas against the numpy oracle
    (which equals Pillow / the reference bit for bit wherever the reference survives — tests/test_oracle_golden.py)."""
    from doc2tex_b200.preprocess import PreprocessError, Preprocessor
    from oracle import make_golden_helpers as mh
    from oracle import preprocess_oracle as po
    specs = [(600, 700, True), (300, 1500, False), (470, 500, True), (500, 1800, True), (20, 90, True), (31, 70, False), (64, 256, True)]
    imgs = [mh.synth_crop(h, w, 3000 + k, dark) for k, (h, w, dark) in enumerate(specs)]
    e = _engine()
    for pad in (True, False):
        opt = dict(OPT, pad=pad, downsample=None)
        use = imgs if pad else [a for a in imgs if a.shape in ((600, 700), (500, 1800), (20, 90), (64, 256), (470, 500))]
        if not pad:   # without the crop the raw sizes must already lead to /32 outputs
            use = [a for a in use if all(v % 32 == 0 for v in po.minmax_size(a, [448, 960], [32, 32]).shape)]
        prep = Preprocessor(e, opt)
        got = _collect(prep(use), len(use))
        for i, a in enumerate(use):
            ref = po.preprocess(a, opt)[0, 0]
            assert got[i].shape == ref.shape, (pad, i, a.shape, got[i].shape, ref.shape)
            assert np.abs(got[i] - ref).max() <= 1e-6, (pad, i)
    blank = np.full((40, 100), 255, dtype=np.uint8)
    with pytest.raises(PreprocessError, match="blank"):
        Preprocessor(e, dict(OPT, pad=True, downsample=None))([blank])
    with pytest.raises(PreprocessError, match="multiple of 32"):
        Preprocessor(e, dict(OPT, pad=False, downsample=None))([np.full((40, 100), 200, dtype=np.uint8)])
    e.close()


def test_preprocessed_buckets_feed_the_recognizer(built_lib):
    """Mixed crops -> buckets of equal (H, W) -> encode + greedy decode per bucket == the same images pushed one by one."""
    from doc2tex_b200.engine import Engine
    from doc2tex_b200.preprocess import Preprocessor
    from oracle import make_golden_helpers as mh
    cfg, sd = state_dict_for("TFM", 1.5)
    e = Engine(cfg, "cuda:0", precision="bf16x3")
    e.load_state_dict(sd)
    imgs = [mh.synth_crop(h, w, 4000 + k) for k, (h, w) in enumerate([(60, 200), (61, 199), (37, 150), (62, 201), (128, 400)])]
    prep = Preprocessor(e, dict(OPT, pad=True, downsample=None, max_dimension=[192, 896]))
    buckets = prep(imgs)
    assert sorted(len(idx) for _, idx in buckets.values()) == [1, 1, 3]
    for (h, w), (batch, idx) in buckets.items():
        ctx, _, _ = e.encode(batch)
        ids, _, steps = e.decode_greedy(ctx, is_test=True, return_logits=False)
        for slot, i in enumerate(idx):
            one = prep([imgs[i]])[(h, w)][0]
            assert torch.equal(one[0], batch[slot])
            c1, _, _ = e.encode(one)
            i1, _, s1 = e.decode_greedy(c1, is_test=True, return_logits=False)
            assert torch.equal(i1[0, :s1], ids[slot, :s1])
    e.close()

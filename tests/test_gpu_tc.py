"""tcgen05 contraction kernel: unit parity against a float64 torch reference, and whole-encoder parity per precision."""
import pytest
import torch

from doc2tex_b200 import synth
from tests.util import rel_err, state_dict_for

pytestmark = pytest.mark.gpu

# relative error (max-abs / max-abs) a contraction may show per precision mode; K up to 4608.
# Measured on B200: bf16 2e-3; bf16x3 4e-6 (K<=576) .. 1.5e-5 (K=4608); tf32x3 1e-6 (K=64) .. 3e-5 (K=4608).
# The growth with K is the tensor core's fp32 accumulator (truncating adds, one per UMMA_K step), which is
# why tf32x3 (twice the accumulation steps of bf16x3) is not more accurate than bf16x3 at large K.
GEMM_TOL = {"bf16": 2e-2, "bf16x3": 1e-4, "tf32x3": 1e-4}
# whole encoder (31 stacked contractions + ViT) vs the fp32 oracle
CTX_TOL = {"bf16": 6e-2, "bf16x3": 1e-3, "tf32x3": 1e-4}


def _engine(precision):
    from doc2tex_b200.engine import Engine
    cfg, sd = state_dict_for("TFM", None)
    e = Engine(cfg, "cuda:0", precision=precision)
    return e, cfg, sd


@pytest.mark.parametrize("precision", ["bf16", "bf16x3", "tf32x3"])
def test_tc_gemm_matches_float64(built_lib, precision):
    e, _, _ = _engine("fp32")
    g = torch.Generator().manual_seed(1)
    shapes = [(128, 64, 64), (300, 256, 256), (1000, 512, 4608), (257, 128, 576), (4096, 64, 288), (130, 504, 256),
              (20000, 256, 2048)]
    for (M, N, K) in shapes:
        a = torch.randn(M, K, generator=g)
        w = torch.randn(N, K, generator=g) / K ** 0.5
        sc = torch.rand(N, generator=g) + 0.5
        sh = torch.randn(N, generator=g)
        ref = torch.relu((a.double() @ w.double().t()) * sc.double() + sh.double()).float()
        out = e.gemm(a.cuda(), w.cuda(), sc.cuda(), sh.cuda(), act=1, precision=precision).cpu()
        err = rel_err(out, ref)
        print(f"{precision} M={M} N={N} K={K}: rel err {err:.3e}")
        assert err < GEMM_TOL[precision], (precision, M, N, K, err)


@pytest.mark.parametrize("precision", ["bf16", "bf16x3", "tf32x3"])
def test_tc_encoder_matches_oracle(built_lib, precision):
    from oracle import oracle_model as om
    e, cfg, sd = _engine(precision)
    e.load_state_dict(sd)
    img = synth.make_images(2, 64, 256, seed=2024)
    e.set_debug(True)
    ctx, _, _ = e.encode(img.cuda())
    taps = {}
    ctx_ref, _, _ = om.encoder_forward(sd, img, taps=taps)
    for name, ref in taps.items():
        print(f"{precision} stage {name}: rel err {rel_err(e.tap(name).cpu(), ref):.3e}")
    err = rel_err(ctx.cpu(), ctx_ref)
    print(f"{precision} ctx rel err {err:.3e}")
    assert err < CTX_TOL[precision]


@pytest.mark.parametrize("precision", ["bf16x3", "tf32x3"])
def test_tc_fp32_parity_modes_decode_token_exact(built_lib, precision):
    """The error-compensated tensor-core modes must reproduce the reference's tokens (fp32 parity gate)."""
    from doc2tex_b200.engine import Engine
    from tests.util import REL_TOL_FP32, end_bias_of, load_golden
    for case in ("tfm_64x256_full", "tfm_64x256_end15"):
        g = load_golden(case)
        cfg, sd = state_dict_for("TFM", end_bias_of(g))
        e = Engine(cfg, "cuda:0", precision=precision)
        e.load_state_dict(sd)
        img = synth.make_images(2, 64, 256, seed=2024)
        ctx, _, _ = e.encode(img.cuda())
        assert rel_err(ctx.cpu(), torch.from_numpy(g["ctx"])) < REL_TOL_FP32
        ids, logits, steps = e.decode_greedy(ctx, is_test=True)
        ref_ids = torch.from_numpy(g["greedy_gen"])
        assert steps == ref_ids.shape[1]
        assert torch.equal(ids[:, :steps].cpu(), ref_ids)
        ref_logits = torch.from_numpy(g["greedy_logits"])
        for j, s in enumerate(g["greedy_logit_steps"].tolist()):
            assert rel_err(logits[:, s].cpu(), ref_logits[:, j]) < REL_TOL_FP32, s
        bids, blen, bscore, _, _, _ = e.decode_beam(ctx, 5)
        for i in range(2):
            n = int(g["beam_len"][i])
            assert int(blen[i]) == n and bids[i, :n].cpu().tolist() == g["beam_seq"][i, :n].tolist()
        e.close()


def test_attnv2_head_on_tensor_cores_matches_golden(built_lib):
    """config/train.yaml default stack with the contractions on the tensor-core path (bf16x3)."""
    from doc2tex_b200.engine import Engine
    from tests.util import REL_TOL_FP32, end_bias_of, load_golden
    for case in ("attnv2_64x256_full", "attnv2_64x256_end"):
        g = load_golden(case)
        cfg, sd = state_dict_for("Attnv2", end_bias_of(g))
        e = Engine(cfg, "cuda:0", precision="bf16x3")
        e.load_state_dict(sd)
        img = synth.make_images(2, 64, 256, seed=2024)
        ctx, _, _ = e.encode(img.cuda())
        ids, logits, steps = e.decode_greedy(ctx, max_steps=151, is_test=True)
        assert torch.equal(ids.cpu(), torch.from_numpy(g["ids"]))
        ref = torch.from_numpy(g["logits"])
        for j, s in enumerate(g["logit_steps"].tolist()):
            if ref[:, j].abs().max() > 0:
                assert rel_err(logits[:, s].cpu(), ref[:, j]) < REL_TOL_FP32, s
        e.close()


def test_bf16_mode_decode_tolerance(built_lib):
    """Single-pass bf16 mode (bf16 operands, bf16 KV cache) against the fp32 FFMA anchor on the same weights: the stated
    tolerance of the mode: ctx relative L2 error <= 2e-2, greedy token agreement >= 99 % over the first 20
    steps, per-step logits within 5e-2 of the anchor's max-abs while the prefixes agree; beam-5 runs and returns the full
    length."""
    from doc2tex_b200.engine import Engine
    cfg, sd = state_dict_for("TFM", -1e4)
    img = synth.make_images(16, 64, 256, seed=77).cuda()
    out = {}
    for prec in ("fp32", "bf16"):
        e = Engine(cfg, "cuda:0", precision=prec)
        e.load_state_dict(sd)
        ctx, _, _ = e.encode(img)
        ids, logits, steps = e.decode_greedy(ctx, max_steps=40, is_test=True)
        b = e.decode_beam(ctx, 5, max_steps=40)
        out[prec] = (ctx.cpu(), ids.cpu(), logits.cpu(), b[1].cpu())
        e.close()
    c0, i0, l0, _ = out["fp32"]
    c1, i1, l1, bl = out["bf16"]
    assert float((c1 - c0).norm() / c0.norm()) <= 2e-2
    assert float((i0[:, :20] == i1[:, :20]).float().mean()) >= 0.99
    same = (i0 == i1).cumprod(dim=1).bool()
    first = torch.ones_like(same)
    first[:, 1:] = same[:, :-1]
    err = ((l0 - l1).abs().amax(dim=2) / l0.abs().amax(dim=2).clamp_min(1e-6))[first]
    assert float(err.max()) < 5e-2, float(err.max())
    assert bool((bl == 40).all())


@pytest.mark.parametrize("precision", ["bf16x3", "bf16"])
@pytest.mark.parametrize("H,W", [(64, 256), (96, 384)])
def test_fused_maxpool_is_bit_identical(built_lib, precision, H, W):
    """Option fuse_pool: max-pools 1 and 2 (resnet.py:214-217, 222-225) run in the epilogue of conv0_2 / conv1 (window-major
    pixel order, BN + ReLU on the four pixels, maximum, operand planes of the POOLED map only).  Same values, same order of
    operations per element as conv -> fp32 map -> pool kernel: the encoder output must be bit-identical."""
    from doc2tex_b200.engine import Engine
    cfg, sd = state_dict_for("TFM", None)
    e = Engine(cfg, "cuda:0", precision=precision)
    e.load_state_dict(sd)
    img = synth.make_images(3, H, W, seed=2024).cuda()
    out = {}
    for mode in (0, 1):
        e.set_option("fuse_pool", mode)
        l0 = e.launch_count()
        ctx, _, _ = e.encode(img)
        torch.cuda.synchronize()
        out[mode] = (ctx.cpu(), e.launch_count() - l0)
    e.close()
    assert torch.equal(out[0][0], out[1][0])
    assert out[0][1] - out[1][1] == 2      # two pool launches fewer


@pytest.mark.parametrize("precision", ["bf16x3", "bf16"])
@pytest.mark.parametrize("B,H,W", [(3, 64, 256), (5, 96, 384), (1, 32, 32), (2, 160, 928)])
def test_pair_and_tma_kernels_are_bit_identical(built_lib, precision, B, H, W):
    """The three ways the stem convolutions receive their operands — producer-warp cp.async gather (tma_a 0), TMA im2col
    loads in the single-CTA kernel (tma_a 1, pair 0) and the CTA-pair kernel with 128-byte rows (pair 2) — accumulate the
    same products in the same order: the encoder output must be bit-identical, including the ragged last tiles of small and
    odd-sized batches (the pair's second CTA may own no pixel at all)."""
    from doc2tex_b200.engine import Engine
    cfg, sd = state_dict_for("TFM", None)
    e = Engine(cfg, "cuda:0", precision=precision)
    e.load_state_dict(sd)
    img = synth.make_images(B, H, W, seed=99).cuda()
    out = []
    try:
        for tma_a, pair in ((0, 0), (1, 0), (1, 2)):
            e.set_option("tma_a", tma_a)
            e.set_option("pair", pair)
            ctx, _, _ = e.encode(img)
            out.append(ctx.cpu())
    finally:
        e.set_option("tma_a", 1)     # process-wide switch: restore the default
        e.close()
    assert torch.equal(out[0], out[1])
    assert torch.equal(out[0], out[2])


@pytest.mark.parametrize("precision", ["bf16x3", "bf16"])
@pytest.mark.parametrize("B,H,W", [(40, 64, 256), (33, 128, 512), (3, 96, 384)])
def test_vit_linears_on_operand_planes_match_gather_path(built_lib, precision, B, H, W):
    """Option vit_planes: the ViT blocks' Linears read bf16 operand planes by TMA on the stem's kernels (1 = auto, 2 = single-CTA
    kernel, 3 = CTA pair wherever its tile count allows) instead of gathering fp32 rows through registers (0).  Same products,
    another summation order: the encoder output agrees far inside the fp32-parity tolerance (bf16x3) / the mode's own noise
    (bf16), including row counts that are no multiple of the 128- / 256-row tiles (33 x 261 = 8 613 rows)."""
    from doc2tex_b200.engine import Engine
    cfg, sd = state_dict_for("TFM", None)
    e = Engine(cfg, "cuda:0", precision=precision)
    e.load_state_dict(sd)
    img = synth.make_images(B, H, W, seed=5).cuda()
    out = {}
    try:
        for mode in (0, 1, 2, 3):
            e.set_option("vit_planes", mode)
            ctx, _, _ = e.encode(img)
            out[mode] = ctx.cpu()
    finally:
        e.close()
    tol = 2e-5 if precision == "bf16x3" else 2e-2
    for mode in (1, 2, 3):
        assert torch.isfinite(out[mode]).all()
        assert rel_err(out[mode], out[0]) < tol, (mode, rel_err(out[mode], out[0]))


def test_plane_paths_fall_back_when_the_plane_kernels_are_off(built_lib):
    """Option tc3 = 0 switches the cp.async / TMA-fed plane kernels off: the ViT Linears (vit_planes) and the wide decode
    projections (wide_decode) must fall back to the register-gather / one-tile kernels instead of handing them a missing fp32
    tensor; results stay within the fp32-parity tolerance of the default path."""
    from doc2tex_b200.engine import Engine
    cfg, sd = state_dict_for("TFM", 1.5)
    img = synth.make_images(4, 64, 256, seed=3).cuda()
    out = {}
    for tc3 in (1, 0):
        e = Engine(cfg, "cuda:0", precision="bf16x3")
        e.load_state_dict(sd)
        e.set_option("tc3", tc3)
        ctx, _, _ = e.encode(img)
        wide = ctx[:1].repeat(2560, 1, 1).contiguous()      # lin1: 20 x 8 = 160 tiles > 148 SMs
        ids, logits, _ = e.decode_greedy(wide, max_steps=3, is_test=False)
        out[tc3] = (ctx.cpu(), ids.cpu(), logits[:8].cpu())
        e.close()
    assert rel_err(out[0][0], out[1][0]) < 1e-4
    assert torch.equal(out[0][1], out[1][1])
    assert rel_err(out[0][2], out[1][2]) < 1e-4

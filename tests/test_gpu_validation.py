"""f2: the batched evaluation loop (doc2tex/engine/inferencing.py::validation_step) on the engine against the LIVE
reference's outputs (tests/golden/validation_step.json, oracle/make_golden.py::validation_case): prediction / label strings
exact, per-sample losses within the fp32 tolerance, accuracy / edit distances / BLEU equal."""
import json
import os
import types

import pytest
import torch

from doc2tex_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
@pytest.mark.parametrize("case", ["tfm_noend", "tfm_end15", "attnv2_end30"])
def test_validation_step_matches_live_reference(built_lib, case, precision):
    from doc2tex_b200.engine_inferencing import validation_step
    from doc2tex_b200.modules.build_model import Model
    from doc2tex_b200.modules.converter import AttnLabelConverter, TFMLabelConverter
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "validation_step.json")))
    rec = g["cases"][case]
    head = rec["head"]
    cfg = synth.make_config(head)
    cfg["engine"] = {"precision": precision}
    sd = synth.make_state_dict(cfg, seed=1111, end_bias=rec["end_bias"])
    model = Model(cfg)
    model.load_state_dict(sd, strict=True)
    model = model.to("cuda:0").eval()
    conv = (TFMLabelConverter if head == "TFM" else AttnLabelConverter)(g["vocab"], "cuda:0")
    loader = []
    for bidx, labels in enumerate(rec["labels_in"]):
        img = synth.make_images(3, 64, 256, seed=5000 + 3 * bidx).cuda()
        loader.append((img, labels, [f"img_{bidx}_{k}.png" for k in range(3)]))
    config = dict(cfg, export_csv=False, use_amp=False, sanity_check=False, token_level="word", postprocess=True)
    crit = torch.nn.CrossEntropyLoss(ignore_index=conv.ignore_idx, reduction="none")
    with torch.no_grad():
        res = validation_step(model, None, crit, loader, conv, config, types.SimpleNamespace(log_path="golden.log"), "cuda",
                              decode_merge=2)
    all_loss, names, mean_loss, acc, bleu, ned, wed, preds, labels, infer_time, n = res
    assert n == rec["n"] and names == rec["names"]
    assert preds == rec["preds"], (preds[0][:80], rec["preds"][0][:80])       # strings exact
    assert labels == rec["labels"]
    assert torch.allclose(torch.tensor(all_loss), torch.tensor(rec["all_loss"]), rtol=1e-3, atol=1e-5)
    assert abs(float(mean_loss) - rec["mean_loss"]) <= 1e-3 * max(1.0, abs(rec["mean_loss"]))
    assert acc == rec["accuracy"] and abs(ned - rec["norm_ED"]) < 1e-9 and abs(wed - rec["word_ED"]) < 1e-9
    assert (bleu is None) == (rec["bleu"] is None) and (bleu is None or abs(bleu - rec["bleu"]) < 1e-6)
    assert infer_time > 0


def test_model_accepts_the_reference_eval_call(built_lib):
    """inferencing.py:151-153 calls model(image, text) with the defaults is_train=True, is_test=False; the TFM head ignores
    is_train (tfm.py:188-195): greedy, all 151 steps.  The LSTM heads' is_train=True is teacher forcing -> refused."""
    from doc2tex_b200.engine import EngineError
    from doc2tex_b200.modules.build_model import Model
    cfg = synth.make_config("TFM")
    m = Model(cfg)
    m.load_state_dict(synth.make_state_dict(cfg, seed=1111, end_bias=1.5), strict=True)
    m = m.to("cuda:0").eval()
    img = synth.make_images(2, 64, 256, seed=2024).cuda()
    text = torch.full((2, 1), 1, dtype=torch.long, device="cuda:0")
    with torch.no_grad():
        ids, logits, _ = m(img, text)
        ids_t, logits_t, _ = m(img, text, is_train=False, is_test=True)
    assert ids.shape == (2, 151) and logits.shape == (2, 151, cfg["num_class"])
    assert torch.equal(ids[:, : ids_t.shape[1]], ids_t)
    cfg2 = synth.make_config("Attnv2")
    m2 = Model(cfg2).to("cuda:0").eval()
    with pytest.raises(EngineError, match="teacher forcing"):
        m2(img, torch.zeros(2, 151, dtype=torch.long, device="cuda:0"))

import sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
from doc2tex_b200 import synth
from doc2tex_b200.engine import Engine
from tests.util import load_golden, state_dict_for, end_bias_of
for case in ("attnv2_beam_64x256_full", "attnv2_beam_64x256_end04"):
    g = load_golden(case)
    cfg, sd = state_dict_for("Attnv2", end_bias_of(g))
    for prec in ("fp32", "bf16x3"):
        e = Engine(cfg, "cuda:0", precision=prec); e.load_state_dict(sd)
        B = int(g["beam_len"].shape[0])
        ctx, _, _ = e.encode(synth.make_images(B, 64, 256, seed=2024).cuda())
        ids, lens, score, steps, tr, trs = e.decode_beam(ctx, 5, trace=True)
        tr, trs = tr.cpu().numpy(), trs.cpu().numpy()
        for i in range(B):
            T = int(g["beam_steps"][i])
            for t in range(T):
                k = int((g["beam_parents"][i, t] >= 0).sum())
                if not (np.array_equal(tr[i, t, :k, 0], g["beam_parents"][i, t, :k]) and np.array_equal(tr[i, t, :k, 1], g["beam_words"][i, t, :k])):
                    print(case, prec, "img", i, "first diff at step", t, "k", k)
                    print("  ref parents", g["beam_parents"][i, t, :k], "words", g["beam_words"][i, t, :k], "scores", g["beam_scores"][i, t, :k])
                    print("  got parents", tr[i, t, :k, 0], "words", tr[i, t, :k, 1], "scores", trs[i, t, :k])
                    print("  prev step ref scores", g["beam_scores"][i, t-1, :k], "got", trs[i, t-1, :k])
                    break
            else:
                print(case, prec, "img", i, "trace identical; score", float(score[i]), "ref", g["beam_score"][i], "len", int(lens[i]))
        e.close()

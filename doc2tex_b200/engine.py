"""Python host of the B200 engine: owns a ``d2t_engine`` handle and passes raw device
pointers of torch tensors through the C ABI (include/doc2tex_b200.h).  PyTorch is used
only for device memory, streams and (in ``dist.py``) ``torch.distributed``.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Tuple

import torch

from . import _lib


class EngineError(RuntimeError):
    pass


def config_from_opt(opt: dict, precision: str = "fp32", use_graphs: bool = True) -> _lib.Config:
    """Translate the reference's YAML/dict surface into ``d2t_config``."""
    sp = opt["SequenceModeling"]["params"]
    if opt["FeatureExtraction"]["name"] != "None" or opt["SequenceModeling"]["name"] != "ViT":
        raise EngineError("only the HybridViT stack (FeatureExtraction None + SequenceModeling ViT) is accelerated")
    if sp.get("patching_style", "2d") != "2d":
        raise EngineError("only patching_style '2d' (ViTEncoder / ViTEncoderV2 / ViTEncoderV3) is supported")
    if sp["backbone"]["name"] != "resnet" or sp["backbone"].get("gcb", False):
        raise EngineError("only the resnet backbone without GlobalContext is supported")
    if list(sp["patch_size"]) != [2, 2]:
        raise EngineError("only patch_size [2, 2] is supported")
    max_h, max_w = (opt["imgH"], opt["max_dimension"][1]) if opt.get("imgH") else opt["max_dimension"]
    fh, fw = max_h // 16 - 1, max_w // 4 + 1
    cfg = _lib.Config()
    cfg.struct_size = C.sizeof(_lib.Config)
    cfg.in_channels = sp["backbone"]["input_channel"]
    cfg.stem_channels = sp["backbone"]["output_channel"]
    cfg.hidden = sp["hidden_size"]
    cfg.depth = sp["depth"]
    cfg.heads = sp["num_heads"]
    cfg.max_tokens = 1 + ((fh + 1) // 2) * ((fw + 1) // 2)
    pred = opt["Prediction"]
    pp = pred.get("params", {})
    name = pred["name"]
    cfg.vocab = int(opt.get("num_class", 0))
    if name == "TFM":
        cfg.head = _lib.HEAD["TFM"]
        if pp["d_model"] != sp["hidden_size"]:
            raise EngineError("TFM d_model must equal the encoder hidden_size")
        cfg.dec_layers, cfg.dec_heads, cfg.dec_ff = pp["num_decoder_layers"], pp["nhead"], pp["dim_feedforward"]
        cfg.max_seq_len = pp["max_seq_len"]
    elif name in ("Attnv2", "Attn"):
        if pp.get("attn_type", "coverage") != "coverage" or not pp.get("embed_target", False) \
                or not pp.get("enc_init", False):
            raise EngineError(f"{name} is supported as shipped: coverage attention, embed_target, enc_init")
        if name == "Attnv2" and pp.get("seqmodel", "ViT") != "TFM":
            raise EngineError("Attnv2 is supported with seqmodel 'TFM' (it drops the cls token, seq2seq_v2.py:27,190)")
        if name == "Attn" and pp.get("seqmodel", "ViT") == "BiLSTM":
            raise EngineError("Attn with seqmodel 'BiLSTM' (mean-pooled init, seq2seq.py:232-234) is not on the accelerated path")
        # 'Attn' (seq2seq.py) attends over every encoder token incl. cls; 'Attnv2' (seq2seq_v2.py) skips the cls token
        cfg.head = _lib.HEAD[name]
        cfg.attn_hidden, cfg.attn_kernel_dim, cfg.attn_kernel_size = pp["hidden_size"], pp["kernel_dim"], pp["kernel_size"]
        cfg.max_seq_len = int(opt["batch_max_length"])
    else:
        raise EngineError(f"Prediction head {name!r} is not on the accelerated path (TFM, Attnv2, Attn)")
    cfg.precision = _lib.PREC[precision]
    cfg.use_graphs = 1 if use_graphs else 0
    return cfg


class Engine:
    def __init__(self, opt: dict, device="cuda:0", precision: str = "fp32", use_graphs: bool = True):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise EngineError("doc2tex_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise EngineError("doc2tex_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.cfg = config_from_opt(opt, precision, use_graphs)
        self.precision = precision
        h = C.c_void_p()
        rc = self.lib.d2t_create(C.byref(self.cfg), self.index, C.byref(h))
        if rc != 0:
            raise EngineError(f"d2t_create failed ({rc}): {self.lib.d2t_last_error(None).decode()}")
        self.h = h
        self.loaded = False
        sp = opt["SequenceModeling"]["params"]
        # encoder variant (vit_encoder.py:301-309): fix_embed -> V3 (sin-cos table, prefix slice); else interpolate_embed
        # (default True) -> ViTEncoder (bicubic resample of the learnable table per image size), False -> V2 (prefix slice)
        if not sp.get("fix_embed", False) and sp.get("interpolate_embed", True):
            max_h, max_w = (opt["imgH"], opt["max_dimension"][1]) if opt.get("imgH") else opt["max_dimension"]
            fh, fw = max_h // 16 - 1, max_w // 4 + 1
            self.set_option("pos_interpolate", 1)
            self.set_option("pos_grid_h", (fh + 1) // 2)
            self.set_option("pos_grid_w", (fw + 1) // 2)

    # ------------------------------------------------------------------
    def _check(self, rc: int, what: str):
        if rc != 0:
            raise EngineError(f"{what} failed ({rc}): {self.lib.d2t_last_error(self.h).decode()}")

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def close(self):
        if getattr(self, "h", None):
            self.lib.d2t_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------
    def load_state_dict(self, sd: Dict[str, torch.Tensor]):
        """Hand every tensor to the engine under its reference key, then fold/repack/upload."""
        for k, v in sd.items():
            t = v.detach().to("cpu").contiguous()
            if t.dtype == torch.int64:
                dt = 1
            else:
                t = t.float().contiguous()
                dt = 0
            shape = (C.c_int64 * max(t.dim(), 1))(*t.shape)
            self._check(self.lib.d2t_load_tensor(self.h, k.encode(), C.c_void_p(t.data_ptr()), shape, t.dim(), dt),
                        f"d2t_load_tensor({k})")
        self._check(self.lib.d2t_finalize_weights(self.h), "d2t_finalize_weights")
        self.loaded = True

    def geometry(self, H: int, W: int):
        v = [C.c_int() for _ in range(5)]
        self._check(self.lib.d2t_encoder_geometry(self.h, H, W, *[C.byref(x) for x in v]), "d2t_encoder_geometry")
        gh, gw, pad_w, pad_h, ntok = [x.value for x in v]
        return (gh, gw), (pad_w, pad_h), ntok

    def _dev(self, t: torch.Tensor, dtype) -> torch.Tensor:
        if t.device != self.device and not (t.is_cuda and t.device.index == self.index):
            raise EngineError(f"tensor on {t.device}, engine on {self.device}")
        return t.to(dtype).contiguous()

    def encode(self, img: torch.Tensor) -> Tuple[torch.Tensor, Tuple[int, int], Tuple[int, int]]:
        """Model.forward_encoder: (B,1,H,W) fp32 -> ctx (B, N+1, hidden), (gh, gw), (pad_W, pad_H)."""
        img = self._dev(img, torch.float32)
        B, Cin, H, W = img.shape
        if Cin != self.cfg.in_channels:
            raise EngineError(f"expected {self.cfg.in_channels} input channel(s), got {Cin}")
        grid, pad, ntok = self.geometry(H, W)
        ctx = torch.empty(B, ntok, self.cfg.hidden, device=self.device, dtype=torch.float32)
        self._check(self.lib.d2t_encode(self.h, img.data_ptr(), B, H, W, ctx.data_ptr(), self._stream()), "d2t_encode")
        return ctx, grid, pad

    def decode_greedy(self, ctx: torch.Tensor, max_steps: Optional[int] = None, is_test: bool = True,
                      return_logits: bool = True):
        """TransformerPrediction.forward_greedy (eval): ids (B, l) int64, logits (B, l, V) or None."""
        ctx = self._dev(ctx, torch.float32)
        B, ntok, _ = ctx.shape
        T = self.cfg.max_seq_len + 1 if max_steps is None else max_steps
        ids = torch.empty(B, T, device=self.device, dtype=torch.int64)
        logits = torch.empty(B, T, self.cfg.vocab, device=self.device, dtype=torch.float32) if return_logits else None
        steps = C.c_int()
        fn = self.lib.d2t_decode_greedy if self.cfg.head == _lib.HEAD["TFM"] else self.lib.d2t_decode_attn_greedy
        self._check(fn(self.h, ctx.data_ptr(), B, ntok, T, 1 if is_test else 0, ids.data_ptr(),
                       logits.data_ptr() if return_logits else None, C.byref(steps), self._stream()), "d2t_decode_greedy")
        return ids, logits, steps.value

    def decode_beam(self, ctx: torch.Tensor, beam: int, max_steps: Optional[int] = None, trace: bool = False,
                    runner_up: Optional[torch.Tensor] = None):
        """Batched TransformerPrediction.forward_beam / AttentionV2.forward_beam: per image best ids (padded), length, score.
        runner_up (TFM head, tests): a (B, T) fp32 device tensor that receives, per step, the score of the best candidate NOT
        selected — with the trace scores it gives the margin of every beam decision (near-tie audit)."""
        ctx = self._dev(ctx, torch.float32)
        B, ntok, _ = ctx.shape
        T = self.cfg.max_seq_len + 1 if max_steps is None else max_steps
        if runner_up is not None:
            assert runner_up.is_cuda and runner_up.dtype == torch.float32 and tuple(runner_up.shape) == (B, T) and runner_up.is_contiguous()
            self._check(self.lib.d2t_debug_beam_runner_up(self.h, runner_up.data_ptr()), "d2t_debug_beam_runner_up")
        ids = torch.empty(B, T, device=self.device, dtype=torch.int64)
        lens = torch.empty(B, device=self.device, dtype=torch.int32)
        score = torch.empty(B, device=self.device, dtype=torch.float32)
        tr = torch.empty(B, T, beam, 2, device=self.device, dtype=torch.int32) if trace else None
        trs = torch.empty(B, T, beam, device=self.device, dtype=torch.float32) if trace else None
        steps = C.c_int()
        fn = self.lib.d2t_decode_beam if self.cfg.head == _lib.HEAD["TFM"] else self.lib.d2t_decode_attn_beam
        self._check(fn(self.h, ctx.data_ptr(), B, ntok, beam, T, ids.data_ptr(), lens.data_ptr(),
                       score.data_ptr(), tr.data_ptr() if trace else None,
                       trs.data_ptr() if trace else None, C.byref(steps), self._stream()),
                    "d2t_decode_beam")
        if runner_up is not None:
            self._check(self.lib.d2t_debug_beam_runner_up(self.h, None), "d2t_debug_beam_runner_up")
        return ids, lens, score, steps.value, tr, trs

    @property
    def end_id(self) -> int:
        """[s]/END token id of the head's converter (tfm_converter.py:8, attn_converter.py:8)."""
        return 2 if self.cfg.head == _lib.HEAD["TFM"] else 1

    def set_option(self, key: str, value: int):
        """Engine knobs (documented at d2t_set_option in include/doc2tex_b200.h)."""
        self._check(self.lib.d2t_set_option(self.h, key.encode(), int(value)), f"d2t_set_option({key})")

    # ---- test / profiling hooks ----
    def set_debug(self, keep_taps: bool):
        self._check(self.lib.d2t_set_debug(self.h, 1 if keep_taps else 0), "d2t_set_debug")

    def tap(self, name: str) -> torch.Tensor:
        numel = C.c_int64()
        shape = (C.c_int64 * 4)()
        self._check(self.lib.d2t_debug_tap(self.h, name.encode(), None, C.byref(numel), shape, self._stream()), "d2t_debug_tap")
        dims = [int(s) for s in shape if s > 0]
        out = torch.empty(dims, device=self.device, dtype=torch.float32)
        assert out.numel() == numel.value
        self._check(self.lib.d2t_debug_tap(self.h, name.encode(), out.data_ptr(), C.byref(numel), shape, self._stream()), "d2t_debug_tap")
        return out

    def gemm(self, a: torch.Tensor, w: torch.Tensor, scale=None, shift=None, act: int = 0, precision: Optional[str] = None):
        a, w = self._dev(a, torch.float32), self._dev(w, torch.float32)
        M, K = a.shape
        N = w.shape[0]
        c = torch.empty(M, N, device=self.device, dtype=torch.float32)
        sc = self._dev(scale, torch.float32) if scale is not None else None
        sh = self._dev(shift, torch.float32) if shift is not None else None
        self._check(self.lib.d2t_debug_gemm(self.h, a.data_ptr(), w.data_ptr(), sc.data_ptr() if sc is not None else None,
                                            sh.data_ptr() if sh is not None else None, c.data_ptr(), M, N, K, act,
                                            _lib.PREC[precision or self.precision], self._stream()), "d2t_debug_gemm")
        return c

    def gemm_bench(self, M: int, N: int, K: int, precision: str, iters: int = 200, interleave: bool = False) -> float:
        """Average microseconds per launch of an (M,N,K) contraction, warm, back to back."""
        a = torch.randn(M, K, device=self.device)
        w = torch.randn(N, K, device=self.device)
        c = torch.empty(M, N, device=self.device)
        ms = C.c_float()
        self._check(self.lib.d2t_debug_gemm_bench(self.h, a.data_ptr(), w.data_ptr(), c.data_ptr(), M, N, K,
                                                  _lib.PREC[precision], iters, 1 if interleave else 0, C.byref(ms),
                                                  self._stream()), "d2t_debug_gemm_bench")
        return ms.value * 1e3

    def conv_time(self):
        """(total ms, launches, FLOPs per launch) of the layer3 3x3 convolution launches timed since the last call
        (enable with set_option("time_conv", 1))."""
        ms, n, fl = C.c_double(), C.c_int64(), C.c_double()
        self._check(self.lib.d2t_debug_conv_time(self.h, C.byref(ms), C.byref(n), C.byref(fl)), "d2t_debug_conv_time")
        return ms.value, n.value, fl.value

    def decode_time(self, kind: int):
        """(total ms, launches, algorithmic bytes) of the decode launches of one kind timed since the last call (enable with
        set_option("time_decode", 1)): 0 self-attention, 1 cross-attention, 2 beam step, 3 greedy pick."""
        ms, n, by = C.c_double(), C.c_int64(), C.c_double()
        self._check(self.lib.d2t_debug_decode_time(self.h, kind, C.byref(ms), C.byref(n), C.byref(by)), "d2t_debug_decode_time")
        return ms.value, n.value, by.value

    def launch_count(self) -> int:
        return int(self.lib.d2t_launch_count(self.h))

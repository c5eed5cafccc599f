/*
 * doc2tex_b200 — C ABI of the B200-native recognizer engine.
 *
 * This is the drop-in boundary for the hot path of duylebkHCM/doc2tex: the batched
 * recognizer forward and its autoregressive decode.  Every entry point replaces a
 * Python/PyTorch call site of the reference (file:line relative to the reference
 * repository).  Only plain pointers, sizes and a CUDA stream handle cross the
 * boundary; the caller (PyTorch on the host side) owns all input/output buffers,
 * the engine owns weights, KV caches, workspaces and CUDA graphs.
 *
 * All functions return 0 on success and a non-zero d2t_status otherwise;
 * d2t_last_error() gives the message.  No C++ exception crosses the boundary.
 * One engine per (process, device); calls on one engine are not re-entrant.
 * All pointers named *_dev are device pointers on the engine's device; work is
 * enqueued on the caller's stream and is asynchronous w.r.t. the host unless noted.
 */
#ifndef DOC2TEX_B200_H_
#define DOC2TEX_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct d2t_engine d2t_engine;
typedef void* d2t_stream; /* cudaStream_t */

enum d2t_status {
  D2T_OK = 0,
  D2T_ERR_INVALID = 1,   /* bad argument / shape / config */
  D2T_ERR_MISSING = 2,   /* a state_dict tensor the config requires was never loaded */
  D2T_ERR_CUDA = 3,      /* CUDA runtime / driver error */
  D2T_ERR_STATE = 4,     /* call order (e.g. encode before finalize) */
  D2T_ERR_UNSUPPORTED = 5
};

/* D2T_HEAD_ATTN = Prediction.name 'Attn' (seq2seq.py:10-345): the same LSTM coverage-attention decoder as 'Attnv2', but it
 * attends over ALL encoder tokens including the cls token (seq2seq.py:236-238), where AttentionV2 with seqmodel 'TFM'
 * drops it (seq2seq_v2.py:27, 190). */
enum d2t_head { D2T_HEAD_NONE = 0, D2T_HEAD_TFM = 1, D2T_HEAD_ATTNV2 = 2, D2T_HEAD_ATTN = 3 };

/* Arithmetic of the dense contractions (conv stem, patch-embed, linear layers). */
enum d2t_precision {
  D2T_PREC_FP32 = 0,    /* fp32 FFMA (CUDA cores): the parity anchor */
  D2T_PREC_TF32X3 = 1,  /* tcgen05 kind::tf32, error-compensated 3-pass split (~2^-21) */
  D2T_PREC_BF16X3 = 2,  /* tcgen05 kind::f16 on bf16 hi/lo planes, 3 passes (~2^-16) */
  D2T_PREC_BF16 = 3     /* tcgen05 kind::f16, single bf16 pass, fp32 accumulate */
};

enum d2t_dtype { D2T_F32 = 0, D2T_I64 = 1 };

/*
 * Mirrors the YAML keys Model(opt) reads on this path
 * (build_model.py:7-34, vit_encoder.py:271-317, tfm.py:36-47, seq2seq.py:11-28).
 */
typedef struct d2t_config {
  int32_t struct_size;      /* sizeof(d2t_config), ABI check */
  int32_t in_channels;      /* SequenceModeling.params.backbone.input_channel (1) */
  int32_t stem_channels;    /* ...backbone.output_channel (512) */
  int32_t hidden;           /* ...hidden_size (256) */
  int32_t depth;            /* ...depth (6) */
  int32_t heads;            /* ...num_heads (8) */
  int32_t max_tokens;       /* rows of pos_embed = 1 + max grid (vit_encoder.py:234-237) */
  int32_t head;             /* d2t_head; Prediction.name */
  int32_t vocab;            /* num_class (build_pred.py:16-26) */
  int32_t dec_layers;       /* TFM: num_decoder_layers */
  int32_t dec_heads;        /* TFM: nhead */
  int32_t dec_ff;           /* TFM: dim_feedforward */
  int32_t max_seq_len;      /* TFM: max_seq_len / Attn: batch_max_length */
  int32_t attn_hidden;      /* Attnv2: hidden_size */
  int32_t attn_kernel_dim;  /* Attnv2: kernel_dim */
  int32_t attn_kernel_size; /* Attnv2: kernel_size (Conv1d taps = 2k+1) */
  int32_t precision;        /* d2t_precision */
  int32_t use_graphs;       /* capture the decode step in a CUDA graph */
} d2t_config;

/* Replaces Model.__init__ (build_model.py:7-34). */
int d2t_create(const d2t_config* cfg, int device, d2t_engine** out);
int d2t_destroy(d2t_engine* e);
const char* d2t_last_error(const d2t_engine* e); /* e may be NULL: last create error */
const char* d2t_version(void);

/*
 * Replaces Model.load_state_dict / utils.model_utils.load_checkpoint
 * (model_utils.py:136-237): hand every state_dict entry to the engine by its
 * reference key, then finalize once.  `data` is a HOST
 * pointer, contiguous, row-major; the engine copies it.
 */
int d2t_load_tensor(d2t_engine* e, const char* key, const void* data,
                    const int64_t* shape, int ndim, int dtype);
/* Folds eval-mode BatchNorm into per-channel scale/shift, repacks conv weights to
 * [Cout][KH][KW][Cin] (K-major, NHWC activations) and uploads. */
int d2t_finalize_weights(d2t_engine* e);

/*
 * Replaces Model.forward_encoder (build_model.py:36-43 -> build_seq.py:59-66 ->
 * vit_encoder.py:249-268 -> patchembed.py:115-141 -> resnet.py:205-245).
 * img_dev: (B,1,H,W) fp32 NCHW, H and W multiples of 32.
 * ctx_dev: (B, ntok, hidden) fp32 out, ntok = 1 + gh*gw (see d2t_encoder_geometry).
 */
int d2t_encode(d2t_engine* e, const float* img_dev, int B, int H, int W,
               float* ctx_dev, d2t_stream stream);
/* (gh, gw) = output_shape, (pad_w, pad_h) = feat_pad of forward_encoder. */
int d2t_encoder_geometry(const d2t_engine* e, int H, int W, int* gh, int* gw,
                         int* pad_w, int* pad_h, int* ntok);

/*
 * Replaces TransformerPrediction.forward_greedy, eval branch (tfm.py:119-143), with a
 * KV cache instead of the reference's full-prefix recompute.
 * ctx_dev (B, ntok, hidden).  ids_dev (B, max_steps) int64: token chosen at each step.
 * logits_dev (B, max_steps, vocab) fp32 or NULL.  Runs until every row has emitted END
 * (when stop_on_all_eos, i.e. is_test=True) or max_steps; *steps_out (host) receives the
 * number of steps the reference would have executed.  Synchronises the stream.
 */
int d2t_decode_greedy(d2t_engine* e, const float* ctx_dev, int B, int ntok,
                      int max_steps, int stop_on_all_eos, int64_t* ids_dev,
                      float* logits_dev, int* steps_out, d2t_stream stream);

/*
 * Replaces TransformerPrediction.forward_beam + tools/beam.py (tfm.py:145-186,
 * beam.py:38-140), batched over B images (the reference is batch-1 only; each image gets
 * a fresh beam).  best_ids_dev (B, max_steps) int64 padded with PAD(0);
 * best_len_dev (B) int32; best_score_dev (B) fp32.
 * trace_dev: optional (B, max_steps, beam, 2) int32 = (parent, word) of the top-k
 * candidates per step in top-k order, -1 where unused; trace_score_dev optional
 * (B, max_steps, beam) fp32.  Synchronises the stream.
 */
int d2t_decode_beam(d2t_engine* e, const float* ctx_dev, int B, int ntok, int beam,
                    int max_steps, int64_t* best_ids_dev, int32_t* best_len_dev,
                    float* best_score_dev, int32_t* trace_dev, float* trace_score_dev,
                    int* steps_out, d2t_stream stream);

/*
 * Replaces AttentionV2.forward_greedy, eval branch (seq2seq_v2.py:176-293).
 * ids_dev (B, max_steps) int64 (0 = [GO] after early exit, like the reference's
 * probs.max(2) over untouched zero rows); logits_dev (B, max_steps, vocab) or NULL.
 */
int d2t_decode_attn_greedy(d2t_engine* e, const float* ctx_dev, int B, int ntok,
                           int max_steps, int stop_on_all_eos, int64_t* ids_dev,
                           float* logits_dev, int* steps_out, d2t_stream stream);

/*
 * Replaces AttentionV2.forward_beam (seq2seq_v2.py:12-174), batched over B images (the reference asserts batch 1, :18-19).
 * Same outputs as d2t_decode_beam.  Reproduces the reference's rules: step 0 ranks row 0 only; hidden state follows the
 * parent while the coverage memory is re-indexed by top-k position; when the last executed step completed nothing the
 * live beam 0 wins, else the first maximum of fp32 score / len(seq incl. GO and END), returned with the maximum
 * completed score.
 */
int d2t_decode_attn_beam(d2t_engine* e, const float* ctx_dev, int B, int ntok, int beam, int max_steps,
                         int64_t* best_ids_dev, int32_t* best_len_dev, float* best_score_dev, int32_t* trace_dev,
                         float* trace_score_dev, int* steps_out, d2t_stream stream);

/*
 * Engine knobs (no reference counterpart; every switch of the engine is set here, none through the environment).
 * Changing a decode-affecting key drops the cached CUDA graphs of the decode step.
 *   "encoder_sms"      SMs the encoder's persistent tensor-core kernels may occupy (default: all) — leaving a few SMs free
 *                      lets the latency-bound decode of batch i overlap the encode of batch i+1 (doc2tex_b200/pipeline.py)
 *   "pdl"              0/1 programmatic dependent launch inside the decode step (default 1)
 *   "pdl_max_rows"     decode calls with more rows than this launch without it (default 2048: it only pays for small launches)
 *   "steps_per_graph"  decode steps captured per CUDA graph, 1..16 (default 8 = the early-exit poll interval)
 *   "decode_groups"    concurrent row groups of one decode call (parallel graph branches; 0/1 = one chain, default)
 *   "split_k"          0 = never, 1 = auto split-K of the LayerNorm-fed decode projections (default 1)
 *   "stack_mma"        0/1 bf16x3 decode projections: two MMAs per k-step against the stacked [W_hi ; W_lo] operand (default 1)
 *   "attn_kpi"         keys in flight per quarter warp and iteration of the decode attention walk: 2, 4 (default), 8
 *   "attn_split"       warps per (row, head) of the decode attention: 0 = auto, 1, 2
 *   "fuse_pick"        0/1 greedy pick also embeds the next token and advances the step counter (default 1)
 *   "kv_bf16"          0/1 bf16 KV caches in the single-pass bf16 mode (default 1; the fp32-parity modes always keep fp32)
 *   "tma_a"            0/1 stem convolutions (stride 1, symmetric padding): the activation operand by TMA im2col loads of the bf16 planes
 *                      instead of the producer warps' cp.async gather (process-wide; default 1)
 *   "pair"             stem convolutions with 256-multiple output channels on a CTA pair (tcgen05 cta_group::2, both operands by
 *                      TMA): 0 = never, 1 = single-pass bf16 mode only, 2 = the 3-pass parity mode too (default; needs tma_a = 1)
 *   "tc3"              0/1 stem convolutions fed from bf16 activation planes by cp.async (default 1)
 *   "lean_acts"        0/1 stem layers write only the representations their consumers read (default 1)
 *   "fuse_pool"        0/1 max-pools 1 and 2 fused into the producing convolution's epilogue (default 1)
 *   "wide_decode"      0/1 decode projections with more 128x128 tiles than SMs (merged calls of thousands of rows) run on the
 *                      stem's persistent TMA-fed kernels instead of the fp32 register-gather kernel (default 1)
 *   "vit_planes"       the ViT blocks' Linears read bf16 operand planes by TMA on the stem's kernels: 0 off (fp32 gather), 1 auto
 *                      (default: CTA pair in bf16x3, single-CTA 128-wide tile in bf16), 2 single-CTA only, 3 CTA pair where it applies
 *   "pos_interpolate", "pos_grid_h", "pos_grid_w"   ViTEncoder (fix_embed: False) bicubic pos-embed resampling
 *   "time_conv", "time_decode"   see d2t_debug_conv_time / d2t_debug_decode_time
 *   "dbg_decode", "dbg_timeline" phase / per-launch timestamps of the last decode step on stderr
 */
int d2t_set_option(d2t_engine* e, const char* key, int value);

/*
 * Image preprocessing on the device.  Replaces, for a LIST of differently sized grey crops, what
 * utils/predict_utils.py::resize (14-115; imgH None, no resizer) does per image on the host with PIL / cv2 / albumentations:
 * optional integer cv2.INTER_AREA down-sampling (:33-44), data_utils.py::pad = min-max stretch + polarity + crop to the ink
 * bounding box + black padding to multiples of 32 (10-47), minmax_size = Pillow LANCZOS shrink to max_dimension / white
 * canvas up to min_dimension (62-82), Normalize(mean, std) and the channel-0 slice (math_transform.py:42-50).
 * Two calls, because the output size of an image depends on its ink box: d2t_prep_measure -> the host plans sizes and
 * buckets (doc2tex_b200/preprocess.py) -> d2t_prep_render.  All pointers are device pointers unless noted.
 */
typedef struct d2t_prep_image {
  int64_t src_off;   /* byte offset of the image (row-major uint8, h0 x w0) in the packed buffer */
  int32_t h0, w0;    /* source size */
  int32_t ds;        /* integer down-sampling ratio applied on the fly (1 = none); h0, w0 must be divisible by it */
  int32_t pad_;
} d2t_prep_image;

typedef struct d2t_prep_plan {
  int32_t use_crop, crop_x, crop_y, crop_w, crop_h, inverted, vmin;   /* data_utils.py::pad (from d2t_prep_measure) */
  int32_t hb, wb;                 /* stage-B image: the padded crop, or the (down-sampled) source */
  int32_t do_resize, rh, rw;      /* Pillow LANCZOS shrink to (rh, rw) */
  int32_t kx_off, kx_ksize, ky_off, ky_ksize;   /* coefficient tables (int32 rows [first, count, k...]) in coefs_dev */
  int32_t out_h, out_w;           /* final canvas (white beyond the image) = the bucket's (H, W) */
  int64_t off_b, off_t, off_r;    /* byte offsets of stage B / horizontal-pass / resized images in the scratch buffer */
  void* dst;                      /* fp32 (out_h, out_w) slot of the bucket tensor */
} d2t_prep_plan;

/* stats_dev: int32 [n][8] = {x, y, w, h, inverted, vmin, status, 0}; status 0 ok, 1 blank image, 2 no pixel crosses the threshold. */
int d2t_prep_measure(d2t_engine* e, const uint8_t* packed_dev, const d2t_prep_image* imgs_dev, int n, int32_t* stats_dev,
                     d2t_stream stream);
/* sub = mean * 255, mul = 1 / (std * 255), both rounded to fp32 by the caller (albumentations' arithmetic). */
int d2t_prep_render(d2t_engine* e, const uint8_t* packed_dev, const d2t_prep_image* imgs_dev, const d2t_prep_plan* plans_dev,
                    int n, const int32_t* coefs_dev, uint8_t* scratch_dev, int any_resize, float sub, float mul, d2t_stream stream);

/*
 * Test / profiling hooks (not part of the reference surface).
 */
/* keep_taps != 0: d2t_encode keeps every intermediate activation alive for d2t_debug_tap. */
int d2t_set_debug(d2t_engine* e, int keep_taps);
/* Copies an intermediate activation of the LAST d2t_encode call to out_dev as NCHW/BND
 * fp32 so it can be compared with the oracle taps.  names: "conv0_1","conv0_2","layer1",
 * "conv1","layer2","conv2","pool3","layer3","conv3","layer4","conv4_1","conv4_2",
 * "patch_embed","block0".."blockN".  *numel receives the element count (query with
 * out_dev == NULL). */
int d2t_debug_tap(d2t_engine* e, const char* name, float* out_dev, int64_t* numel,
                  int64_t* shape4, d2t_stream stream);
/* C[M,N] = act(A[M,K] * W[N,K]^T * scale[n] + shift[n]) through the configured GEMM
 * path (unit parity test of the contraction kernels, including the tcgen05 ones). */
int d2t_debug_gemm(d2t_engine* e, const float* a_dev, const float* w_dev,
                   const float* scale_dev, const float* shift_dev, float* c_dev,
                   int M, int N, int K, int act, int precision, d2t_stream stream);
/* `iters` back-to-back launches of one contraction (optionally interleaved with a LayerNorm launch, the
 * decode-step pattern); *ms_out = average milliseconds per iteration (CUDA events). */
int d2t_debug_gemm_bench(d2t_engine* e, const float* a_dev, const float* w_dev, float* c_dev, int M, int N,
                         int K, int precision, int iters, int interleave, float* ms_out, d2t_stream stream);
/* Live timing of the dominant kernel (bench.py's roofline): after d2t_set_option(e, "time_conv", 1) every d2t_encode
 * brackets the launch of ONE layer3 3x3 convolution (512 -> 512 channels, the shape that carries 80 % of the encoder
 * FLOPs) with CUDA events on the launching stream.  This call synchronises, returns the sum of the
 * bracketed durations (ms), the number of launches timed and the algorithmic FLOPs of one launch (2*M*N*K), and
 * clears the record. */
int d2t_debug_conv_time(d2t_engine* e, double* total_ms, int64_t* launches, double* flops_per_launch);
/* Live timing of the memory-bound decode kernels (bench.py's decode roofline): after d2t_set_option(e, "time_decode", 1) the
 * decode loop runs eagerly (events cannot be timed inside a captured graph) and brackets, with CUDA events on the launching
 * stream, the launches of kind 0 = self-attention and 1 = cross-attention of decoder layer 1 (decode_attention_*_kernel),
 * 2 = beam_step_kernel, 3 = greedy_pick_kernel.  This call synchronises and returns, for one kind, the summed durations (ms),
 * the number of launches and their summed ALGORITHMIC bytes (rows x (t+1) x 2 x d_model x elem for the
 * self-attention of step t; images x ntok x 2 x d_model x elem for the cross-attention), and clears that record. */
int d2t_debug_decode_time(d2t_engine* e, int kind, double* total_ms, int64_t* launches, double* total_bytes);
/* Near-tie audit of the TFM beam search (tests): runner_up_dev = a caller-owned (B, max_steps) fp32 buffer that every later
 * d2t_decode_beam call fills with the cumulative score of the best candidate it did NOT select at each step (-inf when there is
 * none; untouched after an image finished); NULL switches the audit off.  The margin of a decision is the smallest gap between
 * neighbours among trace_score_dev[b, t, :k] and this value. */
int d2t_debug_beam_runner_up(d2t_engine* e, float* runner_up_dev);
/* Number of kernel launches issued by this engine since creation (bench bookkeeping;
 * launches replayed from a CUDA graph are counted per replay). */
int64_t d2t_launch_count(const d2t_engine* e);

#ifdef __cplusplus
}
#endif
#endif /* DOC2TEX_B200_H_ */

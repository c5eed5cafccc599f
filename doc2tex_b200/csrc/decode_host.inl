// Host side of the decode loops (included at the end of engine.cu).
//
// One decode step = a fixed kernel sequence that reads the step index from device memory, so the
// same CUDA graph is replayed for every step; the host only polls a pinned counter every few steps
// for the early-exit conditions of the reference (tfm.py:138-140, 174).

namespace {

constexpr int TFM_PAD = 0, TFM_GO = 1, TFM_END = 2;  // TFMLabelConverter (tfm_converter.py:8)
constexpr int POLL_EVERY = 8;

struct TfmBuffers {
  float *crosskv = nullptr, *selfkv = nullptr;
  float* crosskv_f32 = nullptr;   // bf16 KV mode: fp32 staging of one layer's cross K/V projection
  float *x = nullptr, *x2 = nullptr, *q = nullptr, *att = nullptr, *ffn = nullptr, *logits = nullptr;
  int *tokens = nullptr, *anc = nullptr, *n_live = nullptr, *n_done = nullptr, *finished = nullptr;
  int *done_seq = nullptr, *done_len = nullptr, *ended = nullptr, *counters = nullptr, *trace = nullptr;
  float *scores = nullptr, *done_score = nullptr, *trace_score = nullptr, *logits_out = nullptr;
  float* runner_up = nullptr;   // caller buffer (d2t_debug_beam_runner_up), offset to this group's first image
  long long* ids = nullptr;
  // bf16 hi/lo operand planes of the decoder activations + their TMA maps (tensor-core precisions only)
  long long* dbg = nullptr;   // D2T_DBG_DECODE=1: phase timestamps of one decode-step GEMM
  long long* timeline = nullptr;   // D2T_DBG_TIMELINE=1: per-launch timestamps of a step
  bool planes = false;
  int x2_parts = 1;   // x2 holds up to this many split-K partial outputs [parts][R][D]
  __nv_bfloat16 *x_hi = nullptr, *x_lo = nullptr, *att_hi = nullptr, *att_lo = nullptr, *ffn_hi = nullptr, *ffn_lo = nullptr;
  CUtensorMap map_x_hi, map_x_lo, map_att_hi, map_att_lo, map_ffn_hi, map_ffn_lo;
};

// Rows of one decode call are independent (greedy) or interact only inside one image (beam), so a call is cut into
// `decode_groups` contiguous image slices that run the same step sequence concurrently on side streams: the step is a
// chain of ~37 dependent launches of a few CTAs each, and concurrent chains fill SMs the single chain leaves idle.
struct TfmGroup {
  TfmBuffers b;
  int B0 = 0, Bg = 0;   // first image, number of images
  int R0 = 0, Rg = 0;   // first row, number of rows (images x beam)
};

template <typename T>
int pool_get(d2t_engine* e, T** out, size_t n) {
  cudaError_t st = cudaSuccess;
  *out = (T*)e->dec_pool.get(n * sizeof(T), &st);
  if (!*out) return e->fail(D2T_ERR_CUDA, "cudaMalloc of %zu bytes failed: %s", n * sizeof(T), cudaGetErrorString(st));
  return 0;
}

int dec_linear(d2t_engine* e, ConvGemm p, cudaStream_t s) {
  p.stack = e->stack_mma ? 1 : 0;   // honoured by the TMA-fed-A bf16x3 kernel only
  return run_contraction(e, p, nullptr, e->cfg.precision, s);
}

// KV caches are fp32 except in the single-pass bf16 mode, where K/V are stored as bf16 (half the attention traffic;
// the operands of every projection are bf16 there anyway).  Buffers keep their float* type; element offsets are equal.
bool kv_is_bf16(const d2t_engine* e) { return e->cfg.precision == D2T_PREC_BF16 && e->kv_bf16; }

// option "time_decode": CUDA events around one launch (eager decode only; events cannot be timed inside a captured graph)
struct DecTimer {
  d2t_engine* e; cudaStream_t s; bool on; d2t_engine::DecEvent ev;
  DecTimer(d2t_engine* e_, bool want, int kind, double bytes, cudaStream_t s_) : e(e_), s(s_), on(false) {
    if (!want || !e->time_decode || e->dec_events.size() >= 8192) return;
    ev.kind = kind; ev.bytes = bytes;
    if (cudaEventCreate(&ev.a) != cudaSuccess) return;
    if (cudaEventCreate(&ev.b) != cudaSuccess) { cudaEventDestroy(ev.a); return; }
    cudaEventRecord(ev.a, s);
    on = true;
  }
  ~DecTimer() {
    if (!on) return;
    cudaEventRecord(ev.b, s);
    e->dec_events.push_back(ev);
  }
};

int enqueue_attention(d2t_engine* e, const float* q, const float* kv, long long row_stride, const int* anc,
                      long long anc_parity, int anc_ld, int rows_per_src, const int* step, int n_fixed,
                      float* out, __nv_bfloat16* out_hi, __nv_bfloat16* out_lo, int R, cudaStream_t s) {
  const int heads = e->cfg.dec_heads, D = e->cfg.hidden;
  if (heads > 8) return e->fail(D2T_ERR_UNSUPPORTED, "decode attention supports at most 8 heads per row block");
  // few rows: the loads in flight per SM, not the bandwidth, bound the kernel -> two warps per (row, head)
  const int split = e->attn_split > 0 ? e->attn_split : (R <= 4 * e->num_sms && heads == 8 ? 2 : 1);
  const bool kv16 = kv_is_bf16(e);
  const __nv_bfloat16* kvh = reinterpret_cast<const __nv_bfloat16*>(kv);
  cudaError_t st;
  // Measured alternatives for beam search, both slower than this per-row walk and removed (profiles/
  // r02_ncu_beam_attention_variants.txt): one block per image so that L1 serves the records the hypotheses share (L1 hit
  // rate 8 -> 79 %, 42.7 vs 35.0 us: the walk is bound by dependent round trips x waves, not by L2 traffic), and a kernel
  // that stages every record of an (image, head) in shared memory with cp.async (2 460 instructions per warp: issue-bound).
  // A third, beam-grouped kernel (one warp = all hypotheses of an (image, head): every shared record loaded ONCE into
  // registers) cut the L2 -> SM traffic five-fold and ran exactly as long as this walk at 1 280 and 5 120 rows (step 415 vs
  // 414 us, 1 204 vs 1 190 us): at these sizes the walk is bound by its own dot / shuffle / exp instruction stream.
  // What helps is more loads in flight per warp: attn_kpi keys per quarter warp and iteration (4: -7 .. -11 % per step).
  const int kpi = e->attn_kpi;
  // A fourth measured alternative for beam search, also removed: four lanes per key with 256-bit loads and eight channels per
  // lane (a third fewer instructions per key, 41 instead of 57 % issue utilisation) ran 3-6 % SLOWER (profiles/
  // r02g_attn_lpk_ab.txt): the walk is bound by load latency x resident warps, which is what five blocks per SM address.
#define D2T_ROW_ATTN(SP, HB, TKV, KPI_, GRID, BLOCK, PTR)                                                                         \
  launch_kernel(decode_attention_kernel<32, SP, HB, TKV, KPI_>, GRID, BLOCK, 0, s, q, D, PTR, row_stride, 2 * D, anc, anc_parity,   \
                anc_ld, rows_per_src, step, n_fixed, out, D, out_hi, out_lo)
#define D2T_ROW_ATTN_KPI(SP, HB, TKV, GRID, BLOCK, PTR)                                                                           \
  (kpi >= 8 ? D2T_ROW_ATTN(SP, HB, TKV, 8, GRID, BLOCK, PTR) : kpi >= 4 ? D2T_ROW_ATTN(SP, HB, TKV, 4, GRID, BLOCK, PTR)            \
                                                                      : D2T_ROW_ATTN(SP, HB, TKV, 2, GRID, BLOCK, PTR))
  if (kv16) {
    if (split >= 2 && heads == 8) st = D2T_ROW_ATTN_KPI(2, 2, __nv_bfloat16, dim3(R * 2), dim3(256), kvh);
    else st = D2T_ROW_ATTN_KPI(1, 1, __nv_bfloat16, dim3(R), dim3(heads * 32), kvh);
  } else if (split >= 2 && heads == 8) {
    st = D2T_ROW_ATTN_KPI(2, 2, float, dim3(R * 2), dim3(256), kv);
  } else {
    st = D2T_ROW_ATTN_KPI(1, 1, float, dim3(R), dim3(heads * 32), kv);
  }
#undef D2T_ROW_ATTN_KPI
#undef D2T_ROW_ATTN
  CUDA_TRY(e, st);
  e->launches += 1;
  return 0;
}

// One decoder step for R rows (tfm.py:125-135 / 152-169 with a KV cache):
// nn.TransformerDecoderLayer defaults = post-norm, ReLU, eps 1e-5.
// x = LayerNorm(g) with g = sublayer GEMM + bias + residual (post-norm decoder layer): GEMM -> b.x2 (split-K partials),
// then the LayerNorm kernel, which adds the slices.
int linear_ln(d2t_engine* e, ConvGemm g, const TfmBuffers& b, const float* lw, const float* lb, int R, int D, cudaStream_t s) {
  const bool tc = e->cfg.precision == D2T_PREC_BF16X3 || e->cfg.precision == D2T_PREC_BF16;
  // split-K over extra CTAs: the serial tcgen05.mma chain of one tile (K/16 x passes instructions, ~90 ns each) is the
  // critical path of these projections; the LayerNorm kernel adds the slices
  int parts = 1;
  if (tc && b.planes && e->split_k > 0 && g.a_map_hi != nullptr && e->tcw.count(g.w) && b.x2_parts > 1) {
    const int nkb = g.K / 64, tiles = ((R + TC_BM - 1) / TC_BM) * ((g.N + 63) / 64);
    while (parts * 2 <= b.x2_parts && nkb % (parts * 2) == 0 && tiles * parts * 2 <= e->active_sms) parts *= 2;
  }
  if (parts > 1) { g.k_splits = parts; g.split_stride = (long long)R * D; }
  if (int rc = dec_linear(e, g, s)) return rc;
  return layernorm(e, b.x2, lw, lb, b.x, R, D, 1e-5f, s, b.x_hi, b.x_lo, parts, (long long)R * D);
}

}  // namespace

namespace {

// D2T_DBG_TIMELINE=1: a one-thread kernel between the launches of a step records globaltimer, so the last step's
// per-launch durations (each inflated by one extra launch gap) can be printed.  Debug only.
__global__ void timeline_stamp_kernel(long long* buf, int idx) {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  buf[idx] = t;
}
std::vector<const char*> g_tl_names;
const char* timeline_name(int i) { return i < (int)g_tl_names.size() ? g_tl_names[i] : "?"; }
struct Timeline {
  long long* buf = nullptr; int n = 0; cudaStream_t s = nullptr; std::vector<const char*>* names = nullptr;
  void mark(const char* what) {
    if (!buf || n >= 63) return;
    timeline_stamp_kernel<<<1, 1, 0, s>>>(buf, n);
    if (names && (int)names->size() <= n) names->push_back(what);
    ++n;
  }
};

struct PdlScope {   // kernels enqueued inside the scope are chained with programmatic dependent launch
  bool prev;
  explicit PdlScope(bool on) : prev(pdl_enabled()) { pdl_enabled() = on; }
  ~PdlScope() { pdl_enabled() = prev; }
};

int enqueue_tfm_step(d2t_engine* e, const TfmBuffers& b, int R, int B, int ntok, int beam, int T, bool want_logits,
                     cudaStream_t s) {
  // Programmatic dependent launch pays while the step is a chain of small launches (256 rows: 229 vs 278 us per step);
  // with thousands of rows the pre-launched CTAs of the next kernel only take SM slots from the running one
  // (5 120 beam rows: 1 195 us with, 1 097 us without) -> off above pdl_max_rows.
  PdlScope pdl(e->use_pdl && R <= e->pdl_max_rows);
  const d2t_config& c = e->cfg;
  const int D = c.hidden, F = c.dec_ff, V = c.vocab, L = T + 1;
  int* step = b.counters;
  const long long par = beam > 0 ? (long long)R * L : 0;
  int rc;
  auto from_x = [&](ConvGemm& g) { if (b.planes) { g.a_map_hi = &b.map_x_hi; g.a_map_lo = b.x_lo ? &b.map_x_lo : nullptr; } };
  auto from_att = [&](ConvGemm& g) { if (b.planes) { g.a_map_hi = &b.map_att_hi; g.a_map_lo = b.att_lo ? &b.map_att_lo : nullptr; } };
  auto from_ffn = [&](ConvGemm& g) { if (b.planes) { g.a_map_hi = &b.map_ffn_hi; g.a_map_lo = b.ffn_lo ? &b.map_ffn_lo : nullptr; } };
  // Merged decode calls (thousands of rows): a projection with more 128 x 128 tiles than SMs no longer fits the one-tile-per-CTA
  // kernel and used to fall back to the fp32 register-gather kernel (timeline at 5 120 rows: lin1 34 us, vocab 19 us for
  // 8 / 4 GFLOP).  Such projections read the same operand planes through the stem's persistent kernels instead (TMA-fed A,
  // CTA pair when the tile count fills the machine): rows as the pixels of one image row, like the ViT Linears.
  auto from_x_wide = [&](ConvGemm& g) {
    const bool wide = b.planes && e->wide_decode && e->tc3.count(g.w) &&
                      (long long)((R + TC_BM - 1) / TC_BM) * ((g.N + 127) / 128) > e->active_sms;
    if (!wide) return from_x(g);
    ConvGemm keep = g;
    g.x_hi = b.x_hi; g.x_lo = b.x_lo; g.B = 1; g.W = R; g.OW = R;
    if (!routes_to_tc3(e, g)) { g = keep; return from_x(g); }   // e.g. option "tc3" off: stay on the former path
    // 128-wide single-CTA tiles (two epilogue warps per TMEM quadrant): with K = 256 these problems are bound by their epilogue
    // and the pair kernel's 256 x 256 tiles quantise badly (lin1 at 5 120 rows: 80 tiles on 74 pairs, 34 us — no gain)
    g.single_cta = 1;
  };
  Timeline tl; tl.buf = b.timeline; tl.s = s; tl.names = &g_tl_names;
  tl.mark("start");
  {
  // greedy chain: the pick kernel of step t already wrote x of step t + 1 and advanced the counter (tfm_decode embeds GO once)
  const bool fused_pick = beam == 0 && e->fuse_pick;
  if (!fused_pick) {
  CUDA_TRY(e, launch_kernel(embed_tokens_kernel, dim3((R * D / 4 + 255) / 256), dim3(256), 0, s, b.tokens, L, step, par,
                            e->dev[PRED + "word_embed.weight"], e->dev[PRED + "pos_enc.pe"], b.x, R, D, sqrtf((float)D),
                            b.x_hi, b.x_lo));
  e->launches += 1;
  }
  tl.mark("embed");
  for (int l = 0; l < c.dec_layers; ++l) {
    const std::string p = PRED + "model.layers." + std::to_string(l) + ".";
    // element offset of layer l; a bf16 cache addresses 2-byte elements from the same base
    const size_t kvdiv = kv_is_bf16(e) ? 2 : 1;
    float* selfkv = b.selfkv + (size_t)l * R * T * 2 * D / kvdiv;
    const float* crosskv = b.crosskv + (size_t)l * B * ntok * 2 * D / kvdiv;
    // self-attention: in_proj (q -> b.q, k|v -> cache slot t), attention over the prefix, out_proj + residual, norm1
    {
      ConvGemm g = linear_params(b.x, e->dev[p + "self_attn.in_proj_weight"], e->dev[p + "self_attn.in_proj_bias"], b.q, R, 3 * D, D);
      g.ldc = D; g.n_split = D; g.out2 = selfkv; g.ldc2 = T * 2 * D; g.dyn = step; g.dyn_mul2 = 2 * D;
      g.out2_bf16 = kv_is_bf16(e) ? 1 : 0;
      from_x_wide(g);
      if (l == 1) g.dbg = b.dbg;
      if ((rc = dec_linear(e, g, s))) return rc;
    }
    if (l == 1) tl.mark("qkv");
    {
      // algorithmic bytes: every row reads the K and V records of its t + 1 prefix positions
      DecTimer tm(e, l == 1, 0, (double)R * (e->cur_step + 1) * 2 * D * (kv_is_bf16(e) ? 2 : 4), s);
      if ((rc = enqueue_attention(e, b.q, selfkv, (long long)T * 2 * D, beam > 0 ? b.anc : nullptr, par, L,
                                  beam > 0 ? beam : 1, step, 0, b.att, b.att_hi, b.att_lo, R, s))) return rc;
    }
    if (l == 1) tl.mark("self-attn");
    {
      ConvGemm g = linear_params(b.att, e->dev[p + "self_attn.out_proj.weight"], e->dev[p + "self_attn.out_proj.bias"], b.x2, R, D, D);
      g.res = b.x; g.ldr = D; from_att(g);
      if ((rc = linear_ln(e, g, b, e->dev[p + "norm1.weight"], e->dev[p + "norm1.bias"], R, D, s))) return rc;
    }
    if (l == 1) tl.mark("o1+LN1");
    // cross-attention over the encoder memory (K/V projected once per image, shared by its beams), norm2
    {
      ConvGemm g = linear_params(b.x, e->dev[p + "multihead_attn.in_proj_weight"], e->dev[p + "multihead_attn.in_proj_bias"], b.q, R, D, D);
      from_x(g);
      if ((rc = dec_linear(e, g, s))) return rc;
    }
    if (l == 1) tl.mark("q2");
    {
      // algorithmic bytes: the encoder memory of an image is shared by its hypotheses -> B x ntok records
      DecTimer tm(e, l == 1, 1, (double)B * ntok * 2 * D * (kv_is_bf16(e) ? 2 : 4), s);
      if ((rc = enqueue_attention(e, b.q, crosskv, (long long)ntok * 2 * D, nullptr, 0, 0, beam > 0 ? beam : 1, nullptr,
                                  ntok, b.att, b.att_hi, b.att_lo, R, s))) return rc;
    }
    if (l == 1) tl.mark("cross-attn");
    {
      ConvGemm g = linear_params(b.att, e->dev[p + "multihead_attn.out_proj.weight"], e->dev[p + "multihead_attn.out_proj.bias"], b.x2, R, D, D);
      g.res = b.x; g.ldr = D; from_att(g);
      if ((rc = linear_ln(e, g, b, e->dev[p + "norm2.weight"], e->dev[p + "norm2.bias"], R, D, s))) return rc;
    }
    if (l == 1) tl.mark("o2+LN2");
    // feed-forward, norm3
    {
      ConvGemm g = linear_params(b.x, e->dev[p + "linear1.weight"], e->dev[p + "linear1.bias"], b.ffn, R, F, D);
      g.act = ACT_RELU; from_x_wide(g);
      g.out_hi = b.ffn_hi; g.out_lo = b.ffn_lo;
      // wide path: nobody reads the fp32 copy as long as linear2 still fits its one-tile-per-CTA kernel, which reads the planes
      if (g.x_hi && (long long)((R + TC_BM - 1) / TC_BM) * ((D + 127) / 128) <= e->active_sms) g.out = nullptr;
      if ((rc = dec_linear(e, g, s))) return rc;
    }
    if (l == 1) tl.mark("lin1");
    {
      ConvGemm g = linear_params(b.ffn, e->dev[p + "linear2.weight"], e->dev[p + "linear2.bias"], b.x2, R, D, F);
      g.res = b.x; g.ldr = D; from_ffn(g);
      if ((rc = linear_ln(e, g, b, e->dev[p + "norm3.weight"], e->dev[p + "norm3.bias"], R, D, s))) return rc;
    }
    if (l == 1) tl.mark("lin2+LN3");
    if (l != 1) tl.mark("layer");
  }
  {
    ConvGemm g = linear_params(b.x, e->dev[PRED + "proj.weight"], e->dev[PRED + "proj.bias"], b.logits, R, V, D);
    from_x_wide(g);
    if ((rc = dec_linear(e, g, s))) return rc;
  }
  tl.mark("vocab");
  }  // launch-per-sublayer chain
  if (beam > 0) {
    BeamState st{};
    st.tokens = b.tokens; st.anc = b.anc; st.scores = b.scores; st.n_live = b.n_live; st.n_done = b.n_done;
    st.finished = b.finished; st.done_seq = b.done_seq; st.done_len = b.done_len; st.done_score = b.done_score;
    st.counters = b.counters; st.trace = b.trace; st.trace_score = b.trace_score;
    st.runner_up = b.runner_up;
    st.L = L; st.beam = beam; st.B = B; st.V = V; st.end_id = TFM_END; st.max_steps = T;
    // algorithmic bytes: the logits of every row, and the token + ancestry prefixes read from one buffer and written to the other
    DecTimer tm(e, true, 2, (double)R * V * 4 + 4.0 * R * (e->cur_step + 1) * 4, s);
    CUDA_TRY(e, launch_kernel(beam_step_kernel, dim3(B), dim3(256), (size_t)beam * V * sizeof(float), s, b.logits, st));
  } else if (e->fuse_pick) {
    DecTimer tm(e, true, 3, (double)R * V * 4 * (want_logits ? 2 : 1) + (double)R * D * 4, s);
    CUDA_TRY(e, launch_kernel(greedy_pick_kernel, dim3(R), dim3(128), 0, s, b.logits, V, step, b.tokens, L, b.ids, T,
                              want_logits ? b.logits_out : nullptr, b.ended, b.counters + 1, b.counters + 2, R, TFM_END,
                              e->dev[PRED + "word_embed.weight"], e->dev[PRED + "pos_enc.pe"], b.x, b.x_hi, b.x_lo, D,
                              sqrtf((float)D), step, b.counters + 3));
    e->launches += 1;
    tl.mark("pick+embed+advance");
    return 0;
  } else {
    CUDA_TRY(e, launch_kernel(greedy_pick_kernel, dim3(R), dim3(128), 0, s, b.logits, V, step, b.tokens, L, b.ids, T,
                              want_logits ? b.logits_out : nullptr, b.ended, b.counters + 1, b.counters + 2, R, TFM_END,
                              nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0.f, nullptr, nullptr));
  }
  e->launches += 1;
  tl.mark("pick");
  CUDA_TRY(e, launch_kernel(advance_step_kernel, dim3(1), dim3(1), 0, s, step));
  e->launches += 1;
  tl.mark("advance");
  tl.mark("(stamp only)");
  return 0;
}

int choose_groups(const d2t_engine* e, int B, int R) {
  int g = e->decode_groups;
  // auto = one chain: measured on B200 (B=256, 151 steps) groups of 1/2/4/8 decode in 47.2/49.1/47.1/48.4 ms (greedy) and
  // 70.8/68.1/66.6/67.0 ms (beam-5) — a step's duration is set by the length of the dependent chain, not by its rows
  if (g <= 0) g = 1;
  if (g > D2T_MAX_GROUPS) g = D2T_MAX_GROUPS;
  if (g > B) g = B;
  return g < 1 ? 1 : g;
}

int alloc_group(d2t_engine* e, TfmGroup& grp, int ntok, int beam, int T, bool want_logits, int* counters, cudaStream_t s) {
  const d2t_config& c = e->cfg;
  const int D = c.hidden, F = c.dec_ff, V = c.vocab, L = T + 1, nl = c.dec_layers;
  const int B = grp.Bg, R = grp.Rg;
  TfmBuffers& b = grp.b;
  int rc;
  if ((rc = pool_get(e, &b.crosskv, (size_t)nl * B * ntok * 2 * D))) return rc;
  if ((rc = pool_get(e, &b.selfkv, (size_t)nl * R * T * 2 * D))) return rc;
  if (kv_is_bf16(e) && (rc = pool_get(e, &b.crosskv_f32, (size_t)B * ntok * 2 * D))) return rc;
  if ((rc = pool_get(e, &b.x, (size_t)R * D))) return rc;
  b.x2_parts = 8;
  if ((rc = pool_get(e, &b.x2, (size_t)b.x2_parts * R * D))) return rc;
  if ((rc = pool_get(e, &b.q, (size_t)R * D))) return rc;
  if ((rc = pool_get(e, &b.att, (size_t)R * D))) return rc;
  if ((rc = pool_get(e, &b.ffn, (size_t)R * F))) return rc;
  if ((rc = pool_get(e, &b.logits, (size_t)R * V))) return rc;
  b.counters = counters;
  if (e->dbg_decode && grp.B0 == 0) {
    if ((rc = pool_get(e, &b.dbg, 64))) return rc;
    CUDA_TRY(e, cudaMemsetAsync(b.dbg, 0, 64 * sizeof(long long), s));
  }
  if (e->dbg_timeline && grp.B0 == 0) {
    if ((rc = pool_get(e, &b.timeline, 64))) return rc;
    CUDA_TRY(e, cudaMemsetAsync(b.timeline, 0, 64 * sizeof(long long), s));
  }
  b.planes = c.precision == D2T_PREC_BF16X3 || c.precision == D2T_PREC_BF16;
  if (b.planes) {
    const bool lo = c.precision == D2T_PREC_BF16X3;
    if ((rc = pool_get(e, &b.x_hi, (size_t)R * D))) return rc;
    if ((rc = pool_get(e, &b.att_hi, (size_t)R * D))) return rc;
    if ((rc = pool_get(e, &b.ffn_hi, (size_t)R * F))) return rc;
    if (lo) {
      if ((rc = pool_get(e, &b.x_lo, (size_t)R * D))) return rc;
      if ((rc = pool_get(e, &b.att_lo, (size_t)R * D))) return rc;
      if ((rc = pool_get(e, &b.ffn_lo, (size_t)R * F))) return rc;
    }
    cudaError_t st = tc_make_act_map(b.x_hi, R, D, &b.map_x_hi);
    if (st == cudaSuccess) st = tc_make_act_map(b.att_hi, R, D, &b.map_att_hi);
    if (st == cudaSuccess) st = tc_make_act_map(b.ffn_hi, R, F, &b.map_ffn_hi);
    if (st == cudaSuccess && lo) st = tc_make_act_map(b.x_lo, R, D, &b.map_x_lo);
    if (st == cudaSuccess && lo) st = tc_make_act_map(b.att_lo, R, D, &b.map_att_lo);
    if (st == cudaSuccess && lo) st = tc_make_act_map(b.ffn_lo, R, F, &b.map_ffn_lo);
    if (st != cudaSuccess) return e->fail(D2T_ERR_CUDA, "activation tensor map: %s", cudaGetErrorString(st));
  }
  const int nbuf = beam > 0 ? 2 : 1;
  if ((rc = pool_get(e, &b.tokens, (size_t)nbuf * R * L))) return rc;
  if (beam > 0) {
    if ((rc = pool_get(e, &b.anc, (size_t)2 * R * L))) return rc;
    if ((rc = pool_get(e, &b.scores, (size_t)R))) return rc;
    if ((rc = pool_get(e, &b.n_live, (size_t)B))) return rc;
    if ((rc = pool_get(e, &b.n_done, (size_t)B))) return rc;
    if ((rc = pool_get(e, &b.finished, (size_t)B))) return rc;
    if ((rc = pool_get(e, &b.done_seq, (size_t)R * L))) return rc;
    if ((rc = pool_get(e, &b.done_len, (size_t)R))) return rc;
    if ((rc = pool_get(e, &b.done_score, (size_t)R))) return rc;
    if ((rc = pool_get(e, &b.trace, (size_t)B * T * beam * 2))) return rc;
    if ((rc = pool_get(e, &b.trace_score, (size_t)B * T * beam))) return rc;
    CUDA_TRY(e, cudaMemsetAsync(b.trace, 0xFF, (size_t)B * T * beam * 2 * sizeof(int), s));
    CUDA_TRY(e, cudaMemsetAsync(b.trace_score, 0, (size_t)B * T * beam * sizeof(float), s));
  } else {
    if ((rc = pool_get(e, &b.ended, (size_t)R))) return rc;
    if ((rc = pool_get(e, &b.ids, (size_t)R * T))) return rc;
    CUDA_TRY(e, cudaMemsetAsync(b.ids, 0, (size_t)R * T * sizeof(long long), s));
    if (want_logits) {
      if ((rc = pool_get(e, &b.logits_out, (size_t)R * T * V))) return rc;
      CUDA_TRY(e, cudaMemsetAsync(b.logits_out, 0, (size_t)R * T * V * sizeof(float), s));
    }
  }
  init_decode_state_kernel<<<grid_for((long long)nbuf * R * L, 256, e->num_sms), 256, 0, s>>>(
      b.tokens, (long long)nbuf * R * L, L, R, beam, TFM_GO, b.anc, L, b.scores, b.n_live, b.n_done, b.finished,
      b.ended, b.counters);
  set_go_tokens_kernel<<<(R + 255) / 256, 256, 0, s>>>(b.tokens, L, beam > 0 ? (long long)R * L : 0, R, TFM_GO);
  e->launches += 2;
  CUDA_TRY(e, cudaGetLastError());
  return 0;
}

// Fork the groups' work onto the side streams (group 0 stays on s) and join them back on s.  Works both eagerly and
// under stream capture, where it becomes a graph with one parallel branch per group.
template <typename Fn>
int for_each_group_parallel(d2t_engine* e, std::vector<TfmGroup>& groups, cudaStream_t s, Fn&& fn) {
  const int G = (int)groups.size();
  if (G > 1) {
    CUDA_TRY(e, cudaEventRecord(e->ev_fork, s));
    for (int g = 1; g < G; ++g) CUDA_TRY(e, cudaStreamWaitEvent(e->side[g], e->ev_fork, 0));
  }
  int rc = 0;
  for (int g = 0; g < G && !rc; ++g) rc = fn(groups[g], g == 0 ? s : e->side[g]);
  for (int g = 1; g < G; ++g) {   // always join, also after an error, so that a capture can be ended cleanly
    cudaError_t st = cudaEventRecord(e->ev_join[g], e->side[g]);
    if (st == cudaSuccess) st = cudaStreamWaitEvent(s, e->ev_join[g], 0);
    if (st != cudaSuccess && !rc) rc = e->fail(D2T_ERR_CUDA, "group join failed: %s", cudaGetErrorString(st));
  }
  return rc;
}

int tfm_decode(d2t_engine* e, const float* ctx, int B, int ntok, int beam, int max_steps, bool stop_early,
               bool want_logits, std::vector<TfmGroup>* groups_out, int* steps_out, cudaStream_t s) {
  const d2t_config& c = e->cfg;
  if (c.head != D2T_HEAD_TFM) return e->fail(D2T_ERR_STATE, "engine was not configured with the TFM head");
  if (!e->finalized) return e->fail(D2T_ERR_STATE, "decode before d2t_finalize_weights");
  if (!ctx || B <= 0 || ntok <= 0 || max_steps <= 0) return e->fail(D2T_ERR_INVALID, "bad decode arguments");
  if (max_steps > c.max_seq_len + 1)
    return e->fail(D2T_ERR_INVALID, "max_steps %d exceeds max_seq_len+1 = %d", max_steps, c.max_seq_len + 1);
  if (beam > BEAM_MAX) return e->fail(D2T_ERR_UNSUPPORTED, "beam size %d > %d", beam, BEAM_MAX);
  CUDA_TRY(e, cudaSetDevice(e->device));
  e->active_sms = e->num_sms;
  const int D = c.hidden, T = max_steps;
  const int rows_per_img = beam > 0 ? beam : 1;
  const int nl = c.dec_layers;
  const int G = choose_groups(e, B, B * rows_per_img);
  e->dec_pool.release_all();
  std::vector<TfmGroup>& groups = *groups_out;
  groups.assign(G, TfmGroup());
  int rc;
  int* counters_all = nullptr;
  if ((rc = pool_get(e, &counters_all, (size_t)4 * G))) return rc;
  for (int g = 0; g < G; ++g) {
    TfmGroup& grp = groups[g];
    grp.B0 = (int)((long long)B * g / G);
    grp.Bg = (int)((long long)B * (g + 1) / G) - grp.B0;
    grp.R0 = grp.B0 * rows_per_img;
    grp.Rg = grp.Bg * rows_per_img;
    if ((rc = alloc_group(e, grp, ntok, beam, T, want_logits, counters_all + 4 * g, s))) return rc;
    grp.b.runner_up = (beam > 0 && e->beam_runner_up) ? e->beam_runner_up + (size_t)grp.B0 * T : nullptr;
  }
  // cross-attention K/V of the encoder memory, once per image and layer (the reference re-projects
  // them at every step: nn.MultiheadAttention inside tfm.py:130 / :165)
  rc = for_each_group_parallel(e, groups, s, [&](TfmGroup& grp, cudaStream_t gs) -> int {
    for (int l = 0; l < nl; ++l) {
      const std::string p = PRED + "model.layers." + std::to_string(l) + ".multihead_attn.";
      float* dst = grp.b.crosskv + (size_t)l * grp.Bg * ntok * 2 * D / (grp.b.crosskv_f32 ? 2 : 1);
      ConvGemm g = linear_params(ctx + (size_t)grp.B0 * ntok * D, e->dev[p + "in_proj_weight"] + (size_t)D * D,
                                 e->dev[p + "in_proj_bias"] + D,
                                 grp.b.crosskv_f32 ? grp.b.crosskv_f32 : dst,
                                 grp.Bg * ntok, 2 * D, D);
      if (int r = dec_linear(e, g, gs)) return r;
      if (grp.b.crosskv_f32) {   // bf16 KV cache
        const long long n = (long long)grp.Bg * ntok * 2 * D;
        f32_to_bf16_kernel<<<grid_for(n, 256, e->num_sms), 256, 0, gs>>>(grp.b.crosskv_f32, reinterpret_cast<__nv_bfloat16*>(dst), n);
        e->launches += 1;
        CUDA_TRY(e, cudaGetLastError());
      }
    }
    return 0;
  });
  if (rc) return rc;
  if (beam == 0 && e->fuse_pick) {   // x of step 0 = embedding of GO; later steps: written by the pick kernel
    for (TfmGroup& grp : groups) {
      const TfmBuffers& b = grp.b;
      CUDA_TRY(e, launch_kernel(embed_tokens_kernel, dim3((grp.Rg * D / 4 + 255) / 256), dim3(256), 0, s, b.tokens, T + 1,
                                b.counters, 0LL, e->dev[PRED + "word_embed.weight"], e->dev[PRED + "pos_enc.pe"], b.x, grp.Rg, D,
                                sqrtf((float)D), b.x_hi, b.x_lo));
      e->launches += 1;
    }
  }
  auto enqueue_step = [&]() -> int {
    return for_each_group_parallel(e, groups, s, [&](TfmGroup& grp, cudaStream_t gs) -> int {
      return enqueue_tfm_step(e, grp.b, grp.Rg, grp.Bg, ntok, beam, T, want_logits, gs);
    });
  };

  // ---- step graphs: `spg` consecutive steps per graph (the step index lives on the device, so one graph serves every
  // step), plus a one-step graph for the tail ----
  const int spg = e->steps_per_graph < 1 ? 1 : e->steps_per_graph;
  auto get_graph = [&](int n_steps, cudaGraphExec_t* exec_out, int* nodes_out) -> int {
    std::vector<long long> key = {(long long)B, G, ntok, beam, T, want_logits ? 1 : 0, n_steps};
    for (const TfmGroup& grp : groups) {
      const TfmBuffers& b = grp.b;
      const void* ptrs[] = {b.crosskv, b.selfkv, b.x, b.x2, b.q, b.att, b.ffn, b.logits, b.counters, b.tokens, b.anc,
                            b.scores, b.n_live, b.n_done, b.finished, b.done_seq, b.done_len, b.done_score, b.trace,
                            b.trace_score, b.ended, b.ids, b.logits_out, b.dbg, b.x_hi, b.x_lo, b.att_hi, b.att_lo, b.ffn_hi, b.ffn_lo,
                            b.crosskv_f32, b.runner_up};
      for (const void* q : ptrs) key.push_back((long long)(uintptr_t)q);
    }
    for (auto& g : e->graphs) if (g.key == key) { *exec_out = g.exec; *nodes_out = g.nodes; return 0; }
    cudaGraph_t graph = nullptr;
    CUDA_TRY(e, cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    const int64_t before = e->launches;
    int rc2 = 0;
    for (int i = 0; i < n_steps && !rc2; ++i) rc2 = enqueue_step();
    const int nodes = (int)(e->launches - before);
    e->launches = before;
    cudaError_t st = cudaStreamEndCapture(s, &graph);
    if (rc2) { if (graph) cudaGraphDestroy(graph); return rc2; }
    if (st != cudaSuccess) return e->fail(D2T_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(st));
    cudaGraphExec_t exec = nullptr;
    st = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (st != cudaSuccess) return e->fail(D2T_ERR_CUDA, "graph instantiate failed: %s", cudaGetErrorString(st));
    if (e->graphs.size() >= 32) {  // bounded cache
      cudaGraphExecDestroy(e->graphs.front().exec);
      e->graphs.erase(e->graphs.begin());
    }
    d2t_engine::GraphEntry ge; ge.key = key; ge.exec = exec; ge.nodes = nodes;
    e->graphs.push_back(ge);
    *exec_out = exec; *nodes_out = nodes;
    return 0;
  };
  const bool use_graphs = c.use_graphs && !e->time_decode;
  // the reference stops when EVERY row / image of the batch is done: all groups done, at the latest group's step
  auto all_done_step = [&]() -> int {
    int last = 0;
    for (int g = 0; g < G; ++g) {
      const int d = e->h_counters[4 * g + 2];
      if (d < 0) return -1;
      if (d > last) last = d;
    }
    return last;
  };
  // Early exit (tfm.py:138-140, :174) without draining the stream: every POLL_EVERY steps the counters are copied to
  // pinned memory and an event is recorded, but the host only LOOKS at the copy issued one poll earlier, after it has
  // already enqueued the next POLL_EVERY steps.  The GPU never idles on the round trip; an exit is noticed up to
  // 2 x POLL_EVERY steps late, and the surplus steps change nothing the caller sees (done_step is latched on the device
  // when the last row / image ends, later tokens lie beyond the returned length, finished beams are skipped).
  int executed = 0;
  bool poll_pending = false;
  while (executed < T) {
    int did = 1;
    if (use_graphs) {
      did = (spg > 1 && T - executed >= spg) ? spg : 1;
      cudaGraphExec_t exec = nullptr;
      int nodes = 0;
      if ((rc = get_graph(did, &exec, &nodes))) return rc;   // cached after the first decode of this shape
      CUDA_TRY(e, cudaGraphLaunch(exec, s));
      e->launches += nodes;
    } else {
      e->cur_step = executed;
      if ((rc = enqueue_step())) return rc;
    }
    const int before = executed;
    executed += did;
    if (stop_early && (executed / POLL_EVERY != before / POLL_EVERY) && executed < T) {
      if (poll_pending) {
        CUDA_TRY(e, cudaEventSynchronize(e->ev_poll));
        if (all_done_step() >= 0) break;
      }
      CUDA_TRY(e, cudaMemcpyAsync(e->h_counters, counters_all, (size_t)4 * G * sizeof(int), cudaMemcpyDeviceToHost, s));
      CUDA_TRY(e, cudaEventRecord(e->ev_poll, s));
      poll_pending = true;
    }
  }
  CUDA_TRY(e, cudaMemcpyAsync(e->h_counters, counters_all, (size_t)4 * G * sizeof(int), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(e, cudaStreamSynchronize(s));
  const int done_step = all_done_step();
  if (groups[0].b.dbg) {
    long long h[12];
    cudaMemcpy(h, groups[0].b.dbg, sizeof h, cudaMemcpyDeviceToHost);
    fprintf(stderr, "[decode gemm dbg R=%d] prologue %lld ns, first full +%lld, last full +%lld, last commit +%lld, "
                    "epi start +%lld, epi done +%lld, exit +%lld\n", groups[0].Rg, h[1] - h[0], h[2] - h[0], h[3] - h[0], h[4] - h[0],
            h[5] - h[0], h[6] - h[0], h[7] - h[0]);
  }
  if (groups[0].b.timeline) {
    long long h[64];
    cudaMemcpy(h, groups[0].b.timeline, sizeof h, cudaMemcpyDeviceToHost);
    fprintf(stderr, "[step timeline R=%d, last step; each interval includes one stamp launch]", groups[0].Rg);
    for (int i = 1; i < 64 && h[i] != 0; ++i) fprintf(stderr, " %s %.1f", timeline_name(i), (h[i] - h[i - 1]) * 1e-3);
    fprintf(stderr, "\n");
  }
  *steps_out = (stop_early && done_step >= 0) ? done_step : executed;
  return 0;
}

}  // namespace

extern "C" {

int d2t_decode_greedy(d2t_engine* e, const float* ctx, int B, int ntok, int max_steps, int stop_on_all_eos,
                      int64_t* ids, float* logits, int* steps_out, d2t_stream stream) {
  if (!e) return D2T_ERR_INVALID;
  if (!ids || !steps_out) return e->fail(D2T_ERR_INVALID, "ids_dev and steps_out are required");
  CUDA_TRY(e, cudaSetDevice(e->device));
  WorkStream ws(e, (cudaStream_t)stream);
  cudaStream_t s = ws.get();
  std::vector<TfmGroup> groups;
  int steps = 0;
  if (int rc = tfm_decode(e, ctx, B, ntok, 0, max_steps, stop_on_all_eos != 0, logits != nullptr, &groups, &steps, s)) return rc;
  for (const TfmGroup& grp : groups) {
    CUDA_TRY(e, cudaMemcpyAsync(ids + (size_t)grp.R0 * max_steps, grp.b.ids, (size_t)grp.Rg * max_steps * sizeof(int64_t),
                                cudaMemcpyDeviceToDevice, s));
    if (logits)
      CUDA_TRY(e, cudaMemcpyAsync(logits + (size_t)grp.R0 * max_steps * e->cfg.vocab, grp.b.logits_out,
                                  (size_t)grp.Rg * max_steps * e->cfg.vocab * sizeof(float), cudaMemcpyDeviceToDevice, s));
  }
  CUDA_TRY(e, cudaStreamSynchronize(s));
  *steps_out = steps;
  return D2T_OK;
}

int d2t_decode_beam(d2t_engine* e, const float* ctx, int B, int ntok, int beam, int max_steps, int64_t* best_ids,
                    int32_t* best_len, float* best_score, int32_t* trace, float* trace_score, int* steps_out,
                    d2t_stream stream) {
  if (!e) return D2T_ERR_INVALID;
  if (!best_ids || !best_len || !best_score || !steps_out) return e->fail(D2T_ERR_INVALID, "output pointers are required");
  if (beam < 1) return e->fail(D2T_ERR_INVALID, "beam must be >= 1");
  CUDA_TRY(e, cudaSetDevice(e->device));
  WorkStream ws(e, (cudaStream_t)stream);
  cudaStream_t s = ws.get();
  std::vector<TfmGroup> groups;
  int steps = 0;
  if (int rc = tfm_decode(e, ctx, B, ntok, beam, max_steps, true, false, &groups, &steps, s)) return rc;
  for (const TfmGroup& grp : groups) {
    const TfmBuffers& b = grp.b;
    BeamState st{};
    st.tokens = b.tokens; st.anc = b.anc; st.scores = b.scores; st.n_live = b.n_live; st.n_done = b.n_done;
    st.finished = b.finished; st.done_seq = b.done_seq; st.done_len = b.done_len; st.done_score = b.done_score;
    st.counters = b.counters; st.L = max_steps + 1; st.beam = beam; st.B = grp.Bg; st.V = e->cfg.vocab; st.end_id = TFM_END;
    st.max_steps = max_steps;
    // device step counter == number of executed steps == parity of the live token buffer
    beam_finalize_kernel<<<grp.Bg, 128, 0, s>>>(st, e->h_counters[0], (long long*)best_ids + (size_t)grp.B0 * max_steps, max_steps,
                                                best_len + grp.B0, best_score + grp.B0);
    e->launches += 1;
    CUDA_TRY(e, cudaGetLastError());
    if (trace)
      CUDA_TRY(e, cudaMemcpyAsync(trace + (size_t)grp.B0 * max_steps * beam * 2, b.trace,
                                  (size_t)grp.Bg * max_steps * beam * 2 * sizeof(int), cudaMemcpyDeviceToDevice, s));
    if (trace_score)
      CUDA_TRY(e, cudaMemcpyAsync(trace_score + (size_t)grp.B0 * max_steps * beam, b.trace_score,
                                  (size_t)grp.Bg * max_steps * beam * sizeof(float), cudaMemcpyDeviceToDevice, s));
  }
  CUDA_TRY(e, cudaStreamSynchronize(s));
  *steps_out = steps;
  return D2T_OK;
}

}  // extern "C"

#include "lstm_host.inl"

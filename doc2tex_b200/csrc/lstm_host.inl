// Host side of the Attnv2 (LSTM + coverage attention) greedy decode (included at the end of decode_host.inl).

namespace {

constexpr int ATTN_GO = 0, ATTN_END = 1;  // AttnLabelConverter (attn_converter.py:8)

inline bool is_lstm_head(int head) { return head == D2T_HEAD_ATTNV2 || head == D2T_HEAD_ATTN; }
// first encoder token the decoder attends to: AttentionV2 (seqmodel 'TFM') drops the cls token, Attention keeps it
inline int attn_tok0(const d2t_config& c) { return c.head == D2T_HEAD_ATTN ? 0 : 1; }

// the step kernel stages [taps][256] location weights + the coverage / score rows in dynamic shared memory: beyond the 48 KB
// default (very long location kernels or token sequences) the opt-in limit is raised once
int lstm_attention_prepare(d2t_engine* e, int S, int taps) {
  const size_t need = lstm_attention_smem_bytes(S, taps, 256);
  static size_t granted = 48 * 1024;
  if (need <= granted) return 0;
  if (need > 200 * 1024) return e->fail(D2T_ERR_UNSUPPORTED, "attention step needs %zu bytes of shared memory (taps %d, tokens %d)", need, taps, S);
  CUDA_TRY(e, cudaFuncSetAttribute(lstm_attention_step_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  granted = 200 * 1024;
  return 0;
}

struct AttnBuffers {
  float *keyproj = nullptr, *qp = nullptr, *xcat = nullptr, *gates = nullptr, *h = nullptr, *c = nullptr;
  float *alpha_cum = nullptr, *logits = nullptr, *logits_out = nullptr;
  int *tokens = nullptr, *ended = nullptr, *counters = nullptr;
  long long* ids = nullptr;
};

int enqueue_attn_step(d2t_engine* e, const AttnBuffers& b, const float* ctx, int B, int ntok, int T, bool want_logits,
                      cudaStream_t s) {
  const d2t_config& c = e->cfg;
  const int D = c.hidden, Hs = c.attn_hidden, V = c.vocab, L = T + 1, Kc = 2 * D + Hs;
  const int taps = 2 * c.attn_kernel_size + 1, S = ntok - attn_tok0(c);
  const std::string a = PRED + "attention_cell.attn.";
  int* step = b.counters;
  int rc;
  lstm_embed_kernel<<<(B * D / 4 + 255) / 256, 256, 0, s>>>(b.tokens, L, step, e->dev[PRED + "embedding.weight"], b.xcat, Kc, D, B, D);
  e->launches += 1;
  CUDA_TRY(e, cudaGetLastError());
  {  // query_proj(h_prev): h lives in xcat[:, 2D:2D+Hs]
    ConvGemm g = linear_params(b.h, e->dev[a + "query_proj.weight"], e->dev[a + "query_proj.bias"], b.qp, B, Hs, Hs);
    if ((rc = dec_linear(e, g, s))) return rc;
  }
  if ((rc = lstm_attention_prepare(e, S, taps))) return rc;
  lstm_attention_step_kernel<256><<<B, 256, lstm_attention_smem_bytes(S, taps, 256), s>>>(
      b.keyproj, ctx, ntok, b.qp, e->dev["attn.locM"], e->dev["attn.locc"], taps, e->dev[a + "score.weight"],
      e->dev[a + "score.bias"], b.alpha_cum, b.xcat, Kc, 1, attn_tok0(c));
  e->launches += 1;
  CUDA_TRY(e, cudaGetLastError());
  {  // LSTMCell gates over [context ; embedding ; h]
    ConvGemm g = linear_params(b.xcat, e->dev["lstm.w_cat"], e->dev["lstm.b_sum"], b.gates, B, 4 * Hs, Kc);
    if ((rc = dec_linear(e, g, s))) return rc;
  }
  lstm_pointwise_kernel<<<(B * Hs + 255) / 256, 256, 0, s>>>(b.gates, b.c, b.h, b.xcat, Kc, 2 * D, B, Hs);
  e->launches += 1;
  CUDA_TRY(e, cudaGetLastError());
  {
    ConvGemm g = linear_params(b.h, e->dev[PRED + "attention_cell.generator.weight"], e->dev[PRED + "attention_cell.generator.bias"], b.logits, B, V, Hs);
    if ((rc = dec_linear(e, g, s))) return rc;
  }
  lstm_pick_kernel<<<B, 128, 0, s>>>(b.logits, V, step, b.tokens, L, b.ids, T, want_logits ? b.logits_out : nullptr,
                                     b.ended, b.counters + 1, b.counters + 2, B, ATTN_END, T - 1);
  e->launches += 1;
  CUDA_TRY(e, cudaGetLastError());
  advance_step_kernel<<<1, 1, 0, s>>>(step);
  e->launches += 1;
  CUDA_TRY(e, cudaGetLastError());
  return 0;
}

}  // namespace

// Folds loc_proj(conv1d(.)) into one [Hs][taps] matrix + bias (both are linear, attention1D.py:146-148), and
// sums the two LSTM biases.  Called from d2t_finalize_weights.
int finalize_attn_extras(d2t_engine* e) {
  const d2t_config& c = e->cfg;
  const int Hs = c.attn_hidden, Kd = c.attn_kernel_dim, taps = 2 * c.attn_kernel_size + 1;
  const std::string a = PRED + "attention_cell.attn.";
  const HostTensor *wc, *bc, *wp, *bp, *bih, *bhh;
  int rc;
  if ((rc = need(e, a + "loc_conv.weight", &wc))) return rc;
  if ((rc = need(e, a + "loc_conv.bias", &bc))) return rc;
  if ((rc = need(e, a + "loc_proj.weight", &wp))) return rc;
  if ((rc = need(e, a + "loc_proj.bias", &bp))) return rc;
  if ((rc = need(e, PRED + "attention_cell.rnn.bias_ih", &bih))) return rc;
  if ((rc = need(e, PRED + "attention_cell.rnn.bias_hh", &bhh))) return rc;
  std::vector<float> M((size_t)Hs * taps), cv(Hs), bs(4 * Hs);
  for (int h = 0; h < Hs; ++h) {
    for (int j = 0; j < taps; ++j) {
      double acc = 0.0;
      for (int k = 0; k < Kd; ++k) acc += (double)wp->f[(size_t)h * Kd + k] * (double)wc->f[(size_t)k * taps + j];
      M[(size_t)h * taps + j] = (float)acc;
    }
    double acc = bp->f[h];
    for (int k = 0; k < Kd; ++k) acc += (double)wp->f[(size_t)h * Kd + k] * (double)bc->f[k];
    cv[h] = (float)acc;
  }
  for (int i = 0; i < 4 * Hs; ++i) bs[i] = bih->f[i] + bhh->f[i];
  float* p;
  if ((rc = upload(e, M.data(), M.size(), &p))) return rc;
  e->dev["attn.locM"] = p;
  if ((rc = upload(e, cv.data(), cv.size(), &p))) return rc;
  e->dev["attn.locc"] = p;
  if ((rc = upload(e, bs.data(), bs.size(), &p))) return rc;
  e->dev["lstm.b_sum"] = p;
  return 0;
}

extern "C" int d2t_decode_attn_greedy(d2t_engine* e, const float* ctx, int B, int ntok, int max_steps,
                                      int stop_on_all_eos, int64_t* ids, float* logits, int* steps_out,
                                      d2t_stream stream) {
  if (!e) return D2T_ERR_INVALID;
  const d2t_config& c = e->cfg;
  if (!is_lstm_head(c.head)) return e->fail(D2T_ERR_STATE, "engine was not configured with the Attn / Attnv2 head");
  if (!e->finalized) return e->fail(D2T_ERR_STATE, "decode before d2t_finalize_weights");
  if (!ctx || !ids || !steps_out || B <= 0 || ntok < 2 || max_steps <= 0) return e->fail(D2T_ERR_INVALID, "bad decode arguments");
  if (c.attn_hidden != 256 || c.hidden != 256) return e->fail(D2T_ERR_UNSUPPORTED, "Attnv2 head needs hidden_size == input_size == 256");
  CUDA_TRY(e, cudaSetDevice(e->device));
  WorkStream ws(e, (cudaStream_t)stream);
  cudaStream_t s = ws.get();
  e->active_sms = e->num_sms;
  const int D = c.hidden, Hs = c.attn_hidden, V = c.vocab, T = max_steps, L = T + 1, Kc = 2 * D + Hs, S = ntok - attn_tok0(c);
  const bool want_logits = logits != nullptr;
  e->dec_pool.release_all();
  AttnBuffers b;
  int rc;
  if ((rc = pool_get(e, &b.keyproj, (size_t)B * ntok * Hs))) return rc;
  if ((rc = pool_get(e, &b.qp, (size_t)B * Hs))) return rc;
  if ((rc = pool_get(e, &b.xcat, (size_t)B * Kc))) return rc;
  if ((rc = pool_get(e, &b.gates, (size_t)B * 4 * Hs))) return rc;
  if ((rc = pool_get(e, &b.h, (size_t)B * Hs))) return rc;
  if ((rc = pool_get(e, &b.c, (size_t)B * Hs))) return rc;
  if ((rc = pool_get(e, &b.alpha_cum, (size_t)B * S))) return rc;
  if ((rc = pool_get(e, &b.logits, (size_t)B * V))) return rc;
  if ((rc = pool_get(e, &b.tokens, (size_t)B * L))) return rc;
  if ((rc = pool_get(e, &b.ended, (size_t)B))) return rc;
  if ((rc = pool_get(e, &b.counters, 4))) return rc;
  if ((rc = pool_get(e, &b.ids, (size_t)B * T))) return rc;
  CUDA_TRY(e, cudaMemsetAsync(b.ids, 0, (size_t)B * T * sizeof(long long), s));
  CUDA_TRY(e, cudaMemsetAsync(b.alpha_cum, 0, (size_t)B * S * sizeof(float), s));
  if (want_logits) {
    if ((rc = pool_get(e, &b.logits_out, (size_t)B * T * V))) return rc;
    CUDA_TRY(e, cudaMemsetAsync(b.logits_out, 0, (size_t)B * T * V * sizeof(float), s));
  }
  init_decode_state_kernel<<<grid_for((long long)B * L, 256, e->num_sms), 256, 0, s>>>(
      b.tokens, (long long)B * L, L, B, 0, ATTN_GO, nullptr, L, nullptr, nullptr, nullptr, nullptr, b.ended, b.counters);
  e->launches += 1;
  CUDA_TRY(e, cudaGetLastError());
  const std::string a = PRED + "attention_cell.attn.";
  {  // key_proj(H) hoisted out of the loop (computed for every ctx row; the cls row is simply unused)
    ConvGemm g = linear_params(ctx, e->dev[a + "key_proj.weight"], e->dev[a + "key_proj.bias"], b.keyproj, B * ntok, Hs, D);
    if ((rc = dec_linear(e, g, s))) return rc;
  }
  for (int which = 0; which < 2; ++which) {  // h0 / c0 = proj_init_{h,c}(ctx[:, 0]) (seq2seq_v2.py:196-199)
    const std::string n = which == 0 ? "proj_init_h" : "proj_init_c";
    ConvGemm g = linear_params(ctx, e->dev[PRED + n + ".weight"], e->dev[PRED + n + ".bias"], which == 0 ? b.h : b.c, B, Hs, D);
    g.W = ntok; g.SW = ntok;  // row m reads pixel (b, 0, 0) of a [B,1,ntok,D] tensor = the cls token
    if ((rc = dec_linear(e, g, s))) return rc;
  }
  CUDA_TRY(e, cudaMemcpy2DAsync(b.xcat + 2 * D, (size_t)Kc * sizeof(float), b.h, (size_t)Hs * sizeof(float),
                                (size_t)Hs * sizeof(float), B, cudaMemcpyDeviceToDevice, s));

  cudaGraphExec_t exec = nullptr;
  int nodes = 0;
  if (c.use_graphs) {
    std::vector<long long> key = {-2, B, ntok, T, want_logits ? 1 : 0, (long long)(uintptr_t)ctx};
    const void* ptrs[] = {b.keyproj, b.qp, b.xcat, b.gates, b.h, b.c, b.alpha_cum, b.logits, b.logits_out, b.tokens,
                          b.ended, b.counters, b.ids};
    for (const void* q : ptrs) key.push_back((long long)(uintptr_t)q);
    for (auto& g : e->graphs) if (g.key == key) { exec = g.exec; nodes = g.nodes; }
    if (!exec) {
      cudaGraph_t graph = nullptr;
      CUDA_TRY(e, cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
      const int64_t before = e->launches;
      rc = enqueue_attn_step(e, b, ctx, B, ntok, T, want_logits, s);
      nodes = (int)(e->launches - before);
      e->launches = before;
      cudaError_t st = cudaStreamEndCapture(s, &graph);
      if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
      if (st != cudaSuccess) return e->fail(D2T_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(st));
      st = cudaGraphInstantiate(&exec, graph, 0);
      cudaGraphDestroy(graph);
      if (st != cudaSuccess) return e->fail(D2T_ERR_CUDA, "graph instantiate failed: %s", cudaGetErrorString(st));
      if (e->graphs.size() >= 16) { cudaGraphExecDestroy(e->graphs.front().exec); e->graphs.erase(e->graphs.begin()); }
      d2t_engine::GraphEntry ge; ge.key = key; ge.exec = exec; ge.nodes = nodes;
      e->graphs.push_back(ge);
    }
  }
  int executed = 0;
  bool poll_pending = false;   // early-exit poll checked one poll late (see tfm_decode): the stream never drains
  for (int t = 0; t < T; ++t) {
    if (exec) { CUDA_TRY(e, cudaGraphLaunch(exec, s)); e->launches += nodes; }
    else if ((rc = enqueue_attn_step(e, b, ctx, B, ntok, T, want_logits, s))) return rc;
    executed = t + 1;
    if (stop_on_all_eos && (executed % POLL_EVERY == 0) && executed < T) {
      if (poll_pending) {
        CUDA_TRY(e, cudaEventSynchronize(e->ev_poll));
        if (e->h_counters[2] >= 0) break;
      }
      CUDA_TRY(e, cudaMemcpyAsync(e->h_counters, b.counters, 4 * sizeof(int), cudaMemcpyDeviceToHost, s));
      CUDA_TRY(e, cudaEventRecord(e->ev_poll, s));
      poll_pending = true;
    }
  }
  CUDA_TRY(e, cudaMemcpyAsync(e->h_counters, b.counters, 4 * sizeof(int), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(e, cudaStreamSynchronize(s));
  const int done_step = e->h_counters[2];
  const int steps = (stop_on_all_eos && done_step >= 0) ? done_step : executed;
  // rows past the reference's break stay zero, like its untouched `probs` buffer (seq2seq_v2.py:219-225, 291)
  if (steps < T) {
    CUDA_TRY(e, cudaMemset2DAsync(b.ids + steps, (size_t)T * sizeof(long long), 0, (size_t)(T - steps) * sizeof(long long), B, s));
    if (want_logits)
      CUDA_TRY(e, cudaMemset2DAsync(b.logits_out + (size_t)steps * V, (size_t)T * V * sizeof(float), 0,
                                    (size_t)(T - steps) * V * sizeof(float), B, s));
  }
  CUDA_TRY(e, cudaMemcpyAsync(ids, b.ids, (size_t)B * T * sizeof(int64_t), cudaMemcpyDeviceToDevice, s));
  if (want_logits)
    CUDA_TRY(e, cudaMemcpyAsync(logits, b.logits_out, (size_t)B * T * V * sizeof(float), cudaMemcpyDeviceToDevice, s));
  CUDA_TRY(e, cudaStreamSynchronize(s));
  *steps_out = steps;
  return D2T_OK;
}


// ---------------------------------------------------------------------------------------------
// Batched AttentionV2.forward_beam (seq2seq_v2.py:12-174): what every shipped YAML runs at beam_size 5/10.
// Rows = images x beam hypothesis slots; the beams of an image share key_proj(H) and H.  One CUDA graph per step:
// embed -> query_proj -> fused coverage attention -> LSTM gates GEMM -> pointwise -> generator -> beam step.
// ---------------------------------------------------------------------------------------------
namespace {

struct AttnBeamBuffers {
  float *keyproj = nullptr, *qp = nullptr, *xcat = nullptr, *gates = nullptr, *h = nullptr, *c = nullptr;
  float *alpha_cum = nullptr, *logits = nullptr, *scores = nullptr, *done_score = nullptr, *trace_score = nullptr;
  int *targets = nullptr, *seqs = nullptr, *n_live = nullptr, *n_done = nullptr, *last_complete = nullptr, *finished = nullptr;
  int *done_seq = nullptr, *done_len = nullptr, *counters = nullptr, *trace = nullptr;
};

__global__ void attn_beam_init_kernel(AttnBeamState st, int R, int go_id) {
  const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gsz = gridDim.x * blockDim.x;
  for (long long i = gtid; i < (long long)R * st.L; i += gsz) st.seqs[i] = (i % st.L) == 0 ? go_id : 0;
  for (int i = gtid; i < R; i += gsz) { st.scores[i] = 0.f; st.targets[i] = go_id; }
  for (int i = gtid; i < st.B; i += gsz) { st.n_live[i] = st.beam; st.n_done[i] = 0; st.last_complete[i] = 0; st.finished[i] = 0; }
  if (gtid == 0) { st.counters[0] = 0; st.counters[1] = 0; st.counters[2] = -1; st.counters[3] = 0; }
}

size_t attn_beam_smem(int beam, int V, int Hs, int S, int L) {
  const size_t a = (size_t)beam * V, b = (size_t)beam * (2 * Hs + S + L);
  return (a > b ? a : b) * sizeof(float);
}

AttnBeamState attn_beam_state(const d2t_engine* e, const AttnBeamBuffers& b, int B, int beam, int ntok, int T) {
  const d2t_config& c = e->cfg;
  AttnBeamState st{};
  st.scores = b.scores; st.targets = b.targets; st.seqs = b.seqs; st.h = b.h; st.c = b.c;
  st.xcat = b.xcat; st.ld = 2 * c.hidden + c.attn_hidden; st.hoff = 2 * c.hidden;
  st.alpha_cum = b.alpha_cum; st.n_live = b.n_live; st.n_done = b.n_done; st.last_complete = b.last_complete;
  st.finished = b.finished; st.done_seq = b.done_seq; st.done_len = b.done_len; st.done_score = b.done_score;
  st.counters = b.counters; st.trace = b.trace; st.trace_score = b.trace_score;
  st.L = T + 1; st.beam = beam; st.B = B; st.V = c.vocab; st.S = ntok - attn_tok0(c); st.HS = c.attn_hidden; st.end_id = ATTN_END;
  st.max_steps = T;
  return st;
}

int enqueue_attn_beam_step(d2t_engine* e, const AttnBeamBuffers& b, const float* ctx, int B, int beam, int ntok, int T,
                           cudaStream_t s) {
  const d2t_config& c = e->cfg;
  const int D = c.hidden, Hs = c.attn_hidden, V = c.vocab, Kc = 2 * D + Hs, R = B * beam;
  const int taps = 2 * c.attn_kernel_size + 1, S = ntok - attn_tok0(c);
  const std::string a = PRED + "attention_cell.attn.";
  int rc;
  lstm_embed_cur_kernel<<<(R * D / 4 + 255) / 256, 256, 0, s>>>(b.targets, e->dev[PRED + "embedding.weight"], b.xcat, Kc, D, R, D);
  e->launches += 1;
  CUDA_TRY(e, cudaGetLastError());
  {
    ConvGemm g = linear_params(b.h, e->dev[a + "query_proj.weight"], e->dev[a + "query_proj.bias"], b.qp, R, Hs, Hs);
    if ((rc = dec_linear(e, g, s))) return rc;
  }
  if ((rc = lstm_attention_prepare(e, S, taps))) return rc;
  lstm_attention_step_kernel<256><<<R, 256, lstm_attention_smem_bytes(S, taps, 256), s>>>(
      b.keyproj, ctx, ntok, b.qp, e->dev["attn.locM"], e->dev["attn.locc"], taps, e->dev[a + "score.weight"],
      e->dev[a + "score.bias"], b.alpha_cum, b.xcat, Kc, beam, attn_tok0(c));
  e->launches += 1;
  CUDA_TRY(e, cudaGetLastError());
  {
    ConvGemm g = linear_params(b.xcat, e->dev["lstm.w_cat"], e->dev["lstm.b_sum"], b.gates, R, 4 * Hs, Kc);
    if ((rc = dec_linear(e, g, s))) return rc;
  }
  lstm_pointwise_kernel<<<(R * Hs + 255) / 256, 256, 0, s>>>(b.gates, b.c, b.h, b.xcat, Kc, 2 * D, R, Hs);
  e->launches += 1;
  CUDA_TRY(e, cudaGetLastError());
  {
    ConvGemm g = linear_params(b.h, e->dev[PRED + "attention_cell.generator.weight"], e->dev[PRED + "attention_cell.generator.bias"], b.logits, R, V, Hs);
    if ((rc = dec_linear(e, g, s))) return rc;
  }
  attn_beam_step_kernel<<<B, 256, attn_beam_smem(beam, V, Hs, S, T + 1), s>>>(b.logits, attn_beam_state(e, b, B, beam, ntok, T));
  e->launches += 1;
  CUDA_TRY(e, cudaGetLastError());
  advance_step_kernel<<<1, 1, 0, s>>>(b.counters);
  e->launches += 1;
  CUDA_TRY(e, cudaGetLastError());
  return 0;
}

}  // namespace

extern "C" int d2t_decode_attn_beam(d2t_engine* e, const float* ctx, int B, int ntok, int beam, int max_steps,
                                    int64_t* best_ids, int32_t* best_len, float* best_score, int32_t* trace,
                                    float* trace_score, int* steps_out, d2t_stream stream) {
  if (!e) return D2T_ERR_INVALID;
  const d2t_config& c = e->cfg;
  if (!is_lstm_head(c.head)) return e->fail(D2T_ERR_STATE, "engine was not configured with the Attn / Attnv2 head");
  if (!e->finalized) return e->fail(D2T_ERR_STATE, "decode before d2t_finalize_weights");
  if (!ctx || !best_ids || !best_len || !best_score || !steps_out || B <= 0 || ntok < 2 || max_steps <= 0)
    return e->fail(D2T_ERR_INVALID, "bad decode arguments");
  if (beam < 1 || beam > ATTN_BEAM_MAX) return e->fail(D2T_ERR_UNSUPPORTED, "beam size %d not in [1, %d]", beam, ATTN_BEAM_MAX);
  if (beam > c.vocab) return e->fail(D2T_ERR_INVALID, "beam size exceeds the vocabulary");
  if (c.attn_hidden != 256 || c.hidden != 256) return e->fail(D2T_ERR_UNSUPPORTED, "Attnv2 head needs hidden_size == input_size == 256");
  CUDA_TRY(e, cudaSetDevice(e->device));
  WorkStream ws(e, (cudaStream_t)stream);
  cudaStream_t s = ws.get();
  e->active_sms = e->num_sms;
  const int D = c.hidden, Hs = c.attn_hidden, V = c.vocab, T = max_steps, L = T + 1, Kc = 2 * D + Hs, S = ntok - attn_tok0(c), R = B * beam;
  const size_t smem = attn_beam_smem(beam, V, Hs, S, L);
  if (smem > 200 * 1024) return e->fail(D2T_ERR_UNSUPPORTED, "beam %d x %d tokens needs %zu bytes of shared memory", beam, ntok, smem);
  CUDA_TRY(e, cudaFuncSetAttribute(attn_beam_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  e->dec_pool.release_all();
  AttnBeamBuffers b;
  int rc;
  if ((rc = pool_get(e, &b.keyproj, (size_t)B * ntok * Hs))) return rc;
  if ((rc = pool_get(e, &b.qp, (size_t)R * Hs))) return rc;
  if ((rc = pool_get(e, &b.xcat, (size_t)R * Kc))) return rc;
  if ((rc = pool_get(e, &b.gates, (size_t)R * 4 * Hs))) return rc;
  if ((rc = pool_get(e, &b.h, (size_t)R * Hs))) return rc;
  if ((rc = pool_get(e, &b.c, (size_t)R * Hs))) return rc;
  if ((rc = pool_get(e, &b.alpha_cum, (size_t)R * S))) return rc;
  if ((rc = pool_get(e, &b.logits, (size_t)R * V))) return rc;
  if ((rc = pool_get(e, &b.scores, (size_t)R))) return rc;
  if ((rc = pool_get(e, &b.targets, (size_t)R))) return rc;
  if ((rc = pool_get(e, &b.seqs, (size_t)R * L))) return rc;
  if ((rc = pool_get(e, &b.n_live, (size_t)B))) return rc;
  if ((rc = pool_get(e, &b.n_done, (size_t)B))) return rc;
  if ((rc = pool_get(e, &b.last_complete, (size_t)B))) return rc;
  if ((rc = pool_get(e, &b.finished, (size_t)B))) return rc;
  if ((rc = pool_get(e, &b.done_seq, (size_t)R * L))) return rc;
  if ((rc = pool_get(e, &b.done_len, (size_t)R))) return rc;
  if ((rc = pool_get(e, &b.done_score, (size_t)R))) return rc;
  if ((rc = pool_get(e, &b.counters, 4))) return rc;
  if ((rc = pool_get(e, &b.trace, (size_t)B * T * beam * 2))) return rc;
  if ((rc = pool_get(e, &b.trace_score, (size_t)B * T * beam))) return rc;
  CUDA_TRY(e, cudaMemsetAsync(b.trace, 0xFF, (size_t)B * T * beam * 2 * sizeof(int), s));
  CUDA_TRY(e, cudaMemsetAsync(b.trace_score, 0, (size_t)B * T * beam * sizeof(float), s));
  CUDA_TRY(e, cudaMemsetAsync(b.alpha_cum, 0, (size_t)R * S * sizeof(float), s));
  AttnBeamState st = attn_beam_state(e, b, B, beam, ntok, T);
  attn_beam_init_kernel<<<grid_for((long long)R * L, 256, e->num_sms), 256, 0, s>>>(st, R, ATTN_GO);
  e->launches += 1;
  CUDA_TRY(e, cudaGetLastError());
  const std::string a = PRED + "attention_cell.attn.";
  {  // key_proj(H), once per image, shared by its beams
    ConvGemm g = linear_params(ctx, e->dev[a + "key_proj.weight"], e->dev[a + "key_proj.bias"], b.keyproj, B * ntok, Hs, D);
    if ((rc = dec_linear(e, g, s))) return rc;
  }
  for (int which = 0; which < 2; ++which) {  // h0 / c0 = proj_init_{h,c}(cls token), replicated over the beam (seq2seq_v2.py:33-46)
    const std::string n = which == 0 ? "proj_init_h" : "proj_init_c";
    ConvGemm g = linear_params(ctx, e->dev[PRED + n + ".weight"], e->dev[PRED + n + ".bias"], which == 0 ? b.h : b.c, R, Hs, D);
    g.B = B; g.W = ntok; g.OW = beam; g.SW = 0;  // output pixel (image, slot) reads input pixel (image, 0, 0) = the cls token
    if ((rc = dec_linear(e, g, s))) return rc;
  }
  CUDA_TRY(e, cudaMemcpy2DAsync(b.xcat + 2 * D, (size_t)Kc * sizeof(float), b.h, (size_t)Hs * sizeof(float),
                                (size_t)Hs * sizeof(float), R, cudaMemcpyDeviceToDevice, s));

  cudaGraphExec_t exec = nullptr;
  int nodes = 0;
  if (c.use_graphs) {
    std::vector<long long> key = {-3, B, beam, ntok, T, (long long)(uintptr_t)ctx};
    const void* ptrs[] = {b.keyproj, b.qp, b.xcat, b.gates, b.h, b.c, b.alpha_cum, b.logits, b.scores, b.targets, b.seqs,
                          b.n_live, b.n_done, b.last_complete, b.finished, b.done_seq, b.done_len, b.done_score,
                          b.counters, b.trace, b.trace_score};
    for (const void* q : ptrs) key.push_back((long long)(uintptr_t)q);
    for (auto& g : e->graphs) if (g.key == key) { exec = g.exec; nodes = g.nodes; }
    if (!exec) {
      cudaGraph_t graph = nullptr;
      CUDA_TRY(e, cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
      const int64_t before = e->launches;
      rc = enqueue_attn_beam_step(e, b, ctx, B, beam, ntok, T, s);
      nodes = (int)(e->launches - before);
      e->launches = before;
      cudaError_t stt = cudaStreamEndCapture(s, &graph);
      if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
      if (stt != cudaSuccess) return e->fail(D2T_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(stt));
      stt = cudaGraphInstantiate(&exec, graph, 0);
      cudaGraphDestroy(graph);
      if (stt != cudaSuccess) return e->fail(D2T_ERR_CUDA, "graph instantiate failed: %s", cudaGetErrorString(stt));
      if (e->graphs.size() >= 16) { cudaGraphExecDestroy(e->graphs.front().exec); e->graphs.erase(e->graphs.begin()); }
      d2t_engine::GraphEntry ge; ge.key = key; ge.exec = exec; ge.nodes = nodes;
      e->graphs.push_back(ge);
    }
  }
  int executed = 0;
  bool poll_pending = false;
  for (int t = 0; t < T; ++t) {
    if (exec) { CUDA_TRY(e, cudaGraphLaunch(exec, s)); e->launches += nodes; }
    else if ((rc = enqueue_attn_beam_step(e, b, ctx, B, beam, ntok, T, s))) return rc;
    executed = t + 1;
    if ((executed % POLL_EVERY == 0) && executed < T) {   // every image has exhausted its beam (seq2seq_v2.py:124-126)
      if (poll_pending) {
        CUDA_TRY(e, cudaEventSynchronize(e->ev_poll));
        if (e->h_counters[2] >= 0) break;
      }
      CUDA_TRY(e, cudaMemcpyAsync(e->h_counters, b.counters, 4 * sizeof(int), cudaMemcpyDeviceToHost, s));
      CUDA_TRY(e, cudaEventRecord(e->ev_poll, s));
      poll_pending = true;
    }
  }
  CUDA_TRY(e, cudaMemcpyAsync(e->h_counters, b.counters, 4 * sizeof(int), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(e, cudaStreamSynchronize(s));
  const int done_step = e->h_counters[2];
  attn_beam_finalize_kernel<<<B, 128, 0, s>>>(st, executed, (long long*)best_ids, T, best_len, best_score);
  e->launches += 1;
  CUDA_TRY(e, cudaGetLastError());
  if (trace)
    CUDA_TRY(e, cudaMemcpyAsync(trace, b.trace, (size_t)B * T * beam * 2 * sizeof(int), cudaMemcpyDeviceToDevice, s));
  if (trace_score)
    CUDA_TRY(e, cudaMemcpyAsync(trace_score, b.trace_score, (size_t)B * T * beam * sizeof(float), cudaMemcpyDeviceToDevice, s));
  CUDA_TRY(e, cudaStreamSynchronize(s));
  *steps_out = done_step >= 0 ? done_step : executed;
  return D2T_OK;
}

// Decode-step kernels of the LSTM + coverage location-aware attention head ("Attnv2", the
// config/train.yaml default stack): seq2seq_v2.py:176-293, attention1D.py:121-161, 205-242.
#pragma once
#include "common.cuh"

namespace d2t {

// xcat[b, off : off+D] = E[tokens[b][t]]      (seq2seq.py:62-63; row 0 = [GO] is the zero padding row)
__global__ void lstm_embed_kernel(const int* __restrict__ tokens, int tok_ld, const int* __restrict__ step,
                                  const float* __restrict__ emb, float* __restrict__ xcat, int ld, int off, int B, int D) {
  const int t = *step;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int d4n = D / 4;
  if (idx >= B * d4n) return;
  const int b = idx / d4n, d = (idx % d4n) * 4;
  const int tok = tokens[(size_t)b * tok_ld + t];
  *reinterpret_cast<float4*>(xcat + (size_t)b * ld + off + d) = *reinterpret_cast<const float4*>(emb + (size_t)tok * D + d);
}

// Fused attention step for one image (attention1D.py:136-161, 223-233; seq2seq_v2.py:264-266):
//   e[s]   = score_w . tanh(key_proj(H)[s] + query_proj(h)[.] + loc[s]) + score_b
//   loc[s] = loc_proj(conv1d(alpha_cum))[s] = sum_j M[:, j] * alpha_cum[s + j - pad] + cvec   (M, cvec folded at load)
//   alpha  = softmax_s(e);  context = alpha^T H;  alpha_cum += alpha
// key_proj(H) is hoisted out of the loop (the reference recomputes it every step).  HBM traffic per step
// = keyproj + H rows of the image, read once each as coalesced 1 KB rows.
// The kernel is one block per decoder row and was bound by dependent load latency (ncu: 61 us at 256 rows, 21 % of the issue
// slots, long-scoreboard stalls): a warp now scores FOUR tokens at a time with their 32 key_proj loads in flight before the
// first use, the folded location matrix is staged once in shared memory ([tap][channel], conflict-free), the coverage copy is
// zero-padded (no bounds tests), and the context pass issues 16 loads per thread before its FMAs.  Every sum keeps the
// association order of the first version, so results are bit-identical to it.
inline size_t lstm_attention_smem_bytes(int S, int taps, int HS) {
  return (size_t)(2 * S + 2 * (taps / 2) + taps * HS) * sizeof(float);
}

template <int HS>  // hidden size (256)
__global__ void __launch_bounds__(256)
lstm_attention_step_kernel(const float* __restrict__ keyproj, const float* __restrict__ ctx, int ntok,
                           const float* __restrict__ qp, const float* __restrict__ locM /*[HS][taps]*/,
                           const float* __restrict__ locc /*[HS]*/, int taps, const float* __restrict__ score_w,
                           const float* __restrict__ score_b, float* __restrict__ alpha_cum /*[B][S]*/,
                           float* __restrict__ xcat, int ld, int rows_per_img = 1,
                           int tok0 = 1 /* first attended token: 1 skips the cls token (Attnv2), 0 keeps it (Attn) */) {
  extern __shared__ float sm[];  // [S + 2 pad] zero-padded alpha_cum copy, [S] scores, [taps][HS] folded location matrix
  const int S = ntok - tok0;
  const int pad = taps / 2;
  float* s_ac = sm;                  // s_ac[pad + i] = alpha_cum[i]
  float* s_e = sm + S + 2 * pad;
  float* s_M = s_e + S;
  __shared__ float red[8];
  __shared__ float s_bc[2];
  const int b = blockIdx.x;               // decoder row (hypothesis slot)
  const int img = b / rows_per_img;       // the beams of an image share its encoder memory
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  constexpr int PER = HS / 32;
  for (int i = threadIdx.x; i < S + 2 * pad; i += blockDim.x)
    s_ac[i] = (i >= pad && i < pad + S) ? alpha_cum[(size_t)b * S + i - pad] : 0.f;
  for (int i = threadIdx.x; i < taps * HS; i += blockDim.x) {
    const int j = i / HS, h = i - j * HS;
    s_M[i] = locM[h * taps + j];
  }
  float q[PER], sw[PER];
#pragma unroll
  for (int k = 0; k < PER; ++k) {
    const int h = lane + 32 * k;
    q[k] = qp[(size_t)b * HS + h] + locc[h];
    sw[k] = score_w[h];
  }
  __syncthreads();
  const float sb = score_b[0];
  const float* const kbase = keyproj + ((size_t)img * ntok + tok0) * HS + lane;
  for (int s0 = wid; s0 < S; s0 += 4 * nw) {   // tokens s0, s0 + nw, s0 + 2 nw, s0 + 3 nw of this warp
    float kr[4][PER], loc[4][PER];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int s = min(s0 + u * nw, S - 1);   // a clamped duplicate is computed and dropped
#pragma unroll
      for (int k = 0; k < PER; ++k) {
        kr[u][k] = kbase[(size_t)s * HS + 32 * k];
        loc[u][k] = 0.f;
      }
    }
    for (int j = 0; j < taps; ++j) {
      float m[PER], a[4];
#pragma unroll
      for (int k = 0; k < PER; ++k) m[k] = s_M[j * HS + lane + 32 * k];
#pragma unroll
      for (int u = 0; u < 4; ++u) a[u] = s_ac[min(s0 + u * nw, S - 1) + j];   // = alpha_cum[s + j - pad], 0 outside
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int k = 0; k < PER; ++k) loc[u][k] = fmaf(m[k], a[u], loc[u][k]);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < PER; ++k) acc = fmaf(sw[k], tanhf(kr[u][k] + q[k] + loc[u][k]), acc);
      acc = warp_sum(acc);
      const int s = s0 + u * nw;
      if (lane == 0 && s < S) s_e[s] = acc + sb;
    }
  }
  __syncthreads();
  // softmax over s
  float mx = -INFINITY;
  for (int i = threadIdx.x; i < S; i += blockDim.x) mx = fmaxf(mx, s_e[i]);
  mx = warp_max(mx);
  if (lane == 0) red[wid] = mx;
  __syncthreads();
  if (threadIdx.x == 0) { float m = red[0]; for (int i = 1; i < nw; ++i) m = fmaxf(m, red[i]); s_bc[0] = m; }
  __syncthreads();
  mx = s_bc[0];
  float sum = 0.f;
  for (int i = threadIdx.x; i < S; i += blockDim.x) { const float p = expf(s_e[i] - mx); s_e[i] = p; sum += p; }
  sum = warp_sum(sum);
  __syncthreads();
  if (lane == 0) red[wid] = sum;
  __syncthreads();
  if (threadIdx.x == 0) { float t = 0.f; for (int i = 0; i < nw; ++i) t += red[i]; s_bc[1] = t; }
  __syncthreads();
  const float inv = 1.0f / s_bc[1];
  for (int i = threadIdx.x; i < S; i += blockDim.x) {
    const float a = s_e[i] * inv;
    s_e[i] = a;
    alpha_cum[(size_t)b * S + i] = s_ac[pad + i] + a;  // coverage update AFTER the step (seq2seq_v2.py:264-266)
  }
  __syncthreads();
  // context = alpha^T H, one thread per channel (HS == input channels == 256 here): even tokens into a0, odd into a1
  for (int d = threadIdx.x; d < HS; d += blockDim.x) {
    const float* hp = ctx + ((size_t)img * ntok + tok0) * HS + d;
    float a0 = 0.f, a1 = 0.f;
    int s = 0;
    for (; s + 16 <= S; s += 16) {
      float v[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) v[u] = hp[(size_t)(s + u) * HS];
#pragma unroll
      for (int u = 0; u < 16; u += 2) {
        a0 = fmaf(s_e[s + u], v[u], a0);
        a1 = fmaf(s_e[s + u + 1], v[u + 1], a1);
      }
    }
    for (; s + 2 <= S; s += 2) {
      a0 = fmaf(s_e[s], hp[(size_t)s * HS], a0);
      a1 = fmaf(s_e[s + 1], hp[(size_t)(s + 1) * HS], a1);
    }
    if (s < S) a0 = fmaf(s_e[s], hp[(size_t)s * HS], a0);
    xcat[(size_t)b * ld + d] = a0 + a1;
  }
}

// LSTMCell pointwise part (torch gate order i, f, g, o): c' = sig(f) c + sig(i) tanh(g); h' = sig(o) tanh(c').
__global__ void lstm_pointwise_kernel(const float* __restrict__ gates, float* __restrict__ c, float* __restrict__ h,
                                      float* __restrict__ xcat, int ld, int hoff, int B, int HS) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * HS) return;
  const int b = idx / HS, j = idx % HS;
  const float* g = gates + (size_t)b * 4 * HS;
  const float ig = 1.0f / (1.0f + expf(-g[j]));
  const float fg = 1.0f / (1.0f + expf(-g[HS + j]));
  const float gg = tanhf(g[2 * HS + j]);
  const float og = 1.0f / (1.0f + expf(-g[3 * HS + j]));
  const float cn = fg * c[idx] + ig * gg;
  const float hn = og * tanhf(cn);
  c[idx] = cn;
  h[idx] = hn;
  xcat[(size_t)b * ld + hoff + j] = hn;
}

// Greedy pick over raw logits: next = argmax(logits[b]) (first maximum; seq2seq_v2.py:283-284), records
// logits / ids, END flags and the executed-step count for the early exit (:286-289).
__global__ void lstm_pick_kernel(const float* __restrict__ logits, int V, const int* __restrict__ step,
                                 int* __restrict__ tokens, int tok_ld, long long* __restrict__ ids, int ids_ld,
                                 float* __restrict__ logits_out, int* __restrict__ ended, int* __restrict__ n_ended,
                                 int* __restrict__ done_step, int B, int end_id, int last_step) {
  __shared__ float red_v[32];
  __shared__ int red_i[32];
  const int b = blockIdx.x;
  const int t = *step;
  const float* x = logits + (size_t)b * V;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  float bv = -INFINITY; int bi = 0x7fffffff;
  for (int i = threadIdx.x; i < V; i += blockDim.x) {
    const float v = x[i];
    if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; }
    if (logits_out) logits_out[((size_t)b * ids_ld + t) * V + i] = v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
  }
  if (lane == 0) { red_v[wid] = bv; red_i[wid] = bi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < nw; ++i)
      if (red_v[i] > bv || (red_v[i] == bv && red_i[i] < bi)) { bv = red_v[i]; bi = red_i[i]; }
    ids[(size_t)b * ids_ld + t] = bi;
    tokens[(size_t)b * tok_ld + t + 1] = bi;
    if (t != last_step && bi == end_id && !ended[b]) {
      ended[b] = 1;
      const int n = atomicAdd(n_ended, 1) + 1;
      if (n == B) *done_step = t + 1;
    }
  }
}

// xcat[r, off : off+D] = E[targets[r]]  (beam rows carry their current token instead of a history lookup)
__global__ void lstm_embed_cur_kernel(const int* __restrict__ targets, const float* __restrict__ emb,
                                      float* __restrict__ xcat, int ld, int off, int R, int D) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int d4n = D / 4;
  if (idx >= R * d4n) return;
  const int r = idx / d4n, d = (idx % d4n) * 4;
  *reinterpret_cast<float4*>(xcat + (size_t)r * ld + off + d) =
      *reinterpret_cast<const float4*>(emb + (size_t)targets[r] * D + d);
}

// ---------------------------------------------------------------------------------------------
// Beam step of AttentionV2.forward_beam (seq2seq_v2.py:82-150), one CTA per image, batched over images (the
// reference asserts batch 1, :18-19).  Reproduces its quirks:
//   * step 0 ranks the candidates of row 0 only (:98-99); later steps rank rows < n_live flattened row-major, k = n_live;
//   * candidates are processed in top-k order; word == END -> completed list (ascending top-k position, :115), else
//     the next live set in that order;
//   * hidden (h, c) and the token history follow the PARENT, the coverage memory alpha_cum (already += alpha by the
//     attention kernel) is re-indexed by top-k POSITION (:137-147);
//   * n_complete of the LAST executed step decides the final pick (:152-168, attn_beam_finalize_kernel).
// State is permuted in place: every source row is staged in shared memory before anything is written.
// ---------------------------------------------------------------------------------------------
constexpr int ATTN_BEAM_MAX = 16;

struct AttnBeamState {
  float* scores;        // [R] cumulative log-prob of the live hypotheses
  int* targets;         // [R] current token
  int* seqs;            // [R][L] token history incl. GO at 0
  float* h; float* c;   // [R][HS]
  float* xcat; int ld, hoff;   // LSTM input rows: the h part lives at xcat[r*ld + hoff]
  float* alpha_cum;     // [R][S]
  int* n_live;          // [B]
  int* n_done;          // [B]
  int* last_complete;   // [B] completions of the last executed step
  int* finished;        // [B]
  int* done_seq;        // [R][L] completed hypotheses (GO ... END)
  int* done_len;        // [R]   length incl. GO and END
  float* done_score;    // [R]
  int* counters;        // step, -, done_step, n_finished
  int* trace;           // optional [B][max_steps][beam][2]
  float* trace_score;   // optional [B][max_steps][beam]
  int L, beam, B, V, S, HS, end_id, max_steps;
};

__global__ void __launch_bounds__(256)
attn_beam_step_kernel(const float* __restrict__ logits, AttnBeamState st) {
  extern __shared__ float s_dyn[];   // candidates [n][V]; later the staging area of the in-place permutation
  __shared__ float s_rowmax[ATTN_BEAM_MAX], s_rowlse[ATTN_BEAM_MAX];
  __shared__ float red_v[8];
  __shared__ int red_i[8];
  __shared__ float top_v[ATTN_BEAM_MAX];
  __shared__ int top_i[ATTN_BEAM_MAX];
  __shared__ int inc_pos[ATTN_BEAM_MAX], inc_parent[ATTN_BEAM_MAX], inc_word[ATTN_BEAM_MAX];
  __shared__ int cmp_parent[ATTN_BEAM_MAX];
  __shared__ float inc_score[ATTN_BEAM_MAX], cmp_score[ATTN_BEAM_MAX];
  __shared__ int s_ninc, s_ncmp;
  const int img = blockIdx.x;
  if (st.finished[img]) return;
  const int t = st.counters[0];
  const int V = st.V, beam = st.beam, L = st.L, S = st.S, HS = st.HS;
  const int n = st.n_live[img];
  const int n_eff = t == 0 ? 1 : n;   // step 0: scores[0].topk (all rows are identical)
  const int k = n;
  const int row0 = img * beam;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;

  for (int s = wid; s < n_eff; s += nw) {   // log-softmax per live row
    const float* x = logits + (size_t)(row0 + s) * V;
    float mx = -INFINITY;
    for (int i = lane; i < V; i += 32) mx = fmaxf(mx, x[i]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int i = lane; i < V; i += 32) sum += expf(x[i] - mx);
    sum = warp_sum(sum);
    if (lane == 0) { s_rowmax[s] = mx; s_rowlse[s] = logf(sum); }
  }
  __syncthreads();
  const int ncand = n_eff * V;
  for (int i = threadIdx.x; i < ncand; i += blockDim.x) {
    const int s = i / V, v = i - s * V;
    const float lp = (logits[(size_t)(row0 + s) * V + v] - s_rowmax[s]) - s_rowlse[s];
    s_dyn[i] = st.scores[row0 + s] + lp;
  }
  __syncthreads();
  for (int round = 0; round < k; ++round) {   // k rounds of block-wide argmax (value desc, index asc)
    float bv = -INFINITY; int bi = 0x7fffffff;
    for (int i = threadIdx.x; i < ncand; i += blockDim.x) {
      const float v = s_dyn[i];
      if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) { red_v[wid] = bv; red_i[wid] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int i = 1; i < nw; ++i)
        if (red_v[i] > bv || (red_v[i] == bv && red_i[i] < bi)) { bv = red_v[i]; bi = red_i[i]; }
      top_v[round] = bv; top_i[round] = bi;
      if (bi != 0x7fffffff) s_dyn[bi] = -INFINITY;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    int ninc = 0, ncmp = 0;
    for (int i = 0; i < k; ++i) {
      const int parent = top_i[i] / V, word = top_i[i] - parent * V;
      if (st.trace) {
        int* tr = st.trace + (((size_t)img * st.max_steps + t) * beam + i) * 2;
        tr[0] = parent; tr[1] = word;
        if (st.trace_score) st.trace_score[((size_t)img * st.max_steps + t) * beam + i] = top_v[i];
      }
      if (word != st.end_id) { inc_pos[ninc] = i; inc_parent[ninc] = parent; inc_word[ninc] = word; inc_score[ninc] = top_v[i]; ++ninc; }
      else { cmp_parent[ncmp] = parent; cmp_score[ncmp] = top_v[i]; ++ncmp; }
    }
    s_ninc = ninc; s_ncmp = ncmp;
  }
  __syncthreads();
  const int ninc = s_ninc, ncmp = s_ncmp;
  const int ndone0 = st.n_done[img];
  // completed hypotheses: history of the parent (t + 1 tokens incl. GO) + END
  for (int j = 0; j < ncmp; ++j) {
    const int* src = st.seqs + (size_t)(row0 + cmp_parent[j]) * L;
    int* dst = st.done_seq + (size_t)(row0 + ndone0 + j) * L;
    for (int i = threadIdx.x; i <= t; i += blockDim.x) dst[i] = src[i];
    if (threadIdx.x == 0) {
      dst[t + 1] = st.end_id;
      st.done_len[row0 + ndone0 + j] = t + 2;
      st.done_score[row0 + ndone0 + j] = cmp_score[j];
    }
  }
  // stage the sources of the in-place permutation: [n][HS] h, [n][HS] c, [k][S] alpha_cum, [n][t+1] tokens
  float* s_h = s_dyn;
  float* s_c = s_h + (size_t)n * HS;
  float* s_a = s_c + (size_t)n * HS;
  int* s_q = reinterpret_cast<int*>(s_a + (size_t)n * S);
  for (int i = threadIdx.x; i < n * HS; i += blockDim.x) {
    s_h[i] = st.h[(size_t)row0 * HS + i];
    s_c[i] = st.c[(size_t)row0 * HS + i];
  }
  for (int i = threadIdx.x; i < n * S; i += blockDim.x) s_a[i] = st.alpha_cum[(size_t)row0 * S + i];
  for (int i = threadIdx.x; i < n * (t + 1); i += blockDim.x) {
    const int r = i / (t + 1), p = i - r * (t + 1);
    s_q[i] = st.seqs[(size_t)(row0 + r) * L + p];
  }
  __syncthreads();
  for (int j = 0; j < ninc; ++j) {
    const int par = inc_parent[j], pos = inc_pos[j];
    const size_t r = (size_t)(row0 + j);
    for (int i = threadIdx.x; i < HS; i += blockDim.x) {
      const float hv = s_h[par * HS + i];
      st.h[r * HS + i] = hv;
      st.c[r * HS + i] = s_c[par * HS + i];
      st.xcat[r * st.ld + st.hoff + i] = hv;
    }
    for (int i = threadIdx.x; i < S; i += blockDim.x) st.alpha_cum[r * S + i] = s_a[pos * S + i];   // by position (Q10)
    for (int i = threadIdx.x; i <= t; i += blockDim.x) st.seqs[r * L + i] = s_q[par * (t + 1) + i];
    if (threadIdx.x == 0) {
      if (t + 1 < L) st.seqs[r * L + t + 1] = inc_word[j];
      st.scores[row0 + j] = inc_score[j];
      st.targets[row0 + j] = inc_word[j];
    }
  }
  if (threadIdx.x == 0) {
    st.n_live[img] = ninc;
    st.n_done[img] = ndone0 + ncmp;
    st.last_complete[img] = ncmp;
    if (ninc == 0) {
      st.finished[img] = 1;
      const int nf = atomicAdd(&st.counters[3], 1) + 1;
      if (nf == st.B) st.counters[2] = t + 1;
    }
  }
}

// Final pick (seq2seq_v2.py:152-174): the last executed step completed nothing -> live beam 0 (tokens after GO,
// its running score); else the first maximum of fp32 score / len(seq incl. GO and END) over the completion order,
// returned with the MAXIMUM completed score.
__global__ void attn_beam_finalize_kernel(AttnBeamState st, int steps, long long* __restrict__ best_ids, int ids_ld,
                                          int* __restrict__ best_len, float* __restrict__ best_score) {
  const int img = blockIdx.x;
  const int beam = st.beam, L = st.L;
  const int row0 = img * beam;
  __shared__ int s_best;
  __shared__ float s_score;
  if (threadIdx.x == 0) {
    int best = -1;
    float bv = 0.f, mxs = -INFINITY;
    if (st.last_complete[img] > 0) {
      const int nd = st.n_done[img];
      for (int c = 0; c < nd; ++c) {
        const float sc = st.done_score[row0 + c];
        const float v = sc / (float)st.done_len[row0 + c];
        if (best < 0 || v > bv) { best = c; bv = v; }
        mxs = fmaxf(mxs, sc);
      }
    }
    s_best = best;
    s_score = best < 0 ? st.scores[row0] : mxs;
  }
  __syncthreads();
  const int best = s_best;
  const int* src = best < 0 ? st.seqs + (size_t)row0 * L + 1 : st.done_seq + (size_t)(row0 + best) * L + 1;
  const int len = best < 0 ? steps : st.done_len[row0 + best] - 1;
  for (int i = threadIdx.x; i < ids_ld; i += blockDim.x) best_ids[(size_t)img * ids_ld + i] = i < len ? src[i] : 0;
  if (threadIdx.x == 0) { best_len[img] = len; best_score[img] = s_score; }
}

}  // namespace d2t

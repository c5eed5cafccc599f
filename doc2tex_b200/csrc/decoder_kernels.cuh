// Decode-loop kernels of the transformer head: token embedding, KV-cached self-/cross-attention,
// greedy pick, and the fused log-softmax + top-k + beam bookkeeping kernel.
//
// The reference has no KV cache (tfm.py:125-136 re-runs the whole prefix every step) and does its
// beam bookkeeping in Python on the host (tools/beam.py:68-105).  Here every decoder row keeps its
// keys/values in HBM; a beam "reorder" never moves K/V: each hypothesis carries an ancestry table
// anc[row][pos] = physical row that holds position pos of its prefix, and the reorder rewrites
// that small table only.
#pragma once
#include "common.cuh"

namespace d2t {

// bf16 hi/lo operand planes of an fp32 value (error-compensated tensor-core modes): x ~= hi + lo.
__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

// x[r,:] = E[tok[r][t]] * sqrt(D) + pe[t]    (tfm.py:92-93, position_encoding.py:24-28)
__global__ void embed_tokens_kernel(const int* __restrict__ tokens, int tok_ld, const int* __restrict__ step,
                                    long long parity_stride,  // tokens buffer = tokens + (t&1)*parity_stride (0: single buffer)
                                    const float* __restrict__ emb, const float* __restrict__ pe, float* __restrict__ x,
                                    int R, int D, float mult, __nv_bfloat16* __restrict__ x_hi,
                                    __nv_bfloat16* __restrict__ x_lo) {
  pdl_trigger();
  pdl_wait();
  const int t = *step;
  const int* tk = tokens + (parity_stride ? (long long)(t & 1) * parity_stride : 0);
  const int d4n = D / 4;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= R * d4n) return;
  const int r = idx / d4n, d = (idx % d4n) * 4;
  const int tok = tk[(size_t)r * tok_ld + t];
  const float4 e = *reinterpret_cast<const float4*>(emb + (size_t)tok * D + d);
  const float4 p = *reinterpret_cast<const float4*>(pe + (size_t)t * D + d);
  const float4 o = make_float4(e.x * mult + p.x, e.y * mult + p.y, e.z * mult + p.z, e.w * mult + p.w);
  *reinterpret_cast<float4*>(x + (size_t)r * D + d) = o;
  if (x_hi) {
    const float f[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      __nv_bfloat16 hi, lo;
      split_bf16(f[u], hi, lo);
      x_hi[(size_t)r * D + d + u] = hi;
      if (x_lo) x_lo[(size_t)r * D + d + u] = lo;
    }
  }
}

// Single-query attention for one decoder row and all heads: block = heads warps, warp h = head h.
//   out[r, h*HD + :] = softmax_j(q_h . K_j / sqrt(HD)) V_j
// Key j of row r lives at kv + src(r,j)*row_stride + j*pos_stride (K at +h*HD, V at +D+h*HD), where
//   self-attention, greedy:  src = r                      n_keys = *step + 1
//   self-attention, beam:    src = img*beam + anc[r][j]   n_keys = *step + 1   (ancestry indirection)
//   cross-attention:         src = r / rows_per_src       n_keys = n_fixed      (memory shared by the beams)
// HBM-bound single pass: every lane owns the keys j = lane (mod 32) and streams its key AND value head slices
// as two full 128-byte lines per key (all loads of a key independent -> deep memory-level parallelism), keeping an
// online-softmax state (running max, sum, 32-wide accumulator); the 32 lane states are merged once at the end
// through shared memory.  Optional bf16 hi/lo planes of the output feed the tensor-core out-projection.
template <int HD>
__global__ void __launch_bounds__(256)
decode_attention_kernel(const float* __restrict__ q, int ldq, const float* __restrict__ kv,
                        long long row_stride, int pos_stride, const int* __restrict__ anc,
                        long long anc_parity_stride, int anc_ld, int rows_per_src,
                        const int* __restrict__ step, int n_fixed, float* __restrict__ out, int D,
                        __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo) {
  pdl_trigger();
  pdl_wait();
  static_assert(HD == 32, "one lane per output channel in the merge");
  __shared__ float s_acc[8][32][HD + 1];
  const int r = blockIdx.x;
  const int h = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t = step ? *step : 0;
  const int n_keys = n_fixed > 0 ? n_fixed : t + 1;
  const int* anc_r = anc ? anc + (anc_parity_stride ? (long long)(t & 1) * anc_parity_stride : 0) + (size_t)r * anc_ld : nullptr;
  const int src_base = (r / rows_per_src) * (anc ? rows_per_src : 1);

  float qv[HD], acc[HD];
  const float scale = rsqrtf((float)HD);
#pragma unroll
  for (int d = 0; d < HD; d += 4) {
    const float4 v = *reinterpret_cast<const float4*>(q + (size_t)r * ldq + h * HD + d);
    qv[d] = v.x * scale; qv[d + 1] = v.y * scale; qv[d + 2] = v.z * scale; qv[d + 3] = v.w * scale;
    acc[d] = acc[d + 1] = acc[d + 2] = acc[d + 3] = 0.f;
  }
  float mx = -INFINITY, sum = 0.f;
  for (int j = lane; j < n_keys; j += 32) {
    const int src = anc_r ? src_base + anc_r[j] : src_base;
    const float4* kp = reinterpret_cast<const float4*>(kv + (size_t)src * row_stride + (size_t)j * pos_stride + h * HD);
    const float4* vp = reinterpret_cast<const float4*>(kv + (size_t)src * row_stride + (size_t)j * pos_stride + D + h * HD);
    float4 k4[HD / 4], v4[HD / 4];
#pragma unroll
    for (int i = 0; i < HD / 4; ++i) k4[i] = __ldg(kp + i);
#pragma unroll
    for (int i = 0; i < HD / 4; ++i) v4[i] = __ldg(vp + i);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < HD / 4; ++i) {
      s = fmaf(qv[4 * i], k4[i].x, s); s = fmaf(qv[4 * i + 1], k4[i].y, s);
      s = fmaf(qv[4 * i + 2], k4[i].z, s); s = fmaf(qv[4 * i + 3], k4[i].w, s);
    }
    const float nm = fmaxf(mx, s);
    const float corr = expf(mx - nm);   // 0 on the first key (mx = -inf)
    const float pj = expf(s - nm);
    sum = sum * corr + pj;
#pragma unroll
    for (int i = 0; i < HD / 4; ++i) {
      acc[4 * i] = fmaf(pj, v4[i].x, acc[4 * i] * corr);
      acc[4 * i + 1] = fmaf(pj, v4[i].y, acc[4 * i + 1] * corr);
      acc[4 * i + 2] = fmaf(pj, v4[i].z, acc[4 * i + 2] * corr);
      acc[4 * i + 3] = fmaf(pj, v4[i].w, acc[4 * i + 3] * corr);
    }
    mx = nm;
  }
  // merge the 32 lane states: global max, rescale, sum
  const float gm = warp_max(mx);
  const float sc = (mx == -INFINITY) ? 0.f : expf(mx - gm);
  const float total = warp_sum(sum * sc);
#pragma unroll
  for (int d = 0; d < HD; ++d) s_acc[h][lane][d] = acc[d] * sc;
  __syncwarp();
  float o = 0.f;
#pragma unroll
  for (int l = 0; l < 32; ++l) o += s_acc[h][l][lane];
  o /= total;
  out[(size_t)r * D + h * HD + lane] = o;
  if (out_hi) {
    __nv_bfloat16 hi, lo;
    split_bf16(o, hi, lo);
    out_hi[(size_t)r * D + h * HD + lane] = hi;
    if (out_lo) out_lo[(size_t)r * D + h * HD + lane] = lo;
  }
}

// Greedy pick: next = argmax(softmax(logits[r])) with the lowest index on ties (tfm.py:134-135, quirk Q7).
// Writes ids[r][t], tokens[r][t+1], optional logits copy, END flags; the block that completes the
// "every row has emitted END" condition records the number of executed steps (tfm.py:138-140).
__global__ void greedy_pick_kernel(const float* __restrict__ logits, int V, const int* __restrict__ step,
                                   int* __restrict__ tokens, int tok_ld, long long* __restrict__ ids, int ids_ld,
                                   float* __restrict__ logits_out, int* __restrict__ ended, int* __restrict__ n_ended,
                                   int* __restrict__ done_step, int R, int end_id) {
  pdl_trigger();
  pdl_wait();
  __shared__ float red_v[32];
  __shared__ int red_i[32];
  __shared__ float s_max, s_sum;
  const int r = blockIdx.x;
  const int t = *step;
  const float* x = logits + (size_t)r * V;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  float mx = -INFINITY;
  for (int i = threadIdx.x; i < V; i += blockDim.x) mx = fmaxf(mx, x[i]);
  mx = warp_max(mx);
  if (lane == 0) red_v[wid] = mx;
  __syncthreads();
  if (threadIdx.x == 0) { float m = red_v[0]; for (int i = 1; i < nw; ++i) m = fmaxf(m, red_v[i]); s_max = m; }
  __syncthreads();
  mx = s_max;
  float sum = 0.f;
  for (int i = threadIdx.x; i < V; i += blockDim.x) sum += expf(x[i] - mx);
  sum = warp_sum(sum);
  __syncthreads();
  if (lane == 0) red_v[wid] = sum;
  __syncthreads();
  if (threadIdx.x == 0) { float s = 0.f; for (int i = 0; i < nw; ++i) s += red_v[i]; s_sum = s; }
  __syncthreads();
  sum = s_sum;
  float bv = -INFINITY; int bi = 0x7fffffff;
  for (int i = threadIdx.x; i < V; i += blockDim.x) {
    const float pr = expf(x[i] - mx) / sum;
    if (pr > bv || (pr == bv && i < bi)) { bv = pr; bi = i; }
    if (logits_out) logits_out[((size_t)r * ids_ld + t) * V + i] = x[i];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
  }
  __syncthreads();
  if (lane == 0) { red_v[wid] = bv; red_i[wid] = bi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < nw; ++i)
      if (red_v[i] > bv || (red_v[i] == bv && red_i[i] < bi)) { bv = red_v[i]; bi = red_i[i]; }
    ids[(size_t)r * ids_ld + t] = bi;
    tokens[(size_t)r * tok_ld + t + 1] = bi;
    if (bi == end_id && !ended[r]) {
      ended[r] = 1;
      const int n = atomicAdd(n_ended, 1) + 1;
      if (n == R) *done_step = t + 1;
    }
  }
}

__global__ void advance_step_kernel(int* step) {
  pdl_trigger();
  pdl_wait();
  *step += 1;
}

// Decode-state initialisation (one launch per decode call).
__global__ void init_decode_state_kernel(int* tokens, long long tokens_elems, int tok_ld, int R, int beam, int go_id,
                                         int* anc, int anc_ld, float* scores, int* n_live, int* n_done, int* finished,
                                         int* ended, int* counters /* step, n_ended, done_step, n_finished */) {
  pdl_trigger();
  pdl_wait();
  const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long gsz = (long long)gridDim.x * blockDim.x;
  for (long long i = gtid; i < tokens_elems; i += gsz) tokens[i] = 0;  // PAD
  __syncthreads();
  const int nbuf = beam > 0 ? 2 : 1;
  if (anc)
    for (long long i = gtid; i < (long long)nbuf * R * anc_ld; i += gsz) anc[i] = (int)((i / anc_ld) % R) % beam;
  for (long long i = gtid; i < R; i += gsz) {
    if (ended) ended[i] = 0;
    if (scores) scores[i] = 0.f;
  }
  const int B = beam > 0 ? R / beam : R;
  for (long long i = gtid; i < B; i += gsz) {
    if (n_live) { n_live[i] = 1; n_done[i] = 0; finished[i] = 0; }
  }
  if (gtid == 0) { counters[0] = 0; counters[1] = 0; counters[2] = -1; counters[3] = 0; }
}
__global__ void set_go_tokens_kernel(int* tokens, int tok_ld, long long parity_stride, int R, int go_id) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  tokens[(size_t)r * tok_ld] = go_id;
  if (parity_stride) tokens[parity_stride + (size_t)r * tok_ld] = go_id;
}

// ---------------------------------------------------------------------------------------------
// Fused log-softmax + top-k + beam update, one CTA per image  (tfm.py:167-178 + tools/beam.py:68-105).
//   candidates = hyp_scores[s] + log_softmax(logits[s])  over live rows s, flattened row-major (beam.py:71-74)
//   k = beam - len(completed) (beam.py:70); total order: value descending, flat index ascending
//   processed in top-k order: END -> completed list (seq incl. END, fp32 score); else next live row (beam.py:86-103)
//   done when len(completed) == beam (beam.py:129-130).
// The KV "reorder" is the rewrite of the ancestry table; tokens/anc are ping-ponged by step parity.
// ---------------------------------------------------------------------------------------------
constexpr int BEAM_MAX = 16;

struct BeamState {
  int* tokens;          // [2][B*beam][L]
  int* anc;             // [2][B*beam][L]
  float* scores;        // [B*beam]
  int* n_live;          // [B]
  int* n_done;          // [B]
  int* finished;        // [B]
  int* done_seq;        // [B][beam][L]   completed hypotheses (tokens incl. END)
  int* done_len;        // [B][beam]
  float* done_score;    // [B][beam]
  int* counters;        // step, -, done_step, n_finished
  int* trace;           // optional [B][max_steps][beam][2]
  float* trace_score;   // optional [B][max_steps][beam]
  int L, beam, B, V, end_id, max_steps;
};

__global__ void __launch_bounds__(256)
beam_step_kernel(const float* __restrict__ logits, BeamState st) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float s_cand[];  // [beam][V] candidate scores
  __shared__ float s_rowmax[BEAM_MAX], s_rowlse[BEAM_MAX];
  __shared__ float red_v[8];
  __shared__ int red_i[8];
  __shared__ float top_v[BEAM_MAX];
  __shared__ int top_i[BEAM_MAX];
  __shared__ int new_parent[BEAM_MAX], new_word[BEAM_MAX];
  __shared__ float new_score[BEAM_MAX];
  __shared__ int s_nnew, s_ndone;

  const int img = blockIdx.x;
  if (st.finished[img]) return;
  const int t = st.counters[0];
  const int V = st.V, beam = st.beam, L = st.L;
  const int nlive = st.n_live[img];
  const int ndone0 = st.n_done[img];
  const int k = beam - ndone0;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const long long par = (long long)st.B * beam * L;
  const int* tok_old = st.tokens + (long long)(t & 1) * par + (size_t)img * beam * L;
  int* tok_new = st.tokens + (long long)((t + 1) & 1) * par + (size_t)img * beam * L;
  const int* anc_old = st.anc + (long long)(t & 1) * par + (size_t)img * beam * L;
  int* anc_new = st.anc + (long long)((t + 1) & 1) * par + (size_t)img * beam * L;

  // log-softmax per live row (warp per row): lp = (x - max) - log(sum exp(x - max))
  for (int s = wid; s < nlive; s += nw) {
    const float* x = logits + ((size_t)img * beam + s) * V;
    float mx = -INFINITY;
    for (int i = lane; i < V; i += 32) mx = fmaxf(mx, x[i]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int i = lane; i < V; i += 32) sum += expf(x[i] - mx);
    sum = warp_sum(sum);
    if (lane == 0) { s_rowmax[s] = mx; s_rowlse[s] = logf(sum); }
  }
  __syncthreads();
  const int ncand = nlive * V;
  for (int i = threadIdx.x; i < ncand; i += blockDim.x) {
    const int s = i / V, v = i - s * V;
    const float lp = (logits[((size_t)img * beam + s) * V + v] - s_rowmax[s]) - s_rowlse[s];
    s_cand[i] = st.scores[img * beam + s] + lp;
  }
  __syncthreads();
  // k rounds of block-wide argmax (value desc, index asc)
  for (int round = 0; round < k; ++round) {
    float bv = -INFINITY; int bi = 0x7fffffff;
    for (int i = threadIdx.x; i < ncand; i += blockDim.x) {
      const float v = s_cand[i];
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
      if (bi != 0x7fffffff) s_cand[bi] = -INFINITY;  // exclude from later rounds
    }
    __syncthreads();
  }
  // process candidates in top-k order
  if (threadIdx.x == 0) {
    int nnew = 0, ndone = ndone0;
    for (int i = 0; i < k; ++i) {
      const int idx = top_i[i];
      const int parent = idx / V, word = idx - parent * V;
      if (st.trace) {
        int* tr = st.trace + (((size_t)img * st.max_steps + t) * beam + i) * 2;
        tr[0] = parent; tr[1] = word;
        if (st.trace_score) st.trace_score[((size_t)img * st.max_steps + t) * beam + i] = top_v[i];
      }
      if (word == st.end_id) {
        new_parent[BEAM_MAX - 1 - (ndone - ndone0)] = parent;  // completed ones stacked from the top
        st.done_len[img * beam + ndone] = t + 1;
        st.done_score[img * beam + ndone] = top_v[i];
        ++ndone;
      } else {
        new_parent[nnew] = parent; new_word[nnew] = word; new_score[nnew] = top_v[i];
        ++nnew;
      }
    }
    s_nnew = nnew; s_ndone = ndone;
  }
  __syncthreads();
  const int nnew = s_nnew, ndone = s_ndone;
  // completed hypotheses: tokens[parent][1..t] + END
  for (int c = ndone0; c < ndone; ++c) {
    const int parent = new_parent[BEAM_MAX - 1 - (c - ndone0)];
    int* dst = st.done_seq + ((size_t)img * beam + c) * L;
    for (int i = threadIdx.x; i < t; i += blockDim.x) dst[i] = tok_old[(size_t)parent * L + 1 + i];
    if (threadIdx.x == 0) dst[t] = st.end_id;
  }
  // next live set: copy the parent's prefix, append the word; rewrite the ancestry table
  for (int j = 0; j < beam; ++j) {
    int* tn = tok_new + (size_t)j * L;
    int* an = anc_new + (size_t)j * L;
    if (j < nnew) {
      const int p = new_parent[j];
      for (int i = threadIdx.x; i <= t; i += blockDim.x) {
        tn[i] = tok_old[(size_t)p * L + i];
        an[i] = anc_old[(size_t)p * L + i];
      }
      if (threadIdx.x == 0) {
        if (t + 1 < L) { tn[t + 1] = new_word[j]; an[t + 1] = j; }
        st.scores[img * beam + j] = new_score[j];
      }
    } else {
      for (int i = threadIdx.x; i <= t + 1 && i < L; i += blockDim.x) { tn[i] = 0; an[i] = j; }
    }
  }
  if (threadIdx.x == 0) {
    st.n_live[img] = nnew;
    st.n_done[img] = ndone;
    if (ndone == beam) {
      st.finished[img] = 1;
      const int n = atomicAdd(&st.counters[3], 1) + 1;
      if (n == st.B) st.counters[2] = t + 1;
    }
  }
}

// Final pick (tfm.py:180-186, beam.py:132-140): nothing completed -> live hypothesis 0 (tokens[1:], score[0]);
// else first maximum of score/len in float64 over the completion order.
__global__ void beam_finalize_kernel(BeamState st, int steps, long long* __restrict__ best_ids, int ids_ld,
                                     int* __restrict__ best_len, float* __restrict__ best_score) {
  const int img = blockIdx.x;
  const int beam = st.beam, L = st.L;
  __shared__ int s_best;
  const int ndone = st.n_done[img];
  const long long par = (long long)st.B * beam * L;
  if (threadIdx.x == 0) {
    int best = -1;
    double bv = 0.0;
    for (int c = 0; c < ndone; ++c) {
      const int len = st.done_len[img * beam + c];
      const double v = (double)st.done_score[img * beam + c] / (double)(len > 0 ? len : 1);
      if (best < 0 || v > bv) { best = c; bv = v; }
    }
    s_best = best;
  }
  __syncthreads();
  const int best = s_best;
  const int* src;
  int len;
  float score;
  if (best < 0) {
    src = st.tokens + (long long)(steps & 1) * par + (size_t)img * beam * L + 1;  // slot 0, drop GO
    len = L - 1;
    score = st.scores[img * beam];
  } else {
    src = st.done_seq + ((size_t)img * beam + best) * L;
    len = st.done_len[img * beam + best];
    score = st.done_score[img * beam + best];
  }
  for (int i = threadIdx.x; i < ids_ld; i += blockDim.x) best_ids[(size_t)img * ids_ld + i] = i < len ? src[i] : 0;
  if (threadIdx.x == 0) { best_len[img] = len; best_score[img] = score; }
}

}  // namespace d2t

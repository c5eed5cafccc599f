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
template <int HS>  // hidden size (256)
__global__ void __launch_bounds__(256)
lstm_attention_step_kernel(const float* __restrict__ keyproj, const float* __restrict__ ctx, int ntok,
                           const float* __restrict__ qp, const float* __restrict__ locM /*[HS][taps]*/,
                           const float* __restrict__ locc /*[HS]*/, int taps, const float* __restrict__ score_w,
                           const float* __restrict__ score_b, float* __restrict__ alpha_cum /*[B][S]*/,
                           float* __restrict__ xcat, int ld) {
  extern __shared__ float sm[];  // [S] alpha_cum copy, [S] scores
  const int S = ntok - 1;
  float* s_ac = sm;
  float* s_e = sm + S;
  __shared__ float red[8];
  __shared__ float s_bc[2];
  const int b = blockIdx.x;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  constexpr int PER = HS / 32;
  for (int i = threadIdx.x; i < S; i += blockDim.x) s_ac[i] = alpha_cum[(size_t)b * S + i];
  float q[PER], sw[PER], cc[PER];
#pragma unroll
  for (int k = 0; k < PER; ++k) {
    const int h = lane + 32 * k;
    q[k] = qp[(size_t)b * HS + h] + locc[h];
    sw[k] = score_w[h];
  }
  __syncthreads();
  const int pad = taps / 2;
  const float sb = score_b[0];
  for (int s = wid; s < S; s += nw) {
    const float* kr = keyproj + ((size_t)b * ntok + 1 + s) * HS;
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < PER; ++k) {
      const int h = lane + 32 * k;
      float loc = 0.f;
      for (int j = 0; j < taps; ++j) {
        const int ss = s + j - pad;
        const float a = (ss >= 0 && ss < S) ? s_ac[ss] : 0.f;
        loc = fmaf(locM[h * taps + j], a, loc);
      }
      cc[k] = loc;
      acc = fmaf(sw[k], tanhf(kr[h] + q[k] + cc[k]), acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) s_e[s] = acc + sb;
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
    alpha_cum[(size_t)b * S + i] = s_ac[i] + a;  // coverage update AFTER the step (seq2seq_v2.py:264-266)
  }
  __syncthreads();
  // context = alpha^T H, one thread per channel (HS == input channels == 256 here)
  for (int d = threadIdx.x; d < HS; d += blockDim.x) {
    const float* hp = ctx + ((size_t)b * ntok + 1) * HS + d;
    float a0 = 0.f, a1 = 0.f;
    int s = 0;
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

}  // namespace d2t

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
  pdl_wait();      // PDL: everything above overlapped the predecessor
  pdl_trigger();   // allow exactly one successor to pre-launch (chain depth 1: pre-launched CTAs hold SM resources)
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
// Memory-bound single pass.  A quarter warp (8 lanes x float4) covers one 128-byte key (and value) head slice, so
// every load instruction moves four complete lines; the four quarter warps walk keys j = g (mod 4) with an
// online-softmax state each (running max, sum, 4 output channels per lane), two keys in flight per quarter warp,
// and are merged with shuffles at the end.  Optional bf16 hi/lo planes of the output feed the tensor-core
// out-projection.
// SPLIT > 1: the keys of a (row, head) are dealt to SPLIT warps of the block (8-key groups, round robin) and their
// online-softmax states are merged through shared memory.  At small batch the kernel is bound by the loads each SM
// has in flight, not by bandwidth: twice the warps per row = twice the bytes in flight.
// four consecutive K / V channels of one lane: fp32 cache (16-byte load) or bf16 cache (8-byte load, single-pass bf16 mode)
__device__ __forceinline__ float4 kv_load4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 kv_load4(const __nv_bfloat16* p) {
  const uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
  const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&u.x), b = *reinterpret_cast<const __nv_bfloat162*>(&u.y);
  const float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
  return make_float4(fa.x, fa.y, fb.x, fb.y);
}

// 2^x in one MUFU.EX2 (the walk keeps its scores in log2 units: q is pre-scaled by log2(e) / sqrt(HD))
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// The key walk of one (row, head) by one warp: returns the warp-merged online-softmax state (every lane holds gm / sum,
// lane c holds its 4 output channels summed over the four quarter warps).  A quarter warp takes KPI keys per iteration —
// key j = g + 4 i + 4 KPI (sp + SPLIT it) — and issues all 2 KPI 16-byte loads before it touches any of them: the walk is
// bound by memory latency x iterations, so the loads in flight per warp set its speed (KPI = 2: 13 dependent round trips
// at step 100, KPI = 4: 7).  The pointer array / K loads / V loads are three separate unrolled loops ON PURPOSE: written
// any other way tried (one loop; a warp-uniform trip count with a branch or with selects around the update) nvcc sinks the
// value loads below the score shuffles, every iteration pays two dependent round trips and the kernel is 20 % slower
// (ncu at 5 120 rows: 136.9 vs 115.8 us) although it executes 14 % fewer instructions.
template <int SPLIT, int KPI, typename KV>
__device__ __forceinline__ void attention_walk(const float4 q4, const KV* __restrict__ kbase, long long row_stride,
                                               int pos_stride, const int* anc_r, int src_base, int n_keys,
                                               int g, int sp, int D, float& gm_out, float& sum_out, float4& acc_out) {
  float mx = -INFINITY, sum = 0.f;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  const unsigned gmask = 0xFFu << (g * 8);   // quarter warps run different trip counts: group-local shuffles
  for (int j0 = g + 4 * KPI * sp; j0 < n_keys; j0 += 4 * KPI * SPLIT) {
    const KV* p[KPI];
    float4 k[KPI], v[KPI];
    float d[KPI];
#pragma unroll
    for (int i = 0; i < KPI; ++i) {
      const int j = (j0 + 4 * i < n_keys) ? j0 + 4 * i : j0;
      const int src = anc_r ? src_base + anc_r[j] : src_base;
      p[i] = kbase + (size_t)src * row_stride + (size_t)j * pos_stride;
    }
#pragma unroll
    for (int i = 0; i < KPI; ++i) k[i] = kv_load4(p[i]);
#pragma unroll
    for (int i = 0; i < KPI; ++i) v[i] = kv_load4(p[i] + D);
#pragma unroll
    for (int i = 0; i < KPI; ++i) d[i] = fmaf(q4.x, k[i].x, fmaf(q4.y, k[i].y, fmaf(q4.z, k[i].z, q4.w * k[i].w)));
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
#pragma unroll
      for (int i = 0; i < KPI; ++i) d[i] += __shfl_xor_sync(gmask, d[i], o);
    }
    float nm = mx;
#pragma unroll
    for (int i = 0; i < KPI; ++i) {
      if (j0 + 4 * i >= n_keys) d[i] = -INFINITY;
      nm = fmaxf(nm, d[i]);
    }
    const float corr = fast_exp2(mx - nm);   // 0 on the first iteration (mx = -inf)
    sum *= corr; acc.x *= corr; acc.y *= corr; acc.z *= corr; acc.w *= corr;
#pragma unroll
    for (int i = 0; i < KPI; ++i) {
      const float e = fast_exp2(d[i] - nm);
      sum += e;
      acc.x = fmaf(e, v[i].x, acc.x); acc.y = fmaf(e, v[i].y, acc.y); acc.z = fmaf(e, v[i].z, acc.z); acc.w = fmaf(e, v[i].w, acc.w);
    }
    mx = nm;
  }
  // merge the four quarter-warp states
  __syncwarp();
  float gm = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 8));
  gm = fmaxf(gm, __shfl_xor_sync(0xffffffffu, gm, 16));
  const float sc = (mx == -INFINITY) ? 0.f : fast_exp2(mx - gm);
  sum *= sc; acc.x *= sc; acc.y *= sc; acc.z *= sc; acc.w *= sc;
#pragma unroll
  for (int o = 8; o < 32; o <<= 1) {
    sum += __shfl_xor_sync(0xffffffffu, sum, o);
    acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o);
    acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
    acc.z += __shfl_xor_sync(0xffffffffu, acc.z, o);
    acc.w += __shfl_xor_sync(0xffffffffu, acc.w, o);
  }
  gm_out = gm; sum_out = sum; acc_out = acc;
}

// split 0 folds the states the other key splits left in shared memory ([SPLIT - 1][max, sum, pad, pad, acc[32]]) into its own
template <int SPLIT>
__device__ __forceinline__ void attention_merge_splits(const float* __restrict__ parts, int c, float gm, float& sum, float4& acc) {
  float M = gm;
#pragma unroll
  for (int q = 0; q < SPLIT - 1; ++q) M = fmaxf(M, parts[q * 36]);
  const float w0 = (gm == -INFINITY) ? 0.f : fast_exp2(gm - M);   // the walk's maxima are in log2 units
  sum *= w0; acc.x *= w0; acc.y *= w0; acc.z *= w0; acc.w *= w0;
#pragma unroll
  for (int q = 0; q < SPLIT - 1; ++q) {
    const float* pp = parts + q * 36;
    const float wq = (pp[0] == -INFINITY) ? 0.f : fast_exp2(pp[0] - M);
    const float4 a = *reinterpret_cast<const float4*>(pp + 4 + c);
    sum = fmaf(wq, pp[1], sum);
    acc.x = fmaf(wq, a.x, acc.x); acc.y = fmaf(wq, a.y, acc.y); acc.z = fmaf(wq, a.z, acc.z); acc.w = fmaf(wq, a.w, acc.w);
  }
}

__device__ __forceinline__ void attention_store(const float4 acc, float sum, size_t off, float* __restrict__ out,
                                                __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo) {
  const float inv = 1.0f / sum;
  const float4 o4 = make_float4(acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv);
  *reinterpret_cast<float4*>(out + off) = o4;
  if (out_hi) {
    const float f[4] = {o4.x, o4.y, o4.z, o4.w};
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      __nv_bfloat16 hi, lo;
      split_bf16(f[u], hi, lo);
      out_hi[off + u] = hi;
      if (out_lo) out_lo[off + u] = lo;
    }
  }
}

// HB > 1: the heads of a row are spread over HB blocks (grid = rows x HB, 8 / HB heads each): 512 quarter-size blocks
// balance over 148 SMs better than 256 full ones.
// Five resident blocks per SM (48 registers, no spills): the walk waits on its loads (ncu at 5 120 beam rows: 41-57 % issue,
// 30 % L2, long-scoreboard stalls, 47 % of the warp slots with 4 blocks of 60 registers), so warps in flight are what it
// needs — beam-5 1 048 -> 964 us per step at 5 120 rows, greedy 797 -> 775 at 2 560; six blocks (40 registers) spill and lose.
template <int HD, int SPLIT = 1, int HB = 1, typename KV = float, int KPI = 2>
__global__ void __launch_bounds__(256 * SPLIT / HB, 5)
decode_attention_kernel(const float* __restrict__ q, int ldq, const KV* __restrict__ kv,
                        long long row_stride, int pos_stride, const int* __restrict__ anc,
                        long long anc_parity_stride, int anc_ld, int rows_per_src,
                        const int* __restrict__ step, int n_fixed, float* __restrict__ out, int D,
                        __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo) {
  static_assert(HD == 32, "8 lanes x float4 per head slice");
  pdl_wait();      // PDL: everything above overlapped the predecessor
  pdl_trigger();   // allow exactly one successor to pre-launch (chain depth 1: pre-launched CTAs hold SM resources)
  const int r = HB == 1 ? blockIdx.x : blockIdx.x / HB;
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nheads = blockDim.x / (32 * SPLIT);                       // heads of this block
  const int hl = SPLIT == 1 ? wid : wid % nheads, sp = SPLIT == 1 ? 0 : wid / nheads;
  const int h = HB == 1 ? hl : (blockIdx.x % HB) * nheads + hl;
  const int g = lane >> 3, c = (lane & 7) * 4;
  const int t = step ? *step : 0;
  const int n_keys = n_fixed > 0 ? n_fixed : t + 1;
  const int* anc_r = anc ? anc + (anc_parity_stride ? (long long)(t & 1) * anc_parity_stride : 0) + (size_t)r * anc_ld : nullptr;
  const int src_base = (r / rows_per_src) * (anc ? rows_per_src : 1);
  const float scale = rsqrtf((float)HD) * 1.4426950408889634f;   // scores in log2 units: softmax = 2^(s - max) / sum
  float4 q4 = *reinterpret_cast<const float4*>(q + (size_t)r * ldq + h * HD + c);
  q4.x *= scale; q4.y *= scale; q4.z *= scale; q4.w *= scale;
  float gm, sum;
  float4 acc;
  attention_walk<SPLIT, KPI>(q4, kv + h * HD + c, row_stride, pos_stride, anc_r, src_base, n_keys, g, sp, D, gm, sum, acc);
  if constexpr (SPLIT > 1) {
    __shared__ float s_part[8 * (SPLIT - 1) * 36];   // [head][split - 1][max, sum, pad, pad, acc[32]]
    if (sp > 0) {
      float* pp = s_part + (hl * (SPLIT - 1) + sp - 1) * 36;
      if (g == 0) *reinterpret_cast<float4*>(pp + 4 + c) = acc;
      if (lane == 0) { pp[0] = gm; pp[1] = sum; }
    }
    __syncthreads();
    if (sp > 0) return;
    attention_merge_splits<SPLIT>(s_part + hl * (SPLIT - 1) * 36, c, gm, sum, acc);
  }
  if (g == 0) attention_store(acc, sum, (size_t)r * D + h * HD + c, out, out_hi, out_lo);
}

__global__ void f32_to_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = __float2bfloat16_rn(in[i]);
}

// Greedy pick: next = argmax(softmax(logits[r])) with the lowest index on ties (tfm.py:134-135, quirk Q7).
// Writes ids[r][t], tokens[r][t+1], optional logits copy, END flags; the block that completes the
// "every row has emitted END" condition records the number of executed steps (tfm.py:138-140).
__global__ void greedy_pick_kernel(const float* __restrict__ logits, int V, const int* __restrict__ step,
                                   int* __restrict__ tokens, int tok_ld, long long* __restrict__ ids, int ids_ld,
                                   float* __restrict__ logits_out, int* __restrict__ ended, int* __restrict__ n_ended,
                                   int* __restrict__ done_step, int R, int end_id,
                                   // optional fusion of the NEXT step's token embedding and of the step counter (emb != nullptr):
                                   // x[r] = E[picked] * mult + pe[t + 1] (+ bf16 planes); the last block to finish advances *step_rw
                                   const float* __restrict__ emb = nullptr, const float* __restrict__ pe = nullptr,
                                   float* __restrict__ xn = nullptr, __nv_bfloat16* __restrict__ xn_hi = nullptr,
                                   __nv_bfloat16* __restrict__ xn_lo = nullptr, int D = 0, float mult = 0.f,
                                   int* __restrict__ step_rw = nullptr, int* __restrict__ ticket = nullptr) {
  pdl_wait();      // PDL: everything above overlapped the predecessor
  pdl_trigger();   // allow exactly one successor to pre-launch (chain depth 1: pre-launched CTAs hold SM resources)
  __shared__ float red_v[32];
  __shared__ int red_i[32];
  __shared__ float s_max, s_sum;
  __shared__ int s_tok;
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
  // fed-back token: argmax of the softmax PROBABILITIES, lowest index on ties (tfm.py:134-135, quirk Q7);
  // returned id: argmax of the raw LOGITS (tfm.py:142, preds_index = out.max(2)) — the two differ only where two distinct
  // logits round to the same fp32 probability.
  float bv = -INFINITY; int bi = 0x7fffffff;
  float lv = -INFINITY; int li = 0x7fffffff;
  for (int i = threadIdx.x; i < V; i += blockDim.x) {
    const float xi = x[i];
    const float pr = expf(xi - mx) / sum;
    if (pr > bv || (pr == bv && i < bi)) { bv = pr; bi = i; }
    if (xi > lv || (xi == lv && i < li)) { lv = xi; li = i; }
    if (logits_out) logits_out[((size_t)r * ids_ld + t) * V + i] = xi;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    const float pv = __shfl_xor_sync(0xffffffffu, lv, o);
    const int pi = __shfl_xor_sync(0xffffffffu, li, o);
    if (pv > lv || (pv == lv && pi < li)) { lv = pv; li = pi; }
  }
  __syncthreads();
  if (lane == 0) { red_v[wid] = bv; red_i[wid] = bi; red_v[16 + wid] = lv; red_i[16 + wid] = li; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < nw; ++i) {
      if (red_v[i] > bv || (red_v[i] == bv && red_i[i] < bi)) { bv = red_v[i]; bi = red_i[i]; }
      if (red_v[16 + i] > lv || (red_v[16 + i] == lv && red_i[16 + i] < li)) { lv = red_v[16 + i]; li = red_i[16 + i]; }
    }
    ids[(size_t)r * ids_ld + t] = li;
    tokens[(size_t)r * tok_ld + t + 1] = bi;
    if (bi == end_id && !ended[r]) {
      ended[r] = 1;
      const int n = atomicAdd(n_ended, 1) + 1;
      if (n == R) *done_step = t + 1;
    }
    s_tok = bi;
  }
  if (emb == nullptr) return;
  __syncthreads();
  const int tok = s_tok;
  for (int i = threadIdx.x; i < D / 4; i += blockDim.x) {
    const float4 e4 = *reinterpret_cast<const float4*>(emb + (size_t)tok * D + 4 * i);
    const float4 p4 = *reinterpret_cast<const float4*>(pe + (size_t)(t + 1) * D + 4 * i);
    const float4 o = make_float4(e4.x * mult + p4.x, e4.y * mult + p4.y, e4.z * mult + p4.z, e4.w * mult + p4.w);
    *reinterpret_cast<float4*>(xn + (size_t)r * D + 4 * i) = o;
    if (xn_hi) {
      const float f[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        __nv_bfloat16 hi, lo;
        split_bf16(f[u], hi, lo);
        xn_hi[(size_t)r * D + 4 * i + u] = hi;
        if (xn_lo) xn_lo[(size_t)r * D + 4 * i + u] = lo;
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {   // every block has read *step before it takes a ticket; the last one advances the counter
    __threadfence();
    if (atomicAdd(ticket, 1) == (int)gridDim.x - 1) {
      *ticket = 0;
      *step_rw = t + 1;
    }
  }
}

__global__ void advance_step_kernel(int* step) {
  pdl_wait();      // PDL: everything above overlapped the predecessor
  pdl_trigger();   // allow exactly one successor to pre-launch (chain depth 1: pre-launched CTAs hold SM resources)
  *step += 1;
}

// Decode-state initialisation (one launch per decode call).
__global__ void init_decode_state_kernel(int* tokens, long long tokens_elems, int tok_ld, int R, int beam, int go_id,
                                         int* anc, int anc_ld, float* scores, int* n_live, int* n_done, int* finished,
                                         int* ended, int* counters /* step, n_ended, done_step, n_finished */) {
  pdl_wait();      // PDL: everything above overlapped the predecessor
  pdl_trigger();   // allow exactly one successor to pre-launch (chain depth 1: pre-launched CTAs hold SM resources)
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
constexpr int BEAM_CPL = 16;   // beam_step fast path: candidates per lane held in registers (vocabulary <= 512)

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
  float* runner_up;     // optional [B][max_steps]: score of the best candidate NOT selected at that step (near-tie audit)
  int L, beam, B, V, end_id, max_steps;
};

__global__ void __launch_bounds__(256)
beam_step_kernel(const float* __restrict__ logits, BeamState st) {
  pdl_wait();      // PDL: everything above overlapped the predecessor
  pdl_trigger();   // allow exactly one successor to pre-launch (chain depth 1: pre-launched CTAs hold SM resources)
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

  const int rounds = k + (st.runner_up != nullptr ? 1 : 0);   // one more when the near-tie audit wants the runner-up's score
  if (V <= 32 * BEAM_CPL) {
    // ---- fast path: a warp keeps the candidates of one live row in registers (BEAM_CPL per lane), takes the row's own
    // top `rounds` with warp-level argmax rounds (no block barrier), and warp 0 merges the nlive x rounds survivors.
    // Total order everywhere: value descending, flat index (row * V + word) ascending. ----
    __shared__ float s_lv[BEAM_MAX * (BEAM_MAX + 1)];
    __shared__ int s_li[BEAM_MAX * (BEAM_MAX + 1)];
    for (int s = wid; s < nlive; s += nw) {
      const float* x = logits + ((size_t)img * beam + s) * V;
      float c[BEAM_CPL];
      float mxr = -INFINITY;
#pragma unroll
      for (int i = 0; i < BEAM_CPL; ++i) {
        const int v = lane + 32 * i;
        c[i] = v < V ? x[v] : -INFINITY;
        mxr = fmaxf(mxr, c[i]);
      }
      mxr = warp_max(mxr);
      float sm = 0.f;
#pragma unroll
      for (int i = 0; i < BEAM_CPL; ++i) sm += (lane + 32 * i < V) ? expf(c[i] - mxr) : 0.f;
      sm = warp_sum(sm);
      const float lse = logf(sm), base = st.scores[img * beam + s];
#pragma unroll
      for (int i = 0; i < BEAM_CPL; ++i) c[i] = (lane + 32 * i < V) ? base + ((c[i] - mxr) - lse) : -INFINITY;
      for (int round = 0; round < rounds; ++round) {
        float bv = -INFINITY; int bi = 0x7fffffff;
#pragma unroll
        for (int i = 0; i < BEAM_CPL; ++i) {            // ascending word index: strict > keeps the lowest index on ties
          if (c[i] > bv) { bv = c[i]; bi = lane + 32 * i; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
          const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
          if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if (lane == 0) { s_lv[s * (BEAM_MAX + 1) + round] = bv; s_li[s * (BEAM_MAX + 1) + round] = bi == 0x7fffffff ? bi : s * V + bi; }
        if (bi != 0x7fffffff && (bi & 31) == lane) {     // the owner drops the winner
#pragma unroll
          for (int i = 0; i < BEAM_CPL; ++i) if (i == (bi >> 5)) c[i] = -INFINITY;
        }
      }
    }
    __syncthreads();
    if (wid == 0) {
      const int ncand = nlive * rounds;                   // <= 16 * 17 survivors
      float c[(BEAM_MAX * (BEAM_MAX + 1) + 31) / 32];
      int ci[(BEAM_MAX * (BEAM_MAX + 1) + 31) / 32];
#pragma unroll
      for (int i = 0; i < (BEAM_MAX * (BEAM_MAX + 1) + 31) / 32; ++i) {
        const int e = lane + 32 * i;
        const bool ok = e < ncand;
        const int s = ok ? e / rounds : 0, r_ = ok ? e - s * rounds : 0;
        c[i] = ok ? s_lv[s * (BEAM_MAX + 1) + r_] : -INFINITY;
        ci[i] = ok ? s_li[s * (BEAM_MAX + 1) + r_] : 0x7fffffff;
      }
      for (int round = 0; round < rounds; ++round) {
        float bv = -INFINITY; int bi = 0x7fffffff, bslot = -1;
#pragma unroll
        for (int i = 0; i < (BEAM_MAX * (BEAM_MAX + 1) + 31) / 32; ++i) {
          if (c[i] > bv || (c[i] == bv && ci[i] < bi)) { bv = c[i]; bi = ci[i]; bslot = i; }
        }
        float wv = bv; int wi = bi;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float ov = __shfl_xor_sync(0xffffffffu, wv, o);
          const int oi = __shfl_xor_sync(0xffffffffu, wi, o);
          if (ov > wv || (ov == wv && oi < wi)) { wv = ov; wi = oi; }
        }
        if (lane == 0) {
          if (round < k) { top_v[round] = wv; top_i[round] = wi; }
          else st.runner_up[(size_t)img * st.max_steps + t] = wv;
        }
        if (bslot >= 0 && wi == bi && wv == bv && wi != 0x7fffffff) {   // flat indices are unique: exactly one lane owns it
#pragma unroll
          for (int i = 0; i < (BEAM_MAX * (BEAM_MAX + 1) + 31) / 32; ++i) if (i == bslot) { c[i] = -INFINITY; ci[i] = 0x7fffffff; }
        }
      }
    }
    __syncthreads();
  } else {
  // ---- generic path (large vocabularies): candidates in shared memory, block-wide argmax rounds ----
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
  // `rounds` rounds of block-wide argmax (value desc, index asc)
  for (int round = 0; round < rounds; ++round) {
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
      if (round < k) { top_v[round] = bv; top_i[round] = bi; }
      else st.runner_up[(size_t)img * st.max_steps + t] = bv;
      if (bi != 0x7fffffff) s_cand[bi] = -INFINITY;  // exclude from later rounds
    }
    __syncthreads();
  }
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

// Memory-bound encoder kernels: first conv (Cin=1), max-pools, token assembly, LayerNorm,
// and the flash-style encoder self-attention.  All activations are NHWC / row-major fp32.
#pragma once
#include "common.cuh"

namespace d2t {

// bf16 hi/lo planes of four consecutive fp32 values (A operand of the next tensor-core convolution)
__device__ __forceinline__ void store_planes4(float4 v, __nv_bfloat16* hi, __nv_bfloat16* lo, size_t off) {
  const float f[4] = {v.x, v.y, v.z, v.w};
  uint32_t hw[2], lw[2];
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const __nv_bfloat16 h0 = __float2bfloat16_rn(f[2 * u]), h1 = __float2bfloat16_rn(f[2 * u + 1]);
    hw[u] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
    const __nv_bfloat16 l0 = __float2bfloat16_rn(f[2 * u] - __bfloat162float(h0));
    const __nv_bfloat16 l1 = __float2bfloat16_rn(f[2 * u + 1] - __bfloat162float(h1));
    lw[u] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
  }
  *reinterpret_cast<uint2*>(hi + off) = make_uint2(hw[0], hw[1]);
  if (lo) *reinterpret_cast<uint2*>(lo + off) = make_uint2(lw[0], lw[1]);
}

// conv0_1: Conv2d(1 -> Cout, 3x3, s1, p1, bias=False) + BN(eval) + ReLU  (resnet.py:206-208).
// K = 9 is not tensor-core work: direct conv, output-bandwidth bound.  x: [B,H,W] (NCHW with C=1),
// out: NHWC [B,H,W,Cout].  One thread = one pixel x 4 output channels (float4 store, coalesced).
__global__ void conv0_direct_kernel(const float* __restrict__ x, const float* __restrict__ w /*[Cout][9]*/,
                                    const float* __restrict__ scale, const float* __restrict__ shift,
                                    float* __restrict__ out, int B, int H, int W, int Cout,
                                    __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo) {
  extern __shared__ float sm[];  // [9][Cout] weights, [Cout] scale, [Cout] shift
  float* sw = sm;
  float* ssc = sm + 9 * Cout;
  float* ssh = ssc + Cout;
  for (int i = threadIdx.x; i < 9 * Cout; i += blockDim.x) {
    const int c = i % Cout, t = i / Cout;
    sw[i] = w[c * 9 + t];
  }
  for (int i = threadIdx.x; i < Cout; i += blockDim.x) { ssc[i] = scale[i]; ssh[i] = shift[i]; }
  __syncthreads();
  const int cg = Cout / 4;
  const long long total = (long long)B * H * W * cg;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c4 = (int)(idx % cg) * 4;
    const long long pix = idx / cg;
    const int ow = (int)(pix % W);
    const int oh = (int)((pix / W) % H);
    const int b = (int)(pix / ((long long)W * H));
    const float* xb = x + (size_t)b * H * W;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int ih = oh + kh - 1;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int iw = ow + kw - 1;
        const float v = ((unsigned)ih < (unsigned)H && (unsigned)iw < (unsigned)W) ? __ldg(xb + (size_t)ih * W + iw) : 0.f;
        const float* wr = sw + (kh * 3 + kw) * Cout + c4;
        a0 = fmaf(v, wr[0], a0); a1 = fmaf(v, wr[1], a1); a2 = fmaf(v, wr[2], a2); a3 = fmaf(v, wr[3], a3);
      }
    }
    float4 o;
    o.x = fmaxf(a0 * ssc[c4 + 0] + ssh[c4 + 0], 0.f);
    o.y = fmaxf(a1 * ssc[c4 + 1] + ssh[c4 + 1], 0.f);
    o.z = fmaxf(a2 * ssc[c4 + 2] + ssh[c4 + 2], 0.f);
    o.w = fmaxf(a3 * ssc[c4 + 3] + ssh[c4 + 3], 0.f);
    if (out) *reinterpret_cast<float4*>(out + pix * Cout + c4) = o;   // fp32 copy only when a consumer reads it
    if (out_hi) store_planes4(o, out_hi, out_lo, (size_t)(pix * Cout + c4));
  }
}

// Same layer, register-blocked: one thread = FOUR consecutive pixels of a row x EIGHT output channels.  The thread's 72 filter
// weights and its BN scale / shift stay in registers over the grid-stride loop (the channel group of a thread never changes:
// the stride is a multiple of Cout / 8), the 3 x 6 input window is loaded once for the four pixels, and every pixel's eight
// channels leave as one 16-byte store per plane.  Each output keeps the (kh, kw) accumulation order of the kernel above, so the
// results are bit-identical; the first version spent its time on index arithmetic and 8-byte stores (ncu: 0.42 ms for 537 MB
// of planes = 1.28 TB/s).  Needs Cout % 8 == 0 and W % 4 == 0.
__global__ void __launch_bounds__(256)
conv0_direct4x8_kernel(const float* __restrict__ x, const float* __restrict__ w /*[Cout][9]*/,
                       const float* __restrict__ scale, const float* __restrict__ shift,
                       float* __restrict__ out, int B, int H, int W, int Cout,
                       __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo) {
  const int cg = Cout / 8, wq = W / 4;
  const long long total = (long long)B * H * wq * cg;
  const long long stride = (long long)gridDim.x * blockDim.x;   // host: a multiple of cg
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int c8 = (int)(idx % cg) * 8;
  float wr[9][8], sc[8], sh[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
#pragma unroll
    for (int t = 0; t < 9; ++t) wr[t][c] = __ldg(w + (c8 + c) * 9 + t);
    sc[c] = __ldg(scale + c8 + c);
    sh[c] = __ldg(shift + c8 + c);
  }
  for (; idx < total; idx += stride) {
    const long long pg = idx / cg;
    const int ow0 = (int)(pg % wq) * 4;
    const int oh = (int)((pg / wq) % H);
    const int b = (int)(pg / ((long long)wq * H));
    const float* xb = x + (size_t)b * H * W;
    float v[3][6];
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int ih = oh + kh - 1;
      const bool rok = (unsigned)ih < (unsigned)H;
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        const int iw = ow0 + i - 1;
        v[kh][i] = (rok && (unsigned)iw < (unsigned)W) ? __ldg(xb + (size_t)ih * W + iw) : 0.f;
      }
    }
#pragma unroll
    for (int px = 0; px < 4; ++px) {
      float a[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) a[c] = 0.f;
#pragma unroll
      for (int kh = 0; kh < 3; ++kh)
#pragma unroll
        for (int kw = 0; kw < 3; ++kw)
#pragma unroll
          for (int c = 0; c < 8; ++c) a[c] = fmaf(v[kh][px + kw], wr[kh * 3 + kw][c], a[c]);
      float o[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) o[c] = fmaxf(a[c] * sc[c] + sh[c], 0.f);
      const size_t off = ((size_t)((size_t)b * H + oh) * W + ow0 + px) * Cout + c8;
      if (out) {
        *reinterpret_cast<float4*>(out + off) = make_float4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<float4*>(out + off + 4) = make_float4(o[4], o[5], o[6], o[7]);
      }
      if (out_hi) {
        uint32_t hw[4], lw[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const __nv_bfloat16 h0 = __float2bfloat16_rn(o[2 * u]), h1 = __float2bfloat16_rn(o[2 * u + 1]);
          hw[u] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
          const __nv_bfloat16 l0 = __float2bfloat16_rn(o[2 * u] - __bfloat162float(h0));
          const __nv_bfloat16 l1 = __float2bfloat16_rn(o[2 * u + 1] - __bfloat162float(h1));
          lw[u] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
        }
        *reinterpret_cast<uint4*>(out_hi + off) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
        if (out_lo) *reinterpret_cast<uint4*>(out_lo + off) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
      }
    }
  }
}

// MaxPool2d(kernel 2x2, stride (SH,SW), padding (PH,PW)) on NHWC; padding behaves as -inf
// (resnet.py:97,107,120: maxpool3 is k2 s(2,1) p(0,1)).
__global__ void maxpool2x2_nhwc_kernel(const float* __restrict__ x, float* __restrict__ out, int B, int H, int W,
                                       int C, int OH, int OW, int SH, int SW, int PH, int PW,
                                       __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo) {
  const int c4n = C / 4;
  const long long total = (long long)B * OH * OW * c4n;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c4 = (int)(idx % c4n) * 4;
    long long pix = idx / c4n;
    const int ow = (int)(pix % OW);
    const int oh = (int)((pix / OW) % OH);
    const int b = (int)(pix / ((long long)OW * OH));
    float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
#pragma unroll
    for (int kh = 0; kh < 2; ++kh) {
      const int ih = oh * SH - PH + kh;
      if ((unsigned)ih >= (unsigned)H) continue;
#pragma unroll
      for (int kw = 0; kw < 2; ++kw) {
        const int iw = ow * SW - PW + kw;
        if ((unsigned)iw >= (unsigned)W) continue;
        const float4 v = __ldg(reinterpret_cast<const float4*>(x + (((size_t)b * H + ih) * W + iw) * C + c4));
        m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
      }
    }
    *reinterpret_cast<float4*>(out + pix * C + c4) = m;
    if (out_hi) store_planes4(m, out_hi, out_lo, (size_t)(pix * C + c4));
  }
}

// x[b,0,:] = cls + pos[0];  x[b,1+p,:] = tok[b,p,:] + pos[1+p]   (vit_encoder.py:255-260; pos is the
// PREFIX slice of the max-grid table, quirk Q3).
__global__ void assemble_tokens_kernel(const float* __restrict__ tok, const float* __restrict__ cls,
                                       const float* __restrict__ pos, float* __restrict__ x, int B, int N, int D) {
  const int d4n = D / 4;
  const long long total = (long long)B * (N + 1) * d4n;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int d = (int)(idx % d4n) * 4;
    const long long r = idx / d4n;
    const int t = (int)(r % (N + 1));
    const int b = (int)(r / (N + 1));
    const float4 a = (t == 0) ? *reinterpret_cast<const float4*>(cls + d)
                              : *reinterpret_cast<const float4*>(tok + ((size_t)b * N + (t - 1)) * D + d);
    const float4 pe = *reinterpret_cast<const float4*>(pos + (size_t)t * D + d);
    *reinterpret_cast<float4*>(x + r * D + d) = make_float4(a.x + pe.x, a.y + pe.y, a.z + pe.z, a.w + pe.w);
  }
}

// ViTEncoder.interpolating_pos_embedding (vit_encoder.py:58-95): out[0] = pos[0] (cls), out[1 + y*gw + x] = bicubic
// resample of the [EH, EW, D] patch table at output cell (y, x).  torch upsample_bicubic2d, align_corners=False with an
// explicit scale_factor: src = (dst + 0.5) / scale - 0.5 (not clamped), A = -0.75, taps clamped to the border.
__device__ __forceinline__ void bicubic_coeffs(float t, float (&c)[4]) {
  const float A = -0.75f;
  const float x0 = t + 1.0f, x1 = t, x2 = 1.0f - t, x3 = 2.0f - t;
  c[0] = ((A * x0 - 5.0f * A) * x0 + 8.0f * A) * x0 - 4.0f * A;
  c[1] = ((A + 2.0f) * x1 - (A + 3.0f)) * x1 * x1 + 1.0f;
  c[2] = ((A + 2.0f) * x2 - (A + 3.0f)) * x2 * x2 + 1.0f;
  c[3] = ((A * x3 - 5.0f * A) * x3 + 8.0f * A) * x3 - 4.0f * A;
}
__global__ void pos_embed_bicubic_kernel(const float* __restrict__ pos, int EH, int EW, float* __restrict__ out, int gh,
                                         int gw, float inv_scale_h, float inv_scale_w, int D) {
  const long long total = (long long)(1 + gh * gw) * D;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int d = (int)(idx % D);
    const int t = (int)(idx / D);
    if (t == 0) { out[idx] = pos[d]; continue; }
    const int oy = (t - 1) / gw, ox = (t - 1) % gw;
    const float sy = inv_scale_h * (oy + 0.5f) - 0.5f, sx = inv_scale_w * (ox + 0.5f) - 0.5f;
    const int iy = (int)floorf(sy), ix = (int)floorf(sx);
    float cy[4], cx[4];
    bicubic_coeffs(sy - iy, cy);
    bicubic_coeffs(sx - ix, cx);
    float acc = 0.f;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int yy = min(max(iy - 1 + a, 0), EH - 1);
      float row = 0.f;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int xx = min(max(ix - 1 + b, 0), EW - 1);
        row += cx[b] * pos[(size_t)(1 + yy * EW + xx) * D + d];
      }
      acc += cy[a] * row;
    }
    out[idx] = acc;
  }
}

// LayerNorm over the last dim, one warp per row, values kept in registers (D = 128 * NV4).
// Two-pass mean / biased variance in fp32, y = (x - mean) / sqrt(var + eps) * w + b (torch semantics).
template <int NV4>
__global__ void layernorm_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                                 float* __restrict__ y, int rows, float eps,
                                 __nv_bfloat16* __restrict__ y_hi = nullptr, __nv_bfloat16* __restrict__ y_lo = nullptr,
                                 int nparts = 1, long long part_stride = 0) {
  pdl_wait();      // PDL: everything above overlapped the predecessor
  pdl_trigger();   // allow exactly one successor to pre-launch (chain depth 1: pre-launched CTAs hold SM resources)
  constexpr int D = 128 * NV4;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const float* xr = x + (size_t)warp * D;
  float4 v[NV4];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV4; ++i) {
    v[i] = *reinterpret_cast<const float4*>(xr + (i * 32 + lane) * 4);
    for (int pt = 1; pt < nparts; ++pt) {   // split-K partial sums of the producing GEMM, added in slice order
      const float4 u = *reinterpret_cast<const float4*>(xr + (size_t)pt * part_stride + (i * 32 + lane) * 4);
      v[i].x += u.x; v[i].y += u.y; v[i].z += u.z; v[i].w += u.w;
    }
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  const float mean = warp_sum(s) * (1.0f / D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV4; ++i) {
    const float a = v[i].x - mean, c = v[i].y - mean, d = v[i].z - mean, e = v[i].w - mean;
    q += (a * a + c * c) + (d * d + e * e);
  }
  const float rstd = 1.0f / sqrtf(warp_sum(q) * (1.0f / D) + eps);
  float* yr = y + (size_t)warp * D;
#pragma unroll
  for (int i = 0; i < NV4; ++i) {
    const int o = (i * 32 + lane) * 4;
    const float4 ww = *reinterpret_cast<const float4*>(w + o);
    const float4 bb = *reinterpret_cast<const float4*>(b + o);
    float4 r;
    r.x = (v[i].x - mean) * rstd * ww.x + bb.x;
    r.y = (v[i].y - mean) * rstd * ww.y + bb.y;
    r.z = (v[i].z - mean) * rstd * ww.z + bb.z;
    r.w = (v[i].w - mean) * rstd * ww.w + bb.w;
    if (y) *reinterpret_cast<float4*>(yr + o) = r;   // fp32 copy only when a consumer reads it
    if (y_hi) {  // bf16 hi/lo operand planes for the next tensor-core GEMM
      const float f[4] = {r.x, r.y, r.z, r.w};
      uint32_t hw[2], lw[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const __nv_bfloat16 h0 = __float2bfloat16_rn(f[2 * u]), h1 = __float2bfloat16_rn(f[2 * u + 1]);
        hw[u] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
        const __nv_bfloat16 l0 = __float2bfloat16_rn(f[2 * u] - __bfloat162float(h0));
        const __nv_bfloat16 l1 = __float2bfloat16_rn(f[2 * u + 1] - __bfloat162float(h1));
        lw[u] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
      }
      *reinterpret_cast<uint2*>(y_hi + (size_t)warp * D + o) = make_uint2(hw[0], hw[1]);
      if (y_lo) *reinterpret_cast<uint2*>(y_lo + (size_t)warp * D + o) = make_uint2(lw[0], lw[1]);
    }
  }
}

// Encoder self-attention, flash style: softmax(q k^T * scale) v for one (image, head), no mask
// (vision_transformer.py:61-81; zero-padded patch tokens attend and are attended, quirk Q4).
// qkv: [B*N, 3*D] rows = tokens, q | k | v column blocks, head h at columns h*HD.  out: [B*N, D].
// FOUR threads own one query: each keeps q and a partial output accumulator in registers and walks every fourth key of the
// K/V tile in shared memory (rows padded to 36 floats: the four rows a quarter warp reads lie in disjoint banks), with its
// own running max / sum; the four partial softmaxes are merged by two shuffle rounds at the end.  The 65-token sequence of a
// 64x256 image is 260 threads (9 warps, 90 % of the lanes busy, 17 keys each) where one thread per query left 65 of 128
// lanes busy for 65 serial keys (ncu: 99 us per launch at 8 % of the DRAM roofline, latency-bound).
constexpr int ENC_ATT_QT = 72;              // queries per block
constexpr int ENC_ATT_KT = 72;              // keys per shared-memory tile
constexpr int ENC_ATT_PITCH = 36;           // floats per staged K / V row
template <int HD>
__global__ void __launch_bounds__(4 * ENC_ATT_QT, 2)
encoder_attention_kernel(const float* __restrict__ qkv, float* __restrict__ out, int N, int D, float scale,
                         __nv_bfloat16* __restrict__ out_hi = nullptr, __nv_bfloat16* __restrict__ out_lo = nullptr) {
  static_assert(HD == 32, "head dim 32: each thread of a quad writes 8 output columns");
  constexpr int KT = ENC_ATT_KT, KPT = KT / 4, P = ENC_ATT_PITCH;
  __shared__ __align__(16) float Ks[KT * P];
  __shared__ __align__(16) float Vs[KT * P];
  const int heads = D / HD;
  const int b = blockIdx.y / heads, h = blockIdx.y % heads;
  const int part = threadIdx.x & 3;
  const int qi = blockIdx.x * ENC_ATT_QT + (threadIdx.x >> 2);
  const bool qok = qi < N;
  const size_t ld = 3 * (size_t)D;
  const float* base = qkv + (size_t)b * N * ld;
  float q[HD], acc[HD];
#pragma unroll
  for (int d = 0; d < HD; d += 4) {
    const float4 v = qok ? *reinterpret_cast<const float4*>(base + (size_t)qi * ld + h * HD + d) : make_float4(0, 0, 0, 0);
    q[d] = v.x; q[d + 1] = v.y; q[d + 2] = v.z; q[d + 3] = v.w;
    acc[d] = acc[d + 1] = acc[d + 2] = acc[d + 3] = 0.f;
  }
  float mx = -INFINITY, l = 0.f;
  for (int k0 = 0; k0 < N; k0 += KT) {
    const int kn = min(KT, N - k0);
    __syncthreads();
    for (int i = threadIdx.x; i < kn * (HD / 4); i += blockDim.x) {
      const int r = i / (HD / 4), c = (i % (HD / 4)) * 4;
      const float* src = base + (size_t)(k0 + r) * ld + h * HD + c;
      *reinterpret_cast<float4*>(&Ks[r * P + c]) = *reinterpret_cast<const float4*>(src + D);
      *reinterpret_cast<float4*>(&Vs[r * P + c]) = *reinterpret_cast<const float4*>(src + 2 * D);
    }
    __syncthreads();
    float sc[KPT];
    float cm = -INFINITY;
#pragma unroll
    for (int u = 0; u < KPT; ++u) {
      const int j = part + 4 * u;
      float d0 = -INFINITY;
      if (j < kn) {
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
        for (int d = 0; d < HD; d += 4) {
          const float4 kv = *reinterpret_cast<const float4*>(&Ks[j * P + d]);
          a0 = fmaf(q[d], kv.x, a0); a1 = fmaf(q[d + 1], kv.y, a1); a2 = fmaf(q[d + 2], kv.z, a2); a3 = fmaf(q[d + 3], kv.w, a3);
        }
        d0 = ((a0 + a1) + (a2 + a3)) * scale;
      }
      sc[u] = d0;
      cm = fmaxf(cm, d0);
    }
    const float nm = fmaxf(mx, cm);
    if (nm > -INFINITY) {   // a thread whose key slice of this tile is empty keeps its state
      const float corr = expf(mx - nm);   // exp(-inf) = 0 on the first tile
      l *= corr;
#pragma unroll
      for (int d = 0; d < HD; ++d) acc[d] *= corr;
#pragma unroll
      for (int u = 0; u < KPT; ++u) {
        const int j = part + 4 * u;
        if (j < kn) {
          const float pj = expf(sc[u] - nm);
          l += pj;
#pragma unroll
          for (int d = 0; d < HD; d += 4) {
            const float4 vv = *reinterpret_cast<const float4*>(&Vs[j * P + d]);
            acc[d] = fmaf(pj, vv.x, acc[d]); acc[d + 1] = fmaf(pj, vv.y, acc[d + 1]);
            acc[d + 2] = fmaf(pj, vv.z, acc[d + 2]); acc[d + 3] = fmaf(pj, vv.w, acc[d + 3]);
          }
        }
      }
      mx = nm;
    }
  }
  // merge the four partial softmaxes of a query (lanes 4i .. 4i+3)
  float ma = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
  ma = fmaxf(ma, __shfl_xor_sync(0xffffffffu, ma, 2));
  const float w = mx > -INFINITY ? expf(mx - ma) : 0.f;
  l *= w;
  l += __shfl_xor_sync(0xffffffffu, l, 1);
  l += __shfl_xor_sync(0xffffffffu, l, 2);
  const float inv = 1.0f / l;
  float mine[8];
#pragma unroll
  for (int d = 0; d < HD; ++d) {
    float v = acc[d] * w;
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    if ((d >> 3) == part) mine[d & 7] = v * inv;   // thread `part` of the quad writes columns [8 part, 8 part + 8)
  }
  if (qok) {
    const size_t o0 = ((size_t)b * N + qi) * D + h * HD + 8 * part;
#pragma unroll
    for (int d = 0; d < 8; d += 4) {
      const float f[4] = {mine[d], mine[d + 1], mine[d + 2], mine[d + 3]};
      if (out) *reinterpret_cast<float4*>(out + o0 + d) = make_float4(f[0], f[1], f[2], f[3]);
      if (out_hi) {   // bf16 hi/lo operand planes for the projection that follows (tensor-core precisions)
        uint32_t hw[2], lw[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const __nv_bfloat16 h0 = __float2bfloat16_rn(f[2 * u]), h1 = __float2bfloat16_rn(f[2 * u + 1]);
          hw[u] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
          const __nv_bfloat16 l0 = __float2bfloat16_rn(f[2 * u] - __bfloat162float(h0));
          const __nv_bfloat16 l1 = __float2bfloat16_rn(f[2 * u + 1] - __bfloat162float(h1));
          lw[u] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
        }
        *reinterpret_cast<uint2*>(out_hi + o0 + d) = make_uint2(hw[0], hw[1]);
        if (out_lo) *reinterpret_cast<uint2*>(out_lo + o0 + d) = make_uint2(lw[0], lw[1]);
      }
    }
  }
}

// NHWC [B,H,W,C] -> NCHW [B,C,H,W] (debug taps only).
__global__ void nhwc_to_nchw_kernel(const float* __restrict__ x, float* __restrict__ y, int B, int H, int W, int C) {
  const long long total = (long long)B * H * W * C;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int w = (int)(idx % W);
    const int h = (int)((idx / W) % H);
    const int c = (int)((idx / ((long long)W * H)) % C);
    const int b = (int)(idx / ((long long)W * H * C));
    y[idx] = x[(((size_t)b * H + h) * W + w) * C + c];
  }
}

}  // namespace d2t

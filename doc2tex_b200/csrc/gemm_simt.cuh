// fp32 implicit-GEMM convolution / linear kernel on the FFMA pipe.
//
// This is the parity anchor of the engine (D2T_PREC_FP32): every product and sum is IEEE fp32,
// the K loop is sequential per output element, so results are deterministic and differ from the
// reference's oneDNN/MKL kernels only by summation order.  It serves every contraction of the
// path: the 3x3 / 2x2 / 1x1 convolutions of the ResNet stem (resnet.py:205-245) with the folded
// BatchNorm + residual + ReLU epilogue, the patch-embed conv (patchembed.py:135) and all Linear
// layers (vision_transformer.py:63-79,27-31; nn.TransformerDecoderLayer; tfm.py:133).
#pragma once
#include "common.cuh"

namespace d2t {

template <int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
conv_gemm_simt_kernel(const ConvGemm p) {
  pdl_wait();      // PDL: everything above overlapped the predecessor
  pdl_trigger();   // allow exactly one successor to pre-launch (chain depth 1: pre-launched CTAs hold SM resources)
  constexpr int BK = 16;
  constexpr int NT = (BM / TM) * (BN / TN);
  constexpr int A_F4 = BM * BK / 4;
  constexpr int B_F4 = BN * BK / 4;
  constexpr int A_PER = (A_F4 + NT - 1) / NT;
  constexpr int B_PER = (B_F4 + NT - 1) / NT;
  static_assert(TM == 4 || TM == 8, "TM");
  static_assert(TN == 4 || TN == 8, "TN");

  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Bs[2][BK][BN + 4];

  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;

  // ---- per-thread gather bookkeeping for the A (activation) tile ----
  const float* a_base[A_PER];
  int a_ih0[A_PER], a_iw0[A_PER];
  bool a_ok[A_PER];
#pragma unroll
  for (int i = 0; i < A_PER; ++i) {
    const int f = tid + i * NT;
    const int row = f >> 2;
    const int m = m0 + row;
    a_ok[i] = (f < A_F4) && (m < p.M);
    const int mm = a_ok[i] ? m : 0;
    const int ow = mm % p.OW;
    const int t = mm / p.OW;
    const int oh = t % p.OH;
    const int b = t / p.OH;
    a_ih0[i] = oh * p.SH - p.PH;
    a_iw0[i] = ow * p.SW - p.PW;
    a_base[i] = p.x + (size_t)b * p.H * p.W * p.C + ((f & 3) << 2);
  }
  const float* b_ptr[B_PER];
  bool b_ok[B_PER];
#pragma unroll
  for (int i = 0; i < B_PER; ++i) {
    const int f = tid + i * NT;
    const int n = n0 + (f >> 2);
    b_ok[i] = (f < B_F4) && (n < p.N);
    b_ptr[i] = p.w + (size_t)(b_ok[i] ? n : 0) * p.K + ((f & 3) << 2);
  }

  float4 a_reg[A_PER], b_reg[B_PER];
  auto load_tiles = [&](int kb) {
    const int k0 = kb * BK;
    const int tap = k0 / p.C;
    const int c0 = k0 - tap * p.C;
    const int kh = tap / p.KW;
    const int kw = tap - kh * p.KW;
#pragma unroll
    for (int i = 0; i < A_PER; ++i) {
      const int ih = a_ih0[i] + kh, iw = a_iw0[i] + kw;
      const bool ok = a_ok[i] && (unsigned)ih < (unsigned)p.H && (unsigned)iw < (unsigned)p.W;
      a_reg[i] = ok ? __ldg(reinterpret_cast<const float4*>(a_base[i] + ((size_t)ih * p.W + iw) * p.C + c0))
                    : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int i = 0; i < B_PER; ++i)
      b_reg[i] = b_ok[i] ? __ldg(reinterpret_cast<const float4*>(b_ptr[i] + k0)) : make_float4(0.f, 0.f, 0.f, 0.f);
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int i = 0; i < A_PER; ++i) {
      const int f = tid + i * NT;
      if (f < A_F4) {
        const int row = f >> 2, k = (f & 3) << 2;
        As[buf][k + 0][row] = a_reg[i].x; As[buf][k + 1][row] = a_reg[i].y;
        As[buf][k + 2][row] = a_reg[i].z; As[buf][k + 3][row] = a_reg[i].w;
      }
    }
#pragma unroll
    for (int i = 0; i < B_PER; ++i) {
      const int f = tid + i * NT;
      if (f < B_F4) {
        const int row = f >> 2, k = (f & 3) << 2;
        Bs[buf][k + 0][row] = b_reg[i].x; Bs[buf][k + 1][row] = b_reg[i].y;
        Bs[buf][k + 2][row] = b_reg[i].z; Bs[buf][k + 3][row] = b_reg[i].w;
      }
    }
  };

  const int tx = tid % (BN / TN);
  const int ty = tid / (BN / TN);
  // TM/TN == 8: the 8 rows/cols are two groups of 4 half a tile apart (conflict-free float4 smem reads)
  auto row_of = [&](int i) { return (TM == 8 && i >= 4) ? BM / 2 + ty * 4 + (i - 4) : ty * 4 + i; };
  auto col_of = [&](int j) { return (TN == 8 && j >= 4) ? BN / 2 + tx * 4 + (j - 4) : tx * 4 + j; };

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  const int nkb = p.K / BK;
  load_tiles(0);
  store_tiles(0);
  __syncthreads();
  for (int kb = 0; kb < nkb; ++kb) {
    const int buf = kb & 1;
    if (kb + 1 < nkb) load_tiles(kb + 1);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[TM], b[TN];
#pragma unroll
      for (int g = 0; g < TM / 4; ++g) {
        const float4 v = *reinterpret_cast<const float4*>(&As[buf][k][row_of(g * 4)]);
        a[g * 4 + 0] = v.x; a[g * 4 + 1] = v.y; a[g * 4 + 2] = v.z; a[g * 4 + 3] = v.w;
      }
#pragma unroll
      for (int g = 0; g < TN / 4; ++g) {
        const float4 v = *reinterpret_cast<const float4*>(&Bs[buf][k][col_of(g * 4)]);
        b[g * 4 + 0] = v.x; b[g * 4 + 1] = v.y; b[g * 4 + 2] = v.z; b[g * 4 + 3] = v.w;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kb + 1 < nkb) store_tiles(buf ^ 1);
    __syncthreads();
  }

  // ---- epilogue: folded BN / bias, residual, activation ----
  float* out2 = p.out2;
  if (out2 != nullptr && p.dyn != nullptr) out2 += (long long)(*p.dyn) * p.dyn_mul2;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + row_of(i);
    if (m >= p.M) continue;
#pragma unroll
    for (int g = 0; g < TN / 4; ++g) {
      const int n = n0 + col_of(g * 4);
      if (n >= p.N) continue;
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int nn = n + j;
        float r = acc[i][g * 4 + j];
        if (nn < p.N) {
          if (p.scale) r *= __ldg(p.scale + nn);
          if (p.shift) r += __ldg(p.shift + nn);
          if (p.res) r += __ldg(p.res + (size_t)m * p.ldr + nn);
          r = apply_act(r, p.act);
        }
        v[j] = r;
      }
      float* dst;
      if (p.out2 != nullptr && n >= p.n_split) dst = out2 + (size_t)m * p.ldc2 + (n - p.n_split);
      else dst = p.out + (size_t)m * p.ldc + n;
      if (n + 3 < p.N && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
        *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (n + j < p.N) dst[j] = v[j];
      }
    }
  }
}

// Host-side launch: big tiles for the stem / encoder, small tiles when the grid would not fill the SMs.
inline cudaError_t launch_conv_gemm_simt(const ConvGemm& p, cudaStream_t s, int num_sms) {
  const long long big_ctas = (long long)((p.M + 127) / 128) * ((p.N + 127) / 128);
  if (big_ctas >= num_sms && p.N >= 96) {
    dim3 grid((p.M + 127) / 128, (p.N + 127) / 128);
    return launch_kernel(conv_gemm_simt_kernel<128, 128, 8, 8>, grid, dim3(256), 0, s, p);
  } else if ((long long)((p.M + 63) / 64) * ((p.N + 63) / 64) >= num_sms) {
    dim3 grid((p.M + 63) / 64, (p.N + 63) / 64);
    return launch_kernel(conv_gemm_simt_kernel<64, 64, 4, 4>, grid, dim3(256), 0, s, p);
  } else {
    dim3 grid((p.M + 31) / 32, (p.N + 31) / 32);
    return launch_kernel(conv_gemm_simt_kernel<32, 32, 4, 4>, grid, dim3(64), 0, s, p);
  }
  return cudaGetLastError();
}

}  // namespace d2t

// Shared definitions for the doc2tex_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <math.h>

namespace d2t {

enum Act { ACT_NONE = 0, ACT_RELU = 1, ACT_GELU = 2 };

// One dense contraction in GEMM view: out[m, n] = act((sum_k A[m,k] W[n,k]) * scale[n] + shift[n] + res[m,n]).
// A is gathered on the fly from an NHWC activation tensor (implicit GEMM): row m is the output pixel
// (b, oh, ow), column k = (kh*KW + kw)*C + ci.  A plain linear layer is the KH=KW=1, H=W=1 case.
struct ConvGemm {
  const float* x;      // NHWC input [B,H,W,C]
  const float* w;      // [N][KH*KW*C]
  const float* scale;  // [N] or nullptr (1)
  const float* shift;  // [N] or nullptr (0)
  const float* res;    // [M, ldr] or nullptr
  float* out;          // [M, ldc]  columns [0, n_split)
  float* out2;         // [M, ldc2] columns [n_split, N) (stored at col - n_split); nullptr = unused
  const int* dyn;      // optional device scalar: out2 += (*dyn) * dyn_mul2 elements (KV-cache slot of the step)
  long long dyn_mul2;
  int ldc, ldc2, ldr, n_split;
  int B, H, W, C, KH, KW, SH, SW, PH, PW, OH, OW;
  int M, N, K;
  int act;
};

__device__ __forceinline__ float gelu_erf(float x) {
  // nn.GELU() default = exact erf form (vision_transformer.py:15)
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == ACT_RELU) return fmaxf(v, 0.0f);
  if (act == ACT_GELU) return gelu_erf(v);
  return v;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace d2t

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
  // tensor-core extras (ignored by the FFMA kernel)
  __nv_bfloat16* out_hi;   // optional bf16 hi/lo planes of `out` (A operand of the next tensor-core GEMM)
  __nv_bfloat16* out_lo;
  const void* a_map_hi;    // optional host pointers to CUtensorMap of pre-split A planes [M, K] (TMA-fed A operand)
  const void* a_map_lo;
  long long* dbg;          // optional: CTA 0 writes phase timestamps (globaltimer ns) for latency debugging
  // optional split-K (TMA-fed-A tensor-core path): the K range is cut into k_splits slices, slice s is computed by its own
  // CTA and stored (without activation; bias / residual only in slice 0) at out + s * split_stride; the consumer (the
  // LayerNorm kernel) adds the slices.  A tcgen05.mma retires every ~90 ns whatever its N, so a K=1024 bf16x3 projection
  // is 192 serial MMAs = 17 us on one CTA; eight slices take 2 us each.
  int out2_bf16;           // out2 (the KV-cache slot) holds bf16 elements (single-pass bf16 mode): same element offsets
  int k_splits;            // 0 / 1 = no split
  int stack;               // bf16x3, TMA-fed-A path: 1 = two MMAs per k-step against the stacked [W_hi ; W_lo] operand (see gemm_tc.cuh)
  long long split_stride;  // elements between partial outputs
  // fused 2x2 / stride-2 max-pool (conv_gemm_tc3_kernel only): row m of the GEMM addresses the pixels in WINDOW-major order
  // (m = 4 * window + 2 * dy + dx), the epilogue reduces the four rows of a window after BN + ReLU and writes the pooled
  // NHWC tensor [B, OH/2, OW/2, N] (planes and / or fp32).  OH and OW must be even.
  int pool;
  // optional bf16 hi/lo NHWC planes of the INPUT activation (cp.async-fed A operand of conv_gemm_tc3_kernel)
  const __nv_bfloat16* x_hi;
  const __nv_bfloat16* x_lo;
  int single_cta;          // 1 = never the CTA-pair kernel (short-K problems whose tile count quantises badly over 74 pairs)
};

// Programmatic dependent launch (PDL): a kernel launched with the programmatic-serialization attribute may start
// while its predecessor in the stream is still running; it must execute pdl_wait() before touching anything the
// predecessor wrote.  pdl_trigger() lets the successor's CTAs be scheduled early (its prologue then overlaps our
// execution).  Both are no-ops for ordinary launches.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

inline bool& pdl_enabled() {
  static thread_local bool on = false;
  return on;
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute attr[1];
  int n = 0;
  if (pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

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

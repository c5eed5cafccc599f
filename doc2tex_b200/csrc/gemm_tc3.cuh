// Stem-convolution variant of the tcgen05 implicit-GEMM contraction with an ASYNCHRONOUS A operand.
//
// conv_gemm_tc_kernel gathers fp32 activations through registers (LDG -> split -> STS), one k-block ahead; on the
// large 3x3 convolutions that path, not the tensor pipe, sets the pace (measured: 1.67 us per k-block against a
// 0.81 us MMA floor, 0.34 us in single-pass bf16).  Here the activations already exist in HBM as bf16 hi/lo NHWC
// planes — written by the epilogue of the producing layer — and the eight producer warps only issue 16-byte
// cp.async (LDGSTS, zero-fill at borders / tails) straight into the swizzled operand stage and arm the stage's
// mbarrier with cp.async.mbarrier.arrive: no registers, no conversion, as many k-blocks in flight as there are
// stages.  Stages are half as deep (64-byte rows, 64B swizzle, 32 bf16 per k-block) so that twice as many fit:
// 4 x 48 KB for the 3-pass 256-wide tile.
#pragma once
#include "gemm_tc.cuh"

namespace d2t {
namespace tc {

__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
// TMA im2col load (NHWC tensor map built by cuTensorMapEncodeIm2col): `pixelsPerColumn` consecutive output pixels, starting
// at the filter's base position (w, h) of image n and walking the bounding box row-major, x `channelsPerPixel` channels from
// channel c, for the filter tap (kw, kh); out-of-image taps are zero-filled — the padding of the convolution.
__device__ __forceinline__ void tma_load_im2col_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c, int w, int h, int n,
                                                   uint16_t kw, uint16_t kh) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(n), "h"(kw), "h"(kh) : "memory");
}
// K-major operand tile with 64-byte rows and 64B swizzle: 8-row atoms of 512 B
__device__ __forceinline__ uint64_t make_smem_desc_sw64(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;   // SWIZZLE_64B
  return d;
}

}  // namespace tc

// MT = 128-row m-tiles per CTA.  MT = 2 keeps TWO accumulators (2 x 256 TMEM columns) and runs both against the same
// weight stage, halving the weight bytes each SM has to pull from L2 per FLOP (the per-SM L2->SM fill rate, ~83 GB/s,
// is what bounds these kernels); the price is a non-overlapped epilogue (no spare TMEM for double buffering).
template <int PASSES, int BN, int MT = 1>
struct Tc3Cfg {
  static constexpr int PLANES = PASSES == 1 ? 1 : 2;
  static constexpr int KB_ELEMS = 32, CH_ELEMS = 8;        // 64-byte operand rows = 4 chunks of 16 B
  static constexpr int A_HALF_BYTES = TC_BM * 64;          // one m-tile, one plane
  static constexpr int A_BYTES = MT * A_HALF_BYTES;
  static constexpr int B_BYTES = BN * 64;
  static constexpr int STAGE_BYTES = PLANES * (A_BYTES + B_BYTES);
  // epilogue warps: the 64- / 128-wide tiles of the early layers are bound by their epilogue traffic (ncu), so they get
  // two warps per TMEM lane quadrant; the 256-wide tiles hide the epilogue behind a 58 us main loop and keep the 4th stage
  static constexpr int EPI_WARPS = (BN <= 128 && MT == 1) ? 8 : 4;
  static constexpr int THREADS = (EPI_WARPS + 8 + 2) * 32;
  static constexpr int EPI_STAGE_BYTES = EPI_WARPS * 32 * TC_EPI_PITCH * 4;
  static constexpr int STAGES_RAW = (225 * 1024 - EPI_STAGE_BYTES - 1280) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int TMEM_COLS = 2 * BN;                 // MT=1: two buffers of one tile; MT=2: one buffer of two tiles
  static constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_BYTES + EPI_STAGE_BYTES + 1024 + 256;
  static_assert(STAGES >= 2, "need at least a double buffer");
  static_assert(MT == 1 || MT == 2, "MT");
};

// A_TMA: the activation operand arrives by TMA im2col loads of the bf16 NHWC planes (one instruction per 128-pixel tile,
// plane and k-block, issued by the TMA thread next to the weight tile; hardware zero-fill = the convolution's padding)
// instead of the producer warps' cp.async gather — both operands are then "fed by TMA" and the eight producer warps idle.
// Stride-1 convolutions with symmetric padding and C % 32 == 0 in row-major pixel order (not the fused-pool order).
template <int PASSES, int BN, int MT, bool A_TMA = false>
__global__ void __launch_bounds__(Tc3Cfg<PASSES, BN, MT>::THREADS, 1)
conv_gemm_tc3_kernel(const ConvGemm p, const __grid_constant__ CUtensorMap map_hi,
                     const __grid_constant__ CUtensorMap map_lo, int tiles_m, int tiles_n,
                     const __grid_constant__ CUtensorMap amap_hi, const __grid_constant__ CUtensorMap amap_lo) {
  using Cfg = Tc3Cfg<PASSES, BN, MT>;
  constexpr int ROWS = MT * TC_BM;                          // output rows per CTA tile
  constexpr int STAGES = Cfg::STAGES, PLANES = Cfg::PLANES;
  constexpr int EPI_WARPS = Cfg::EPI_WARPS, TMA_WARP = EPI_WARPS + 8, MMA_WARP = EPI_WARPS + 9;
  constexpr int NSLOT = EPI_WARPS / 4;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - tc::smem_u32(smem_raw));
  const uint32_t bars = smem_base + STAGES * Cfg::STAGE_BYTES + Cfg::EPI_STAGE_BYTES;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bars + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bars + 8u * (2 * STAGES + 2 + a); };
  const uint32_t tmem_slot = bars + 8u * (2 * STAGES + 4);
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + STAGES * Cfg::STAGE_BYTES + Cfg::EPI_STAGE_BYTES + 8 * (2 * STAGES + 4));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = tiles_m * tiles_n;
  const int nkb = (p.K + Cfg::KB_ELEMS - 1) / Cfg::KB_ELEMS;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      tc::mbar_init(full_bar(s), A_TMA ? 1 : TC_PROD_THREADS + 1);   // 256 cp.async completion arrivals + the TMA thread's expect_tx
      tc::mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      tc::mbar_init(tfull_bar(a), 1);
      tc::mbar_init(tempty_bar(a), EPI_WARPS * 32);
    }
    tc::fence_barrier_init();
  }
  if (warp == MMA_WARP) tc::tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  if (warp == TMA_WARP && lane == 0) {
    tc::tma_prefetch_desc(&map_hi);
    if (PLANES == 2) tc::tma_prefetch_desc(&map_lo);
    if (A_TMA) {
      tc::tma_prefetch_desc(&amap_hi);
      if (PLANES == 2) tc::tma_prefetch_desc(&amap_lo);
    }
  }
  tc::tcgen05_before_sync();
  __syncthreads();
  tc::tcgen05_after_sync();
  const uint32_t tmem_base = *tmem_slot_gen;
  pdl_wait();
  pdl_trigger();

  if (warp < EPI_WARPS) {
    // =========================== epilogue ===========================
    const int quad = warp & 3, slot = warp >> 2;   // two warps of a quadrant alternate over the 32-column chunks
    float* const stg = reinterpret_cast<float*>(smem_gen + (size_t)STAGES * Cfg::STAGE_BYTES) + warp * (32 * TC_EPI_PITCH);
    const int sub_r = lane >> 3, c4 = (lane & 7) * 4;
    const int M = p.M, N = p.N, ldc = p.ldc, ldr = p.ldr, act = p.act & 15;
    const float* const scale = p.scale;
    const float* const shift = p.shift;
    const float* const res = p.res;
    float* const out = p.out;
    __nv_bfloat16* const out_hi = p.out_hi;
    __nv_bfloat16* const out_lo = p.out_lo;
    // columns >= n_split go to a second tensor (the KV-cache slot of the current decode step; whole 32-column chunks)
    const int n_split = p.n_split, ldc2 = p.ldc2;
    float* out2 = nullptr;
    if (p.out2) out2 = p.out2 + (p.dyn ? (long long)(*p.dyn) * p.dyn_mul2 : 0) - n_split;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int tm = tile / tiles_n, tn = tile - tm * tiles_n;
      const int acc = MT == 1 ? (it & 1) : 0;
      tc::mbar_wait(tfull_bar(acc), MT == 1 ? ((it >> 1) & 1) : (it & 1));
      tc::tcgen05_after_sync();
#pragma unroll 1
      for (int hj = slot; hj < MT * (BN / 32); hj += NSLOT) {
        const int h = hj / (BN / 32), j = hj - h * (BN / 32);   // m-tile of the CTA tile, 32-column chunk
        const int m_first = tm * ROWS + h * TC_BM + quad * 32 + sub_r;
        uint32_t r[32];
        tc::tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)((MT == 1 ? acc : h) * BN + j * 32), r);
        tc::tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 8; ++q)
          *reinterpret_cast<uint4*>(stg + lane * TC_EPI_PITCH + q * 4) = make_uint4(r[q * 4], r[q * 4 + 1], r[q * 4 + 2], r[q * 4 + 3]);
        __syncwarp();
        const int n = tn * BN + j * 32 + c4;
        if (p.pool) {
          // rows 4 w .. 4 w + 3 of this 32-row chunk are the four pixels of pooling window w: BN scale / shift (the scale may
          // be negative, so the affine comes first) + ReLU on each, then the maximum; lane (sub_r, c4) takes windows sub_r, sub_r + 4
          if (n < N) {
            float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
            if (scale) sc = __ldg(reinterpret_cast<const float4*>(scale + n));
            if (shift) sh = __ldg(reinterpret_cast<const float4*>(shift + n));
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              const int w = sub_r + 4 * i;
              const int m0 = tm * ROWS + h * TC_BM + quad * 32 + 4 * w;     // first GEMM row of the window
              if (m0 < M) {
                float4 best = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                  float4 v = *reinterpret_cast<const float4*>(stg + (4 * w + u) * TC_EPI_PITCH + c4);
                  v.x = fmaf(v.x, sc.x, sh.x); v.y = fmaf(v.y, sc.y, sh.y); v.z = fmaf(v.z, sc.z, sh.z); v.w = fmaf(v.w, sc.w, sh.w);
                  best.x = fmaxf(best.x, v.x); best.y = fmaxf(best.y, v.y); best.z = fmaxf(best.z, v.z); best.w = fmaxf(best.w, v.w);
                }
                if (act == ACT_RELU) {
                  best.x = fmaxf(best.x, 0.f); best.y = fmaxf(best.y, 0.f); best.z = fmaxf(best.z, 0.f); best.w = fmaxf(best.w, 0.f);
                }
                const size_t o = (size_t)(m0 >> 2) * ldc + n;
                if (out) *reinterpret_cast<float4*>(out + o) = best;
                if (out_hi) {
                  const float f[4] = {best.x, best.y, best.z, best.w};
                  uint32_t hw[2], lw[2];
#pragma unroll
                  for (int u = 0; u < 2; ++u) {
                    const __nv_bfloat16 h0 = __float2bfloat16_rn(f[2 * u]), h1 = __float2bfloat16_rn(f[2 * u + 1]);
                    hw[u] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
                    const __nv_bfloat16 l0 = __float2bfloat16_rn(f[2 * u] - __bfloat162float(h0));
                    const __nv_bfloat16 l1 = __float2bfloat16_rn(f[2 * u + 1] - __bfloat162float(h1));
                    lw[u] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
                  }
                  *reinterpret_cast<uint2*>(out_hi + o) = make_uint2(hw[0], hw[1]);
                  if (out_lo) *reinterpret_cast<uint2*>(out_lo + o) = make_uint2(lw[0], lw[1]);
                }
              }
            }
          }
        } else
        if (n < N) {
          float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
          if (scale) sc = __ldg(reinterpret_cast<const float4*>(scale + n));
          if (shift) sh = __ldg(reinterpret_cast<const float4*>(shift + n));
          float4 rr[8];   // residual rows of this chunk: all eight loads in flight before the first use
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int m = m_first + 4 * i;
            rr[i] = (res && m < M) ? __ldg(reinterpret_cast<const float4*>(res + (size_t)m * ldr + n)) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int m = m_first + 4 * i;
            if (m < M) {
              float4 v = *reinterpret_cast<const float4*>(stg + (sub_r + 4 * i) * TC_EPI_PITCH + c4);
              v.x = fmaf(v.x, sc.x, sh.x) + rr[i].x; v.y = fmaf(v.y, sc.y, sh.y) + rr[i].y;
              v.z = fmaf(v.z, sc.z, sh.z) + rr[i].z; v.w = fmaf(v.w, sc.w, sh.w) + rr[i].w;
              if (act == ACT_RELU) {
                v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
              } else if (act == ACT_GELU) {
                v = tc::gelu_erf4(v);
              }
              if (out2 != nullptr && n >= n_split) {
                float* dst = out2 + (size_t)m * ldc2 + n;
                if (p.out2_bf16) {   // bf16 KV cache: same element offsets, 2-byte elements
                  const __nv_bfloat162 lo2 = __floats2bfloat162_rn(v.x, v.y), hi2 = __floats2bfloat162_rn(v.z, v.w);
                  uint2 pk;
                  pk.x = *reinterpret_cast<const uint32_t*>(&lo2);
                  pk.y = *reinterpret_cast<const uint32_t*>(&hi2);
                  *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.out2) + (dst - p.out2)) = pk;
                } else {
                  *reinterpret_cast<float4*>(dst) = v;
                }
                continue;
              }
              const size_t o = (size_t)m * ldc + n;
              if (out) *reinterpret_cast<float4*>(out + o) = v;   // fp32 copy only when a consumer reads it
              if (out_hi) {
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
                *reinterpret_cast<uint2*>(out_hi + o) = make_uint2(hw[0], hw[1]);
                if (out_lo) *reinterpret_cast<uint2*>(out_lo + o) = make_uint2(lw[0], lw[1]);
              }
            }
          }
        }
        __syncwarp();
      }
      tc::tcgen05_before_sync();
      tc::mbar_arrive(tempty_bar(acc));
    }
  } else if (warp < TMA_WARP) {
    // =========================== A producers: cp.async from the bf16 NHWC planes ===========================
    // ROWS rows x 4 chunks (16 B) per plane and k-block: 2 * MT rows per thread.
    if constexpr (!A_TMA) {
    constexpr int RPT = 2 * MT;
    const int pt = threadIdx.x - EPI_WARPS * 32;   // 0..255
    const int chunk = pt & 3;                      // 16-byte chunk of the 64-byte operand row
    const int rg = pt >> 2;                        // rows rg + 64*i
    const __nv_bfloat16* const xh = p.x_hi;
    const __nv_bfloat16* const xl = p.x_lo;
    int kit = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int tm = tile / tiles_n;
      long long base[RPT];
      int ih0[RPT], iw0[RPT];
      bool ok[RPT];
      uint32_t soff[RPT];
#pragma unroll
      for (int i = 0; i < RPT; ++i) {
        const int r = rg + 64 * i;
        const int m = tm * ROWS + r;
        ok[i] = m < p.M;
        const int mm = ok[i] ? m : 0;
        int ow, oh, b;
        if (p.pool) {   // window-major pixel order: m = 4 * (pooled pixel) + 2 * dy + dx
          const int wdw = mm >> 2, qd = mm & 3, pw_ = p.OW >> 1, ph_ = p.OH >> 1;
          const int px = wdw % pw_, t = wdw / pw_;
          ow = 2 * px + (qd & 1); oh = 2 * (t % ph_) + (qd >> 1); b = t / ph_;
        } else {
          ow = mm % p.OW;
          const int t = mm / p.OW;
          oh = t % p.OH;
          b = t / p.OH;
        }
        ih0[i] = oh * p.SH - p.PH;
        iw0[i] = ow * p.SW - p.PW;
        base[i] = (long long)b * p.H * p.W * p.C;
        soff[i] = (uint32_t)(r >> 3) * 512u + (uint32_t)(r & 7) * 64u + (uint32_t)((chunk ^ ((r >> 1) & 3)) << 4);   // m-tiles are contiguous 8 KB blocks
      }
      for (int kb = 0; kb < nkb; ++kb, ++kit) {
        const int s = kit % STAGES;
        const int k = kb * Cfg::KB_ELEMS + chunk * Cfg::CH_ELEMS;
        const bool kok = k < p.K;
        const int tap = kok ? k / p.C : 0;
        const int ci = k - tap * p.C;
        const int kh = tap / p.KW, kw = tap - kh * p.KW;
        tc::mbar_wait(empty_bar(s), ((kit / STAGES) & 1) ^ 1);
        const uint32_t a_hi = smem_base + s * Cfg::STAGE_BYTES;
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
          const int ih = ih0[i] + kh, iw = iw0[i] + kw;
          const bool valid = kok && ok[i] && (unsigned)ih < (unsigned)p.H && (unsigned)iw < (unsigned)p.W;
          const long long e = valid ? base[i] + ((long long)ih * p.W + iw) * p.C + ci : 0;
          tc::cp_async_16(a_hi + soff[i], xh + e, valid ? 16u : 0u);
          if (PLANES == 2) tc::cp_async_16(a_hi + Cfg::A_BYTES + soff[i], xl + e, valid ? 16u : 0u);
        }
        tc::cp_async_mbar_arrive_noinc(full_bar(s));   // arrives when this thread's copies above have landed
      }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    }  // !A_TMA
  } else if (warp == TMA_WARP) {
    // =========================== W producer (TMA, 64-byte rows) [+ the A operand by TMA im2col] ===========================
    if (lane == 0) {
      int kit = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int tm = tile / tiles_n, tn = tile - tm * tiles_n;
        int aw[MT], ah[MT], an[MT];
        if constexpr (A_TMA) {
#pragma unroll
          for (int h = 0; h < MT; ++h) {   // base filter position of the first pixel of m-tile h
            const int m0 = tm * ROWS + h * TC_BM;
            const int ow = m0 % p.OW, t = m0 / p.OW;
            aw[h] = ow * p.SW - p.PW; ah[h] = (t % p.OH) * p.SH - p.PH; an[h] = t / p.OH;
          }
        }
        for (int kb = 0; kb < nkb; ++kb, ++kit) {
          const int s = kit % STAGES;
          tc::mbar_wait(empty_bar(s), ((kit / STAGES) & 1) ^ 1);
          int a_tiles = 0;   // m-tiles of this CTA tile that hold pixels (the second half of a ragged MT = 2 tile may be empty)
          if constexpr (A_TMA) {
#pragma unroll
            for (int h = 0; h < MT; ++h) a_tiles += an[h] < p.B ? 1 : 0;
          }
          tc::mbar_arrive_expect_tx(full_bar(s), PLANES * (Cfg::B_BYTES + a_tiles * Cfg::A_HALF_BYTES));
          if constexpr (A_TMA) {
            const int k = kb * Cfg::KB_ELEMS;
            const int tap = k / p.C, ci = k - tap * p.C;
            const int kh = tap / p.KW, kw = tap - kh * p.KW;
            const uint32_t a_hi = smem_base + s * Cfg::STAGE_BYTES;
#pragma unroll
            for (int h = 0; h < MT; ++h) {
              if (an[h] < p.B) {
                tc::tma_load_im2col_4d(a_hi + h * Cfg::A_HALF_BYTES, &amap_hi, full_bar(s), ci, aw[h], ah[h], an[h], (uint16_t)kw, (uint16_t)kh);
                if (PLANES == 2)
                  tc::tma_load_im2col_4d(a_hi + Cfg::A_BYTES + h * Cfg::A_HALF_BYTES, &amap_lo, full_bar(s), ci, aw[h], ah[h], an[h], (uint16_t)kw, (uint16_t)kh);
              }
            }
          }
          const uint32_t b_hi = smem_base + s * Cfg::STAGE_BYTES + PLANES * Cfg::A_BYTES;
          tc::tma_load_2d(b_hi, &map_hi, full_bar(s), kb * Cfg::KB_ELEMS, tn * BN);
          if (PLANES == 2) tc::tma_load_2d(b_hi + Cfg::B_BYTES, &map_lo, full_bar(s), kb * Cfg::KB_ELEMS, tn * BN);
        }
      }
    }
    __syncwarp();
  } else {
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      constexpr uint32_t idesc = tc::make_idesc(1, BN);
      int kit = 0, it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int acc = MT == 1 ? (it & 1) : 0;
        tc::mbar_wait(tempty_bar(acc), (MT == 1 ? ((it >> 1) & 1) : (it & 1)) ^ 1);
        tc::tcgen05_after_sync();
        for (int kb = 0; kb < nkb; ++kb, ++kit) {
          const int s = kit % STAGES;
          tc::mbar_wait(full_bar(s), (kit / STAGES) & 1);
          tc::tcgen05_after_sync();
          const uint32_t a_hi = smem_base + s * Cfg::STAGE_BYTES;
          const uint32_t b_hi = a_hi + PLANES * Cfg::A_BYTES;
          const uint64_t db_hi = tc::make_smem_desc_sw64(b_hi), db_lo = tc::make_smem_desc_sw64(b_hi + Cfg::B_BYTES);
#pragma unroll
          for (int h = 0; h < MT; ++h) {   // the m-tiles of this CTA share the weight stage
            const uint32_t d = tmem_base + (uint32_t)((MT == 1 ? acc : h) * BN);
            const uint64_t da_hi = tc::make_smem_desc_sw64(a_hi + h * Cfg::A_HALF_BYTES);
            const uint64_t da_lo = tc::make_smem_desc_sw64(a_hi + Cfg::A_BYTES + h * Cfg::A_HALF_BYTES);
#pragma unroll
            for (int k = 0; k < 2; ++k)   // 2 x 32 B = one 64-byte row
              tc::umma<false>(d, da_hi + 2 * k, db_hi + 2 * k, idesc, (kb | k) != 0);
            if constexpr (PASSES == 3) {
#pragma unroll
              for (int k = 0; k < 2; ++k) tc::umma<false>(d, da_lo + 2 * k, db_hi + 2 * k, idesc, 1u);
#pragma unroll
              for (int k = 0; k < 2; ++k) tc::umma<false>(d, da_hi + 2 * k, db_lo + 2 * k, idesc, 1u);
            }
          }
          tc::umma_commit(empty_bar(s));
        }
        tc::umma_commit(tfull_bar(acc));
      }
    }
    __syncwarp();
  }
  tc::tcgen05_before_sync();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc::tcgen05_after_sync();
    tc::tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// host side ---------------------------------------------------------------------------------------------------------
// Extra weight tensor maps with 64-byte boxes (32 bf16) and 64B swizzle, for 64 / 128 / 256 weight rows.
struct Tc3Maps {
  bool ready = false;
  CUtensorMap hi[3], lo[3];
};

inline cudaError_t tc3_prepare_maps(const TcWeight& w, Tc3Maps* out) {
  out->ready = false;
  if (!w.ready || (w.precision != 2 && w.precision != 3)) return cudaSuccess;
  PFN_encodeTiled enc = tc_encode_fn();
  if (!enc) return cudaErrorNotSupported;
  const cuuint64_t gdim[2] = {(cuuint64_t)w.K, (cuuint64_t)w.N};
  const cuuint64_t gstr[1] = {(cuuint64_t)w.K * 2};
  const cuuint32_t estr[2] = {1, 1};
  for (int i = 0; i < 3; ++i) {
    const cuuint32_t box[2] = {32, (cuuint32_t)(64 << i)};
    CUresult r = enc(&out->hi[i], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, w.hi, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return cudaErrorInvalidValue;
    out->lo[i] = out->hi[i];
    if (w.lo) {
      r = enc(&out->lo[i], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, w.lo, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return cudaErrorInvalidValue;
    }
  }
  out->ready = true;
  return cudaSuccess;
}

inline bool tc3_supported(const ConvGemm& p, int precision) {
  if (precision != 2 && precision != 3) return false;
  if (p.x_hi == nullptr || (precision == 2 && p.x_lo == nullptr)) return false;
  if (p.pool && (p.res != nullptr || (p.OH & 1) || (p.OW & 1) || p.M % 4 != 0)) return false;
  if (p.out2 != nullptr && (p.pool || p.n_split % 32 != 0 || p.ldc2 % 4 != 0 || p.dyn_mul2 % 4 != 0)) return false;
  return p.a_map_hi == nullptr && p.C % 8 == 0 && p.K % 8 == 0 && p.ldc % 4 == 0 &&
         (p.res == nullptr || p.ldr % 4 == 0) && p.N % 4 == 0;
}

typedef CUresult (*PFN_encodeIm2col)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                     const int*, const int*, cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave,
                                     CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeIm2col tc3_encode_im2col_fn() {
  static PFN_encodeIm2col fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &ptr, cudaEnableDefault, &q) == cudaSuccess && ptr) fn = (PFN_encodeIm2col)ptr;
  }
  return fn;
}

// The activation operand by TMA im2col: stride 1, symmetric padding, whole k-blocks inside one filter tap, row-major pixels.
inline bool tc3_a_tma_supported(const ConvGemm& p) {
  return !p.pool && p.SH == 1 && p.SW == 1 && p.PH == p.PW && p.KH == p.KW && p.C % 32 == 0 && p.PH < 128 &&
         p.OH == p.H + 2 * p.PH - p.KH + 1 && p.OW == p.W + 2 * p.PW - p.KW + 1 && tc3_encode_im2col_fn() != nullptr;
}

// Tensor map over one bf16 NHWC plane [B, H, W, C]: bounding box of the filter's base position = [-pad, dim + pad - k],
// 32 channels (one 64-byte operand row) x 128 pixels per load, 64B swizzle = the layout the UMMA descriptor reads.
inline cudaError_t tc3_make_a_map(const __nv_bfloat16* plane, const ConvGemm& p, CUtensorMap* out, int channels = 32) {
  const cuuint64_t gdim[4] = {(cuuint64_t)p.C, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)p.B};
  const cuuint64_t gstr[3] = {(cuuint64_t)p.C * 2, (cuuint64_t)p.W * p.C * 2, (cuuint64_t)p.H * p.W * p.C * 2};
  const int lower[2] = {-p.PW, -p.PH};
  const int upper[2] = {p.PW - (p.KW - 1), p.PH - (p.KH - 1)};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = tc3_encode_im2col_fn()(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<__nv_bfloat16*>(plane), gdim, gstr, lower,
                                      upper, (cuuint32_t)channels, TC_BM, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                      channels == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

inline bool& tc3_use_a_tma() {
  static bool on = true;
  return on;
}

template <int PASSES, int BN, int MT = 1>
inline cudaError_t tc3_launch_one(const ConvGemm& p, const Tc3Maps& m, cudaStream_t s, int num_sms) {
  using Cfg = Tc3Cfg<PASSES, BN, MT>;
  static bool attr_set = false, attr_set_t = false;
  const int tiles_m_ = (p.M + MT * TC_BM - 1) / (MT * TC_BM), tiles_n_ = (p.N + BN - 1) / BN;
  constexpr int mi_ = BN == 64 ? 0 : (BN == 128 ? 1 : 2);
  if (tc3_use_a_tma() && tc3_a_tma_supported(p)) {
    auto kern_t = conv_gemm_tc3_kernel<PASSES, BN, MT, true>;
    if (!attr_set_t) {
      cudaError_t st = cudaFuncSetAttribute(kern_t, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM_BYTES);
      if (st != cudaSuccess) return st;
      attr_set_t = true;
    }
    CUtensorMap ah, al;
    cudaError_t st = tc3_make_a_map(p.x_hi, p, &ah);
    if (st != cudaSuccess) return st;
    al = ah;
    if (PASSES == 3 && (st = tc3_make_a_map(p.x_lo, p, &al)) != cudaSuccess) return st;
    const int grid_t = tiles_m_ * tiles_n_ < num_sms ? tiles_m_ * tiles_n_ : num_sms;
    return launch_kernel(kern_t, dim3(grid_t), dim3(Cfg::THREADS), Cfg::SMEM_BYTES, s, p, m.hi[mi_], m.lo[mi_], tiles_m_, tiles_n_, ah, al);
  }
  auto kern = conv_gemm_tc3_kernel<PASSES, BN, MT>;
  if (!attr_set) {
    cudaError_t st = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM_BYTES);
    if (st != cudaSuccess) return st;
    attr_set = true;
  }
  const int tiles_m = (p.M + MT * TC_BM - 1) / (MT * TC_BM), tiles_n = (p.N + BN - 1) / BN;
  const int grid = tiles_m * tiles_n < num_sms ? tiles_m * tiles_n : num_sms;
  constexpr int mi = BN == 64 ? 0 : (BN == 128 ? 1 : 2);
  return launch_kernel(kern, dim3(grid), dim3(Cfg::THREADS), Cfg::SMEM_BYTES, s, p, m.hi[mi], m.lo[mi], tiles_m, tiles_n, m.hi[mi], m.lo[mi]);
}

inline cudaError_t launch_conv_gemm_tc3(const ConvGemm& p, const Tc3Maps& m, int precision, cudaStream_t s, int num_sms) {
  int bn = tc_pick_bn(p.M, p.N, num_sms);
  // short-K problems (the ViT blocks' Linears) are bound by their epilogue, not by the main loop: the 128-wide tile has two
  // epilogue warps per TMEM quadrant
  if (bn == 256 && p.K <= 1024 && (p.KH == 1 && p.KW == 1)) bn = 128;
  // Two m-tiles per CTA (both accumulators against one weight stage) pay in exactly one place: single-pass bf16 on the
  // 512-channel 3x3 convolutions without residual (measured 0.828 -> 0.734 ms per launch; there the weight stage is re-read
  // for half as many MMAs).  Everywhere else, and in the 3-pass mode, the lost epilogue overlap costs more (encoder
  // 24.4 -> 27.4 ms in bf16, 34.7 -> 39.9 ms in bf16x3 when applied everywhere).
  const bool want = precision == 3 && p.N >= 512 && p.K >= 4608 && p.res == nullptr;
  const bool mt2 = want && bn == 256 && (long long)((p.M + 255) / 256) * (p.N / 256) >= 2LL * num_sms;
  if (mt2) return tc3_launch_one<1, 256, 2>(p, m, s, num_sms);
  if (precision == 2) {
    switch (bn) {
      case 256: return tc3_launch_one<3, 256>(p, m, s, num_sms);
      case 128: return tc3_launch_one<3, 128>(p, m, s, num_sms);
      default: return tc3_launch_one<3, 64>(p, m, s, num_sms);
    }
  }
  switch (bn) {
    case 256: return tc3_launch_one<1, 256>(p, m, s, num_sms);
    case 128: return tc3_launch_one<1, 128>(p, m, s, num_sms);
    default: return tc3_launch_one<1, 64>(p, m, s, num_sms);
  }
}

}  // namespace d2t

// tcgen05 / TMEM / TMA implicit-GEMM contraction for sm_100a.
//
//   out[m, n] = act((sum_k A[m,k] W[n,k]) * scale[n] + shift[n] + res[m,n])
//
// A is gathered on the fly from the fp32 NHWC activation tensor (implicit GEMM over (kh, kw, cin)),
// W is the pre-packed K-major weight matrix.  One persistent CTA per SM, warp-specialised:
//
//   warps 0-3   epilogue      tcgen05.ld accumulator rows -> folded BN / bias / residual / activation -> HBM
//   warps 4-11  A producers   LDG (coalesced 128 B runs of the NHWC row) -> split / convert -> swizzled STS
//   warp  12    W producer    TMA (cp.async.bulk.tensor, 128B swizzle) of the weight tile, one per operand plane
//   warp  13    MMA issuer    one thread issues tcgen05.mma (M=128, N=BLOCK_N, fp32 accumulators in TMEM)
//
// Pipelines: a STAGES-deep smem ring (full/empty mbarriers) between producers and the MMA thread, and a
// two-deep TMEM accumulator ring (tmem_full/tmem_empty) between the MMA thread and the epilogue, so the
// epilogue of tile i overlaps the main loop of tile i+1.
//
// Precisions (d2t_precision):
//   BF16     one pass:   bf16(A) . bf16(W)                                   kind::f16
//   BF16X3   3 passes:   A_hi.W_hi + A_lo.W_hi + A_hi.W_lo, hi/lo bf16       kind::f16   (~2^-16 relative)
//   TF32X3   3 passes:   same split with tf32 hi/lo kept in fp32 containers  kind::tf32  (~2^-21 relative)
// The split of A happens in registers inside the producer warps, W planes are split once at load time.
#pragma once
#include <cuda.h>
#include <vector>
#include "common.cuh"

namespace d2t {

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
namespace tc {

__device__ __noinline__ float4 gelu_erf4(float4 v) {
  return make_float4(gelu_erf(v.x), gelu_erf(v.y), gelu_erf(v.z), gelu_erf(v.w));
}
__device__ __forceinline__ long long gtime() {
  long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n.reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n}\n"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {}
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
template <bool TF32>
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if constexpr (TF32) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
  } else {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
  }
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, 128-byte-swizzled operand tile: rows of 128 B, 8-row atoms of 1024 B (SBO), version 1 (sm_100).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);        // start address, bits [0,14)
  d |= (uint64_t)1 << 16;                         // leading byte offset (unused for swizzled K-major), bits [16,30)
  d |= (uint64_t)(1024 >> 4) << 32;               // stride byte offset = 1024 B, bits [32,46)
  d |= (uint64_t)1 << 46;                         // descriptor version, bits [46,48)
  d |= (uint64_t)2 << 61;                         // layout type SWIZZLE_128B, bits [61,64)
  return d;
}
// Instruction descriptor: fp32 accumulate, A/B K-major, formats BF16 (1) or TF32 (2), M=128.
__host__ __device__ constexpr uint32_t make_idesc(int fmt, int n) {
  return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

}  // namespace tc

// ------------------------------------------------------------------------------------------------
// Kernel configuration
// ------------------------------------------------------------------------------------------------
constexpr int TC_BM = 128;
// warp roles: gather variant = 4 epilogue + 8 A-producer + TMA + MMA warps (448 threads);
//             TMA-fed-A variant = 8 epilogue + TMA + MMA warps (320 threads; nothing to gather)
constexpr int TC_PROD_WARPS = 8;
constexpr int TC_PROD_THREADS = TC_PROD_WARPS * 32;                  // 256
constexpr int TC_ROWS_PER_THREAD = TC_BM * 8 / TC_PROD_THREADS;      // 4 (8 x 16-byte chunks per 128-byte row)
constexpr int TC_EPI_PITCH = 36;                                     // floats per staged accumulator row (32 + pad)
__host__ __device__ constexpr int tc_epi_warps(bool a_tma) { return a_tma ? 8 : 4; }
__host__ __device__ constexpr int tc_prod_warps(bool a_tma) { return a_tma ? 0 : TC_PROD_WARPS; }
__host__ __device__ constexpr int tc_threads(bool a_tma) { return (tc_epi_warps(a_tma) + tc_prod_warps(a_tma) + 2) * 32; }

template <bool TF32, int PASSES, int BN, bool A_TMA>
struct TcCfg {
  static constexpr int PLANES = PASSES == 1 ? 1 : 2;
  static constexpr int KB_ELEMS = TF32 ? 32 : 64;   // elements per 128-byte operand row
  static constexpr int CH_ELEMS = TF32 ? 4 : 8;     // elements per 16-byte chunk
  static constexpr int A_BYTES = TC_BM * 128;       // per plane
  static constexpr int B_BYTES = BN * 128;          // per plane
  static constexpr int STAGE_BYTES = PLANES * (A_BYTES + B_BYTES);
  // accumulator transpose tiles of the epilogue.  The TMA-fed-A variant runs ONE tile per CTA (host guarantees
  // grid == tiles), so by epilogue time every operand stage is free and the transpose tiles alias stage memory:
  // that buys a 4th stage, i.e. all of a K=256 tile's operands in flight at once.
  static constexpr int EPI_TILE_BYTES = tc_epi_warps(A_TMA) * 32 * TC_EPI_PITCH * 4;
  static constexpr int EPI_STAGE_BYTES = A_TMA ? 0 : EPI_TILE_BYTES;
  static constexpr int SMEM_BUDGET = 225 * 1024 - EPI_STAGE_BYTES - 1280;
  static constexpr int STAGES_RAW = SMEM_BUDGET / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 6 ? 6 : STAGES_RAW;
  static constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;   // two accumulators; power of two (BN in {64,128,256})
  static constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_BYTES + EPI_STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
  static_assert(STAGES >= 2, "need at least a double buffer");
  static_assert(!A_TMA || STAGES * STAGE_BYTES >= EPI_TILE_BYTES, "epilogue tiles must fit in the aliased stage memory");
};

struct TcWeight {
  bool ready = false;
  int precision = 0, N = 0, K = 0;
  void* hi = nullptr;
  void* lo = nullptr;
  CUtensorMap map_hi[3], map_lo[3];  // TMA boxes of 64 / 128 / 256 weight rows
};

// A_TMA: the A operand planes already exist in HBM as bf16 [M, K] matrices (written by the producing kernel's
// epilogue) and are fetched by TMA like the weights; the gather/convert warps then have nothing to do.  This is the
// low-latency path for the small decode-step GEMMs: every k-block of a tile is in flight at once.
// STACK (TMA-fed-A, bf16x3): a tcgen05.mma retires every ~90 ns whatever its N (<= 256), so the three passes of a k-step
// are issued as TWO instructions: A_hi x [W_hi ; W_lo] (the two weight planes are adjacent in the stage = one 2*BN-row B
// operand, accumulator columns [0, 2*BN)) and A_lo x W_hi (columns [0, BN)); the epilogue adds the two column halves.
// The one-tile-per-CTA variant never uses its second accumulator, so TMEM holds the wide one for free.  A third fewer
// serial MMAs on the critical path of every decode-step projection.
template <bool TF32, int PASSES, int BN, bool A_TMA, bool STACK = false>
__global__ void __launch_bounds__(tc_threads(A_TMA), 1)
conv_gemm_tc_kernel(const ConvGemm p, const __grid_constant__ CUtensorMap map_hi,
                    const __grid_constant__ CUtensorMap map_lo, const __grid_constant__ CUtensorMap map_a_hi,
                    const __grid_constant__ CUtensorMap map_a_lo, int tiles_m, int tiles_n) {
  using Cfg = TcCfg<TF32, PASSES, BN, A_TMA>;
  constexpr int STAGES = Cfg::STAGES, PLANES = Cfg::PLANES;
  constexpr int TC_EPI_WARPS = tc_epi_warps(A_TMA), TC_EPI_STAGE_BYTES = Cfg::EPI_STAGE_BYTES;
  constexpr int TMA_WARP = tc_epi_warps(A_TMA) + tc_prod_warps(A_TMA), MMA_WARP = TMA_WARP + 1;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - tc::smem_u32(smem_raw));
  // stage s: [A plane 0][A plane 1][B plane 0][B plane 1]
  const uint32_t bars = smem_base + STAGES * Cfg::STAGE_BYTES + TC_EPI_STAGE_BYTES;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bars + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bars + 8u * (2 * STAGES + 2 + a); };
  const uint32_t tmem_slot = bars + 8u * (2 * STAGES + 4);
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + STAGES * Cfg::STAGE_BYTES + TC_EPI_STAGE_BYTES + 8 * (2 * STAGES + 4));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool dbg = p.dbg != nullptr && blockIdx.x == 0;
  if (dbg && threadIdx.x == 0) p.dbg[0] = tc::gtime();
  const int KS = (A_TMA && p.k_splits > 1) ? p.k_splits : 1;   // split-K slices (TMA-fed-A variant only)
  const int tiles_mn = tiles_m * tiles_n;
  const int num_tiles = tiles_mn * KS;
  const int nkb_total = (p.K + Cfg::KB_ELEMS - 1) / Cfg::KB_ELEMS;
  const int nkb = nkb_total / KS;   // host guarantees divisibility

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      tc::mbar_init(full_bar(s), A_TMA ? 1 : TC_PROD_THREADS + 1);   // A-producer arrivals + the TMA thread's expect_tx arrive
      tc::mbar_init(empty_bar(s), 1);                     // one tcgen05.commit
    }
    for (int a = 0; a < 2; ++a) {
      tc::mbar_init(tfull_bar(a), 1);                     // one tcgen05.commit
      tc::mbar_init(tempty_bar(a), TC_EPI_WARPS * 32);    // every epilogue thread
    }
    tc::fence_barrier_init();
  }
  if (warp == MMA_WARP) tc::tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  if (warp == TMA_WARP && lane == 0) {
    tc::tma_prefetch_desc(&map_hi);
    if (PLANES == 2) tc::tma_prefetch_desc(&map_lo);
    if (A_TMA) {
      tc::tma_prefetch_desc(&map_a_hi);
      if (PLANES == 2) tc::tma_prefetch_desc(&map_a_lo);
    }
  }
  tc::tcgen05_before_sync();
  __syncthreads();
  tc::tcgen05_after_sync();
  const uint32_t tmem_base = *tmem_slot_gen;
  // TMA-fed-A variant (one tile per CTA): the weight tiles do not depend on the previous kernel, so their loads for the
  // first stages are issued BEFORE the dependency wait and overlap the predecessor's tail; the activation tiles follow.
  int w_prefetched = 0;
  if constexpr (A_TMA) {
    w_prefetched = nkb < STAGES ? nkb : STAGES;
    if (warp == TMA_WARP && lane == 0 && !(p.act & 64) && (int)blockIdx.x < num_tiles) {
      const int tile = blockIdx.x;
      const int ks = tile / tiles_mn, tmn = tile - ks * tiles_mn;
      const int tn = tmn % tiles_n;
      for (int kb = 0; kb < w_prefetched; ++kb) {
        const int kcoord = (ks * nkb + kb) * Cfg::KB_ELEMS;
        tc::mbar_arrive_expect_tx(full_bar(kb), PLANES * (Cfg::B_BYTES + Cfg::A_BYTES));
        const uint32_t b_hi = smem_base + kb * Cfg::STAGE_BYTES + PLANES * Cfg::A_BYTES;
        tc::tma_load_2d(b_hi, &map_hi, full_bar(kb), kcoord, tn * BN);
        if (PLANES == 2) tc::tma_load_2d(b_hi + Cfg::B_BYTES, &map_lo, full_bar(kb), kcoord, tn * BN);
      }
    }
  }
  pdl_wait();      // everything above (barrier init, TMEM alloc, descriptor prefetch, weight tiles) overlapped the previous kernel
  pdl_trigger();   // now let ONE successor pre-launch (pre-launched CTAs pin 200 KB of smem each while they wait)
  if (dbg && threadIdx.x == 0) p.dbg[1] = tc::gtime();

  if (warp < TC_EPI_WARPS) {
    // =========================== epilogue ===========================
    // Warp w owns TMEM lane quadrant (w & 3) = accumulator rows 32*(w&3)..+31; with 8 epilogue warps the 32-column
    // chunks of a tile alternate between the two warps of a quadrant.  Row-per-thread chunks are transposed through
    // a per-warp smem tile so global traffic is coalesced (8 lanes = one 128-byte row segment, 4 rows per
    // instruction).  Parameters are hoisted into registers and the row loop is kept rolled: this code is
    // instruction-latency bound (one warp per scheduler), not bandwidth bound.
    const int quad = warp & 3, slot = warp >> 2;
    constexpr int NSLOT = TC_EPI_WARPS / 4;
    float* const stg = reinterpret_cast<float*>(smem_gen + (A_TMA ? (size_t)0 : (size_t)STAGES * Cfg::STAGE_BYTES)) + warp * (32 * TC_EPI_PITCH);
    const int sub_r = lane >> 3, c4 = (lane & 7) * 4;
    const int M = p.M, N = p.N, ldc = p.ldc, ldr = p.ldr, ldc2 = p.ldc2, n_split = p.n_split, act = p.act & 15;
    const float* const scale = p.scale;
    float* out2 = nullptr;   // columns >= n_split (KV-cache slot of the current decode step)
    if (p.out2) out2 = p.out2 + (p.dyn ? (long long)(*p.dyn) * p.dyn_mul2 : 0) - n_split;
    __nv_bfloat16* const out_hi = p.out_hi;
    __nv_bfloat16* const out_lo = p.out_lo;
    // Epilogue operands that do not depend on the accumulator (bias / BN scale+shift, residual rows) are fetched
    // BEFORE waiting for the MMA, and for the following chunk while the current one is being written, so their
    // L2 latency is off the critical path of the small decode-step GEMMs.
    float4 sc, sh, rr0, rr1, rr2, rr3, rr4, rr5, rr6, rr7;
#define TC_EPI_PREFETCH(TM_, TN_, J_)                                                                              \
    do {                                                                                                             \
      const int n_ = (TN_) * BN + (J_) * 32 + c4;                                                                    \
      const int m0_ = (TM_) * TC_BM + quad * 32 + sub_r;                                                             \
      const float4 z_ = make_float4(0.f, 0.f, 0.f, 0.f);                                                             \
      sc = make_float4(1.f, 1.f, 1.f, 1.f);                                                                          \
      sh = z_;                                                                                                       \
      if (n_ < N) {                                                                                                  \
        if (scale) sc = __ldg(reinterpret_cast<const float4*>(scale + n_));                                          \
        if (shift) sh = __ldg(reinterpret_cast<const float4*>(shift + n_));                                          \
      }                                                                                                              \
      const bool ok_ = res != nullptr && n_ < N;                                                                     \
      const float* rp_ = res + (size_t)m0_ * ldr + n_;                                                               \
      rr0 = (ok_ && m0_ + 0 < M) ? __ldg(reinterpret_cast<const float4*>(rp_)) : z_;                                 \
      rr1 = (ok_ && m0_ + 4 < M) ? __ldg(reinterpret_cast<const float4*>(rp_ + (size_t)4 * ldr)) : z_;               \
      rr2 = (ok_ && m0_ + 8 < M) ? __ldg(reinterpret_cast<const float4*>(rp_ + (size_t)8 * ldr)) : z_;               \
      rr3 = (ok_ && m0_ + 12 < M) ? __ldg(reinterpret_cast<const float4*>(rp_ + (size_t)12 * ldr)) : z_;             \
      rr4 = (ok_ && m0_ + 16 < M) ? __ldg(reinterpret_cast<const float4*>(rp_ + (size_t)16 * ldr)) : z_;             \
      rr5 = (ok_ && m0_ + 20 < M) ? __ldg(reinterpret_cast<const float4*>(rp_ + (size_t)20 * ldr)) : z_;             \
      rr6 = (ok_ && m0_ + 24 < M) ? __ldg(reinterpret_cast<const float4*>(rp_ + (size_t)24 * ldr)) : z_;             \
      rr7 = (ok_ && m0_ + 28 < M) ? __ldg(reinterpret_cast<const float4*>(rp_ + (size_t)28 * ldr)) : z_;             \
    } while (0)
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int ks = tile / tiles_mn, tmn = tile - ks * tiles_mn;
      const int tm = tmn / tiles_n, tn = tmn - tm * tiles_n;
      const int acc = it & 1;
      // split-K: slice ks > 0 stores the bare partial sum at its own offset
      const float* const shift = ks == 0 ? p.shift : nullptr;
      const float* const res = ks == 0 ? p.res : nullptr;
      float* const out = p.out + (size_t)ks * p.split_stride;
      if (slot < BN / 32) TC_EPI_PREFETCH(tm, tn, slot);
      tc::mbar_wait(tfull_bar(acc), (it >> 1) & 1);
      tc::tcgen05_after_sync();
      if (dbg && threadIdx.x == 0) p.dbg[5] = tc::gtime();
      const int m_first = tm * TC_BM + quad * 32 + sub_r;
#pragma unroll 1
      for (int j = slot; j < BN / 32; j += NSLOT) {
        uint32_t r[32];
        tc::tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN + j * 32), r);
        if constexpr (STACK) {   // columns [BN, 2*BN) hold A_hi x W_lo of the same outputs
          uint32_t r2[32];
          tc::tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(BN + j * 32), r2);
          tc::tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 32; ++q) r[q] = __float_as_uint(__uint_as_float(r[q]) + __uint_as_float(r2[q]));
        } else {
          tc::tmem_ld_wait();
        }
#pragma unroll
        for (int q = 0; q < 8; ++q)
          *reinterpret_cast<uint4*>(stg + lane * TC_EPI_PITCH + q * 4) = make_uint4(r[q * 4], r[q * 4 + 1], r[q * 4 + 2], r[q * 4 + 3]);
        __syncwarp();
        const int n = tn * BN + j * 32 + c4;
        float4 v[8];
        const float4 rr[8] = {rr0, rr1, rr2, rr3, rr4, rr5, rr6, rr7};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          v[i] = *reinterpret_cast<const float4*>(stg + (sub_r + 4 * i) * TC_EPI_PITCH + c4);
          v[i].x = fmaf(v[i].x, sc.x, sh.x) + rr[i].x; v[i].y = fmaf(v[i].y, sc.y, sh.y) + rr[i].y;
          v[i].z = fmaf(v[i].z, sc.z, sh.z) + rr[i].z; v[i].w = fmaf(v[i].w, sc.w, sh.w) + rr[i].w;
        }
        __syncwarp();
        if (j + NSLOT < BN / 32) TC_EPI_PREFETCH(tm, tn, j + NSLOT);   // operands of the next chunk, in flight during the stores
        if (n < N) {
          const bool second = out2 != nullptr && n >= n_split;
          float* dst = second ? out2 + (size_t)m_first * ldc2 + n : out + (size_t)m_first * ldc + n;
          const size_t dst_step = (size_t)4 * (second ? ldc2 : ldc);
          const bool planes = out_hi != nullptr && !second;
          if (act == ACT_RELU) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              v[i].x = fmaxf(v[i].x, 0.f); v[i].y = fmaxf(v[i].y, 0.f); v[i].z = fmaxf(v[i].z, 0.f); v[i].w = fmaxf(v[i].w, 0.f);
            }
          } else if (act == ACT_GELU) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = tc::gelu_erf4(v[i]);   // out-of-line: keeps the unrolled epilogue small
          }
          if (second && p.out2_bf16) {   // bf16 KV cache: same element offsets, 2-byte elements
            __nv_bfloat16* dst16 = reinterpret_cast<__nv_bfloat16*>(p.out2) + (dst - p.out2);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              if (m_first + 4 * i < M) {
                const __nv_bfloat162 lo2 = __floats2bfloat162_rn(v[i].x, v[i].y), hi2 = __floats2bfloat162_rn(v[i].z, v[i].w);
                uint2 pk;
                pk.x = *reinterpret_cast<const uint32_t*>(&lo2);
                pk.y = *reinterpret_cast<const uint32_t*>(&hi2);
                *reinterpret_cast<uint2*>(dst16 + i * dst_step) = pk;
              }
            }
          } else {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (m_first + 4 * i < M) *reinterpret_cast<float4*>(dst + i * dst_step) = v[i];
          }
          if (planes) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              if (m_first + 4 * i < M) {
                const float f[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
                uint32_t hw[2], lw[2];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                  const __nv_bfloat16 h0 = __float2bfloat16_rn(f[2 * u]), h1 = __float2bfloat16_rn(f[2 * u + 1]);
                  hw[u] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
                  const __nv_bfloat16 l0 = __float2bfloat16_rn(f[2 * u] - __bfloat162float(h0));
                  const __nv_bfloat16 l1 = __float2bfloat16_rn(f[2 * u + 1] - __bfloat162float(h1));
                  lw[u] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
                }
                const size_t o = (size_t)(m_first + 4 * i) * ldc + n;
                *reinterpret_cast<uint2*>(out_hi + o) = make_uint2(hw[0], hw[1]);
                if (out_lo) *reinterpret_cast<uint2*>(out_lo + o) = make_uint2(lw[0], lw[1]);
              }
            }
          }
        }
      }
      tc::tcgen05_before_sync();
      tc::mbar_arrive(tempty_bar(acc));
      if (dbg && threadIdx.x == 0) p.dbg[6] = tc::gtime();
    }
#undef TC_EPI_PREFETCH
  } else if (warp < TMA_WARP) {
    // =========================== A producers (implicit-GEMM gather) ===========================
    if constexpr (!A_TMA) {
    const int pt = threadIdx.x - TC_EPI_WARPS * 32;       // 0..255
    const int chunk = pt & 7;                             // 16-byte chunk within the 128-byte operand row
    const int rg = pt >> 3;                               // rows rg + 32*i
    constexpr int RPT = TC_ROWS_PER_THREAD;
    constexpr int LD4 = TF32 ? 1 : 2;                     // float4 loads per (row, chunk)
    int kit = 0;                                          // running k-block counter across tiles (ring position)
    for (int tile = blockIdx.x; tile < num_tiles && !(p.act & 64); tile += gridDim.x) {
      const int tm = tile / tiles_n;
      const float* base[RPT];
      int ih0[RPT], iw0[RPT];
      bool ok[RPT];
#pragma unroll
      for (int i = 0; i < RPT; ++i) {
        const int m = tm * TC_BM + rg + 32 * i;
        ok[i] = m < p.M;
        const int mm = ok[i] ? m : 0;
        const int ow = mm % p.OW;
        const int t = mm / p.OW;
        const int oh = t % p.OH;
        const int b = t / p.OH;
        ih0[i] = oh * p.SH - p.PH;
        iw0[i] = ow * p.SW - p.PW;
        base[i] = p.x + (size_t)b * p.H * p.W * p.C;
      }
      float4 v[2][RPT][LD4];
      auto load = [&](int kb, float4 (&dst)[RPT][LD4]) {
        const int k = kb * Cfg::KB_ELEMS + chunk * Cfg::CH_ELEMS;
        const bool kok = k < p.K;
        const int tap = kok ? k / p.C : 0;
        const int ci = k - tap * p.C;
        const int kh = tap / p.KW, kw = tap - kh * p.KW;
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
          const int ih = ih0[i] + kh, iw = iw0[i] + kw;
          const bool valid = kok && ok[i] && (unsigned)ih < (unsigned)p.H && (unsigned)iw < (unsigned)p.W;
          const float4* src = reinterpret_cast<const float4*>(base[i] + ((size_t)ih * p.W + iw) * p.C + ci);
#pragma unroll
          for (int q = 0; q < LD4; ++q) dst[i][q] = valid ? __ldg(src + q) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      };
      auto store = [&](int s, const float4 (&src)[RPT][LD4]) {
        uint8_t* a_hi = smem_gen + (size_t)s * Cfg::STAGE_BYTES;
        uint8_t* a_lo = a_hi + Cfg::A_BYTES;
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
          const int r = rg + 32 * i;
          const uint32_t off = (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u + (uint32_t)((chunk ^ (r & 7)) << 4);
          if constexpr (TF32) {
            const float4 x = src[i][0];
            float4 h;
            h.x = __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u);
            h.y = __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u);
            h.z = __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u);
            h.w = __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u);
            *reinterpret_cast<float4*>(a_hi + off) = h;
            if constexpr (PLANES == 2)
              *reinterpret_cast<float4*>(a_lo + off) = make_float4(x.x - h.x, x.y - h.y, x.z - h.z, x.w - h.w);
          } else {
            const float f[8] = {src[i][0].x, src[i][0].y, src[i][0].z, src[i][0].w,
                                src[i][1].x, src[i][1].y, src[i][1].z, src[i][1].w};
            uint32_t hw[4], lw[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const __nv_bfloat16 h0 = __float2bfloat16_rn(f[2 * q]), h1 = __float2bfloat16_rn(f[2 * q + 1]);
              hw[q] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
              if constexpr (PLANES == 2) {
                const __nv_bfloat16 l0 = __float2bfloat16_rn(f[2 * q] - __bfloat162float(h0));
                const __nv_bfloat16 l1 = __float2bfloat16_rn(f[2 * q + 1] - __bfloat162float(h1));
                lw[q] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
              }
            }
            *reinterpret_cast<uint4*>(a_hi + off) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
            if constexpr (PLANES == 2) *reinterpret_cast<uint4*>(a_lo + off) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
          }
        }
      };
      load(0, v[0]);
      for (int kb = 0; kb < nkb; kb += 2) {
        // unrolled by two so that the register double buffer is statically indexed
        if (kb + 1 < nkb) load(kb + 1, v[1]);
        {
          const int s = kit % STAGES;
          tc::mbar_wait(empty_bar(s), ((kit / STAGES) & 1) ^ 1);
          store(s, v[0]);
          tc::fence_proxy_async();
          tc::mbar_arrive(full_bar(s));
          ++kit;
        }
        if (kb + 1 < nkb) {
          if (kb + 2 < nkb) load(kb + 2, v[0]);
          const int s = kit % STAGES;
          tc::mbar_wait(empty_bar(s), ((kit / STAGES) & 1) ^ 1);
          store(s, v[1]);
          tc::fence_proxy_async();
          tc::mbar_arrive(full_bar(s));
          ++kit;
        }
      }
    }
    }  // !A_TMA
  } else if (warp == TMA_WARP) {
    // =========================== W producer (TMA) ===========================
    if (lane == 0) {
      int kit = 0;
      for (int tile = blockIdx.x; tile < num_tiles && !(p.act & 64); tile += gridDim.x) {
        const int ks = tile / tiles_mn, tmn = tile - ks * tiles_mn;
        const int tm = tmn / tiles_n, tn = tmn - tm * tiles_n;
        for (int kb = 0; kb < nkb; ++kb, ++kit) {
          const int s = kit % STAGES;
          const int kcoord = (ks * nkb + kb) * Cfg::KB_ELEMS;
          const bool pre = A_TMA && kit < w_prefetched;   // barrier armed and weight tile already in flight
          if (!pre) {
            tc::mbar_wait(empty_bar(s), ((kit / STAGES) & 1) ^ 1);
            tc::mbar_arrive_expect_tx(full_bar(s), PLANES * (Cfg::B_BYTES + (A_TMA ? Cfg::A_BYTES : 0)));
          }
          if constexpr (A_TMA) {
            const uint32_t a_hi = smem_base + s * Cfg::STAGE_BYTES;
            tc::tma_load_2d(a_hi, &map_a_hi, full_bar(s), kcoord, tm * TC_BM);
            if (PLANES == 2) tc::tma_load_2d(a_hi + Cfg::A_BYTES, &map_a_lo, full_bar(s), kcoord, tm * TC_BM);
          }
          if (pre) continue;
          const uint32_t b_hi = smem_base + s * Cfg::STAGE_BYTES + PLANES * Cfg::A_BYTES;
          tc::tma_load_2d(b_hi, &map_hi, full_bar(s), kcoord, tn * BN);
          if (PLANES == 2) tc::tma_load_2d(b_hi + Cfg::B_BYTES, &map_lo, full_bar(s), kcoord, tn * BN);
        }
      }
    }
    __syncwarp();
  } else {
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      constexpr uint32_t idesc = tc::make_idesc(TF32 ? 2 : 1, BN);
      int kit = 0, it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        tc::mbar_wait(tempty_bar(acc), ((it >> 1) & 1) ^ 1);
        tc::tcgen05_after_sync();
        const uint32_t d = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < nkb; ++kb, ++kit) {
          const int s = kit % STAGES;
          if (!(p.act & 64)) tc::mbar_wait(full_bar(s), (kit / STAGES) & 1);
          tc::tcgen05_after_sync();
          if (dbg && kb == 0) p.dbg[2] = tc::gtime();
          if (dbg && kb == nkb - 1) p.dbg[3] = tc::gtime();
          const uint32_t a_hi = smem_base + s * Cfg::STAGE_BYTES;
          const uint32_t b_hi = a_hi + PLANES * Cfg::A_BYTES;
          const uint64_t da_hi = tc::make_smem_desc(a_hi), db_hi = tc::make_smem_desc(b_hi);
          const uint64_t da_lo = tc::make_smem_desc(a_hi + Cfg::A_BYTES), db_lo = tc::make_smem_desc(b_hi + Cfg::B_BYTES);
          if constexpr (STACK) {
            constexpr uint32_t idesc2 = tc::make_idesc(1, 2 * BN);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              tc::umma<false>(d, da_hi + 2 * k, db_hi + 2 * k, idesc2, (kb | k) != 0);   // [hi.hi | hi.lo]
              tc::umma<false>(d, da_lo + 2 * k, db_hi + 2 * k, idesc, 1u);               // += lo.hi
            }
          } else {
#pragma unroll
          for (int k = 0; k < 4; ++k)  // 4 x (UMMA_K * elem) = 4 x 32 B = one 128-byte row
            tc::umma<TF32>(d, da_hi + 2 * k, db_hi + 2 * k, idesc, (kb | k) != 0);
          }
          if constexpr (PASSES == 3 && !STACK) {
#pragma unroll
            for (int k = 0; k < 4; ++k) tc::umma<TF32>(d, da_lo + 2 * k, db_hi + 2 * k, idesc, 1u);
#pragma unroll
            for (int k = 0; k < 4; ++k) tc::umma<TF32>(d, da_hi + 2 * k, db_lo + 2 * k, idesc, 1u);
          }
          tc::umma_commit(empty_bar(s));   // frees the smem stage once the MMAs above have read it
        }
        tc::umma_commit(tfull_bar(acc));    // accumulator complete -> epilogue
        if (dbg) p.dbg[4] = tc::gtime();
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
  if (dbg && threadIdx.x == 0) p.dbg[7] = tc::gtime();
}

// ------------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------------
__global__ void tc_split_weight_kernel(const float* __restrict__ w, void* __restrict__ hi, void* __restrict__ lo,
                                       long long n, int tf32) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float x = w[i];
    if (tf32) {
      const float h = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
      reinterpret_cast<float*>(hi)[i] = h;
      if (lo) reinterpret_cast<float*>(lo)[i] = x - h;
    } else {
      const __nv_bfloat16 h = __float2bfloat16_rn(x);
      reinterpret_cast<__nv_bfloat16*>(hi)[i] = h;
      if (lo) reinterpret_cast<__nv_bfloat16*>(lo)[i] = __float2bfloat16_rn(x - __bfloat162float(h));
    }
  }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled tc_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && p) fn = (PFN_encodeTiled)p;
  }
  return fn;
}

// Wide tiles (best operand reuse) when they still fill the machine, else narrower tiles for more CTAs.
inline int tc_pick_bn(int M, int N, int num_sms) {
  const int tiles_m = (M + TC_BM - 1) / TC_BM;
  int bn = N % 256 == 0 ? 256 : (N <= 64 ? 64 : 128);
  while (bn > 64 && tiles_m * ((N + bn - 1) / bn) < num_sms) bn >>= 1;
  return bn;
}

inline bool tc_supported(const ConvGemm& p) {
  // fp32 NHWC activations with 16-byte chunks inside one filter tap; KV-cache split outputs stay on the FFMA kernel
  return p.C % 8 == 0 && p.K % 8 == 0 && p.ldc % 4 == 0 && (p.res == nullptr || p.ldr % 4 == 0) && p.N % 4 == 0 &&
         (p.out2 == nullptr || (p.n_split % 32 == 0 && p.ldc2 % 4 == 0 && p.dyn_mul2 % 4 == 0));
}

inline cudaError_t tc_prepare_weight(const float* w_dev, int N, int K, int precision, TcWeight* out,
                                     std::vector<void*>* owned) {
  out->ready = false;
  if (precision == 0) return cudaSuccess;
  PFN_encodeTiled enc = tc_encode_fn();
  if (!enc) return cudaErrorNotSupported;
  if (K % 8 != 0) return cudaSuccess;  // unsupported shape: caller falls back to the FFMA kernel
  const bool tf32 = precision == 1;
  const bool two = precision != 3;
  const size_t esz = tf32 ? 4 : 2;
  const long long n = (long long)N * K;
  cudaError_t st;
  if ((st = cudaMalloc(&out->hi, n * esz)) != cudaSuccess) return st;
  owned->push_back(out->hi);
  out->lo = nullptr;
  if (two) {
    if ((st = cudaMalloc(&out->lo, n * esz)) != cudaSuccess) return st;
    owned->push_back(out->lo);
  }
  tc_split_weight_kernel<<<(int)((n + 255) / 256 > 4096 ? 4096 : (n + 255) / 256), 256>>>(w_dev, out->hi, out->lo, n, tf32 ? 1 : 0);
  if ((st = cudaGetLastError()) != cudaSuccess) return st;
  out->precision = precision; out->N = N; out->K = K;
  const cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)N};
  const cuuint64_t gstr[1] = {(cuuint64_t)K * esz};
  const cuuint32_t estr[2] = {1, 1};
  const CUtensorMapDataType dt = tf32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  for (int i = 0; i < 3; ++i) {
    const cuuint32_t box[2] = {(cuuint32_t)(tf32 ? 32 : 64), (cuuint32_t)(64 << i)};
    CUresult r = enc(&out->map_hi[i], dt, 2, out->hi, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return cudaErrorInvalidValue;
    out->map_lo[i] = out->map_hi[i];
    if (two) {
      r = enc(&out->map_lo[i], dt, 2, out->lo, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return cudaErrorInvalidValue;
    }
  }
  out->ready = true;
  return cudaSuccess;
}

// Tensor map over a bf16 activation plane [M, K] (row-major), box = 64 elements (128 B) x 128 rows, 128B swizzle.
inline cudaError_t tc_make_act_map(const void* plane, int M, int K, CUtensorMap* out) {
  PFN_encodeTiled enc = tc_encode_fn();
  if (!enc) return cudaErrorNotSupported;
  const cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)M};
  const cuuint64_t gstr[1] = {(cuuint64_t)K * 2};
  const cuuint32_t box[2] = {64, (cuuint32_t)TC_BM};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(plane), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

template <bool TF32, int PASSES, int BN, bool A_TMA, bool STACK = false>
inline cudaError_t tc_launch_one(const ConvGemm& p, const TcWeight& w, cudaStream_t s, int num_sms) {
  using Cfg = TcCfg<TF32, PASSES, BN, A_TMA>;
  static_assert(!STACK || (A_TMA && PASSES == 3 && !TF32 && BN <= 128), "stacked operand: bf16x3, one tile per CTA, N <= 256");
  static bool attr_set = false;
  auto kern = conv_gemm_tc_kernel<TF32, PASSES, BN, A_TMA, STACK>;
  if (!attr_set) {
    cudaError_t st = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM_BYTES);
    if (st != cudaSuccess) return st;
    attr_set = true;
  }
  const int tiles_m = (p.M + TC_BM - 1) / TC_BM, tiles_n = (p.N + BN - 1) / BN;
  const int tiles = tiles_m * tiles_n * ((A_TMA && p.k_splits > 1) ? p.k_splits : 1);
  const int grid = tiles < num_sms ? tiles : num_sms;
  constexpr int mi = BN == 64 ? 0 : (BN == 128 ? 1 : 2);
  const CUtensorMap* ah = A_TMA ? reinterpret_cast<const CUtensorMap*>(p.a_map_hi) : &w.map_hi[mi];
  const CUtensorMap* al = (A_TMA && p.a_map_lo) ? reinterpret_cast<const CUtensorMap*>(p.a_map_lo) : ah;
  return launch_kernel(kern, dim3(grid), dim3(tc_threads(A_TMA)), Cfg::SMEM_BYTES, s, p, w.map_hi[mi], w.map_lo[mi], *ah, *al,
                       tiles_m, tiles_n);
}

template <bool TF32, int PASSES>
inline cudaError_t tc_launch_bn(const ConvGemm& p, const TcWeight& w, cudaStream_t s, int num_sms) {
  if constexpr (!TF32) {
    // pre-split A planes + plain [M,K] operand -> TMA-fed A (decode-step GEMMs), one tile per CTA
    if (p.a_map_hi != nullptr && p.KH == 1 && p.KW == 1 && p.H == 1 && p.W == 1 && (PASSES == 1 || p.a_map_lo != nullptr)) {
      const int tiles_m = (p.M + TC_BM - 1) / TC_BM;
      const int ks = p.k_splits > 1 ? p.k_splits : 1;
      if constexpr (PASSES == 3) {
        if (p.stack) {
          if (tiles_m * ((p.N + 63) / 64) * ks <= num_sms) return tc_launch_one<false, 3, 64, true, true>(p, w, s, num_sms);
          if (ks == 1 && tiles_m * ((p.N + 127) / 128) <= num_sms) return tc_launch_one<false, 3, 128, true, true>(p, w, s, num_sms);
        }
      }
      if (tiles_m * ((p.N + 63) / 64) * ks <= num_sms) return tc_launch_one<false, PASSES, 64, true>(p, w, s, num_sms);
      if (ks == 1 && tiles_m * ((p.N + 127) / 128) <= num_sms) return tc_launch_one<false, PASSES, 128, true>(p, w, s, num_sms);
    }
  }
  if (p.k_splits > 1) return cudaErrorInvalidValue;   // split-K exists on the TMA-fed-A path only: the caller must not ask
  switch (tc_pick_bn(p.M, p.N, num_sms)) {
    case 256: return tc_launch_one<TF32, PASSES, 256, false>(p, w, s, num_sms);
    case 128: return tc_launch_one<TF32, PASSES, 128, false>(p, w, s, num_sms);
    case 64: return tc_launch_one<TF32, PASSES, 64, false>(p, w, s, num_sms);
  }
  return cudaErrorInvalidValue;
}

inline cudaError_t launch_conv_gemm_tc(const ConvGemm& p, const TcWeight& w, int precision, cudaStream_t s, int num_sms) {
  if (!w.ready || w.precision != precision) return cudaErrorInvalidValue;
  switch (precision) {
    case 1: return tc_launch_bn<true, 3>(p, w, s, num_sms);
    case 2: return tc_launch_bn<false, 3>(p, w, s, num_sms);
    case 3: return tc_launch_bn<false, 1>(p, w, s, num_sms);
  }
  return cudaErrorInvalidValue;
}

}  // namespace d2t

// CTA-pair (tcgen05 cta_group::2) stem convolution with BOTH operands delivered by TMA.
//
// Why a pair: on one SM the tensor core reads A (4 KB) and B (8 KB) from shared memory for every M=128 x N=256 x K=16
// instruction = 96 B/clk, and the operand fill writes as much again; shared-memory bandwidth is 128 B/clk, so a single-pass
// bf16 convolution cannot exceed ~2/3 of the tensor pipe (ncu: 44 %).  Two CTAs on the two SMs of a TPC execute ONE
// instruction with M = 256: CTA r owns output rows [256 tile + 128 r, +128) (its A tile, its TMEM accumulator rows) and HALF
// of the weight tile (rows [256 tn + 128 r, +128) of W); every operand byte is fetched and read once for both tensor cores:
// per SM 64 B/clk of reads + 64 B/clk of fill.
//
// Why TMA on both operands: round 1 built this pair with register-fed (gemm_tc2) and cp.async-fed (gemm_tc4) activations
// and both lost to the single-CTA kernel — the peer's "my A stage landed" had to be forwarded to the leader's barrier in
// software, a round trip longer than the stages buffer.  Here the activation tile arrives by TMA im2col (tc3_make_a_map)
// and the weight half by a tiled TMA load, both with .cta_group::2 completion straight on the LEADER's mbarrier: no thread
// of the peer takes part in the hand-off.
//
// Roles per CTA (192 threads): warps 0-3 epilogue (own 128 rows x 256 columns), warp 4 TMA thread, warp 5 MMA thread (leader).
//   full[s]    leader  2 expect_tx arrivals (one per CTA, the peer's remote) + the bytes of both CTAs' A and B tiles
//   empty[s]   both    tcgen05.commit multicast
//   tfull[a]   both    tcgen05.commit multicast
//   tempty[a]  leader  4 + 4 epilogue-warp arrivals (the peer's remote)
#pragma once
#include "gemm_tc3.cuh"

namespace d2t {
namespace tc {

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same smem offset in CTA 0 (the leader) of the pair
__device__ __forceinline__ uint32_t leader_addr(uint32_t local_addr) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(0u));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx_cluster(uint32_t cluster_addr, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cluster.b64 _, [%0], %1;" ::"r"(cluster_addr), "r"(bytes) : "memory");
}
// Waits on barriers that the other CTA of the pair arrives on use the ordinary (CTA-scope) try_wait: a .cluster-scope acquire
// costs an L1 invalidate (CCTL.IVALL) per poll and a .cluster-scope release on the arrive a memory barrier (ERRBAR) per
// k-block — ncu showed the first build of this kernel spending its time in exactly those (profiles/r02b_ncu_conv_tc5_*);
// what the barriers order here travels through the async proxy (TMA bytes, tcgen05 commits), not through generic loads.
// TMA loads into THIS CTA's shared memory whose transaction bytes are counted on the LEADER CTA's mbarrier
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(leader_addr(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_im2col_4d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c, int w, int h, int n,
                                                       uint16_t kw, uint16_t kh) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
      ::"r"(dst), "l"(map), "r"(leader_addr(bar)), "r"(c), "r"(w), "r"(h), "r"(n), "h"(kw), "h"(kh) : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar) {   // arrive on `bar` in BOTH CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void umma_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// fp32 accumulate, A/B K-major bf16, M = 256 (pair), N = n
__host__ __device__ constexpr uint32_t make_idesc_2sm(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}

}  // namespace tc

template <int PASSES>
struct Tc5Cfg {
  static constexpr int PLANES = PASSES == 1 ? 1 : 2;
  static constexpr int BN = 256;                       // columns per pair
  // 128-byte operand rows (64 channels of one filter tap, 128B swizzle): the TMA unit moves one ROW per request, so the
  // 64-byte rows of the single-CTA kernel cost twice the requests per byte (measured: encoder 21.5 ms with 32-element
  // k-blocks here)
  static constexpr int KB_ELEMS = 64;
  static constexpr int A_BYTES = TC_BM * 128;          // per plane: this CTA's 128 pixel rows
  static constexpr int B_BYTES = 128 * 128;            // per plane: this CTA's 128 weight rows
  static constexpr int STAGE_BYTES = PLANES * (A_BYTES + B_BYTES);
  static constexpr int EPI_WARPS = 4, THREADS = (EPI_WARPS + 2) * 32;
  static constexpr int EPI_STAGE_BYTES = EPI_WARPS * 32 * TC_EPI_PITCH * 4;
  static constexpr int STAGES_RAW = (225 * 1024 - EPI_STAGE_BYTES - 1280) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 10 ? 10 : STAGES_RAW;
  static constexpr int TMEM_COLS = 512;                // two 256-column accumulators
  static constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_BYTES + EPI_STAGE_BYTES + 1024 + 512;
};

template <int PASSES>
__global__ void __launch_bounds__(Tc5Cfg<PASSES>::THREADS, 1)
conv_gemm_tc5_kernel(const ConvGemm p, const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo,
                     const __grid_constant__ CUtensorMap amap_hi, const __grid_constant__ CUtensorMap amap_lo, int tiles_m2, int tiles_n) {
  using Cfg = Tc5Cfg<PASSES>;
  constexpr int STAGES = Cfg::STAGES, PLANES = Cfg::PLANES, BN = Cfg::BN;
  constexpr int EPI_WARPS = Cfg::EPI_WARPS, TMA_WARP = EPI_WARPS, MMA_WARP = EPI_WARPS + 1;
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
  const uint32_t rank = tc::cluster_ctarank();
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  const int num_tiles = tiles_m2 * tiles_n;
  const int nkb = p.K / Cfg::KB_ELEMS;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      tc::mbar_init(full_bar(s), 2);      // one expect_tx arrival per CTA of the pair (used on the leader only)
      tc::mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      tc::mbar_init(tfull_bar(a), 1);
      tc::mbar_init(tempty_bar(a), 2 * EPI_WARPS);
    }
    tc::fence_barrier_init();
  }
  if (warp == MMA_WARP) tc::tmem_alloc_2sm(tmem_slot, Cfg::TMEM_COLS);
  if (warp == TMA_WARP && lane == 0) {
    tc::tma_prefetch_desc(&map_hi);
    tc::tma_prefetch_desc(&amap_hi);
    if (PLANES == 2) { tc::tma_prefetch_desc(&map_lo); tc::tma_prefetch_desc(&amap_lo); }
  }
  tc::tcgen05_before_sync();
  tc::cluster_sync_all();
  tc::tcgen05_after_sync();
  const uint32_t tmem_base = *tmem_slot_gen;
  pdl_wait();
  pdl_trigger();

  if (warp < EPI_WARPS) {
    // =========================== epilogue: own 128 rows, all 256 columns ===========================
    const int quad = warp & 3;
    float* const stg = reinterpret_cast<float*>(smem_gen + (size_t)STAGES * Cfg::STAGE_BYTES) + warp * (32 * TC_EPI_PITCH);
    const int sub_r = lane >> 3, c4 = (lane & 7) * 4;
    const int M = p.M, N = p.N, ldc = p.ldc, ldr = p.ldr, act = p.act & 15;
    const float* const scale = p.scale;
    const float* const shift = p.shift;
    const float* const res = p.res;
    float* const out = p.out;
    __nv_bfloat16* const out_hi = p.out_hi;
    __nv_bfloat16* const out_lo = p.out_lo;
    int it = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs, ++it) {
      const int tm2 = tile / tiles_n, tn = tile - tm2 * tiles_n;
      const int acc = it & 1;
      tc::mbar_wait(tfull_bar(acc), (it >> 1) & 1);
      tc::tcgen05_after_sync();
      const int m_first = tm2 * 256 + (int)rank * 128 + quad * 32 + sub_r;
#pragma unroll 1
      for (int j = 0; j < BN / 32; ++j) {
        uint32_t r[32];
        tc::tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN + j * 32), r);
        tc::tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 8; ++q)
          *reinterpret_cast<uint4*>(stg + lane * TC_EPI_PITCH + q * 4) = make_uint4(r[q * 4], r[q * 4 + 1], r[q * 4 + 2], r[q * 4 + 3]);
        __syncwarp();
        const int n = tn * BN + j * 32 + c4;
        if (n < N) {
          float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
          if (scale) sc = __ldg(reinterpret_cast<const float4*>(scale + n));
          if (shift) sh = __ldg(reinterpret_cast<const float4*>(shift + n));
          float4 rr[8];
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
              const size_t o = (size_t)m * ldc + n;
              if (out) *reinterpret_cast<float4*>(out + o) = v;
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
      __syncwarp();
      if (lane == 0) tc::mbar_arrive_cluster(tc::leader_addr(tempty_bar(acc)));
    }
  } else if (warp == TMA_WARP) {
    // =========================== both operands by TMA: this CTA's 128 pixel rows and its half of the weight tile ===========================
    if (lane == 0) {
      int kit = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        const int tm2 = tile / tiles_n, tn = tile - tm2 * tiles_n;
        const int m0 = tm2 * 256 + (int)rank * 128;
        const int ow = m0 % p.OW, t = m0 / p.OW;
        const int aw = ow * p.SW - p.PW, ah = (t % p.OH) * p.SH - p.PH, an = t / p.OH;
        const bool have_a = an < p.B;      // the peer's half of the last tile may lie wholly past the last image
        const int n_row = tn * BN + (int)rank * 128;
        for (int kb = 0; kb < nkb; ++kb, ++kit) {
          const int s = kit % STAGES;
          tc::mbar_wait(empty_bar(s), ((kit / STAGES) & 1) ^ 1);
          tc::mbar_arrive_expect_tx_cluster(tc::leader_addr(full_bar(s)), PLANES * (Cfg::B_BYTES + (have_a ? Cfg::A_BYTES : 0)));
          const int k = kb * Cfg::KB_ELEMS;
          const int tap = k / p.C, ci = k - tap * p.C;
          const int kh = tap / p.KW, kw = tap - kh * p.KW;
          const uint32_t a_hi = smem_base + s * Cfg::STAGE_BYTES;
          const uint32_t b_hi = a_hi + PLANES * Cfg::A_BYTES;
          if (have_a) {
            tc::tma_load_im2col_4d_2sm(a_hi, &amap_hi, full_bar(s), ci, aw, ah, an, (uint16_t)kw, (uint16_t)kh);
            if (PLANES == 2) tc::tma_load_im2col_4d_2sm(a_hi + Cfg::A_BYTES, &amap_lo, full_bar(s), ci, aw, ah, an, (uint16_t)kw, (uint16_t)kh);
          }
          tc::tma_load_2d_2sm(b_hi, &map_hi, full_bar(s), k, n_row);
          if (PLANES == 2) tc::tma_load_2d_2sm(b_hi + Cfg::B_BYTES, &map_lo, full_bar(s), k, n_row);
        }
      }
    }
    __syncwarp();
  } else {
    if (lane == 0 && rank == 0) {
      // =========================== MMA issuer (leader CTA only) ===========================
      constexpr uint32_t idesc = tc::make_idesc_2sm(BN);
      int kit = 0, it = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs, ++it) {
        const int acc = it & 1;
        tc::mbar_wait(tempty_bar(acc), ((it >> 1) & 1) ^ 1);
        tc::tcgen05_after_sync();
        const uint32_t d = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < nkb; ++kb, ++kit) {
          const int s = kit % STAGES;
          tc::mbar_wait(full_bar(s), (kit / STAGES) & 1);
          tc::tcgen05_after_sync();
          const uint32_t a_hi = smem_base + s * Cfg::STAGE_BYTES;
          const uint32_t b_hi = a_hi + PLANES * Cfg::A_BYTES;
          const uint64_t da_hi = tc::make_smem_desc(a_hi), db_hi = tc::make_smem_desc(b_hi);
          const uint64_t da_lo = tc::make_smem_desc(a_hi + Cfg::A_BYTES), db_lo = tc::make_smem_desc(b_hi + Cfg::B_BYTES);
          // per 32-element half: hi.hi, lo.hi, hi.lo — the accumulation order of the single-CTA kernel's 32-element k-blocks,
          // so the two kernels round identically (the A/B tool asserts equal outputs)
#pragma unroll
          for (int h = 0; h < 2; ++h) {
#pragma unroll
            for (int k = 2 * h; k < 2 * h + 2; ++k) tc::umma_2sm(d, da_hi + 2 * k, db_hi + 2 * k, idesc, (kb | k) != 0);
            if constexpr (PASSES == 3) {
#pragma unroll
              for (int k = 2 * h; k < 2 * h + 2; ++k) tc::umma_2sm(d, da_lo + 2 * k, db_hi + 2 * k, idesc, 1u);
#pragma unroll
              for (int k = 2 * h; k < 2 * h + 2; ++k) tc::umma_2sm(d, da_hi + 2 * k, db_lo + 2 * k, idesc, 1u);
            }
          }
          tc::umma_commit_2sm(empty_bar(s));
        }
        tc::umma_commit_2sm(tfull_bar(acc));
      }
    }
    __syncwarp();
  }
  tc::tcgen05_before_sync();
  tc::cluster_sync_all();
  if (warp == MMA_WARP) {
    tc::tcgen05_after_sync();
    tc::tmem_dealloc_2sm(tmem_base, Cfg::TMEM_COLS);
  }
}

// host side ---------------------------------------------------------------------------------------------------------
inline bool tc5_supported(const ConvGemm& p, int precision, int num_sms) {
  if (!tc3_supported(p, precision) || !tc3_a_tma_supported(p) || p.N % 256 != 0 || p.C % 64 != 0 || p.out2 != nullptr) return false;
  // a GELU epilogue over a short K is bound by its erf evaluations (ncu: fc1 of the ViT blocks 96 us here with 4 epilogue warps
  // per CTA against 60 us on the gather kernel): such problems go to the single-CTA kernel's 128-wide tile (8 epilogue warps)
  if ((p.act & 15) == ACT_GELU && p.K <= 1024) return false;
  // enough pair tiles to fill (7/8 of) the machine in one wave: the ViT projections of a 256-image batch have 65
  return 8LL * ((p.M + 255) / 256) * (p.N / 256) >= 7LL * (num_sms / 2);
}

template <int PASSES>
inline cudaError_t tc5_launch(const ConvGemm& p, const TcWeight& w, cudaStream_t s, int num_sms) {
  using Cfg = Tc5Cfg<PASSES>;
  static bool attr_set = false;
  auto kern = conv_gemm_tc5_kernel<PASSES>;
  if (!attr_set) {
    cudaError_t st = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM_BYTES);
    if (st != cudaSuccess) return st;
    attr_set = true;
  }
  CUtensorMap ah, al;
  cudaError_t st = tc3_make_a_map(p.x_hi, p, &ah, 64);
  if (st != cudaSuccess) return st;
  al = ah;
  if (PASSES == 3 && (st = tc3_make_a_map(p.x_lo, p, &al, 64)) != cudaSuccess) return st;
  const int tiles_m2 = (p.M + 255) / 256, tiles_n = p.N / 256;
  int pairs = tiles_m2 * tiles_n;
  if (pairs > num_sms / 2) pairs = num_sms / 2;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * pairs); cfg.blockDim = dim3(Cfg::THREADS); cfg.dynamicSmemBytes = Cfg::SMEM_BYTES; cfg.stream = s;
  cudaLaunchAttribute attr[2];
  int na = 0;
  attr[na].id = cudaLaunchAttributeClusterDimension;
  attr[na].val.clusterDim.x = 2; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
  ++na;
  if (pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr; cfg.numAttrs = na;
  // weight maps with 128-row x 128-byte boxes (index 1): each CTA fetches its half of the 256-row tile
  return cudaLaunchKernelEx(&cfg, kern, p, w.map_hi[1], w.map_lo[1], ah, al, tiles_m2, tiles_n);
}

inline cudaError_t launch_conv_gemm_tc5(const ConvGemm& p, const TcWeight& w, int precision, cudaStream_t s, int num_sms) {
  return precision == 2 ? tc5_launch<3>(p, w, s, num_sms) : tc5_launch<1>(p, w, s, num_sms);
}

}  // namespace d2t

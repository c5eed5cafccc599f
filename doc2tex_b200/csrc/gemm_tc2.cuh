// CTA-pair (cta_group::2) variant of the tcgen05 implicit-GEMM contraction, for the large stem convolutions.
//
// Two CTAs on the two SMs of a TPC form a cluster and execute ONE tcgen05.mma with M = 256, N = 256: CTA r owns
// output rows [256*tile_m + 128*r, +128) (its A tile and its TMEM accumulator rows) and HALF of the weight tile
// (rows [256*tile_n + 128*r, +128) of W).  The tensor core reads A and B from both CTAs' shared memory, so each
// SM streams only 16 KB (A) + 16 KB (B) per operand plane and k-block instead of 16 + 32 KB: the per-SM L2->SM
// fill that caps the single-CTA kernel at 63 % tensor-pipe utilisation drops by a third, and the smaller stage
// (64 KB for the 3-pass mode) buys a third pipeline stage.
//
// Roles per CTA (448 threads) are those of conv_gemm_tc_kernel; only CTA 0 (the leader) issues MMAs.
//   full[s]   (leader only)  8+8 producer-warp arrivals + 2 TMA expect_tx arrivals, tx bytes from both CTAs
//   empty[s]  (both CTAs)    tcgen05.commit multicast to both
//   tfull[a]  (both CTAs)    tcgen05.commit multicast to both
//   tempty[a] (leader only)  4+4 epilogue-warp arrivals
#pragma once
#include "gemm_tc.cuh"

namespace d2t {
namespace tc {

// shared::cluster address of the same smem offset in CTA 0 (the leader) of the pair
__device__ __forceinline__ uint32_t leader_addr(uint32_t local_addr) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(0u));
  return r;
}

__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx_cluster(uint32_t cluster_addr, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.release.cluster.shared::cluster.b64 _, [%0], %1;" ::"r"(cluster_addr), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n.reg .pred p;\n"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n}\n"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok;
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait_cluster(bar, parity)) {}
}
// TMA load whose transaction bytes are counted on the LEADER CTA's mbarrier (same smem offset in CTA 0)
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(leader_addr(bar)), "r"(c0), "r"(c1) : "memory");
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
struct Tc2Cfg {
  static constexpr int PLANES = PASSES == 1 ? 1 : 2;
  static constexpr int BN = 256;                    // columns per PAIR; each CTA holds 128 weight rows
  static constexpr int KB_ELEMS = 64, CH_ELEMS = 8;
  static constexpr int A_BYTES = TC_BM * 128;       // per plane: this CTA's 128 activation rows
  static constexpr int B_BYTES = 128 * 128;         // per plane: this CTA's half of the weight tile
  static constexpr int STAGE_BYTES = PLANES * (A_BYTES + B_BYTES);
  static constexpr int EPI_STAGE_BYTES = 4 * 32 * TC_EPI_PITCH * 4;
  static constexpr int STAGES_RAW = (225 * 1024 - EPI_STAGE_BYTES - 1280) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 6 ? 6 : STAGES_RAW;
  static constexpr int TMEM_COLS = 512;             // two 256-column accumulators
  static constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_BYTES + EPI_STAGE_BYTES + 1024 + 256;
};

template <int PASSES>
__global__ void __launch_bounds__(448, 1)
conv_gemm_tc2_kernel(const ConvGemm p, const __grid_constant__ CUtensorMap map_hi,
                     const __grid_constant__ CUtensorMap map_lo, int tiles_m2, int tiles_n) {
  using Cfg = Tc2Cfg<PASSES>;
  constexpr int STAGES = Cfg::STAGES, PLANES = Cfg::PLANES, BN = Cfg::BN;
  constexpr int EPI_WARPS = 4, TMA_WARP = 12, MMA_WARP = 13;
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
  const uint32_t rank = tc::cluster_ctarank();        // 0 = leader
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  const int num_tiles = tiles_m2 * tiles_n;
  const int nkb = (p.K + Cfg::KB_ELEMS - 1) / Cfg::KB_ELEMS;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      tc::mbar_init(full_bar(s), 2 * TC_PROD_WARPS + 2);
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
    if (PLANES == 2) tc::tma_prefetch_desc(&map_lo);
  }
  tc::tcgen05_before_sync();
  tc::cluster_sync_all();          // barrier inits of BOTH CTAs are visible before any remote arrive
  tc::tcgen05_after_sync();
  const uint32_t tmem_base = *tmem_slot_gen;
  pdl_wait();
  pdl_trigger();

  if (warp < EPI_WARPS) {
    // =========================== epilogue (own 128 rows, all 256 columns of the pair's tile) ===========================
    const int quad = warp & 3;
    float* const stg = reinterpret_cast<float*>(smem_gen + (size_t)STAGES * Cfg::STAGE_BYTES) + warp * (32 * TC_EPI_PITCH);
    const int sub_r = lane >> 3, c4 = (lane & 7) * 4;
    const int M = p.M, N = p.N, ldc = p.ldc, ldr = p.ldr, act = p.act & 15;
    const float* const scale = p.scale;
    const float* const shift = p.shift;
    const float* const res = p.res;
    float* const out = p.out;
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
#pragma unroll 2
          for (int i = 0; i < 8; ++i) {
            const int m = m_first + 4 * i;
            if (m < M) {
              float4 v = *reinterpret_cast<const float4*>(stg + (sub_r + 4 * i) * TC_EPI_PITCH + c4);
              v.x = fmaf(v.x, sc.x, sh.x); v.y = fmaf(v.y, sc.y, sh.y); v.z = fmaf(v.z, sc.z, sh.z); v.w = fmaf(v.w, sc.w, sh.w);
              if (res) {
                const float4 rr = __ldg(reinterpret_cast<const float4*>(res + (size_t)m * ldr + n));
                v.x += rr.x; v.y += rr.y; v.z += rr.z; v.w += rr.w;
              }
              if (act == ACT_RELU) {
                v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
              } else if (act == ACT_GELU) {
                v = tc::gelu_erf4(v);
              }
              *reinterpret_cast<float4*>(out + (size_t)m * ldc + n) = v;
            }
          }
        }
        __syncwarp();
      }
      tc::tcgen05_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive_cluster(tc::leader_addr(tempty_bar(acc)));   // leader's barrier
    }
  } else if (warp < TMA_WARP) {
    // =========================== A producers: this CTA's 128 rows ===========================
    const int pt = threadIdx.x - EPI_WARPS * 32;
    const int chunk = pt & 7, rg = pt >> 3;
    constexpr int RPT = TC_ROWS_PER_THREAD;
    int kit = 0;
    for (int tile = pair; tile < num_tiles && !(p.act & 64); tile += num_pairs) {
      const int tm2 = tile / tiles_n;
      const float* base[RPT];
      int ih0[RPT], iw0[RPT];
      bool ok[RPT];
#pragma unroll
      for (int i = 0; i < RPT; ++i) {
        const int m = tm2 * 256 + (int)rank * 128 + rg + 32 * i;
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
      float4 v[2][RPT][2];
      auto load = [&](int kb, float4 (&dst)[RPT][2]) {
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
          for (int q = 0; q < 2; ++q) dst[i][q] = valid ? __ldg(src + q) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      };
      auto store = [&](int s, const float4 (&src)[RPT][2]) {
        uint8_t* a_hi = smem_gen + (size_t)s * Cfg::STAGE_BYTES;
        uint8_t* a_lo = a_hi + Cfg::A_BYTES;
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
          const int r = rg + 32 * i;
          const uint32_t off = (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u + (uint32_t)((chunk ^ (r & 7)) << 4);
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
      };
      auto publish = [&](int s) {   // this warp's part of stage s is in smem: make it visible to the tensor core, tell the leader
        tc::fence_proxy_async();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive_cluster(tc::leader_addr(full_bar(s)));
      };
      load(0, v[0]);
      for (int kb = 0; kb < nkb; kb += 2) {
        if (kb + 1 < nkb) load(kb + 1, v[1]);
        {
          const int s = kit % STAGES;
          tc::mbar_wait(empty_bar(s), ((kit / STAGES) & 1) ^ 1);
          store(s, v[0]);
          publish(s);
          ++kit;
        }
        if (kb + 1 < nkb) {
          if (kb + 2 < nkb) load(kb + 2, v[0]);
          const int s = kit % STAGES;
          tc::mbar_wait(empty_bar(s), ((kit / STAGES) & 1) ^ 1);
          store(s, v[1]);
          publish(s);
          ++kit;
        }
      }
    }
  } else if (warp == TMA_WARP) {
    // =========================== W producer: this CTA's 128 of the tile's 256 weight rows ===========================
    if (lane == 0) {
      int kit = 0;
      for (int tile = pair; tile < num_tiles && !(p.act & 64); tile += num_pairs) {
        const int tm2 = tile / tiles_n, tn = tile - tm2 * tiles_n;
        for (int kb = 0; kb < nkb; ++kb, ++kit) {
          const int s = kit % STAGES;
          tc::mbar_wait(empty_bar(s), ((kit / STAGES) & 1) ^ 1);
          tc::mbar_arrive_expect_tx_cluster(tc::leader_addr(full_bar(s)), PLANES * Cfg::B_BYTES);
          const uint32_t b_hi = smem_base + s * Cfg::STAGE_BYTES + PLANES * Cfg::A_BYTES;
          const int n_row = tn * BN + (int)rank * 128;
          tc::tma_load_2d_2sm(b_hi, &map_hi, full_bar(s), kb * Cfg::KB_ELEMS, n_row);
          if (PLANES == 2) tc::tma_load_2d_2sm(b_hi + Cfg::B_BYTES, &map_lo, full_bar(s), kb * Cfg::KB_ELEMS, n_row);
        }
      }
    }
    __syncwarp();
  } else {
    // =========================== MMA issuer: leader CTA only ===========================
    if (rank == 0 && lane == 0) {
      constexpr uint32_t idesc = tc::make_idesc_2sm(BN);
      int kit = 0, it = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs, ++it) {
        const int acc = it & 1;
        tc::mbar_wait_cluster(tempty_bar(acc), ((it >> 1) & 1) ^ 1);
        tc::tcgen05_after_sync();
        const uint32_t d = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < nkb; ++kb, ++kit) {
          const int s = kit % STAGES;
          if (!(p.act & 64)) tc::mbar_wait_cluster(full_bar(s), (kit / STAGES) & 1);
          tc::tcgen05_after_sync();
          const uint32_t a_hi = smem_base + s * Cfg::STAGE_BYTES;
          const uint32_t b_hi = a_hi + PLANES * Cfg::A_BYTES;
          const uint64_t da_hi = tc::make_smem_desc(a_hi), db_hi = tc::make_smem_desc(b_hi);
          const uint64_t da_lo = tc::make_smem_desc(a_hi + Cfg::A_BYTES), db_lo = tc::make_smem_desc(b_hi + Cfg::B_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k) tc::umma_2sm(d, da_hi + 2 * k, db_hi + 2 * k, idesc, (kb | k) != 0);
          if constexpr (PASSES == 3) {
#pragma unroll
            for (int k = 0; k < 4; ++k) tc::umma_2sm(d, da_lo + 2 * k, db_hi + 2 * k, idesc, 1u);
#pragma unroll
            for (int k = 0; k < 4; ++k) tc::umma_2sm(d, da_hi + 2 * k, db_lo + 2 * k, idesc, 1u);
          }
          tc::umma_commit_2sm(empty_bar(s));
        }
        tc::umma_commit_2sm(tfull_bar(acc));
      }
    }
    __syncwarp();
  }
  tc::tcgen05_before_sync();
  tc::cluster_sync_all();   // nobody leaves while the peer may still touch this CTA's smem / barriers / TMEM
  if (warp == MMA_WARP) {
    tc::tcgen05_after_sync();
    tc::tmem_dealloc_2sm(tmem_base, Cfg::TMEM_COLS);
  }
}

// host side ---------------------------------------------------------------------------------------------------------
inline bool tc2_supported(const ConvGemm& p, const TcWeight& w, int precision, int num_sms) {
  if (precision != 2 && precision != 3) return false;                 // bf16x3 / bf16
  if (p.N % 256 != 0 || p.out2 != nullptr || p.out_hi != nullptr || p.a_map_hi != nullptr || p.ln_w != nullptr) return false;
  if (!tc_supported(p) || !w.ready) return false;
  const int tiles = ((p.M + 255) / 256) * (p.N / 256);
  return tiles >= num_sms / 2;                                        // at least one tile per CTA pair
}

template <int PASSES>
inline cudaError_t tc2_launch(const ConvGemm& p, const TcWeight& w, cudaStream_t s, int num_sms) {
  using Cfg = Tc2Cfg<PASSES>;
  static bool attr_set = false;
  auto kern = conv_gemm_tc2_kernel<PASSES>;
  if (!attr_set) {
    cudaError_t st = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM_BYTES);
    if (st != cudaSuccess) return st;
    attr_set = true;
  }
  const int tiles_m2 = (p.M + 255) / 256, tiles_n = p.N / 256;
  int pairs = tiles_m2 * tiles_n;
  if (pairs > num_sms / 2) pairs = num_sms / 2;
  launch_cluster_x() = 2;
  // weight map with a 128-row box (index 1): each CTA loads its half of the 256-row tile
  return launch_kernel(kern, dim3(2 * pairs), dim3(448), Cfg::SMEM_BYTES, s, p, w.map_hi[1], w.map_lo[1], tiles_m2, tiles_n);
}

inline cudaError_t launch_conv_gemm_tc2(const ConvGemm& p, const TcWeight& w, int precision, cudaStream_t s, int num_sms) {
  return precision == 2 ? tc2_launch<3>(p, w, s, num_sms) : tc2_launch<1>(p, w, s, num_sms);
}

}  // namespace d2t

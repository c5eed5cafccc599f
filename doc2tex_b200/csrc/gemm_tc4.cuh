// CTA-pair (cta_group::2) stem-convolution kernel with the asynchronous A operand: the combination of
// conv_gemm_tc2_kernel (M = 256 over two SMs, each CTA holds its own 128 activation rows and HALF of the weight tile,
// leader-CTA MMA issue, multicast commits) and conv_gemm_tc3_kernel (bf16 hi/lo activation planes, 16-byte cp.async
// into 64B-swizzled stages of 32 bf16 per k-block).
//
// Why: on the single-CTA kernels shared-memory bandwidth is the binding resource (operand fill + tcgen05 operand
// reads > 128 B/clk per SM).  In a pair every operand byte is read from shared memory once for BOTH tensor cores, so
// per SM and k-block: fill 32 KB + reads 48 KB per 768 MMA cycles = 104 B/clk (3-pass), and a stage is only 32 KB
// -> six stages in flight.
//
// Barriers (s = stage, a = accumulator buffer):
//   full[s]    leader  256 cp.async arrivals (leader's producers) + 1 forwarded arrival (peer's A stage) +
//                      2 TMA expect_tx arrivals (one per CTA; the peer's is remote), tx bytes of both weight halves
//   afull[s]   peer    256 cp.async arrivals of the peer's producers; the peer's idle MMA warp forwards each phase to
//                      the leader's full[s]
//   empty[s]   both    tcgen05.commit multicast
//   tfull[a]   both    tcgen05.commit multicast
//   tempty[a]  leader  4 + 4 epilogue-warp arrivals (the peer's are remote)
#pragma once
#include "gemm_tc2.cuh"
#include "gemm_tc3.cuh"

namespace d2t {

template <int PASSES>
struct Tc4Cfg {
  static constexpr int PLANES = PASSES == 1 ? 1 : 2;
  static constexpr int BN = 256;                       // columns per pair
  static constexpr int KB_ELEMS = 32, CH_ELEMS = 8;
  static constexpr int A_BYTES = TC_BM * 64;           // per plane: this CTA's 128 rows
  static constexpr int B_BYTES = 128 * 64;             // per plane: this CTA's 128 weight rows
  static constexpr int STAGE_BYTES = PLANES * (A_BYTES + B_BYTES);
  static constexpr int EPI_STAGE_BYTES = 4 * 32 * TC_EPI_PITCH * 4;
  static constexpr int STAGES_RAW = (225 * 1024 - EPI_STAGE_BYTES - 1280) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int TMEM_COLS = 512;
  static constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_BYTES + EPI_STAGE_BYTES + 1024 + 512;
};

template <int PASSES>
__global__ void __launch_bounds__(448, 1)
conv_gemm_tc4_kernel(const ConvGemm p, const __grid_constant__ CUtensorMap map_hi,
                     const __grid_constant__ CUtensorMap map_lo, int tiles_m2, int tiles_n) {
  using Cfg = Tc4Cfg<PASSES>;
  constexpr int STAGES = Cfg::STAGES, PLANES = Cfg::PLANES, BN = Cfg::BN;
  constexpr int EPI_WARPS = 4, TMA_WARP = 12, MMA_WARP = 13;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - tc::smem_u32(smem_raw));
  const uint32_t bars = smem_base + STAGES * Cfg::STAGE_BYTES + Cfg::EPI_STAGE_BYTES;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };
  auto afull_bar = [&](int s) { return bars + 8u * (2 * STAGES + s); };
  auto tfull_bar = [&](int a) { return bars + 8u * (3 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bars + 8u * (3 * STAGES + 2 + a); };
  const uint32_t tmem_slot = bars + 8u * (3 * STAGES + 4);
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + STAGES * Cfg::STAGE_BYTES + Cfg::EPI_STAGE_BYTES + 8 * (3 * STAGES + 4));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = tc::cluster_ctarank();
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  const int num_tiles = tiles_m2 * tiles_n;
  const int nkb = (p.K + Cfg::KB_ELEMS - 1) / Cfg::KB_ELEMS;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      tc::mbar_init(full_bar(s), TC_PROD_THREADS + 1 + 2);
      tc::mbar_init(afull_bar(s), TC_PROD_THREADS);
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
              const size_t o = (size_t)m * ldc + n;
              *reinterpret_cast<float4*>(out + o) = v;
              if (out_hi) store_planes4(v, out_hi, out_lo, o);
            }
          }
        }
        __syncwarp();
      }
      tc::tcgen05_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive_cluster(tc::leader_addr(tempty_bar(acc)));
    }
  } else if (warp < TMA_WARP) {
    // =========================== A producers: cp.async, this CTA's 128 rows ===========================
    const int pt = threadIdx.x - EPI_WARPS * 32;
    const int chunk = pt & 3, rg = pt >> 2;
    const __nv_bfloat16* const xh = p.x_hi;
    const __nv_bfloat16* const xl = p.x_lo;
    const bool leader = rank == 0;
    int kit = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs) {
      const int tm2 = tile / tiles_n;
      long long base[2];
      int ih0[2], iw0[2];
      bool ok[2];
      uint32_t soff[2];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int r = rg + 64 * i;
        const int m = tm2 * 256 + (int)rank * 128 + r;
        ok[i] = m < p.M;
        const int mm = ok[i] ? m : 0;
        const int ow = mm % p.OW;
        const int t = mm / p.OW;
        const int oh = t % p.OH;
        const int b = t / p.OH;
        ih0[i] = oh * p.SH - p.PH;
        iw0[i] = ow * p.SW - p.PW;
        base[i] = (long long)b * p.H * p.W * p.C;
        soff[i] = (uint32_t)(r >> 3) * 512u + (uint32_t)(r & 7) * 64u + (uint32_t)((chunk ^ ((r >> 1) & 3)) << 4);
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
        for (int i = 0; i < 2; ++i) {
          const int ih = ih0[i] + kh, iw = iw0[i] + kw;
          const bool valid = kok && ok[i] && (unsigned)ih < (unsigned)p.H && (unsigned)iw < (unsigned)p.W;
          const long long e = valid ? base[i] + ((long long)ih * p.W + iw) * p.C + ci : 0;
          tc::cp_async_16(a_hi + soff[i], xh + e, valid ? 16u : 0u);
          if (PLANES == 2) tc::cp_async_16(a_hi + Cfg::A_BYTES + soff[i], xl + e, valid ? 16u : 0u);
        }
        tc::cp_async_mbar_arrive_noinc(leader ? full_bar(s) : afull_bar(s));
      }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
  } else if (warp == TMA_WARP) {
    // =========================== W producer: this CTA's half of the weight tile ===========================
    if (lane == 0) {
      int kit = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
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
    if (lane == 0) {
      if (rank == 0) {
        // =========================== MMA issuer (leader) ===========================
        constexpr uint32_t idesc = tc::make_idesc_2sm(BN);
        int kit = 0, it = 0;
        for (int tile = pair; tile < num_tiles; tile += num_pairs, ++it) {
          const int acc = it & 1;
          tc::mbar_wait_cluster(tempty_bar(acc), ((it >> 1) & 1) ^ 1);
          tc::tcgen05_after_sync();
          const uint32_t d = tmem_base + (uint32_t)(acc * BN);
          for (int kb = 0; kb < nkb; ++kb, ++kit) {
            const int s = kit % STAGES;
            tc::mbar_wait_cluster(full_bar(s), (kit / STAGES) & 1);
            tc::tcgen05_after_sync();
            const uint32_t a_hi = smem_base + s * Cfg::STAGE_BYTES;
            const uint32_t b_hi = a_hi + PLANES * Cfg::A_BYTES;
            const uint64_t da_hi = tc::make_smem_desc_sw64(a_hi), db_hi = tc::make_smem_desc_sw64(b_hi);
            const uint64_t da_lo = tc::make_smem_desc_sw64(a_hi + Cfg::A_BYTES), db_lo = tc::make_smem_desc_sw64(b_hi + Cfg::B_BYTES);
#pragma unroll
            for (int k = 0; k < 2; ++k) tc::umma_2sm(d, da_hi + 2 * k, db_hi + 2 * k, idesc, (kb | k) != 0);
            if constexpr (PASSES == 3) {
#pragma unroll
              for (int k = 0; k < 2; ++k) tc::umma_2sm(d, da_lo + 2 * k, db_hi + 2 * k, idesc, 1u);
#pragma unroll
              for (int k = 0; k < 2; ++k) tc::umma_2sm(d, da_hi + 2 * k, db_lo + 2 * k, idesc, 1u);
            }
            tc::umma_commit_2sm(empty_bar(s));
          }
          tc::umma_commit_2sm(tfull_bar(acc));
        }
      } else {
        // =========================== peer: forward "my A stage is in smem" to the leader ===========================
        int kit = 0;
        for (int tile = pair; tile < num_tiles; tile += num_pairs) {
          for (int kb = 0; kb < nkb; ++kb, ++kit) {
            const int s = kit % STAGES;
            tc::mbar_wait(afull_bar(s), (kit / STAGES) & 1);
            tc::mbar_arrive_cluster(tc::leader_addr(full_bar(s)));
          }
        }
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
inline bool tc4_supported(const ConvGemm& p, int precision, int num_sms) {
  if (!tc3_supported(p, precision) || p.N % 256 != 0) return false;
  return (long long)((p.M + 255) / 256) * (p.N / 256) >= num_sms / 2;
}

template <int PASSES>
inline cudaError_t tc4_launch(const ConvGemm& p, const Tc3Maps& m, cudaStream_t s, int num_sms) {
  using Cfg = Tc4Cfg<PASSES>;
  static bool attr_set = false;
  auto kern = conv_gemm_tc4_kernel<PASSES>;
  if (!attr_set) {
    cudaError_t st = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM_BYTES);
    if (st != cudaSuccess) return st;
    attr_set = true;
  }
  const int tiles_m2 = (p.M + 255) / 256, tiles_n = p.N / 256;
  int pairs = tiles_m2 * tiles_n;
  if (pairs > num_sms / 2) pairs = num_sms / 2;
  launch_cluster_x() = 2;
  return launch_kernel(kern, dim3(2 * pairs), dim3(448), Cfg::SMEM_BYTES, s, p, m.hi[1], m.lo[1], tiles_m2, tiles_n);
}

inline cudaError_t launch_conv_gemm_tc4(const ConvGemm& p, const Tc3Maps& m, int precision, cudaStream_t s, int num_sms) {
  return precision == 2 ? tc4_launch<3>(p, m, s, num_sms) : tc4_launch<1>(p, m, s, num_sms);
}

}  // namespace d2t

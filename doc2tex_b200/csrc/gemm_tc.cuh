// tcgen05 / TMEM / TMA implicit-GEMM contraction path (placeholder until the kernel lands).
#pragma once
#include <vector>
#include "common.cuh"

namespace d2t {

struct TcWeight {
  bool ready = false;
};

inline bool tc_supported(const ConvGemm&) { return false; }
inline cudaError_t tc_prepare_weight(const float*, int, int, int, TcWeight*, std::vector<void*>*) { return cudaSuccess; }
inline cudaError_t launch_conv_gemm_tc(const ConvGemm&, const TcWeight&, int, cudaStream_t, int) { return cudaErrorNotSupported; }

}  // namespace d2t

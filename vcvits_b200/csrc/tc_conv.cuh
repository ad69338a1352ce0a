// tcgen05 / TMEM / TMA implicit-GEMM convolution kernels (bf16 operands, fp32 accumulation).
#pragma once
#include <atomic>

#include "plan.h"

namespace vcd {

inline void tc_layer_eligibility(Layer& L) {
  L.tc_ok_fwd = L.tc_ok_dgr = L.tc_ok_wgr = false;
  L.nt_fwd = L.nt_dgr = 8;
}
inline int tc_plan_init(vcd_plan*) { return 0; }
inline int tc_run_conv(vcd_plan*, const Layer&, bool, const void*, int, int, int, int, const Epilogue&, cudaStream_t,
                       std::atomic<uint64_t>&, char*, size_t) { return 1; }
inline int tc_run_wgrad(vcd_plan*, const Layer&, const void*, const void*, float*, int, int, int, int, cudaStream_t,
                        std::atomic<uint64_t>&, char*, size_t) { return 1; }

}  // namespace vcd

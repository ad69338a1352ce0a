// Host side of the tcgen05 kernels: layer eligibility, tile configuration, TMA descriptors, launches.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>

#include "plan.h"
#include "tc_kernels.cuh"

namespace vcd {

// Debug switches (environment): VCD_TC_FWD / VCD_TC_DGRAD / VCD_TC_WGRAD = 0 force the FFMA kernels for that
// direction in bf16 mode; VCD_TC_MINLAYER / VCD_TC_MAXLAYER restrict the tensor-core path to a layer range.
inline int tc_env_int(const char* name, int dflt) {
  const char* s = getenv(name);
  return s && *s ? atoi(s) : dflt;
}

inline int tc_col_tile(int creal) {
  if (creal % 128 == 0) return 128;
  if (creal == 64 || creal == 32) return creal;
  return 0;
}

inline bool tc_geo_ok(const ConvGeo& g) {
  if (g.is != 1) return false;                      // strided gathers (ConvTranspose dgrad) stay on FFMA
  if (g.K % 16) return false;
  if (g.K > 64 && g.K % 64) return false;
  if (!tc_col_tile(g.creal)) return false;
  const int halo = (g.taps - 1) * (g.step < 0 ? -g.step : g.step);
  if (128 + halo > 256) return false;               // TMA box limit on the row dimension
  return true;
}

inline void tc_layer_eligibility(Layer& L) {
  static const int en_fwd = tc_env_int("VCD_TC_FWD", 1), en_dgr = tc_env_int("VCD_TC_DGRAD", 1),
                   en_wgr = tc_env_int("VCD_TC_WGRAD", 1);
  L.tc_ok_fwd = en_fwd && tc_geo_ok(L.fwd);
  L.tc_ok_dgr = en_dgr && tc_geo_ok(L.dgr);
  static const int en_m64 = tc_env_int("VCD_TC_WGRAD_M64", 1);
  {
    const ConvGeo& g = L.wgr;
    const int halo = (g.taps - 1) * (g.step < 0 ? -g.step : g.step);
    const bool shape_ok = g.is == 1 && g.os == 1 && g.p == 0 && g.creal == g.N && g.N % 16 == 0 &&
                          (g.N <= 256 || g.N % 256 == 0) && 64 + halo <= 256;
    const bool k_ok = (g.K % 128 == 0) || (en_m64 && (g.K == 64 || g.K == 32));
    L.tc_ok_wgr = en_wgr && shape_ok && k_ok;
  }
  L.nt_fwd = L.tc_ok_fwd ? tc_col_tile(L.fwd.creal) : 8;
  L.nt_dgr = L.tc_ok_dgr ? tc_col_tile(L.dgr.creal) : 8;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled tc_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

inline int tc_plan_init(vcd_plan* p) {
  // restrict the tensor-core path to a layer range (debug bisect)
  const int lo = tc_env_int("VCD_TC_MINLAYER", 0), hi = tc_env_int("VCD_TC_MAXLAYER", 1 << 30);
  (void)lo; (void)hi; (void)p;
  cudaError_t e = cudaFuncSetAttribute(tc::conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) return 1;
  e = cudaFuncSetAttribute(tc::wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) return 1;
  return tc_encode_fn() ? 0 : 1;
}

// bf16 blocked activation [B][C/8][L][8] viewed as a 4-D tensor (8, L, C/8, B); box = (8, rows, kchunks, 1).
inline bool tc_make_act_map(CUtensorMap* map, const void* base, int B, int C, int L, int box_rows, int box_chunks) {
  const cuuint64_t dims[4] = {8, static_cast<cuuint64_t>(L), static_cast<cuuint64_t>(C / 8), static_cast<cuuint64_t>(B)};
  const cuuint64_t strides[3] = {16, static_cast<cuuint64_t>(L) * 16, static_cast<cuuint64_t>(C / 8) * L * 16};
  const cuuint32_t box[4] = {8, static_cast<cuuint32_t>(box_rows), static_cast<cuuint32_t>(box_chunks), 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  return tc_encode_fn()(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

inline int tc_run_conv(vcd_plan* p, const Layer& L, bool dgrad, const void* in, int B, int Lin, int Lq, int Lout,
                       const Epilogue& e, cudaStream_t stream, std::atomic<uint64_t>& launches, char* err, size_t errn) {
  const ConvGeo& g = dgrad ? L.dgr : L.fwd;
  tc::ConvParams P{};
  P.g = g;
  P.e = e;
  P.w = p->d_bf16 + (dgrad ? L.tc_dgr : L.tc_fwd);
  P.B = B; P.Lin = Lin; P.Lq = Lq; P.Lout = Lout;
  P.BN = tc_col_tile(g.creal);
  P.KB = g.K < 64 ? g.K : 64;
  const int astep = g.step < 0 ? -g.step : g.step;
  P.RA = (128 + (g.taps - 1) * astep + 7) / 8 * 8;
  P.minshift = g.step < 0 ? (g.taps - 1) * g.step : 0;
  P.n_tiles_n = g.N / P.BN;
  const int mtiles = (Lq + 127) / 128;
  int MT = 1;
  for (int cand : {4, 2}) {
    if (2 * cand * P.BN > 512 || cand > mtiles) continue;
    const long long tiles = 1LL * ((mtiles + cand - 1) / cand) * P.n_tiles_n * B;
    if (tiles >= p->num_sms) { MT = cand; break; }
  }
  P.MT = MT;
  P.n_mgroups = (mtiles + MT - 1) / MT;
  P.total_tiles = P.n_mgroups * P.n_tiles_n * B;
  uint32_t cols = 32;
  while (cols < static_cast<uint32_t>(2 * MT * P.BN)) cols <<= 1;
  P.tmem_cols = cols;
  const size_t a_stage = static_cast<size_t>(MT) * (P.KB / 8) * P.RA * 16;
  const size_t w_tap = static_cast<size_t>(P.KB / 8) * P.BN * 16;
  const size_t w_all = w_tap * g.taps * (g.K / P.KB);
  const size_t budget = 220 * 1024;
  P.NA = (g.K / P.KB) > 1 ? 3 : 2;
  if (P.NA * a_stage > budget / 2) P.NA = 2;
  P.w_resident = (P.n_tiles_n == 1 && w_all <= 100 * 1024 && P.NA * a_stage + w_all <= budget) ? 1 : 0;
  size_t w_region;
  if (P.w_resident) {
    P.TPS = g.taps; P.NW = 1;
    w_region = w_all;
  } else {
    int tps = static_cast<int>((32 * 1024) / w_tap);
    if (tps < 1) tps = 1;
    if (tps > g.taps) tps = g.taps;
    int nw = static_cast<int>((budget - P.NA * a_stage) / (tps * w_tap));
    while (nw < 2 && tps > 1) { --tps; nw = static_cast<int>((budget - P.NA * a_stage) / (tps * w_tap)); }
    if (nw > 6) nw = 6;
    if (nw < 2) {
      snprintf(err, errn, "tc_run_conv(%s): no room for the weight ring", L.name.c_str());
      return 1;
    }
    P.TPS = tps; P.NW = nw;
    w_region = static_cast<size_t>(tps) * nw * w_tap;
  }
  const size_t smem = 128 + P.NA * a_stage + w_region + (2 * P.NA + 16 + 4) * 8 + 16;
  if (smem > 227 * 1024) {
    snprintf(err, errn, "tc_run_conv(%s): shared memory budget exceeded (%zu bytes)", L.name.c_str(), smem);
    return 1;
  }
  CUtensorMap tmA;
  if (!tc_make_act_map(&tmA, in, B, g.K, Lin, P.RA, P.KB / 8)) {
    snprintf(err, errn, "tc_run_conv(%s): cuTensorMapEncodeTiled failed", L.name.c_str());
    return 1;
  }
  const int grid = P.total_tiles < p->num_sms ? P.total_tiles : p->num_sms;
  tc::conv_kernel<<<grid, tc::kConvThreads, smem, stream>>>(tmA, P);
  launches.fetch_add(1, std::memory_order_relaxed);
  const cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) {
    snprintf(err, errn, "launch of tc::conv_kernel(%s) failed: %s", L.name.c_str(), cudaGetErrorString(ce));
    return 1;
  }
  return 0;
}

inline int tc_run_wgrad(vcd_plan* p, const Layer& L, const void* in, const void* dout, float* dwp, int B, int Lin,
                        int Ld, cudaStream_t stream, std::atomic<uint64_t>& launches, char* err, size_t errn) {
  const ConvGeo& g = L.wgr;
  tc::WgradParams P{};
  P.dwp = dwp;
  P.taps = g.taps; P.K = g.K; P.N = g.N; P.step = g.step; P.off0 = g.off0;
  P.minshift = g.step < 0 ? (g.taps - 1) * g.step : 0;
  P.B = B; P.L = Ld;
  if (g.K % 128 == 0) { P.G = 1; P.mch = 16; P.n_mtiles = g.K / 128; }
  else if (g.K == 64) { P.G = 2; P.mch = 8; P.n_mtiles = 1; }
  else { P.G = 4; P.mch = 4; P.n_mtiles = 1; }
  P.NT = g.N < 256 ? g.N : 256;
  P.n_ntiles = g.N / P.NT;
  P.n_slots = (g.taps + P.G - 1) / P.G;
  const int cap = 512 / P.NT;
  P.TG = cap < P.n_slots ? cap : P.n_slots;
  P.n_tgroups = (P.n_slots + P.TG - 1) / P.TG;
  uint32_t need = static_cast<uint32_t>(P.TG * P.NT), cols = 32;
  while (cols < need) cols <<= 1;
  P.tmem_cols = cols;
  const int astep = g.step < 0 ? -g.step : g.step;
  const int halo = (g.taps - 1) * astep;
  auto stage_bytes = [&](int tk) {
    const int ri = (tk + halo + 7) / 8 * 8;
    return static_cast<size_t>(P.mch) * ri * 16 * P.G + static_cast<size_t>(P.NT / 8) * tk * 16;
  };
  P.TK = (128 + halo <= 256 && 3 * stage_bytes(128) <= 200 * 1024) ? 128 : 64;
  P.RI = (P.TK + halo + 7) / 8 * 8;
  const size_t stage = stage_bytes(P.TK);
  int NS = static_cast<int>((200 * 1024) / stage);
  P.NS = NS > 6 ? 6 : (NS < 2 ? 2 : NS);
  const size_t smem = 128 + P.NS * stage + (2 * P.NS + 1) * 8 + 16;
  if (smem > 227 * 1024) {
    snprintf(err, errn, "tc_run_wgrad(%s): shared memory budget exceeded (%zu bytes)", L.name.c_str(), smem);
    return 1;
  }
  // splits of the (batch, time) contraction: enough CTAs for ~1/3 of the SMs per layer (independent layers run
  // concurrently on side streams), never more than one split per 8 time blocks
  static const int target_ctas = tc_env_int("VCD_WGRAD_CTAS", 48);
  const long long base_ctas = 1LL * P.n_mtiles * P.n_ntiles * P.n_tgroups;
  P.kb_per_item = (Ld + P.TK - 1) / P.TK;
  const long long total_kb = 1LL * B * P.kb_per_item;
  long long want = (target_ctas + base_ctas - 1) / base_ctas;
  const long long max_splits = (total_kb + 7) / 8;
  if (want > max_splits) want = max_splits;
  if (want < 1) want = 1;
  P.kb_per_split = static_cast<int>((total_kb + want - 1) / want);
  P.n_splits = static_cast<int>((total_kb + P.kb_per_split - 1) / P.kb_per_split);
  CUtensorMap tmIn, tmD;
  if (!tc_make_act_map(&tmIn, in, B, g.K, Lin, P.RI, P.mch) || !tc_make_act_map(&tmD, dout, B, g.N, Ld, P.TK, P.NT / 8)) {
    snprintf(err, errn, "tc_run_wgrad(%s): cuTensorMapEncodeTiled failed", L.name.c_str());
    return 1;
  }
  const long long grid = base_ctas * P.n_splits;
  tc::wgrad_kernel<<<static_cast<unsigned>(grid), tc::kThreads, smem, stream>>>(tmIn, tmD, P);
  launches.fetch_add(1, std::memory_order_relaxed);
  const cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) {
    snprintf(err, errn, "launch of tc::wgrad_kernel(%s) failed: %s", L.name.c_str(), cudaGetErrorString(ce));
    return 1;
  }
  return 0;
}

}  // namespace vcd

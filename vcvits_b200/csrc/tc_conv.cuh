// Host side of the tcgen05 kernels: layer eligibility, tile configuration, TMA descriptors, launches.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>

#include "plan.h"
#include <algorithm>

#include "tc_kernels.cuh"
#include "tc_pair.cuh"

namespace vcd {

// Debug switches: VCD_TC_FWD / VCD_TC_DGRAD / VCD_TC_WGRAD = 0 (environment, or vcd_debug_tc_paths() at run time for
// plans created afterwards) force the FFMA kernels for that direction in bf16 mode -- same bf16 operands, fp32 FFMA
// arithmetic: the on-device cross-check of the tcgen05 kernels.
inline int tc_env_int(const char* name, int dflt) {
  const char* s = getenv(name);
  return s && *s ? atoi(s) : dflt;
}
inline int* tc_path_flags() {   // {forward, data gradient, weight gradient}
  static int f[3] = {tc_env_int("VCD_TC_FWD", 1), tc_env_int("VCD_TC_DGRAD", 1), tc_env_int("VCD_TC_WGRAD", 1)};
  return f;
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
  const int left = g.off0 + (g.step < 0 ? (g.taps - 1) * g.step : 0);       // most negative row offset
  const int right = g.off0 + (g.step > 0 ? (g.taps - 1) * g.step : 0);     // most positive row offset
  if (-left > kPadL) return false;                  // halo must fit the zero pads of the row-padded layout
  if (127 + right + 8 > kPadR) return false;
  (void)halo;
  return true;
}

inline void tc_layer_eligibility(Layer& L) {
  const int en_fwd = tc_path_flags()[0], en_dgr = tc_path_flags()[1], en_wgr = tc_path_flags()[2];
  L.tc_ok_fwd = en_fwd && tc_geo_ok(L.fwd);
  L.tc_ok_dgr = en_dgr && tc_geo_ok(L.dgr);
  static const int en_m64 = tc_env_int("VCD_TC_WGRAD_M64", 1);
  {
    const ConvGeo& g = L.wgr;
    const int halo = (g.taps - 1) * (g.step < 0 ? -g.step : g.step);
    const int left = g.off0 + (g.step < 0 ? (g.taps - 1) * g.step : 0);
    const int right = g.off0 + (g.step > 0 ? (g.taps - 1) * g.step : 0);
    const bool shape_ok = g.is == 1 && g.os == 1 && g.p == 0 && g.creal == g.N && g.N % 16 == 0 &&
                          (g.N <= 128 || g.N % 128 == 0) && -left <= kPadL && 127 + right + 8 <= kPadR && (64 + halo + 7) / 8 * 8 <= 128;
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

// Row-padded blocked bf16 tensor [B][C/8][padded_len(L)][8] viewed as 8-byte elements:
// dims (2*padded_len(L), C/8, B), box (2*rows, groups, 1) -> shared memory [groups][rows][16 B].
inline bool tc_make_rows_map(CUtensorMap* map, const void* base, int B, int C, int L, int box_rows, int box_groups) {
  const cuuint64_t lp = static_cast<cuuint64_t>(padded_len(L));
  const cuuint64_t dims[3] = {2 * lp, static_cast<cuuint64_t>(C / 8), static_cast<cuuint64_t>(B)};
  const cuuint64_t strides[2] = {lp * 16, static_cast<cuuint64_t>(C / 8) * lp * 16};
  const cuuint32_t box[3] = {static_cast<cuuint32_t>(2 * box_rows), static_cast<cuuint32_t>(box_groups), 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  return tc_encode_fn() && tc_encode_fn()(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<void*>(base), dims, strides, box, estr,
                                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Epilogue-specialised instantiations of the convolution kernel.  P == nullptr: only raise the dynamic shared
// memory limit of instantiation `f` (plan creation); otherwise launch it.
template <int F, int UW = 16, bool LEAN = false, bool WG = false, bool CL = false>
inline cudaError_t tc_conv_launch_one(const tc::ConvParams* P, int grid, size_t smem, cudaStream_t stream, bool pdl) {
  if (!P) return cudaFuncSetAttribute(tc::conv_kernel<F, UW, LEAN, WG, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  const bool cluster = CL;
  // programmatic dependent launch: the prologue (barrier init, TMEM allocation, weight loads) overlaps the tail of the
  // previous kernel of the stream; the kernel's activation / epilogue-operand readers call griddepcontrol.wait
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(static_cast<unsigned>(grid));
  cfg.blockDim = dim3(tc::kConvThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (cluster) {   // pairs of CTAs sharing their weight stages by multicast (ConvParams::cl)
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 2;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, tc::conv_kernel<F, UW, LEAN, WG, CL>, *P);
}
// f: epilogue feature set (tc::EPI_*); bit 5 (32) selects the 32-column epilogue units
constexpr int kUw32 = 32, kLean = 64, kWg = 128;   // kWg: fused weight gradient (data-gradient launches, lean + mask staged)
constexpr int kCl = 256;                            // 2-CTA clusters with multicast weight stages (generic epilogues only)
inline cudaError_t tc_conv_dispatch(int f, const tc::ConvParams* P, int grid, size_t smem, cudaStream_t stream, bool pdl) {
  switch (f) {
    case kCl | 0: return tc_conv_launch_one<0, 16, false, false, true>(P, grid, smem, stream, pdl);
    case kCl | 1: return tc_conv_launch_one<1, 16, false, false, true>(P, grid, smem, stream, pdl);
    case kCl | 2: return tc_conv_launch_one<2, 16, false, false, true>(P, grid, smem, stream, pdl);
    case kCl | 3: return tc_conv_launch_one<3, 16, false, false, true>(P, grid, smem, stream, pdl);
    case kCl | 6: return tc_conv_launch_one<6, 16, false, false, true>(P, grid, smem, stream, pdl);
    case kCl | 7: return tc_conv_launch_one<7, 16, false, false, true>(P, grid, smem, stream, pdl);
    case kCl | 8: return tc_conv_launch_one<8, 16, false, false, true>(P, grid, smem, stream, pdl);
    case kCl | 14: return tc_conv_launch_one<14, 16, false, false, true>(P, grid, smem, stream, pdl);
    case kCl | 15: return tc_conv_launch_one<15, 16, false, false, true>(P, grid, smem, stream, pdl);
    case kWg | kLean | 17: return tc_conv_launch_one<17, 16, true, true>(P, grid, smem, stream, pdl);
    case kWg | kLean | 19: return tc_conv_launch_one<19, 16, true, true>(P, grid, smem, stream, pdl);
    case kWg | kLean | kUw32 | 17: return tc_conv_launch_one<17, 32, true, true>(P, grid, smem, stream, pdl);
    case kWg | kLean | kUw32 | 19: return tc_conv_launch_one<19, 32, true, true>(P, grid, smem, stream, pdl);
    case kLean | 0: return tc_conv_launch_one<0, 16, true>(P, grid, smem, stream, pdl);
    case kLean | 17: return tc_conv_launch_one<17, 16, true>(P, grid, smem, stream, pdl);
    case kLean | 18: return tc_conv_launch_one<18, 16, true>(P, grid, smem, stream, pdl);
    case kLean | 19: return tc_conv_launch_one<19, 16, true>(P, grid, smem, stream, pdl);
    case kLean | kUw32 | 0: return tc_conv_launch_one<0, 32, true>(P, grid, smem, stream, pdl);
    case kLean | kUw32 | 17: return tc_conv_launch_one<17, 32, true>(P, grid, smem, stream, pdl);
    case kLean | kUw32 | 18: return tc_conv_launch_one<18, 32, true>(P, grid, smem, stream, pdl);
    case kLean | kUw32 | 19: return tc_conv_launch_one<19, 32, true>(P, grid, smem, stream, pdl);
    case kUw32 | 0: return tc_conv_launch_one<0, 32>(P, grid, smem, stream, pdl);
    case kUw32 | 17: return tc_conv_launch_one<17, 32>(P, grid, smem, stream, pdl);
    case kUw32 | 18: return tc_conv_launch_one<18, 32>(P, grid, smem, stream, pdl);
    case kUw32 | 19: return tc_conv_launch_one<19, 32>(P, grid, smem, stream, pdl);
    case 0: return tc_conv_launch_one<0>(P, grid, smem, stream, pdl);
    case 1: return tc_conv_launch_one<1>(P, grid, smem, stream, pdl);
    case 2: return tc_conv_launch_one<2>(P, grid, smem, stream, pdl);
    case 3: return tc_conv_launch_one<3>(P, grid, smem, stream, pdl);
    case 6: return tc_conv_launch_one<6>(P, grid, smem, stream, pdl);
    case 7: return tc_conv_launch_one<7>(P, grid, smem, stream, pdl);
    case 8: return tc_conv_launch_one<8>(P, grid, smem, stream, pdl);
    case 14: return tc_conv_launch_one<14>(P, grid, smem, stream, pdl);
    case 15: return tc_conv_launch_one<15>(P, grid, smem, stream, pdl);
    case 17: return tc_conv_launch_one<17>(P, grid, smem, stream, pdl);   // EPI_SMEM variants of 1, 2, 3, 7, 15
    case 18: return tc_conv_launch_one<18>(P, grid, smem, stream, pdl);
    case 19: return tc_conv_launch_one<19>(P, grid, smem, stream, pdl);
    case 23: return tc_conv_launch_one<23>(P, grid, smem, stream, pdl);
    case 31: return tc_conv_launch_one<31>(P, grid, smem, stream, pdl);
    default: return P ? cudaErrorInvalidValue : cudaSuccess;
  }
}

inline int tc_plan_init(vcd_plan* p) {
  cudaError_t e = cudaSuccess;
  for (int f = 0; f < 512 && e == cudaSuccess; ++f) e = tc_conv_dispatch(f, nullptr, 0, 0, 0, false);
  if (e != cudaSuccess) return 1;
  e = cudaFuncSetAttribute(tc::wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) return 1;
  e = cudaFuncSetAttribute(tc::pair_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) return 1;
  e = cudaFuncSetAttribute(tc::pair_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) return 1;
  e = cudaFuncSetAttribute(tc::pair_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) return 1;
  if (getenv("VCD_KTRACE")) {
    if (cudaMalloc(&p->d_trace, 64 * sizeof(unsigned long long)) != cudaSuccess) return 1;
    cudaMemset(p->d_trace, 0, 64 * sizeof(unsigned long long));
  }
  return 0;
}

// Request to accumulate the layer's weight (+ bias) gradient inside its data-gradient launch (see ConvParams::wg_*).
struct WgFuse {
  float* dwp = nullptr;     // [taps][cin][cout] fp32, zeroed by the caller
  float* dbias = nullptr;   // [cout] or null
  bool fused = false;       // out: the launch took it (else the caller runs the weight-gradient kernel)
};

inline int tc_run_conv(vcd_plan* p, const Layer& L, bool dgrad, const void* in, int B, int Lin, int Lq, int Lout,
                       const Epilogue& e, cudaStream_t stream, std::atomic<uint64_t>& launches, char* err, size_t errn,
                       WgFuse* wg = nullptr) {
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
  // Rows per CTA tile (MT x 128) from a small cost model (clocks per CTA): the weight stream of a tile is paid once
  // per MT*128 rows, so larger MT trades CTA-level parallelism for less L2->SM traffic per MMA.
  //   t_mma  = MT * taps * K/16 * (128+BN)/4    (128 x BN x 16 UMMA, shared-memory operand fetch bound; tools/mma_rate.cu)
  //   t_load = bytes(A + streamed W) / 27 B/clk  (measured per-SM bulk-copy rate with ~160 KB in flight)
  //   t_epi  = MT * BN/16 units * 450 clk        (8 epilogue warps)
  const double w_tile_bytes = 2.0 * g.taps * g.K * P.BN;
  const bool can_reside = (P.n_tiles_n == 1 && w_tile_bytes <= 100 * 1024);
  // 2-CTA clusters with multicast weight stages for the streamed-weight launches (ConvParams::cl).  VCD_CONV_CLUSTER:
  // 0 (default) = off, 1 = whenever eligible, 2 = whenever eligible and the cost model below counts on the halved weight
  // traffic when it picks the rows per CTA tile, -1 = like 2 but only for long launches (>= 8 tiles per SM).  Bit-identical
  // results (tests/test_options_gpu.py).  Measured (DESIGN.md, round-2 notes): the >= 128-channel launches of the 10 s
  // inference batch run 5 % faster in isolation (938 -> 985 TFLOP/s), the concurrent step does not (71.8 -> 72.5 ms), and
  // at the training shapes a CTA owns one or two tiles, so the cluster start-up costs more than it saves (2.78 -> 2.80 ms).
  static const int cl_env = tc_env_int("VCD_CONV_CLUSTER", 0);
  const bool cl_long = 1LL * mtiles * P.n_tiles_n * B >= 8LL * p->num_sms;
  const bool cl_want = cl_env > 0 || (cl_env < 0 && cl_long);
  const bool cl_model = (cl_env >= 2 || cl_env < 0) && cl_want && !can_reside && (B % 2 == 0);
  int MT = 1, bufs = 2;
  double best = 1e30;
  for (int cand : {1, 2, 4}) {
    if (cand * P.BN > 512 || cand > mtiles) continue;
    const int cb = 2 * cand * P.BN <= 512 ? 2 : 1;
    const long long tiles = 1LL * ((mtiles + cand - 1) / cand) * P.n_tiles_n * B;
    // streamed-weight launches of a training step share the SMs with the other ResBlock branches and the trailing weight
    // gradients: plan their waves on a share of the machine (VCD_CONV_SHARE; 1 = the whole machine)
    static const int share = tc_env_int("VCD_CONV_SHARE", 1);
    const int eff_sms = can_reside || share <= 1 ? p->num_sms : (p->num_sms + share - 1) / share;
    const double waves = static_cast<double>((tiles + eff_sms - 1) / eff_sms);
    const double t_mma = 1.0 * cand * g.taps * (g.K / 16) * ((128 + P.BN) / 4);   // operand fetch at 128 B/clk
    const double a_bytes = 2.0 * cand * P.RA * g.K;
    const double t_load = (a_bytes + (can_reside ? 0.0 : w_tile_bytes / (cl_model ? 2.0 : 1.0))) / 27.0;
    const double t_epi = cand * (P.BN / 16) * 450.0;
    const double body = t_mma > t_load ? t_mma : t_load;
    const double per_tile = cb == 2 ? (body > t_epi ? body : t_epi) : body + t_epi;
    const double total = waves * per_tile + (cb == 2 ? t_epi : 0.0) + (can_reside ? w_tile_bytes / 27.0 : 0.0);
    if (total < best) { best = total; MT = cand; bufs = cb; }
  }
  {
    static const int force_mt = tc_env_int("VCD_CONV_MT", 0), force_mt_small = tc_env_int("VCD_CONV_MT_SMALL", 0);
    const int fm = can_reside && force_mt_small > 0 ? force_mt_small : force_mt;
    if (fm > 0 && fm * P.BN <= 512 && fm <= mtiles) { MT = fm; bufs = 2 * MT * P.BN <= 512 ? 2 : 1; }
  }
  // epilogue feature set -> instantiation (a superset is always valid: unused operands are null-checked or zero)
  int f = (e.mask ? tc::EPI_MASK : 0) | (e.res_t ? tc::EPI_RES : 0) | (e.res2 ? tc::EPI_RES2 : 0) | (e.out_raw ? tc::EPI_RAW : 0);
  if (f & (tc::EPI_RES2 | tc::EPI_RAW)) {
    if (f & tc::EPI_RES) f |= tc::EPI_RES2 | tc::EPI_RAW;      // running-sum variants: {RES,RES2,RAW} (+MASK)
  }
  if (f == (tc::EPI_RAW | tc::EPI_RES2)) f = tc::EPI_RES | tc::EPI_RES2 | tc::EPI_RAW;
  {
    const bool known = f == 0 || f == 1 || f == 2 || f == 3 || f == 6 || f == 7 || f == 8 || f == 14 || f == 15;
    if (!known) f = 15;
  }
  // EPI_SMEM candidates: bf16 operands in the output's own layout (plain convolution geometry), instantiations 1, 2, 3, 7, 15
  // Ring depths of the resident-weight (<= 64-channel) launches.  Short launches (a few tiles per CTA, operands in L2:
  // the training shapes) gain more from a small footprint -- a co-resident CTA of another ResBlock branch -- than from
  // depth; long launches (inference on seconds of audio: operands stream from HBM) need the bytes in flight.
  static const int esmem_on = tc_env_int("VCD_CONV_ESMEM", 1), ne_env = tc_env_int("VCD_CONV_NE", 0),
                   na_env = tc_env_int("VCD_CONV_NA_SMALL", 0);
  const bool long_launch = 1LL * ((mtiles + MT - 1) / MT) * P.n_tiles_n * B >= 32LL * p->num_sms;
  const int ne_want = ne_env > 0 ? ne_env : (long_launch ? 4 : 3);
  const int na_small = na_env > 0 ? na_env : (long_launch ? 8 : 3);
  const int e_nops = ((f & tc::EPI_MASK) && e.mask ? 1 : 0) + ((f & tc::EPI_RES) && e.res_t ? 1 : 0);
  const bool e_cand = esmem_on && e_nops > 0 && g.os == 1 && g.p == 0 && g.creal == g.N &&
                      (f == 1 || f == 2 || f == 3 || f == 7 || f == 15) &&
                      ((f & tc::EPI_MASK) == 0 || e.mask) && ((f & tc::EPI_RES) == 0 || e.res_t);
  P.NE = 0; P.e_ops = 0; P.e_stage_bytes = 0;
  const int kb0 = P.KB;
  size_t a_stage = 0, w_region = 0;
  // Streamed-weight (>= 128-channel) launches own one tile per CTA at training shapes, so a second accumulator buffer
  // buys nothing there; a CTA that keeps to half of the SM's shared memory and TMEM lets a CTA of another ResBlock
  // branch (or weight-gradient kernel) share the SM, whose epilogue / prologue then overlaps this one's MMAs.
  static const int bufs1_big = tc_env_int("VCD_CONV_BUFS1_BIG", 0);
  for (;; MT /= 2) {  // a row-tile count whose rings do not fit the shared-memory budget falls back to the next smaller one
  bufs = 2 * MT * P.BN <= 512 ? 2 : 1;
  if (bufs1_big && !can_reside && MT * P.BN <= 256) bufs = 1;
  P.KB = kb0;
  P.MT = MT;
  P.acc_bufs = bufs;
  P.n_mgroups = (mtiles + MT - 1) / MT;
  P.total_tiles = P.n_mgroups * P.n_tiles_n * B;
  P.d_tiles_n.init(P.n_tiles_n);
  P.d_mgroups.init(P.n_mgroups);
  P.d_creal.init(g.creal);
  P.units_shift = 0;
  while ((16 << P.units_shift) < P.BN) ++P.units_shift;
  uint32_t cols = 32;
  while (cols < static_cast<uint32_t>(bufs * MT * P.BN)) cols <<= 1;
  P.tmem_cols = cols;
  // keep two activation stages within ~96 KB so the weight ring keeps most of the shared memory
  // (streamed weights: keep them even smaller -- bytes in flight of the weight ring are what bounds those layers)
  static const int amax_kb = tc_env_int("VCD_CONV_AMAX_KB", 96);
  const size_t a_cap = (can_reside ? 96 : static_cast<size_t>(amax_kb)) * 1024;
  while (P.KB > 16 && 2 * static_cast<size_t>(MT) * (P.KB / 8) * P.RA * 16 > a_cap) P.KB /= 2;
  a_stage = static_cast<size_t>(MT) * (P.KB / 8) * P.RA * 16;
  const size_t w_tap = static_cast<size_t>(P.KB / 8) * P.BN * 16;
  const size_t w_all = w_tap * g.taps * (g.K / P.KB);
  static const int smem_kb = tc_env_int("VCD_CONV_SMEM_KB", 220), smem_kb_big = tc_env_int("VCD_CONV_SMEM_KB_BIG", 0);
  const size_t budget = static_cast<size_t>(!can_reside && smem_kb_big > 0 ? smem_kb_big : smem_kb) * 1024;
  P.NA = (g.K / P.KB) > 1 ? 3 : 2;
  {
    static const int force_na = tc_env_int("VCD_CONV_NA", 0);
    if (force_na > 0) P.NA = force_na;
  }
  if (P.NA * a_stage > budget / 2) P.NA = 2;
  P.w_resident = (P.n_tiles_n == 1 && w_all <= 100 * 1024 && P.NA * a_stage + w_all <= budget) ? 1 : 0;
  P.NE = 0; P.e_ops = 0; P.e_stage_bytes = 0;
  if (P.w_resident) {
    P.TPS = g.taps; P.NW = 1;
    w_region = w_all;
    // Small-channel layers are bound by the LATENCY of their loads (a 128-row tile of 32 channels is 8 KB): keep more
    // tiles in flight -- a deeper activation ring and the epilogue operands staged by the bulk-copy engine.
    size_t used = P.NA * a_stage + w_all;
    if (e_cand) {
      const size_t e_stage = static_cast<size_t>(e_nops) * MT * P.BN * 256;
      int ne = ne_want;
      while (ne >= 2 && used + ne * e_stage > budget) --ne;
      if (ne >= 2) { P.NE = ne > 8 ? 8 : ne; P.e_ops = e_nops; P.e_stage_bytes = static_cast<uint32_t>(e_stage); used += P.NE * e_stage; }
    }
    while (P.NA < na_small && used + a_stage <= budget && (P.NA + 1) * a_stage <= 96 * 1024) { ++P.NA; used += a_stage; }
  } else {
    if (P.NA * a_stage + 2 * w_tap > budget) {  // not even two one-tap weight stages beside the activation stages
      if (MT > 1) continue;
      snprintf(err, errn, "tc_run_conv(%s): no room for the weight ring", L.name.c_str());
      return 1;
    }
    int tps = static_cast<int>((32 * 1024) / w_tap);
    if (tps < 1) tps = 1;
    if (tps > g.taps) tps = g.taps;
    int nw = static_cast<int>((budget - P.NA * a_stage) / (tps * w_tap));
    while (nw < 2 && tps > 1) { --tps; nw = static_cast<int>((budget - P.NA * a_stage) / (tps * w_tap)); }
    if (nw > 6) nw = 6;
    if (nw < 2) {
      if (MT > 1) continue;
      snprintf(err, errn, "tc_run_conv(%s): no room for the weight ring", L.name.c_str());
      return 1;
    }
    P.TPS = tps; P.NW = nw;
    w_region = static_cast<size_t>(tps) * nw * w_tap;
  }
  break;
  }
  size_t smem = 128 + P.NA * a_stage + w_region + static_cast<size_t>(P.NE) * P.e_stage_bytes + (2 * P.NA + 16 + 4 + 16) * 8 + 16 + 2 * 128 * 4;
  if (P.NE > 0) f |= tc::EPI_SMEM;
  // 32-column epilogue units: resident-weight (<= 64-channel) layers whose epilogue has no global operands; both warps
  // of a TMEM lane quadrant need at least one unit per tile
  static const int uw32_on = tc_env_int("VCD_CONV_UW32", 1);
  const bool uw32 = uw32_on && P.w_resident && (f == 0 || f == 17 || f == 18 || f == 19) && P.BN % 32 == 0 && MT * P.BN >= 64;
  if (smem > 227 * 1024) {
    snprintf(err, errn, "tc_run_conv(%s): shared memory budget exceeded (%zu bytes)", L.name.c_str(), smem);
    return 1;
  }
  P.in = static_cast<const bf16*>(in);
  P.trace = nullptr;
  {
    static const char* want = getenv("VCD_KTRACE");  // e.g. "resblocks.5.convs1.0:fwd"
    if (want && p->d_trace && (L.name + (dgrad ? ":dgrad" : ":fwd")) == want) {
      P.trace = p->d_trace;
      fprintf(stderr, "[ktrace] %s%s bufs=%d BN=%d MT=%d KB=%d RA=%d NA=%d resident=%d TPS=%d NW=%d tiles=%d taps=%d K=%d\n", L.name.c_str(),
              dgrad ? ":dgrad" : ":fwd", P.acc_bufs, P.BN, P.MT, P.KB, P.RA, P.NA, P.w_resident, P.TPS, P.NW, P.total_tiles, g.taps, g.K);
    }
  }
  // two co-resident CTAs per SM when both fit (small-channel layers): one CTA's prologue / epilogue overlaps the
  // other's MMAs.  Needs <= ~110 KB shared memory, <= 256 TMEM columns and a <= 102-register instantiation.
  static const int occ2 = tc_env_int("VCD_CONV_OCC2", 0);
  const int ctas_per_sm = (occ2 && smem <= 110 * 1024 && P.tmem_cols <= 256) ? 2 : 1;
  const int max_ctas = p->num_sms * ctas_per_sm;
  int grid = P.total_tiles < max_ctas ? P.total_tiles : max_ctas;
  // 2-CTA clusters: streamed weights, generic epilogue with global operands, an even number of (batch item, row group)s
  P.cl = (cl_want && !P.w_resident && P.NE == 0 && !(wg != nullptr && wg->dwp != nullptr) &&
          (static_cast<long long>(B) * P.n_mgroups) % 2 == 0 && P.total_tiles >= 2) ? 1 : 0;
  if (P.cl) grid &= ~1;
  if (blk_elems(B, g.creal, Lout) >= (1ull << 31)) {  // the epilogue addresses its operands with 32-bit element offsets
    snprintf(err, errn, "tc_run_conv(%s): output tensor of %zu elements exceeds the 2^31-element limit of the tensor-core path",
             L.name.c_str(), blk_elems(B, g.creal, Lout));
    return 1;
  }
  if (!e.mask && e.scale != 1.f) { snprintf(err, errn, "tc_run_conv(%s): scale without mask is not supported", L.name.c_str()); return 1; }
  if ((f & tc::EPI_MASK) && !e.mask) { snprintf(err, errn, "tc_run_conv(%s): internal epilogue mismatch", L.name.c_str()); return 1; }
  if ((f & tc::EPI_RES) && !e.res_t) { snprintf(err, errn, "tc_run_conv(%s): internal epilogue mismatch", L.name.c_str()); return 1; }
  // VCD_PDL bit 0: forward launches, bit 1: data-gradient launches.  In the backward pass an early-started successor
  // holds shared memory / TMEM that the concurrent weight-gradient CTAs need (measured: +7 % on the segment), so
  // only the forward chain uses it by default.
  static const int pdl_mask = tc_env_int("VCD_PDL", 1), pdl_late_mask = tc_env_int("VCD_PDL_LATE", 2);
  P.pdl_late = (pdl_late_mask & (dgrad ? 2 : 1)) != 0 ? 1 : 0;
  // lean epilogue: plain geometry, launch-constant bias, operand-free or smem-staged epilogue, both warps of a quadrant busy
  static const int lean_on = tc_env_int("VCD_CONV_LEAN", 1);
  const bool lean = lean_on && P.w_resident && (f == 0 || f == 17 || f == 18 || f == 19) && g.os == 1 && g.p == 0 && g.creal == g.N &&
                    P.n_tiles_n == 1 && e.bias2 == nullptr && e.zu == 0 && (MT * P.BN) / (uw32 ? 32 : 16) >= 2;
  // Fused weight gradient (ConvParams::wg_*): lean data-gradient launch whose staged mask operand IS the layer's forward
  // input, one 128-row tile per CTA tile, one K block, square <= 64-channel layer, odd tap count (free slot for the bias).
  bool wg_on = false;
  P.wg_dwp = nullptr; P.wg_dbias = nullptr; P.wg_r0 = P.wg_rstep = P.wg_rc = 0; P.wg_col0 = 0;
  static const int wg_env = tc_env_int("VCD_WG_FUSE", 0);   // off: measured slower (see DESIGN.md, round-2 negative results)
  if (wg != nullptr) wg->fused = false;
  if (wg != nullptr && wg->dwp != nullptr && wg_env && !p->deterministic && dgrad && lean && (f & tc::EPI_MASK) && (f & tc::EPI_SMEM) && MT == 1 && P.KB == g.K &&
      g.K == g.N && g.N == P.BN && g.K <= 64 && (g.taps & 1) && L.tc_ok_wgr) {
    const ConvGeo& gw = L.wgr;
    const int r0 = -gw.off0 - g.off0 - P.minshift, rstep = -gw.step, rc = -g.off0 - P.minshift;
    bool ok = gw.taps == g.taps && gw.K == g.N && gw.N == g.K && gw.is == 1 && gw.os == 1 && rc >= 0 && rc + 128 <= P.RA;
    for (int j = 0; ok && j < g.taps; ++j) ok = r0 + j * rstep >= 0 && r0 + j * rstep + 128 <= P.RA;
    const uint32_t col0 = static_cast<uint32_t>(P.acc_bufs * MT * P.BN);
    const uint32_t need = col0 + static_cast<uint32_t>(((g.taps + 2) / 2) * P.BN);
    ok = ok && need <= 512 && P.NA * a_stage + w_region >= 8 * 32 * 36 * sizeof(float);   // (write-out scratch aliases the rings)
    const size_t extra = 2048 + 128 + (g.K < 64 ? 8 * 1024 : 0);   // ones tile; slack: the M = 64 operand over a 32-channel tile
    ok = ok && smem + extra <= 227 * 1024;
    if (ok) {
      wg_on = true;
      smem += extra;
      P.wg_dwp = wg->dwp; P.wg_dbias = wg->dbias;
      P.wg_r0 = r0; P.wg_rstep = rstep; P.wg_rc = rc; P.wg_col0 = col0;
      uint32_t cols = 32;
      while (cols < need) cols <<= 1;
      P.tmem_cols = cols;
      wg->fused = true;
    }
  }
  if (P.cl && (uw32 || lean || wg_on || (f & tc::EPI_SMEM))) { P.cl = 0; grid = P.total_tiles < max_ctas ? P.total_tiles : max_ctas; }
  const cudaError_t ce = tc_conv_dispatch(f | (uw32 ? kUw32 : 0) | (lean ? kLean : 0) | (wg_on ? kWg : 0) | (P.cl ? kCl : 0), &P, grid, smem, stream,
                                          (pdl_mask & (dgrad ? 2 : 1)) != 0);
  launches.fetch_add(1, std::memory_order_relaxed);
  if (ce != cudaSuccess) {
    snprintf(err, errn, "launch of tc::conv_kernel(%s) failed: %s", L.name.c_str(), cudaGetErrorString(ce));
    return 1;
  }
  return 0;
}

// Shared-memory bytes of the fused pair kernel: phase-A halo hA (input region), phase-B tap over-read `slack` (mid region),
// an NA-deep activation ring and MT 128-row tiles per CTA tile.
inline size_t tc_pair_smem(int C, int taps, int hA, int slack, int NA, int MT) {
  const int RA = (MT * 128 + 2 * hA + 7) / 8 * 8;
  const size_t a_stage = static_cast<size_t>(C / 8) * RA * 16, w = static_cast<size_t>(taps) * (C / 8) * C * 16;
  const size_t mid = static_cast<size_t>(C / 8) * (MT * 128 + slack) * 16;
  return 128 + NA * a_stage + 2 * w + 2 * mid + 32 * 8 + 16 + 2 * 64 * 4 + 128;
}
inline int tc_pair_slack(int hB) { const int s = (2 * hB + 7) / 8 * 8; return s < 16 ? 16 : s; }

// Can the pair (L1 = dilated c1, L2 = c2 with dilation 1) of a ResBlock1 run as ONE fused launch?  bwd: the two data
// gradients (phase A = c2's, phase B = c1's) instead of the two forward convolutions.
inline bool tc_pair_ok(const Layer& L1, const Layer& L2, bool bwd = false) {
  static const int on = tc_env_int("VCD_PAIR", 1), on_bwd = tc_env_int("VCD_PAIR_BWD", 1);
  if (!on || (bwd && !on_bwd)) return false;
  if (bwd ? (!L1.tc_ok_dgr || !L2.tc_ok_dgr) : (!L1.tc_ok_fwd || !L2.tc_ok_fwd)) return false;
  if (L1.kind != LK_CONV || L2.kind != LK_CONV) return false;
  const int C = L1.cin;
  if (!(C == 32 || C == 64) || L1.cout != C || L2.cin != C || L2.cout != C) return false;
  if (L1.k != L2.k || !(L1.k & 1) || L1.k < 3 || L1.k > 11 || L2.dil != 1 || L1.dil < 1) return false;
  if (L1.p_b < 0 || L2.p_b < 0) return false;
  const int h1 = L1.dil * (L1.k - 1) / 2, h2 = (L1.k - 1) / 2;
  if (h1 + h2 > kPadL || 128 + h1 + h2 + 8 > kPadR) return false;   // the tile halo must stay inside the zero pads
  if (bwd ? (L1.nt_dgr != C || L2.nt_dgr != C) : (L1.nt_fwd != C || L2.nt_fwd != C)) return false;   // one column tile: [tap][C/8][C][8]
  const int hA = bwd ? h2 : h1, hB = bwd ? h1 : h2;
  return tc_pair_smem(C, L1.k, hA, tc_pair_slack(hB), 2, 1) <= 227 * 1024;
}

// One fused pair launch.  Forward: out / out_raw / res2 / tscale / out_slope as in tc::PairParams (a non-final pair stores
// lrelu(.) with the ResBlock slope; the final pair of a branch joins the fp32 running sum over the branches).  Backward
// (bwd): in = gradient w.r.t. the pair's output, mask1 = stored lrelu(c1 out), mask2 = stored lrelu(pair input),
// mid_out = gradient w.r.t. c1's output (read by c1's weight gradient), out = gradient w.r.t. the pair's input.
struct PairCall {
  const void* in = nullptr;
  void* mid_out = nullptr;
  void* out = nullptr;
  const float* res2 = nullptr;
  float* out_raw = nullptr;
  float tscale = 1.f, out_slope = 1.f;
  const float* bias1 = nullptr;
  const float* bias2 = nullptr;
  float act_slope = 1.f, res_inv = 1.f;
  bool bwd = false;
  const void* mask1 = nullptr;
  const void* mask2 = nullptr;
  float mask_slope = 1.f;
};

inline int tc_run_pair(vcd_plan* p, const Layer& L1, const Layer& L2, const PairCall& a, int B, int Lrows, cudaStream_t stream,
                       std::atomic<uint64_t>& launches, char* err, size_t errn) {
  tc::PairParams P{};
  P.in = static_cast<const bf16*>(a.in);
  // phase A / phase B weights: forward c1, c2; backward c2's and c1's data-gradient formats
  P.w1 = p->d_bf16 + (a.bwd ? L2.tc_dgr : L1.tc_fwd);
  P.w2 = p->d_bf16 + (a.bwd ? L1.tc_dgr : L2.tc_fwd);
  P.bias1 = a.bias1; P.bias2 = a.bias2;
  P.mid_out = static_cast<bf16*>(a.mid_out);
  P.out = static_cast<bf16*>(a.out);
  P.res2 = a.res2; P.out_raw = a.out_raw; P.tscale = a.tscale; P.out_slope = a.out_slope;
  P.mask1 = static_cast<const bf16*>(a.mask1); P.mask2 = static_cast<const bf16*>(a.mask2); P.mask_slope = a.mask_slope;
  P.B = B; P.L = Lrows; P.C = L1.cin; P.taps = L1.k;
  const int h1 = L1.dil * (L1.k - 1) / 2, h2 = (L1.k - 1) / 2;
  P.dilA = a.bwd ? 1 : L1.dil; P.dilB = a.bwd ? L1.dil : 1;
  P.hA = a.bwd ? h2 : h1; P.hB = a.bwd ? h1 : h2;
  P.mid_slack = tc_pair_slack(P.hB);
  if (a.bwd && (!a.mid_out || !a.out || !a.mask1 || !a.mask2)) { snprintf(err, errn, "tc_run_pair(%s): backward pair needs dm, out and both masks", L1.name.c_str()); return 1; }
  // MT 128-row MMA tiles per CTA tile: the largest that fits TMEM (4 * MT * C columns) and shared memory (two
  // activation stages at least) while every SM still gets a few CTA tiles
  static const int mt_env = tc_env_int("VCD_PAIR_MT", 0), na_want = tc_env_int("VCD_PAIR_NA", 3);
  P.MT = 1;
  for (int mt : {4, 2}) {
    if (mt_env > 0 && mt != mt_env) continue;
    if (4 * mt * P.C > 512 || tc_pair_smem(P.C, P.taps, P.hA, P.mid_slack, 2, mt) > 220 * 1024) continue;
    const int r = mt * 128 - 2 * P.hB;
    const long long tiles = 1LL * B * ((Lrows + r - 1) / r);
    if (mt_env == 0 && tiles < 3LL * p->num_sms) continue;
    P.MT = mt;
    break;
  }
  P.R = P.MT * 128 - 2 * P.hB;
  P.RA = (P.MT * 128 + 2 * P.hA + 7) / 8 * 8;
  P.NA = na_want < 2 ? 2 : (na_want > 8 ? 8 : na_want);
  while (P.NA > 2 && tc_pair_smem(P.C, P.taps, P.hA, P.mid_slack, P.NA, P.MT) > 220 * 1024) --P.NA;
  const size_t smem = tc_pair_smem(P.C, P.taps, P.hA, P.mid_slack, P.NA, P.MT);
  if (smem > 227 * 1024) { snprintf(err, errn, "tc_run_pair(%s): shared memory budget exceeded", L1.name.c_str()); return 1; }
  P.tiles_per_item = (Lrows + P.R - 1) / P.R;
  P.total_tiles = P.tiles_per_item * B;
  P.d_tiles.init(P.tiles_per_item);
  P.act_slope = a.act_slope; P.res_inv = a.res_inv;
  uint32_t cols = 32;
  while (cols < static_cast<uint32_t>(4 * P.MT * P.C)) cols <<= 1;
  P.tmem_cols = cols;
  if (blk_elems(B, P.C, Lrows) >= (1ull << 31)) { snprintf(err, errn, "tc_run_pair(%s): tensor too large", L1.name.c_str()); return 1; }
  const int grid = P.total_tiles < p->num_sms ? P.total_tiles : p->num_sms;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(static_cast<unsigned>(grid));
  cfg.blockDim = dim3(tc::kPairThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  static const int pdl_mask = tc_env_int("VCD_PDL", 1);
  cfg.numAttrs = (pdl_mask & (a.bwd ? 2 : 1)) ? 1 : 0;
  cudaError_t ce;
  if (a.bwd) ce = cudaLaunchKernelEx(&cfg, tc::pair_kernel<true, true>, P);
  else ce = a.mid_out ? cudaLaunchKernelEx(&cfg, tc::pair_kernel<true, false>, P) : cudaLaunchKernelEx(&cfg, tc::pair_kernel<false, false>, P);
  launches.fetch_add(1, std::memory_order_relaxed);
  if (ce != cudaSuccess) {
    snprintf(err, errn, "launch of tc::pair_kernel(%s) failed: %s", L1.name.c_str(), cudaGetErrorString(ce));
    return 1;
  }
  return 0;
}

inline int tc_run_wgrad(vcd_plan* p, const Layer& L, const void* in, const void* dout, float* dwp, float* dbias, int B, int Lin,
                        int Ld, bool tail, cudaStream_t stream, std::atomic<uint64_t>& launches, char* err, size_t errn) {
  const ConvGeo& g = L.wgr;
  tc::WgradParams P{};
  P.dwp = dwp;
  P.taps = g.taps; P.K = g.K; P.N = g.N; P.step = g.step; P.off0 = g.off0;
  P.minshift = g.step < 0 ? (g.taps - 1) * g.step : 0;
  P.B = B; P.L = Ld;
  P.NT = g.N <= 128 ? g.N : 128;
  P.n_ntiles = g.N / P.NT;
  const int astep = g.step < 0 ? -g.step : g.step;
  const int halo = (g.taps - 1) * astep;
  // (M, G): instruction M and the number of tap copies stacked along it.  Small-N MMAs are bound by the operand
  // reads from shared memory (~(M + NT)/4 clk), the stage load by ~27 B/clk: pick the cheaper bottleneck.
  struct Cand { int M, G; };
  Cand cands[5];
  int nc = 0;
  if (g.K % 128 == 0) cands[nc++] = {128, 1};
  else if (g.K == 64) { cands[nc++] = {64, 1}; cands[nc++] = {128, 2}; }
  else { cands[nc++] = {64, 1}; cands[nc++] = {64, 2}; cands[nc++] = {128, 4}; }
  double best = 1e30;
  for (int i = 0; i < nc; ++i) {
    const int M = cands[i].M, G = cands[i].G;
    const int mch = (g.K < 128 ? g.K : 128) / 8;
    const double slots = (g.taps + G - 1) / G;
    // per 128 rows: an MMA costs max(issue ~55 clk, operand fetch (M + NT)/4 clk, NT/2 clk); loads run at ~22 B/clk
    double per_mma = (M + P.NT) / 4.0;
    if (per_mma < P.NT / 2.0) per_mma = P.NT / 2.0;
    static const int issue_floor = tc_env_int("VCD_WGRAD_ISSUE_FLOOR", 55);
    if (per_mma < issue_floor) per_mma = issue_floor;
    const double mma = slots * 8.0 * per_mma;
    static const double load_rate = tc_env_int("VCD_WGRAD_LOAD_RATE", 22);
    const double load = (static_cast<double>(G) * mch * (128 + halo) + (P.NT / 8) * 128.0) * 16.0 / load_rate;
    const double cost = mma > load ? mma : load;
    if (cost < best) { best = cost; P.M = M; P.G = G; P.mch = mch; }
  }
  P.n_mtiles = g.K >= 128 ? g.K / 128 : 1;
  P.n_slots = (g.taps + P.G - 1) / P.G;
  const int cap = P.M == 128 ? 512 / P.NT : 2 * (512 / P.NT);
  P.TG = cap < P.n_slots ? cap : P.n_slots;
  P.n_tgroups = (P.n_slots + P.TG - 1) / P.TG;
  uint32_t need = static_cast<uint32_t>(P.M == 128 ? P.TG * P.NT : ((P.TG + 1) / 2) * P.NT), cols = 32;
  while (cols < need) cols <<= 1;
  P.tmem_cols = cols;
  auto stage_bytes = [&](int tk) {
    const int ri = (tk + halo + 7) / 8 * 8;
    return static_cast<size_t>(P.G) * P.mch * ri * 16 + static_cast<size_t>(P.NT / 8) * tk * 16;
  };
  // one tensor-TMA box per operand: at most 128 rows (256 8-byte elements) including the tap halo
  P.TK = 64;
  for (int tk : {112, 96, 80}) {
    if ((tk + halo + 7) / 8 * 8 <= 128 && 4 * stage_bytes(tk) <= 200 * 1024) { P.TK = tk; break; }
  }
  P.RI = (P.TK + halo + 7) / 8 * 8;
  const size_t stage = stage_bytes(P.TK);
  // small-channel layers keep the pipeline shallow: a CTA that fills the SM's shared memory cannot share the SM with
  // the data-gradient CTAs of the other ResBlock branches, whose persistent grids then wait for whole wgrad CTAs
  static const int ws_kb_big = tc_env_int("VCD_WGRAD_SMEM_KB", 200), ws_kb_small = tc_env_int("VCD_WGRAD_SMEM_KB_SMALL", 96);
  const bool small_layer = static_cast<long long>(g.taps) * g.K * g.N <= 64 * 64 * 11;
  int NS = static_cast<int>((static_cast<size_t>(small_layer ? ws_kb_small : ws_kb_big) * 1024) / stage);
  P.NS = NS > 8 ? 8 : (NS < 2 ? 2 : NS);
  // slack: an M = 64 operand over a 32-channel tile reads 4 channel groups past the tile (garbage rows, unused)
  const size_t smem = 128 + P.NS * stage + (2 * P.NS + 1) * 8 + 16 + 8 * 1024;
  if (smem > 227 * 1024) {
    snprintf(err, errn, "tc_run_wgrad(%s): shared memory budget exceeded (%zu bytes)", L.name.c_str(), smem);
    return 1;
  }
  // splits of the (batch, time) contraction: enough CTAs for ~1/3 of the SMs per layer (independent layers run
  // concurrently on side streams), never more than one split per 8 time blocks
  // ~48 CTAs per layer for large weight tensors (the split partials are combined with fp32 reductions), up to one
  // CTA per SM when the weight gradient is small (C <= 64: the kernel is bound by the per-SM load rate)
  static const int env_ctas = tc_env_int("VCD_WGRAD_CTAS", 0);
  static const int env_small = tc_env_int("VCD_WGRAD_CTAS_SMALL", 48), env_big = tc_env_int("VCD_WGRAD_CTAS_BIG", 32);
  // tail launches (the last weight gradients of a segment: nothing else is left to share the SMs with) may spread wider
  static const int env_tail = tc_env_int("VCD_WGRAD_CTAS_TAIL", 96);
  int target_ctas = env_ctas > 0 ? env_ctas : (static_cast<long long>(g.taps) * g.K * g.N <= 64 * 64 * 11 ? env_small : env_big);
  if (tail && env_tail > 0) target_ctas = env_tail;
  const long long base_ctas = 1LL * P.n_mtiles * P.n_ntiles * P.n_tgroups;
  P.kb_per_item = (Ld + P.TK - 1) / P.TK;
  const long long total_kb = 1LL * B * P.kb_per_item;
  long long want = (target_ctas + base_ctas - 1) / base_ctas;
  const long long max_splits = (total_kb + 7) / 8;
  if (want > max_splits) want = max_splits;
  // deterministic mode: the splits of a tile add their partial sums in a fixed order (see the turnstile in wgrad_kernel);
  // fewer splits keep the serialised reduction passes short
  static const int det_splits = tc_env_int("VCD_DET_SPLITS", 16);
  if (p->deterministic && want > det_splits) want = det_splits;
  if (want < 1) want = 1;
  P.kb_per_split = static_cast<int>((total_kb + want - 1) / want);
  P.n_splits = static_cast<int>((total_kb + P.kb_per_split - 1) / P.kb_per_split);
  static const int dbg_direct = tc_env_int("VCD_WGRAD_DEBUG_DIRECT", 0);   // wrong results: timing experiment only
  P.direct = (P.n_splits == 1 || dbg_direct) ? 1 : 0;
  P.in = static_cast<const bf16*>(in);
  P.dout = static_cast<const bf16*>(dout);
  P.Lin = Lin;
  P.dbias = dbias;
  P.cmod = L.cout;
  P.turn = nullptr;
  if (p->deterministic && P.n_splits > 1 && !dbg_direct) {
    const long long li = &L - p->layers.data();
    const int passes = P.M == 128 ? P.TG : (P.TG + 1) / 2;
    P.turn_stride = 1 + passes;
    if (!p->d_turn || base_ctas * P.turn_stride > vcd_plan::kTurnInts) {
      snprintf(err, errn, "tc_run_wgrad(%s): deterministic mode needs the turn counters (%lld tiles)", L.name.c_str(), base_ctas);
      return 1;
    }
    P.turn = p->d_turn + li * vcd_plan::kTurnInts;
  }
  P.trace = nullptr;
  {
    static const char* want = getenv("VCD_KTRACE");
    if (want && p->d_trace && (L.name + ":wgrad") == want) {
      P.trace = p->d_trace;
      fprintf(stderr, "[ktrace] %s:wgrad M=%d G=%d mch=%d NT=%d TG=%d tgroups=%d mtiles=%d ntiles=%d TK=%d RI=%d NS=%d splits=%d kb/split=%d taps=%d K=%d N=%d stage=%zu\n",
              L.name.c_str(), P.M, P.G, P.mch, P.NT, P.TG, P.n_tgroups, P.n_mtiles, P.n_ntiles, P.TK, P.RI, P.NS, P.n_splits,
              P.kb_per_split, g.taps, g.K, g.N, stage);
    }
  }
  CUtensorMap tmIn, tmD;
  if (!tc_make_rows_map(&tmIn, in, B, g.K, Lin, P.RI, P.mch) || !tc_make_rows_map(&tmD, dout, B, g.N, Ld, P.TK, P.NT / 8)) {
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

// Common device/host helpers for the vcd kernels.
//
// HBM data layout used by every kernel ("blocked channels-last"):
//     act[b][c / 8][t][c % 8]
// i.e. for a fixed batch item and a fixed group of 8 channels, consecutive time steps are consecutive
// 8-element vectors (16 B in bf16, 32 B in fp32).  Reasons:
//   * one time row of one channel group is exactly the 16-byte unit of the tcgen05 no-swizzle core matrix,
//     so a [rows x C] tile lands in shared memory as [C/8][rows][8] and a dilated tap is a +16*shift byte
//     offset on the UMMA descriptor start address (fwd/dgrad: K-major operand; wgrad: MN-major operand);
//   * the epilogue thread <-> TMEM lane <-> time row mapping writes 16 B per thread, 512 B contiguous per
//     warp per channel group -- coalesced without a shared-memory transpose.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace vcd {

using bf16 = __nv_bfloat16;

struct __align__(16) bf16x8 {
  __nv_bfloat162 v[4];
};

// ---- 8-wide vector load/store of one (time row, channel group) -------------------------------------
template <typename T>
__device__ __forceinline__ void load8(const T* p, float (&f)[8]);

template <>
__device__ __forceinline__ void load8<float>(const float* p, float (&f)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
  f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

template <>
__device__ __forceinline__ void load8<bf16>(const bf16* p, float (&f)[8]) {
  const uint4 raw = __ldg(reinterpret_cast<const uint4*>(p));
  const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}

template <typename T>
__device__ __forceinline__ void store8(T* p, const float (&f)[8]);

template <>
__device__ __forceinline__ void store8<float>(float* p, const float (&f)[8]) {
  reinterpret_cast<float4*>(p)[0] = make_float4(f[0], f[1], f[2], f[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(f[4], f[5], f[6], f[7]);
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&h);
}

template <>
__device__ __forceinline__ void store8<bf16>(bf16* p, const float (&f)[8]) {
  uint4 raw;
  raw.x = pack_bf16x2(f[0], f[1]);
  raw.y = pack_bf16x2(f[2], f[3]);
  raw.z = pack_bf16x2(f[4], f[5]);
  raw.w = pack_bf16x2(f[6], f[7]);
  *reinterpret_cast<uint4*>(p) = raw;
}

__device__ __forceinline__ float lrelu(float v, float slope) { return v > 0.f ? v : v * slope; }

template <typename T>
__device__ __forceinline__ float to_float(T v);
template <>
__device__ __forceinline__ float to_float<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_float<bf16>(bf16 v) { return __bfloat162float(v); }

template <typename T>
__device__ __forceinline__ T from_float(float v);
template <>
__device__ __forceinline__ float from_float<float>(float v) { return v; }
template <>
__device__ __forceinline__ bf16 from_float<bf16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Every (batch item, channel group) row array carries kPadL zero rows before row 0 and kPadR zero rows after
// row L-1.  The pads ARE the convolution's zero padding: a tile with its dilation halo is then one contiguous
// run of 16-byte rows per channel group, fetched with a single 1-D bulk copy (cp.async.bulk) -- no per-row TMA
// requests, no out-of-bounds handling, and taps never bleed across batch items.  kPadR also absorbs the
// overhang of the last 128-row tile / time block of an item.  Pads are (re)zeroed by pad_zero_kernel.
constexpr int kPadL = 32;
constexpr int kPadR = 160;
__host__ __device__ __forceinline__ int padded_len(int L) { return L + kPadL + kPadR; }

// Element offset of row t (may be in [-kPadL, L + kPadR)) of channel group cg of batch item b.
__host__ __device__ __forceinline__ size_t blk_row(int b, int cg, int t, int C, int L) {
  return ((static_cast<size_t>(b) * (C >> 3) + cg) * padded_len(L) + (t + kPadL)) * 8;
}
// Element offset of (b, c, t) in a blocked channels-last tensor with C channels and length L.
__host__ __device__ __forceinline__ size_t blk_off(int b, int c, int t, int C, int L) {
  return blk_row(b, c >> 3, t, C, L) + (c & 7);
}
// Elements of a blocked tensor (including pads).
__host__ __device__ __forceinline__ size_t blk_elems(int B, int C, int L) {
  return static_cast<size_t>(B) * C * padded_len(L);
}

// ---- generalised convolution geometry (shared by SIMT and tensor-core kernels) ----------------------
//   out[b][q*os + n/creal - p][n % creal] = sum_{j<taps} sum_{c<K} Wp[j][c][n] * in[b][q*is + off0 + j*step][c]
// covers Conv1d forward / data-gradient (os=1) and ConvTranspose1d forward (scatter form, os=u) and its
// data-gradient (is=u).  DESIGN.md §4 derives the four instances.
struct ConvGeo {
  int taps, K, N;        // logical packed weight Wp[taps][K][N]
  int is, off0, step;    // input row  = q*is + off0 + j*step
  int os, p, creal;      // output row = q*os + n/creal - p ; output channel = n % creal
};

// Epilogue shared by all conv kernels (T = activation storage type):
//   v = acc + bias[ch] + bias2[b][ch];  v *= (mask > 0 ? 1 : mask_slope);  v *= scale;
//   v += inv_lrelu(res_t) + res2;       inv_lrelu(r) = r > 0 ? r : r * res_inv
//   out_raw = v (fp32);  out_t = T(lrelu(v * tscale, act_slope))
// Only ACTIVATED tensors are stored between layers: the raw residual stream x of a ResBlock
// (modules.py:203-216, `x = xt + x`) is recovered from the stored leaky_relu(x) through the exact inverse of
// the (monotonic) leaky_relu, which removes one tensor write + read per residual add.
struct Epilogue {
  const float* bias;     // [creal] or null
  const float* bias2;    // [B][creal] or null
  const void* mask;      // T, same layout as the output, or null
  const void* res_t;     // T, same layout as the output, or null
  const float* res2;     // fp32, same layout as the output, or null
  float* out_raw;        // fp32 or null
  void* out_t;           // T or null
  float mask_slope, scale, res_inv, tscale, act_slope;
  // Optional phase-packed ("Z") store of out_t, the operand layout of the ConvTranspose1d backward GEMMs:
  // element (row, ch) goes to Z[b][(r*creal + ch)][q] with q = (row + zp) / zu, r = (row + zp) % zu.
  int zu, zp, zLq;
};

template <typename T>
__device__ __forceinline__ void apply_epilogue(const Epilogue& e, const ConvGeo& g, int b, int ro, int ch0,
                                               int Lout, float (&v)[8]) {
  const size_t o = blk_off(b, ch0, ro, g.creal, Lout);
  if (e.bias) {
#pragma unroll
    for (int n = 0; n < 8; ++n) v[n] += __ldg(e.bias + ch0 + n);
  }
  if (e.bias2) {
#pragma unroll
    for (int n = 0; n < 8; ++n) v[n] += __ldg(e.bias2 + static_cast<size_t>(b) * g.creal + ch0 + n);
  }
  if (e.mask) {
    float m[8];
    load8<T>(reinterpret_cast<const T*>(e.mask) + o, m);
#pragma unroll
    for (int n = 0; n < 8; ++n) v[n] *= (m[n] > 0.f ? 1.f : e.mask_slope);
  }
#pragma unroll
  for (int n = 0; n < 8; ++n) v[n] *= e.scale;
  if (e.res_t) {
    float t[8];
    load8<T>(reinterpret_cast<const T*>(e.res_t) + o, t);
#pragma unroll
    for (int n = 0; n < 8; ++n) v[n] += (t[n] > 0.f ? t[n] : t[n] * e.res_inv);
  }
  if (e.res2) {
    float t[8];
    load8<float>(e.res2 + o, t);
#pragma unroll
    for (int n = 0; n < 8; ++n) v[n] += t[n];
  }
  if (e.out_raw) store8<float>(e.out_raw + o, v);
  if (e.out_t) {
    float a[8];
#pragma unroll
    for (int n = 0; n < 8; ++n) a[n] = lrelu(v[n] * e.tscale, e.act_slope);
    if (e.zu > 0) {
      const int q = (ro + e.zp) / e.zu, r = (ro + e.zp) - q * e.zu;
      store8<T>(reinterpret_cast<T*>(e.out_t) + blk_off(b, r * g.creal + ch0, q, e.zu * g.creal, e.zLq), a);
    } else {
      store8<T>(reinterpret_cast<T*>(e.out_t) + o, a);
    }
  }
}

}  // namespace vcd

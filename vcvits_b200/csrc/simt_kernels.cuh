// CUDA-core (FFMA) kernels on the blocked channels-last layout.
//
// Role: (1) the complete fp32 parity path (VCD_MODE_FP32, fp32 storage + FFMA math, bit-for-bit
// independent of the tensor-core kernels so it doubles as their on-device cross-check); (2) in bf16 mode,
// the layers that are too small or too thin for an implicit-GEMM tile (conv_post+tanh, cond, layout
// transposes, bias/column reductions).  All kernels are templated on the storage type T (float | bf16) and
// accumulate in fp32.
#pragma once
#include "common.cuh"

namespace vcd {

// -------------------------------------------------------------------------------------------------
// Generalised convolution, direct form.  One thread = ROWS output positions (q, q+128, ...) x 8 output
// columns; the 8x8 weight block of each (tap, channel group) is a block-uniform load (L1 broadcast).
// grid = (ceil(Lq / (128*ROWS)), N/8, B), block = 128.
// -------------------------------------------------------------------------------------------------
template <typename T, int ROWS>
__global__ void __launch_bounds__(128)
gconv_simt_kernel(const T* __restrict__ in, const float* __restrict__ w, ConvGeo g, Epilogue e,
                  int Lin, int Lq, int Lout) {
  const int b = blockIdx.z;
  const int n0 = blockIdx.y * 8;
  const int q0 = blockIdx.x * (128 * ROWS) + threadIdx.x;
  const int kch = g.K >> 3;

  float acc[ROWS][8];
#pragma unroll
  for (int i = 0; i < ROWS; ++i)
#pragma unroll
    for (int n = 0; n < 8; ++n) acc[i][n] = 0.f;

  const T* in_b = in + blk_row(b, 0, 0, g.K, Lin);
  const size_t cg_stride = static_cast<size_t>(padded_len(Lin)) * 8;
  for (int j = 0; j < g.taps; ++j) {
    int row[ROWS];
    bool ok[ROWS];
    bool any = false;
#pragma unroll
    for (int i = 0; i < ROWS; ++i) {
      const int q = q0 + i * 128;
      row[i] = q * g.is + g.off0 + j * g.step;
      ok[i] = (q < Lq) && (row[i] >= 0) && (row[i] < Lin);
      any |= ok[i];
    }
    if (!__syncthreads_or(any)) continue;
    const float* wj = w + static_cast<size_t>(j) * g.K * g.N + n0;
    for (int cc = 0; cc < kch; ++cc) {
      float xin[ROWS][8];
#pragma unroll
      for (int i = 0; i < ROWS; ++i) {
        if (ok[i]) {
          load8<T>(in_b + cc * cg_stride + static_cast<size_t>(row[i]) * 8, xin[i]);
        } else {
#pragma unroll
          for (int c = 0; c < 8; ++c) xin[i][c] = 0.f;
        }
      }
      const float* wp = wj + static_cast<size_t>(cc) * 8 * g.N;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float4 w0 = __ldg(reinterpret_cast<const float4*>(wp + static_cast<size_t>(c) * g.N));
        const float4 w1 = __ldg(reinterpret_cast<const float4*>(wp + static_cast<size_t>(c) * g.N) + 1);
#pragma unroll
        for (int i = 0; i < ROWS; ++i) {
          const float x = xin[i][c];
          acc[i][0] = fmaf(x, w0.x, acc[i][0]);
          acc[i][1] = fmaf(x, w0.y, acc[i][1]);
          acc[i][2] = fmaf(x, w0.z, acc[i][2]);
          acc[i][3] = fmaf(x, w0.w, acc[i][3]);
          acc[i][4] = fmaf(x, w1.x, acc[i][4]);
          acc[i][5] = fmaf(x, w1.y, acc[i][5]);
          acc[i][6] = fmaf(x, w1.z, acc[i][6]);
          acc[i][7] = fmaf(x, w1.w, acc[i][7]);
        }
      }
    }
  }

  // epilogue
  const int r = n0 / g.creal;
  const int ch0 = n0 - r * g.creal;
#pragma unroll
  for (int i = 0; i < ROWS; ++i) {
    const int q = q0 + i * 128;
    if (q >= Lq) continue;
    const int ro = q * g.os + r - g.p;
    if (ro < 0 || ro >= Lout) continue;
    float v[8];
#pragma unroll
    for (int n = 0; n < 8; ++n) v[n] = acc[i][n];
    apply_epilogue<T>(e, g, b, ro, ch0, Lout, v);
  }
}

// -------------------------------------------------------------------------------------------------
// Weight gradient of the generalised convolution (fwd geometry g):
//   dWp[j][c][n] += sum_{b,q} in[b][q*is + off0 + j*step][c] * dout[b][q*os + n/creal - p][n % creal]
// grid = (taps * K/8 * N/8, splits), block = 256; one 8x8 block of dWp per CTA, fp32 atomics to combine
// the time splits (dWp must be zeroed by the caller).
// -------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
gconv_wgrad_simt_kernel(const T* __restrict__ in, const T* __restrict__ dout, float* __restrict__ dwp,
                        ConvGeo g, int B, int Lin, int Lq, int Lout, int rows_per_split) {
  const int kch = g.K >> 3, nch = g.N >> 3;
  int id = blockIdx.x;
  const int nc = id % nch; id /= nch;
  const int cc = id % kch; id /= kch;
  const int j = id;
  const int n0 = nc * 8;
  const int r = n0 / g.creal;
  const int ch0 = n0 - r * g.creal;

  const long long total = static_cast<long long>(B) * Lq;
  const long long begin = static_cast<long long>(blockIdx.y) * rows_per_split;
  long long end = begin + rows_per_split;
  if (end > total) end = total;

  float acc[8][8];
#pragma unroll
  for (int c = 0; c < 8; ++c)
#pragma unroll
    for (int n = 0; n < 8; ++n) acc[c][n] = 0.f;

  for (long long idx = begin + threadIdx.x; idx < end; idx += 256) {
    const int b = static_cast<int>(idx / Lq);
    const int q = static_cast<int>(idx - static_cast<long long>(b) * Lq);
    const int ri = q * g.is + g.off0 + j * g.step;
    const int ro = q * g.os + r - g.p;
    if (ri < 0 || ri >= Lin || ro < 0 || ro >= Lout) continue;
    float x[8], d[8];
    load8<T>(in + blk_off(b, cc * 8, ri, g.K, Lin), x);
    load8<T>(dout + blk_off(b, ch0, ro, g.creal, Lout), d);
#pragma unroll
    for (int c = 0; c < 8; ++c)
#pragma unroll
      for (int n = 0; n < 8; ++n) acc[c][n] = fmaf(x[c], d[n], acc[c][n]);
  }

  __shared__ float red[8][64];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int c = 0; c < 8; ++c)
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      const float s = warp_sum(acc[c][n]);
      if (lane == 0) red[warp][c * 8 + n] = s;
    }
  __syncthreads();
  if (threadIdx.x < 64) {
    float s = 0.f;
#pragma unroll
    for (int wv = 0; wv < 8; ++wv) s += red[wv][threadIdx.x];
    const int c = threadIdx.x >> 3, n = threadIdx.x & 7;
    atomicAdd(dwp + (static_cast<size_t>(j) * g.K + cc * 8 + c) * g.N + n0 + n, s);
  }
}

// -------------------------------------------------------------------------------------------------
// Column sums over time (bias gradients): out[(per_batch ? b*cmod : 0) + c % cmod] += sum_t d[b][c][t]
// (cmod < C folds the u phase groups of a phase-packed ConvTranspose gradient onto the real channels).
// grid = (C/8, B, splits), block = 256.  `out` must be zeroed by the caller.
// colsum_det_kernel + sum_parts_kernel: the same sums for the deterministic mode (vcd_set_deterministic).  Grid
// (cmod / 8, 1, parts): a block walks ALL fold copies and all batch items of its slice of the time range in a fixed
// order and stores its partial sums parts[z][c]; sum_parts_kernel then adds the parts in index order.  No atomics.
// -------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
colsum_det_kernel(const T* __restrict__ d, float* __restrict__ parts, int C, int L, int B, int cmod) {
  const int cgm = blockIdx.x;
  const int chunk = (L + gridDim.z - 1) / gridDim.z;
  const int t0 = blockIdx.z * chunk, t1 = min(L, t0 + chunk);
  float acc[8];
#pragma unroll
  for (int n = 0; n < 8; ++n) acc[n] = 0.f;
  for (int b = 0; b < B; ++b) {
    for (int cg = cgm; cg < C / 8; cg += cmod / 8) {
      const T* base = d + blk_row(b, cg, 0, C, L);
      for (int t = t0 + threadIdx.x; t < t1; t += 256) {
        float v[8];
        load8<T>(base + static_cast<size_t>(t) * 8, v);
#pragma unroll
        for (int n = 0; n < 8; ++n) acc[n] += v[n];
      }
    }
  }
  __shared__ float red[8][8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    const float s = warp_sum(acc[n]);
    if (lane == 0) red[warp][n] = s;
  }
  __syncthreads();
  if (threadIdx.x < 8) {
    float s = 0.f;
#pragma unroll
    for (int wv = 0; wv < 8; ++wv) s += red[wv][threadIdx.x];
    parts[static_cast<size_t>(blockIdx.z) * cmod + cgm * 8 + threadIdx.x] = s;
  }
}

__global__ void __launch_bounds__(256) sum_parts_kernel(const float* __restrict__ parts, int nparts, int n, float* __restrict__ out) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  float s = 0.f;
  for (int p = 0; p < nparts; ++p) s += parts[static_cast<size_t>(p) * n + i];
  out[i] = s;
}

// -------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
colsum_kernel(const T* __restrict__ d, float* __restrict__ out, int C, int L, int per_batch, int cmod) {
  const int cg = blockIdx.x, b = blockIdx.y;
  const int chunk = (L + gridDim.z - 1) / gridDim.z;
  const int t0 = blockIdx.z * chunk;
  const int t1 = min(L, t0 + chunk);
  float acc[8];
#pragma unroll
  for (int n = 0; n < 8; ++n) acc[n] = 0.f;
  const T* base = d + blk_row(b, cg, 0, C, L);
  for (int t = t0 + threadIdx.x; t < t1; t += 256) {
    float v[8];
    load8<T>(base + static_cast<size_t>(t) * 8, v);
#pragma unroll
    for (int n = 0; n < 8; ++n) acc[n] += v[n];
  }
  __shared__ float red[8][8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    const float s = warp_sum(acc[n]);
    if (lane == 0) red[warp][n] = s;
  }
  __syncthreads();
  if (threadIdx.x < 8) {
    float s = 0.f;
#pragma unroll
    for (int wv = 0; wv < 8; ++wv) s += red[wv][threadIdx.x];
    atomicAdd(out + (per_batch ? static_cast<size_t>(b) * cmod : 0) + (cg * 8 + threadIdx.x) % cmod, s);
  }
}

// -------------------------------------------------------------------------------------------------
// Zero the pad rows of a list of blocked tensors (see kPadL / kPadR in common.cuh).
// grid = (total row arrays over all tensors), block = 64; one CTA per (tensor, batch item, channel group).
// -------------------------------------------------------------------------------------------------
struct PadJob {
  long long off;      // byte offset of the tensor in the workspace
  int arrays;         // B * C/8
  int L;
  int esize;          // 2 | 4
  int first_block;
};

__global__ void __launch_bounds__(64)
pad_zero_kernel(const PadJob* __restrict__ jobs, int njobs, char* __restrict__ ws) {
  int lo = 0, hi = njobs - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (jobs[mid].first_block <= static_cast<int>(blockIdx.x)) lo = mid; else hi = mid - 1;
  }
  const PadJob jb = jobs[lo];
  const int arr = blockIdx.x - jb.first_block;
  const size_t row_bytes = 8 * jb.esize;
  char* base = ws + jb.off + static_cast<size_t>(arr) * padded_len(jb.L) * row_bytes;
  const int vec_per_row = static_cast<int>(row_bytes / 16);
  uint4* left = reinterpret_cast<uint4*>(base);
  uint4* right = reinterpret_cast<uint4*>(base + static_cast<size_t>(kPadL + jb.L) * row_bytes);
  const uint4 z = make_uint4(0, 0, 0, 0);
  for (int i = threadIdx.x; i < kPadL * vec_per_row; i += 64) left[i] = z;
  for (int i = threadIdx.x; i < kPadR * vec_per_row; i += 64) right[i] = z;
}

// -------------------------------------------------------------------------------------------------
// Boundary layout changes.
// -------------------------------------------------------------------------------------------------
// x: fp32 [B, C, *] with element strides -> blocked T of length L.  grid = (ceil(L/128), C/8, B), block = 128.
// `starts` (optional, device): per-item first frame -- the segment gather of commons.slice_segments /
// rand_slice_segments (vits/commons.py:48-64) folded into the decoder's input load.
template <typename T>
__global__ void __launch_bounds__(128)
ncl_to_blocked_kernel(const float* __restrict__ x, long long sb, long long sc, long long st,
                      T* __restrict__ out, int C, int L, const long long* __restrict__ starts) {
  const int t = blockIdx.x * 128 + threadIdx.x;
  if (t >= L) return;
  const int cg = blockIdx.y, b = blockIdx.z;
  const long long t_src = t + (starts ? starts[b] : 0);
  float v[8];
#pragma unroll
  for (int n = 0; n < 8; ++n) v[n] = __ldg(x + b * sb + (cg * 8 + n) * sc + t_src * st);
  store8<T>(out + blk_off(b, cg * 8, t, C, L), v);
}

// blocked fp32 of length L -> contiguous fp32 [B, C, Lfull] at frame offset starts[b] (the scatter that is the backward
// of the segment gather; the caller zero-fills `out` when Lfull > L).
__global__ void __launch_bounds__(128)
blocked_to_ncl_kernel(const float* __restrict__ in, float* __restrict__ out, int C, int L, long long Lfull,
                      const long long* __restrict__ starts) {
  const int t = blockIdx.x * 128 + threadIdx.x;
  if (t >= L) return;
  const int cg = blockIdx.y, b = blockIdx.z;
  const long long t_dst = t + (starts ? starts[b] : 0);
  float v[8];
  load8<float>(in + blk_off(b, cg * 8, t, C, L), v);
#pragma unroll
  for (int n = 0; n < 8; ++n) out[(static_cast<size_t>(b) * C + cg * 8 + n) * Lfull + t_dst] = v[n];
}

// -------------------------------------------------------------------------------------------------
// cond (speaker conditioning, 1x1 conv on a length-1 input): cb[b][n] = bias[n] + sum_c w[n][c] g[b][c]
// grid = (ceil(N/8), B), block = 256: one warp per output (lanes walk the G-long weight row, coalesced).
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
cond_fwd_kernel(const float* __restrict__ w, const float* __restrict__ bias, const float* __restrict__ gv,
                float* __restrict__ cb, int N, int G) {
  const int lane = threadIdx.x & 31;
  const int n = blockIdx.x * 8 + (threadIdx.x >> 5), b = blockIdx.y;
  if (n >= N) return;
  float s = 0.f;
  for (int c = lane; c < G; c += 32) s = fmaf(__ldg(w + static_cast<size_t>(n) * G + c), __ldg(gv + static_cast<size_t>(b) * G + c), s);
  s = warp_sum(s);
  if (lane == 0) cb[static_cast<size_t>(b) * N + n] = s + bias[n];
}

// Backward of cond + conv_pre.bias from dcb[b][n] = sum_t d0[b][t][n].
//   d_pre_bias[n] = sum_b dcb ; d_cond_bias[n] = same ; d_cond_w[n][c] = sum_b dcb[b][n] g[b][c]
//   dg[b][c] = sum_n dcb[b][n] w[n][c]
// grid = (ceil(max(N*G, B*G, N)/256)), block = 256.  Null pointers skip the respective output.
__global__ void __launch_bounds__(256)
cond_bwd_kernel(const float* __restrict__ dcb, const float* __restrict__ w, const float* __restrict__ gv,
                float* __restrict__ d_pre_bias, float* __restrict__ d_cond_bias, float* __restrict__ d_cond_w,
                int B, int N, int G, float gscale) {
  const long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (i < N) {
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += dcb[static_cast<size_t>(b) * N + i];
    if (d_pre_bias) d_pre_bias[i] = s * gscale;
    if (d_cond_bias) d_cond_bias[i] = s * gscale;
  }
  if (d_cond_w && i < static_cast<long long>(N) * G) {
    const int n = static_cast<int>(i / G), c = static_cast<int>(i % G);
    float s = 0.f;
    for (int b = 0; b < B; ++b) s = fmaf(dcb[static_cast<size_t>(b) * N + n], gv[static_cast<size_t>(b) * G + c], s);
    d_cond_w[i] = s * gscale;
  }
}

// dg[b][c] = sum_n dcb[b][n] * w[n][c]  (gradient w.r.t. the speaker embedding).  grid = (ceil(G/32), B), block = 256:
// the eight warps split n, so no thread walks a 512-long dependent chain of L2-latency loads.
__global__ void __launch_bounds__(256)
cond_dg_kernel(const float* __restrict__ dcb, const float* __restrict__ w, float* __restrict__ dg, int N, int G) {
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  const int b = blockIdx.y, c = blockIdx.x * 32 + lane;
  float s = 0.f;
  if (c < G) {
#pragma unroll 4
    for (int n = wp; n < N; n += 8) s = fmaf(__ldg(dcb + static_cast<size_t>(b) * N + n), __ldg(w + static_cast<size_t>(n) * G + c), s);
  }
  __shared__ float red[8][33];
  red[wp][lane] = s;
  __syncthreads();
  if (wp == 0 && c < G) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k][lane];
    dg[static_cast<size_t>(b) * G + c] = t;
  }
}

// -------------------------------------------------------------------------------------------------
// conv_post (C -> 1, k = 7, no bias) + tanh.  HBM-bound: 2*C bytes read per output sample (bf16).
// One thread owns one INPUT ROW (time step): it loads the row once (C/8 16-byte loads, consecutive threads read
// consecutive 16-byte rows of a channel group: perfectly coalesced, no halo re-reads) and forms the seven per-tap
// partial dot products  p_j[r] = sum_c a[r][c] * w[c][j];  the output  y[t] = tanh(sum_j p_j[t + j - 3])  then
// gathers seven floats through shared memory.  A CTA of 256 rows produces 250 outputs (3-row halo each side; the
// zero pad rows of the layout ARE the convolution's zero padding).  The round-1 kernel re-read every row seven
// times through L1 (one thread per output, seven overlapping row loads).
// w is the raw parameter [1][C][7].  grid = (ceil(L/250), B), block = 256, smem = C*8 floats.
// -------------------------------------------------------------------------------------------------
constexpr int kPostOut = 250;

template <typename T>
__global__ void __launch_bounds__(256)
conv_post_fwd_kernel(const T* __restrict__ a, const float* __restrict__ w, float* __restrict__ y, int C, int L) {
  extern __shared__ float ws[];          // [C][8]: the seven taps of a channel, padded to two float4
  __shared__ float part[7][256 + 8];
  for (int i = threadIdx.x; i < C * 8; i += 256) {
    const int c = i >> 3, j = i & 7;
    ws[i] = j < 7 ? __ldg(w + c * 7 + j) : 0.f;
  }
  __syncthreads();
  const int b = blockIdx.y;
  const int r = blockIdx.x * kPostOut - 3 + static_cast<int>(threadIdx.x);   // this thread's input row
  float p[7];
#pragma unroll
  for (int j = 0; j < 7; ++j) p[j] = 0.f;
  if (r < L + kPadR) {                    // r >= -3 >= -kPadL always
    for (int cg = 0; cg < (C >> 3); ++cg) {
      float v[8];
      load8<T>(a + blk_row(b, cg, r, C, L), v);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float4 w0 = *reinterpret_cast<const float4*>(ws + (cg * 8 + c) * 8);
        const float4 w1 = *reinterpret_cast<const float4*>(ws + (cg * 8 + c) * 8 + 4);
        p[0] = fmaf(v[c], w0.x, p[0]); p[1] = fmaf(v[c], w0.y, p[1]); p[2] = fmaf(v[c], w0.z, p[2]);
        p[3] = fmaf(v[c], w0.w, p[3]); p[4] = fmaf(v[c], w1.x, p[4]); p[5] = fmaf(v[c], w1.y, p[5]);
        p[6] = fmaf(v[c], w1.z, p[6]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 7; ++j) part[j][threadIdx.x] = p[j];
  __syncthreads();
  const int t = blockIdx.x * kPostOut + static_cast<int>(threadIdx.x);
  if (threadIdx.x < kPostOut && t < L) {
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < 7; ++j) acc += part[j][threadIdx.x + j];   // row t + j - 3 is thread (t - t0) + j
    y[static_cast<size_t>(b) * L + t] = tanhf(acc);
  }
}

// Data gradient of conv_post + tanh, fused with the final leaky_relu(0.01) mask and the 1/num_kernels of the
// branch mean:  G[b][t][c] = mask(a) * scale * sum_j w[c][j] * dpost[t + 3 - j],  dpost = dy * (1 - y^2).
// grid = (ceil(L/128), C/8, B), block = 128.
template <typename T>
__global__ void __launch_bounds__(128)
conv_post_dgrad_kernel(const float* __restrict__ dy, const float* __restrict__ y, const float* __restrict__ w,
                       const T* __restrict__ a, float mask_slope, float scale, float* __restrict__ out_raw,
                       T* __restrict__ out_t, int C, int L) {
  const int t = blockIdx.x * 128 + threadIdx.x;
  const int cg = blockIdx.y, b = blockIdx.z;
  __shared__ float wsm[56];   // the channel group's [8][7] weights
  if (threadIdx.x < 56) wsm[threadIdx.x] = __ldg(w + cg * 56 + threadIdx.x);
  __syncthreads();
  if (t >= L) return;
  // branch-free operand fetch: every load is issued before the first dependent instruction
  float yy[7], dd[7], dp[7];
#pragma unroll
  for (int j = 0; j < 7; ++j) {
    const int s = t + 3 - j;
    const bool ok = s >= 0 && s < L;
    yy[j] = __ldg(y + static_cast<size_t>(b) * L + (ok ? s : t));
    dd[j] = __ldg(dy + static_cast<size_t>(b) * L + (ok ? s : t));
  }
  const size_t o = blk_off(b, cg * 8, t, C, L);
  float m[8], v[8];
  load8<T>(a + o, m);
#pragma unroll
  for (int j = 0; j < 7; ++j) {
    const int s = t + 3 - j;
    dp[j] = (s >= 0 && s < L) ? dd[j] * (1.f - yy[j] * yy[j]) : 0.f;
  }
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 7; ++j) s = fmaf(wsm[c * 7 + j], dp[j], s);
    v[c] = s * (m[c] > 0.f ? 1.f : mask_slope) * scale;
  }
  if (out_raw) store8<float>(out_raw + o, v);
  if (out_t) store8<T>(out_t + o, v);
}

// Weight gradient of conv_post: dw[c][j] += sum_{b,t} dpost[b][t] * a[b][t + j - 3][c]   (param layout [1][C][7]).
// grid = (C/8, B, splits), block = 256 (deterministic mode: (C/8, 1, parts) with per-part partial sums).  dw must be zeroed.
template <typename T>
__global__ void __launch_bounds__(256)
conv_post_wgrad_kernel(const float* __restrict__ dy, const float* __restrict__ y, const T* __restrict__ a,
                       float* __restrict__ dw, int C, int L, int B, float* __restrict__ parts) {
  const int cg = blockIdx.x;
  const int chunk = (L + gridDim.z - 1) / gridDim.z;
  const int r0 = blockIdx.z * chunk, r1 = min(L, r0 + chunk);
  float acc[8][7];
#pragma unroll
  for (int c = 0; c < 8; ++c)
#pragma unroll
    for (int j = 0; j < 7; ++j) acc[c][j] = 0.f;
  // gridDim.y == B: one item per block; gridDim.y == 1 (deterministic mode): every item, in order
  for (int b = blockIdx.y; b < B; b += gridDim.y) {
  const T* base = a + blk_row(b, cg, 0, C, L);
  for (int r = r0 + threadIdx.x; r < r1; r += 256) {
    float v[8];
    load8<T>(base + static_cast<size_t>(r) * 8, v);
#pragma unroll
    for (int j = 0; j < 7; ++j) {
      const int t = r - j + 3;  // output position that reads row r with tap j
      float dp = 0.f;
      if (t >= 0 && t < L) {
        const float yy = __ldg(y + static_cast<size_t>(b) * L + t);
        dp = __ldg(dy + static_cast<size_t>(b) * L + t) * (1.f - yy * yy);
      }
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[c][j] = fmaf(v[c], dp, acc[c][j]);
    }
  }
  }
  __shared__ float red[8][56];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int c = 0; c < 8; ++c)
#pragma unroll
    for (int j = 0; j < 7; ++j) {
      const float s = warp_sum(acc[c][j]);
      if (lane == 0) red[warp][c * 7 + j] = s;
    }
  __syncthreads();
  if (threadIdx.x < 56) {
    float s = 0.f;
#pragma unroll
    for (int wv = 0; wv < 8; ++wv) s += red[wv][threadIdx.x];
    if (parts) parts[static_cast<size_t>(blockIdx.z) * (C * 7) + cg * 56 + threadIdx.x] = s;   // deterministic mode: summed by sum_parts_kernel
    else atomicAdd(dw + cg * 56 + threadIdx.x, s);
  }
}

}  // namespace vcd

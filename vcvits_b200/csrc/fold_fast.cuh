// Bandwidth-shaped weight-norm fold / unfold for the tensor-core path (bf16 mode).
//
// The generic kernels in fold.cuh address the packed layouts element by element through WeightMap: every thread of
// wn_pack_kernel gathers 8 values that lie k (or cin*k) floats apart and wn_unfold_kernel reads dWp[j][c][n] along
// its slowest direction -- 13x / 9x more L2 sectors than bytes used (ncu, profiles/r01_ncu_launch_summary.md).
// Here one CTA owns EIGHT ROWS of a parameter (rows = dim 0 of weight_v: output channels of a Conv1d, input channels
// of a ConvTranspose1d -- the dimension weight_norm normalises over, modules.py:10) and stages them in shared memory:
//   pack   : rows are read once, contiguously (row_len floats each); the eight warps compute the eight row norms
//            (the separate norm pass disappears), then BOTH operand formats (forward and data-gradient B operands,
//            fold.cuh FMT_TC) are written as 16-byte units in runs of >= 128 contiguous bytes;
//   unfold : the packed gradient dWp[tap][c][n] is gathered as whole 32-byte sectors (8 rows = 8 consecutive n for a
//            Conv1d; a contiguous run of n for a ConvTranspose1d) and transposed through shared memory; each warp then
//            owns one row: <dw, v>, weight_g gradient, weight_v gradient -- row-contiguous reads and writes.
// Both walk a device-side job table (one launch for all layers of the model / of a backward segment).
#pragma once
#include "common.cuh"
#include "fold.cuh"

namespace vcd {

// Parameter tensor [rows][inner][k] (Conv1d: [cout][cin][k]; ConvTranspose1d: [cin][cout][k]).  Two packed formats
// (FMT_TC, bf16 [n/NT][tap][K/8][n%NT][c%8]):
//   format R ("row is the GEMM column"):      Conv1d forward,          ConvTranspose1d data gradient
//        n = row,  c = r*inner + i,  tap = s        (jj = s*u + r;  u = 1 for Conv1d, i.e. c = i, tap = jj)
//   format C ("row is the contraction index"): Conv1d data gradient,    ConvTranspose1d forward
//        n = r*inner + i,  c = row,  tap = Conv1d ? k-1-jj : s
struct FastPackJob {
  int p_w, p_g;          // parameter indices; p_g < 0 (or a NULL pointer in the table): no weight norm
  int norm_off;          // where the row norms go (weight-normed only)
  int rows, inner, k, u; // u = 1: Conv1d
  int is_convt;
  int kshift;            // log2(k) for even (power-of-two) k, 0 for odd k
  int nt_r, nt_c;        // column tiles of the two formats
  long long dst_r, dst_c;  // element offsets into the bf16 arena (-1: format not needed)
  int first_block;       // first blockIdx.x (one block per 8 rows)
};

constexpr int kFoldRows = 8;
constexpr int kFoldThreads = 512;

// Shared-memory position of parameter element idx = i*k + jj of a row.  Odd k: the natural order is already conflict-free
// for both access directions (pos == idx, no division anywhere); even k (ConvTranspose1d: powers of two) pads every inner
// index to k + 1 elements.
__host__ __device__ __forceinline__ int fold_ks(int k) { return k | 1; }
__host__ __device__ __forceinline__ int fold_srow(int inner, int k) { return inner * fold_ks(k) + 4; }   // +4: rows land in distinct banks, stay 16-byte aligned
__device__ __forceinline__ int fold_pos(int idx, int k, int kshift) {
  return (k & 1) ? idx : (((idx >> kshift) * (k + 1)) + (idx & (k - 1)));
}
__host__ inline int fold_kshift(int k) {   // log2(k) for the even (power-of-two) kernel sizes, -1 otherwise
  if (k & 1) return 0;
  int s = 0;
  while ((1 << s) < k) ++s;
  return (1 << s) == k ? s : -1;
}

// grid = sum over jobs of rows/8, block = 512, dynamic smem = 8 * fold_srow floats (max over jobs).
__global__ void __launch_bounds__(kFoldThreads)
wn_pack_fast_kernel(const FastPackJob* __restrict__ jobs, int njobs, const float* const* __restrict__ params,
                    float* __restrict__ norms, bf16* __restrict__ arena) {
  extern __shared__ __align__(16) float sm[];
  __shared__ float s_part[16];
  __shared__ float s_scale[kFoldRows];
  int lo = 0, hi = njobs - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (jobs[mid].first_block <= static_cast<int>(blockIdx.x)) lo = mid; else hi = mid - 1;
  }
  const FastPackJob jb = jobs[lo];
  const int row0 = (static_cast<int>(blockIdx.x) - jb.first_block) * kFoldRows;
  const int k = jb.k, ks = fold_ks(k), inner = jb.inner, kshift = jb.kshift;
  const int row_len = inner * k, srow = fold_srow(inner, k);
  const float* __restrict__ w = params[jb.p_w] + static_cast<size_t>(row0) * row_len;
  const float* __restrict__ gv = jb.p_g >= 0 ? params[jb.p_g] : nullptr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- load 8 contiguous rows; four independent 16-byte loads in flight per thread ----
  const int vec_per_row = row_len >> 2, nvec = kFoldRows * vec_per_row;
  for (int v0 = threadIdx.x; v0 < nvec; v0 += 4 * kFoldThreads) {
    float4 x[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int v = v0 + q * kFoldThreads;
      if (v < nvec) x[q] = __ldg(reinterpret_cast<const float4*>(w) + v);   // the 8 rows are one contiguous block
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int v = v0 + q * kFoldThreads;
      if (v >= nvec) continue;
      const int r = v / vec_per_row, idx = (v - r * vec_per_row) * 4;
      float* dst = sm + r * srow;
      if (k & 1) {
        *reinterpret_cast<float4*>(dst + idx) = x[q];
      } else {               // 4 consecutive elements share the inner index (k is a multiple of 4)
        float* d2 = dst + fold_pos(idx, k, kshift);
        d2[0] = x[q].x; d2[1] = x[q].y; d2[2] = x[q].z; d2[3] = x[q].w;
      }
    }
  }
  __syncthreads();
  // ---- row norms: two warps per row ----
  {
    const int r = warp >> 1, half = warp & 1;
    float s = 0.f;
    for (int idx = half * 32 + lane; idx < row_len; idx += 64) {
      const float x = sm[r * srow + fold_pos(idx, k, kshift)];
      s = fmaf(x, x, s);
    }
    s = warp_sum(s);
    if (lane == 0) s_part[warp] = s;
  }
  __syncthreads();
  if (threadIdx.x < kFoldRows) {
    float scale = 1.f;
    if (gv) {
      const float nrm = sqrtf(s_part[2 * threadIdx.x] + s_part[2 * threadIdx.x + 1]);
      norms[jb.norm_off + row0 + threadIdx.x] = nrm;
      scale = __ldg(gv + row0 + threadIdx.x) / nrm;
    }
    s_scale[threadIdx.x] = scale;
  }
  __syncthreads();
  const int u = jb.u, taps = k / u;
  // ---- format R: 16-byte unit = (tap, 8 consecutive c, one row n); consecutive rows are consecutive n ----
  if (jb.dst_r >= 0) {
    const int K = u * inner, kgroups = K >> 3;
    bf16* __restrict__ dst = arena + jb.dst_r;
    const int units = taps * kgroups * kFoldRows;
    for (int t = threadIdx.x; t < units; t += kFoldThreads) {
      const int r = t & 7;
      const int cg = (t >> 3) % kgroups, s = (t >> 3) / kgroups;
      const int c0 = cg * 8, rr = c0 / inner, i0 = c0 - rr * inner;   // inner % 8 == 0: the 8 c share rr
      const int jj = s * u + rr;
      const float sc = s_scale[r];
      float v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = sm[r * srow + (i0 + e) * ks + jj] * sc;
      const int n = row0 + r, nt = n / jb.nt_r, nl = n - nt * jb.nt_r;
      store8<bf16>(dst + ((static_cast<size_t>(nt) * taps + s) * kgroups + cg) * (static_cast<size_t>(jb.nt_r) * 8) + nl * 8, v);
    }
  }
  // ---- format C: 16-byte unit = (tap, the tile's 8 rows = one contraction group, one column n); consecutive i are consecutive n ----
  if (jb.dst_c >= 0) {
    const int kgroups = jb.rows >> 3, cg = row0 >> 3;
    bf16* __restrict__ dst = arena + jb.dst_c;
    const int units = k * inner;            // (jj, i)
    float sc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) sc[e] = s_scale[e];
    for (int t = threadIdx.x; t < units; t += kFoldThreads) {
      const int jj = t / inner, i = t - jj * inner;
      int tap, n;
      if (jb.is_convt) { tap = jj / u; n = (jj - tap * u) * inner + i; }
      else { tap = k - 1 - jj; n = i; }
      float v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = sm[e * srow + i * ks + jj] * sc[e];
      const int nt = n / jb.nt_c, nl = n - nt * jb.nt_c;
      store8<bf16>(dst + ((static_cast<size_t>(nt) * taps + tap) * kgroups + cg) * (static_cast<size_t>(jb.nt_c) * 8) + nl * 8, v);
    }
  }
}

// Backward: dWp[tap][c][n] (forward geometry of the layer; Conv1d: [jj][ci][co], ConvTranspose1d: [s][ci][r*cout+co])
//   -> weight_g / weight_v (or plain weight) gradients.  `gscale` multiplies every gradient written (1/world_size for
//   data-parallel averaging before a SUM all-reduce).
struct FastUnfoldJob {
  int kind;              // 0 = 8-row weight job, 1 = copy job (biases, tensors already in parameter layout)
  int p_w, p_g, norm_off;
  int rows, inner, k, u, is_convt, kshift;
  int K, N;              // logical dims of dWp
  long long src_off;     // offset (floats) into the gradient scratch
  long long numel;       // copy job
  int first_block;
};

__global__ void __launch_bounds__(kFoldThreads)
wn_unfold_fast_kernel(const FastUnfoldJob* __restrict__ jobs, int njobs, const float* const* __restrict__ params,
                      float* const* __restrict__ dparams, const float* __restrict__ norms,
                      const float* __restrict__ scratch, int block_base, float gscale) {
  extern __shared__ __align__(16) float sm[];
  __shared__ float s_part[16];
  const int blk = static_cast<int>(blockIdx.x) + block_base;
  int lo = 0, hi = njobs - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (jobs[mid].first_block <= blk) lo = mid; else hi = mid - 1;
  }
  const FastUnfoldJob jb = jobs[lo];
  const int rel = blk - jb.first_block;
  if (jb.kind == 1) {
    const long long e0 = static_cast<long long>(rel) * 2048 + threadIdx.x;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const long long e = e0 + q * kFoldThreads;
      if (e < jb.numel) dparams[jb.p_w][e] = scratch[jb.src_off + e] * gscale;
    }
    return;
  }
  const int row0 = rel * kFoldRows;
  const int k = jb.k, ks = fold_ks(k), inner = jb.inner, u = jb.u, kshift = jb.kshift;
  const int row_len = inner * k, srow = fold_srow(inner, k);
  const float* __restrict__ dwp = scratch + jb.src_off;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (!jb.is_convt) {
    // dWp[jj][ci][co0 .. co0+7]: one full 32-byte sector per (jj, ci); four sectors in flight per thread
    for (int t0 = threadIdx.x; t0 < row_len; t0 += 4 * kFoldThreads) {
      float4 a[4], b[4];
      int pos[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int t = t0 + q * kFoldThreads;
        pos[q] = -1;
        if (t < row_len) {
          const int jj = t / inner, i = t - jj * inner;   // consecutive threads: consecutive ci (N floats apart, one sector each)
          const float* src = dwp + (static_cast<size_t>(jj) * jb.K + i) * jb.N + row0;
          a[q] = __ldg(reinterpret_cast<const float4*>(src));
          b[q] = __ldg(reinterpret_cast<const float4*>(src) + 1);
          pos[q] = i * ks + jj;
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (pos[q] < 0) continue;
        float* d = sm + pos[q];
        d[0] = a[q].x; d[srow] = a[q].y; d[2 * srow] = a[q].z; d[3 * srow] = a[q].w;
        d[4 * srow] = b[q].x; d[5 * srow] = b[q].y; d[6 * srow] = b[q].z; d[7 * srow] = b[q].w;
      }
    }
  } else {
    // dWp[s][ci0 + r][n], n = rr*inner + co contiguous: coalesced rows
    const int taps = k / u, N = jb.N;
    const int total = kFoldRows * taps * N;
    for (int t0 = threadIdx.x; t0 < total; t0 += 4 * kFoldThreads) {
      float x[4];
      int pos[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int t = t0 + q * kFoldThreads;
        pos[q] = -1;
        if (t < total) {
          const int rs = t / N, n = t - rs * N, r = rs / taps, s = rs - r * taps;
          const int rr = n / inner, i = n - rr * inner;
          x[q] = __ldg(dwp + (static_cast<size_t>(s) * jb.K + row0 + r) * N + n);
          pos[q] = r * srow + i * ks + s * u + rr;
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (pos[q] >= 0) sm[pos[q]] = x[q];
    }
  }
  __syncthreads();
  // two warps per row: warp pair (2r, 2r+1) owns row row0 + r
  const int r = warp >> 1, half = warp & 1;
  const int row = row0 + r;
  float* __restrict__ dw = dparams[jb.p_w] + static_cast<size_t>(row) * row_len;
  const float* srcrow = sm + r * srow;
  const bool normed = jb.p_g >= 0 && params[jb.p_g] != nullptr;
  if (!normed) {
#pragma unroll 4
    for (int idx = half * 32 + lane; idx < row_len; idx += 64) dw[idx] = srcrow[fold_pos(idx, k, kshift)] * gscale;
    return;
  }
  const float* __restrict__ v = params[jb.p_w] + static_cast<size_t>(row) * row_len;
  float dot = 0.f;
#pragma unroll 4
  for (int idx = half * 32 + lane; idx < row_len; idx += 64) dot = fmaf(srcrow[fold_pos(idx, k, kshift)], __ldg(v + idx), dot);
  dot = warp_sum(dot);
  if (lane == 0) s_part[warp] = dot;
  __syncthreads();
  dot = s_part[2 * r] + s_part[2 * r + 1];
  const float inv = 1.f / norms[jb.norm_off + row];
  const float gval = __ldg(params[jb.p_g] + row);
  if (half == 0 && lane == 0) dparams[jb.p_g][row] = dot * inv * gscale;
  const float a = gval * inv * gscale, bcoef = dot * inv * inv;
#pragma unroll 4
  for (int idx = half * 32 + lane; idx < row_len; idx += 64) dw[idx] = a * (srcrow[fold_pos(idx, k, kshift)] - __ldg(v + idx) * bcoef);
}

}  // namespace vcd

// Weight-norm fold (w = g * v / ||v||, reference: old-style torch.nn.utils.weight_norm used at
// vits/model/modules.py:10,190-199,229-230) into the packed operand layouts, and its backward.
//
// Every trainable convolution has one *logical* packed weight per direction,
//     Wp[tap][k][n]      (k = contraction channel, n = GEMM column; see ConvGeo in common.cuh)
// stored in up to two physical formats:
//     FMT_F32 : float [tap][k][n]                          (CUDA-core kernels)
//     FMT_TC  : bf16  [n / NT][tap][k / 8][n % NT][k % 8]  (tcgen05 B operand, K-major no-swizzle core
//                                                           matrices: 8 k-values = 16 B contiguous)
// All layers are handled by ONE launch per phase through a device-side job table (launch count matters:
// 76 weight-normed tensors per step).
#pragma once
#include "common.cuh"

namespace vcd {

enum : int { SRC_CONV_FWD = 0, SRC_CONV_DGRAD = 1, SRC_CONVT_FWD = 2, SRC_CONVT_DGRAD = 3 };
enum : int { FMT_F32 = 0, FMT_TC = 1 };

// Map a logical packed index (j, c, n) to the flat index into the parameter tensor (and its dim-0 row).
// Returns false if the element is structurally zero (ConvTranspose phase padding).
struct WeightMap {
  int src;           // SRC_*
  int cin, cout, k;  // parameter geometry: Conv1d [cout][cin][k], ConvTranspose1d [cin][cout][k]
  int u;             // upsample rate (ConvTranspose only)
  __host__ __device__ bool operator()(int j, int c, int n, int& row, size_t& idx) const {
    switch (src) {
      case SRC_CONV_FWD:  // Wp[j][ci][co] = w[co][ci][j]
        row = n;
        idx = (static_cast<size_t>(n) * cin + c) * k + j;
        return true;
      case SRC_CONV_DGRAD:  // Wp[j][co][ci] = w[co][ci][k-1-j]
        row = c;
        idx = (static_cast<size_t>(c) * cin + n) * k + (k - 1 - j);
        return true;
      case SRC_CONVT_FWD: {  // Wp[s][ci][r*cout+co] = w[ci][co][s*u+r]
        const int r = n / cout, co = n - r * cout, jj = j * u + r;
        if (jj >= k) return false;
        row = c;
        idx = (static_cast<size_t>(c) * cout + co) * k + jj;
        return true;
      }
      default: {  // SRC_CONVT_DGRAD: Wp[s][r*cout+co][ci] = w[ci][co][s*u+r]
        const int r = c / cout, co = c - r * cout, jj = j * u + r;
        if (jj >= k) return false;
        row = n;
        idx = (static_cast<size_t>(n) * cout + co) * k + jj;
        return true;
      }
    }
  }
};

struct NormJob {       // one per weight-normed parameter
  int p_v;             // index of weight_v in the parameter table
  int rows, row_len;   // dim0, prod(other dims)
  int norm_off;        // offset into the norms arena
  int first_block;     // first blockIdx.x of this job (one block per row)
};

struct PackJob {
  WeightMap map;
  int p_w, p_g;        // parameter indices (p_g < 0: not weight-normed)
  int norm_off;
  int taps, K, N, NT;  // logical dims; NT = tensor-core column tile (FMT_TC only)
  int fmt;
  int round_bf16;      // FMT_F32 in bf16 mode: values rounded through bf16 (the FFMA kernels then see the tensor cores' operands)
  long long dst_off;   // element offset into the fp32 or bf16 packed arena
  long long numel;     // taps*K*N
  int first_block;     // first blockIdx.x (256 elements per block)
};

// grid = total rows, block = 128.
__global__ void __launch_bounds__(128)
wn_norm_kernel(const NormJob* __restrict__ jobs, int njobs, const float* const* __restrict__ params,
               float* __restrict__ norms) {
  int lo = 0, hi = njobs - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (jobs[mid].first_block <= static_cast<int>(blockIdx.x)) lo = mid; else hi = mid - 1;
  }
  const NormJob jb = jobs[lo];
  const int row = blockIdx.x - jb.first_block;
  const float* v = params[jb.p_v] + static_cast<size_t>(row) * jb.row_len;
  float s = 0.f;
  for (int i = threadIdx.x; i < jb.row_len; i += 128) s = fmaf(v[i], v[i], s);
  __shared__ float red[4];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) norms[jb.norm_off + row] = sqrtf(red[0] + red[1] + red[2] + red[3]);
}

// grid = total 256-thread blocks over all jobs, block = 256.  FMT_TC jobs: one thread per 8-element (16-byte) group of
// the destination, so the bf16 stores are fully coalesced; FMT_F32 jobs: one thread per element.
__global__ void __launch_bounds__(256)
wn_pack_kernel(const PackJob* __restrict__ jobs, int njobs, const float* const* __restrict__ params,
               const float* __restrict__ norms, float* __restrict__ arena_f32, bf16* __restrict__ arena_bf16) {
  int lo = 0, hi = njobs - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (jobs[mid].first_block <= static_cast<int>(blockIdx.x)) lo = mid; else hi = mid - 1;
  }
  const PackJob jb = jobs[lo];
  const long long e = static_cast<long long>(blockIdx.x - jb.first_block) * 256 + threadIdx.x;
  const float* __restrict__ w = params[jb.p_w];
  const float* __restrict__ gv = jb.p_g >= 0 ? params[jb.p_g] : nullptr;   // NULL weight_g: baked weight, no scale
  if (jb.fmt == FMT_F32) {
    if (e >= jb.numel) return;
    const int n = static_cast<int>(e % jb.N);
    const long long r = e / jb.N;
    const int c = static_cast<int>(r % jb.K);
    const int j = static_cast<int>(r / jb.K);
    int row; size_t idx;
    float val = 0.f;
    if (jb.map(j, c, n, row, idx)) {
      val = w[idx];
      if (gv) val *= gv[row] / norms[jb.norm_off + row];
    }
    if (jb.round_bf16) val = __bfloat162float(__float2bfloat16_rn(val));
    arena_f32[jb.dst_off + e] = val;
    return;
  }
  // [n/NT][tap][k/8][n%NT][k%8], 8 consecutive k per thread
  if (e * 8 >= jb.numel) return;
  long long r = e;
  const int nl = static_cast<int>(r % jb.NT); r /= jb.NT;
  const int cg = static_cast<int>(r % (jb.K >> 3)); r /= (jb.K >> 3);
  const int j = static_cast<int>(r % jb.taps);
  const int nt = static_cast<int>(r / jb.taps);
  const int n = nt * jb.NT + nl;
  float v[8];
#pragma unroll
  for (int c8 = 0; c8 < 8; ++c8) {
    int row; size_t idx;
    float val = 0.f;
    if (jb.map(j, cg * 8 + c8, n, row, idx)) {
      val = __ldg(w + idx);
      if (gv) val *= __ldg(gv + row) / norms[jb.norm_off + row];
    }
    v[c8] = val;
  }
  store8<bf16>(arena_bf16 + jb.dst_off + e * 8, v);
}

// ---- backward: packed weight gradients -> parameter gradients -----------------------------------------
//   weight-normed:  dg[row] = <dw_row, v_row> / ||v_row|| ;  dv = (g/||v||) * (dw - v * <dw,v> / ||v||^2)
//   plain weight :  dw gathered from the packed gradient
//   copy job     :  dst = src (biases and tensors whose gradient is produced in parameter layout)
struct UnfoldJob {
  WeightMap map;       // SRC_CONV_FWD or SRC_CONVT_FWD (gradients are produced in forward geometry)
  int kind;            // 0 = weight row job, 1 = copy job
  int p_w, p_g;        // destination parameter indices
  int norm_off;
  int rows, row_len;   // weight row job: dim0 and prod(other dims); copy job: rows = ceil(numel/256)
  int K, N;            // logical dims of dWp
  long long src_off;   // offset into the gradient scratch arena
  long long numel;     // copy job
  int first_block;
};

__device__ __forceinline__ float unfold_fetch(const UnfoldJob& jb, const float* __restrict__ dwp, int row, int i) {
  // parameter element (row, i) -> packed (j, c, n)
  int j, c, n;
  if (jb.map.src == SRC_CONV_FWD) {       // w[co=row][ci][jj]
    const int ci = i / jb.map.k, jj = i - ci * jb.map.k;
    j = jj; c = ci; n = row;
  } else {                                // w[ci=row][co][jj] ; Wp[s][ci][r*cout+co]
    const int co = i / jb.map.k, jj = i - co * jb.map.k;
    j = jj / jb.map.u;
    c = row;
    n = (jj - j * jb.map.u) * jb.map.cout + co;
  }
  return dwp[jb.src_off + (static_cast<size_t>(j) * jb.K + c) * jb.N + n];
}

// grid = total blocks over all jobs, block = 256.
__global__ void __launch_bounds__(256)
wn_unfold_kernel(const UnfoldJob* __restrict__ jobs, int njobs, const float* const* __restrict__ params,
                 float* const* __restrict__ dparams, const float* __restrict__ norms,
                 const float* __restrict__ scratch, int block_base, float gscale) {
  // block_base: first_block of jobs[0] when only a tail slice of a segment's table is launched
  const int blk = static_cast<int>(blockIdx.x) + block_base;
  int lo = 0, hi = njobs - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (jobs[mid].first_block <= blk) lo = mid; else hi = mid - 1;
  }
  const UnfoldJob jb = jobs[lo];
  const int row = blk - jb.first_block;
  if (jb.kind == 1) {
    const long long e = static_cast<long long>(row) * 256 + threadIdx.x;
    if (e < jb.numel) dparams[jb.p_w][e] = scratch[jb.src_off + e] * gscale;
    return;
  }
  float* dw = dparams[jb.p_w] + static_cast<size_t>(row) * jb.row_len;
  if (jb.p_g < 0 || params[jb.p_g] == nullptr) {   // plain weight (never weight-normed, or remove_weight_norm was called)
    for (int i = threadIdx.x; i < jb.row_len; i += 256) dw[i] = unfold_fetch(jb, scratch, row, i) * gscale;
    return;
  }
  const float* v = params[jb.p_w] + static_cast<size_t>(row) * jb.row_len;
  float dot = 0.f;
  for (int i = threadIdx.x; i < jb.row_len; i += 256) dot = fmaf(unfold_fetch(jb, scratch, row, i), v[i], dot);
  __shared__ float red[8];
  dot = warp_sum(dot);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = dot;
  __syncthreads();
  dot = red[0] + red[1] + red[2] + red[3] + red[4] + red[5] + red[6] + red[7];
  const float norm = norms[jb.norm_off + row];
  const float gval = params[jb.p_g][row];
  const float inv = 1.f / norm;
  if (threadIdx.x == 0) dparams[jb.p_g][row] = dot * inv * gscale;
  const float a = gval * inv * gscale, bcoef = dot * inv * inv;
  for (int i = threadIdx.x; i < jb.row_len; i += 256) dw[i] = a * (unfold_fetch(jb, scratch, row, i) - v[i] * bcoef);
}

}  // namespace vcd

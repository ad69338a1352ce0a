// Mel / STFT loss tail of the generator step, directly behind the decoder's tanh:
//     y_spec_hat = spectrogram_torch_audio(y_hat, n_fft, sr, hop, win, center=False)   (vits/mel_processing.py:76-95)
//     y_mel_hat  = spec_to_mel_torch(y_spec_hat, n_fft, n_mel, sr, fmin, fmax)         (mel_processing.py:97-112)
//     loss_mel   = F.l1_loss(y_mel_hat, y_mel_slice) * c_mel                           (vits/light/vcvits.py:96-115)
// and its backward, which produces the decoder's upstream gradient dy.
//
// The 2048-point STFT of hop-512 frames is computed as a dense fp32 GEMM against a window-folded DFT basis (the
// magnitudes feed a log and the loss compares against fp32 targets, so the arithmetic stays fp32 FFMA: a bf16 DFT
// loses the low-energy bins):
//     S[r][2k], S[r][2k+1] = sum_n frame_r[n] * w[n] * (cos, -sin)(2 pi k n / n_fft)       (gemm 1, A gathered from y
//                                                                                           with the reflect padding)
//     mag = sqrt(re^2 + im^2 + 1e-6)                                                       (gemm 1 epilogue)
//     M = mag . melW^T;  lm = log(max(M, 1e-5));  |lm - target|, dM                        (gemm 2 + epilogue)
//     dmag = dM . melW;  dS = dmag * (re, im) / mag                                        (gemm 3 + epilogue, in place)
//     dframe = dS . basis                                                                   (gemm 4)
//     dy[t] = overlap-add of dframe over the frames (and reflected pad positions) that read sample t
// Every reduction has a fixed order: results are bit-identical run to run.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace vcd {
namespace mel {

constexpr int kBK = 16;

// ---- operand loaders: four consecutive elements along the contiguous direction ---------------------
// kContig = true : the operand is [rows][K] with K contiguous; load(row, k) -> element (row, k..k+3)
// kContig = false: the operand is [K][cols] with the cols contiguous; load(k, col) -> element (k, col..col+3)
struct RowsK {     // [rows][K], leading dimension ld (a multiple of 4, 16-byte aligned base)
  static constexpr bool kContig = true;
  const float* p; int ld; int rows; int K;
  __device__ __forceinline__ void load(int row, int k, float (&v)[4]) const {
    if (row < rows && k + 3 < K) {
      const float4 q = __ldg(reinterpret_cast<const float4*>(p + static_cast<size_t>(row) * ld + k));
      v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] = (row < rows && k + i < K) ? __ldg(p + static_cast<size_t>(row) * ld + k + i) : 0.f;
    }
  }
};
struct KCols {     // [K][cols], leading dimension ld (a multiple of 4)
  static constexpr bool kContig = false;
  const float* p; int ld; int K; int cols;
  __device__ __forceinline__ void load(int k, int col, float (&v)[4]) const {
    if (k < K && col + 3 < cols) {
      const float4 q = __ldg(reinterpret_cast<const float4*>(p + static_cast<size_t>(k) * ld + col));
      v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] = (k < K && col + i < cols) ? __ldg(p + static_cast<size_t>(k) * ld + col + i) : 0.f;
    }
  }
};
// STFT frames gathered from the waveform: row r = (b, f) reads y[b][reflect(f*hop + n - pad)]
// (torch.nn.functional.pad(..., mode='reflect') of mel_processing.py:90 / torchaudio.functional.spectrogram's pad)
struct Frames {
  static constexpr bool kContig = true;
  const float* y; int T; int F; int hop; int pad; int rows; int n_fft;
  __device__ __forceinline__ void load(int row, int k, float (&v)[4]) const {
    if (row >= rows) { v[0] = v[1] = v[2] = v[3] = 0.f; return; }
    const int b = row / F, f = row - b * F;
    const float* yb = y + static_cast<size_t>(b) * T;
    const int p0 = f * hop + k - pad;
    if (p0 >= 0 && p0 + 3 < T && k + 3 < n_fft) {
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] = __ldg(yb + p0 + i);
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int p = p0 + i;
        if (p < 0) p = -p;
        if (p >= T) p = 2 * (T - 1) - p;
        v[i] = (k + i < n_fft) ? __ldg(yb + p) : 0.f;
      }
    }
  }
};

// ---- epilogues: called with (row m, first of four consecutive columns n, the four sums) -----------
struct EpiSpectrum {   // gemm 1: columns (2k, 2k+1) = (re, im) of bin k
  static constexpr bool kReduce = false;
  float* S; int ldS;        // [rows][ldS] re/im interleaved, or null (no backward wanted)
  float* mag; int ldMag;    // [rows][ldMag]
  int rows; int n_bins;
  __device__ __forceinline__ float operator()(int m, int n, const float (&v)[4]) const {
    if (m >= rows) return 0.f;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int bin = (n >> 1) + h;
      if (bin < n_bins) {
        const float re = v[2 * h], im = v[2 * h + 1];
        mag[static_cast<size_t>(m) * ldMag + bin] = sqrtf(re * re + im * im + 1e-6f);
        if (S) *reinterpret_cast<float2*>(S + static_cast<size_t>(m) * ldS + 2 * bin) = make_float2(re, im);
      }
    }
    return 0.f;
  }
};
struct EpiLogMel {     // gemm 2: log-mel, optionally the L1 loss against the target and its gradient w.r.t. the mel energies
  static constexpr bool kReduce = true;
  float* out;            // [B][n_mel][F] log-mel (spectrogram mode) or null
  const float* target;   // [B][n_mel][F] or null
  float* dM; int ldM;    // [rows][ldM] gradient w.r.t. the (pre-log) mel energies, or null
  float scale;           // c_mel / (B * n_mel * F)
  int rows, n_mel, F;
  const long long* starts; int F_tgt;   // target [B][n_mel][F_tgt] read at frame starts[b] + f (slice_segments folded in), or null / F
  __device__ __forceinline__ float operator()(int m, int n, const float (&v)[4]) const {
    if (m >= rows) return 0.f;
    const int b = m / F, f = m - b * F;
    const long long ft = starts ? __ldg(starts + b) + f : f;
    float part = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = n + i;
      if (c >= n_mel) break;
      const float e = v[i];
      const float lm = logf(fmaxf(e, 1e-5f));            // dynamic_range_compression_torch (mel_processing.py:22-28)
      const size_t o = (static_cast<size_t>(b) * n_mel + c) * F + f;
      if (out) out[o] = lm;
      if (target) {
        const float d = lm - __ldg(target + (static_cast<size_t>(b) * n_mel + c) * F_tgt + ft);
        part += fabsf(d);
        const float sg = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
        dM[static_cast<size_t>(m) * ldM + c] = e >= 1e-5f ? sg * scale / e : 0.f;
      }
    }
    return part;
  }
};
struct EpiMagGrad {    // gemm 3: dmag -> (dre, dim) = dmag * (re, im) / mag, written over S
  static constexpr bool kReduce = false;
  float* S; int ldS;
  const float* mag; int ldMag;
  int rows, n_bins;
  __device__ __forceinline__ float operator()(int m, int n, const float (&v)[4]) const {
    if (m >= rows) return 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int bin = n + i;
      if (bin >= n_bins) break;
      float2* s = reinterpret_cast<float2*>(S + static_cast<size_t>(m) * ldS + 2 * bin);
      const float2 ri = *s;
      const float g = v[i] / __ldg(mag + static_cast<size_t>(m) * ldMag + bin);
      *s = make_float2(g * ri.x, g * ri.y);
    }
    return 0.f;
  }
};
struct EpiStore {      // gemm 4: plain store
  static constexpr bool kReduce = false;
  float* D; int ld; int rows, cols;
  __device__ __forceinline__ float operator()(int m, int n, const float (&v)[4]) const {
    if (m >= rows) return 0.f;
    if (n + 3 < cols) {
      *reinterpret_cast<float4*>(D + static_cast<size_t>(m) * ld + n) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (n + i < cols) D[static_cast<size_t>(m) * ld + n + i] = v[i];
    }
    return 0.f;
  }
};

// ---- tiled fp32 GEMM: D[M x N] = A[M x K] . B[K x N], 256 threads, (BM x BN) tile, (TM x TN) sums per thread -------
// TN = 8 is held as two groups of four columns BN/2 apart, so that the 16 threads of a tile row read 256 contiguous
// bytes of the B tile per shared-memory load (no bank conflicts).
template <int BM, int BN, int TM, int TN, class AL, class BL, class EP>
__global__ void __launch_bounds__(256) gemm_kernel(int K, AL al, BL bl, EP ep, float* partials) {
  static_assert((BM / TM) * (BN / TN) == 256, "256 threads");
  static_assert(TM == 4 && (TN == 4 || TN == 8), "micro tile");
  constexpr int TXN = BN / TN;          // threads along N
  constexpr int ACH = BM * kBK / 4 / 256, BCH = BN * kBK / 4 / 256;   // float4 chunks per thread and tile
  static_assert(ACH >= 1 && BCH >= 1, "tile too small");
  __shared__ __align__(16) float As[2][kBK][BM + 4];
  __shared__ __align__(16) float Bs[2][kBK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid % TXN, ty = tid / TXN;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;

  float ra[ACH][4], rb[BCH][4];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int c = 0; c < ACH; ++c) {
      const int ch = tid + c * 256;
      if constexpr (AL::kContig) al.load(m0 + ch / (kBK / 4), k0 + (ch % (kBK / 4)) * 4, ra[c]);
      else al.load(k0 + ch / (BM / 4), m0 + (ch % (BM / 4)) * 4, ra[c]);
    }
#pragma unroll
    for (int c = 0; c < BCH; ++c) {
      const int ch = tid + c * 256;
      if constexpr (BL::kContig) bl.load(n0 + ch / (kBK / 4), k0 + (ch % (kBK / 4)) * 4, rb[c]);
      else bl.load(k0 + ch / (BN / 4), n0 + (ch % (BN / 4)) * 4, rb[c]);
    }
  };
  auto stash = [&](int buf) {
#pragma unroll
    for (int c = 0; c < ACH; ++c) {
      const int ch = tid + c * 256;
      if constexpr (AL::kContig) {
        const int r = ch / (kBK / 4), kc = (ch % (kBK / 4)) * 4;
#pragma unroll
        for (int i = 0; i < 4; ++i) As[buf][kc + i][r] = ra[c][i];
      } else {
        const int k = ch / (BM / 4), mc = (ch % (BM / 4)) * 4;
        *reinterpret_cast<float4*>(&As[buf][k][mc]) = make_float4(ra[c][0], ra[c][1], ra[c][2], ra[c][3]);
      }
    }
#pragma unroll
    for (int c = 0; c < BCH; ++c) {
      const int ch = tid + c * 256;
      if constexpr (BL::kContig) {
        const int r = ch / (kBK / 4), kc = (ch % (kBK / 4)) * 4;
#pragma unroll
        for (int i = 0; i < 4; ++i) Bs[buf][kc + i][r] = rb[c][i];
      } else {
        const int k = ch / (BN / 4), nc = (ch % (BN / 4)) * 4;
        *reinterpret_cast<float4*>(&Bs[buf][k][nc]) = make_float4(rb[c][0], rb[c][1], rb[c][2], rb[c][3]);
      }
    }
  };

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  const int KT = (K + kBK - 1) / kBK;
  fetch(0);
  stash(0);
  __syncthreads();
  for (int kt = 0; kt < KT; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < KT) fetch((kt + 1) * kBK);
#pragma unroll
    for (int k = 0; k < kBK; ++k) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[buf][k][ty * TM]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w};
      float b[TN];
#pragma unroll
      for (int h = 0; h < TN / 4; ++h) {
        const float4 b4 = *reinterpret_cast<const float4*>(&Bs[buf][k][h * (BN / 2) + tx * 4]);
        b[4 * h] = b4.x; b[4 * h + 1] = b4.y; b[4 * h + 2] = b4.z; b[4 * h + 3] = b4.w;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < KT) stash(buf ^ 1);
    __syncthreads();
  }

  float part = 0.f;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
#pragma unroll
    for (int h = 0; h < TN / 4; ++h) {
      const float v[4] = {acc[i][4 * h], acc[i][4 * h + 1], acc[i][4 * h + 2], acc[i][4 * h + 3]};
      part += ep(m0 + ty * TM + i, n0 + h * (BN / 2) + tx * 4, v);
    }
  }
  if constexpr (EP::kReduce) {
    // fixed-order block sum: lanes by shuffle tree, then the eight warp sums in index order
    __shared__ float red[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if ((tid & 31) == 0) red[tid >> 5] = part;
    __syncthreads();
    if (tid == 0) {
      float s = 0.f;
      for (int w = 0; w < 8; ++w) s += red[w];
      partials[blockIdx.y * gridDim.x + blockIdx.x] = s;
    }
  }
}

// sum of the per-block loss partials in a fixed order (one block)
__global__ void __launch_bounds__(256) loss_finalize_kernel(const float* partials, int n, float scale, float* loss) {
  __shared__ float red[256];
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += 256) s += partials[i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (static_cast<int>(threadIdx.x) < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *loss = red[0] * scale;
}

// Backward of framing + reflect padding: dy[b][t] = sum over the padded positions p that read sample t (t + pad, and
// the mirrored positions in the two pads) of the overlap-add of dframe over the frames containing p.
__global__ void __launch_bounds__(256) overlap_add_kernel(const float* dframe, int n_fft, int hop, int pad, int F, int T,
                                                          long long total, float* dy) {
  const long long idx = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (idx >= total) return;
  const int b = static_cast<int>(idx / T), t = static_cast<int>(idx - static_cast<long long>(b) * T);
  int ps[3];
  int np = 0;
  ps[np++] = t + pad;
  if (t >= 1 && t <= pad) ps[np++] = pad - t;
  if (t >= T - 1 - pad && t <= T - 2) ps[np++] = pad + 2 * (T - 1) - t;
  float s = 0.f;
  for (int q = 0; q < np; ++q) {
    const int p = ps[q];
    int f_hi = p / hop;
    if (f_hi > F - 1) f_hi = F - 1;
    int f_lo = (p - n_fft + hop) / hop;          // smallest f with p - f*hop < n_fft
    if (p - n_fft + 1 <= 0) f_lo = 0;
    for (int f = f_lo; f <= f_hi; ++f) {
      const int n = p - f * hop;
      if (n >= 0 && n < n_fft) s += __ldg(dframe + (static_cast<size_t>(b) * F + f) * n_fft + n);
    }
  }
  dy[idx] = s;
}

// Window-folded DFT basis [2 * n_bins][n_fft]: row 2k = w[n] cos(2 pi k n / N), row 2k+1 = -w[n] sin(2 pi k n / N);
// w = periodic Hann window of win samples centred in the frame (torch.hann_window + torch.stft's window padding).
__global__ void __launch_bounds__(256) basis_kernel(float* basis, int n_fft, int win, int n_bins) {
  const long long idx = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (idx >= static_cast<long long>(n_bins) * n_fft) return;
  const int k = static_cast<int>(idx / n_fft), n = static_cast<int>(idx - static_cast<long long>(k) * n_fft);
  const int left = (n_fft - win) / 2;
  double w = 0.0;
  if (n >= left && n < left + win) w = 0.5 - 0.5 * cospi(2.0 * (n - left) / static_cast<double>(win));
  const long long kn = (static_cast<long long>(k) * n) % n_fft;
  double s, c;
  sincospi(2.0 * static_cast<double>(kn) / n_fft, &s, &c);
  basis[(2LL * k) * n_fft + n] = static_cast<float>(w * c);
  basis[(2LL * k + 1) * n_fft + n] = static_cast<float>(-w * s);
}


// ---------------------------------------------------------------------------------------------------
// Fast path (n_fft a power of two): ONE CTA per frame does the whole tail of that frame in shared memory --
//   gather + reflect pad + window  ->  radix-2 Stockham FFT (natural order in and out, one barrier per stage)
//   -> magnitudes -> banded mel filterbank (each filter only touches its own bins) -> log / L1 / d(mel energy)
//   -> banded transpose -> d(re, im) -> the SAME FFT applied to conj(dS) (the adjoint of the one-sided real DFT:
//      dframe[n] = Re sum_k (dRe_k - j dIm_k) e^(-2 pi j k n / N)) -> window -> dframe.
// The dense GEMM kernels above stay as the general path (any n_fft) and as the in-library cross-check.
// ---------------------------------------------------------------------------------------------------
struct FftParams {
  const float* y; int T, F, hop, pad, rows, n_fft, log2n, n_bins, n_mel;
  const float* window;        // [n_fft]
  const float2* tw;           // [n_fft - 1] per-stage twiddles: entry (ns - 1) + k = e^(-2 pi j k / (2 ns)), ns = 1, 2, 4, ... n_fft/2
  const int* f_lo; const int* f_cnt; const int* f_off; const float* f_val;   // per filter: first bin, bins, offset into f_val
  const int* b_lo; const int* b_cnt; const int* b_off; const float* b_val;   // per bin: first filter, filters, offset into b_val
  float* out;                 // [B][n_mel][F] log-mel, or null
  const float* target;        // [B][n_mel][F_tgt] read at frame starts[b] + f, or null
  const long long* starts;    // per-item first target frame (ids_slice), or null
  int F_tgt;
  float scale;
  float* dframe;              // [rows][n_fft], or null (no gradient wanted)
  float* partials;            // [rows]
};

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

// in-place view: transforms `src` (n complex values) using `dst` as the second buffer; returns the buffer holding the result
__device__ __forceinline__ float2* fft_stockham(float2* src, float2* dst, const float2* tw, int n, int log2n) {
  const int half = n >> 1;
  for (int s = 0; s < log2n; ++s) {
    const int ns = 1 << s;
    for (int j = threadIdx.x; j < half; j += blockDim.x) {
      const int k = j & (ns - 1);
      const float2 w = tw[ns - 1 + k];     // stage tables are contiguous in k: conflict-free (a strided n_fft/2 table is up to 32-way conflicted)
      const float2 v0 = src[j];
      const float2 v1 = cmul(src[j + half], w);
      const int j0 = ((j - k) << 1) + k;
      dst[j0] = make_float2(v0.x + v1.x, v0.y + v1.y);
      dst[j0 + ns] = make_float2(v0.x - v1.x, v0.y - v1.y);
    }
    __syncthreads();
    float2* t = src; src = dst; dst = t;
  }
  return src;
}

__global__ void __launch_bounds__(256) frame_fft_kernel(const FftParams P) {
  extern __shared__ __align__(16) uint8_t fft_smem[];
  const int n = P.n_fft, half = n >> 1;
  float2* bufA = reinterpret_cast<float2*>(fft_smem);
  float2* bufB = bufA + n;
  float2* tw = bufB + n;                                  // [n - 1] (+1 pad)
  float* mag = reinterpret_cast<float*>(tw + n);          // [n_bins]
  float* dMs = mag + ((P.n_bins + 3) & ~3);               // [n_mel]
  __shared__ float red[8];
  const int row = blockIdx.x, tid = threadIdx.x;
  const int b = row / P.F, f = row - b * P.F;
  const float* yb = P.y + static_cast<size_t>(b) * P.T;
  for (int t = tid; t < n - 1; t += 256) tw[t] = __ldg(P.tw + t);
  for (int i = tid; i < n; i += 256) {
    int p = f * P.hop + i - P.pad;
    if (p < 0) p = -p;
    if (p >= P.T) p = 2 * (P.T - 1) - p;
    bufA[i] = make_float2(__ldg(yb + p) * __ldg(P.window + i), 0.f);
  }
  __syncthreads();
  float2* X = fft_stockham(bufA, bufB, tw, n, P.log2n);
  float2* other = X == bufA ? bufB : bufA;
  for (int k = tid; k < P.n_bins; k += 256) mag[k] = sqrtf(X[k].x * X[k].x + X[k].y * X[k].y + 1e-6f);
  __syncthreads();
  float part = 0.f;
  for (int m = tid; m < P.n_mel; m += 256) {
    const int lo = __ldg(P.f_lo + m), cnt = __ldg(P.f_cnt + m);
    const float* v = P.f_val + __ldg(P.f_off + m);
    float e = 0.f;
    for (int i = 0; i < cnt; ++i) e = fmaf(__ldg(v + i), mag[lo + i], e);
    const float lm = logf(fmaxf(e, 1e-5f));
    const size_t o = (static_cast<size_t>(b) * P.n_mel + m) * P.F + f;
    if (P.out) P.out[o] = lm;
    if (P.target) {
      const long long ft = P.starts ? __ldg(P.starts + b) + f : f;
      const float d = lm - __ldg(P.target + (static_cast<size_t>(b) * P.n_mel + m) * P.F_tgt + ft);
      part += fabsf(d);
      const float sg = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
      dMs[m] = e >= 1e-5f ? sg * P.scale / e : 0.f;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  if ((tid & 31) == 0) red[tid >> 5] = part;
  __syncthreads();                       // also: dMs complete
  if (tid == 0) {
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += red[w];
    P.partials[row] = s;
  }
  if (!P.dframe) return;
  for (int k = tid; k < n; k += 256) {
    float2 z = make_float2(0.f, 0.f);
    if (k < P.n_bins) {
      const int lo = __ldg(P.b_lo + k), cnt = __ldg(P.b_cnt + k);
      const float* v = P.b_val + __ldg(P.b_off + k);
      float dmag = 0.f;
      for (int i = 0; i < cnt; ++i) dmag = fmaf(__ldg(v + i), dMs[lo + i], dmag);
      const float g = dmag / mag[k];
      z = make_float2(g * X[k].x, -g * X[k].y);
    }
    other[k] = z;
  }
  __syncthreads();
  const float2* R = fft_stockham(other, X, tw, n, P.log2n);
  float* dst = P.dframe + static_cast<size_t>(row) * n;
  for (int i = tid; i < n; i += 256) dst[i] = R[i].x * __ldg(P.window + i);
}

// window [n_fft] and per-stage twiddles [n_fft - 1] of the fast path
__global__ void __launch_bounds__(256) fft_tables_kernel(float* window, float2* tw, int n_fft, int win) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n_fft) return;
  const int left = (n_fft - win) / 2;
  double w = 0.0;
  if (i >= left && i < left + win) w = 0.5 - 0.5 * cospi(2.0 * (i - left) / static_cast<double>(win));
  window[i] = static_cast<float>(w);
  if (i < n_fft - 1) {       // entry i = (ns - 1) + k with ns the largest power of two <= i + 1
    int ns = 1;
    while (2 * ns <= i + 1) ns *= 2;
    const int k = i - (ns - 1);
    double s, c;
    sincospi(-static_cast<double>(k) / ns, &s, &c);          // e^(-2 pi j k / (2 ns))
    tw[i] = make_float2(static_cast<float>(c), static_cast<float>(s));
  }
}

}  // namespace mel
}  // namespace vcd

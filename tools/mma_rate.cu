// Microbenchmark: tcgen05.mma issue/execute rate for the operand layouts the conv / wgrad kernels use.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build_tools/mma_rate tools/mma_rate.cu   (also built by
//   __graft_entry__.build()); run on the GPU box: gpurun -- ./build_tools/mma_rate
// Prints clocks per MMA (issue loop, and issue loop + drain) for M x N x 16 bf16 MMAs with both operands in shared memory,
// for 1 and 2 CTAs per SM, 1..4 issuing warps, and three issue-loop shapes: unroll=0 runtime kk loop with one elect per MMA,
// unroll=1 one elected block of two MMAs per tap, unroll=2 one elected block per tile.  Results on B200 (this round):
//   128xNx16: 40 / 48 / 64 / 128 clk at N = 32 / 64 / 128 / 256 (= (128+N)/4: operand fetch at 128 B/clk; math-bound from N = 128),
//   never below ~23 clk per MMA, 85-87 clk per MMA with unroll=0 whatever the shape.
#include <cstdio>
#include <cstdlib>

#include "../vcvits_b200/csrc/tc_kernels.cuh"

using namespace vcd;
using namespace vcd::tc;

struct RateParams {
  int M, N, a_mn, b_mn;       // shape and operand majors
  uint32_t a_lbo, a_sbo, b_lbo, b_sbo;
  int swizzle;                // 0 none, 2 = 128B (layout_type field)
  int n_mma, taps, step16;    // MMAs per measurement; A start address cycles through `taps` shifts of step16 (16-B units)
  int kk, a_kk16, b_kk16;     // kk MMAs per tap advance descriptors by a_kk16/b_kk16
  int ncols;
  int nwarps;                 // issuing warps per CTA (each its own accumulator)
  int unroll;                 // 0: runtime loops, 1: kk unrolled (template), 2: taps x kk unrolled
  long long* out;             // [grid] clocks
};

__global__ void __launch_bounds__(128, 1) rate_kernel(const RateParams P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bars[4];
  __shared__ uint32_t slot;
  const int warp = __shfl_sync(~0u, static_cast<int>(threadIdx.x >> 5), 0);
  if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bars[i], 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(&slot, 512 / (gridDim.x > 148 ? 2 : 1));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = __shfl_sync(~0u, slot, 0) + warp * 2 * P.ncols;
  if (warp < P.nwarps) {
    const uint32_t idesc = make_idesc(P.M, P.N, P.a_mn, P.b_mn);
    uint64_t ad0 = make_desc(0, P.a_lbo, P.a_sbo) | (static_cast<uint64_t>(P.swizzle) << 61);
    uint64_t bd0 = make_desc(0, P.b_lbo, P.b_sbo) | (static_cast<uint64_t>(P.swizzle) << 61);
    const uint32_t a_hi = ad0 >> 32, b_hi = bd0 >> 32;
    const uint32_t a_base = static_cast<uint32_t>(ad0) + (smem_u32(smem) >> 4);
    const uint32_t b_base = static_cast<uint32_t>(bd0) + (smem_u32(smem + 64 * 1024) >> 4);
    const long long t0 = clock64();
    const int tiles = P.n_mma / (P.taps * P.kk);
    if (P.unroll == 0) {
      for (int it = 0; it < tiles; ++it) {
        const uint32_t d_tmem = tmem + static_cast<uint32_t>((it & 1) * P.ncols);
        uint32_t a_tap = a_base;
        for (int j = 0; j < P.taps; ++j) {
          uint32_t ad = a_tap, bd = b_base;
          for (int k = 0; k < P.kk; ++k) {
            if (elect_one()) umma_bf16_split(d_tmem, ad, a_hi, bd, b_hi, idesc, (j | k) ? 1u : 0u);
            ad += P.a_kk16;
            bd += P.b_kk16;
          }
          a_tap += P.step16;
        }
      }
    } else if (P.unroll == 1) {
      for (int it = 0; it < tiles; ++it) {
        const uint32_t d_tmem = tmem + static_cast<uint32_t>((it & 1) * P.ncols);
        uint32_t a_tap = a_base;
        for (int j = 0; j < P.taps; ++j) {
          if (elect_one()) {
            umma_bf16_split(d_tmem, a_tap, a_hi, b_base, b_hi, idesc, j ? 1u : 0u);
            umma_bf16_split(d_tmem, a_tap + P.a_kk16, a_hi, b_base + P.b_kk16, b_hi, idesc, 1u);
          }
          a_tap += P.step16;
        }
      }
    } else {
      for (int it = 0; it < tiles; ++it) {
        const uint32_t d_tmem = tmem + static_cast<uint32_t>((it & 1) * P.ncols);
        if (elect_one()) {
#pragma unroll
          for (int j = 0; j < 11; ++j) {
            umma_bf16_split(d_tmem, a_base + j * P.step16, a_hi, b_base, b_hi, idesc, j ? 1u : 0u);
            umma_bf16_split(d_tmem, a_base + j * P.step16 + P.a_kk16, a_hi, b_base + P.b_kk16, b_hi, idesc, 1u);
          }
        }
        __syncwarp();
      }
    }
    const long long t1 = clock64();
    if (elect_one()) umma_commit(&bars[warp]);
    __syncwarp();
    mbar_wait(&bars[warp], 0);
    const long long t2 = clock64();
    if (threadIdx.x == 0) { P.out[2 * blockIdx.x] = t1 - t0; P.out[2 * blockIdx.x + 1] = t2 - t0; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(__shfl_sync(~0u, slot, 0), 512 / (gridDim.x > 148 ? 2 : 1));
}

static void run(const char* name, RateParams P, int ctas_per_sm) {
  long long* d;
  const int grid = 148 * ctas_per_sm;
  cudaMalloc(&d, grid * 2 * sizeof(long long));
  P.out = d;
  const int smem = ctas_per_sm == 1 ? 200 * 1024 : 110 * 1024;
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  // operands: A region at +0, B region at +64 KB (both fit the 110 KB of the two-CTAs-per-SM runs)
  for (int rep = 0; rep < 2; ++rep) rate_kernel<<<grid, 128, smem>>>(P);
  cudaError_t err = cudaDeviceSynchronize();
  if (err != cudaSuccess) { printf("%-44s ERROR %s\n", name, cudaGetErrorString(err)); exit(1); }
  long long h[2 * 296];
  cudaMemcpy(h, d, grid * 2 * sizeof(long long), cudaMemcpyDeviceToHost);
  double issue = 0, total = 0;
  for (int i = 0; i < grid; ++i) { issue += h[2 * i]; total += h[2 * i + 1]; }
  issue /= grid; total /= grid;
  const double flop = 2.0 * P.M * P.N * 16;
  printf("%-44s ctas/sm=%d  issue %6.1f clk/mma  total %6.1f clk/mma  (%5.0f flop/clk/SM)\n", name, ctas_per_sm,
         issue / P.n_mma, total / P.n_mma, flop * P.n_mma * ctas_per_sm * P.nwarps / total);
  cudaFree(d);
}

int main() {
  const int RA = 144;
  for (int occ = 1; occ <= 2; ++occ)
    for (int nw = 1; nw <= 2; ++nw)
      for (int unroll = 0; unroll <= 2; ++unroll)
        for (int N : {32, 64}) {
          RateParams P{};
          P.M = 128; P.N = N; P.a_lbo = RA * 16; P.a_sbo = 128; P.b_lbo = N * 16; P.b_sbo = 128;
          P.n_mma = 22 * 40; P.taps = 11; P.step16 = 1; P.kk = 2; P.a_kk16 = 2 * RA; P.b_kk16 = 2 * N; P.ncols = N;
          P.nwarps = nw; P.unroll = unroll;
          char name[96];
          snprintf(name, sizeof name, "conv M128 N%-3d K32 warps=%d unroll=%d", N, nw, unroll);
          run(name, P, occ);
        }
  // weight-gradient operand layout: both operands MN-major (time rows are the contraction, 16-byte rows, LBO = 128 B,
  // SBO = channel-group stride), as tc::wgrad_kernel reads them
  for (int M : {64, 128})
    for (int N : {32, 64, 128, 256}) {
      RateParams P{};
      P.M = M; P.N = N; P.a_mn = 1; P.b_mn = 1;
      P.a_lbo = 128; P.a_sbo = 2048; P.b_lbo = 128; P.b_sbo = 2048;
      P.n_mma = 22 * 40; P.taps = 11; P.step16 = 0; P.kk = 2; P.a_kk16 = 16; P.b_kk16 = 16; P.ncols = N;
      P.nwarps = 1; P.unroll = 2;
      char name[96];
      snprintf(name, sizeof name, "wgrad MN-major M%-3d N%-3d", M, N);
      run(name, P, 1);
    }
  // same shapes, both operands K-major (for comparison)
  for (int M : {64, 128})
    for (int N : {32, 64, 128, 256}) {
      RateParams P{};
      P.M = M; P.N = N;
      P.a_lbo = 2048; P.a_sbo = 128; P.b_lbo = N * 16; P.b_sbo = 128;
      P.n_mma = 22 * 40; P.taps = 11; P.step16 = 0; P.kk = 2; P.a_kk16 = 256; P.b_kk16 = 2 * N; P.ncols = N;
      P.nwarps = 1; P.unroll = 2;
      char name[96];
      snprintf(name, sizeof name, "K-major M%-3d N%-3d", M, N);
      run(name, P, 1);
    }
  // tiny MMAs: pure issue rate
  for (int nw = 1; nw <= 4; ++nw)
    for (int unroll = 0; unroll <= 2; ++unroll) {
      RateParams P{};
      P.M = 64; P.N = 8; P.a_lbo = RA * 16; P.a_sbo = 128; P.b_lbo = 8 * 16; P.b_sbo = 128;
      P.n_mma = 22 * 40; P.taps = 11; P.step16 = 1; P.kk = 2; P.a_kk16 = 2 * RA; P.b_kk16 = 2 * 8; P.ncols = 8;
      P.nwarps = nw; P.unroll = unroll;
      char name[96];
      snprintf(name, sizeof name, "tiny M64 N8 warps=%d unroll=%d", nw, unroll);
      run(name, P, 1);
    }
  return 0;
}

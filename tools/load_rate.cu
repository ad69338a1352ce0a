// Microbenchmark: per-SM global -> shared copy throughput of the bulk-copy engine on B200, with the issue pattern the
// conv / wgrad kernels use (converged warp, elect.sync-predicated issue), as a function of
//   * the copy size (1-D cp.async.bulk of 1 KB .. 64 KB; 3-D tensor-map boxes of [groups][128 rows][16 B]),
//   * the number of issuing warps per CTA (each with its own ring and barriers),
//   * the number of CTAs pulling (1 .. 148) and the ring depth.
// The conv / wgrad cost models assumed ~22-27 B/clk per SM; this tool shows where that number comes from.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build_tools/load_rate tools/load_rate.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../vcvits_b200/csrc/tc_conv.cuh"

using namespace vcd;
using namespace vcd::tc;

struct LoadParams {
  const uint8_t* src;
  size_t span;            // bytes walked per warp before wrapping
  uint32_t slot_bytes, copy_bytes;
  int depth, iters, nwarps;
  int tensor;             // 1: tensor-map boxes (copy_bytes = groups * 2048)
  int groups;
  long long* out;         // [grid * 4] clocks per warp
};

__global__ void __launch_bounds__(128, 1) stream_kernel(const __grid_constant__ CUtensorMap tm, const LoadParams P) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int warp = __shfl_sync(~0u, static_cast<int>(threadIdx.x >> 5), 0);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw) + warp * 32;
  uint8_t* ring = smem_raw + 1024 + static_cast<size_t>(warp) * P.depth * P.slot_bytes;
  if ((threadIdx.x & 31) == 0) {
    for (int i = 0; i < P.depth; ++i) mbar_init(&bars[i], 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (warp >= P.nwarps) return;
  const int gw = blockIdx.x * P.nwarps + warp;
  const uint8_t* base = P.src + static_cast<size_t>(gw) * P.span;
  const int per_slot = P.slot_bytes / P.copy_bytes;
  size_t off = 0;
  int rb = 0;                                   // tensor mode: row block (128 rows) inside batch item gw
  const int rblocks = static_cast<int>(P.span / (static_cast<size_t>(P.groups) * 2048));
  auto issue = [&](int s) {
    if (elect_one()) {
      mbar_expect_tx(&bars[s], P.slot_bytes);
      for (int c = 0; c < per_slot; ++c) {
        uint8_t* dst = ring + static_cast<size_t>(s) * P.slot_bytes + c * P.copy_bytes;
        if (P.tensor) tma_load_3d(&tm, &bars[s], dst, 2 * 128 * ((rb + c) % rblocks), 0, gw);
        else bulk_load(dst, base + ((off + static_cast<size_t>(c) * P.copy_bytes) % P.span), P.copy_bytes, &bars[s]);
      }
    }
    __syncwarp();
    off += P.slot_bytes;
    rb += per_slot;
  };
  const long long t0 = clock64();
  for (int s = 0; s < P.depth && s < P.iters; ++s) issue(s);
  for (int it = 0; it < P.iters; ++it) {
    const int s = it % P.depth;
    mbar_wait(&bars[s], (it / P.depth) & 1);
    if (it + P.depth < P.iters) issue(s);
  }
  if ((threadIdx.x & 31) == 0) P.out[blockIdx.x * 4 + warp] = clock64() - t0;
}

int main() {
  const size_t total = size_t(1) << 30;
  uint8_t* src;
  cudaMalloc(&src, total);
  cudaMemset(src, 1, total);
  long long* d_clk;
  cudaMalloc(&d_clk, 4096 * sizeof(long long));
  cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  printf("# mode grid warps copy_B slot_KB depth | GB/s total | GB/s per CTA | B/clk per CTA | clk per copy (per warp)\n");
  struct Cfg { int tensor, groups; uint32_t copy, slot; int depth, nwarps; };
  std::vector<Cfg> cfgs;
  for (uint32_t copy : {1024u, 2048u, 4096u, 8192u, 16384u, 32768u, 65536u}) {
    const uint32_t slot = copy < 16384 ? 16384 : copy;
    cfgs.push_back({0, 0, copy, slot, static_cast<int>((192 * 1024) / slot), 1});
  }
  cfgs.push_back({0, 0, 2048, 2048, 16, 1});        // one copy per barrier round
  cfgs.push_back({0, 0, 2048, 8192, 2, 1});         // shallow ring
  cfgs.push_back({0, 0, 16384, 16384, 2, 1});
  for (int w : {2, 4}) {
    cfgs.push_back({0, 0, 2048, 16384, 48 / w / 4, w});
    cfgs.push_back({0, 0, 16384, 16384, 12 / w, w});
    cfgs.push_back({0, 0, 32768, 32768, 6 / w, w});
  }
  for (int g : {1, 4, 8, 16}) cfgs.push_back({1, g, static_cast<uint32_t>(g) * 2048, static_cast<uint32_t>(g) * 2048 * (g == 1 ? 8 : 1), g == 1 ? 8 : 96 / g, 1});
  cfgs.push_back({1, 4, 8192, 8192, 4, 4});
  cfgs.push_back({1, 8, 16384, 16384, 3, 4});
  for (int grid : {1, 74, 148})
    for (const Cfg& c : cfgs) {
      LoadParams P{};
      P.src = src;
      P.tensor = c.tensor;
      P.groups = c.tensor ? c.groups : 1;
      P.span = size_t(1) << 20;                   // 1 MB per warp (tensor mode: one batch item of `groups` x 512 rows ... see map)
      P.slot_bytes = c.slot; P.copy_bytes = c.copy; P.depth = c.depth; P.nwarps = c.nwarps;
      const size_t bytes_per_warp = (size_t(8) << 20) / c.nwarps;
      P.iters = static_cast<int>(bytes_per_warp / c.slot);
      P.out = d_clk;
      CUtensorMap tm{};
      if (c.tensor) {
        // tensor viewed as [B = 1024][groups][rows = span / (groups * 16)][16 B]; box = [groups][128 rows][16 B]
        const cuuint64_t rows = P.span / (static_cast<size_t>(c.groups) * 16);
        const cuuint64_t dims[3] = {2 * rows, static_cast<cuuint64_t>(c.groups), 1024};
        const cuuint64_t strides[2] = {rows * 16, P.span};
        const cuuint32_t box[3] = {256, static_cast<cuuint32_t>(c.groups), 1};
        const cuuint32_t estr[3] = {1, 1, 1};
        if (!tc_encode_fn() || tc_encode_fn()(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, src, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
          printf("tensor map encode failed\n");
          return 1;
        }
      }
      const size_t smem = 1024 + size_t(c.depth) * c.slot * c.nwarps;
      if (smem > 220 * 1024 || size_t(grid) * c.nwarps * P.span > total) { printf("skip\n"); continue; }
      float best = 1e30f;
      long long clk = 0;
      for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        stream_kernel<<<grid, 128, smem>>>(tm, P);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) {
          best = ms;
          std::vector<long long> h(grid * 4);
          cudaMemcpy(h.data(), d_clk, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
          clk = 0;
          for (int i = 0; i < grid; ++i)
            for (int w = 0; w < c.nwarps; ++w) clk = h[i * 4 + w] > clk ? h[i * 4 + w] : clk;
        }
      }
      const cudaError_t e = cudaGetLastError();
      if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
      const double bytes_cta = double(bytes_per_warp) * c.nwarps;
      const double gbs = bytes_cta * grid / (best * 1e-3) / 1e9;
      printf("%s %4d %d %6u %4u %3d | %8.1f | %7.1f | %6.1f | %7.1f\n", c.tensor ? "tensor" : "bulk  ", grid, c.nwarps, c.copy, c.slot >> 10, c.depth, gbs,
             gbs / grid, bytes_cta / double(clk), double(clk) / (double(bytes_per_warp) / c.copy));
    }
  return 0;
}

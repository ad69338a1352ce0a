// tcgen05 / TMEM / TMA implicit-GEMM kernels for sm_100a (bf16 operands, fp32 accumulation in TMEM).
//
// Shared-memory operand layout (no swizzle, 8x(16 B) core matrices, see common.cuh):
//     A tile  : [K_blk/8][RA rows][8]   -- K-major operand, row r of channel group c at (c*RA + r)*16 B.
//               A dilated tap is the SAME tile read through a descriptor whose start address is advanced by
//               shift*16 B: descriptor SBO = 128 B makes rows uniformly 16 B apart, LBO = RA*16 B.
//     W stage : [K_blk/8][BN cols][8]   -- K-major B operand, LBO = BN*16 B, SBO = 128 B.
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer,
// warps 2..5 = epilogue (TMEM -> registers -> fused bias/mask/residual/leaky_relu -> global).
// The kernel is persistent: each CTA walks a static tile list; TMEM accumulators are double-buffered so the
// epilogue of tile i overlaps the MMAs of tile i+1.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace vcd {
namespace tc {

constexpr int kThreads = 192;
constexpr uint32_t kSpinLimit = 1u << 28;  // bounded waits: a broken pipeline traps instead of hanging the GPU

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0, spins = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) break;
    if (++spins > kSpinLimit) __trap();
  }
}

__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}

// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 x bf16 -> fp32, issued by ONE thread.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrives when every previously issued tcgen05.mma of this thread has completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// Shared-memory matrix descriptor, SWIZZLE_NONE, version 1 (sm_100).  All byte quantities multiples of 16.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr >> 4) & 0x3fffu);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3fffu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3fffu) << 32;
  d |= 1ull << 46;  // descriptor version
  return d;
}

// Instruction descriptor for kind::f16: D = fp32, A = B = bf16; majors: 0 = K-major, 1 = MN-major.
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(a_mn_major) << 15) |
         (static_cast<uint32_t>(b_mn_major) << 16) | (static_cast<uint32_t>(N >> 3) << 17) |
         (static_cast<uint32_t>(M >> 4) << 24);
}

struct Pipe {
  int stage = 0;
  uint32_t phase = 0;
  __device__ __forceinline__ void advance(int n) {
    if (++stage == n) { stage = 0; phase ^= 1; }
  }
};

// ---------------------------------------------------------------------------------------------------
// Forward / data-gradient convolution (generalised geometry with is == 1).
// ---------------------------------------------------------------------------------------------------
struct ConvParams {
  ConvGeo g;
  Epilogue e;
  const bf16* w;          // packed [N/BN][taps][K/8][BN][8]
  int B, Lin, Lq, Lout;
  int BN, MT, KB, RA;     // column tile, 128-row tiles per CTA tile, K block (channels), rows per A region
  int NA, NW;             // pipeline depths
  int n_tiles_n, n_mgroups, total_tiles;
  int minshift;           // min over taps of j*step (<= 0)
  uint32_t tmem_cols;
};

__global__ void __launch_bounds__(kThreads, 1)
conv_kernel(const __grid_constant__ CUtensorMap tmA, const ConvParams P) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const uint32_t a_region_bytes = static_cast<uint32_t>(P.KB / 8) * P.RA * 16;
  const uint32_t a_stage_bytes = a_region_bytes * P.MT;
  const uint32_t w_stage_bytes = static_cast<uint32_t>(P.KB / 8) * P.BN * 16;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  uint8_t* a_smem = smem;
  uint8_t* w_smem = a_smem + static_cast<size_t>(P.NA) * a_stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(w_smem + static_cast<size_t>(P.NW) * w_stage_bytes);
  uint64_t* fullA = bars;
  uint64_t* emptyA = fullA + P.NA;
  uint64_t* fullW = emptyA + P.NA;
  uint64_t* emptyW = fullW + P.NW;
  uint64_t* acc_full = emptyW + P.NW;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  if (threadIdx.x == 0) {
    for (int i = 0; i < P.NA; ++i) { mbar_init(&fullA[i], 1); mbar_init(&emptyA[i], 1); }
    for (int i = 0; i < P.NW; ++i) { mbar_init(&fullW[i], 1); mbar_init(&emptyW[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, P.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int kblocks = P.g.K / P.KB;
  const int kk_per_block = P.KB / 16;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      Pipe pa, pw;
      for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
        const int nt = tile % P.n_tiles_n;
        const int rest = tile / P.n_tiles_n;
        const int mg = rest % P.n_mgroups;
        const int b = rest / P.n_mgroups;
        const int row0 = mg * 128 * P.MT + P.g.off0 + P.minshift;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&emptyA[pa.stage], pa.phase ^ 1);
          mbar_expect_tx(&fullA[pa.stage], a_stage_bytes);
          for (int mt = 0; mt < P.MT; ++mt)
            tma_load_4d(&tmA, &fullA[pa.stage], a_smem + static_cast<size_t>(pa.stage) * a_stage_bytes + mt * a_region_bytes,
                        0, row0 + mt * 128, kb * (P.KB / 8), b);
          pa.advance(P.NA);
          for (int j = 0; j < P.g.taps; ++j) {
            mbar_wait(&emptyW[pw.stage], pw.phase ^ 1);
            mbar_expect_tx(&fullW[pw.stage], w_stage_bytes);
            const bf16* src = P.w + ((static_cast<size_t>(nt) * P.g.taps + j) * (P.g.K / 8) + static_cast<size_t>(kb) * (P.KB / 8)) * P.BN * 8;
            bulk_load(w_smem + static_cast<size_t>(pw.stage) * w_stage_bytes, src, w_stage_bytes, &fullW[pw.stage]);
            pw.advance(P.NW);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      const uint32_t idesc = make_idesc(128, P.BN, 0, 0);
      const uint32_t a_lbo = static_cast<uint32_t>(P.RA) * 16, w_lbo = static_cast<uint32_t>(P.BN) * 16;
      Pipe pa, pw;
      int it = 0;
      for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        mbar_wait(&acc_empty[buf], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&fullA[pa.stage], pa.phase);
          tc_fence_after();
          const uint32_t a_base = smem_u32(a_smem + static_cast<size_t>(pa.stage) * a_stage_bytes);
          for (int j = 0; j < P.g.taps; ++j) {
            mbar_wait(&fullW[pw.stage], pw.phase);
            tc_fence_after();
            const uint32_t w_base = smem_u32(w_smem + static_cast<size_t>(pw.stage) * w_stage_bytes);
            const uint32_t shift_bytes = static_cast<uint32_t>(j * P.g.step - P.minshift) * 16;
            for (int mt = 0; mt < P.MT; ++mt) {
              const uint32_t d_tmem = tmem_base + static_cast<uint32_t>((buf * P.MT + mt) * P.BN);
              for (int kk = 0; kk < kk_per_block; ++kk) {
                const uint64_t ad = make_desc(a_base + mt * a_region_bytes + kk * 2 * a_lbo + shift_bytes, a_lbo, 128);
                const uint64_t bd = make_desc(w_base + kk * 2 * w_lbo, w_lbo, 128);
                umma_bf16(d_tmem, ad, bd, idesc, (kb | j | kk) != 0 ? 1u : 0u);
              }
            }
            umma_commit(&emptyW[pw.stage]);
            pw.advance(P.NW);
          }
          umma_commit(&emptyA[pa.stage]);
          pa.advance(P.NA);
        }
        umma_commit(&acc_full[buf]);
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int quad = warp & 3;               // TMEM lane quadrant this warp may access
    const int row_in_tile = quad * 32 + lane;
    const Epilogue& e = P.e;
    int it = 0;
    for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x, ++it) {
      const int nt = tile % P.n_tiles_n;
      const int rest = tile / P.n_tiles_n;
      const int mg = rest % P.n_mgroups;
      const int b = rest / P.n_mgroups;
      const int buf = it & 1;
      mbar_wait(&acc_full[buf], (it >> 1) & 1);
      tc_fence_after();
      for (int mt = 0; mt < P.MT; ++mt) {
        const int q = (mg * P.MT + mt) * 128 + row_in_tile;
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>((buf * P.MT + mt) * P.BN);
        for (int c16 = 0; c16 < P.BN / 16; ++c16) {
          float acc[16];
          tmem_ld16(t_row + c16 * 16, acc);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int n0 = nt * P.BN + c16 * 16 + h * 8;
            const int r = n0 / P.g.creal;
            const int ch0 = n0 - r * P.g.creal;
            const int ro = q * P.g.os + r - P.g.p;
            if (q >= P.Lq || ro < 0 || ro >= P.Lout) continue;
            const size_t o = blk_off(b, ch0, ro, P.g.creal, P.Lout);
            float v[8];
#pragma unroll
            for (int n = 0; n < 8; ++n) v[n] = acc[h * 8 + n];
            if (e.bias) {
#pragma unroll
              for (int n = 0; n < 8; ++n) v[n] += __ldg(e.bias + ch0 + n);
            }
            if (e.bias2) {
#pragma unroll
              for (int n = 0; n < 8; ++n) v[n] += __ldg(e.bias2 + static_cast<size_t>(b) * P.g.creal + ch0 + n);
            }
            if (e.mask) {
              float m[8];
              load8<bf16>(reinterpret_cast<const bf16*>(e.mask) + o, m);
#pragma unroll
              for (int n = 0; n < 8; ++n) v[n] *= (m[n] > 0.f ? 1.f : e.mask_slope);
            }
#pragma unroll
            for (int n = 0; n < 8; ++n) v[n] *= e.scale;
            if (e.res) {
              float t[8];
              load8<float>(e.res + o, t);
#pragma unroll
              for (int n = 0; n < 8; ++n) v[n] += t[n];
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
              store8<bf16>(reinterpret_cast<bf16*>(e.out_t) + o, a);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, P.tmem_cols);
}

// ---------------------------------------------------------------------------------------------------
// Weight gradient:  dWp[j][c][n] += sum_t in[b][t + off0 + j*step][c] * dout[b][t][n]     (is = os = 1)
// GEMM view: D[M = c][N = n] per tap, contraction over time.  Both operands are read MN-major straight from
// the blocked layout (time rows are the 16-byte-strided K direction: LBO = 128 B, SBO = channel-group stride),
// a tap is again a +16*shift byte offset on the A descriptor.  One CTA = (channel tile, column tile, tap group,
// time slab); partial sums of different slabs are combined with fp32 atomics (dWp zeroed by the caller).
//   M == 128 : accumulator of tap tl at TMEM columns tl*NT, all 128 lanes.
//   M == 64  : two accumulators share NT columns (lanes +0 / +16 of every 32-lane quadrant).
//   pair     : K_conv == 32 -- rows 0..31 of a 64-row accumulator are tap 2*tl, rows 32..63 tap 2*tl+1 (the
//              A tile holds a second copy of the 4 channel groups displaced by `step` rows).
// ---------------------------------------------------------------------------------------------------
struct WgradParams {
  float* dwp;             // [taps][K][N] fp32
  int taps, K, N, step, off0, minshift;
  int B, L;
  int M, mch, n_mtiles;   // instruction M (128 | 64), channel groups per A tile, channel tiles
  int NT, n_ntiles;       // column tile (<= 256)
  int pair;               // 1: K == 32 tap pairing
  int TG, n_tgroups;      // accumulator slots (taps, or tap pairs) per CTA and number of groups
  int n_slots;            // total slots = pair ? ceil(taps/2) : taps
  int TK, RI;             // time rows per pipeline stage, rows of the A tile (TK + halo)
  int slab_rows, slabs_per_item;
  int NS;
  uint32_t tmem_cols;
};

__global__ void __launch_bounds__(kThreads, 1)
wgrad_kernel(const __grid_constant__ CUtensorMap tmIn, const __grid_constant__ CUtensorMap tmD, const WgradParams P) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t in_copy_bytes = static_cast<uint32_t>(P.mch) * P.RI * 16;
  const uint32_t in_bytes = in_copy_bytes * (P.pair ? 2 : 1);
  const uint32_t d_bytes = static_cast<uint32_t>(P.NT / 8) * P.TK * 16;
  const uint32_t stage_bytes = in_bytes + d_bytes;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + static_cast<size_t>(P.NS) * stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + P.NS;
  uint64_t* acc_full = empty + P.NS;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  if (threadIdx.x == 0) {
    for (int i = 0; i < P.NS; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, P.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // CTA coordinates
  int id = blockIdx.x;
  const int slab = id % P.slabs_per_item; id /= P.slabs_per_item;
  const int b = id % P.B; id /= P.B;
  const int tg = id % P.n_tgroups; id /= P.n_tgroups;
  const int ntile = id % P.n_ntiles; id /= P.n_ntiles;
  const int mtile = id;
  const int t_begin = slab * P.slab_rows;
  const int t_end = min(P.L, t_begin + P.slab_rows);
  const int kblocks = (t_end - t_begin + P.TK - 1) / P.TK;
  const int slot0 = tg * P.TG;
  const int nslots = min(P.TG, P.n_slots - slot0);

  if (warp == 0) {
    if (lane == 0) {
      Pipe ps;
      for (int kb = 0; kb < kblocks; ++kb) {
        const int t0 = t_begin + kb * P.TK;
        mbar_wait(&empty[ps.stage], ps.phase ^ 1);
        mbar_expect_tx(&full[ps.stage], stage_bytes);
        uint8_t* st = smem + static_cast<size_t>(ps.stage) * stage_bytes;
        tma_load_4d(&tmIn, &full[ps.stage], st, 0, t0 + P.off0 + P.minshift, mtile * P.mch, b);
        if (P.pair) tma_load_4d(&tmIn, &full[ps.stage], st + in_copy_bytes, 0, t0 + P.off0 + P.minshift + P.step, mtile * P.mch, b);
        tma_load_4d(&tmD, &full[ps.stage], st + in_bytes, 0, t0, ntile * (P.NT / 8), b);
        ps.advance(P.NS);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc(P.M, P.NT, 1, 1);
      const uint32_t a_sbo = static_cast<uint32_t>(P.RI) * 16, b_sbo = static_cast<uint32_t>(P.TK) * 16;
      Pipe ps;
      for (int kb = 0; kb < kblocks; ++kb) {
        mbar_wait(&full[ps.stage], ps.phase);
        tc_fence_after();
        const uint32_t a_base = smem_u32(smem + static_cast<size_t>(ps.stage) * stage_bytes);
        const uint32_t b_base = a_base + in_bytes;
        for (int tl = 0; tl < nslots; ++tl) {
          const int j = (slot0 + tl) * (P.pair ? 2 : 1);
          const uint32_t shift_bytes = static_cast<uint32_t>(j * P.step - P.minshift) * 16;
          uint32_t d_tmem;
          if (P.M == 128) d_tmem = tmem_base + static_cast<uint32_t>(tl * P.NT);
          else d_tmem = tmem_base + static_cast<uint32_t>((tl >> 1) * P.NT) + (static_cast<uint32_t>((tl & 1) * 16) << 16);
          for (int kk = 0; kk < P.TK / 16; ++kk) {
            const uint64_t ad = make_desc(a_base + shift_bytes + kk * 256, 128, a_sbo);
            const uint64_t bd = make_desc(b_base + kk * 256, 128, b_sbo);
            umma_bf16(d_tmem, ad, bd, idesc, (kb | kk) != 0 ? 1u : 0u);
          }
        }
        umma_commit(&empty[ps.stage]);
        ps.advance(P.NS);
      }
      umma_commit(acc_full);
    }
  } else {
    const int quad = warp & 3;
    mbar_wait(acc_full, 0);
    tc_fence_after();
    for (int tl = 0; tl < nslots; ++tl) {
      int j, c;            // tap and conv-K channel handled by this thread for this slot (j < 0: nothing)
      uint32_t col0;
      bool active = true;
      if (P.M == 128) {
        j = slot0 + tl;
        c = mtile * 128 + quad * 32 + lane;
        col0 = static_cast<uint32_t>(tl * P.NT);
      } else {
        // lanes 16..31 of a quadrant hold the odd slot of the column block
        if ((tl & 1) != (lane >> 4)) active = false;
        const int row = quad * 16 + (lane & 15);
        col0 = static_cast<uint32_t>((tl >> 1) * P.NT);
        if (P.pair) {
          j = (slot0 + tl) * 2 + (row >> 5);
          c = row & 31;
        } else {
          j = slot0 + tl;
          c = mtile * 64 + row;
        }
      }
      if (j >= P.taps) active = false;
      for (int c16 = 0; c16 < P.NT / 16; ++c16) {
        float acc[16];
        tmem_ld16(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + col0 + c16 * 16, acc);
        if (active) {
          float* dst = P.dwp + (static_cast<size_t>(j) * P.K + c) * P.N + ntile * P.NT + c16 * 16;
#pragma unroll
          for (int n = 0; n < 16; ++n) atomicAdd(dst + n, acc[n]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, P.tmem_cols);
}

}  // namespace tc
}  // namespace vcd

// tcgen05 / TMEM / TMA implicit-GEMM kernels for sm_100a (bf16 operands, fp32 accumulation in TMEM).
//
// Operands reach shared memory through the TMA bulk-copy engine (cp.async.bulk, SASS UBLKCP) completing on
// mbarriers: in the row-padded blocked layout every (channel group, row range) of a tile is one contiguous run.
// Shared-memory operand layout (no swizzle, 8x(16 B) core matrices, see common.cuh):
//     A tile  : [K_blk/8][RA rows][8]   -- K-major operand, row r of channel group c at (c*RA + r)*16 B.
//               A dilated tap is the SAME tile read through a descriptor whose start address is advanced by
//               shift*16 B: descriptor SBO = 128 B makes rows uniformly 16 B apart, LBO = RA*16 B.
//     W stage : [K_blk/8][BN cols][8]   -- K-major B operand, LBO = BN*16 B, SBO = 128 B.
// Two kernels share these helpers: conv_kernel (forward / data gradient; 352 threads: activation producer, weight
// producer, MMA issuer, eight epilogue warps) and wgrad_kernel (weight gradient; 192 threads: tensor-TMA producer,
// MMA issuer, four epilogue warps).  Both are persistent over a static tile / split list; conv_kernel's TMEM
// accumulators are double-buffered so the epilogue of tile i overlaps the MMAs of tile i+1.  The MMA issue loops are
// compile-time shaped (one elected block of straight-line UTCHMMA per tap): see tools/mma_rate.cu for the measured
// issue / operand-fetch rates that motivate it.
#pragma once
#include <cuda.h>

#include <type_traits>

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

// Tensor TMA, 3-D map over a row-padded blocked tensor viewed as 8-byte elements (dims: 2*rows, channel group,
// batch): ONE instruction fetches [groups][rows][16 B]; every (group) row run is a contiguous burst.
__device__ __forceinline__ void tma_load_3d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
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
// Same, with the descriptors given as (lo, hi) halves: the hi words (LBO/SBO/version) are loop invariant, so the
// issue loop only performs 32-bit adds on the start-address words.
__device__ __forceinline__ void umma_bf16_split(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                                uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
      ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrives when every previously issued tcgen05.mma of this thread has completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// ---- thread-block cluster helpers (weight stages shared by the CTAs of a cluster, see ConvParams::cl) ----
// The same arrive on the barrier at this shared-memory offset in every CTA of `mask`.
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
// One global read, delivered (data and complete_tx) to the same shared-memory offsets of every CTA of `mask`.
__device__ __forceinline__ void bulk_load_mc(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint16_t mask) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
               ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar)), "h"(mask)
               : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  // the wait is tied to the loaded registers (in/out operands) instead of a memory clobber, so independent
  // global loads may be scheduled across it
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]));
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// 32 consecutive accumulator columns of this thread's TMEM lane.
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// Split form: issue several TMEM loads back to back, then wait once.  tmem_wait* ties the wait to the loaded registers
// (in/out operands), so no use of them can be scheduled above it; a second wait after the first returns immediately.
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait16(uint32_t (&r)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]));
}

// Shared-memory matrix descriptor, SWIZZLE_NONE, version 1 (sm_100).  All byte quantities multiples of 16.
// Start-address field (16-byte units, 14 bits) of a shared-memory matrix descriptor.  In a cluster launch the shared::cta
// address of a CTA with a non-zero cluster rank carries the rank above the 256 KB window offset: adding the unmasked
// address to a descriptor word corrupts its leading-byte-offset field (found the hard way: rank 1 computed garbage).
template <bool CL>
__device__ __forceinline__ uint32_t desc_addr16(const void* p) { return CL ? (smem_u32(p) & 0x3ffffu) >> 4 : smem_u32(p) >> 4; }

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

// One elected lane of a CONVERGED warp.  The producer / MMA warps keep their control flow warp-uniform and
// predicate only the issuing instruction with this: descriptors, addresses and loop state then live in
// uniform registers (a divergent `if (lane == 0)` region forces an R2UR + per-lane ELECT loop around every
// UTCHMMA / UBLKCP -- measured ~150 clk per MMA issue and ~500 clk per bulk copy on B200).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

// Programmatic dependent launch: a kernel launched with the programmatic-stream-serialization attribute starts while
// its stream predecessor drains; pdl_wait() blocks until the predecessor has completed and its writes are visible
// (no-op without the attribute), pdl_launch() lets the successor's CTAs start their prologue.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

struct Pipe {
  int stage = 0;
  uint32_t phase = 0;
  __device__ __forceinline__ void advance(int n) {
    if (++stage == n) { stage = 0; phase ^= 1; }
  }
};

// ---------------------------------------------------------------------------------------------------
// Forward / data-gradient convolution (generalised geometry with is == 1).
//
// Warp roles (kConvThreads = 352): warp 0 = activation producer, warp 10 = weight producer, warp 1 = TMEM allocator + MMA issuer,
// warps 2..9 = epilogue (two warps per TMEM lane quadrant, interleaved over 16-column units).
// Weights: if the whole [taps][K][BN] set of the column tile fits (and every tile of the launch uses the same
// column tile) it is loaded ONCE per CTA and stays resident; otherwise it streams through a ring whose stages
// hold several taps of one K block (fewer barrier round trips than one tap per stage).
// ---------------------------------------------------------------------------------------------------
constexpr int kConvThreads = 352;

// Division by a launch-constant through a host-computed multiplier (valid for dividends < 2^31): the tile -> (batch
// item, row group, column tile) decode sits on every role's per-tile path, and a hardware integer division is a
// ~100-clock dependent chain for a lone warp.
struct FastDiv {
  uint32_t mul, shr, div;
  __host__ void init(int d) {
    div = static_cast<uint32_t>(d);
    if (d == 1) { mul = 0; shr = 0; return; }
    uint32_t lg = 0;
    while ((1u << lg) < div) ++lg;
    const uint32_t pw = 31 + lg;
    mul = static_cast<uint32_t>(((1ull << pw) + div - 1) / div);
    shr = pw - 32;
  }
  __device__ __forceinline__ int quot(int n) const {
    return div == 1 ? n : static_cast<int>(__umulhi(static_cast<uint32_t>(n), mul) >> shr);
  }
  __device__ __forceinline__ void divmod(int n, int& q, int& r) const {
    q = quot(n);
    r = n - q * static_cast<int>(div);
  }
};

struct ConvParams {
  ConvGeo g;
  Epilogue e;
  const bf16* in;         // blocked, row-padded activation [B][K/8][Lin + pads][8]
  const bf16* w;          // packed [N/BN][taps][K/8][BN][8]
  int B, Lin, Lq, Lout;
  int BN, MT, KB, RA;     // column tile, 128-row tiles per CTA tile, K block (channels), rows per A region
  int NA;                 // activation pipeline depth
  int w_resident;         // 1: all weights of the column tile live in shared memory for the whole kernel
  int TPS, NW;            // ring mode: taps per stage, stages
  int n_tiles_n, n_mgroups, total_tiles;
  FastDiv d_tiles_n, d_mgroups, d_creal;   // dividers by n_tiles_n, n_mgroups, g.creal
  int units_shift;        // log2(BN / 16)
  int minshift;           // min over taps of j*step (<= 0)
  int acc_bufs;           // 2: accumulators double-buffered (epilogue overlaps the next tile's MMAs); 1: MT*BN > 256
  uint32_t tmem_cols;
  int cl;                 // 1: launched as clusters of two CTAs that own the SAME column tile and neighbouring row groups: each
                          // CTA fetches half of every streamed weight stage and multicasts it to both (half the L2 -> SM weight
                          // traffic per CTA).  Tile index v -> pair q = v >> 1 (column tile, row-group pair), rank v & 1.
  int pdl_late;           // 1: release the stream successor after this CTA's last MMAs are issued (default: at its last tile's loads)
  int NE;                 // EPI_SMEM: stages of the epilogue-operand ring (mask / residual tiles fetched by the bulk-copy engine)
  int e_ops;              // operands per stage (mask, residual)
  uint32_t e_stage_bytes; // e_ops * MT * (BN/8) * 128 rows * 16 B
  // WG instantiations (data-gradient launches of the <= 64-channel ResBlock layers): the layer's WEIGHT gradient is
  // accumulated by the same CTA.  Both of its operands are already in shared memory -- the gradient tile with its tap
  // halo (this kernel's A operand) and the layer's forward input (staged as the leaky-ReLU mask operand) -- so the
  // separate weight-gradient launch and its second pass over the two tensors disappear.
  float* wg_dwp;          // [taps][cin][cout] fp32 accumulated with reductions (zeroed by the caller)
  float* wg_dbias;        // [cout] bias gradient (column sums of the gradient tile), or null
  int wg_r0, wg_rstep;    // row of the A region that pairs with input row 0 of the tile for tap 0, and its step per tap
  int wg_rc;              // row of the A region that is output row 0 of the tile (bias gradient)
  uint32_t wg_col0;       // first TMEM column of the weight-gradient accumulators
  unsigned long long* trace;  // debug: %globaltimer stamps of CTA 0 (null = off)
};

// tile index -> (batch item * row groups + row group, column tile)
template <bool CL>
__device__ __forceinline__ void tile_decode(const ConvParams& P, int tile, int& rest, int& nt) {
  if constexpr (CL) {
    int r2;
    P.d_tiles_n.divmod(tile >> 1, r2, nt);
    rest = 2 * r2 + (tile & 1);
  } else {
    P.d_tiles_n.divmod(tile, rest, nt);
  }
}

__device__ __forceinline__ void ktrace(unsigned long long* buf, int slot) {
  if (buf != nullptr && blockIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    buf[slot] = t;
  }
}

// Compile-time epilogue features: unused operand paths (and their registers) vanish from the specialisation.
// EPI_SMEM: the bf16 epilogue operands (mask, residual) of a tile are staged in shared memory by a producer warp
// (one 2 KB bulk copy per channel group and 128-row tile, several tiles ahead) instead of being loaded from global
// memory by the epilogue threads: the epilogue of the <= 64-channel layers is bound by the latency of those loads and
// by their 64-bit address arithmetic, not by a pipe (profiles/r01_ncu_conv_c32_full_summary.md).
enum : int { EPI_MASK = 1, EPI_RES = 2, EPI_RES2 = 4, EPI_RAW = 8, EPI_SMEM = 16 };

struct EpiLoads {
  uint4 mask, rest;
  float4 r2a, r2b;
};

__device__ __forceinline__ void unpack8(const uint4& raw, float (&f)[8]) {
  const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}

template <int F>
__device__ __forceinline__ void epi_prefetch(const Epilogue& e, uint32_t o, bool valid, EpiLoads& L) {
  if (!valid) return;
  if constexpr (!(F & EPI_SMEM)) {
    if constexpr (F & EPI_MASK) L.mask = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(e.mask) + o));
    if constexpr (F & EPI_RES) L.rest = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(e.res_t) + o));
  }
  if constexpr (F & EPI_RES2) {
    if (e.res2) {
      L.r2a = __ldg(reinterpret_cast<const float4*>(e.res2 + o));
      L.r2b = __ldg(reinterpret_cast<const float4*>(e.res2 + o) + 1);
    } else {
      L.r2a = make_float4(0.f, 0.f, 0.f, 0.f);
      L.r2b = L.r2a;
    }
  }
}

// bias8: the 8 bias values (bias + per-batch bias) of this channel group, staged in shared memory per tile -- a
// global bias load here would put a full memory latency on the critical path of every 16-column unit; null when the
// launch has no bias (data gradients).  plain_out: out_t = bf16(v), no activation (tscale == act_slope == 1).
// o: 32-bit element offset (the host refuses tensors of 2^31 elements or more).
template <int F>
__device__ __forceinline__ void epi_finish(const Epilogue& e, const ConvGeo& g, int b, int ro, int ch0, uint32_t o,
                                           const EpiLoads& L, const float* bias8, bool plain_out, float s_pos, float s_neg,
                                           float (&v)[8]) {
  if (bias8 != nullptr) {
    const float4 b0 = *reinterpret_cast<const float4*>(bias8);
    const float4 b1 = *reinterpret_cast<const float4*>(bias8 + 4);
    v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
    v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
  }
  if constexpr (F & EPI_MASK) {
    float m[8];
    unpack8(L.mask, m);
#pragma unroll
    for (int n = 0; n < 8; ++n) v[n] *= (m[n] > 0.f ? s_pos : s_neg);   // s_pos = scale, s_neg = mask_slope * scale
  }
  if constexpr (F & EPI_RES) {
    float t[8];
    unpack8(L.rest, t);
    if (e.res_inv == 1.f) {   // backward: the identity path of `x = xt + x` adds the gradient as is
#pragma unroll
      for (int n = 0; n < 8; ++n) v[n] += t[n];
    } else {                  // forward: residual stream recovered from the stored leaky_relu(x)
#pragma unroll
      for (int n = 0; n < 8; ++n) v[n] += (t[n] > 0.f ? t[n] : t[n] * e.res_inv);
    }
  }
  if constexpr (F & EPI_RES2) {
    v[0] += L.r2a.x; v[1] += L.r2a.y; v[2] += L.r2a.z; v[3] += L.r2a.w;
    v[4] += L.r2b.x; v[5] += L.r2b.y; v[6] += L.r2b.z; v[7] += L.r2b.w;
  }
  if constexpr (F & EPI_RAW) {
    if (e.out_raw) store8<float>(e.out_raw + o, v);
  }
  if (e.out_t) {
    if (!plain_out) {
#pragma unroll
      for (int n = 0; n < 8; ++n) {   // leaky_relu with 0 < slope <= 1 is max(x, slope * x)
        const float x = v[n] * e.tscale;
        v[n] = fmaxf(x, x * e.act_slope);
      }
    }
    if (e.zu > 0) {
      const int q = (ro + e.zp) / e.zu, r = (ro + e.zp) - q * e.zu;
      store8<bf16>(reinterpret_cast<bf16*>(e.out_t) + blk_off(b, r * g.creal + ch0, q, e.zu * g.creal, e.zLq), v);
    } else {
      store8<bf16>(reinterpret_cast<bf16*>(e.out_t) + o, v);
    }
  }
}

// UW: accumulator columns an epilogue warp handles per step (16 | 32).  The epilogue of the <= 64-channel layers is bound
// by instruction issue; one 32-column TMEM load per step halves the per-step bookkeeping (tile / unit decode, barrier
// checks, pipeline state).  UW = 32 is only instantiated for the variants without global epilogue operands (none, or all
// staged in shared memory) -- the register prefetch buffers of the other variants would double.
// LEAN: epilogue role for the plain convolution geometry of the resident-weight (<= 64-channel) ResBlock layers (os == 1,
// p == 0, one column tile, launch-constant bias, no global epilogue operands: F == 0 or EPI_SMEM without running sums).
// The generic epilogue decodes (tile -> batch item, row group, scatter phase) with three multiplier divisions and
// re-derives every offset per 16-column unit; the compiler places that warp-uniform arithmetic on the uniform datapath,
// whose long dependent chains (~100 UIMAD / USEL / LDCU per unit) cost more than the arithmetic of a 128 x 32 tile
// (in-kernel trace: 0.67 us per tile of which 0.13 us MMA).  Here the tile walk is incremental (b, row group advance by
// the grid stride), offsets are one IMAD per unit, and a warp sweeps all its units of a tile in one straight loop.
// CL: the launch runs as 2-CTA clusters with multicast weight stages (ConvParams::cl; streamed weights, generic epilogue).
// Compile-time so that the plain instantiations carry none of it (as runtime branches it cost 1.9 % of the training step).
template <int F, int UW = 16, bool LEAN = false, bool WG = false, bool CL = false>
__global__ void __launch_bounds__(kConvThreads, 1)
conv_kernel(const ConvParams P) {
  static_assert(!WG || (LEAN && (F & EPI_SMEM) && (F & EPI_MASK)), "WG rides on the lean, smem-staged data-gradient epilogues");
  static_assert(!CL || (!LEAN && !WG && !(F & EPI_SMEM)), "clusters: streamed weights, generic epilogue");
  extern __shared__ __align__(128) uint8_t smem_raw[];
  // weights resident in shared memory for the whole launch?  Known at compile time for the lean (always resident) and
  // cluster (always streamed) instantiations: their dead ring / resident code and its checks vanish.
  const int w_res = LEAN ? 1 : (CL ? 0 : P.w_resident);
  // shfl-broadcast warp index: provably warp-uniform, so role branches are uniform and the compiler may use the
  // uniform datapath (UR registers) for descriptor / address arithmetic inside them
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  if (threadIdx.x == 0) ktrace(P.trace, 0);

  const uint32_t a_region_bytes = static_cast<uint32_t>(P.KB / 8) * P.RA * 16;
  const uint32_t a_stage_bytes = a_region_bytes * P.MT;
  const uint32_t w_tap_bytes = static_cast<uint32_t>(P.KB / 8) * P.BN * 16;   // one tap of one K block
  const int kblocks = P.g.K / P.KB;
  const int kk_per_block = P.KB / 16;
  const int ngroups_w = (P.g.taps + P.TPS - 1) / P.TPS;   // weight stages per K block (ring mode)
  const uint32_t w_region_bytes = w_res ? w_tap_bytes * P.g.taps * kblocks : w_tap_bytes * P.TPS * P.NW;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  uint8_t* a_smem = smem;
  uint8_t* w_smem = a_smem + static_cast<size_t>(P.NA) * a_stage_bytes;
  uint8_t* e_smem = w_smem + w_region_bytes;                                     // EPI_SMEM: NE stages of epilogue operands
  uint64_t* bars = reinterpret_cast<uint64_t*>(e_smem + static_cast<size_t>(P.NE) * P.e_stage_bytes);
  uint64_t* fullA = bars;
  uint64_t* emptyA = fullA + P.NA;
  uint64_t* fullW = emptyA + P.NA;      // ring mode: NW entries; resident mode: entry 0 = "weights loaded"
  uint64_t* emptyW = fullW + 8;
  uint64_t* acc_full = emptyW + 8;
  uint64_t* acc_empty = acc_full + 2;
  uint64_t* fullE = acc_empty + 2;
  uint64_t* emptyE = fullE + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(emptyE + 8);
  float* bias_s = reinterpret_cast<float*>(tmem_slot + 4);   // [2][128]: per-tile bias (+ per-batch bias), double-buffered
  // WG: [8 channel groups][16 rows][16 B] of bf16 ones -- the A operand whose product with the gradient tile is its column sum
  uint8_t* ones_s = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(bias_s + 256) + 127) & ~uintptr_t(127));

  if (threadIdx.x == 0) {
    for (int i = 0; i < P.NA; ++i) { mbar_init(&fullA[i], 1); mbar_init(&emptyA[i], 1); }
    for (int i = 0; i < 8; ++i) { mbar_init(&fullW[i], 1); mbar_init(&emptyW[i], CL ? 2 : 1); }   // cluster: both CTAs' MMA warps release a slot
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 8); }
    for (int i = 0; i < 8; ++i) { mbar_init(&fullE[i], 1); mbar_init(&emptyE[i], WG ? 9 : 8); }   // WG: + the MMA warp's commit
    fence_barrier_init();
  }
  if constexpr (WG) {
    if (threadIdx.x >= 64 && threadIdx.x < 64 + 128)
      reinterpret_cast<uint4*>(ones_s)[threadIdx.x - 64] = make_uint4(0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u);
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc(tmem_slot, P.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if constexpr (CL) cluster_sync_all();   // the peer's barriers exist before anything is multicast to them
  // shfl-broadcast: provably warp-uniform, so TMEM addresses stay in uniform registers in the MMA issue loop
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  if (threadIdx.x == 0) ktrace(P.trace, 1);

  if (warp == 0) {
    // ===================== activation producer: converged warp, elected lane issues =====================
    // An A region is (KB/8) channel groups x RA rows; in the row-padded global layout each channel group's rows
    // are one contiguous run, so the region is KB/8 1-D bulk copies (TMA bulk engine, mbarrier complete_tx).
    const int cgs = P.KB / 8;
    const uint32_t cg_bytes = static_cast<uint32_t>(P.RA) * 16;
    const size_t cg_stride_g = static_cast<size_t>(padded_len(P.Lin)) * 8;   // elements between channel groups
    Pipe pa;
    pdl_wait();   // the activations are the stream predecessor's output (weights are older: their producer does not wait)
    for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
      // last tile of this CTA: the successor kernel may start its prologue (barriers, TMEM, weight loads) now
      if (!P.pdl_late && tile + static_cast<int>(gridDim.x) >= P.total_tiles) pdl_launch();
      int b, mg, rest_a, nt_a;
      tile_decode<CL>(P, tile, rest_a, nt_a);
      P.d_mgroups.divmod(rest_a, b, mg);
      const int row0 = mg * 128 * P.MT + P.g.off0 + P.minshift;
      // 128-row tiles that start beyond the last output row are not loaded (their accumulators are never stored)
      int mt_live = P.MT;
      while (mt_live > 1 && (mg * P.MT + mt_live - 1) * 128 >= P.Lq) --mt_live;
      for (int kb = 0; kb < kblocks; ++kb) {
        mbar_wait(&emptyA[pa.stage], pa.phase ^ 1);
        uint8_t* stage = a_smem + static_cast<size_t>(pa.stage) * a_stage_bytes;
        const bf16* src0 = P.in + blk_row(b, kb * cgs, row0, P.g.K, P.Lin);
        if (elect_one()) {
          mbar_expect_tx(&fullA[pa.stage], mt_live * cgs * cg_bytes);
          for (int mt = 0; mt < mt_live; ++mt)
            for (int cg = 0; cg < cgs; ++cg)
              bulk_load(stage + mt * a_region_bytes + cg * cg_bytes, src0 + cg * cg_stride_g + static_cast<size_t>(mt) * 128 * 8,
                        cg_bytes, &fullA[pa.stage]);
        }
        __syncwarp();
        if (tile == static_cast<int>(blockIdx.x) && kb == 0 && lane == 0) ktrace(P.trace, 2);
        pa.advance(P.NA);
      }
    }
  } else if (warp == 10) {
    // ===================== weight producer (own warp: never queues behind the activation copies) =====================
    if (w_res) {  // whole weight set of column tile 0 (n_tiles_n == 1), loaded once
      const uint32_t total = w_region_bytes;
      if (elect_one()) {
        mbar_expect_tx(&fullW[0], total);
        for (uint32_t off = 0; off < total; off += 65536u)
          bulk_load(w_smem + off, reinterpret_cast<const uint8_t*>(P.w) + off, min(65536u, total - off), &fullW[0]);
      }
      __syncwarp();
      if constexpr (F & EPI_SMEM) {
        // ===== epilogue-operand producer: the mask / residual tiles of every output tile, NE tiles ahead =====
        // os == 1, p == 0, one column tile (host-checked): output row == accumulator row, and a (channel group, 128-row
        // tile) of an operand is one contiguous 2 KB run of the row-padded layout.
        const int groups = P.BN / 8;
        const size_t cg_stride_e = static_cast<size_t>(padded_len(P.Lout)) * 8;
        const bf16* ops[2] = {nullptr, nullptr};
        int nops = 0;
        if constexpr (F & EPI_MASK) ops[nops++] = reinterpret_cast<const bf16*>(P.e.mask);
        if constexpr (F & EPI_RES) ops[nops++] = reinterpret_cast<const bf16*>(P.e.res_t);
        Pipe pe;
        pdl_wait();
        for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
          int b, mg;
          P.d_mgroups.divmod(tile, b, mg);
          int mt_live = P.MT;
          while (mt_live > 1 && (mg * P.MT + mt_live - 1) * 128 >= P.Lq) --mt_live;
          mbar_wait(&emptyE[pe.stage], pe.phase ^ 1);
          uint8_t* stage = e_smem + static_cast<size_t>(pe.stage) * P.e_stage_bytes;
          if (elect_one()) {
            mbar_expect_tx(&fullE[pe.stage], static_cast<uint32_t>(nops * mt_live * groups) * 2048u);
            for (int op = 0; op < nops; ++op) {
              const bf16* src0 = ops[op] + blk_row(b, 0, mg * P.MT * 128, P.g.creal, P.Lout);
              for (int mt = 0; mt < mt_live; ++mt)
                for (int cg = 0; cg < groups; ++cg)
                  bulk_load(stage + ((static_cast<size_t>(op) * P.MT + mt) * groups + cg) * 2048, src0 + cg * cg_stride_e + static_cast<size_t>(mt) * 128 * 8,
                            2048u, &fullE[pe.stage]);
            }
          }
          __syncwarp();
          pe.advance(P.NE);
        }
      }
    } else {
      Pipe pw;
      const size_t tap_stride_g = static_cast<size_t>(P.g.K / 8) * P.BN * 8;
      const uint32_t crank = CL ? cluster_ctarank() : 0u;
      for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
        int rest, nt;
        tile_decode<CL>(P, tile, rest, nt);
        for (int kb = 0; kb < kblocks; ++kb) {
          for (int gi = 0; gi < ngroups_w; ++gi) {
            const int j0 = gi * P.TPS;
            const int nj = min(P.TPS, P.g.taps - j0);
            mbar_wait(&emptyW[pw.stage], pw.phase ^ 1);   // cluster: released by the MMA warps of BOTH CTAs
            uint8_t* dst = w_smem + static_cast<size_t>(pw.stage) * P.TPS * w_tap_bytes;
            const bf16* src = P.w + ((static_cast<size_t>(nt) * P.g.taps + j0) * (P.g.K / 8) +
                                     static_cast<size_t>(kb) * (P.KB / 8)) * P.BN * 8;
            if (elect_one()) {
              mbar_expect_tx(&fullW[pw.stage], nj * w_tap_bytes);
              if constexpr (CL) {          // this CTA's half of every copy, delivered to both CTAs (whose barriers each expect the whole stage)
                if (kblocks == 1) {
                  const uint32_t hb = (nj * w_tap_bytes) >> 1;
                  bulk_load_mc(dst + crank * hb, reinterpret_cast<const uint8_t*>(src) + crank * hb, hb, &fullW[pw.stage], 3);
                } else {
                  const uint32_t hb = w_tap_bytes >> 1;
                  for (int jj = 0; jj < nj; ++jj)
                    bulk_load_mc(dst + jj * w_tap_bytes + crank * hb,
                                 reinterpret_cast<const uint8_t*>(src + jj * tap_stride_g) + crank * hb, hb, &fullW[pw.stage], 3);
                }
              } else if (kblocks == 1) {  // taps are contiguous in the packed weights: one copy per stage
                bulk_load(dst, src, nj * w_tap_bytes, &fullW[pw.stage]);
              } else {
                for (int jj = 0; jj < nj; ++jj)
                  bulk_load(dst + jj * w_tap_bytes, src + jj * tap_stride_g, w_tap_bytes, &fullW[pw.stage]);
              }
            }
            __syncwarp();
            pw.advance(P.NW);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: converged warp, elected lane issues =====================
    const uint32_t idesc = make_idesc(128, P.BN, 0, 0);
    const uint32_t a_lbo = static_cast<uint32_t>(P.RA) * 16, w_lbo = static_cast<uint32_t>(P.BN) * 16;
    // descriptor arithmetic in 16-byte units on the low word (start-address field)
    const uint32_t a_region16 = a_region_bytes >> 4, a_kk16 = (2 * a_lbo) >> 4, w_kk16 = (2 * w_lbo) >> 4;
    const uint32_t w_tap16 = w_tap_bytes >> 4;
    const uint64_t a_desc0 = make_desc(0, a_lbo, 128), w_desc0 = make_desc(0, w_lbo, 128);
    Pipe pa, pw;
    int it = 0;
    if (w_res) {
      mbar_wait(&fullW[0], 0);
      tc_fence_after();
      if (lane == 0) ktrace(P.trace, 3);
    }
    // KK (MMAs per K block and tap) and MT (row tiles per CTA tile) are compile-time in the common cases: the issue
    // loop is then one elected region of MT*KK straight-line MMAs per tap with loop-invariant uniform increments.
    // A per-MMA elect + runtime kk/mt loops cost 85-125 clk per MMA; this form issues at the tensor-pipe rate
    // (40 clk for 128x32x16, operand-fetch bound -- tools/mma_rate.cu).
    auto issue_tiles = [&](auto kk_tag, auto mt_tag) {
      constexpr int KK = decltype(kk_tag)::value;
      constexpr int MTC = decltype(mt_tag)::value;     // 0 = runtime P.MT
      const uint32_t a_hi = static_cast<uint32_t>(a_desc0 >> 32), w_hi = static_cast<uint32_t>(w_desc0 >> 32);
      const uint32_t tap_stride16 = w_res ? w_tap16 * kblocks : w_tap16;   // resident layout: [tap][K/8][BN][8]
      const uint32_t w_res_lo = static_cast<uint32_t>(w_desc0) + desc_addr16<CL>(w_smem);
      const uint32_t a_step = static_cast<uint32_t>(P.g.step), bn = static_cast<uint32_t>(P.BN);
      const int taps = P.g.taps;
      for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x, ++it) {
        const int buf = P.acc_bufs == 2 ? (it & 1) : 0;
        const int use = P.acc_bufs == 2 ? (it >> 1) : it;      // how many times this buffer was used before
        mbar_wait(&acc_empty[buf], (use & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tile = tmem_base + static_cast<uint32_t>(buf * P.MT * P.BN);
        const int ngroups = w_res ? 1 : ngroups_w;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&fullA[pa.stage], pa.phase);
          tc_fence_after();
          if (it == 0 && kb == 0 && lane == 0) ktrace(P.trace, 4);
          const uint32_t a_stage_lo = static_cast<uint32_t>(a_desc0) + desc_addr16<CL>(a_smem + static_cast<size_t>(pa.stage) * a_stage_bytes);
          for (int gi = 0; gi < ngroups; ++gi) {
            int nj, j0;
            uint32_t w_lo;
            if (w_res) {
              nj = taps;
              j0 = 0;
              w_lo = w_res_lo + static_cast<uint32_t>(kb) * w_tap16;
            } else {
              j0 = gi * P.TPS;
              nj = min(P.TPS, taps - j0);
              mbar_wait(&fullW[pw.stage], pw.phase);
              tc_fence_after();
              w_lo = static_cast<uint32_t>(w_desc0) + desc_addr16<CL>(w_smem + static_cast<size_t>(pw.stage) * P.TPS * w_tap_bytes);
            }
            uint32_t a_tap = a_stage_lo + static_cast<uint32_t>(-P.minshift) + static_cast<uint32_t>(j0) * a_step;
            uint32_t first = (kb | gi) != 0 ? 1u : 0u;    // accumulate flag of the first MMA of this tap group
#pragma unroll 1
            for (int jj = 0; jj < nj; ++jj) {
              if constexpr (MTC > 0) {
                if (elect_one()) {
#pragma unroll
                  for (int mt = 0; mt < MTC; ++mt) {
#pragma unroll
                    for (int kk = 0; kk < KK; ++kk)
                      umma_bf16_split(d_tile + static_cast<uint32_t>(mt) * bn,
                                      a_tap + static_cast<uint32_t>(mt) * a_region16 + static_cast<uint32_t>(kk) * a_kk16, a_hi,
                                      w_lo + static_cast<uint32_t>(kk) * w_kk16, w_hi, idesc, kk == 0 ? first : 1u);
                  }
                }
              } else {
                uint32_t d_tmem = d_tile, a_mt = a_tap;
                for (int mt = 0; mt < P.MT; ++mt) {
                  if (elect_one()) {
#pragma unroll
                    for (int kk = 0; kk < KK; ++kk)
                      umma_bf16_split(d_tmem, a_mt + static_cast<uint32_t>(kk) * a_kk16, a_hi,
                                      w_lo + static_cast<uint32_t>(kk) * w_kk16, w_hi, idesc, kk == 0 ? first : 1u);
                  }
                  d_tmem += bn;
                  a_mt += a_region16;
                }
              }
              first = 1u;
              w_lo += tap_stride16;
              a_tap += a_step;
            }
            if (!w_res) {
              if (elect_one()) {
                if constexpr (CL) umma_commit_mc(&emptyW[pw.stage], 3);   // the slot is free for both producers once both consumers are done
                else umma_commit(&emptyW[pw.stage]);
              }
              pw.advance(P.NW);
            }
          }
          if constexpr (WG) {
            // (host: one K block, one 128-row tile per CTA tile)  The epilogue may start on the data gradient now; the
            // weight-gradient MMAs of this tile follow in the tensor pipe.
            if (elect_one()) umma_commit(&acc_full[buf]);
            const int es = it % P.NE;
            mbar_wait(&fullE[es], static_cast<uint32_t>(it / P.NE) & 1u);
            tc_fence_after();
            // MN-major operands (time rows are the contraction): A = the layer's forward input [cin][rows] from the staged
            // mask operand ([group][128 rows][16 B]), B = the gradient tile [cout][rows] from the activation ring, displaced
            // by the tap.  M = 64: accumulator slots pair up on the lane halves of a column block (see wgrad_kernel).
            const uint32_t idesc_w = make_idesc(64, P.g.K, 1, 1);   // (host: K == N == BN <= 64)
            const uint64_t x_desc = make_desc(smem_u32(e_smem + static_cast<size_t>(es) * P.e_stage_bytes), 128, 2048);
            const uint64_t g_desc = make_desc(smem_u32(a_smem + static_cast<size_t>(pa.stage) * a_stage_bytes), 128, a_lbo);
            const uint32_t x_lo = static_cast<uint32_t>(x_desc), x_hi = static_cast<uint32_t>(x_desc >> 32);
            const uint32_t g_hi = static_cast<uint32_t>(g_desc >> 32);
            const uint32_t acc0 = it != 0 ? 1u : 0u;
            uint32_t g_lo = static_cast<uint32_t>(g_desc) + static_cast<uint32_t>(P.wg_r0);
            const uint32_t wg_base = tmem_base + P.wg_col0;
#pragma unroll 1
            for (int j = 0; j < taps; ++j) {
              const uint32_t d_w = wg_base + static_cast<uint32_t>((j >> 1) * P.BN) + (static_cast<uint32_t>((j & 1) * 16) << 16);
              if (elect_one()) {
#pragma unroll
                for (int kk = 0; kk < 8; ++kk)
                  umma_bf16_split(d_w, x_lo + 16u * kk, x_hi, g_lo + 16u * kk, g_hi, idesc_w, kk == 0 ? acc0 : 1u);
              }
              g_lo += static_cast<uint32_t>(P.wg_rstep);
            }
            if (P.wg_dbias != nullptr) {   // bias gradient: ones x gradient tile -> column sums in slot `taps` (taps is odd: a free lane half)
              const uint32_t d_b = wg_base + static_cast<uint32_t>((taps >> 1) * P.BN) + (static_cast<uint32_t>((taps & 1) * 16) << 16);
              const uint64_t o_desc = make_desc(smem_u32(ones_s), 128, 256);
              const uint32_t gb_lo = static_cast<uint32_t>(g_desc) + static_cast<uint32_t>(P.wg_rc);
              if (elect_one()) {
#pragma unroll
                for (int kk = 0; kk < 8; ++kk)
                  umma_bf16_split(d_b, static_cast<uint32_t>(o_desc), static_cast<uint32_t>(o_desc >> 32), gb_lo + 16u * kk, g_hi, idesc_w,
                                  kk == 0 ? acc0 : 1u);
              }
            }
            if (elect_one()) {
              umma_commit(&emptyE[es]);
              umma_commit(&emptyA[pa.stage]);
            }
          } else {
            if (elect_one()) umma_commit(&emptyA[pa.stage]);
          }
          pa.advance(P.NA);
        }
        if constexpr (!WG) {
          if (elect_one()) umma_commit(&acc_full[buf]);
        }
        if (it < 4 && lane == 0) ktrace(P.trace, 8 + it);       // MMAs of tile `it` issued
      }
      if constexpr (WG) {
        if (elect_one()) umma_commit(&fullW[1]);   // every weight-gradient MMA of this CTA has completed
      }
    };
    using std::integral_constant;
    if (kk_per_block == 4) {
      if (P.MT == 1) issue_tiles(integral_constant<int, 4>{}, integral_constant<int, 1>{});
      else if (P.MT == 2) issue_tiles(integral_constant<int, 4>{}, integral_constant<int, 2>{});
      else if (P.MT == 4) issue_tiles(integral_constant<int, 4>{}, integral_constant<int, 4>{});
      else issue_tiles(integral_constant<int, 4>{}, integral_constant<int, 0>{});
    } else if (kk_per_block == 2) {
      if (P.MT == 1) issue_tiles(integral_constant<int, 2>{}, integral_constant<int, 1>{});
      else if (P.MT == 2) issue_tiles(integral_constant<int, 2>{}, integral_constant<int, 2>{});
      else if (P.MT == 4) issue_tiles(integral_constant<int, 2>{}, integral_constant<int, 4>{});
      else issue_tiles(integral_constant<int, 2>{}, integral_constant<int, 0>{});
    } else {
      issue_tiles(integral_constant<int, 1>{}, integral_constant<int, 0>{});
    }
    if (P.pdl_late) pdl_launch();
    if (lane == 0) ktrace(P.trace, 5);
  } else if (LEAN && warp >= 2 && warp <= 9) {
    // ===================== epilogue warps, plain geometry (see LEAN above) =====================
    if constexpr (LEAN) {
      static_assert(F == 0 || ((F & EPI_SMEM) && !(F & (EPI_RES2 | EPI_RAW))), "LEAN needs operand-free or smem-staged epilogues");
      constexpr int NG = UW / 8;
      const int quad = warp & 3, half = (warp - 2) >> 2;
      const int row_in_tile = quad * 32 + lane;
      const Epilogue& e = P.e;
      const int ushift = P.units_shift - (UW == 32 ? 1 : 0);
      const int n_units = P.MT << ushift;
      const int groups = P.BN >> 3;
      const uint32_t chunk_stride = static_cast<uint32_t>(padded_len(P.Lout)) * 8;
      const uint32_t b_stride = static_cast<uint32_t>(P.g.creal >> 3) * chunk_stride;
      const uint32_t tile_rows8 = static_cast<uint32_t>(P.MT) * 128u * 8u;
      const bool plain_out = e.tscale == 1.f && e.act_slope == 1.f;
      const float s_pos = e.scale, s_neg = e.mask_slope * e.scale;
      const bool has_bias = e.bias != nullptr;
      const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
      const uint32_t op_stride = static_cast<uint32_t>(P.MT) * groups * 2048u;
      const int grid = static_cast<int>(gridDim.x);
      if (has_bias) {     // launch-constant bias vector (host-checked: one column tile, no per-batch bias)
        const int et = static_cast<int>(threadIdx.x) - 64;
        if (et < P.BN) bias_s[et] = __ldg(e.bias + et);
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
      int b, mg;
      P.d_mgroups.divmod(static_cast<int>(blockIdx.x), b, mg);
      pdl_wait();
      int it = 0;
      for (int tile = blockIdx.x; tile < P.total_tiles; tile += grid, ++it) {
        const int buf = P.acc_bufs == 2 ? (it & 1) : 0;
        const int use_n = P.acc_bufs == 2 ? (it >> 1) : it;
        const int q0 = mg * P.MT * 128 + row_in_tile;
        const uint32_t tile_o = static_cast<uint32_t>(b) * b_stride + static_cast<uint32_t>(mg) * tile_rows8 +
                                static_cast<uint32_t>(row_in_tile + kPadL) * 8u;
        const uint32_t t_tile = t_lane + static_cast<uint32_t>(buf * P.MT * P.BN);
        int e_stage = 0;
        const uint8_t* e_base = nullptr;
        if constexpr (F & EPI_SMEM) {
          e_stage = it % P.NE;
          mbar_wait(&fullE[e_stage], static_cast<uint32_t>(it / P.NE) & 1u);
          e_base = e_smem + static_cast<size_t>(e_stage) * P.e_stage_bytes + static_cast<uint32_t>(row_in_tile) * 16u;
        }
        mbar_wait(&acc_full[buf], use_n & 1);
        tc_fence_after();
        for (int u = half; u < n_units; u += 2) {
          const int mt = u >> ushift, cu = u - (mt << ushift);
          float acc[UW];
          const uint32_t taddr = t_tile + static_cast<uint32_t>(mt * P.BN + cu * UW);
          if constexpr (UW == 16) tmem_ld16(taddr, acc);
          else tmem_ld32(taddr, acc);
          if (q0 + mt * 128 < P.Lq) {
            const uint32_t o = tile_o + static_cast<uint32_t>(mt) * 1024u + static_cast<uint32_t>(cu * NG) * chunk_stride;
#pragma unroll
            for (int h = 0; h < NG; ++h) {
              float v[8];
#pragma unroll
              for (int n = 0; n < 8; ++n) v[n] = acc[h * 8 + n];
              EpiLoads L;
              if constexpr (F & EPI_SMEM) {
                const uint8_t* ep = e_base + static_cast<uint32_t>(mt * groups + cu * NG + h) * 2048u;
                if constexpr (F & EPI_MASK) {
                  L.mask = *reinterpret_cast<const uint4*>(ep);
                  ep += op_stride;
                }
                if constexpr (F & EPI_RES) L.rest = *reinterpret_cast<const uint4*>(ep);
              }
              epi_finish<F>(e, P.g, 0, 0, 0, o + static_cast<uint32_t>(h) * chunk_stride, L,
                            has_bias ? bias_s + cu * UW + h * 8 : nullptr, plain_out, s_pos, s_neg, v);
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&acc_empty[buf]);
          if constexpr (F & EPI_SMEM) mbar_arrive(&emptyE[e_stage]);
        }
        mg += grid;
        while (mg >= P.n_mgroups) { mg -= P.n_mgroups; ++b; }
      }
      if constexpr (WG) {
        // ---- weight-gradient accumulators -> global (fp32 reductions; the CTAs of the launch hold partial sums) ----
        // Slot s (tap s; slot `taps` = bias gradient) lives at columns (s >> 1) * BN, lanes +16 * (s & 1) of every quadrant;
        // row i of the 64-row D: lane i % 16 of quadrant i / 16.  Each 32-lane x 32-column block is transposed through a
        // per-warp scratch (the drained activation ring) so one red.v4 covers 4 rows x 128 contiguous bytes.
        mbar_wait(&fullW[1], 0);
        tc_fence_after();
        asm volatile("bar.sync 2, 256;" ::: "memory");   // every epilogue warp is done with the rings before they become scratch
        float* scratch = reinterpret_cast<float*>(a_smem) + (warp - 2) * (32 * 36);
        const int taps = P.g.taps;
        const int nslots = taps + (P.wg_dbias != nullptr ? 1 : 0);
        const int passes = (nslots + 1) >> 1;
        const int K = P.g.N, N = P.g.K;   // weight gradient [tap][cin = this launch's output channels][cout = its contraction]
        for (int ps = half; ps < passes; ps += 2) {
          const int slot = 2 * ps + (lane >> 4);
          const int row = quad * 16 + (lane & 15);
          bool active = false;
          const float* row_dst = P.wg_dwp;
          if (slot < taps) {
            active = row < K;
            row_dst = P.wg_dwp + (static_cast<size_t>(slot) * K + (active ? row : 0)) * N;
          } else if (slot == taps && P.wg_dbias != nullptr) {
            active = row == 0;
            row_dst = P.wg_dbias;
          }
          const unsigned long long row_ptr = reinterpret_cast<unsigned long long>(row_dst);
          for (int c32 = 0; c32 < N / 32; ++c32) {
            float acc[32];
            tmem_ld32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + P.wg_col0 + static_cast<uint32_t>(ps * N + c32 * 32), acc);
#pragma unroll
            for (int n = 0; n < 32; n += 4)
              *reinterpret_cast<float4*>(scratch + lane * 36 + n) = make_float4(acc[n], acc[n + 1], acc[n + 2], acc[n + 3]);
            __syncwarp();
            float4 v[8];
            unsigned long long base[8];
            int act[8];
#pragma unroll
            for (int i8 = 0; i8 < 8; ++i8) {
              const int r = i8 * 4 + (lane >> 3);
              v[i8] = *reinterpret_cast<const float4*>(scratch + r * 36 + (lane & 7) * 4);
              base[i8] = __shfl_sync(0xffffffffu, row_ptr, r);
              act[i8] = __shfl_sync(0xffffffffu, active ? 1 : 0, r);
            }
#pragma unroll
            for (int i8 = 0; i8 < 8; ++i8) {
              float* dst = reinterpret_cast<float*>(base[i8]) + c32 * 32 + (lane & 7) * 4;
              if (act[i8])
                asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v[i8].x), "f"(v[i8].y),
                             "f"(v[i8].z), "f"(v[i8].w) : "memory");
            }
            __syncwarp();
          }
        }
      }
    }
  } else if (warp >= 2 && warp <= 9) {
    // ===================== epilogue warps =====================
    const int quad = warp & 3;               // TMEM lane quadrant this warp may access
    const int half = (warp - 2) >> 2;        // two warps per quadrant interleave over 16-column units
    const int row_in_tile = quad * 32 + lane;
    const Epilogue& e = P.e;
    static_assert(UW == 16 || UW == 32, "unit width");
    static_assert(UW == 16 || F == 0 || ((F & EPI_SMEM) && !(F & (EPI_RES2 | EPI_RAW))), "UW = 32 needs operand-free or smem-staged epilogues");
    constexpr int NG = UW / 8;                                   // 8-channel groups per unit
    const int ushift = P.units_shift - (UW == 32 ? 1 : 0);      // log2(BN / UW)
    const int units_per_mt = P.BN / UW;
    const int n_units = P.MT * units_per_mt;
    const uint32_t chunk_stride = static_cast<uint32_t>(padded_len(P.Lout)) * 8;  // next 8-channel group, same row (elements)
    const uint32_t b_stride = static_cast<uint32_t>(P.g.creal >> 3) * chunk_stride;   // next batch item
    const bool plain_out = e.tscale == 1.f && e.act_slope == 1.f;
    const float s_pos = e.scale, s_neg = e.mask_slope * e.scale;
    // per-tile state of this thread's row
    struct TileC { int b, r_phase, ch_tile, q_first; uint32_t t_lane; };
    auto tile_coords = [&](int tile, int it) {
      TileC t;
      int rest, nt, mg;
      tile_decode<CL>(P, tile, rest, nt);
      P.d_mgroups.divmod(rest, t.b, mg);
      const int buf = P.acc_bufs == 2 ? (it & 1) : 0;
      t.r_phase = P.d_creal.quot(nt * P.BN);            // scatter phase of this column tile (os > 1)
      t.ch_tile = nt * P.BN - t.r_phase * P.g.creal;    // first output channel of the tile
      t.q_first = mg * P.MT * 128 + row_in_tile;
      t.t_lane = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(buf * P.MT * P.BN);
      return t;
    };
    struct UnitC { int ro, ch; bool valid; uint32_t o; uint32_t taddr; };
    auto unit_coords = [&](const TileC& t, int u) {
      UnitC c;
      const int mt = u >> ushift, cu = u - (mt << ushift);
      const int q = t.q_first + mt * 128;
      c.ro = q * P.g.os + t.r_phase - P.g.p;
      c.valid = q < P.Lq && c.ro >= 0 && c.ro < P.Lout;
      c.ch = t.ch_tile + cu * UW;
      c.o = static_cast<uint32_t>(t.b) * b_stride + static_cast<uint32_t>(c.ch >> 3) * chunk_stride +
            static_cast<uint32_t>((c.valid ? c.ro : 0) + kPadL) * 8u;
      c.taddr = t.t_lane + static_cast<uint32_t>(mt * P.BN + cu * UW);
      return c;
    };

    // Software pipeline over this warp's units ACROSS tiles: the global operands (mask / residuals) of the next
    // unit -- the first unit of the next tile when the current tile is exhausted -- are requested before the
    // current unit is finished (and before the wait for its accumulator), so their latency is off the per-tile
    // critical path.  The two operand buffers alternate by code duplication (step(A, B); step(B, A)): a register
    // copy `cur = nxt` would wait for the just-issued loads and serialise the pipeline again.
    pdl_wait();   // masks / residuals / running sums may be the stream predecessor's output
    EpiLoads bufA[NG], bufB[NG];
    UnitC uc{};
    TileC tcur = tile_coords(blockIdx.x, 0);
    int tile = blockIdx.x, it = 0, u = half;
    const int grid = static_cast<int>(gridDim.x);
    if (tile < P.total_tiles) {
      uc = unit_coords(tcur, u);
#pragma unroll
      for (int h = 0; h < NG; ++h) epi_prefetch<F>(e, uc.o + h * chunk_stride, uc.valid, bufA[h]);
    }
    const bool has_bias = e.bias != nullptr || e.bias2 != nullptr;
    // one column tile and no per-batch bias: the bias vector is the same for every tile of the launch -- stage it once
    const bool bias_once = has_bias && P.n_tiles_n == 1 && e.bias2 == nullptr;
    if (bias_once) {
      const int et = static_cast<int>(threadIdx.x) - 64;
      if (et < P.BN) bias_s[et] = __ldg(e.bias + et);
      asm volatile("bar.sync 1, 256;" ::: "memory");
    }
    auto step = [&](EpiLoads (&use)[NG], EpiLoads (&fill)[NG]) {
      const int buf = P.acc_bufs == 2 ? (it & 1) : 0;
      const int use_n = P.acc_bufs == 2 ? (it >> 1) : it;
      const bool first = u == half;                      // first unit of this warp in the tile
      const bool last = u + 2 >= n_units;                // last unit of this warp in the tile
      if (first && has_bias && !bias_once) {  // stage this tile's bias vector (named barrier 1 = the 256 epilogue threads)
        const int et = static_cast<int>(threadIdx.x) - 64;
        if (et < P.BN) {
          float bv = e.bias ? __ldg(e.bias + tcur.ch_tile + et) : 0.f;
          if (e.bias2) bv += __ldg(e.bias2 + static_cast<size_t>(tcur.b) * P.g.creal + tcur.ch_tile + et);
          bias_s[(it & 1) * 128 + et] = bv;
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
      // next unit of this warp: same tile, or the first unit of the next tile
      const int n_tile = last ? tile + grid : tile, n_it = last ? it + 1 : it, n_u = last ? half : u + 2;
      TileC tnext = tcur;
      UnitC un{};
      if (n_tile < P.total_tiles) {
        if (last) tnext = tile_coords(n_tile, n_it);
        un = unit_coords(tnext, n_u);
#pragma unroll
        for (int h = 0; h < NG; ++h) epi_prefetch<F>(e, un.o + h * chunk_stride, un.valid, fill[h]);
      }
      int e_stage = 0;
      if constexpr (F & EPI_SMEM) {
        e_stage = it % P.NE;
        if (first) mbar_wait(&fullE[e_stage], static_cast<uint32_t>(it / P.NE) & 1u);
      }
      if (first) {
        mbar_wait(&acc_full[buf], use_n & 1);
        tc_fence_after();
        if (warp == 2 && lane == 0 && it < 4) ktrace(P.trace, 12 + it);   // accumulators of tile `it` complete
      }
      float acc[UW];
      if constexpr (UW == 16) tmem_ld16(uc.taddr, acc);
      else tmem_ld32(uc.taddr, acc);
      if (uc.valid) {
#pragma unroll
        for (int h = 0; h < NG; ++h) {
          float v[8];
#pragma unroll
          for (int n = 0; n < 8; ++n) v[n] = acc[h * 8 + n];
          if constexpr (F & EPI_SMEM) {   // this row's 16 bytes of each staged operand: [op][mt][channel group][128 rows][16 B]
            const int mt = u >> ushift;
            const int cg = (u - (mt << ushift)) * NG + h;
            const int groups = P.BN >> 3;
            const uint8_t* ep = e_smem + static_cast<size_t>(e_stage) * P.e_stage_bytes +
                                (static_cast<size_t>(mt * groups + cg) * 128 + row_in_tile) * 16;
            const size_t op_stride = static_cast<size_t>(P.MT) * groups * 2048;
            if constexpr (F & EPI_MASK) {
              use[h].mask = *reinterpret_cast<const uint4*>(ep);
              ep += op_stride;
            }
            if constexpr (F & EPI_RES) use[h].rest = *reinterpret_cast<const uint4*>(ep);
          }
          epi_finish<F>(e, P.g, tcur.b, uc.ro, uc.ch + h * 8, uc.o + h * chunk_stride, use[h],
                        has_bias ? bias_s + (bias_once ? 0 : (it & 1) * 128) + (uc.ch - tcur.ch_tile) + h * 8 : nullptr, plain_out,
                        s_pos, s_neg, v);
        }
      }
      if (last) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[buf]);
        if constexpr (F & EPI_SMEM) {
          if (lane == 0) mbar_arrive(&emptyE[e_stage]);
        }
        if (warp == 2 && lane == 0 && it < 4) ktrace(P.trace, 16 + it);   // epilogue of tile `it` done
      }
      tile = n_tile; it = n_it; u = n_u;
      tcur = tnext;
      uc = un;
    };
    while (tile < P.total_tiles) {
      step(bufA, bufB);
      if (tile >= P.total_tiles) break;
      step(bufB, bufA);
    }
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (CL) cluster_sync_all();   // the peer's last multicast arrivals on this CTA's barriers precede its own arrival here
  if (warp == 1) tmem_dealloc(tmem_base, P.tmem_cols);
  if (threadIdx.x == 0) ktrace(P.trace, 6);
}

// ---------------------------------------------------------------------------------------------------
// Weight gradient:  dWp[j][c][n] += sum_t in[b][t + off0 + j*step][c] * dout[b][t][n]     (is = os = 1)
// GEMM view: D[M = c][N = n] per tap with the contraction over time.  Both operands are read MN-major straight
// from the blocked layout (time rows are the 16-byte-strided K direction: LBO = 128 B, SBO = channel-group
// stride); a tap is again a +16*shift byte offset on the A descriptor -- the `in` tile is loaded ONCE per stage.
//   K_conv >= 128 : M = 128 (one 128-channel tile per CTA), accumulator of tap slot tl at TMEM columns tl*NT.
//   K_conv <= 64  : M = 64; two accumulators share NT columns (lanes +0 / +16 of every 32-lane quadrant).  For
//                   K_conv = 32 only rows 0..31 are meaningful (the upper half of the M = 64 operand reads
//                   whatever follows the tile in shared memory; rows of D are independent).
// One CTA = (channel tile, column tile, tap group, split of the flattened (batch, time block) range); a single
// split stores its result directly, several splits combine with fp32 atomics (dWp zeroed by the caller).  Few,
// long CTAs per layer are preferred: the host runs the weight-gradient kernels of independent layers
// concurrently on side streams.
// ---------------------------------------------------------------------------------------------------
struct WgradParams {
  const bf16* in;         // blocked, row-padded [B][K/8][Lin + pads][8]
  const bf16* dout;       // blocked, row-padded [B][N/8][L + pads][8]
  int Lin;
  float* dwp;             // [taps][K][N] fp32
  int taps, K, N, step, off0, minshift;
  int B, L;               // L = rows of dout per batch item (the contraction length)
  int M, mch, n_mtiles;   // instruction M (128 | 64), channel groups loaded per tile copy, channel tiles
  int G;                  // tap copies stacked along M (G*K <= M): copy g is the tile displaced by g*step rows
  int n_slots;            // accumulator slots in total = ceil(taps / G)
  int NT, n_ntiles;       // column tile (<= 256)
  int TG, n_tgroups;      // tap slots per CTA and number of tap groups
  int TK, RI;             // time rows per pipeline stage, rows of the A tile (TK + halo)
  int kb_per_item, n_splits, kb_per_split;  // the flattened (batch, time block) range is cut into n_splits
  int direct;             // 1: plain stores of the accumulators (n_splits == 1; debug: timing without reductions)
  int NS;
  uint32_t tmem_cols;
  float* dbias;           // bias gradient [cmod] (column sums of dout folded modulo cmod), or null
  int cmod;
  int* turn;              // deterministic mode (n_splits > 1): per output tile of this layer a ticket counter followed by one
                          // turn counter per accumulator pass (turn_stride ints per tile), else null
  int turn_stride;
  unsigned long long* trace;
};

// Deterministic split reduction ("turnstile"): the CTAs of one output tile take their split index from a ticket counter
// (arrival order, so the holder of ticket s - 1 is already resident when ticket s waits for it -- no dependence on the
// block dispatch order) and add their partial sums one after the other in ticket = time-range order.  The mainloops
// still run concurrently; the reduction passes (one per accumulator slot) form a wavefront: split s + 1 adds pass q
// while split s adds pass q + 1.
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__global__ void __launch_bounds__(kThreads, 1)
wgrad_kernel(const __grid_constant__ CUtensorMap tmIn, const __grid_constant__ CUtensorMap tmD, const WgradParams P) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  if (threadIdx.x == 0) ktrace(P.trace, 0);
  const uint32_t in_copy_bytes = static_cast<uint32_t>(P.mch) * P.RI * 16;
  const uint32_t in_bytes = in_copy_bytes * P.G;
  const uint32_t d_bytes = static_cast<uint32_t>(P.NT / 8) * P.TK * 16;
  const uint32_t stage_bytes = in_bytes + d_bytes;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + static_cast<size_t>(P.NS) * stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + P.NS;
  uint64_t* acc_full = empty + P.NS;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  // CTA coordinates
  int id = blockIdx.x;
  int split = id % P.n_splits; id /= P.n_splits;
  const int out_tile = id;               // (mtile, ntile, tap group): the CTAs that reduce into the same elements
  const int tg = id % P.n_tgroups; id /= P.n_tgroups;
  const int ntile = id % P.n_ntiles; id /= P.n_ntiles;
  const int mtile = id;
  // The bias gradient (column sums of dout) is folded into this kernel: the (mtile 0, tap group 0) CTA of every
  // (column tile, split) sums its dout stages with the four warps that otherwise idle until the epilogue.
  const bool do_colsum = P.dbias != nullptr && mtile == 0 && tg == 0;

  if (threadIdx.x == 0) {
    for (int i = 0; i < P.NS; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], do_colsum ? 5 : 1); }
    mbar_init(acc_full, 1);
    fence_barrier_init();
    if (P.turn) tmem_slot[1] = static_cast<uint32_t>(atomicAdd(P.turn + out_tile * P.turn_stride, 1));   // ticket = split index
  }
  if (warp == 1) tmem_alloc(tmem_slot, P.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  if (P.turn) split = static_cast<int>(tmem_slot[1]);
  if (threadIdx.x == 0) ktrace(P.trace, 1);
  const int f_begin = split * P.kb_per_split;
  const int f_end = min(P.B * P.kb_per_item, f_begin + P.kb_per_split);
  const int kblocks = f_end - f_begin;
  const int slot0 = tg * P.TG;
  const int nslots = min(P.TG, P.n_slots - slot0);

  if (warp == 0) {
    // TMA producer (converged warp, elected lane issues).  A stage is G + 1 tensor-TMA instructions: issuing one
    // 1-D bulk copy per (channel group) row run is limited to ~25 M copies/s per SM (measured), i.e. ~28 GB/s for
    // the 1-2 KB runs of a 64-row time block.
    Pipe ps;
    for (int kb = 0; kb < kblocks; ++kb) {
      const int f = f_begin + kb;
      const int b = f / P.kb_per_item;
      const int t0 = (f - b * P.kb_per_item) * P.TK;
      mbar_wait(&empty[ps.stage], ps.phase ^ 1);
      uint8_t* st = smem + static_cast<size_t>(ps.stage) * stage_bytes;
      if (elect_one()) {
        mbar_expect_tx(&full[ps.stage], stage_bytes);
        for (int g = 0; g < P.G; ++g)
          tma_load_3d(&tmIn, &full[ps.stage], st + g * in_copy_bytes, 2 * (t0 + P.off0 + P.minshift + g * P.step + kPadL),
                      mtile * P.mch, b);
        tma_load_3d(&tmD, &full[ps.stage], st + in_bytes, 2 * (t0 + kPadL), ntile * (P.NT / 8), b);
      }
      __syncwarp();
      ps.advance(P.NS);
    }
  } else if (warp == 1) {
    // MMA issuer: converged warp, elected lane issues
    const uint32_t idesc = make_idesc(P.M, P.NT, 1, 1);
    const uint64_t a_desc0 = make_desc(0, 128, static_cast<uint32_t>(P.RI) * 16);
    const uint64_t b_desc0 = make_desc(0, 128, static_cast<uint32_t>(P.TK) * 16);
    const int kks = P.TK / 16;
    Pipe ps;
    auto issue_stages = [&](auto kks_tag) {
      constexpr int KKS = decltype(kks_tag)::value;
      for (int kb = 0; kb < kblocks; ++kb) {
        mbar_wait(&full[ps.stage], ps.phase);
        tc_fence_after();
        if (lane == 0 && kb < 8) ktrace(P.trace, 8 + kb);        // stage kb landed
        const uint32_t a_base = smem_u32(smem + static_cast<size_t>(ps.stage) * stage_bytes);
        const uint32_t a_hi = static_cast<uint32_t>(a_desc0 >> 32), b_hi = static_cast<uint32_t>(b_desc0 >> 32);
        const uint32_t b_stage = static_cast<uint32_t>(b_desc0) + ((a_base + in_bytes) >> 4);
        uint32_t a_slot = static_cast<uint32_t>(a_desc0) + (a_base >> 4) + static_cast<uint32_t>(slot0 * P.G * P.step - P.minshift);
        for (int tl = 0; tl < nslots; ++tl) {
          uint32_t d_tmem;
          if (P.M == 128) d_tmem = tmem_base + static_cast<uint32_t>(tl * P.NT);
          else d_tmem = tmem_base + static_cast<uint32_t>((tl >> 1) * P.NT) + (static_cast<uint32_t>((tl & 1) * 16) << 16);
          // one elected region of KKS straight-line MMAs per tap slot (see the conv kernel / tools/mma_rate.cu)
          const uint32_t accum0 = kb != 0 ? 1u : 0u;
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < KKS; ++kk)
              umma_bf16_split(d_tmem, a_slot + 16u * kk, a_hi, b_stage + 16u * kk, b_hi, idesc, kk == 0 ? accum0 : 1u);
          }
          a_slot += static_cast<uint32_t>(P.G * P.step);
        }
        if (elect_one()) umma_commit(&empty[ps.stage]);
        ps.advance(P.NS);
      }
    };
    switch (kks) {
      case 7: issue_stages(std::integral_constant<int, 7>{}); break;
      case 6: issue_stages(std::integral_constant<int, 6>{}); break;
      case 5: issue_stages(std::integral_constant<int, 5>{}); break;
      case 4: issue_stages(std::integral_constant<int, 4>{}); break;
      default: __trap();
    }
    if (elect_one()) umma_commit(acc_full);
    if (lane == 0) ktrace(P.trace, 5);
  } else {
    const int quad = warp & 3;
    float bs[8];
    const int et = static_cast<int>(threadIdx.x) - 64;
    const int n_cg = P.NT / 8, RG = 128 / n_cg;
    const int cg = et / RG, rg = et - cg * RG;
    auto add_bias = [&]() {
      if (rg == 0) {
#pragma unroll
        for (int n = 0; n < 8; ++n) atomicAdd(P.dbias + (ntile * P.NT + cg * 8 + n) % P.cmod, bs[n]);
      }
    };
    if (do_colsum) {
      // thread -> (channel group cg, row group rg); consecutive threads read consecutive 16-byte rows (conflict-free)
#pragma unroll
      for (int n = 0; n < 8; ++n) bs[n] = 0.f;
      Pipe ps;
      for (int kb = 0; kb < kblocks; ++kb) {
        mbar_wait(&full[ps.stage], ps.phase);
        const uint8_t* dt = smem + static_cast<size_t>(ps.stage) * stage_bytes + in_bytes + static_cast<size_t>(cg) * P.TK * 16;
        for (int r = rg; r < P.TK; r += RG) {
          float v[8];
          unpack8(*reinterpret_cast<const uint4*>(dt + r * 16), v);
#pragma unroll
          for (int n = 0; n < 8; ++n) bs[n] += v[n];
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[ps.stage]);
        ps.advance(P.NS);
      }
      // reduce over the RG row groups (RG consecutive lanes, RG in {8, 16, 32}) and add to the bias gradient
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        for (int o = RG >> 1; o > 0; o >>= 1) bs[n] += __shfl_xor_sync(0xffffffffu, bs[n], o);
      }
      if (!P.turn) add_bias();
    }
    mbar_wait(acc_full, 0);
    tc_fence_after();
    if (warp == 2 && lane == 0) ktrace(P.trace, 12);
    // the scratch below aliases pipeline stage 0: every epilogue warp must be done with its column sums first
    asm volatile("bar.sync 2, 128;" ::: "memory");
    int* const turns = P.turn ? P.turn + out_tile * P.turn_stride : nullptr;
    const bool last_split = split == P.n_splits - 1;
    // Accumulators -> global.  A thread owns one accumulator row (TMEM lane); writing it out directly would scatter
    // every warp-level reduction over 32 rows (32 half-used sectors per instruction).  Each 32-row x 32-column block
    // is transposed through a per-warp scratch in the (drained) pipeline stages, so one red.v4 covers four rows x
    // 128 contiguous bytes: full sectors, 8x fewer memory wavefronts.
    float* scratch = reinterpret_cast<float*>(smem) + (warp - 2) * (32 * 36);   // [32 rows][36]: conflict-free v4 access
    const int passes = P.M == 128 ? nslots : (nslots + 1) / 2;
    for (int ps = 0; ps < passes; ++ps) {
      if (turns) {   // wait until the previous time range has added this pass
        if (threadIdx.x == 64) {
          uint32_t spins = 0;
          while (ld_acquire_gpu(turns + 1 + ps) != split) {
            __nanosleep(32);
            if (++spins > kSpinLimit) __trap();
          }
          if (ps == 0 && last_split) turns[0] = 0;   // every ticket of this launch has been taken: reset for the next one
        }
        asm volatile("bar.sync 2, 128;" ::: "memory");
        if (ps == 0 && do_colsum) add_bias();
      }
      // this lane's accumulator row: tap slot, tap j, input channel c
      int tl, row;
      uint32_t col0;
      if (P.M == 128) {
        tl = ps;
        row = quad * 32 + lane;
        col0 = static_cast<uint32_t>(tl * P.NT);
      } else {  // lanes 16..31 of a quadrant hold the odd slot of the column block; row i of D: lane i%16 of quadrant i/16
        tl = 2 * ps + (lane >> 4);
        row = quad * 16 + (lane & 15);
        col0 = static_cast<uint32_t>(ps * P.NT);
      }
      bool active = tl < nslots;
      int j, c;
      if (P.G == 1) {
        j = slot0 + tl;
        c = mtile * P.M + row;
        if (c >= P.K) active = false;
      } else {             // G copies of a K-channel tile stacked along M
        const int g = row / P.K;
        j = (slot0 + tl) * P.G + g;
        c = row - g * P.K;
        if (g >= P.G) active = false;
      }
      if (j >= P.taps) active = false;
      const float* row_dst = P.dwp + (static_cast<size_t>(active ? j : 0) * P.K + (active ? c : 0)) * P.N + ntile * P.NT;
      const unsigned long long row_ptr = reinterpret_cast<unsigned long long>(row_dst);
      for (int c32 = 0; c32 < P.NT / 32; ++c32) {
        float acc[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + col0 + c32 * 32, acc);
#pragma unroll
        for (int n = 0; n < 32; n += 4)
          *reinterpret_cast<float4*>(scratch + lane * 36 + n) = make_float4(acc[n], acc[n + 1], acc[n + 2], acc[n + 3]);
        __syncwarp();
        // gather first (shared-memory loads and shuffles of all eight row groups in flight together), then issue the
        // eight reductions back to back: the volatile reduction asm is an ordering point for the loads around it
        float4 v[8];
        unsigned long long base[8];
        int act[8];
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int r = it * 4 + (lane >> 3);
          v[it] = *reinterpret_cast<const float4*>(scratch + r * 36 + (lane & 7) * 4);
          base[it] = __shfl_sync(0xffffffffu, row_ptr, r);
          act[it] = __shfl_sync(0xffffffffu, active ? 1 : 0, r);
        }
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          float* dst = reinterpret_cast<float*>(base[it]) + c32 * 32 + (lane & 7) * 4;
          if (act[it]) {
            if (P.direct) {
              *reinterpret_cast<float4*>(dst) = v[it];
            } else {
              asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v[it].x), "f"(v[it].y),
                           "f"(v[it].z), "f"(v[it].w) : "memory");
            }
          }
        }
        __syncwarp();
      }
      if (turns) {   // pass this pass on (every thread's reductions are ordered before the release)
        __threadfence();
        asm volatile("bar.sync 2, 128;" ::: "memory");
        if (threadIdx.x == 64) st_release_gpu(turns + 1 + ps, last_split ? 0 : split + 1);
      }
    }
    if (warp == 2 && lane == 0) ktrace(P.trace, 21);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, P.tmem_cols);
  if (threadIdx.x == 0) ktrace(P.trace, 6);
}

}  // namespace tc
}  // namespace vcd

// Fused forward of one ResBlock1 pair for the <= 64-channel stages (vits/model/modules.py:203-216):
//     xt = c1(leaky_relu(x));  xt = c2(leaky_relu(xt));  x = xt + x
// in ONE launch.  These stages are bound by HBM / L2 traffic (5 tensor passes per pair as two launches: read x, write
// mid, read mid, read x, write out); here x is read once, `mid` goes through shared memory (and is written to global
// memory only when the backward pass needs it), and the output is written once: 3 passes in training, 2 in inference.
//
// Tile = R = 128 - (k - 1) output rows [t0, t0 + R).  c2 (dilation 1) needs mid on the 128 rows [t0 - h2, t0 - h2 + 128),
// h2 = (k - 1) / 2: exactly one 128-row MMA tile of c1, evaluated from the input rows [t0 - h2 - h1, ... + 128 + 2 h1),
// h1 = dil (k - 1) / 2 -- one bulk copy per channel group, as in conv_kernel.  Phase A: c1 MMAs -> TMEM acc1 ->
// epilogue-1 warps (bias, leaky ReLU, bf16; rows outside [0, L) forced to zero: c2 zero-pads the INTERMEDIATE) -> a
// shared-memory `mid` tile in A-operand layout (+ global store of the tile's own rows).  Phase B: c2 MMAs read `mid`
// through tap-shifted descriptors (rows R..127 of its accumulator are surplus and dropped) -> TMEM acc2 -> epilogue-2
// warps: bias, residual recovered from the activation tile already in shared memory, leaky ReLU, bf16 store.
// acc1 / acc2 / mid are double-buffered; the issuer interleaves c1(i + 1) before c2(i), so epilogue-1 of a tile overlaps
// the tensor-pipe work of its neighbours.  Both weight sets stay resident in shared memory for the CTA's lifetime.
// A CTA tile is MT (1 | 2 | 4) consecutive 128-row MMA tiles = MT * 128 - (k - 1) output rows: the four barrier
// hand-overs per tile (MMA -> epilogue-1 -> MMA -> epilogue-2, ~0.4 us each with two buffers in flight) are what
// bounds a 128-row tile of these layers, not bandwidth; MT amortises them (and the halo) over more rows.
//
// BWD instantiation = the data gradients of the same pair in one launch:
//     dm = mask(mid) * dgrad_c2(G);   G' = G + mask(x_in) * dgrad_c1(dm)        (mask(a) = a > 0 ? 1 : slope)
// Same structure with the roles of the dilations swapped (phase A = c2's data gradient, dilation 1, small input halo;
// phase B = c1's, dilated, reading `dm` from shared memory: R = MT * 128 - 2 h1 output rows), masks instead of bias +
// activation (loaded from global memory ahead of the accumulator wait), the identity path of `x = xt + x` taken from the
// gradient tile already in shared memory, `dm` always stored (c1's weight gradient reads it).
//
// Warp roles (576 threads): 0 = producer (weights once, activation tiles), 1 = TMEM allocator + MMA issuer,
// 2..9 = epilogue-1, 10..17 = epilogue-2 (two warps per TMEM lane quadrant each, splitting the columns: with one warp
// per quadrant the two epilogues, not the memory system, set the time per tile).
#pragma once
#include "tc_kernels.cuh"

namespace vcd {
namespace tc {

constexpr int kPairThreads = 576;
constexpr int kMidSlack = 16;    // tap over-read of c2 past the last mid row (k - 1 <= 10 rows, only ever feeds dropped accumulator rows)

struct PairParams {
  const bf16* in;        // blocked, row-padded activated input lrelu(x) [B][C/8][L + pads][8]
  const bf16* w1;        // packed [tap][C/8][C][8] (forward format of c1, one column tile)
  const bf16* w2;
  const float* bias1;
  const float* bias2;
  bf16* mid_out;         // lrelu(c1(.)) [B][C/8][L + pads][8], or null (inference)
  bf16* out;             // lrelu((x + c2(.) [+ res2]) * tscale, out_slope) [B][C/8][L + pads][8], or null
  // final pair of a ResBlock branch: the running fp32 sum over the branches (same blocked layout, 4-byte elements)
  const float* res2;     // sum of the previous branches, or null
  float* out_raw;        // x + c2(.) [+ res2] in fp32 (input of the next branch's final pair), or null
  float tscale, out_slope;
  int B, L, C, taps;
  int dilA, dilB;        // dilation of phase A / phase B (forward: c1's dilation, 1; backward: 1, c1's dilation)
  int hA, hB;            // halos: dilA * (taps - 1) / 2, dilB * (taps - 1) / 2
  int mid_slack;         // rows phase B's taps read past the last `mid` row (>= 2 * hB, multiple of 8)
  const bf16* mask1;     // BWD: stored lrelu(c1 out) -> mask of dm;   mask2: stored lrelu(x_in) -> mask of c1's data gradient
  const bf16* mask2;
  float mask_slope;
  int MT;                // 128-row MMA tiles per CTA tile
  int R;                 // output rows per CTA tile = MT * 128 - 2 * hB
  int RA;                // rows of an activation region = MT * 128 + 2 * hA, rounded up to 8
  int NA;                // activation ring depth
  int tiles_per_item, total_tiles;
  FastDiv d_tiles;       // divider by tiles_per_item
  float act_slope, res_inv;
  uint32_t tmem_cols;
};

template <bool SAVE_MID, bool BWD = false>
__global__ void __launch_bounds__(kPairThreads, 1)
pair_kernel(const PairParams P) {
  static_assert(!BWD || SAVE_MID, "the backward pair always stores dm");
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int C = P.C, cgs = C >> 3, taps = P.taps;
  const uint32_t a_stage_bytes = static_cast<uint32_t>(cgs) * P.RA * 16;
  const uint32_t w_bytes = static_cast<uint32_t>(taps) * cgs * C * 16;
  const int mid_rows = P.MT * 128 + P.mid_slack;
  const uint32_t mid_bytes = static_cast<uint32_t>(cgs) * mid_rows * 16;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  uint8_t* a_smem = smem;
  uint8_t* w1_smem = a_smem + static_cast<size_t>(P.NA) * a_stage_bytes;
  uint8_t* w2_smem = w1_smem + w_bytes;
  uint8_t* mid_smem = w2_smem + w_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(mid_smem + 2 * mid_bytes);
  uint64_t* fullA = bars;                 // [8]
  uint64_t* emptyA = fullA + 8;           // [8]   c1's MMAs (commit) + the eight epilogue-2 warps (residual reads)
  uint64_t* w_full = emptyA + 8;          // [1]
  uint64_t* acc1_full = w_full + 1;       // [2]
  uint64_t* acc1_empty = acc1_full + 2;   // [2]   eight epilogue-1 warps
  uint64_t* mid_full = acc1_empty + 2;    // [2]   eight epilogue-1 warps
  uint64_t* mid_empty = mid_full + 2;     // [2]   c2's MMAs (commit)
  uint64_t* acc2_full = mid_empty + 2;    // [2]
  uint64_t* acc2_empty = acc2_full + 2;   // [2]   eight epilogue-2 warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc2_empty + 2);
  float* bias_s = reinterpret_cast<float*>(tmem_slot + 4);   // [2][64]

  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) { mbar_init(&fullA[i], 1); mbar_init(&emptyA[i], 9); }
    mbar_init(w_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc1_full[i], 1); mbar_init(&acc1_empty[i], 8);
      mbar_init(&mid_full[i], 8); mbar_init(&mid_empty[i], 1);
      mbar_init(&acc2_full[i], 1); mbar_init(&acc2_empty[i], 8);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, P.tmem_cols);
  // biases (older than the stream predecessor's output: no dependency wait) and the never-written slack rows of `mid`
  if (!BWD && threadIdx.x >= 64 && threadIdx.x < 64 + 2 * 64) {
    const int i = threadIdx.x - 64, which = i >> 6, ch = i & 63;
    if (ch < C) bias_s[which * 64 + ch] = __ldg((which ? P.bias2 : P.bias1) + ch);
  }
  for (uint32_t o = threadIdx.x * 16u; o < 2 * mid_bytes; o += kPairThreads * 16u)
    *reinterpret_cast<uint4*>(mid_smem + o) = make_uint4(0u, 0u, 0u, 0u);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  const int grid = static_cast<int>(gridDim.x);

  if (warp == 0) {
    // ===================== producer =====================
    if (elect_one()) {
      mbar_expect_tx(w_full, 2 * w_bytes);
      for (uint32_t off = 0; off < w_bytes; off += 65536u) {
        bulk_load(w1_smem + off, reinterpret_cast<const uint8_t*>(P.w1) + off, min(65536u, w_bytes - off), w_full);
        bulk_load(w2_smem + off, reinterpret_cast<const uint8_t*>(P.w2) + off, min(65536u, w_bytes - off), w_full);
      }
    }
    __syncwarp();
    const uint32_t cg_bytes = static_cast<uint32_t>(P.RA) * 16;
    const size_t cg_stride_g = static_cast<size_t>(padded_len(P.L)) * 8;
    Pipe pa;
    pdl_wait();   // the activations are the stream predecessor's output
    for (int tile = blockIdx.x; tile < P.total_tiles; tile += grid) {
      if (tile + grid >= P.total_tiles) pdl_launch();
      int b, ti;
      P.d_tiles.divmod(tile, b, ti);
      const int row0 = ti * P.R - P.hB - P.hA;
      mbar_wait(&emptyA[pa.stage], pa.phase ^ 1);
      uint8_t* stage = a_smem + static_cast<size_t>(pa.stage) * a_stage_bytes;
      const bf16* src0 = P.in + blk_row(b, 0, row0, C, P.L);
      // the last CTA tile of an item may reach past the zero pad that follows row L - 1: copy only the rows that exist
      // (the rest of the region keeps stale, finite data; it only feeds mid rows >= L, which are forced to zero)
      const uint32_t copy_bytes = static_cast<uint32_t>(min(P.RA, P.L + kPadR - row0)) * 16;
      if (elect_one()) {
        mbar_expect_tx(&fullA[pa.stage], cgs * copy_bytes);
        for (int cg = 0; cg < cgs; ++cg) bulk_load(stage + cg * cg_bytes, src0 + cg * cg_stride_g, copy_bytes, &fullA[pa.stage]);
      }
      __syncwarp();
      pa.advance(P.NA);
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = make_idesc(128, C, 0, 0);
    const uint32_t a_lbo = static_cast<uint32_t>(P.RA) * 16, m_lbo = static_cast<uint32_t>(mid_rows) * 16, w_lbo = static_cast<uint32_t>(C) * 16;
    const uint32_t a_kk16 = (2 * a_lbo) >> 4, m_kk16 = (2 * m_lbo) >> 4, w_kk16 = (2 * w_lbo) >> 4;
    const uint32_t w_tap16 = (static_cast<uint32_t>(cgs) * C * 16) >> 4;
    const uint64_t a_desc0 = make_desc(0, a_lbo, 128), m_desc0 = make_desc(0, m_lbo, 128), w_desc0 = make_desc(0, w_lbo, 128);
    const uint32_t a_hi = static_cast<uint32_t>(a_desc0 >> 32), m_hi = static_cast<uint32_t>(m_desc0 >> 32), w_hi = static_cast<uint32_t>(w_desc0 >> 32);
    const uint32_t w1_lo = static_cast<uint32_t>(w_desc0) + (smem_u32(w1_smem) >> 4);
    const uint32_t w2_lo = static_cast<uint32_t>(w_desc0) + (smem_u32(w2_smem) >> 4);
    const uint32_t dilA = static_cast<uint32_t>(P.dilA), dilB = static_cast<uint32_t>(P.dilB);
    mbar_wait(w_full, 0);
    tc_fence_after();
    int n_tiles = 0;
    for (int tile = blockIdx.x; tile < P.total_tiles; tile += grid) ++n_tiles;
    Pipe pa;
    // per tap one elected block of MT * KK straight-line MMAs (KK = C / 16 and MT compile-time)
    auto issue = [&](auto kk_tag, auto mt_tag) {
      constexpr int KK = decltype(kk_tag)::value, MT = decltype(mt_tag)::value;
      const uint32_t cmt = static_cast<uint32_t>(C);
      for (int i = 0; i <= n_tiles; ++i) {
        if (i < n_tiles) {   // ---- c1 of tile i ----
          const int buf = i & 1, use = i >> 1;
          mbar_wait(&acc1_empty[buf], (use & 1) ^ 1);
          mbar_wait(&fullA[pa.stage], pa.phase);
          tc_fence_after();
          const uint32_t d1 = tmem_base + static_cast<uint32_t>(buf * MT * C);
          uint32_t a_tap = static_cast<uint32_t>(a_desc0) + (smem_u32(a_smem + static_cast<size_t>(pa.stage) * a_stage_bytes) >> 4);
          uint32_t w_lo = w1_lo;
#pragma unroll 1
          for (int j = 0; j < taps; ++j) {
            if (elect_one()) {
#pragma unroll
              for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
                for (int kk = 0; kk < KK; ++kk)
                  umma_bf16_split(d1 + static_cast<uint32_t>(mt) * cmt, a_tap + static_cast<uint32_t>(mt) * 128u + static_cast<uint32_t>(kk) * a_kk16, a_hi,
                                  w_lo + static_cast<uint32_t>(kk) * w_kk16, w_hi, idesc, (j | kk) != 0 ? 1u : 0u);
              }
            }
            a_tap += dilA;
            w_lo += w_tap16;
          }
          if (elect_one()) {
            umma_commit(&acc1_full[buf]);
            umma_commit(&emptyA[pa.stage]);
          }
          __syncwarp();
          pa.advance(P.NA);
        }
        if (i >= 1) {        // ---- c2 of tile i - 1 ----
          const int t = i - 1, buf = t & 1, use = t >> 1;
          mbar_wait(&acc2_empty[buf], (use & 1) ^ 1);
          mbar_wait(&mid_full[buf], use & 1);
          tc_fence_after();
          const uint32_t d2 = tmem_base + static_cast<uint32_t>((2 + buf) * MT * C);
          uint32_t m_tap = static_cast<uint32_t>(m_desc0) + (smem_u32(mid_smem + static_cast<size_t>(buf) * mid_bytes) >> 4);
          uint32_t w_lo = w2_lo;
#pragma unroll 1
          for (int j = 0; j < taps; ++j) {
            if (elect_one()) {
#pragma unroll
              for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
                for (int kk = 0; kk < KK; ++kk)
                  umma_bf16_split(d2 + static_cast<uint32_t>(mt) * cmt, m_tap + static_cast<uint32_t>(mt) * 128u + static_cast<uint32_t>(kk) * m_kk16, m_hi,
                                  w_lo + static_cast<uint32_t>(kk) * w_kk16, w_hi, idesc, (j | kk) != 0 ? 1u : 0u);
              }
            }
            m_tap += dilB;
            w_lo += w_tap16;
          }
          if (elect_one()) {
            umma_commit(&acc2_full[buf]);
            umma_commit(&mid_empty[buf]);
          }
          __syncwarp();
        }
      }
    };
    using std::integral_constant;
    auto by_mt = [&](auto kk_tag) {
      if (P.MT == 1) issue(kk_tag, integral_constant<int, 1>{});
      else if (P.MT == 2) issue(kk_tag, integral_constant<int, 2>{});
      else issue(kk_tag, integral_constant<int, 4>{});
    };
    if (C == 64) by_mt(integral_constant<int, 4>{});
    else by_mt(integral_constant<int, 2>{});
  } else if (warp <= 9) {
    // ===================== epilogue 1: acc1 -> mid (shared memory A operand of c2, + global) =====================
    const int quad = warp & 3, half = (warp - 2) >> 2, r = quad * 32 + lane;      // accumulator row = TMEM lane
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const size_t chunk_stride = static_cast<size_t>(padded_len(P.L)) * 8;
    auto run = [&](auto ncol_tag, auto mt_tag) {
      constexpr int NCOL = decltype(ncol_tag)::value;       // columns per warp = C / 2
      constexpr int MT = decltype(mt_tag)::value;
      const int col0 = half * NCOL;
      int i = 0;
      for (int tile = blockIdx.x; tile < P.total_tiles; tile += grid, ++i) {
        const int buf = i & 1, use = i >> 1;
        int b, ti;
        P.d_tiles.divmod(tile, b, ti);
        // BWD: the leaky-ReLU masks of this thread's rows (stored lrelu(c1 out)), requested ahead of the accumulator wait
        uint4 mk[BWD ? MT : 1][BWD ? NCOL / 8 : 1];
        if constexpr (BWD) {
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            const int t = ti * P.R - P.hB + mt * 128 + r;
            const bool inside = t >= 0 && t < P.L;
            const bf16* mrow_g = P.mask1 + blk_row(b, col0 >> 3, inside ? t : 0, C, P.L);
#pragma unroll
            for (int h = 0; h < NCOL / 8; ++h) mk[mt][h] = __ldg(reinterpret_cast<const uint4*>(mrow_g + h * chunk_stride));
          }
        }
        mbar_wait(&mid_empty[buf], (use & 1) ^ 1);
        mbar_wait(&acc1_full[buf], use & 1);
        tc_fence_after();
        // all TMEM loads of the CTA tile first, ONE wait: the load latency is paid once per tile, not once per row tile
        uint32_t raw[MT][NCOL];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
          for (int c16 = 0; c16 < NCOL / 16; ++c16)
            tmem_ld16_issue(t_lane + static_cast<uint32_t>((buf * MT + mt) * C + col0 + c16 * 16), *reinterpret_cast<uint32_t(*)[16]>(&raw[mt][c16 * 16]));
        }
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
          for (int c16 = 0; c16 < NCOL / 16; ++c16) tmem_wait16(*reinterpret_cast<uint32_t(*)[16]>(&raw[mt][c16 * 16]));
        }
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          const int rr = mt * 128 + r;                        // row of the CTA tile's mid region
          const int t = ti * P.R - P.hB + rr;                 // its global row
          const bool inside = t >= 0 && t < P.L;
          const bool own = SAVE_MID && inside && rr >= P.hB && rr < P.hB + P.R;   // the tile stores its own R rows
          uint8_t* mrow = mid_smem + static_cast<size_t>(buf) * mid_bytes + static_cast<size_t>(rr) * 16;
          bf16* grow = SAVE_MID ? P.mid_out + blk_row(b, 0, inside ? t : 0, C, P.L) : nullptr;
#pragma unroll
          for (int h = 0; h < NCOL / 8; ++h) {
            const int cg = (col0 >> 3) + h;
            float v[8];
            if constexpr (BWD) {
              float m[8];
              unpack8(mk[mt][h], m);
#pragma unroll
              for (int n = 0; n < 8; ++n)     // the intermediate gradient outside [0, L) is zero (c2 zero-pads its input)
                v[n] = inside ? __uint_as_float(raw[mt][h * 8 + n]) * (m[n] > 0.f ? 1.f : P.mask_slope) : 0.f;
            } else {
#pragma unroll
              for (int n = 0; n < 8; ++n) {
                const float x = __uint_as_float(raw[mt][h * 8 + n]) + bias_s[cg * 8 + n];
                v[n] = inside ? fmaxf(x, x * P.act_slope) : 0.f;   // c2 zero-pads the intermediate outside [0, L)
              }
            }
            uint4 pk;
            pk.x = pack_bf16x2(v[0], v[1]); pk.y = pack_bf16x2(v[2], v[3]);
            pk.z = pack_bf16x2(v[4], v[5]); pk.w = pack_bf16x2(v[6], v[7]);
            *reinterpret_cast<uint4*>(mrow + static_cast<size_t>(cg) * mid_rows * 16) = pk;
            if (own) *reinterpret_cast<uint4*>(grow + cg * chunk_stride) = pk;
          }
        }
        fence_proxy_async();       // the generic-proxy writes of `mid` must be visible to the tensor core's (async proxy) reads
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&mid_full[buf]);
          mbar_arrive(&acc1_empty[buf]);
        }
      }
    };
    using std::integral_constant;
    if (C == 64) {   // (host: 4 * MT * C <= 512 TMEM columns, so MT <= 2 at 64 channels)
      if (P.MT == 1) run(integral_constant<int, 32>{}, integral_constant<int, 1>{});
      else run(integral_constant<int, 32>{}, integral_constant<int, 2>{});
    } else {
      if (P.MT == 1) run(integral_constant<int, 16>{}, integral_constant<int, 1>{});
      else if (P.MT == 2) run(integral_constant<int, 16>{}, integral_constant<int, 2>{});
      else run(integral_constant<int, 16>{}, integral_constant<int, 4>{});
    }
  } else {
    // ===================== epilogue 2: acc2 + bias + residual -> out =====================
    const int quad = warp & 3, half = (warp - 10) >> 2, r = quad * 32 + lane;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const size_t chunk_stride = static_cast<size_t>(padded_len(P.L)) * 8;
    const uint32_t cg_bytes = static_cast<uint32_t>(P.RA) * 16;
    auto run = [&](auto ncol_tag, auto mt_tag) {
      constexpr int NCOL = decltype(ncol_tag)::value;
      constexpr int MT = decltype(mt_tag)::value;
      const int col0 = half * NCOL;
      Pipe pa;
      int i = 0;
      for (int tile = blockIdx.x; tile < P.total_tiles; tile += grid, ++i) {
        const int buf = i & 1, use = i >> 1;
        int b, ti;
        P.d_tiles.divmod(tile, b, ti);
        uint4 mk[BWD ? MT : 1][BWD ? NCOL / 8 : 1];   // BWD: masks of this thread's output rows (stored lrelu(x_in))
        if constexpr (BWD) {
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            const int rr = mt * 128 + r, t = ti * P.R + rr;
            const bool valid = rr < P.R && t < P.L;
            const bf16* mrow_g = P.mask2 + blk_row(b, col0 >> 3, valid ? t : 0, C, P.L);
#pragma unroll
            for (int h = 0; h < NCOL / 8; ++h) mk[mt][h] = __ldg(reinterpret_cast<const uint4*>(mrow_g + h * chunk_stride));
          }
        }
        mbar_wait(&fullA[pa.stage], pa.phase);              // (long complete; orders this warp's reads after the bulk copies)
        mbar_wait(&acc2_full[buf], use & 1);
        tc_fence_after();
        uint32_t raw[MT][NCOL];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
          for (int c16 = 0; c16 < NCOL / 16; ++c16)
            tmem_ld16_issue(t_lane + static_cast<uint32_t>(((2 + buf) * MT + mt) * C + col0 + c16 * 16), *reinterpret_cast<uint32_t(*)[16]>(&raw[mt][c16 * 16]));
        }
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
          for (int c16 = 0; c16 < NCOL / 16; ++c16) tmem_wait16(*reinterpret_cast<uint32_t(*)[16]>(&raw[mt][c16 * 16]));
        }
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          const int rr = mt * 128 + r;
          const int t = ti * P.R + rr;                        // global output row
          const bool valid = rr < P.R && t < P.L;
          // residual: x recovered from the stored lrelu(x) (BWD: the incoming gradient itself); its row in the activation
          // region is rr + hB + hA
          const uint8_t* xrow = a_smem + static_cast<size_t>(pa.stage) * a_stage_bytes + static_cast<size_t>(min(rr + P.hB + P.hA, P.RA - 1)) * 16;
          const size_t o0 = blk_row(b, 0, valid ? t : 0, C, P.L);
          if (valid) {
#pragma unroll
            for (int h = 0; h < NCOL / 8; ++h) {
              const int cg = (col0 >> 3) + h;
              float xr[8], v[8];
              unpack8(*reinterpret_cast<const uint4*>(xrow + static_cast<size_t>(cg) * cg_bytes), xr);
              const size_t o = o0 + cg * chunk_stride;
              if constexpr (BWD) {   // G' = G + mask * dgrad_c1(dm), stored as is
                float m[8];
                unpack8(mk[mt][h], m);
#pragma unroll
                for (int n = 0; n < 8; ++n) v[n] = xr[n] + __uint_as_float(raw[mt][h * 8 + n]) * (m[n] > 0.f ? 1.f : P.mask_slope);
                store8<bf16>(P.out + o, v);
                continue;
              }
#pragma unroll
              for (int n = 0; n < 8; ++n) v[n] = __uint_as_float(raw[mt][h * 8 + n]) + bias_s[64 + cg * 8 + n] + (xr[n] > 0.f ? xr[n] : xr[n] * P.res_inv);
              if (P.res2 != nullptr) {
                const float4 r0 = __ldg(reinterpret_cast<const float4*>(P.res2 + o)), r1 = __ldg(reinterpret_cast<const float4*>(P.res2 + o) + 1);
                v[0] += r0.x; v[1] += r0.y; v[2] += r0.z; v[3] += r0.w;
                v[4] += r1.x; v[5] += r1.y; v[6] += r1.z; v[7] += r1.w;
              }
              if (P.out_raw != nullptr) store8<float>(P.out_raw + o, v);
              if (P.out != nullptr) {
#pragma unroll
                for (int n = 0; n < 8; ++n) {
                  const float x = v[n] * P.tscale;
                  v[n] = fmaxf(x, x * P.out_slope);
                }
                store8<bf16>(P.out + o, v);
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&acc2_empty[buf]);
          mbar_arrive(&emptyA[pa.stage]);
        }
        pa.advance(P.NA);
      }
    };
    using std::integral_constant;
    if (C == 64) {   // (host: 4 * MT * C <= 512 TMEM columns, so MT <= 2 at 64 channels)
      if (P.MT == 1) run(integral_constant<int, 32>{}, integral_constant<int, 1>{});
      else run(integral_constant<int, 32>{}, integral_constant<int, 2>{});
    } else {
      if (P.MT == 1) run(integral_constant<int, 16>{}, integral_constant<int, 1>{});
      else if (P.MT == 2) run(integral_constant<int, 16>{}, integral_constant<int, 2>{});
      else run(integral_constant<int, 16>{}, integral_constant<int, 4>{});
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, P.tmem_cols);
}

}  // namespace tc
}  // namespace vcd

// Host-side plan: layer table, parameter table, packed-weight arenas, workspace layout.
#pragma once
#include <cuda_runtime.h>

#include <map>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/vcd.h"
#include "common.cuh"
#include "fold.cuh"
#include "fold_fast.cuh"

namespace vcd {

struct ParamInfo {
  std::string name;
  int64_t shape[3];
  int ndim;
  int64_t numel;
};

enum : int { LK_CONV = 0, LK_CONVT = 1 };

struct Layer {
  std::string name;
  int kind;                   // LK_*
  int cin, cout, k, dil, u, pad;
  bool wn;
  int p_w, p_g, p_b;          // parameter indices (p_w: weight or weight_v), -1 if absent
  ConvGeo fwd, dgr, wgr;      // generalised geometries (common.cuh); wgr = weight-gradient view of fwd
  WeightMap map_fwd, map_dgr;
  long long f32_fwd, f32_dgr; // offsets (elements) into the fp32 packed arena
  long long tc_fwd, tc_dgr;   // offsets (elements) into the bf16 packed arena, -1 if no tensor-core path
  int nt_fwd, nt_dgr;         // tensor-core column tile of the packed bf16 format
  long long dwp;              // offset (floats) of dWp[taps][K][N] (fwd geometry) in the gradient scratch
  long long dbias;            // offset (floats) of the bias gradient in the gradient scratch, -1 if none
  int norm_off;               // offset into the norms arena (weight-normed only)
  int segment;                // backward segment that finalises this layer's gradients
  bool tc_ok_fwd, tc_ok_dgr, tc_ok_wgr;  // eligible for the tcgen05 kernels
};

struct SegmentJobs {
  UnfoldJob* d_jobs = nullptr;
  int njobs = 0, nblocks = 0;
  // bandwidth-shaped variant (fold_fast.cuh), used when every layer of the segment qualifies
  FastUnfoldJob* d_fast = nullptr;
  int fast_njobs = 0, fast_nblocks = 0, fast_lead_jobs = 0, fast_lead_blocks = 0;
  size_t fast_smem = 0;
  int lead_jobs = 0, lead_blocks = 0;            // jobs / blocks ahead of the first layer's (conv_post.weight in segment 0)
  long long scratch_begin = 0, scratch_end = 0;  // float range of the gradient scratch to zero
  std::vector<int> params;
};

// Identity of a captured launch sequence (see run_graphed in vcd_api.cu).
struct GraphKey {
  int region;          // 0 = forward core, 1 + s = backward segment s
  int mode, B, T, save, has_g, has_dx;
  const void* ws;
  uint64_t params_version;
  bool operator<(const GraphKey& o) const {
    return std::tie(region, mode, B, T, save, has_g, has_dx, ws, params_version) <
           std::tie(o.region, o.mode, o.B, o.T, o.save, o.has_g, o.has_dx, o.ws, o.params_version);
  }
};

struct StageDesc {
  int cin, cout, u, k;        // upsample conv
  int up_layer;               // index into layers
  // resblock branches: layers[branch][pair][0|1] (ResBlock2: [pair][0] only)
  std::vector<std::vector<std::vector<int>>> convs;
};

}  // namespace vcd

struct vcd_plan {
  vcd_config cfg;
  int device = 0, num_sms = 0;
  int hop = 1;
  std::vector<vcd::ParamInfo> params;
  std::vector<vcd::Layer> layers;
  std::vector<vcd::StageDesc> stages;
  int l_pre = -1;
  int p_post_w = -1, p_cond_w = -1, p_cond_b = -1;
  long long post_dw = -1;     // gradient scratch offset of conv_post.weight (parameter layout)
  long long cond_dw = -1, cond_db = -1;

  // device arenas (library-owned)
  float* d_f32 = nullptr;     // packed fp32 weights
  vcd::bf16* d_bf16 = nullptr;  // packed bf16 weights
  float* d_norms = nullptr;
  float* d_gscratch = nullptr;  // packed weight gradients + bias gradients
  long long n_f32 = 0, n_bf16 = 0, n_norms = 0, n_gscratch = 0;
  const float** d_params = nullptr;  // device copy of the parameter pointer table
  float** d_dparams = nullptr;
  std::vector<const float*> h_params;
  std::vector<float*> h_dparams;
  std::vector<char> is_weight_g;     // parameter i is a weight_g (may be NULL after remove_weight_norm)
  bool folded[2] = {false, false};

  vcd::NormJob* d_norm_jobs = nullptr;
  int n_norm_jobs = 0, n_norm_blocks = 0;
  vcd::PackJob* d_pack_jobs[2] = {nullptr, nullptr};  // per mode
  int n_pack_jobs[2] = {0, 0}, n_pack_blocks[2] = {0, 0};
  // bf16 mode: layers whose two tensor-core operand formats are written by wn_pack_fast_kernel (norms included);
  // the generic norm / pack tables of that mode then only hold the remaining layers
  vcd::FastPackJob* d_fast_pack = nullptr;
  int n_fast_pack_jobs = 0, n_fast_pack_blocks = 0;
  size_t fast_pack_smem = 0;
  vcd::NormJob* d_norm_jobs_bf16 = nullptr;
  int n_norm_jobs_bf16 = 0, n_norm_blocks_bf16 = 0;
  float grad_scale = 1.f;                  // multiplies every parameter gradient (vcd_set_gradient_scale)
  bool deterministic = false;              // vcd_set_deterministic: fixed-order reductions of every parameter gradient
  static constexpr int kTurnInts = 2048;   // ticket / turn counters per layer for the weight-gradient turnstile
  int* d_turn = nullptr;                   // [layers][kTurnInts], zero between launches
  static constexpr int kDetParts = 64, kDetFloats = kDetParts * 512;
  float* d_det = nullptr;                  // [layers + 1][kDetFloats] partial column sums of the deterministic mode
  std::vector<vcd::SegmentJobs> segments;

  // auxiliary streams / events for intra-step concurrency (ResBlock branches, weight-gradient kernels)
  static constexpr int kMaxAux = 14, kMaxEvents = 1024;   // 7 branch + 4 weight-gradient side streams + 3 helpers
  cudaStream_t aux[kMaxAux] = {};
  cudaStream_t own = nullptr;              // stands in for the legacy default stream (not capturable)
  cudaEvent_t hop_in = nullptr, hop_out = nullptr;
  // whole-backward mode: "gradients of segment i are final" (recorded inside the captured graph as external events) and
  // "conv_post's weight gradient is final" (segment 0, outside the graph); valid after a whole-mode vcd_backward
  cudaEvent_t seg_done[VCD_MAX_UPSAMPLES + 2] = {};
  cudaEvent_t post_done = nullptr;
  bool seg_events_valid = false;
  cudaEvent_t events[kMaxEvents] = {};
  int next_event = 0;

  unsigned long long* d_trace = nullptr;   // VCD_KTRACE debug buffer

  // CUDA graph cache
  uint64_t params_version = 0;
  std::map<vcd::GraphKey, cudaGraphExec_t> graphs;
  std::map<vcd::GraphKey, uint64_t> graph_kernels;

  // device copies of the pad-zeroing job tables, keyed by (mode, B, T, save, phase, backward)
  std::map<std::tuple<int, int, int, int, int, int>, void*> pad_tables;
};

// C ABI of the B200-native VCVITS HiFi-GAN decoder (include/vcd.h).  Host-side schedule + launches.
//
// Reference interfaces replaced (all in /root/reference): Generator.__init__ / forward behind
// SynthesizerTTS.dec (vits/model/synthesizers/synthesizer_tts.py:71-78,140), ResBlock1/ResBlock2
// (vits/model/modules.py:186-247), weight_norm pre-hooks (modules.py:10), autograd backward of the train.py
// step (vits/light/vcvits.py:54-148).
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <memory>
#include <vector>

#include "plan.h"
#include "simt_kernels.cuh"
#include "tc_conv.cuh"

using namespace vcd;

// ---------------------------------------------------------------------------------------------------
// error handling / counters
// ---------------------------------------------------------------------------------------------------
static thread_local char g_err[1024] = "";
static std::atomic<uint64_t> g_launches{0};

static int fail(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return 1;
}

#define CU_TRY(expr)                                                                           \
  do {                                                                                         \
    cudaError_t e__ = (expr);                                                                  \
    if (e__ != cudaSuccess) return fail("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)

#define LAUNCH_CHECK(what)                                                                     \
  do {                                                                                         \
    g_launches.fetch_add(1, std::memory_order_relaxed);                                        \
    cudaError_t e__ = cudaGetLastError();                                                      \
    if (e__ != cudaSuccess) return fail("launch of %s failed: %s", what, cudaGetErrorString(e__)); \
  } while (0)

#define TRY(expr)            \
  do {                       \
    int rc__ = (expr);       \
    if (rc__) return rc__;   \
  } while (0)

// ---------------------------------------------------------------------------------------------------
// optional per-launch profiler (CUDA events around every kernel, grouped by kernel class)
// ---------------------------------------------------------------------------------------------------
namespace {
// tensor-core launches are split at 64 channels: below, a layer moves 48-352 FLOP per byte (under the ~215 FLOP/B ridge
// of the measured peaks: HBM / L2 roofline); above, the tensor pipe is the roofline.  Both carry FLOPs and bytes.
enum : int { PC_TC_CONV = 0, PC_TC_WGRAD, PC_FFMA_CONV, PC_FFMA_WGRAD, PC_POST, PC_FOLD, PC_MISC, PC_TC_CONV_S, PC_TC_WGRAD_S, PC_COUNT };
const char* kProfNames[PC_COUNT] = {"tc_conv_c>=128(fwd+dgrad)", "tc_wgrad_c>=128", "ffma_conv", "ffma_wgrad", "conv_post", "weight_norm_fold", "misc",
                                    "tc_conv_c<=64(fwd+dgrad,pair)", "tc_wgrad_c<=64"};
struct ProfRec { int cls; cudaEvent_t a, b; double flops, bytes; std::string tag; };
bool g_prof_on = false;
std::vector<ProfRec> g_prof;
std::vector<cudaEvent_t> g_prof_pool;
cudaEvent_t prof_event() {
  cudaEvent_t e;
  if (!g_prof_pool.empty()) { e = g_prof_pool.back(); g_prof_pool.pop_back(); return e; }
  cudaEventCreate(&e);
  return e;
}
struct ProfScope {
  cudaStream_t s; bool on;
  ProfScope(int cls, double flops, double bytes, cudaStream_t stream, const char* tag = "") : s(stream), on(g_prof_on) {
    if (!on) return;
    ProfRec r{cls, prof_event(), prof_event(), flops, bytes, tag};
    cudaEventRecord(r.a, s);
    g_prof.push_back(r);
  }
  ~ProfScope() { if (on) cudaEventRecord(g_prof.back().b, s); }
};
}  // namespace

// Phase timer (VCD_PHASES=1): CUDA events on the caller-visible stream at the boundaries of the forward core and
// of every backward segment, in the real (graph-replay, multi-stream) execution mode.
namespace {
struct PhaseRec { std::string name; cudaEvent_t a, b; };
std::vector<PhaseRec> g_phases;
bool phases_on() { static const bool on = getenv("VCD_PHASES") != nullptr; return on; }
struct PhaseScope {
  cudaStream_t s; bool on;
  PhaseScope(const char* name, cudaStream_t stream) : s(stream), on(phases_on()) {
    if (!on) return;
    PhaseRec r{name, nullptr, nullptr};
    cudaEventCreate(&r.a); cudaEventCreate(&r.b);
    cudaEventRecord(r.a, s);
    g_phases.push_back(r);
  }
  ~PhaseScope() { if (on) cudaEventRecord(g_phases.back().b, s); }
};
}  // namespace
extern "C" int vcd_phase_dump(int reset) {
  cudaDeviceSynchronize();
  for (const PhaseRec& r : g_phases) {
    float t = 0.f;
    cudaEventElapsedTime(&t, r.a, r.b);
    fprintf(stderr, "[phase] %-28s %8.3f ms\n", r.name.c_str(), t);
  }
  if (reset) {
    for (const PhaseRec& r : g_phases) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    g_phases.clear();
  }
  return 0;
}

extern "C" int vcd_profile_enable(int on) { g_prof_on = on != 0; return 0; }
extern "C" int vcd_profile_num_classes(void) { return PC_COUNT; }
extern "C" const char* vcd_profile_class_name(int c) { return c >= 0 && c < PC_COUNT ? kProfNames[c] : nullptr; }
// Sums over the records since the last reset: device ms, launches, algorithmic flops and bytes per class.
extern "C" int vcd_profile_read(int reset, double* ms, uint64_t* launches, double* flops, double* bytes) {
  cudaDeviceSynchronize();
  for (int c = 0; c < PC_COUNT; ++c) { ms[c] = 0; launches[c] = 0; flops[c] = 0; bytes[c] = 0; }
  for (const ProfRec& r : g_prof) {
    float t = 0.f;
    cudaEventElapsedTime(&t, r.a, r.b);
    ms[r.cls] += t; launches[r.cls] += 1; flops[r.cls] += r.flops; bytes[r.cls] += r.bytes;
  }
  if (reset) {
    for (const ProfRec& r : g_prof) { g_prof_pool.push_back(r.a); g_prof_pool.push_back(r.b); }
    g_prof.clear();
  }
  return 0;
}

// Writes one CSV line per recorded launch (class, tag, ms, gflop) to `path`; does not reset.
extern "C" int vcd_profile_dump(const char* path) {
  cudaDeviceSynchronize();
  FILE* f = fopen(path, "w");
  if (!f) return fail("vcd_profile_dump: cannot open %s", path);
  fprintf(f, "class,tag,ms,gflop\n");
  for (const ProfRec& r : g_prof) {
    float t = 0.f;
    cudaEventElapsedTime(&t, r.a, r.b);
    fprintf(f, "%s,%s,%.6f,%.4f\n", kProfNames[r.cls], r.tag.c_str(), t, r.flops / 1e9);
  }
  fclose(f);
  return 0;
}

extern "C" const char* vcd_version(void) { return "vcd 0.1 (sm_100a; tcgen05/TMA + FFMA paths)"; }
extern "C" const char* vcd_last_error(void) { return g_err; }
extern "C" uint64_t vcd_launch_count(int reset) {
  return reset ? g_launches.exchange(0) : g_launches.load();
}

static inline size_t esize(int mode) { return mode == VCD_MODE_FP32 ? 4 : 2; }
static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ---------------------------------------------------------------------------------------------------
// plan construction
// ---------------------------------------------------------------------------------------------------
static int add_param(vcd_plan* p, const std::string& name, std::initializer_list<int64_t> shape) {
  ParamInfo pi;
  pi.name = name;
  pi.ndim = static_cast<int>(shape.size());
  pi.numel = 1;
  int i = 0;
  pi.shape[0] = pi.shape[1] = pi.shape[2] = 1;
  for (int64_t s : shape) {
    pi.shape[i++] = s;
    pi.numel *= s;
  }
  p->params.push_back(pi);
  return static_cast<int>(p->params.size()) - 1;
}

static int add_conv(vcd_plan* p, const std::string& name, int cin, int cout, int k, int dil, bool wn, bool bias) {
  Layer L{};
  L.name = name;
  L.kind = LK_CONV;
  L.cin = cin; L.cout = cout; L.k = k; L.dil = dil; L.u = 1;
  L.pad = dil * (k - 1) / 2;
  L.wn = wn;
  // state_dict order of old-style weight_norm: bias, weight_g, weight_v ; plain conv: weight, bias
  if (wn) {
    L.p_b = bias ? add_param(p, name + ".bias", {cout}) : -1;
    L.p_g = add_param(p, name + ".weight_g", {cout, 1, 1});
    L.p_w = add_param(p, name + ".weight_v", {cout, cin, k});
  } else {
    L.p_w = add_param(p, name + ".weight", {cout, cin, k});
    L.p_g = -1;
    L.p_b = bias ? add_param(p, name + ".bias", {cout}) : -1;
  }
  L.fwd = ConvGeo{k, cin, cout, 1, -L.pad, dil, 1, 0, cout};
  L.dgr = ConvGeo{k, cout, cin, 1, -L.pad, dil, 1, 0, cin};
  L.wgr = L.fwd;
  L.map_fwd = WeightMap{SRC_CONV_FWD, cin, cout, k, 1};
  L.map_dgr = WeightMap{SRC_CONV_DGRAD, cin, cout, k, 1};
  p->layers.push_back(L);
  return static_cast<int>(p->layers.size()) - 1;
}

static int add_convt(vcd_plan* p, const std::string& name, int cin, int cout, int k, int u) {
  Layer L{};
  L.name = name;
  L.kind = LK_CONVT;
  L.cin = cin; L.cout = cout; L.k = k; L.dil = 1; L.u = u;
  L.pad = (k - u) / 2;
  L.wn = true;
  L.p_b = add_param(p, name + ".bias", {cout});
  L.p_g = add_param(p, name + ".weight_g", {cin, 1, 1});
  L.p_w = add_param(p, name + ".weight_v", {cin, cout, k});
  const int m = (k + u - 1) / u;
  // forward, scatter form: Z[q][(r,co)] = sum_{s<m} x[q-s] W[:,co,s*u+r]  ->  y[q*u + r - pad][co]
  L.fwd = ConvGeo{m, cin, u * cout, 1, 0, -1, u, L.pad, cout};
  // data gradient on the phase-packed ("Z") output gradient DY[q][(r,co)] = dy[q*u + r - pad][co]:
  //   dx[i][ci] = sum_{s<m} sum_{(r,co)} DY[i+s][(r,co)] W[ci][co][s*u+r]      -- a plain m-tap convolution
  L.dgr = ConvGeo{m, u * cout, cin, 1, 0, 1, 1, 0, cin};
  // weight gradient: same taps as fwd, DY taken as an ordinary [u*cout]-channel tensor
  L.wgr = ConvGeo{m, cin, u * cout, 1, 0, -1, 1, 0, u * cout};
  L.map_fwd = WeightMap{SRC_CONVT_FWD, cin, cout, k, u};
  L.map_dgr = WeightMap{SRC_CONVT_DGRAD, cin, cout, k, u};
  p->layers.push_back(L);
  return static_cast<int>(p->layers.size()) - 1;
}

template <typename J>
static int upload_jobs(const std::vector<J>& jobs, J** dptr) {
  *dptr = nullptr;
  if (jobs.empty()) return 0;
  CU_TRY(cudaMalloc(dptr, jobs.size() * sizeof(J)));
  CU_TRY(cudaMemcpy(*dptr, jobs.data(), jobs.size() * sizeof(J), cudaMemcpyHostToDevice));
  return 0;
}

extern "C" int vcd_plan_create(const vcd_config* cfg, vcd_plan** out_plan) {
  if (!cfg || !out_plan) return fail("vcd_plan_create: null argument");
  *out_plan = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail("vcd_plan_create: no CUDA device available (this library has no CPU fallback)");
  const int S = cfg->num_upsamples, NB = cfg->num_kernels;
  if (S < 1 || S > VCD_MAX_UPSAMPLES) return fail("num_upsamples must be in [1,%d]", VCD_MAX_UPSAMPLES);
  if (NB < 1 || NB > VCD_MAX_KERNELS) return fail("num_kernels must be in [1,%d]", VCD_MAX_KERNELS);
  if (cfg->initial_channel % 8) return fail("initial_channel (%d) must be a multiple of 8", cfg->initial_channel);
  if (cfg->gin_channels < 0) return fail("gin_channels must be >= 0");
  if (cfg->upsample_initial_channel % (8 << S))
    return fail("upsample_initial_channel (%d) must be a multiple of %d so that every stage keeps a multiple "
                "of 8 channels (blocked channels-last layout)", cfg->upsample_initial_channel, 8 << S);
  for (int i = 0; i < S; ++i) {
    const int u = cfg->upsample_rates[i], k = cfg->upsample_kernel_sizes[i];
    if (u < 1 || k < u || ((k - u) & 1))
      return fail("upsample %d: need kernel >= rate and (kernel - rate) even (got k=%d, u=%d)", i, k, u);
  }
  for (int j = 0; j < NB; ++j) {
    if (!(cfg->resblock_kernel_sizes[j] & 1)) return fail("resblock kernel sizes must be odd");
    for (int d = 0; d < (cfg->resblock == 1 ? 3 : 2); ++d)
      if (cfg->resblock_dilation_sizes[j][d] < 1) return fail("resblock dilations must be >= 1");
  }

  vcd_plan* p = new vcd_plan();
  p->cfg = *cfg;
  CU_TRY(cudaGetDevice(&p->device));
  CU_TRY(cudaDeviceGetAttribute(&p->num_sms, cudaDevAttrMultiProcessorCount, p->device));

  // ---- layers + parameters, in the reference's state_dict order ----
  const int C0 = cfg->upsample_initial_channel;
  p->l_pre = add_conv(p, "conv_pre", cfg->initial_channel, C0, 7, 1, false, true);
  p->stages.resize(S);
  int ch = C0;
  p->hop = 1;
  for (int i = 0; i < S; ++i) {
    StageDesc& st = p->stages[i];
    st.cin = ch; st.cout = ch / 2; st.u = cfg->upsample_rates[i]; st.k = cfg->upsample_kernel_sizes[i];
    st.up_layer = add_convt(p, "ups." + std::to_string(i), st.cin, st.cout, st.k, st.u);
    ch /= 2;
    p->hop *= st.u;
  }
  ch = C0;
  const int npairs = cfg->resblock == 1 ? 3 : 2;
  for (int i = 0; i < S; ++i) {
    ch /= 2;
    StageDesc& st = p->stages[i];
    st.convs.resize(NB);
    for (int j = 0; j < NB; ++j) {
      const std::string rb = "resblocks." + std::to_string(i * NB + j);
      const int k = cfg->resblock_kernel_sizes[j];
      st.convs[j].resize(npairs);
      if (cfg->resblock == 1) {
        // state_dict order: convs1.0..2 then convs2.0..2 (modules.py:189-201)
        for (int q = 0; q < npairs; ++q)
          st.convs[j][q].push_back(add_conv(p, rb + ".convs1." + std::to_string(q), ch, ch, k,
                                            cfg->resblock_dilation_sizes[j][q], true, true));
        for (int q = 0; q < npairs; ++q)
          st.convs[j][q].push_back(add_conv(p, rb + ".convs2." + std::to_string(q), ch, ch, k, 1, true, true));
      } else {
        for (int q = 0; q < npairs; ++q)
          st.convs[j][q].push_back(add_conv(p, rb + ".convs." + std::to_string(q), ch, ch, k,
                                            cfg->resblock_dilation_sizes[j][q], true, true));
      }
    }
  }
  p->p_post_w = add_param(p, "conv_post.weight", {1, ch, 7});
  if (cfg->gin_channels) {
    p->p_cond_w = add_param(p, "cond.weight", {C0, cfg->gin_channels, 1});
    p->p_cond_b = add_param(p, "cond.bias", {C0});
  }

  // ---- backward segments ----
  p->layers[p->l_pre].segment = S;
  for (int i = 0; i < S; ++i) {
    const int seg = S - 1 - i;
    p->layers[p->stages[i].up_layer].segment = seg;
    for (auto& br : p->stages[i].convs)
      for (auto& pr : br)
        for (int l : pr) p->layers[l].segment = seg;
  }

  // ---- arenas ----
  long long nf32 = 0, nbf = 0, nnorm = 0;
  for (Layer& L : p->layers) {
    const long long nf = 1LL * L.fwd.taps * L.fwd.K * L.fwd.N, nd = 1LL * L.dgr.taps * L.dgr.K * L.dgr.N;
    L.f32_fwd = nf32; nf32 += nf;
    L.f32_dgr = nf32; nf32 += nd;
    tc_layer_eligibility(L);
    L.tc_fwd = L.tc_dgr = -1;
    if (L.tc_ok_fwd) { L.tc_fwd = nbf; nbf += nf; }
    if (L.tc_ok_dgr) { L.tc_dgr = nbf; nbf += nd; }
    if (L.wn) { L.norm_off = static_cast<int>(nnorm); nnorm += p->params[L.p_w].shape[0]; }
  }
  // gradient scratch, grouped by segment so each segment zeroes one contiguous range
  p->segments.resize(S + 1);
  long long ng = 0;
  for (int seg = 0; seg <= S; ++seg) {
    p->segments[seg].scratch_begin = ng;
    if (seg == 0) { p->post_dw = ng; ng += p->params[p->p_post_w].numel; }
    for (Layer& L : p->layers) {
      if (L.segment != seg) continue;
      L.dwp = ng; ng += 1LL * L.fwd.taps * L.fwd.K * L.fwd.N;
      L.dbias = -1;
      if (L.p_b >= 0 && &L != &p->layers[p->l_pre]) { L.dbias = ng; ng += L.cout; }
      ng = (ng + 63) / 64 * 64;
    }
    p->segments[seg].scratch_end = ng;
  }
  p->n_f32 = nf32; p->n_bf16 = nbf; p->n_norms = nnorm; p->n_gscratch = ng;
  CU_TRY(cudaMalloc(&p->d_f32, std::max<long long>(nf32, 1) * sizeof(float)));
  CU_TRY(cudaMalloc(&p->d_bf16, std::max<long long>(nbf, 1) * sizeof(bf16)));
  CU_TRY(cudaMalloc(&p->d_norms, std::max<long long>(nnorm, 1) * sizeof(float)));
  CU_TRY(cudaMalloc(&p->d_gscratch, std::max<long long>(ng, 1) * sizeof(float)));
  const size_t np = p->params.size();
  CU_TRY(cudaMalloc(&p->d_params, np * sizeof(float*)));
  CU_TRY(cudaMalloc(&p->d_dparams, np * sizeof(float*)));
  p->h_params.assign(np, nullptr);
  p->h_dparams.assign(np, nullptr);
  p->is_weight_g.assign(np, 0);
  for (const Layer& L : p->layers)
    if (L.p_g >= 0) p->is_weight_g[L.p_g] = 1;

  // ---- fold job tables ----
  // layers handled by the bandwidth-shaped fold kernels (fold_fast.cuh): both operand formats on the tensor-core path,
  // whole taps per upsample phase (k % u == 0), rows in groups of 8
  static const int fast_fold = tc_env_int("VCD_FAST_FOLD", 1);
  auto fast_shape = [&](const Layer& L) {
    const ParamInfo& pv = p->params[L.p_w];
    return fast_fold && L.k % L.u == 0 && pv.shape[0] % 8 == 0 && pv.shape[1] % 8 == 0 && (pv.numel / pv.shape[0]) % 4 == 0 &&
           fold_kshift(L.k) >= 0 && ((L.k & 1) || L.k % 4 == 0);
  };
  auto fast_pack_ok = [&](const Layer& L) { return fast_shape(L) && L.tc_ok_fwd && L.tc_ok_dgr; };
  for (int bf = 0; bf < 2; ++bf) {
    std::vector<NormJob> nj;
    int blk = 0;
    for (const Layer& L : p->layers) {
      if (!L.wn || (bf && fast_pack_ok(L))) continue;
      const ParamInfo& pv = p->params[L.p_w];
      NormJob j{L.p_w, static_cast<int>(pv.shape[0]), static_cast<int>(pv.numel / pv.shape[0]), L.norm_off, blk};
      blk += j.rows;
      nj.push_back(j);
    }
    if (!bf) {
      p->n_norm_jobs = static_cast<int>(nj.size());
      p->n_norm_blocks = blk;
      TRY(upload_jobs(nj, &p->d_norm_jobs));
    } else {
      p->n_norm_jobs_bf16 = static_cast<int>(nj.size());
      p->n_norm_blocks_bf16 = blk;
      TRY(upload_jobs(nj, &p->d_norm_jobs_bf16));
    }
  }
  {
    std::vector<FastPackJob> fj;
    int blk = 0;
    for (const Layer& L : p->layers) {
      if (!fast_pack_ok(L)) continue;
      const ParamInfo& pv = p->params[L.p_w];
      FastPackJob j{};
      j.p_w = L.p_w; j.p_g = L.p_g; j.norm_off = L.wn ? L.norm_off : 0;
      j.rows = static_cast<int>(pv.shape[0]); j.inner = static_cast<int>(pv.shape[1]); j.k = L.k; j.u = L.u;
      j.is_convt = L.kind == LK_CONVT ? 1 : 0;
      j.kshift = fold_kshift(L.k);
      // format R = "row is the GEMM column": Conv1d forward / ConvTranspose1d data gradient
      j.nt_r = j.is_convt ? L.nt_dgr : L.nt_fwd;
      j.nt_c = j.is_convt ? L.nt_fwd : L.nt_dgr;
      j.dst_r = j.is_convt ? L.tc_dgr : L.tc_fwd;
      j.dst_c = j.is_convt ? L.tc_fwd : L.tc_dgr;
      j.first_block = blk;
      blk += j.rows / kFoldRows;
      p->fast_pack_smem = std::max(p->fast_pack_smem, sizeof(float) * kFoldRows * static_cast<size_t>(fold_srow(j.inner, j.k)));
      fj.push_back(j);
    }
    p->n_fast_pack_jobs = static_cast<int>(fj.size());
    p->n_fast_pack_blocks = blk;
    TRY(upload_jobs(fj, &p->d_fast_pack));
    if (p->fast_pack_smem > 200 * 1024) return fail("internal: fast fold tile of %zu bytes", p->fast_pack_smem);
    if (blk) CU_TRY(cudaFuncSetAttribute(wn_pack_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(p->fast_pack_smem)));
  }
  for (int mode = 0; mode < 2; ++mode) {
    std::vector<PackJob> pj;
    int blk = 0;
    auto push = [&](const Layer& L, bool dgrad, int fmt) {
      const ConvGeo& g = dgrad ? L.dgr : L.fwd;
      PackJob j{};
      j.map = dgrad ? L.map_dgr : L.map_fwd;
      j.p_w = L.p_w; j.p_g = L.p_g; j.norm_off = L.wn ? L.norm_off : 0;
      j.taps = g.taps; j.K = g.K; j.N = g.N;
      j.NT = dgrad ? L.nt_dgr : L.nt_fwd;
      j.fmt = fmt;
      j.round_bf16 = (mode == VCD_MODE_BF16 && fmt == FMT_F32) ? 1 : 0;
      j.dst_off = fmt == FMT_F32 ? (dgrad ? L.f32_dgr : L.f32_fwd) : (dgrad ? L.tc_dgr : L.tc_fwd);
      j.numel = 1LL * g.taps * g.K * g.N;
      j.first_block = blk;
      blk += static_cast<int>(((fmt == FMT_TC ? j.numel / 8 : j.numel) + 255) / 256);
      pj.push_back(j);
    };
    for (const Layer& L : p->layers) {
      if (mode == VCD_MODE_BF16 && fast_pack_ok(L)) continue;   // written by wn_pack_fast_kernel
      const bool tcf = mode == VCD_MODE_BF16 && L.tc_ok_fwd, tcd = mode == VCD_MODE_BF16 && L.tc_ok_dgr;
      push(L, false, tcf ? FMT_TC : FMT_F32);
      push(L, true, tcd ? FMT_TC : FMT_F32);
    }
    p->n_pack_jobs[mode] = static_cast<int>(pj.size());
    p->n_pack_blocks[mode] = blk;
    TRY(upload_jobs(pj, &p->d_pack_jobs[mode]));
  }
  // ---- unfold job tables (per segment) ----
  size_t unfold_smem_max = 0;
  for (int seg = 0; seg <= S; ++seg) {
    std::vector<UnfoldJob> uj;
    int blk = 0;
    SegmentJobs& sj = p->segments[seg];
    auto copy_job = [&](int param, long long src) {
      UnfoldJob j{};
      j.kind = 1; j.p_w = param; j.p_g = -1; j.src_off = src; j.numel = p->params[param].numel;
      j.rows = static_cast<int>((j.numel + 255) / 256);
      j.first_block = blk; blk += j.rows;
      uj.push_back(j);
      sj.params.push_back(param);
    };
    if (seg == 0) copy_job(p->p_post_w, p->post_dw);
    sj.lead_jobs = static_cast<int>(uj.size());
    sj.lead_blocks = blk;
    for (const Layer& L : p->layers) {
      if (L.segment != seg) continue;
      const ParamInfo& pw = p->params[L.p_w];
      UnfoldJob j{};
      j.map = L.map_fwd; j.kind = 0; j.p_w = L.p_w; j.p_g = L.p_g; j.norm_off = L.wn ? L.norm_off : 0;
      j.rows = static_cast<int>(pw.shape[0]); j.row_len = static_cast<int>(pw.numel / pw.shape[0]);
      j.K = L.fwd.K; j.N = L.fwd.N; j.src_off = L.dwp;
      j.first_block = blk; blk += j.rows;
      uj.push_back(j);
      sj.params.push_back(L.p_w);
      if (L.p_g >= 0) sj.params.push_back(L.p_g);
      if (L.dbias >= 0) copy_job(L.p_b, L.dbias);
    }
    if (seg == S) {  // written directly by cond_bwd_kernel
      sj.params.push_back(p->layers[p->l_pre].p_b);
      if (p->p_cond_w >= 0) { sj.params.push_back(p->p_cond_w); sj.params.push_back(p->p_cond_b); }
    }
    sj.njobs = static_cast<int>(uj.size());
    sj.nblocks = blk;
    TRY(upload_jobs(uj, &sj.d_jobs));
    // the same jobs for wn_unfold_fast_kernel (8 rows per block, 1024 elements per copy block)
    bool all_fast = true;
    for (const Layer& L : p->layers)
      if (L.segment == seg && !fast_shape(L)) all_fast = false;
    if (all_fast) {
      std::vector<FastUnfoldJob> fj;
      int fblk = 0;
      for (size_t n = 0; n < uj.size(); ++n) {
        const UnfoldJob& o = uj[n];
        FastUnfoldJob j{};
        j.kind = o.kind; j.p_w = o.p_w; j.p_g = o.p_g; j.norm_off = o.norm_off; j.src_off = o.src_off; j.numel = o.numel;
        j.first_block = fblk;
        if (o.kind == 1) {
          fblk += static_cast<int>((o.numel + 2047) / 2048);
        } else {
          const ParamInfo& pv = p->params[o.p_w];
          j.rows = o.rows; j.inner = static_cast<int>(pv.shape[1]); j.k = o.map.k;
          j.is_convt = o.map.src == SRC_CONVT_FWD ? 1 : 0;
          j.u = j.is_convt ? o.map.u : 1;
          j.kshift = fold_kshift(j.k);
          j.K = o.K; j.N = o.N;
          fblk += j.rows / kFoldRows;
          sj.fast_smem = std::max(sj.fast_smem, sizeof(float) * kFoldRows * static_cast<size_t>(fold_srow(j.inner, j.k)));
        }
        if (static_cast<int>(n) + 1 == sj.lead_jobs) { sj.fast_lead_jobs = sj.lead_jobs; sj.fast_lead_blocks = fblk; }
        fj.push_back(j);
      }
      sj.fast_njobs = static_cast<int>(fj.size());
      sj.fast_nblocks = fblk;
      TRY(upload_jobs(fj, &sj.d_fast));
      unfold_smem_max = std::max(unfold_smem_max, sj.fast_smem);
    }
  }
  if (unfold_smem_max > 200 * 1024) return fail("internal: fast unfold tile of %zu bytes", unfold_smem_max);
  if (unfold_smem_max) CU_TRY(cudaFuncSetAttribute(wn_unfold_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(unfold_smem_max)));
  for (int i = 0; i < vcd_plan::kMaxAux; ++i) CU_TRY(cudaStreamCreateWithFlags(&p->aux[i], cudaStreamNonBlocking));
  CU_TRY(cudaStreamCreateWithFlags(&p->own, cudaStreamNonBlocking));
  CU_TRY(cudaEventCreateWithFlags(&p->hop_in, cudaEventDisableTiming));
  CU_TRY(cudaEventCreateWithFlags(&p->hop_out, cudaEventDisableTiming));
  for (auto& ev : p->seg_done) CU_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  CU_TRY(cudaEventCreateWithFlags(&p->post_done, cudaEventDisableTiming));
  for (int i = 0; i < vcd_plan::kMaxEvents; ++i) CU_TRY(cudaEventCreateWithFlags(&p->events[i], cudaEventDisableTiming));
  TRY(tc_plan_init(p));
  *out_plan = p;
  return 0;
}

extern "C" void vcd_plan_destroy(vcd_plan* p) {
  if (!p) return;
  cudaFree(p->d_f32); cudaFree(p->d_bf16); cudaFree(p->d_norms); cudaFree(p->d_gscratch);
  cudaFree(p->d_params); cudaFree(p->d_dparams); cudaFree(p->d_norm_jobs);
  cudaFree(p->d_pack_jobs[0]); cudaFree(p->d_pack_jobs[1]);
  for (auto& s : p->segments) { cudaFree(s.d_jobs); cudaFree(s.d_fast); }
  cudaFree(p->d_fast_pack); cudaFree(p->d_norm_jobs_bf16); cudaFree(p->d_turn); cudaFree(p->d_det);
  for (auto& kv : p->graphs) cudaGraphExecDestroy(kv.second);
  for (auto& kv : p->pad_tables) cudaFree(kv.second);
  if (p->own) cudaStreamDestroy(p->own);
  if (p->hop_in) cudaEventDestroy(p->hop_in);
  if (p->hop_out) cudaEventDestroy(p->hop_out);
  for (auto& ev : p->seg_done) if (ev) cudaEventDestroy(ev);
  if (p->post_done) cudaEventDestroy(p->post_done);
  for (int i = 0; i < vcd_plan::kMaxAux; ++i) if (p->aux[i]) cudaStreamDestroy(p->aux[i]);
  for (int i = 0; i < vcd_plan::kMaxEvents; ++i) if (p->events[i]) cudaEventDestroy(p->events[i]);
  delete p;
}

extern "C" int vcd_num_params(const vcd_plan* p) { return p ? static_cast<int>(p->params.size()) : 0; }
extern "C" int vcd_param_info(const vcd_plan* p, int i, const char** name, int64_t shape[3], int* ndim) {
  if (!p || i < 0 || i >= static_cast<int>(p->params.size())) return fail("vcd_param_info: bad index");
  const ParamInfo& pi = p->params[i];
  if (name) *name = pi.name.c_str();
  if (shape) { shape[0] = pi.shape[0]; shape[1] = pi.shape[1]; shape[2] = pi.shape[2]; }
  if (ndim) *ndim = pi.ndim;
  return 0;
}
extern "C" int64_t vcd_total_param_elems(const vcd_plan* p) {
  int64_t n = 0;
  if (p) for (auto& pi : p->params) n += pi.numel;
  return n;
}
extern "C" int vcd_hop(const vcd_plan* p) { return p ? p->hop : 0; }
extern "C" int vcd_num_backward_segments(const vcd_plan* p) { return p ? static_cast<int>(p->segments.size()) : 0; }
extern "C" int vcd_segment_params(const vcd_plan* p, int seg, int* idx, int cap) {
  if (!p || seg < 0 || seg >= static_cast<int>(p->segments.size())) return -1;
  const auto& v = p->segments[seg].params;
  for (int i = 0; i < static_cast<int>(v.size()) && i < cap; ++i) idx[i] = v[i];
  return static_cast<int>(v.size());
}

// ---------------------------------------------------------------------------------------------------
// workspace layout
// ---------------------------------------------------------------------------------------------------
// Only ACTIVATED tensors (storage type T) are kept between layers; fp32 is used for the running branch
// sums, the per-batch conditioning bias and the boundary tensors.  With save_for_backward the activations of
// every layer stay resident (they are the wgrad operands and the leaky_relu masks of the backward pass);
// without it every stage recycles one scratch set per ResBlock branch.
namespace {
struct StageWs {
  size_t ua = 0;
  std::vector<std::vector<size_t>> ma, xa;  // [branch][pair]
};
struct WsLayout {
  size_t xin = 0, cb = 0, dcb = 0;
  std::vector<size_t> a;
  std::vector<StageWs> st;
  size_t sum[2] = {0, 0};                   // fp32 running sums over ResBlock branches (fwd) / branch gradients (bwd)
  size_t Gi[3] = {0, 0, 0};                 // T: gradient w.r.t. the output of stage i lives in Gi[i % 3] (three sets: the
                                            // trailing weight-gradient kernels of segment s-1 still read theirs while segment s writes)
  // T [stage parity][branch][pair]: residual-stream gradient / mid gradient.  Two sets, alternating between consecutive
  // stages: the weight-gradient kernels of a backward segment trail behind the next segment's data-gradient chain
  // (they read these tensors), so a segment must not overwrite its predecessor's set.
  std::vector<std::vector<size_t>> Gt[2], dm[2];
  size_t duz[2] = {0, 0};                   // T: phase-packed gradient w.r.t. the upsample output (same alternation)
  size_t d0 = 0, dxb = 0;
  size_t total = 0;
  // pad-row zeroing jobs: forward phase 0 = boundary tensors, phase 1+i = stage i; backward phase = segment
  std::vector<std::vector<PadJob>> pad_fwd, pad_bwd;
};

WsLayout make_layout(const vcd_plan* p, int mode, int B, int T, bool save) {
  WsLayout w;
  const size_t es = esize(mode);
  size_t top = 0;
  auto alloc = [&](size_t bytes) {
    const size_t o = top;
    top = align_up(top + bytes, 256);
    return o;
  };
  const int S = static_cast<int>(p->stages.size()), NB = p->cfg.num_kernels;
  const int npairs = p->cfg.resblock == 1 ? 3 : 2;
  const int C0 = p->cfg.upsample_initial_channel, Cin0 = p->cfg.initial_channel;
  auto pad = [&](std::vector<PadJob>& v, size_t off, int C, int L) {
    if (!off && &v != &w.pad_fwd[0]) return;  // unallocated slot
    PadJob j{static_cast<long long>(off), B * (C / 8), L, static_cast<int>(es), 0};
    j.first_block = v.empty() ? 0 : v.back().first_block + v.back().arrays;
    v.push_back(j);
  };
  w.pad_fwd.resize(S + 1);
  w.pad_bwd.resize(S + 1);
  w.xin = alloc(es * blk_elems(B, Cin0, T));
  w.cb = alloc(4 * static_cast<size_t>(B) * C0);
  w.dcb = alloc(4 * static_cast<size_t>(B) * C0);
  w.a.resize(S + 1);
  w.a[0] = alloc(es * blk_elems(B, C0, T));
  pad(w.pad_fwd[0], w.xin, Cin0, T);
  pad(w.pad_fwd[0], w.a[0], C0, T);
  size_t emax = 0, zmax = 0;
  std::vector<size_t> E(S);
  std::vector<int> Ls(S + 1);
  Ls[0] = T;
  for (int i = 0; i < S; ++i) {
    const Layer& U = p->layers[p->stages[i].up_layer];
    zmax = std::max(zmax, blk_elems(B, U.wgr.N, Ls[i] + U.fwd.taps - 1));
    Ls[i + 1] = Ls[i] * p->stages[i].u;
    E[i] = blk_elems(B, p->stages[i].cout, Ls[i + 1]);
    emax = std::max(emax, E[i]);
    w.a[i + 1] = alloc(es * E[i]);
  }
  w.st.resize(S);
  std::vector<size_t> sh_ma(NB, 0), sh_xa0(NB, 0), sh_xa1(NB, 0);
  size_t sh_ua = 0;
  if (!save) {
    sh_ua = alloc(es * emax);
    for (int j = 0; j < NB; ++j) {
      if (p->cfg.resblock == 1) sh_ma[j] = alloc(es * emax);
      sh_xa0[j] = alloc(es * emax);
      sh_xa1[j] = alloc(es * emax);
    }
  }
  for (int i = 0; i < S; ++i) {
    StageWs& s = w.st[i];
    const int C = p->stages[i].cout, L = Ls[i + 1];
    std::vector<PadJob>& pj = w.pad_fwd[i + 1];
    s.ua = save ? alloc(es * E[i]) : sh_ua;
    pad(pj, s.ua, C, L);
    s.ma.assign(NB, std::vector<size_t>(npairs, 0));
    s.xa.assign(NB, std::vector<size_t>(npairs, 0));
    for (int j = 0; j < NB; ++j)
      for (int q = 0; q < npairs; ++q) {
        if (p->cfg.resblock == 1) {
          s.ma[j][q] = save ? alloc(es * E[i]) : sh_ma[j];
          if (save || q == 0) pad(pj, s.ma[j][q], C, L);
        }
        if (q < npairs - 1) {
          s.xa[j][q] = save ? alloc(es * E[i]) : ((q & 1) ? sh_xa1[j] : sh_xa0[j]);
          if (save || q < 2) pad(pj, s.xa[j][q], C, L);
        }
      }
    pad(pj, w.a[i + 1], C, L);
  }
  w.sum[0] = alloc(4 * emax);
  w.sum[1] = alloc(4 * emax);
  if (save) {
    for (int t = 0; t < 3; ++t) w.Gi[t] = alloc(es * emax);
    for (int par = 0; par < 2; ++par) {
      // set `par` serves the stages i with (i & 1) == par
      size_t em = 0, zm = 0;
      for (int i = par; i < S; i += 2) {
        const Layer& U = p->layers[p->stages[i].up_layer];
        em = std::max(em, E[i]);
        zm = std::max(zm, blk_elems(B, U.wgr.N, Ls[i] + U.fwd.taps - 1));
      }
      w.Gt[par].assign(NB, std::vector<size_t>(npairs, 0));
      w.dm[par].assign(NB, std::vector<size_t>(npairs, 0));
      if (em == 0) continue;
      for (int j = 0; j < NB; ++j)
        for (int q = 0; q < npairs; ++q) {
          if (q > 0) w.Gt[par][j][q] = alloc(es * em);
          if (p->cfg.resblock == 1) w.dm[par][j][q] = alloc(es * em);
        }
      w.duz[par] = alloc(es * zm);
    }
    w.d0 = alloc(es * blk_elems(B, C0, T));
    w.dxb = alloc(4 * blk_elems(B, Cin0, T));
    // backward pads: segment s works on stage i = S-1-s and produces the stage gradient consumed by segment s+1
    pad(w.pad_fwd[S], w.Gi[(S - 1) % 3], p->stages[S - 1].cout, Ls[S]);  // written by conv_post's backward
    pad(w.pad_fwd[0], w.d0, C0, T);
    for (int seg = 0; seg < S; ++seg) {
      const int i = S - 1 - seg;
      const int C = p->stages[i].cout, L = Ls[i + 1];
      for (int j = 0; j < NB; ++j)
        for (int q = 0; q < npairs; ++q) {
          if (q > 0) pad(w.pad_bwd[seg], w.Gt[i & 1][j][q], C, L);
          if (p->cfg.resblock == 1) pad(w.pad_bwd[seg], w.dm[i & 1][j][q], C, L);
        }
      if (i > 0) pad(w.pad_bwd[seg], w.Gi[(i - 1) % 3], p->stages[i - 1].cout, Ls[i]);
    }
  }
  w.total = top;
  return w;
}

// Pad-zeroing job tables are uploaded once per layout (outside any stream capture) and cached in the plan.
int ensure_pad_tables(vcd_plan* p, const WsLayout& w, int mode, int B, int T, bool save) {
  if (p->pad_tables.size() >= 1024) {  // many distinct shapes (variable-length inference): start over.  Captured graphs
    CU_TRY(cudaDeviceSynchronize());   // hold the table addresses, so they go first.
    for (auto& kv : p->graphs) cudaGraphExecDestroy(kv.second);
    p->graphs.clear();
    p->graph_kernels.clear();
    for (auto& kv : p->pad_tables) cudaFree(kv.second);
    p->pad_tables.clear();
  }
  for (int bwd = 0; bwd < 2; ++bwd) {
    const auto& phases = bwd ? w.pad_bwd : w.pad_fwd;
    for (size_t ph = 0; ph < phases.size(); ++ph) {
      const std::vector<PadJob>& jobs = phases[ph];
      if (jobs.empty()) continue;
      const auto key = std::make_tuple(mode, B, T, save ? 1 : 0, static_cast<int>(ph), bwd);
      if (p->pad_tables.count(key)) continue;
      PadJob* d = nullptr;
      CU_TRY(cudaMalloc(&d, jobs.size() * sizeof(PadJob)));
      CU_TRY(cudaMemcpy(d, jobs.data(), jobs.size() * sizeof(PadJob), cudaMemcpyHostToDevice));
      p->pad_tables.emplace(key, static_cast<void*>(d));
    }
  }
  return 0;
}

// Launch the pad-zeroing kernel for one phase.
int launch_pads(vcd_plan* p, const std::vector<PadJob>& jobs, int mode, int B, int T, bool save, int phase, bool bwd,
                char* ws, cudaStream_t stream) {
  if (jobs.empty()) return 0;
  const auto key = std::make_tuple(mode, B, T, save ? 1 : 0, phase, bwd ? 1 : 0);
  auto it = p->pad_tables.find(key);
  if (it == p->pad_tables.end()) return fail("internal: pad table missing");
  const int blocks = jobs.back().first_block + jobs.back().arrays;
  ProfScope ps__(PC_MISC, 0, 0, stream);
  pad_zero_kernel<<<blocks, 64, 0, stream>>>(static_cast<const PadJob*>(it->second), static_cast<int>(jobs.size()), ws);
  LAUNCH_CHECK("pad_zero_kernel");
  return 0;
}
}  // namespace

extern "C" size_t vcd_workspace_bytes(const vcd_plan* p, int mode, int B, int T, int save) {
  if (!p || B < 1 || T < 1) return 0;
  return make_layout(p, mode, B, T, save != 0).total;
}

// ---------------------------------------------------------------------------------------------------
// fold
// ---------------------------------------------------------------------------------------------------
extern "C" int vcd_fold_weights(vcd_plan* p, int mode, const float* const* params, void* stream_) {
  if (!p || !params) return fail("vcd_fold_weights: null argument");
  if (mode != VCD_MODE_FP32 && mode != VCD_MODE_BF16) return fail("bad mode %d", mode);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const size_t np = p->params.size();
  bool changed = false;
  for (size_t i = 0; i < np; ++i) {
    // a NULL weight_g means "this layer carries no weight norm" (remove_weight_norm, modules.py:218-222): its weight_v
    // slot then holds the baked weight and is used as is
    if (!params[i] && !p->is_weight_g[i]) return fail("vcd_fold_weights: parameter %s is null", p->params[i].name.c_str());
    changed |= p->h_params[i] != params[i];
  }
  if (changed) {
    std::copy(params, params + np, p->h_params.begin());
    ++p->params_version;  // bias pointers are baked into captured kernel parameters
    // pageable->device async copy is staged by the runtime before returning, so h_params may change later
    CU_TRY(cudaMemcpyAsync(p->d_params, p->h_params.data(), np * sizeof(float*), cudaMemcpyHostToDevice, stream));
  }
  PhaseScope ph__("fold", stream);
  ProfScope ps__(PC_FOLD, 0, 4.0 * p->n_f32, stream);
  // the float4 row loads of the fast kernel need 16-byte aligned parameter tensors (true for separately allocated
  // tensors; a parameter that is a view into a flat buffer may not be): fall back to the generic tables otherwise
  bool fast = mode == VCD_MODE_BF16 && p->n_fast_pack_blocks > 0;
  for (size_t i = 0; fast && i < np; ++i)
    if (reinterpret_cast<uintptr_t>(params[i]) & 15) fast = false;
  if (mode == VCD_MODE_BF16 && p->n_fast_pack_blocks > 0 && !fast)
    return fail("vcd_fold_weights: parameter tensors must be 16-byte aligned");
  const bool bf = mode == VCD_MODE_BF16;
  const int nblk = bf ? p->n_norm_blocks_bf16 : p->n_norm_blocks;
  if (nblk) {
    wn_norm_kernel<<<nblk, 128, 0, stream>>>(bf ? p->d_norm_jobs_bf16 : p->d_norm_jobs, bf ? p->n_norm_jobs_bf16 : p->n_norm_jobs,
                                             p->d_params, p->d_norms);
    LAUNCH_CHECK("wn_norm_kernel");
  }
  if (fast) {
    wn_pack_fast_kernel<<<p->n_fast_pack_blocks, kFoldThreads, p->fast_pack_smem, stream>>>(p->d_fast_pack, p->n_fast_pack_jobs, p->d_params,
                                                                                  p->d_norms, p->d_bf16);
    LAUNCH_CHECK("wn_pack_fast_kernel");
  }
  if (p->n_pack_blocks[mode]) {
    wn_pack_kernel<<<p->n_pack_blocks[mode], 256, 0, stream>>>(p->d_pack_jobs[mode], p->n_pack_jobs[mode],
                                                                p->d_params, p->d_norms, p->d_f32, p->d_bf16);
    LAUNCH_CHECK("wn_pack_kernel");
  }
  p->folded[mode] = true;
  p->folded[1 - mode] = false;   // the fp32 arena is shared between the modes (bf16 mode stores bf16-rounded values in it)
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// launch helpers
// ---------------------------------------------------------------------------------------------------
namespace {

struct Ctx {
  vcd_plan* p;
  int mode, B;
  cudaStream_t main;
  char* ws;
  bool serial;  // profiling: everything on the caller's stream so per-launch event times are isolated

  // stream j of the branch set (0 = caller's stream) / weight-gradient side streams
  cudaStream_t branch(int j) const { return (serial || j == 0) ? main : p->aux[j - 1]; }
  cudaStream_t side(int i) const { return serial ? main : p->aux[VCD_MAX_KERNELS - 1 + (i & 3)]; }
  cudaEvent_t record(cudaStream_t s) const {
    cudaEvent_t e = p->events[p->next_event];
    p->next_event = (p->next_event + 1) % vcd_plan::kMaxEvents;
    cudaEventRecord(e, s);
    return e;
  }
  void wait(cudaStream_t s, cudaEvent_t e) const { cudaStreamWaitEvent(s, e, 0); }
  void order(cudaStream_t from, cudaStream_t to) const {
    if (from != to) wait(to, record(from));
  }
};

template <typename T>
int launch_gconv_simt(const Ctx& c, cudaStream_t st, const void* in, const float* w, const ConvGeo& g, const Epilogue& e,
                      int Lin, int Lq, int Lout, double flops, const char* tag) {
  ProfScope ps__(PC_FFMA_CONV, flops, 0, st, tag);
  if (Lq <= 128) {
    dim3 grid((Lq + 127) / 128, g.N / 8, c.B);
    gconv_simt_kernel<T, 1><<<grid, 128, 0, st>>>(static_cast<const T*>(in), w, g, e, Lin, Lq, Lout);
  } else {
    dim3 grid((Lq + 255) / 256, g.N / 8, c.B);
    gconv_simt_kernel<T, 2><<<grid, 128, 0, st>>>(static_cast<const T*>(in), w, g, e, Lin, Lq, Lout);
  }
  LAUNCH_CHECK("gconv_simt_kernel");
  return 0;
}

// Algorithmic FLOPs of one pass over layer L: 2*Cin*Cout*k per forward-input position
// (ConvTranspose1d) or per output position (Conv1d).
double layer_flops(const Layer& L, int B, int Lfwd_in) {
  return 2.0 * L.cin * L.cout * L.k * static_cast<double>(B) * Lfwd_in;
}

// One convolution (forward or data-gradient direction) of layer L on stream st.
int run_conv(const Ctx& c, cudaStream_t st, const Layer& L, bool dgrad, const void* in, int Lin, int Lq, int Lout,
             Epilogue e, double flops, WgFuse* wg = nullptr) {
  const ConvGeo& g = dgrad ? L.dgr : L.fwd;
  // timing experiments only (results are wrong): bit 0 skips forward, bit 1 data-gradient, bit 2 weight-gradient launches
  static const int dbg_skip = tc_env_int("VCD_DEBUG_SKIP", 0);
  if (dbg_skip & (dgrad ? 2 : 1)) return 0;
  if (c.mode == VCD_MODE_BF16 && (dgrad ? L.tc_ok_dgr : L.tc_ok_fwd)) {
    // algorithmic bytes: the input tensor, every bf16 / fp32 epilogue operand and output once, the weights once
    const double esz = 2.0;
    const double n_out = static_cast<double>(c.B) * Lout * g.creal;
    const double bytes = esz * c.B * static_cast<double>(Lin) * g.K + esz * n_out * ((e.out_t ? 1 : 0) + (e.mask ? 1 : 0) + (e.res_t ? 1 : 0)) +
                         4.0 * n_out * ((e.res2 ? 1 : 0) + (e.out_raw ? 1 : 0)) + esz * g.taps * g.K * g.N;
    ProfScope ps__(L.cin <= 64 && L.cout <= 64 ? PC_TC_CONV_S : PC_TC_CONV, flops, bytes, st, (L.name + (dgrad ? ":dgrad" : ":fwd")).c_str());
    const int rc = tc_run_conv(c.p, L, dgrad, in, c.B, Lin, Lq, Lout, e, st, g_launches, g_err, sizeof(g_err), wg);
    if (wg && wg->fused && ps__.on) {   // the launch also produced the weight gradient: twice the algorithmic FLOPs
      g_prof.back().flops *= 2.0;
      g_prof.back().tag += "+wgrad";
    }
    return rc;
  }
  const float* w = c.p->d_f32 + (dgrad ? L.f32_dgr : L.f32_fwd);
  const std::string tag = L.name + (dgrad ? ":dgrad" : ":fwd");
  if (c.mode == VCD_MODE_FP32) return launch_gconv_simt<float>(c, st, in, w, g, e, Lin, Lq, Lout, flops, tag.c_str());
  return launch_gconv_simt<bf16>(c, st, in, w, g, e, Lin, Lq, Lout, flops, tag.c_str());
}

// Weight gradient (+ bias gradient) of layer L: in = layer input (length Lin), dout = gradient w.r.t. the layer
// output in the layer's wgr view (length Ld; for ConvTranspose1d the phase-packed tensor).
int run_wgrad(const Ctx& c, cudaStream_t st, const Layer& L, const void* in, const void* dout, int Lin, int Ld, double flops,
              bool tail = false) {
  const ConvGeo& g = L.wgr;
  float* dwp = c.p->d_gscratch + L.dwp;
  static const int dbg_skip = tc_env_int("VCD_DEBUG_SKIP", 0);
  if (dbg_skip & 4) return 0;
  if (c.mode == VCD_MODE_BF16 && L.tc_ok_wgr) {
    const double bytes = 2.0 * c.B * (static_cast<double>(Lin) * g.K + static_cast<double>(Ld) * g.N) + 4.0 * g.taps * g.K * g.N;
    ProfScope ps__(L.cin <= 64 && L.cout <= 64 ? PC_TC_WGRAD_S : PC_TC_WGRAD, flops, bytes, st, (L.name + ":wgrad").c_str());
    // the bias gradient (column sums of dout) is produced by the same kernel
    // deterministic mode: the in-kernel column sums add once per (column tile, split); a phase-packed ConvTranspose
    // gradient folds u column tiles onto one bias element, which would be more than two contributions
    const bool bias_in_kernel = L.dbias >= 0 && !(c.p->deterministic && g.N != L.cout);
    TRY(tc_run_wgrad(c.p, L, in, dout, dwp, bias_in_kernel ? c.p->d_gscratch + L.dbias : nullptr, c.B, Lin, Ld, tail, st, g_launches,
                     g_err, sizeof(g_err)));
    if (bias_in_kernel || L.dbias < 0) return 0;
  } else {
    const long long total = 1LL * c.B * Ld;
    const int blocks_x = g.taps * (g.K / 8) * (g.N / 8);
    long long splits = std::max<long long>(1, std::min<long long>((total + 2047) / 2048,
                                                                    std::max(1, 4 * c.p->num_sms * 8 / blocks_x)));
    if (c.p->deterministic && splits > 2) splits = 2;
    const int rows_per_split = static_cast<int>((total + splits - 1) / splits);
    splits = (total + rows_per_split - 1) / rows_per_split;
    dim3 grid(blocks_x, static_cast<unsigned>(splits));
    ProfScope ps__(PC_FFMA_WGRAD, flops, 0, st, (L.name + ":wgrad").c_str());
    if (c.mode == VCD_MODE_FP32)
      gconv_wgrad_simt_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float*>(in),
          static_cast<const float*>(dout), dwp, g, c.B, Lin, Ld, Ld, rows_per_split);
    else
      gconv_wgrad_simt_kernel<bf16><<<grid, 256, 0, st>>>(static_cast<const bf16*>(in),
          static_cast<const bf16*>(dout), dwp, g, c.B, Lin, Ld, Ld, rows_per_split);
    LAUNCH_CHECK("gconv_wgrad_simt_kernel");
  }
  if (L.dbias >= 0) {
    const int splits = std::max(1, std::min(64, Ld / 2048));
    dim3 grid(g.N / 8, c.B, splits);
    ProfScope ps__(PC_MISC, 0, esize(c.mode) * static_cast<double>(c.B) * g.N * Ld, st);
    if (c.p->deterministic) {   // per-part partial sums (plain stores), then one fixed-order sum
      float* parts = c.p->d_det + (&L - c.p->layers.data()) * vcd_plan::kDetFloats;
      const int nparts = std::max(1, std::min(vcd_plan::kDetParts, Ld / 256));
      if (L.cout > 512) return fail("deterministic bias gradient: %d channels", L.cout);
      dim3 gdet(L.cout / 8, 1, nparts);
      if (c.mode == VCD_MODE_FP32)
        colsum_det_kernel<float><<<gdet, 256, 0, st>>>(static_cast<const float*>(dout), parts, g.N, Ld, c.B, L.cout);
      else
        colsum_det_kernel<bf16><<<gdet, 256, 0, st>>>(static_cast<const bf16*>(dout), parts, g.N, Ld, c.B, L.cout);
      LAUNCH_CHECK("colsum_det_kernel");
      sum_parts_kernel<<<(L.cout + 255) / 256, 256, 0, st>>>(parts, nparts, L.cout, c.p->d_gscratch + L.dbias);
      LAUNCH_CHECK("sum_parts_kernel");
      return 0;
    }
    if (c.mode == VCD_MODE_FP32)
      colsum_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float*>(dout), c.p->d_gscratch + L.dbias, g.N, Ld, 0, L.cout);
    else
      colsum_kernel<bf16><<<grid, 256, 0, st>>>(static_cast<const bf16*>(dout), c.p->d_gscratch + L.dbias, g.N, Ld, 0, L.cout);
    LAUNCH_CHECK("colsum_kernel");
  }
  return 0;
}

Epilogue epi() {
  Epilogue e{};
  e.mask_slope = 1.f; e.scale = 1.f; e.res_inv = 1.f; e.tscale = 1.f; e.act_slope = 1.f;
  return e;
}

int check_common(vcd_plan* p, int mode, int B, int T, void* ws, size_t ws_bytes, bool save) {
  if (!p) return fail("null plan");
  if (mode != VCD_MODE_FP32 && mode != VCD_MODE_BF16) return fail("bad mode %d", mode);
  if (B < 1 || T < 1) return fail("B and T must be >= 1 (got B=%d, T=%d)", B, T);
  if (!p->folded[mode]) return fail("vcd_fold_weights(mode=%d) must be called before forward/backward", mode);
  const size_t need = make_layout(p, mode, B, T, save).total;
  if (!ws || ws_bytes < need) return fail("workspace too small: need %zu bytes, got %zu", need, ws_bytes);
  if (reinterpret_cast<uintptr_t>(ws) % 256) return fail("workspace must be 256-byte aligned");
  return 0;
}

// The legacy default stream cannot be captured into a CUDA graph.  When the caller hands us stream 0 the work
// runs on a plan-owned non-blocking stream instead, ordered after / before the caller's stream with events.
struct StreamHop {
  vcd_plan* p;
  cudaStream_t user, run;
  bool hop;
  StreamHop(vcd_plan* plan, void* stream_) : p(plan), user(static_cast<cudaStream_t>(stream_)) {
    hop = (user == nullptr || user == cudaStreamLegacy || user == cudaStreamPerThread);
    run = hop ? p->own : user;
    if (hop) {
      cudaEventRecord(p->hop_in, user);
      cudaStreamWaitEvent(run, p->hop_in, 0);
    }
  }
  ~StreamHop() {
    if (hop) {
      cudaEventRecord(p->hop_out, run);
      cudaStreamWaitEvent(user, p->hop_out, 0);
    }
  }
};

// CUDA-graph cache: the launch sequence of a call region (kernels, memsets, the fork/join pattern of the
// auxiliary streams, the TMA descriptors baked into kernel parameters) only depends on GraphKey, so it is
// captured once from the SAME enqueue code and replayed afterwards.  Kernels that touch caller-owned tensors
// (x, g, y, dy, dx, dg) stay outside the captured regions because those addresses may change per call.
template <class F>
int run_graphed(vcd_plan* p, const GraphKey& key, cudaStream_t stream, F&& enqueue) {
  static const int enabled = tc_env_int("VCD_GRAPHS", 1);
  if (!enabled || g_prof_on) return enqueue();
  {  // the caller may itself be capturing this stream (its own CUDA graph of the training step): just enqueue
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(stream, &st) != cudaSuccess || st != cudaStreamCaptureStatusNone) return enqueue();
  }
  auto it = p->graphs.find(key);
  if (it != p->graphs.end()) {
    CU_TRY(cudaGraphLaunch(it->second, stream));
    g_launches.fetch_add(p->graph_kernels[key], std::memory_order_relaxed);
    return 0;
  }
  const uint64_t before = g_launches.load();
  CU_TRY(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
  const int rc = enqueue();
  cudaGraph_t graph = nullptr;
  const cudaError_t ce = cudaStreamEndCapture(stream, &graph);
  if (rc != 0 || ce != cudaSuccess || !graph) {
    if (graph) cudaGraphDestroy(graph);
    if (rc != 0) return rc;
    return fail("CUDA graph capture failed: %s", cudaGetErrorString(ce));
  }
  cudaGraphExec_t exec = nullptr;
  const cudaError_t ie = cudaGraphInstantiate(&exec, graph, 0);
  cudaGraphDestroy(graph);
  if (ie != cudaSuccess) return fail("cudaGraphInstantiate failed: %s", cudaGetErrorString(ie));
  if (p->graphs.size() >= 64) {  // bounded cache: drop everything (shapes / workspaces keep changing)
    // replays still in flight keep their executable alive until they finish (cudaGraphExecDestroy defers the release)
    for (auto& kv : p->graphs) cudaGraphExecDestroy(kv.second);
    p->graphs.clear();
    p->graph_kernels.clear();
  }
  p->graphs[key] = exec;
  p->graph_kernels[key] = g_launches.load() - before;
  CU_TRY(cudaGraphLaunch(exec, stream));
  return 0;
}

constexpr float kSlope = 0.1f;          // LRELU_SLOPE, vits/model/modules.py:16
constexpr float kInvSlope = 10.f;       // exact inverse used to recover the residual stream from lrelu(x)
constexpr float kFinalSlope = 0.01f;    // F.leaky_relu default before conv_post (upstream Generator)

}  // namespace

// ---------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------
static int forward_impl(vcd_plan* p, int mode, const float* x, int64_t xs_b, int64_t xs_c, int64_t xs_t,
                        const int64_t* starts, const float* gvec, float* y, void* ws, size_t ws_bytes, int B, int T, int save_,
                        void* stream_);

extern "C" int vcd_forward(vcd_plan* p, int mode, const float* x, int64_t xs_b, int64_t xs_c, int64_t xs_t,
                           const float* gvec, float* y, void* ws, size_t ws_bytes, int B, int T, int save_,
                           void* stream_) {
  return forward_impl(p, mode, x, xs_b, xs_c, xs_t, nullptr, gvec, y, ws, ws_bytes, B, T, save_, stream_);
}

extern "C" int vcd_forward_sliced(vcd_plan* p, int mode, const float* z, int64_t zs_b, int64_t zs_c, int64_t zs_t,
                                  const int64_t* starts_dev, const float* gvec, float* y, void* ws, size_t ws_bytes, int B,
                                  int T, int save_, void* stream_) {
  if (!starts_dev) return fail("vcd_forward_sliced: null starts");
  return forward_impl(p, mode, z, zs_b, zs_c, zs_t, starts_dev, gvec, y, ws, ws_bytes, B, T, save_, stream_);
}

static int forward_impl(vcd_plan* p, int mode, const float* x, int64_t xs_b, int64_t xs_c, int64_t xs_t,
                        const int64_t* starts, const float* gvec, float* y, void* ws, size_t ws_bytes, int B, int T, int save_,
                        void* stream_) {
  const bool save = save_ != 0;
  TRY(check_common(p, mode, B, T, ws, ws_bytes, save));
  if (!x || !y) return fail("vcd_forward: null x or y");
  if (gvec && !p->cfg.gin_channels) return fail("vcd_forward: g given but the plan has gin_channels = 0");
  const WsLayout w = make_layout(p, mode, B, T, save);
  TRY(ensure_pad_tables(p, w, mode, B, T, save));
  StreamHop hop(p, stream_);
  Ctx c{p, mode, B, hop.run, static_cast<char*>(ws), g_prof_on || tc_env_int("VCD_SERIAL", 0) != 0};
  cudaStream_t stream = c.main;
  const bool f32 = mode == VCD_MODE_FP32;
  const int S = static_cast<int>(p->stages.size()), NB = p->cfg.num_kernels;
  const int npairs = p->cfg.resblock == 1 ? 3 : 2;
  const int C0 = p->cfg.upsample_initial_channel, Cin0 = p->cfg.initial_channel;
  auto P = [&](size_t off) { return static_cast<void*>(c.ws + off); };
  auto PF = [&](size_t off) { return reinterpret_cast<float*>(c.ws + off); };

  {  // latent -> blocked layout
    ProfScope ps__(PC_MISC, 0, 0, stream);
    dim3 grid((T + 127) / 128, Cin0 / 8, B);
    const long long* st0 = reinterpret_cast<const long long*>(starts);
    if (f32) ncl_to_blocked_kernel<float><<<grid, 128, 0, stream>>>(x, xs_b, xs_c, xs_t, static_cast<float*>(P(w.xin)), Cin0, T, st0);
    else ncl_to_blocked_kernel<bf16><<<grid, 128, 0, stream>>>(x, xs_b, xs_c, xs_t, static_cast<bf16*>(P(w.xin)), Cin0, T, st0);
    LAUNCH_CHECK("ncl_to_blocked_kernel");
  }
  if (gvec) {
    ProfScope ps__(PC_MISC, 0, 0, stream);
    dim3 grid((C0 + 7) / 8, B);
    cond_fwd_kernel<<<grid, 256, 0, stream>>>(p->h_params[p->p_cond_w], p->h_params[p->p_cond_b], gvec, PF(w.cb), C0,
                                              p->cfg.gin_channels);
    LAUNCH_CHECK("cond_fwd_kernel");
  }
  int Lcur = T;
  auto core = [&]() -> int {
  TRY(launch_pads(p, w.pad_fwd[0], mode, B, T, save, 0, false, c.ws, stream));
  {  // conv_pre (+ cond) -> a[0] = lrelu(., 0.1)
    const Layer& L = p->layers[p->l_pre];
    Epilogue e = epi();
    e.bias = p->h_params[L.p_b];
    e.bias2 = gvec ? PF(w.cb) : nullptr;
    e.out_t = P(w.a[0]);
    e.act_slope = kSlope;
    TRY(run_conv(c, stream, L, false, P(w.xin), T, T, T, e, layer_flops(L, B, T)));
  }
  for (int i = 0; i < S; ++i) {
    const StageDesc& sd = p->stages[i];
    const StageWs& sw = w.st[i];
    const Layer& U = p->layers[sd.up_layer];
    // per-stage device time (VCD_PHASES=1 together with VCD_GRAPHS=0: timing events cannot live inside a captured graph)
    static const bool stage_phases = phases_on() && tc_env_int("VCD_GRAPHS", 1) == 0;
    std::unique_ptr<PhaseScope> stage_ph__;
    if (stage_phases) stage_ph__.reset(new PhaseScope(("forward stage " + std::to_string(i)).c_str(), stream));
    TRY(launch_pads(p, w.pad_fwd[i + 1], mode, B, T, save, i + 1, false, c.ws, stream));
    {  // x = ups[i](lrelu(x)) ; stored as ua = lrelu(x, 0.1)
      Epilogue e = epi();
      e.bias = p->h_params[U.p_b];
      e.out_t = P(sw.ua);
      e.act_slope = kSlope;
      TRY(run_conv(c, stream, U, false, P(w.a[i]), Lcur, Lcur + U.fwd.taps - 1, Lcur * sd.u, e, layer_flops(U, B, Lcur)));
    }
    Lcur *= sd.u;
    // ResBlock branches run concurrently; the running sum over branches chains their last convolutions
    for (int j = 1; j < NB; ++j) c.order(stream, c.branch(j));
    cudaEvent_t prev_last = nullptr;
    static const int dbg_skip_fwd = tc_env_int("VCD_DEBUG_SKIP", 0) & 1;
    const bool pair_off = dbg_skip_fwd != 0;
    for (int j = 0; j < NB; ++j) {
      cudaStream_t sj = c.branch(j);
      const void* xin_t = P(sw.ua);
      for (int q = 0; q < npairs; ++q) {
        const bool last = q == npairs - 1;
        const void* conv_in = xin_t;
        // <= 64-channel stages: both convolutions of a non-final pair in ONE launch (tc_pair.cuh); the final pair joins
        // the running sum over branches and keeps its own epilogues
        if (p->cfg.resblock == 1 && mode == VCD_MODE_BF16 && !pair_off &&
            tc_pair_ok(p->layers[sd.convs[j][q][0]], p->layers[sd.convs[j][q][1]])) {
          const Layer& L1 = p->layers[sd.convs[j][q][0]];
          const Layer& L2 = p->layers[sd.convs[j][q][1]];
          void* out_t = nullptr;
          const float* res2 = nullptr;
          float* out_raw = nullptr;
          float tscale = 1.f, out_slope = kSlope;
          if (!last) {
            out_t = P(sw.xa[j][q]);
          } else {   // the final pair joins the running sum over branches (same chaining as the unfused path below)
            if (j > 0) {
              res2 = PF(w.sum[(j - 1) & 1]);
              if (sj != c.branch(j - 1)) c.wait(sj, prev_last);
            }
            if (j < NB - 1) {
              out_raw = PF(w.sum[j & 1]);
            } else {
              out_t = P(w.a[i + 1]);
              tscale = 1.f / NB;
              out_slope = (i == S - 1) ? kFinalSlope : kSlope;
            }
          }
          const double units = (save ? 3 : 2) + (res2 ? 2 : 0) + (out_raw ? 1 : 0);   // bf16 tensor passes (fp32 = 2)
          const double pbytes = 2.0 * B * static_cast<double>(Lcur) * L1.cin * units + 4.0 * L1.k * L1.cin * L1.cout;
          {
            ProfScope ps__(PC_TC_CONV_S, layer_flops(L1, B, Lcur) + layer_flops(L2, B, Lcur), pbytes, sj, (L1.name + "+" + L2.name + ":fwd").c_str());
            PairCall pc;
            pc.in = xin_t; pc.mid_out = save ? P(sw.ma[j][q]) : nullptr; pc.out = out_t;
            pc.res2 = res2; pc.out_raw = out_raw; pc.tscale = tscale; pc.out_slope = out_slope;
            pc.bias1 = p->h_params[L1.p_b]; pc.bias2 = p->h_params[L2.p_b];
            pc.act_slope = kSlope; pc.res_inv = kInvSlope;
            TRY(tc_run_pair(p, L1, L2, pc, B, Lcur, sj, g_launches, g_err, sizeof(g_err)));
          }
          if (!last) xin_t = P(sw.xa[j][q]);
          else if (!c.serial) prev_last = c.record(sj);
          continue;
        }
        if (p->cfg.resblock == 1) {
          const Layer& L1 = p->layers[sd.convs[j][q][0]];
          Epilogue e = epi();
          e.bias = p->h_params[L1.p_b];
          e.out_t = P(sw.ma[j][q]);
          e.act_slope = kSlope;
          TRY(run_conv(c, sj, L1, false, xin_t, Lcur, Lcur, Lcur, e, layer_flops(L1, B, Lcur)));
          conv_in = P(sw.ma[j][q]);
        }
        const Layer& L2 = p->layers[sd.convs[j][q].back()];
        Epilogue e = epi();
        e.bias = p->h_params[L2.p_b];
        e.res_t = xin_t;  // x recovered from lrelu(x)
        e.res_inv = kInvSlope;
        if (!last) {
          e.out_t = P(sw.xa[j][q]);
          e.act_slope = kSlope;
          xin_t = P(sw.xa[j][q]);
        } else {
          if (j > 0) {
            e.res2 = PF(w.sum[(j - 1) & 1]);
            if (sj != c.branch(j - 1)) c.wait(sj, prev_last);
          }
          if (j < NB - 1) {
            e.out_raw = PF(w.sum[j & 1]);
          } else {  // mean over branches + the next leaky_relu, fused
            e.out_t = P(w.a[i + 1]);
            e.tscale = 1.f / NB;
            e.act_slope = (i == S - 1) ? kFinalSlope : kSlope;
          }
        }
        TRY(run_conv(c, sj, L2, false, conv_in, Lcur, Lcur, Lcur, e, layer_flops(L2, B, Lcur)));
        if (last && !c.serial) prev_last = c.record(sj);
      }
    }
    if (!c.serial && c.branch(NB - 1) != stream) c.wait(stream, prev_last);
  }
  return 0;
  };
  {
    PhaseScope ph__("forward core", stream);
    TRY(run_graphed(p, GraphKey{0, mode, B, T, save ? 1 : 0, gvec ? 1 : 0, 0, ws, p->params_version}, stream, core));
  }
  Lcur = T * p->hop;  // `core` is skipped on a graph replay: do not rely on its side effects
  {  // conv_post + tanh
    const int C = p->stages[S - 1].cout;
    dim3 grid((Lcur + kPostOut - 1) / kPostOut, B);
    const size_t smem = sizeof(float) * C * 8;
    ProfScope ps__(PC_POST, 2.0 * C * 7 * B * Lcur, static_cast<double>(B) * Lcur * (C * esize(mode) + 4), stream);
    if (f32) conv_post_fwd_kernel<float><<<grid, 256, smem, stream>>>(static_cast<const float*>(P(w.a[S])), p->h_params[p->p_post_w], y, C, Lcur);
    else conv_post_fwd_kernel<bf16><<<grid, 256, smem, stream>>>(static_cast<const bf16*>(P(w.a[S])), p->h_params[p->p_post_w], y, C, Lcur);
    LAUNCH_CHECK("conv_post_fwd_kernel");
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------------
static int backward_impl(vcd_plan* p, int mode, const float* dy, const float* y, const float* gvec, float* dx, int64_t T_full,
                         const int64_t* starts, float* dg, float* const* dparams, void* ws, size_t ws_bytes, int B, int T,
                         uint32_t segment_mask, void* stream_);

extern "C" int vcd_backward(vcd_plan* p, int mode, const float* dy, const float* y, const float* gvec, float* dx,
                            float* dg, float* const* dparams, void* ws, size_t ws_bytes, int B, int T,
                            uint32_t segment_mask, void* stream_) {
  return backward_impl(p, mode, dy, y, gvec, dx, T, nullptr, dg, dparams, ws, ws_bytes, B, T, segment_mask, stream_);
}

extern "C" int vcd_backward_sliced(vcd_plan* p, int mode, const float* dy, const float* y, const float* gvec, float* dz,
                                   int64_t T_full, const int64_t* starts_dev, float* dg, float* const* dparams, void* ws,
                                   size_t ws_bytes, int B, int T, uint32_t segment_mask, void* stream_) {
  if (!starts_dev || T_full < T) return fail("vcd_backward_sliced: null starts or T_full < T");
  return backward_impl(p, mode, dy, y, gvec, dz, T_full, starts_dev, dg, dparams, ws, ws_bytes, B, T, segment_mask, stream_);
}

static int backward_impl(vcd_plan* p, int mode, const float* dy, const float* y, const float* gvec, float* dx, int64_t T_full,
                         const int64_t* starts, float* dg, float* const* dparams, void* ws, size_t ws_bytes, int B, int T,
                         uint32_t segment_mask, void* stream_) {
  TRY(check_common(p, mode, B, T, ws, ws_bytes, true));
  if (!dy || !y || !dparams) return fail("vcd_backward: null dy, y or dparams");
  const WsLayout w = make_layout(p, mode, B, T, true);
  TRY(ensure_pad_tables(p, w, mode, B, T, true));
  StreamHop hop(p, stream_);
  Ctx c{p, mode, B, hop.run, static_cast<char*>(ws), g_prof_on || tc_env_int("VCD_SERIAL", 0) != 0};
  cudaStream_t stream = c.main;
  const bool f32 = mode == VCD_MODE_FP32;
  const size_t es = esize(mode);
  const int S = static_cast<int>(p->stages.size()), NB = p->cfg.num_kernels;
  const int npairs = p->cfg.resblock == 1 ? 3 : 2;
  const int C0 = p->cfg.upsample_initial_channel, Cin0 = p->cfg.initial_channel;
  auto P = [&](size_t off) { return static_cast<void*>(c.ws + off); };
  auto PF = [&](size_t off) { return reinterpret_cast<float*>(c.ws + off); };

  const size_t np = p->params.size();
  bool changed = false;
  for (size_t i = 0; i < np; ++i) {
    if (!dparams[i] && !(p->is_weight_g[i] && !p->h_params[i]))
      return fail("vcd_backward: gradient pointer of %s is null", p->params[i].name.c_str());
    changed |= p->h_dparams[i] != dparams[i];
  }
  if (changed) {
    std::copy(dparams, dparams + np, p->h_dparams.begin());
    CU_TRY(cudaMemcpyAsync(p->d_dparams, p->h_dparams.data(), np * sizeof(float*), cudaMemcpyHostToDevice, stream));
  }

  std::vector<int> Ls(S + 1);
  Ls[0] = T;
  for (int i = 0; i < S; ++i) Ls[i + 1] = Ls[i] * p->stages[i].u;

  // Helper streams outside the captured graphs: the gradient-scratch zero-fill of the NEXT requested segment and
  // conv_post's weight gradient run beside the current segment's graph instead of in front of it.
  cudaStream_t zst = c.serial ? stream : p->aux[vcd_plan::kMaxAux - 1];
  cudaStream_t wst = c.serial ? stream : p->aux[vcd_plan::kMaxAux - 2];
  auto zero_scratch = [&](int sg, cudaStream_t st) -> int {
    const SegmentJobs& z = p->segments[sg];
    if (z.scratch_end > z.scratch_begin)
      CU_TRY(cudaMemsetAsync(p->d_gscratch + z.scratch_begin, 0, (z.scratch_end - z.scratch_begin) * sizeof(float), st));
    return 0;
  };
  // Whole-backward mode (every segment requested in one call): ONE captured graph in which the data-gradient chain runs
  // through all stages without waiting for the weight gradients; the weight-gradient kernels of a segment and its
  // weight-norm backward trail behind on the side streams / the unfold stream while the next segment's chain already
  // runs (per-segment calls join them at every segment end: ~0.15 ms of idle chain per segment on base.json).
  const uint32_t all_mask = (1u << (S + 1)) - 1;
  static const int whole_on = tc_env_int("VCD_BWD_WHOLE", 1);
  const bool whole = whole_on && !c.serial && (segment_mask & all_mask) == all_mask;
  cudaStream_t ust = c.serial ? stream : p->aux[vcd_plan::kMaxAux - 3];   // weight-norm backward of trailing segments
  std::vector<cudaEvent_t> join_ev(S + 1, nullptr);   // whole mode: weight gradients of segment s complete
  bool post_forked = false;
  auto pre_part = [&](int seg) -> int {
    const SegmentJobs& sj = p->segments[seg];
    if (seg == 0) {  // conv_post + tanh backward -> gradient w.r.t. the last stage output
      const int C = p->stages[S - 1].cout, L = Ls[S];
      const int splits = std::max(1, std::min(64, L / 2048));
      dim3 gw(C / 8, B, splits);
      float* post_parts = nullptr;
      if (p->deterministic) {   // per-part partial sums, then one fixed-order sum
        gw = dim3(C / 8, 1, std::max(1, std::min(static_cast<int>(vcd_plan::kDetParts), L / 256)));
        post_parts = p->d_det + p->layers.size() * vcd_plan::kDetFloats;
        if (C * 7 > 512) return fail("deterministic conv_post gradient: %d channels", C);
      }
      dim3 gd((L + 127) / 128, C / 8, B);
      const float* wpost = p->h_params[p->p_post_w];
      const int slot = (S - 1) % 3;
      ProfScope ps__(PC_POST, 4.0 * C * 7 * B * L, static_cast<double>(B) * L * (C * 3 * es + 16), stream);
      // weight gradient (+ its copy into the caller's tensor) beside the data-gradient chain
      c.order(stream, wst);
      post_forked = wst != stream;
      if (f32) {
        conv_post_wgrad_kernel<float><<<gw, 256, 0, wst>>>(dy, y, static_cast<const float*>(P(w.a[S])), p->d_gscratch + p->post_dw, C, L, B, post_parts);
        LAUNCH_CHECK("conv_post_wgrad_kernel");
        if (post_parts) {
          sum_parts_kernel<<<(C * 7 + 255) / 256, 256, 0, wst>>>(post_parts, static_cast<int>(gw.z), C * 7, p->d_gscratch + p->post_dw);
          LAUNCH_CHECK("sum_parts_kernel");
        }
        conv_post_dgrad_kernel<float><<<gd, 128, 0, stream>>>(dy, y, wpost, static_cast<const float*>(P(w.a[S])), kFinalSlope, 1.f / NB,
                                                               nullptr, static_cast<float*>(P(w.Gi[slot])), C, L);
      } else {
        conv_post_wgrad_kernel<bf16><<<gw, 256, 0, wst>>>(dy, y, static_cast<const bf16*>(P(w.a[S])), p->d_gscratch + p->post_dw, C, L, B, post_parts);
        LAUNCH_CHECK("conv_post_wgrad_kernel");
        if (post_parts) {
          sum_parts_kernel<<<(C * 7 + 255) / 256, 256, 0, wst>>>(post_parts, static_cast<int>(gw.z), C * 7, p->d_gscratch + p->post_dw);
          LAUNCH_CHECK("sum_parts_kernel");
        }
        conv_post_dgrad_kernel<bf16><<<gd, 128, 0, stream>>>(dy, y, wpost, static_cast<const bf16*>(P(w.a[S])), kFinalSlope, 1.f / NB,
                                                              nullptr, static_cast<bf16*>(P(w.Gi[slot])), C, L);
      }
      LAUNCH_CHECK("conv_post_dgrad_kernel");
      if (sj.fast_nblocks > 0 && sj.fast_lead_blocks > 0) {
        wn_unfold_fast_kernel<<<sj.fast_lead_blocks, kFoldThreads, sj.fast_smem, wst>>>(sj.d_fast, sj.fast_lead_jobs, p->d_params, p->d_dparams,
                                                                              p->d_norms, p->d_gscratch, 0, p->grad_scale);
        LAUNCH_CHECK("wn_unfold_fast_kernel");
      } else if (sj.lead_blocks > 0) {
        wn_unfold_kernel<<<sj.lead_blocks, 256, 0, wst>>>(sj.d_jobs, sj.lead_jobs, p->d_params, p->d_dparams, p->d_norms, p->d_gscratch, 0,
                                                          p->grad_scale);
        LAUNCH_CHECK("wn_unfold_kernel");
      }
    }
    return 0;
  };
  auto core = [&](int seg) -> int {
    const SegmentJobs& sj = p->segments[seg];
    int side_rr = 0;
    bool side_used[4] = {false, false, false, false};
    // weight-gradient kernels run on side streams: they only need their two operands, never feed the
    // data-gradient chain, and are joined before the segment's weight-norm backward
    auto side_after = [&](cudaEvent_t ready) {
      const int k = side_rr++ & 3;
      cudaStream_t s = c.side(k);
      if (!c.serial) { c.wait(s, ready); side_used[k] = true; }
      return s;
    };
    auto wg_request = [&](const Layer& Lx) {
      WgFuse r;
      if (c.mode == VCD_MODE_BF16 && Lx.tc_ok_wgr && Lx.tc_ok_dgr) {
        r.dwp = p->d_gscratch + Lx.dwp;
        r.dbias = Lx.dbias >= 0 ? p->d_gscratch + Lx.dbias : nullptr;
      }
      return r;
    };
    // whole mode: this segment reuses the gradient workspaces of segment seg-2, whose weight gradients may still trail
    if (whole && seg >= 2 && join_ev[seg - 2]) c.wait(stream, join_ev[seg - 2]);
    if (seg < S) {
      const int i = S - 1 - seg;
      const StageDesc& sd = p->stages[i];
      const StageWs& sw = w.st[i];
      const int L = Ls[i + 1], Lprev = Ls[i];
      const Layer& U = p->layers[sd.up_layer];
      const int Lz = Lprev + U.fwd.taps - 1;  // rows of the phase-packed gradient
      const void* Gi = P(w.Gi[i % 3]);
      // unwritten edge slots (and the pad rows) of the phase-packed tensor must read as zero
      CU_TRY(cudaMemsetAsync(P(w.duz[i & 1]), 0, es * blk_elems(B, U.wgr.N, Lz), stream));
      TRY(launch_pads(p, w.pad_bwd[seg], mode, B, T, true, seg, true, c.ws, stream));
      cudaEvent_t ev0 = c.serial ? nullptr : c.record(stream);
      for (int j = 1; j < NB; ++j) if (!c.serial) c.wait(c.branch(j), ev0);
      cudaEvent_t prev_final = nullptr;
      for (int j = 0; j < NB; ++j) {
        cudaStream_t sjs = c.branch(j);
        const void* Gt_cur = Gi;
        cudaEvent_t ev_cur = ev0;
        for (int q = npairs - 1; q >= 0; --q) {
          const void* in_first = q == 0 ? P(sw.ua) : P(sw.xa[j][q - 1]);  // input of the pair's first conv
          const void* d_first = Gt_cur;                                   // gradient w.r.t. that conv's output
          cudaEvent_t ev_first = ev_cur;
          // <= 64-channel stages, pairs behind the first one: both data gradients in ONE launch (tc_pair.cuh, BWD); the
          // first pair of a branch joins the running sum over branches and keeps its own epilogues
          static const int dbg_skip_dgrad = tc_env_int("VCD_DEBUG_SKIP", 0) & 2;
          if (p->cfg.resblock == 1 && q > 0 && c.mode == VCD_MODE_BF16 && !dbg_skip_dgrad &&
              tc_pair_ok(p->layers[sd.convs[j][q][0]], p->layers[sd.convs[j][q][1]], true)) {
            const Layer& L1 = p->layers[sd.convs[j][q][0]];
            const Layer& L2 = p->layers[sd.convs[j][q][1]];
            TRY(run_wgrad(c, side_after(ev_cur), L2, P(sw.ma[j][q]), Gt_cur, L, L, layer_flops(L2, B, L)));
            {
              const double pbytes = 2.0 * B * static_cast<double>(L) * L1.cin * 5 + 4.0 * L1.k * L1.cin * L1.cout;   // G, two masks in; dm, G' out
              ProfScope ps__(PC_TC_CONV_S, layer_flops(L1, B, L) + layer_flops(L2, B, L), pbytes, sjs, (L2.name + "+" + L1.name + ":dgrad").c_str());
              PairCall pc;
              pc.bwd = true;
              pc.in = Gt_cur; pc.mask1 = P(sw.ma[j][q]); pc.mask2 = in_first; pc.mask_slope = kSlope;
              pc.mid_out = P(w.dm[i & 1][j][q]); pc.out = P(w.Gt[i & 1][j][q]);
              TRY(tc_run_pair(p, L1, L2, pc, B, L, sjs, g_launches, g_err, sizeof(g_err)));
            }
            cudaEvent_t ev_pair = c.serial ? nullptr : c.record(sjs);
            TRY(run_wgrad(c, side_after(ev_pair), L1, in_first, P(w.dm[i & 1][j][q]), L, L, layer_flops(L1, B, L)));
            Gt_cur = P(w.Gt[i & 1][j][q]);
            ev_cur = ev_pair;
            continue;
          }
          if (p->cfg.resblock == 1) {
            const Layer& L2 = p->layers[sd.convs[j][q][1]];
            Epilogue e = epi();
            e.mask = P(sw.ma[j][q]);
            e.mask_slope = kSlope;
            e.out_t = P(w.dm[i & 1][j][q]);
            // the weight gradient rides in the data-gradient launch when it can (its `in` operand is the mask operand)
            WgFuse wg2 = wg_request(L2);
            TRY(run_conv(c, sjs, L2, true, Gt_cur, L, L, L, e, layer_flops(L2, B, L), wg2.dwp ? &wg2 : nullptr));
            if (!wg2.fused) TRY(run_wgrad(c, side_after(ev_cur), L2, P(sw.ma[j][q]), Gt_cur, L, L, layer_flops(L2, B, L)));
            d_first = P(w.dm[i & 1][j][q]);
            ev_first = c.serial ? nullptr : c.record(sjs);
          }
          const Layer& L1 = p->layers[sd.convs[j][q][0]];
          // (the first conv of a branch joins the running sum over branches: its epilogue is not a lean one, no fusion)
          WgFuse wg1 = q > 0 ? wg_request(L1) : WgFuse{};
          Epilogue e = epi();
          e.mask = in_first;
          e.mask_slope = kSlope;
          e.res_t = Gt_cur;  // identity path of `x = xt + x`
          if (q > 0) {
            e.out_t = P(w.Gt[i & 1][j][q]);
            Gt_cur = P(w.Gt[i & 1][j][q]);
          } else {
            if (j > 0) {
              e.res2 = PF(w.sum[(j - 1) & 1]);
              if (!c.serial && sjs != c.branch(j - 1)) c.wait(sjs, prev_final);
            }
            if (j < NB - 1) {
              e.out_raw = PF(w.sum[j & 1]);
            } else {  // sum over branches, written phase-packed for the upsample conv's backward GEMMs
              e.out_t = P(w.duz[i & 1]);
              e.zu = sd.u; e.zp = U.pad; e.zLq = Lz;
            }
          }
          TRY(run_conv(c, sjs, L1, true, d_first, L, L, L, e, layer_flops(L1, B, L), wg1.dwp ? &wg1 : nullptr));
          if (!wg1.fused)
            TRY(run_wgrad(c, side_after(ev_first), L1, in_first, d_first, L, L, layer_flops(L1, B, L), q == 0 && j == NB - 1));
          if (!c.serial) {
            if (q > 0) ev_cur = c.record(sjs);
            else prev_final = c.record(sjs);
          }
        }
      }
      if (!c.serial && c.branch(NB - 1) != stream) c.wait(stream, prev_final);
      // upsample conv: weight/bias gradients (side stream) and data gradient, fused with the lrelu mask and the
      // 1/NB of the previous stage's branch mean
      cudaEvent_t ev_du = c.serial ? nullptr : c.record(stream);
      TRY(run_wgrad(c, side_after(ev_du), U, P(w.a[i]), P(w.duz[i & 1]), Lprev, Lz, layer_flops(U, B, Lprev), true));
      Epilogue e = epi();
      e.mask = P(w.a[i]);
      e.mask_slope = kSlope;
      if (i > 0) {
        e.scale = 1.f / NB;
        e.out_t = P(w.Gi[(i - 1) % 3]);
      } else {
        e.out_t = P(w.d0);
      }
      TRY(run_conv(c, stream, U, true, P(w.duz[i & 1]), Lz, Lprev, Lprev, e, layer_flops(U, B, Lprev)));
    } else {  // conv_pre: weight gradient, per-batch column sums (bias / cond gradients), data gradient
      const Layer& L = p->layers[p->l_pre];
      cudaEvent_t ev0 = c.serial ? nullptr : c.record(stream);
      TRY(run_wgrad(c, side_after(ev0), L, P(w.xin), P(w.d0), T, T, layer_flops(L, B, T)));
      CU_TRY(cudaMemsetAsync(PF(w.dcb), 0, sizeof(float) * B * C0, stream));
      {
        ProfScope ps__(PC_MISC, 0, 0, stream);
        const int splits = std::max(1, std::min(p->deterministic ? 2 : 64, T / 2048));
        dim3 grid(C0 / 8, B, splits);
        if (f32) colsum_kernel<float><<<grid, 256, 0, stream>>>(static_cast<const float*>(P(w.d0)), PF(w.dcb), C0, T, 1, C0);
        else colsum_kernel<bf16><<<grid, 256, 0, stream>>>(static_cast<const bf16*>(P(w.d0)), PF(w.dcb), C0, T, 1, C0);
        LAUNCH_CHECK("colsum_kernel");
      }
      if (dx) {
        Epilogue e = epi();
        e.out_raw = PF(w.dxb);
        TRY(run_conv(c, stream, L, true, P(w.d0), T, T, T, e, layer_flops(L, B, T)));
      }
    }
    // join the side streams, then the weight-norm backward of this segment's parameters (whole mode: on the unfold
    // stream, behind the chain)
    cudaStream_t fst = whole ? ust : stream;
    for (int k = 0; k < 4; ++k)
      if (side_used[k]) c.order(c.side(k), fst);
    if (whole) join_ev[seg] = c.record(fst);
    if (sj.fast_nblocks > sj.fast_lead_blocks) {  // (the jobs ahead of the first layer ran with conv_post's weight gradient)
      ProfScope ps__(PC_FOLD, 0, 12.0 * (sj.scratch_end - sj.scratch_begin), fst);   // dWp read + v read + gradient write
      wn_unfold_fast_kernel<<<sj.fast_nblocks - sj.fast_lead_blocks, kFoldThreads, sj.fast_smem, fst>>>(
          sj.d_fast + sj.fast_lead_jobs, sj.fast_njobs - sj.fast_lead_jobs, p->d_params, p->d_dparams, p->d_norms, p->d_gscratch,
          sj.fast_lead_blocks, p->grad_scale);
      LAUNCH_CHECK("wn_unfold_fast_kernel");
    } else if (sj.fast_nblocks == 0 && sj.nblocks > sj.lead_blocks) {
      ProfScope ps__(PC_FOLD, 0, 12.0 * (sj.scratch_end - sj.scratch_begin), fst);   // dWp read + v read + gradient write
      wn_unfold_kernel<<<sj.nblocks - sj.lead_blocks, 256, 0, fst>>>(sj.d_jobs + sj.lead_jobs, sj.njobs - sj.lead_jobs, p->d_params,
                                                                       p->d_dparams, p->d_norms, p->d_gscratch, sj.lead_blocks, p->grad_scale);
      LAUNCH_CHECK("wn_unfold_kernel");
    }
    if (whole) {   // gradients of this segment are final: visible to streams outside the captured graph
      cudaStreamCaptureStatus cst = cudaStreamCaptureStatusNone;
      cudaStreamIsCapturing(fst, &cst);
      CU_TRY(cudaEventRecordWithFlags(p->seg_done[seg], fst, cst == cudaStreamCaptureStatusActive ? cudaEventRecordExternal : cudaEventRecordDefault));
    }
    return 0;
  };
  auto post_part = [&](int seg) -> int {
    if (seg == S) {  // kernels that touch caller-owned tensors: cond / conv_pre.bias gradients, dg, dx
      const Layer& L = p->layers[p->l_pre];
      {
        ProfScope ps__(PC_MISC, 0, 0, stream);
        const int G = p->cfg.gin_channels;
        const bool have_g = gvec != nullptr && G > 0;
        long long n = C0;
        if (have_g) n = std::max<long long>(n, 1LL * C0 * G);
        cond_bwd_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(
            PF(w.dcb), have_g ? p->h_params[p->p_cond_w] : nullptr, gvec, p->h_dparams[L.p_b],
            have_g ? p->h_dparams[p->p_cond_b] : nullptr, have_g ? p->h_dparams[p->p_cond_w] : nullptr,
            B, C0, std::max(G, 1), p->grad_scale);
        LAUNCH_CHECK("cond_bwd_kernel");
        if (have_g && dg) {
          cond_dg_kernel<<<dim3((G + 31) / 32, B), 256, 0, stream>>>(PF(w.dcb), p->h_params[p->p_cond_w], dg, C0, G);
          LAUNCH_CHECK("cond_dg_kernel");
        }
        if (!have_g && G > 0) {
          CU_TRY(cudaMemsetAsync(p->h_dparams[p->p_cond_w], 0, sizeof(float) * p->params[p->p_cond_w].numel, stream));
          CU_TRY(cudaMemsetAsync(p->h_dparams[p->p_cond_b], 0, sizeof(float) * p->params[p->p_cond_b].numel, stream));
          if (dg) CU_TRY(cudaMemsetAsync(dg, 0, sizeof(float) * B * G, stream));
        }
      }
      if (dx) {
        ProfScope ps__(PC_MISC, 0, 0, stream);
        dim3 grid((T + 127) / 128, Cin0 / 8, B);
        // sliced call: the gradient w.r.t. the full-length latent is zero outside each item's segment
        if (starts) CU_TRY(cudaMemsetAsync(dx, 0, sizeof(float) * static_cast<size_t>(B) * Cin0 * T_full, stream));
        blocked_to_ncl_kernel<<<grid, 128, 0, stream>>>(PF(w.dxb), dx, Cin0, T, T_full, reinterpret_cast<const long long*>(starts));
        LAUNCH_CHECK("blocked_to_ncl_kernel");
      }
    }
    return 0;
  };
  uint32_t scale_bits;
  memcpy(&scale_bits, &p->grad_scale, sizeof(scale_bits));   // baked into the captured unfold launches
  const uint64_t gver = p->params_version ^ (static_cast<uint64_t>(scale_bits) << 32) ^ (p->deterministic ? 1ull << 31 : 0ull);
  if (whole) {
    for (int seg = 0; seg <= S; ++seg) TRY(zero_scratch(seg, stream));
    TRY(pre_part(0));
    {
      PhaseScope ph__("backward (all segments)", stream);
      TRY(run_graphed(p, GraphKey{1 + 64, mode, B, T, 1, gvec ? 1 : 0, dx ? 1 : 0, ws, gver}, stream, [&]() -> int {
        for (int seg = 0; seg <= S; ++seg) TRY(core(seg));
        c.order(ust, stream);   // the last trailing weight-norm backward
        return 0;
      }));
    }
    if (post_forked) {
      CU_TRY(cudaEventRecord(p->post_done, wst));
      CU_TRY(cudaStreamWaitEvent(stream, p->post_done, 0));
    } else {
      CU_TRY(cudaEventRecord(p->post_done, stream));
    }
    TRY(post_part(S));
    CU_TRY(cudaEventRecord(p->seg_done[S], stream));   // the last segment's gradients include the kernels outside the graph
    p->seg_events_valid = true;
    return 0;
  }
  p->seg_events_valid = false;
  bool first_seg = true;
  for (int seg = 0; seg <= S; ++seg) {
    if (!(segment_mask & (1u << seg))) continue;
    if (first_seg) TRY(zero_scratch(seg, stream));   // later segments: zeroed beside the previous segment (below)
    first_seg = false;
    int next_seg = -1;
    for (int k = seg + 1; k <= S; ++k)
      if (segment_mask & (1u << k)) { next_seg = k; break; }
    if (next_seg >= 0) {
      c.order(stream, zst);
      TRY(zero_scratch(next_seg, zst));
    }
    post_forked = false;
    TRY(pre_part(seg));
    {
      PhaseScope ph__(("backward segment " + std::to_string(seg)).c_str(), stream);
      TRY(run_graphed(p, GraphKey{1 + seg, mode, B, T, 1, gvec ? 1 : 0, dx ? 1 : 0, ws, gver}, stream, [&]() -> int { return core(seg); }));
    }
    if (post_forked) c.order(wst, stream);
    if (next_seg >= 0) c.order(zst, stream);
    TRY(post_part(seg));
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// host-buffer entry point (infer.py path)
// ---------------------------------------------------------------------------------------------------
extern "C" size_t vcd_host_call_extra_bytes(const vcd_plan* p, int B, int T) {
  if (!p || B < 1 || T < 1) return 0;
  size_t n = align_up(sizeof(float) * B * p->cfg.initial_channel * T, 256);
  n += align_up(sizeof(float) * B * std::max(p->cfg.gin_channels, 1), 256);
  n += align_up(sizeof(float) * static_cast<size_t>(B) * T * p->hop, 256);
  return n;
}

extern "C" int vcd_synthesize_host(vcd_plan* p, int mode, const float* x_host, const float* g_host, float* y_host,
                                   void* ws, size_t ws_bytes, int B, int T, void* stream_) {
  if (!p || !x_host || !y_host) return fail("vcd_synthesize_host: null argument");
  if (B < 1 || T < 1) return fail("B and T must be >= 1");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const size_t core = align_up(vcd_workspace_bytes(p, mode, B, T, 0), 256);
  const size_t extra = vcd_host_call_extra_bytes(p, B, T);
  if (!ws || ws_bytes < core + extra) return fail("workspace too small: need %zu bytes, got %zu", core + extra, ws_bytes);
  char* base = static_cast<char*>(ws) + core;
  const int Cin = p->cfg.initial_channel, G = p->cfg.gin_channels;
  float* xd = reinterpret_cast<float*>(base);
  float* gd = reinterpret_cast<float*>(base + align_up(sizeof(float) * B * Cin * T, 256));
  float* yd = reinterpret_cast<float*>(reinterpret_cast<char*>(gd) + align_up(sizeof(float) * B * std::max(G, 1), 256));
  CU_TRY(cudaMemcpyAsync(xd, x_host, sizeof(float) * B * Cin * T, cudaMemcpyHostToDevice, stream));
  if (g_host) {
    if (!G) return fail("g_host given but the plan has gin_channels = 0");
    CU_TRY(cudaMemcpyAsync(gd, g_host, sizeof(float) * B * G, cudaMemcpyHostToDevice, stream));
  }
  TRY(vcd_forward(p, mode, xd, static_cast<int64_t>(Cin) * T, T, 1, g_host ? gd : nullptr, yd, ws, core, B, T, 0, stream_));
  CU_TRY(cudaMemcpyAsync(y_host, yd, sizeof(float) * static_cast<size_t>(B) * T * p->hop, cudaMemcpyDeviceToHost, stream));
  CU_TRY(cudaStreamSynchronize(stream));
  return 0;
}

// Debug: copy the 64 %globaltimer stamps written by a VCD_KTRACE-selected kernel (see tc_kernels.cuh).
extern "C" int vcd_debug_read_trace(vcd_plan* p, unsigned long long* out64) {
  if (!p || !p->d_trace) return fail("tracing is off (set VCD_KTRACE=<layer>:<fwd|dgrad>)");
  CU_TRY(cudaDeviceSynchronize());
  CU_TRY(cudaMemcpy(out64, p->d_trace, 64 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  return 0;
}

// Debug / test hook: location of one internal tensor inside the caller-owned workspace (blocked channels-last, row
// padded: [B][C/8][pad_l + L + pad_r][8] elements of 2 (bf16 mode) or 4 (fp32 mode) bytes).  Names:
//   xin | a<i> | ua<i> | ma<i>.<j>.<q> | xa<i>.<j>.<q>            (forward; i = stage, j = ResBlock branch, q = pair)
//   Gi<i> | Gt<i>.<j>.<q> | dm<i>.<j>.<q> | duz<i> | d0            (backward; buffers shared between stages: valid
//                                                                   right after the segment of stage i has run)
// The parity tests use it to check every kernel launch against the oracle on the kernel's own stored operands.
extern "C" int vcd_debug_ws_tensor(const vcd_plan* p, int mode, int B, int T, int save, const char* name, size_t* offset,
                                   int* C, int* L, int* pad_l, int* pad_r) {
  if (!p || !name || !offset || !C || !L) return fail("vcd_debug_ws_tensor: null argument");
  const WsLayout w = make_layout(p, mode, B, T, save != 0);
  const int S = static_cast<int>(p->stages.size()), NB = p->cfg.num_kernels;
  const int npairs = p->cfg.resblock == 1 ? 3 : 2;
  std::vector<int> Ls(S + 1);
  Ls[0] = T;
  for (int i = 0; i < S; ++i) Ls[i + 1] = Ls[i] * p->stages[i].u;
  if (pad_l) *pad_l = kPadL;
  if (pad_r) *pad_r = kPadR;
  int i = -1, j = -1, q = -1;
  auto stage_ok = [&]() { return i >= 0 && i < S; };
  auto bq_ok = [&]() { return stage_ok() && j >= 0 && j < NB && q >= 0 && q < npairs; };
  size_t off = 0;
  bool found = false;
  if (!strcmp(name, "xin")) { off = w.xin; *C = p->cfg.initial_channel; *L = T; found = true; }
  else if (!strcmp(name, "d0")) { off = w.d0; *C = p->cfg.upsample_initial_channel; *L = T; found = save != 0; }
  else if (sscanf(name, "ma%d.%d.%d", &i, &j, &q) == 3) { if (bq_ok() && p->cfg.resblock == 1) { off = w.st[i].ma[j][q]; found = true; } }
  else if (sscanf(name, "xa%d.%d.%d", &i, &j, &q) == 3) { if (bq_ok() && q < npairs - 1) { off = w.st[i].xa[j][q]; found = true; } }
  else if (sscanf(name, "Gt%d.%d.%d", &i, &j, &q) == 3) { if (bq_ok() && q > 0 && save) { off = w.Gt[i & 1][j][q]; found = true; } }
  else if (sscanf(name, "dm%d.%d.%d", &i, &j, &q) == 3) { if (bq_ok() && save && p->cfg.resblock == 1) { off = w.dm[i & 1][j][q]; found = true; } }
  else if (sscanf(name, "ua%d", &i) == 1) { if (stage_ok()) { off = w.st[i].ua; found = true; } }
  else if (sscanf(name, "Gi%d", &i) == 1) { if (stage_ok() && save) { off = w.Gi[i % 3]; found = true; } }
  else if (sscanf(name, "duz%d", &i) == 1) {
    if (stage_ok() && save) {
      const Layer& U = p->layers[p->stages[i].up_layer];
      *offset = w.duz[i & 1]; *C = U.wgr.N; *L = Ls[i] + U.fwd.taps - 1;
      return 0;
    }
  }
  else if (sscanf(name, "a%d", &i) == 1) {
    if (i >= 0 && i <= S) {
      *offset = w.a[i]; *C = i == 0 ? p->cfg.upsample_initial_channel : p->stages[i - 1].cout; *L = Ls[i];
      return 0;
    }
  }
  if (!found) return fail("vcd_debug_ws_tensor: unknown or unavailable tensor '%s'", name);
  *offset = off;
  if (stage_ok()) { *C = p->stages[i].cout; *L = Ls[i + 1]; }
  return 0;
}

// Debug / test hook: choose, for plans created AFTER this call, which directions of bf16 mode run on the tcgen05
// kernels (1) or on the FFMA kernels with the same bf16 operands (0); a negative value leaves a flag unchanged.
extern "C" int vcd_set_gradient_scale(vcd_plan* p, float scale) {
  if (!p) return fail("vcd_set_gradient_scale: null plan");
  p->grad_scale = scale;
  return 0;
}

extern "C" int vcd_set_deterministic(vcd_plan* p, int on) {
  if (!p) return fail("vcd_set_deterministic: null plan");
  if (on && !p->d_turn) {
    const size_t n = p->layers.size() * vcd_plan::kTurnInts * sizeof(int);
    CU_TRY(cudaMalloc(&p->d_turn, n));
    CU_TRY(cudaMemset(p->d_turn, 0, n));
    CU_TRY(cudaMalloc(&p->d_det, (p->layers.size() + 1) * vcd_plan::kDetFloats * sizeof(float)));
  }
  p->deterministic = on != 0;
  return 0;
}

// Whole-backward mode (vcd_backward with every segment requested, outside profiling): the gradients of segment `seg`
// become final while later segments still run.  Makes `stream` wait for them (the caller then issues that segment's
// gradient all-reduce on it).  vcd_segment_events_valid: 1 if the last vcd_backward recorded these events.
extern "C" int vcd_segment_events_valid(const vcd_plan* p) { return p && p->seg_events_valid ? 1 : 0; }
extern "C" int vcd_stream_wait_segment(vcd_plan* p, int seg, void* stream_) {
  if (!p || seg < 0 || seg >= static_cast<int>(p->segments.size())) return fail("vcd_stream_wait_segment: bad plan or segment");
  if (!p->seg_events_valid) return fail("vcd_stream_wait_segment: the last vcd_backward did not run in whole-backward mode");
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  CU_TRY(cudaStreamWaitEvent(st, p->seg_done[seg], 0));
  if (seg == 0) CU_TRY(cudaStreamWaitEvent(st, p->post_done, 0));
  return 0;
}

extern "C" int vcd_debug_tc_paths(int fwd, int dgrad, int wgrad) {
  int* f = tc_path_flags();
  if (fwd >= 0) f[0] = fwd != 0;
  if (dgrad >= 0) f[1] = dgrad != 0;
  if (wgrad >= 0) f[2] = wgrad != 0;
  return 0;
}

extern "C" const char* vcd_layer_path(const vcd_plan* p, int mode, int index) {
  if (!p || index < 0 || index >= static_cast<int>(p->layers.size())) return nullptr;
  static thread_local char buf[256];
  const Layer& L = p->layers[index];
  const bool bf = mode == VCD_MODE_BF16;
  snprintf(buf, sizeof(buf), "%s fwd=%s dgrad=%s wgrad=%s", L.name.c_str(),
           bf && L.tc_ok_fwd ? "tcgen05-bf16" : (bf ? "ffma-bf16io" : "ffma-fp32"),
           bf && L.tc_ok_dgr ? "tcgen05-bf16" : (bf ? "ffma-bf16io" : "ffma-fp32"),
           bf && L.tc_ok_wgr ? "tcgen05-bf16" : (bf ? "ffma-bf16io" : "ffma-fp32"));
  return buf;
}

// ---------------------------------------------------------------------------------------------------
// Mel / STFT loss tail (include/vcd.h "mel loss tail")
// ---------------------------------------------------------------------------------------------------
#include "mel_loss.cuh"

struct vcd_mel_plan {
  vcd_mel_config cfg;
  int device = 0;
  int n_bins = 0, ld_bins = 0, ld_spec = 0;   // n_fft/2 + 1; padded leading dimensions of mag / melW and of S
  float* d_basis = nullptr;                   // [2 * n_bins][n_fft] window-folded DFT basis
  float* d_melw = nullptr;                    // [n_mel][ld_bins] filterbank (pad columns zero)
  // fast path (n_fft a power of two): one CTA per frame, shared-memory FFT + banded filterbank
  bool fft_ok = false, use_gemm = false;
  int log2n = 0;
  size_t fft_smem = 0;
  float* d_window = nullptr;
  float2* d_tw = nullptr;
  int* d_tab = nullptr;                       // f_lo, f_cnt, f_off [n_mel] | b_lo, b_cnt, b_off [n_bins]
  float* d_vals = nullptr;                    // f_val | b_val
  size_t f_val_n = 0;
};

namespace {
inline size_t mel_align(size_t v) { return (v + 255) & ~static_cast<size_t>(255); }
struct MelWs {
  size_t S, mag, dM, dframe, partials, total;
  int frames, rows, n_partials;
};
// frames per item of torch.stft(center=False) on the reflect-padded signal (mel_processing.py:90-93)
inline int mel_frames(const vcd_mel_config& c, int T) {
  const int pad = (c.n_fft - c.hop) / 2;
  const int padded = T + 2 * pad;
  return padded < c.n_fft ? 0 : (padded - c.n_fft) / c.hop + 1;
}
inline MelWs mel_ws_layout(const vcd_mel_plan* p, int B, int T) {
  MelWs w{};
  w.frames = mel_frames(p->cfg, T);
  w.rows = B * w.frames;
  w.n_partials = ((p->cfg.n_mel + 63) / 64) * ((w.rows + 63) / 64);
  if (w.n_partials < w.rows) w.n_partials = w.rows;      // fast path: one partial per frame
  size_t o = 0;
  w.S = o; o += mel_align(sizeof(float) * static_cast<size_t>(w.rows) * p->ld_spec);
  w.mag = o; o += mel_align(sizeof(float) * static_cast<size_t>(w.rows) * p->ld_bins);
  w.dM = o; o += mel_align(sizeof(float) * static_cast<size_t>(w.rows) * p->cfg.n_mel);
  w.dframe = o; o += mel_align(sizeof(float) * static_cast<size_t>(w.rows) * p->cfg.n_fft);
  w.partials = o; o += mel_align(sizeof(float) * static_cast<size_t>(w.n_partials));
  w.total = o;
  return w;
}
}  // namespace

extern "C" int vcd_mel_plan_create(const vcd_mel_config* cfg, const float* mel_basis_host, vcd_mel_plan** out) {
  if (!cfg || !mel_basis_host || !out) return fail("vcd_mel_plan_create: null argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail("vcd_mel_plan_create: no CUDA device available (this library has no CPU fallback)");
  if (cfg->n_fft < 16 || (cfg->n_fft & 3)) return fail("n_fft (%d) must be a multiple of 4 and >= 16", cfg->n_fft);
  if (cfg->win < 1 || cfg->win > cfg->n_fft) return fail("win (%d) must be in [1, n_fft]", cfg->win);
  if (cfg->hop < 1 || cfg->hop > cfg->n_fft || ((cfg->n_fft - cfg->hop) & 1))
    return fail("hop (%d) must be in [1, n_fft] with n_fft - hop even (the reflect pad is (n_fft - hop) / 2 per side)", cfg->hop);
  if (cfg->n_mel < 1 || (cfg->n_mel & 3)) return fail("n_mel (%d) must be a positive multiple of 4", cfg->n_mel);
  std::unique_ptr<vcd_mel_plan> p(new vcd_mel_plan());
  p->cfg = *cfg;
  CU_TRY(cudaGetDevice(&p->device));
  p->n_bins = cfg->n_fft / 2 + 1;
  p->ld_bins = (p->n_bins + 3) & ~3;
  p->ld_spec = (2 * p->n_bins + 3) & ~3;
  const size_t nb = static_cast<size_t>(2) * p->n_bins * cfg->n_fft, nm = static_cast<size_t>(cfg->n_mel) * p->ld_bins;
  CU_TRY(cudaMalloc(&p->d_basis, nb * sizeof(float)));
  if (cudaMalloc(&p->d_melw, nm * sizeof(float)) != cudaSuccess) {
    cudaFree(p->d_basis);
    return fail("vcd_mel_plan_create: cudaMalloc of the filterbank failed");
  }
  std::vector<float> padded(nm, 0.f);
  for (int m = 0; m < cfg->n_mel; ++m)
    memcpy(&padded[static_cast<size_t>(m) * p->ld_bins], mel_basis_host + static_cast<size_t>(m) * p->n_bins, sizeof(float) * p->n_bins);
  cudaError_t e = cudaMemcpy(p->d_melw, padded.data(), nm * sizeof(float), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    const long long n = static_cast<long long>(p->n_bins) * cfg->n_fft;
    mel::basis_kernel<<<static_cast<unsigned>((n + 255) / 256), 256>>>(p->d_basis, cfg->n_fft, cfg->win, p->n_bins);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
  }
  if (e != cudaSuccess) {
    cudaFree(p->d_basis);
    cudaFree(p->d_melw);
    return fail("vcd_mel_plan_create: building the DFT basis failed: %s", cudaGetErrorString(e));
  }
  if ((cfg->n_fft & (cfg->n_fft - 1)) == 0) {
    // banded view of the filterbank: per filter its bin range, per bin its filter range (zeros inside a range are kept,
    // so any matrix is handled exactly; a triangular filterbank touches ~2 filters per bin)
    const int nm = cfg->n_mel, nbins = p->n_bins;
    std::vector<int> tab(3 * static_cast<size_t>(nm) + 3 * static_cast<size_t>(nbins), 0);
    int* f_lo = tab.data(); int* f_cnt = f_lo + nm; int* f_off = f_cnt + nm;
    int* b_lo = f_off + nm; int* b_cnt = b_lo + nbins; int* b_off = b_cnt + nbins;
    std::vector<float> vals;
    auto W = [&](int m, int k) { return mel_basis_host[static_cast<size_t>(m) * nbins + k]; };
    for (int m = 0; m < nm; ++m) {
      int lo = nbins, hi = -1;
      for (int k = 0; k < nbins; ++k)
        if (W(m, k) != 0.f) { if (k < lo) lo = k; hi = k; }
      f_lo[m] = hi < 0 ? 0 : lo;
      f_cnt[m] = hi < 0 ? 0 : hi - lo + 1;
      f_off[m] = static_cast<int>(vals.size());
      for (int k = f_lo[m]; k < f_lo[m] + f_cnt[m]; ++k) vals.push_back(W(m, k));
    }
    p->f_val_n = vals.size();
    for (int k = 0; k < nbins; ++k) {
      int lo = nm, hi = -1;
      for (int m = 0; m < nm; ++m)
        if (W(m, k) != 0.f) { if (m < lo) lo = m; hi = m; }
      b_lo[k] = hi < 0 ? 0 : lo;
      b_cnt[k] = hi < 0 ? 0 : hi - lo + 1;
      b_off[k] = static_cast<int>(vals.size() - p->f_val_n);
      for (int m = b_lo[k]; m < b_lo[k] + b_cnt[k]; ++m) vals.push_back(W(m, k));
    }
    if (vals.empty()) vals.push_back(0.f);
    int l2 = 0;
    while ((1 << l2) < cfg->n_fft) ++l2;
    p->log2n = l2;
    p->fft_smem = sizeof(float2) * (3 * static_cast<size_t>(cfg->n_fft)) +
                  sizeof(float) * (static_cast<size_t>((nbins + 3) & ~3) + nm);
    cudaError_t e2 = cudaMalloc(&p->d_window, sizeof(float) * cfg->n_fft);
    if (e2 == cudaSuccess) e2 = cudaMalloc(&p->d_tw, sizeof(float2) * cfg->n_fft);
    if (e2 == cudaSuccess) e2 = cudaMalloc(&p->d_tab, sizeof(int) * tab.size());
    if (e2 == cudaSuccess) e2 = cudaMalloc(&p->d_vals, sizeof(float) * vals.size());
    if (e2 == cudaSuccess) e2 = cudaMemcpy(p->d_tab, tab.data(), sizeof(int) * tab.size(), cudaMemcpyHostToDevice);
    if (e2 == cudaSuccess) e2 = cudaMemcpy(p->d_vals, vals.data(), sizeof(float) * vals.size(), cudaMemcpyHostToDevice);
    // opt-in limit of the kernel = the largest request of any plan (a later, smaller plan must not lower it)
    if (e2 == cudaSuccess && p->fft_smem <= 200 * 1024)
      e2 = cudaFuncSetAttribute(mel::frame_fft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e2 == cudaSuccess) {
      mel::fft_tables_kernel<<<(cfg->n_fft + 255) / 256, 256>>>(p->d_window, p->d_tw, cfg->n_fft, cfg->win);
      g_launches.fetch_add(1, std::memory_order_relaxed);
      e2 = cudaGetLastError();
      if (e2 == cudaSuccess) e2 = cudaDeviceSynchronize();
    }
    if (e2 != cudaSuccess) {
      cudaGetLastError();
      vcd_mel_plan_destroy(p.release());
      return fail("vcd_mel_plan_create: building the FFT tables failed: %s", cudaGetErrorString(e2));
    }
    p->fft_ok = p->fft_smem <= 200 * 1024;
  }
  *out = p.release();
  return 0;
}

extern "C" void vcd_mel_plan_destroy(vcd_mel_plan* p) {
  if (!p) return;
  cudaFree(p->d_basis);
  cudaFree(p->d_melw);
  cudaFree(p->d_window);
  cudaFree(p->d_tw);
  cudaFree(p->d_tab);
  cudaFree(p->d_vals);
  delete p;
}

extern "C" int vcd_mel_debug_path(vcd_mel_plan* p, int use_gemm) {
  if (!p) return fail("vcd_mel_debug_path: null plan");
  if (!use_gemm && !p->fft_ok) return fail("vcd_mel_debug_path: n_fft (%d) is not a power of two: only the GEMM path exists", p->cfg.n_fft);
  p->use_gemm = use_gemm != 0;
  return 0;
}

// fast path: one launch does everything per frame (see mel_loss.cuh)
static int mel_fft_launch(vcd_mel_plan* p, const float* y, const float* target, const long long* starts, int F_tgt, float* mel_out,
                          float scale, float* dframe, float* partials, const MelWs& w, int T, cudaStream_t st) {
  const vcd_mel_config& c = p->cfg;
  const int nm = c.n_mel, nbins = p->n_bins;
  mel::FftParams P{};
  P.y = y; P.T = T; P.F = w.frames; P.hop = c.hop; P.pad = (c.n_fft - c.hop) / 2; P.rows = w.rows;
  P.n_fft = c.n_fft; P.log2n = p->log2n; P.n_bins = nbins; P.n_mel = nm;
  P.window = p->d_window; P.tw = p->d_tw;
  P.f_lo = p->d_tab; P.f_cnt = P.f_lo + nm; P.f_off = P.f_cnt + nm;
  P.b_lo = P.f_off + nm; P.b_cnt = P.b_lo + nbins; P.b_off = P.b_cnt + nbins;
  P.f_val = p->d_vals; P.b_val = p->d_vals + p->f_val_n;
  P.out = mel_out; P.target = target; P.starts = starts; P.F_tgt = F_tgt; P.scale = scale; P.dframe = dframe; P.partials = partials;
  mel::frame_fft_kernel<<<static_cast<unsigned>(w.rows), 256, p->fft_smem, st>>>(P);
  LAUNCH_CHECK("mel::frame_fft_kernel");
  return 0;
}

extern "C" int vcd_mel_frames(const vcd_mel_plan* p, int T) { return p ? mel_frames(p->cfg, T) : 0; }

extern "C" size_t vcd_mel_workspace_bytes(const vcd_mel_plan* p, int B, int T) {
  if (!p || B < 1 || T < 1) return 0;
  return mel_ws_layout(p, B, T).total;
}

// forward part shared by the two entry points: gemm 1 (spectrum magnitudes) and gemm 2 (log-mel / loss / dM)
static int mel_forward(vcd_mel_plan* p, const float* y, const float* target, const long long* starts, int F_tgt, float* mel_out,
                       float scale, bool want_grad, uint8_t* ws, const MelWs& w, int B, int T, cudaStream_t st) {
  const vcd_mel_config& c = p->cfg;
  const int pad = (c.n_fft - c.hop) / 2;
  float* S = reinterpret_cast<float*>(ws + w.S);
  float* mag = reinterpret_cast<float*>(ws + w.mag);
  {
    mel::Frames A{y, T, w.frames, c.hop, pad, w.rows, c.n_fft};
    mel::RowsK Bm{p->d_basis, c.n_fft, 2 * p->n_bins, c.n_fft};
    mel::EpiSpectrum E{want_grad ? S : nullptr, p->ld_spec, mag, p->ld_bins, w.rows, p->n_bins};
    dim3 grid((2 * p->n_bins + 127) / 128, (w.rows + 63) / 64);
    mel::gemm_kernel<64, 128, 4, 8><<<grid, 256, 0, st>>>(c.n_fft, A, Bm, E, nullptr);
    LAUNCH_CHECK("mel::gemm_kernel (stft)");
  }
  {
    mel::RowsK A{mag, p->ld_bins, w.rows, p->n_bins};
    mel::RowsK Bm{p->d_melw, p->ld_bins, c.n_mel, p->n_bins};
    mel::EpiLogMel E{mel_out, target, reinterpret_cast<float*>(ws + w.dM), c.n_mel, scale, w.rows, c.n_mel, w.frames, starts, F_tgt};
    dim3 grid((c.n_mel + 63) / 64, (w.rows + 63) / 64);
    mel::gemm_kernel<64, 64, 4, 4><<<grid, 256, 0, st>>>(p->n_bins, A, Bm, E, reinterpret_cast<float*>(ws + w.partials));
    LAUNCH_CHECK("mel::gemm_kernel (mel)");
  }
  return 0;
}

static int mel_check(const char* what, vcd_mel_plan* p, const void* a, const void* b, const void* ws, size_t ws_bytes, int B,
                     int T, MelWs* w) {
  if (!p || !a || !b || !ws) return fail("%s: null argument", what);
  if (B < 1 || T < 1) return fail("%s: B and T must be positive", what);
  const int pad = (p->cfg.n_fft - p->cfg.hop) / 2;
  if (T <= pad) return fail("%s: T (%d) must exceed the reflect pad (%d samples)", what, T, pad);
  *w = mel_ws_layout(p, B, T);
  if (w->frames < 1) return fail("%s: T (%d) is shorter than one frame", what, T);
  if (ws_bytes < w->total) return fail("%s: workspace too small (%zu < %zu bytes)", what, ws_bytes, w->total);
  return 0;
}

extern "C" int vcd_mel_spectrogram(vcd_mel_plan* p, const float* y_dev, float* mel_dev, void* ws_dev, size_t ws_bytes, int B,
                                   int T, void* stream) {
  MelWs w;
  TRY(mel_check("vcd_mel_spectrogram", p, y_dev, mel_dev, ws_dev, ws_bytes, B, T, &w));
  if (p->fft_ok && !p->use_gemm)
    return mel_fft_launch(p, y_dev, nullptr, nullptr, 0, mel_dev, 0.f, nullptr,
                          reinterpret_cast<float*>(static_cast<uint8_t*>(ws_dev) + w.partials), w, T, static_cast<cudaStream_t>(stream));
  return mel_forward(p, y_dev, nullptr, nullptr, 0, mel_dev, 0.f, false, static_cast<uint8_t*>(ws_dev), w, B, T,
                     static_cast<cudaStream_t>(stream));
}

static int mel_loss_impl(vcd_mel_plan* p, const float* y_hat_dev, const float* mel_target_dev, const long long* starts, int F_tgt,
                         float c_mel, float* loss_dev, float* dy_dev, void* ws_dev, size_t ws_bytes, int B, int T, void* stream);

extern "C" int vcd_mel_loss(vcd_mel_plan* p, const float* y_hat_dev, const float* mel_target_dev, float c_mel, float* loss_dev,
                            float* dy_dev, void* ws_dev, size_t ws_bytes, int B, int T, void* stream) {
  return mel_loss_impl(p, y_hat_dev, mel_target_dev, nullptr, 0, c_mel, loss_dev, dy_dev, ws_dev, ws_bytes, B, T, stream);
}

extern "C" int vcd_mel_loss_sliced(vcd_mel_plan* p, const float* y_hat_dev, const float* mel_full_dev, int64_t frames_full,
                                   const int64_t* starts_dev, float c_mel, float* loss_dev, float* dy_dev, void* ws_dev,
                                   size_t ws_bytes, int B, int T, void* stream) {
  if (!starts_dev) return fail("vcd_mel_loss_sliced: null starts");
  if (!p || frames_full < mel_frames(p->cfg, T)) return fail("vcd_mel_loss_sliced: the full-length mel has fewer frames than one segment");
  return mel_loss_impl(p, y_hat_dev, mel_full_dev, reinterpret_cast<const long long*>(starts_dev), static_cast<int>(frames_full), c_mel,
                       loss_dev, dy_dev, ws_dev, ws_bytes, B, T, stream);
}

static int mel_loss_impl(vcd_mel_plan* p, const float* y_hat_dev, const float* mel_target_dev, const long long* starts, int F_tgt,
                         float c_mel, float* loss_dev, float* dy_dev, void* ws_dev, size_t ws_bytes, int B, int T, void* stream) {
  MelWs w;
  TRY(mel_check("vcd_mel_loss", p, y_hat_dev, mel_target_dev, ws_dev, ws_bytes, B, T, &w));
  if (!loss_dev) return fail("vcd_mel_loss: null loss pointer");
  if (!starts) F_tgt = w.frames;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* ws = static_cast<uint8_t*>(ws_dev);
  const vcd_mel_config& c = p->cfg;
  const float scale = c_mel / (static_cast<float>(B) * c.n_mel * w.frames);
  float* dframe = reinterpret_cast<float*>(ws + w.dframe);
  const bool fast = p->fft_ok && !p->use_gemm;
  if (fast) TRY(mel_fft_launch(p, y_hat_dev, mel_target_dev, starts, F_tgt, nullptr, scale, dy_dev ? dframe : nullptr,
                               reinterpret_cast<float*>(ws + w.partials), w, T, st));
  else TRY(mel_forward(p, y_hat_dev, mel_target_dev, starts, F_tgt, nullptr, scale, dy_dev != nullptr, ws, w, B, T, st));
  const int n_part = fast ? w.rows : ((c.n_mel + 63) / 64) * ((w.rows + 63) / 64);
  mel::loss_finalize_kernel<<<1, 256, 0, st>>>(reinterpret_cast<const float*>(ws + w.partials), n_part, scale, loss_dev);
  LAUNCH_CHECK("mel::loss_finalize_kernel");
  if (!dy_dev) return 0;
  float* S = reinterpret_cast<float*>(ws + w.S);
  float* mag = reinterpret_cast<float*>(ws + w.mag);
  float* dM = reinterpret_cast<float*>(ws + w.dM);
  if (!fast) {
    mel::RowsK A{dM, c.n_mel, w.rows, c.n_mel};
    mel::KCols Bm{p->d_melw, p->ld_bins, c.n_mel, p->n_bins};
    mel::EpiMagGrad E{S, p->ld_spec, mag, p->ld_bins, w.rows, p->n_bins};
    dim3 grid((p->n_bins + 63) / 64, (w.rows + 63) / 64);
    mel::gemm_kernel<64, 64, 4, 4><<<grid, 256, 0, st>>>(c.n_mel, A, Bm, E, nullptr);
    LAUNCH_CHECK("mel::gemm_kernel (mel backward)");
  }
  if (!fast) {
    mel::RowsK A{S, p->ld_spec, w.rows, 2 * p->n_bins};
    mel::KCols Bm{p->d_basis, c.n_fft, 2 * p->n_bins, c.n_fft};
    mel::EpiStore E{dframe, c.n_fft, w.rows, c.n_fft};
    dim3 grid((c.n_fft + 127) / 128, (w.rows + 63) / 64);
    mel::gemm_kernel<64, 128, 4, 8><<<grid, 256, 0, st>>>(2 * p->n_bins, A, Bm, E, nullptr);
    LAUNCH_CHECK("mel::gemm_kernel (stft backward)");
  }
  {
    const long long total = static_cast<long long>(B) * T;
    mel::overlap_add_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(dframe, c.n_fft, c.hop, (c.n_fft - c.hop) / 2,
                                                                                        w.frames, T, total, dy_dev);
    LAUNCH_CHECK("mel::overlap_add_kernel");
  }
  return 0;
}

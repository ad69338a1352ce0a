/*
 * vcd.h -- C ABI of the B200-native VCVITS HiFi-GAN waveform decoder ("vcd" = vcvits decoder).
 *
 * The reference has no FFI / plugin layer for this path: the decoder is a plain torch.nn.Module attribute
 * `net_g.dec` (reference: vits/model/synthesizers/synthesizer_tts.py:71-78 constructor call,
 * synthesizer_tts.py:140,166,176 and synthesizer_svc.py:87,108,118 forward call sites).  The entry points
 * below are therefore the operator boundary a maintainer would bind with ctypes from the module that
 * replaces `Generator` (see INTEGRATION.md for the binding); each one names the reference interface it
 * stands in for.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types cross the boundary.
 *   - every pointer named *_dev is device memory on the current CUDA device; *_host is host memory.
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).  All work is enqueued
 *     asynchronously on it; nothing synchronises the device unless stated.
 *   - return value 0 = success; non-zero = error, message via vcd_last_error() (thread-local).
 *     Nothing throws or aborts.
 *   - the caller owns every tensor and the workspace; the library owns only the plan.
 *   - one plan per (device, process); a plan must not be used from two host threads at once.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point returns an error.
 */
#ifndef VCD_H_
#define VCD_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VCD_MAX_KERNELS 8
#define VCD_MAX_DILATIONS 3
#define VCD_MAX_UPSAMPLES 8

/* Arithmetic mode of the decoder activations / tensor-core operands. */
#define VCD_MODE_FP32 0 /* fp32 storage, fp32 FFMA math: the parity mode (<=1e-4 max-abs vs. the oracle)   */
#define VCD_MODE_BF16 1 /* bf16 operands on tcgen05 tensor cores, fp32 accumulation in TMEM.  Every tensor kept between
                         * layers is bf16(leaky_relu(.)); the raw residual stream of a ResBlock is NOT stored but recovered
                         * from that bf16 tensor through the exact inverse of leaky_relu; fp32 survives
                         * only in the running sums over ResBlock branches and the boundary tensors. */

/* Constructor arguments of `Generator` (synthesizer_tts.py:71-78; values from configs/base.json:55-67). */
typedef struct vcd_config {
  int32_t initial_channel;                                   /* inter_channels                          */
  int32_t resblock;                                          /* 1 -> ResBlock1 (modules.py:186), else 2 */
  int32_t num_kernels;                                       /* len(resblock_kernel_sizes)              */
  int32_t resblock_kernel_sizes[VCD_MAX_KERNELS];
  int32_t resblock_dilation_sizes[VCD_MAX_KERNELS][VCD_MAX_DILATIONS];
  int32_t num_upsamples;                                     /* len(upsample_rates)                     */
  int32_t upsample_rates[VCD_MAX_UPSAMPLES];
  int32_t upsample_kernel_sizes[VCD_MAX_UPSAMPLES];
  int32_t upsample_initial_channel;
  int32_t gin_channels;                                      /* 0 -> no `cond` conv                     */
} vcd_config;

typedef struct vcd_plan vcd_plan;

/* Library / build info: returns e.g. "vcd 0.1 sm_100a".  Never fails. */
const char* vcd_version(void);

/* Last error message of the calling thread ("" if none). */
const char* vcd_last_error(void);

/* Stands in for Generator.__init__ (synthesizer_tts.py:71-78): validates the configuration, builds the
 * layer table and the parameter table.  Needs a CUDA device (queries SM count / smem limits). */
int vcd_plan_create(const vcd_config* cfg, vcd_plan** out_plan);
void vcd_plan_destroy(vcd_plan* plan);

/* Parameter table == the module's state_dict: name uses the reference's
 * checkpoint keys ("conv_pre.weight", "ups.0.weight_g", "resblocks.3.convs1.0.weight_v", ...). */
int vcd_num_params(const vcd_plan* plan);
int vcd_param_info(const vcd_plan* plan, int index, const char** name, int64_t shape[3], int* ndim);
int64_t vcd_total_param_elems(const vcd_plan* plan);

/* Output samples per latent frame = prod(upsample_rates) (hop_length, configs/base.json:33). */
int vcd_hop(const vcd_plan* plan);

/* Bytes of caller-owned device workspace for a [B, initial_channel, T] call.  `save_for_backward` != 0
 * keeps every activation the backward pass needs (training step); 0 recycles buffers (infer.py path). */
size_t vcd_workspace_bytes(const vcd_plan* plan, int mode, int B, int T, int save_for_backward);

/* Stands in for the weight_norm forward pre-hooks (modules.py:10,190-199,229-230: w = g * v / ||v||):
 * folds all parameters into the packed operand layouts the kernels consume.  params_dev_ptrs: host array
 * of vcd_num_params() DEVICE pointers (fp32, contiguous, shapes per vcd_param_info).  Must be called after
 * every parameter update and before forward.
 * `remove_weight_norm` (modules.py:218-222): a NULL pointer in a `weight_g` slot says that layer carries no
 * weight norm any more -- its `weight_v` slot then holds the baked weight, which is used as is (and whose
 * plain gradient vcd_backward returns in the `weight_v` slot; the `weight_g` gradient pointer may be NULL). */
int vcd_fold_weights(vcd_plan* plan, int mode, const float* const* params_dev_ptrs, void* stream);

/* Stands in for Generator.forward(x, g) (call sites synthesizer_tts.py:140, synthesizer_svc.py:87,108).
 *   x_dev : fp32 [B, initial_channel, T] with element strides (xs_b, xs_c, xs_t) -- a non-contiguous
 *           slice like (z*y_mask)[:, :, :max_len] (synthesizer_tts.py:166) is accepted as is.
 *   g_dev : fp32 [B, gin_channels] contiguous (the [B, gin, 1] speaker embedding) or NULL.
 *   y_dev : fp32 [B, 1, T*hop] contiguous, output waveform.
 *   ws_dev/ws_bytes : workspace of at least vcd_workspace_bytes(...).  With save_for_backward the same
 *           workspace must be handed, untouched, to vcd_backward. */
int vcd_forward(vcd_plan* plan, int mode, const float* x_dev, int64_t xs_b, int64_t xs_c, int64_t xs_t,
                const float* g_dev, float* y_dev, void* ws_dev, size_t ws_bytes, int B, int T,
                int save_for_backward, void* stream);

/* Stands in for autograd through Generator.forward (the backward of the train.py step: vcvits.py:54-148):
 *   dy_dev      : fp32 [B, 1, T*hop] upstream gradient.
 *   y_dev       : the waveform vcd_forward produced (tanh backward needs it); g_dev as in vcd_forward.
 *   dx_dev      : fp32 [B, initial_channel, T] contiguous, or NULL to skip.
 *   dg_dev      : fp32 [B, gin_channels], or NULL.
 *   dparams_dev_ptrs : host array of vcd_num_params() DEVICE pointers receiving the parameter gradients
 *                 (overwritten, not accumulated): weight_g / weight_v / bias exactly like autograd through
 *                 old-style weight_norm.  Needs the params table of the preceding vcd_fold_weights.
 *   segment_mask: bit i set -> run backward segment i (see vcd_num_backward_segments); ~0u runs everything.
 *                 Segments must be executed in increasing order; gradients of segment i are final when the
 *                 call that ran it returns (enqueued), which lets the caller overlap a gradient all-reduce
 *                 of finished segments with the rest of backward (replaces the DDP reducer, train.py:99-100). */
int vcd_backward(vcd_plan* plan, int mode, const float* dy_dev, const float* y_dev, const float* g_dev,
                 float* dx_dev, float* dg_dev, float* const* dparams_dev_ptrs, void* ws_dev, size_t ws_bytes,
                 int B, int T, uint32_t segment_mask, void* stream);

/* the training-time segment gather in front of the decoder,
 *     z_slice, ids = commons.rand_slice_segments(z, lengths, segment_size)   (vits/commons.py:48-64; call sites
 *     synthesizer_tts.py:138, synthesizer_svc.py:86)
 * folded into the decoder's input load: `z_dev` is the FULL-length latent [B, initial_channel, T_full] (element
 * strides as for vcd_forward), `starts_dev` a device array of B int64 first frames (ids_slice), T the segment length.
 * Equivalent to vcd_forward on slice_segments(z, starts, T) without materialising the slice.  vcd_backward_sliced
 * returns the gradient w.r.t. the full-length latent ([B, initial_channel, T_full] contiguous: zero outside each
 * item's segment -- the scatter-add that autograd derives from the per-item copies of slice_segments). */
int vcd_forward_sliced(vcd_plan* plan, int mode, const float* z_dev, int64_t zs_b, int64_t zs_c, int64_t zs_t,
                       const int64_t* starts_dev, const float* g_dev, float* y_dev, void* ws_dev, size_t ws_bytes, int B,
                       int T, int save_for_backward, void* stream);
int vcd_backward_sliced(vcd_plan* plan, int mode, const float* dy_dev, const float* y_dev, const float* g_dev,
                        float* dz_dev, int64_t T_full, const int64_t* starts_dev, float* dg_dev,
                        float* const* dparams_dev_ptrs, void* ws_dev, size_t ws_bytes, int B, int T, uint32_t segment_mask,
                        void* stream);

/* Data-parallel training: every PARAMETER gradient written by vcd_backward is multiplied by `scale` (not dx / dg).
 * With scale = 1 / world_size a plain SUM all-reduce of the gradient buffers yields the DDP average
 * (train.py:99-100) without a separate scaling pass.  Default 1. */
int vcd_set_gradient_scale(vcd_plan* plan, float scale);

/* Run-to-run reproducibility of the parameter gradients.  By default the split partial sums of a weight gradient (and
 * the bias column sums) are combined with fp32 atomics in arrival order: fast, but the last bits vary between runs.
 * on != 0: the splits of a weight-gradient tile take a ticket and add their partial sums one after the other in
 * time-range order (the main loops still overlap); the remaining column-sum / conv_post reductions receive at most TWO
 * atomic contributions onto a zeroed value (a + b == b + a).  All 233 gradients, dx and dg are then bit-identical
 * from run to run, at a cost in step time (measured in DESIGN.md). */
int vcd_set_deterministic(vcd_plan* plan, int on);

/* When vcd_backward runs every segment in ONE call (segment_mask = ~0u) the data-gradient chain runs through all stages
 * while the weight gradients / weight-norm backward of finished segments trail behind it; the library then records one
 * event per segment ("its gradients are final").  vcd_stream_wait_segment makes `stream` wait for segment `segment`, so a
 * caller can overlap that segment's gradient all-reduce (the DDP reducer of train.py:99-100) with the rest of backward
 * without splitting the call.  vcd_segment_events_valid: 1 if the last vcd_backward recorded the events (it does not under
 * the per-launch profiler or VCD_SERIAL=1). */
int vcd_segment_events_valid(const vcd_plan* plan);
int vcd_stream_wait_segment(vcd_plan* plan, int segment, void* stream);

/* Backward segments, in execution order: segment 0 = conv_post + last upsample stage, ...,
 * last segment = conv_pre + cond.  vcd_segment_params lists the parameter indices finalised by a segment. */
int vcd_num_backward_segments(const vcd_plan* plan);
int vcd_segment_params(const vcd_plan* plan, int segment, int* indices, int cap);

/* infer.py-style end-to-end call with HOST buffers (infer.py:83-91: latent in, waveform out):
 * copies x_host (+ g_host) to the device, runs forward without saving activations, copies the waveform back
 * and synchronises `stream`.  ws_dev as for vcd_forward with save_for_backward = 0, plus the staging the
 * function reports through vcd_host_call_extra_bytes(). */
size_t vcd_host_call_extra_bytes(const vcd_plan* plan, int B, int T);
int vcd_synthesize_host(vcd_plan* plan, int mode, const float* x_host, const float* g_host, float* y_host,
                        void* ws_dev, size_t ws_bytes, int B, int T, void* stream);

/* Counters for bench.py: kernels launched by this library since the last reset. */
uint64_t vcd_launch_count(int reset);

/* Optional per-launch profiler for bench.py's roofline object: when enabled every kernel launch is bracketed
 * by CUDA events on the launch stream and accumulated per kernel class (adds overhead: never enabled while the
 * throughput value is timed).  vcd_profile_read synchronises the device and fills arrays of
 * vcd_profile_num_classes() entries: device milliseconds, launches, algorithmic FLOPs, algorithmic bytes. */
int vcd_profile_enable(int on);
int vcd_profile_num_classes(void);
const char* vcd_profile_class_name(int class_id);
int vcd_profile_read(int reset, double* ms, uint64_t* launches, double* flops, double* bytes);
int vcd_profile_dump(const char* csv_path); /* one line per recorded launch: class, layer tag, ms, GFLOP */

/* Debug only (VCD_PHASES=1): print the device time of the fold, the forward core and every backward segment. */
int vcd_phase_dump(int reset);

/* Debug only: 64 in-kernel %globaltimer stamps of the kernel selected with the VCD_KTRACE environment variable. */
int vcd_debug_read_trace(vcd_plan* plan, unsigned long long* out64);

/* Debug / tests only: where an internal tensor lives inside the caller-owned workspace.  Layout: blocked
 * channels-last with row pads, [B][C/8][pad_l + L + pad_r][8] elements of 2 bytes (bf16 mode) or 4 (fp32 mode).
 * Names: xin, a<i>, ua<i>, ma<i>.<j>.<q>, xa<i>.<j>.<q> (forward activations; i = stage, j = ResBlock branch,
 * q = conv pair) and Gi<i>, Gt<i>.<j>.<q>, dm<i>.<j>.<q>, duz<i> (phase-packed), d0 (gradients; the buffers are
 * shared between stages, so they are valid right after the backward segment of stage i ran).  The parity tests
 * use it to check every kernel launch against the oracle on that launch's own stored operands. */
int vcd_debug_ws_tensor(const vcd_plan* plan, int mode, int B, int T, int save_for_backward, const char* name,
                        size_t* offset_bytes, int* C, int* L, int* pad_l, int* pad_r);

/* Debug / tests only: for plans created AFTER the call, run the forward / data-gradient / weight-gradient
 * convolutions of bf16 mode on the tcgen05 kernels (1) or on the FFMA kernels with the same bf16 operands (0);
 * a negative value leaves that flag unchanged.  The environment variables VCD_TC_FWD / VCD_TC_DGRAD /
 * VCD_TC_WGRAD give the initial values. */
int vcd_debug_tc_paths(int fwd, int dgrad, int wgrad);

/* Per-layer timing / debugging: name of the arithmetic path ("simt-fp32", "tcgen05-bf16", ...) used by
 * layer `index` of the forward schedule in `mode`; NULL past the end. */
const char* vcd_layer_path(const vcd_plan* plan, int mode, int index);

/* ---------------------------------------------------------------------------------------------------------------
 * Mel / STFT loss tail of the generator step: the consumer of the decoder's waveform.
 * Replaces, with its backward,
 *     y_spec_hat = spectrogram_torch_audio(y_hat, filter_length, sr, hop_length, win_length, center=False)
 *                                                                       (vits/mel_processing.py:76-95; call vits/light/vcvits.py:96-100)
 *     y_mel_hat  = spec_to_mel_torch(y_spec_hat, filter_length, n_mel_channels, sr, mel_fmin, mel_fmax)
 *                                                                       (mel_processing.py:97-112; call vcvits.py:102-108)
 *     loss_mel   = F.l1_loss(y_mel_hat, y_mel_slice) * c_mel            (vcvits.py:115)
 * fp32 throughout (the reference runs torch.stft in fp32).  All reductions have a fixed order: bit-identical results
 * from run to run.  The Slaney filterbank comes from the caller (the reference takes it from librosa.filters.mel,
 * mel_processing.py:101-103); the Hann window and the DFT basis are built by the library. */
typedef struct vcd_mel_config {
  int32_t n_fft;          /* filter_length  (configs/base.json:32) */
  int32_t hop;            /* hop_length     (base.json:33)         */
  int32_t win;            /* win_length     (base.json:34), <= n_fft: the periodic Hann window is centred in the frame */
  int32_t n_mel;          /* n_mel_channels (base.json:35)         */
} vcd_mel_config;

typedef struct vcd_mel_plan vcd_mel_plan;

/* mel_basis_host: HOST fp32 [n_mel][n_fft/2 + 1], row-major (what librosa.filters.mel returns). */
int vcd_mel_plan_create(const vcd_mel_config* cfg, const float* mel_basis_host, vcd_mel_plan** out_plan);
void vcd_mel_plan_destroy(vcd_mel_plan* plan);

/* Frames per item for T samples: reflect pad (n_fft - hop) / 2 per side, center=False -> T / hop when hop divides T. */
int vcd_mel_frames(const vcd_mel_plan* plan, int T);
size_t vcd_mel_workspace_bytes(const vcd_mel_plan* plan, int B, int T);

/* mel_spectrogram_torch(y, ...) (mel_processing.py:115-142) = spec_to_mel_torch(spectrogram_torch_audio(y)):
 *   y_dev [B, T] fp32 -> mel_dev [B, n_mel, frames] fp32 log-mel. */
int vcd_mel_spectrogram(vcd_mel_plan* plan, const float* y_dev, float* mel_dev, void* ws_dev, size_t ws_bytes, int B, int T,
                        void* stream);

/* The loss tail and its gradient in one call:
 *   y_hat_dev [B, T] (the decoder output [B, 1, T]), mel_target_dev [B, n_mel, frames] (y_mel_slice),
 *   *loss_dev = c_mel * mean |logmel(y_hat) - mel_target|,
 *   dy_dev [B, T] = d loss / d y_hat (the decoder's upstream gradient; NULL: loss only). */
int vcd_mel_loss(vcd_mel_plan* plan, const float* y_hat_dev, const float* mel_target_dev, float c_mel, float* loss_dev,
                 float* dy_dev, void* ws_dev, size_t ws_bytes, int B, int T, void* stream);

/* The same with y_mel_slice = commons.slice_segments(y_mel, ids_slice, segment_size // hop_length) (vits/commons.py:48-55;
 * call vcvits.py:110) folded into the target read: mel_full_dev [B, n_mel, frames_full] is the whole-utterance log-mel,
 * starts_dev B int64 first frames (ids_slice, on the device); frames starts[b] .. starts[b] + frames - 1 are compared.
 * The caller guarantees starts[b] + frames <= frames_full (rand_slice_segments does). */
int vcd_mel_loss_sliced(vcd_mel_plan* plan, const float* y_hat_dev, const float* mel_full_dev, int64_t frames_full,
                        const int64_t* starts_dev, float c_mel, float* loss_dev, float* dy_dev, void* ws_dev, size_t ws_bytes,
                        int B, int T, void* stream);

/* Debug / tests only.  Two implementations exist: a per-frame shared-memory FFT kernel (n_fft a power of two; the
 * default when available) and dense fp32 GEMMs against a DFT basis (any n_fft).  use_gemm != 0 forces the second, so
 * that the tests can check one against the other. */
int vcd_mel_debug_path(vcd_mel_plan* plan, int use_gemm);

#ifdef __cplusplus
}
#endif
#endif /* VCD_H_ */

/*
 * b200sep -- C ABI of the B200-native Conv-TasNet separation + PIT SI-SDR backend.
 *
 * This is the drop-in boundary for the reference's separation stage.  Plain pointers and
 * sizes only (no torch types).  Every entry point names the reference interface it replaces
 * (paths relative to the reference repo NatsuiroGinga/audio-classification).
 *
 * Conventions
 *   - All `d_*` pointers are DEVICE pointers on the model's device; everything else is host.
 *   - Every function returning int returns 0 on success, non-zero on failure; the message is
 *     available from b200sep_last_error() (thread-local).  There is no CPU fallback: a missing
 *     sm_100 device is an error.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  All work is
 *     enqueued asynchronously on it; no call synchronises the device.
 *   - Segments ("overlap segments" / "chunks") are independent units.  A batch is a ragged
 *     concatenation: segment i is `lens[i]` fp32 samples starting at `d_mix + offsets[i]`
 *     (offsets == NULL means tightly packed: offsets[i] = sum(lens[0..i-1])).
 *     Separated output for segment i, source s is `lens[i]` fp32 samples at
 *     `d_out + n_src * offsets[i] + s * lens[i]`  -- i.e. exactly the (n_src, T) array the
 *     reference returns per call, concatenated.
 */
#ifndef B200SEP_H_
#define B200SEP_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200SEP_VERSION 200 /* 0.2.0 */

/* mask activation of TDConvNet (asteroid mask_act) */
#define B200SEP_MASK_SIGMOID 0
#define B200SEP_MASK_RELU 1

/* arithmetic of the 1x1-conv GEMMs and activation storage */
#define B200SEP_PREC_FP32 0      /* fp32 storage, CUDA-core fp32 GEMMs ("exact" mode, >= 100 dB vs oracle) */
#define B200SEP_PREC_FP16_TC 1   /* fp16 operands, fp32 accumulate, tcgen05/TMEM/TMA GEMMs (default)       */
#define B200SEP_PREC_FP16_SIMT 2 /* fp16 storage, CUDA-core GEMMs (validation of the tensor-core path)     */

/*
 * Hyper-parameters of asteroid.models.ConvTasNet (constructor kwargs; replaces the model
 * built at src/osd/separation.py:44-50).  Supported: n_src in {2,3}; n_filters=512,
 * bn_chan=128, hid_chan=512, skip_chan=128, conv_kernel_size=3; (kernel_size, stride) in
 * {(16,8),(32,16)}; any n_blocks<=10, n_repeats>=1; gLN; sigmoid or relu mask.
 */
typedef struct b200sep_config {
  int32_t n_src;
  int32_t n_filters;        /* N */
  int32_t kernel_size;      /* L */
  int32_t stride;
  int32_t bn_chan;          /* B */
  int32_t hid_chan;         /* H */
  int32_t skip_chan;        /* Sc */
  int32_t conv_kernel_size; /* P */
  int32_t n_blocks;         /* X */
  int32_t n_repeats;        /* R */
  int32_t mask_act;         /* B200SEP_MASK_* */
  int32_t precision;        /* B200SEP_PREC_* */
} b200sep_config;

/*
 * Raw fp32 HOST tensors of an Asteroid ConvTasNet state_dict (key names in comments; this is
 * what torch.load(checkpoint)["state_dict"] holds, src/osd/separation.py:53-56).  Per-block
 * arrays have n_blocks*n_repeats entries, block i = masker.TCN.{i}.  The library copies,
 * folds the gLN affine parameters into the following convolution and converts to the GEMM
 * operand type at create time; the caller's buffers are not referenced afterwards.
 */
typedef struct b200sep_weights {
  const float* enc_filters;     /* encoder.filterbank._filters        (N,1,L) */
  const float* dec_filters;     /* decoder.filterbank._filters        (N,1,L) */
  const float* bn_gamma;        /* masker.bottleneck.0.gamma          (N)     */
  const float* bn_beta;         /* masker.bottleneck.0.beta           (N)     */
  const float* bn_weight;       /* masker.bottleneck.1.weight         (B,N,1) */
  const float* bn_bias;         /* masker.bottleneck.1.bias           (B)     */
  const float* const* in_weight;   /* TCN.i.shared_block.0.weight     (H,B,1) */
  const float* const* in_bias;     /* TCN.i.shared_block.0.bias       (H)     */
  const float* const* prelu1;      /* TCN.i.shared_block.1.weight     (1)     */
  const float* const* norm1_gamma; /* TCN.i.shared_block.2.gamma      (H)     */
  const float* const* norm1_beta;  /* TCN.i.shared_block.2.beta       (H)     */
  const float* const* dw_weight;   /* TCN.i.shared_block.3.weight     (H,1,P) */
  const float* const* dw_bias;     /* TCN.i.shared_block.3.bias       (H)     */
  const float* const* prelu2;      /* TCN.i.shared_block.4.weight     (1)     */
  const float* const* norm2_gamma; /* TCN.i.shared_block.5.gamma      (H)     */
  const float* const* norm2_beta;  /* TCN.i.shared_block.5.beta       (H)     */
  const float* const* res_weight;  /* TCN.i.res_conv.weight           (B,H,1) */
  const float* const* res_bias;    /* TCN.i.res_conv.bias             (B)     */
  const float* const* skip_weight; /* TCN.i.skip_conv.weight          (Sc,H,1)*/
  const float* const* skip_bias;   /* TCN.i.skip_conv.bias            (Sc)    */
  const float* mask_prelu;      /* masker.mask_net.0.weight           (1)     */
  const float* mask_weight;     /* masker.mask_net.1.weight           (S*N,Sc,1) */
  const float* mask_bias;       /* masker.mask_net.1.bias             (S*N)   */
  /*
   * eps of every GlobLN (asteroid masknn/norms.py: 1e-8), in model order: [0] masker.bottleneck.0, then for block i
   * [1 + 2 i] shared_block.2 and [2 + 2 i] shared_block.5.  NULL = 1e-8 everywhere.  Not a checkpoint tensor: the
   * host-side range conditioning (weights.condition_state_dict) rescales the tensors feeding a gLN by 2^p -- which
   * the norm removes again -- and passes eps * 4^p here so that the conditioned model computes EXACTLY the function
   * of the original one while its pre-norm activations sit in the fp16 range.
   */
  const float* gln_eps;         /* (1 + 2 * n_blocks * n_repeats) or NULL */
} b200sep_weights;

typedef struct b200sep_model b200sep_model; /* opaque: device weights, TMA descriptors */

/* Library / ABI version (B200SEP_VERSION). */
int b200sep_version(void);

/*
 * Host-only helper of the multi-GPU evaluation sweep (no reference counterpart: the reference runs its per-segment loop,
 * scripts/osd/overlap3_core.py:604-744, in one process): deterministic longest-processing-time-first partition of n
 * independent segments over `world` ranks.  Segments in descending `costs` (ties: ascending index) each go to the least
 * loaded rank (ties: lowest rank); owner[i] = rank of segment i.  Every rank computes the same partition.
 */
int b200sep_lpt_partition(const int64_t* costs, int32_t n, int32_t world, int32_t* owner);

/* Thread-local message of the last failing call on this thread ("" if none). */
const char* b200sep_last_error(void);

/*
 * Replaces Separator._init_asteroid (src/osd/separation.py:36-64): build the model on CUDA
 * device `device` from host weights.  Fails (no fallback) if the device is not sm_100.
 */
int b200sep_model_create(const b200sep_config* cfg, const b200sep_weights* w, int device,
                         b200sep_model** out);

void b200sep_model_destroy(b200sep_model* m);

/*
 * Bytes of device scratch b200sep_separate_batch needs for this batch (activations of every
 * layer for all segments + tables).  The caller owns the scratch; nothing is allocated inside
 * the timed call.  Returns 0 and sets the error on invalid input (e.g. a segment shorter than
 * kernel_size, which the reference's conv1d would also reject).
 */
size_t b200sep_workspace_bytes(const b200sep_model* m, const int32_t* lens, int32_t n_seg);

/*
 * Replaces the forward at src/osd/separation.py:82-83 (`est = self._model(x)`, x = (1,T)),
 * batched over independent segments: ConvTasNet forward of every segment, each normalised
 * (gLN) over its own frames exactly as a batch-1 call would be.  Bytes of d_out between
 * segments (when offsets leave gaps) are zero-filled.  `d_ws` must be 256-byte aligned.
 *
 * `h_flags` (NULL, or B200SEP_MAX_SPLIT int32 words of page-locked HOST memory owned by this call until `stream` has
 * been synchronised): the call's device flags, one word per concurrent sub-batch, copied back at the end of the
 * stream work -- OR them together.  Flags are per CALL (they live in the call's table block inside `d_ws`): nothing is
 * shared between engines, threads or calls and nothing is sticky.
 */
#define B200SEP_MAX_SPLIT 4
#define B200SEP_FLAG_WAIT_TIMEOUT 1 /* a bounded device-side pipeline wait ran out: results invalid (protocol bug)    */
#define B200SEP_FLAG_NONFINITE 2    /* the separated output holds inf / NaN samples (fp16 range exceeded or bad input) */
#define B200SEP_FLAG_EPS_DOMINATED 4 /* a gLN saw var < eps / 16: the signal is a ripple on the biases, beyond fp16 storage */
int b200sep_separate_batch(b200sep_model* m, const float* d_mix, const int64_t* offsets,
                           const int32_t* lens, int32_t n_seg, float* d_out, void* d_ws,
                           size_t ws_bytes, void* stream, int32_t* h_flags);

/* PIT enumeration / tie-break order to reproduce */
#define B200SEP_PIT_OVERLAP3 0 /* scripts/osd/overlap3_core.py:66-117                         */
#define B200SEP_PIT_EVALSRC 1  /* scripts/osd/evaluate_with_sources.py:334-441                */

/*
 * Number of concurrent sub-batches b200sep_separate_batch[_host] will cut this batch into (segments are independent;
 * a large batch runs as tile-balanced sub-batches on separate streams so that the persistent kernels of one fill the
 * SMs the other has drained; 1 while profiling / debugging hooks are on; B200SEP_SPLIT=<n> overrides).  -1 on error.
 */
int b200sep_split_count(const b200sep_model* m, const int32_t* lens, int32_t n_seg);
/*
 * Cap on the number of concurrent sub-batches the two forward entry points cut a call into (0: automatic, the default).
 * For callers that already overlap several batches on their own streams -- the pipelined forms of Separator.separate
 * (src/osd/separation.py:66-89) in audio-classification_b200/separator.py -- where a second level of splitting costs
 * more than it fills.  Results never depend on it.
 */
int b200sep_set_max_split(b200sep_model* m, int32_t n);

/*
 * Host boundary of the same forward: replaces Separator.separate (src/osd/separation.py:66-89: numpy in ->
 * `.to(device)` -> model -> `.cpu().numpy()`), batched.  `h_segs[i]` points to `lens[i]` fp32 samples in ANY host
 * memory; `h_out` (page-locked memory recommended) receives, for segment i and source s, `lens[i]` samples at
 * `h_out + n_src * sum(lens[0..i-1]) + s * lens[i]`.  The call gathers the segments into the library's pinned
 * staging, and queues H2D copy, forward and D2H copy of each concurrent sub-batch on its own stream; it returns when
 * everything is QUEUED -- synchronise `stream` before reading `h_out`.  `d_ws`: 256-byte aligned device scratch of
 * at least b200sep_host_workspace_bytes() bytes (forward scratch + packed device input / output).
 */
size_t b200sep_host_workspace_bytes(const b200sep_model* m, const int32_t* lens, int32_t n_seg);
int b200sep_separate_batch_host(b200sep_model* m, const float* const* h_segs, const int32_t* lens,
                                int32_t n_seg, float* h_out, void* d_ws, size_t ws_bytes, void* stream,
                                int32_t* h_flags);

/* bytes of device scratch needed by b200sep_pit_sisdr_batch for N <= 3 estimates per segment / for N <= 8 */
size_t b200sep_pit_workspace_bytes(const int32_t* lens, int32_t n_seg);
size_t b200sep_pit_workspace_bytes_n(const int32_t* lens, int32_t n_seg, int32_t N);

/*
 * Replaces _compute_sdr_improvement_pit_k / _pit_best_si_sdr_k / _si_sdr
 * (scripts/osd/overlap3_core.py:48-117) and compute_sdr_improvement_pit[_k]
 * (scripts/osd/evaluate_with_sources.py:377-441), batched: one pass over the samples
 * accumulates all first/second moments in fp64, then every pair SI-SDR and every permutation
 * is scored on the device.
 *   d_refs : K  reference sources per segment, segment i source k at d_refs + K*offsets[i] + k*lens[i]
 *   d_ests : N  estimates per segment (same layout with N; the layout b200sep_separate_batch writes)
 *   d_mix  : mixture, segment i at d_mix + offsets[i]
 *   K in {2,3}, K <= N <= 8 (the reference enumerates itertools.combinations(range(N), K) for any N >= K,
 *   overlap3_core.py:81, evaluate_with_sources.py:352-355 / :421; N > 3 costs one extra pass over the samples per
 *   further group of three estimates and needs b200sep_pit_workspace_bytes_n(lens, n_seg, N) of scratch).
 * Outputs (device):
 *   d_pair  [n_seg][3][W] fp64 : W = max(N, 3) + 1 (N <= 3: [3][4]); pair[i][j] = SI-SDR(ref_i, est_j) dB for j<N,
 *                                pair[i][W-1] = SI-SDR(ref_i, mix); unused entries NaN
 *   d_sisdr [n_seg] fp64       : best mean SI-SDR over permutations
 *   d_sisdri[n_seg] fp64       : best - mean_i SI-SDR(ref_i, mix)
 *   d_perm  [n_seg][4] int32   : perm[i] = estimate index assigned to reference i (i<K);
 *                                perm[3] = "swapped" flag of the K=2 EVALSRC variant, else 0
 *   d_gap   [n_seg] fp64 or NULL: best minus second-best mean SI-SDR over the distinct assignments (dB; +inf when
 *                                there is only one): the margin by which the permutation choice is decided -- the
 *                                kernel's fp64 moments differ from the reference's fp32 reductions by up to ~2e-3 dB,
 *                                so "identical permutation" is only defined where the gap exceeds that
 */
int b200sep_pit_sisdr_batch(const float* d_refs, const float* d_ests, const float* d_mix,
                            const int64_t* offsets, const int32_t* lens, int32_t n_seg,
                            int32_t K, int32_t N, int32_t variant, double* d_pair,
                            double* d_sisdr, double* d_sisdri, int32_t* d_perm, double* d_gap, void* d_ws,
                            size_t ws_bytes, void* stream);

/* ---- either side of the separation path: resampling and GT overlap masks ------------ */
#define B200SEP_RESAMPLE_LINEAR 0 /* Separator._ensure_sr, src/osd/separation.py:91-103 (np.interp, fp64 -> fp32)     */
#define B200SEP_RESAMPLE_SINC 1   /* torchaudio.functional.resample defaults: scripts/osd/overlap3_core.py:136-143,    */
                                  /* scripts/osd/evaluate_with_sources.py:163-178                                      */
/*
 * Output lengths of a resample of `lens[i]` samples from sr_in to sr_out.  LINEAR: int(round(len * sr_out / sr_in))
 * with the reference's pass-through rules (same rate, len <= 1 or target <= 1 -> unchanged, separation.py:92-100);
 * SINC: ceil(new * len / orig) on the gcd-reduced rates (len <= 1 -> unchanged, overlap3_core.py:139).
 */
int b200sep_resample_out_lens(const int32_t* lens, int32_t n_seg, int32_t sr_in, int32_t sr_out, int32_t method,
                              int32_t* out_lens);
/*
 * Resamples a tightly packed ragged batch of fp32 signals on the device (`d_in`: signal i at sum(lens[0..i-1]);
 * `d_out`: signal i at sum(out_lens[0..i-1])).  LINEAR replays np.interp operation by operation in fp64 and is
 * bit-identical to the reference for fp32 inputs; SINC applies the same polyphase windowed-sinc filter bank as
 * torchaudio (fp32, agreement to fp32 rounding).
 */
int b200sep_resample_batch(const float* d_in, const int32_t* lens, int32_t n_seg, int32_t sr_in, int32_t sr_out,
                           int32_t method, float* d_out, void* stream);
/*
 * frame_rms (scripts/osd/evaluate_with_sources.py:181-197), batched: frames of win_s samples every hop_s samples,
 * rms = sqrt(mean(frame^2) + 1e-12); signal i yields len(range(0, max(len - win_s + 1, 1), hop_s)) values, packed.
 */
int b200sep_frame_rms_out_lens(const int32_t* lens, int32_t n_seg, int32_t win_s, int32_t hop_s, int32_t* out_lens);
int b200sep_frame_rms_batch(const float* d_wav, const int32_t* lens, int32_t n_seg, int32_t win_s, int32_t hop_s,
                            float* d_rms, void* stream);
/*
 * build_gt_overlap_mask (scripts/osd/evaluate_with_sources.py:226-241), batched over pairs of RMS tracks with the
 * same packed layout: peak = max(max(rms1), max(rms2), 1e-9); mask = (rms1 > peak * thr) & (rms2 > peak * thr).
 */
int b200sep_gt_overlap_mask_batch(const float* d_rms1, const float* d_rms2, const int32_t* n_frames, int32_t n_seg,
                                  float thr_ratio, uint8_t* d_mask, void* stream);

/*
 * Development/validation hook: raw GEMM  C[M,N] = A * B^T  through the same tcgen05/TMA
 * kernel the model uses (A given row-major [M,K] fp32 on device, B [N,K] fp32 on device,
 * both rounded to fp16; C fp32 row-major).  M multiple of 128, K multiple of 64,
 * N multiple of `bn` (128 or 256).  Used by tests to check the tensor-core path in isolation.
 */
int b200sep_debug_gemm(const float* d_a, const float* d_b, float* d_c, int32_t M, int32_t N,
                       int32_t K, int32_t bn, int32_t use_tc, void* stream);

/*
 * Validation hook of the MMA form the block kernel K_A uses: A operand resident in tensor memory, B from the shared
 * memory of one CTA (pair = 0: M = N = 128) or of a CTA pair (pair = 1: tcgen05 cta_group::2, M = N = 256).
 * out[c][f] = sum_k W[c][k] * H[f][k], K = 64; W [256][64], H [256][64], out [256][256] fp32 DEVICE arrays.
 */
int b200sep_debug_ts_probe(const float* d_w, const float* d_h, float* d_out, int32_t pair, void* stream);

/*
 * Per-launch timing for the roofline report (bench.py): when enabled, b200sep_separate_batch
 * brackets every kernel launch with CUDA events on the caller's stream; b200sep_profile_read
 * returns the durations (ms) of the last call in launch order: encoder, bottleneck GEMM, then per
 * TCN block (in-conv GEMM, depthwise conv, res/skip GEMM), mask GEMM, decoder.
 */
int b200sep_profile_enable(b200sep_model* m, int on);
int b200sep_profile_read(b200sep_model* m, float* ms_out, int32_t cap, int32_t* n_out);

/*
 * Per-launch timing of the PIT entry point (bench.py, tools/pit_target.py): after b200sep_pit_profile_enable(1) every
 * b200sep_pit_sisdr_batch call brackets its two kernels with CUDA events on the caller's stream;
 * b200sep_pit_profile_read waits for the last call on the current device and returns ms_out[0] = moments kernel,
 * ms_out[1] = finalize kernel.
 */
int b200sep_pit_profile_enable(int on);
int b200sep_pit_profile_read(float* ms_out);

/* Test hooks: stop the forward after the k-th kernel launch (0 = run everything), and report the
 * byte offsets of the intermediate activations inside the workspace:
 * out[0..10] = enc, x, xa, skip, skipa, h1, h2, masked, n_tiles, total, g. */
int b200sep_debug_set_stop(b200sep_model* m, int stop_after);
int b200sep_debug_layout(const b200sep_model* m, const int32_t* lens, int32_t n_seg, uint64_t* out);

/* Debug: register (nullptr: remove) a device buffer of n_sm * 8 * 128 int64 that the fused kernels stamp
 * with clock64 values per (CTA, tile visit, slot); used by tests/timeline.py to see the pipeline. */
int b200sep_debug_timeline(void* d_buf);

/* Launch sequence of the model: 0 = encoder, bottleneck GEMM, per TCN block (in-conv GEMM, depthwise
 * conv, res/skip GEMM), mask GEMM, decoder; 1 = encoder, bottleneck GEMM, per TCN block (apply + in-conv
 * GEMM, depthwise + res/skip GEMM), final skip, mask GEMM, decoder (fused fp16 tensor-core mode). */
int b200sep_pipeline(const b200sep_model* m);

#ifdef __cplusplus
}
#endif
#endif /* B200SEP_H_ */

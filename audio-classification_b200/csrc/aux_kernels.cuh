// Bandwidth kernels either side of the separation path:
//   * resample_linear_kernel  -- Separator._ensure_sr (src/osd/separation.py:91-103): np.interp over np.linspace,
//                                fp64 arithmetic -> fp32, replayed operation by operation (no FMA contraction) so the
//                                result is bit-identical to NumPy's;
//   * resample_sinc_kernel    -- torchaudio.functional.resample with its defaults, the resampler the reference's
//                                loaders call (scripts/osd/overlap3_core.py:136-143, evaluate_with_sources.py:163-178):
//                                polyphase windowed-sinc FIR, filter bank built on the host in fp64;
//   * frame_rms_kernel        -- frame_rms (scripts/osd/evaluate_with_sources.py:181-197);
//   * gt_overlap_mask_kernel  -- build_gt_overlap_mask (scripts/osd/evaluate_with_sources.py:226-241).
// All are HBM-bound one-pass kernels over a ragged batch of independent signals.
#pragma once
#include "common.cuh"

namespace b200sep {

struct RaggedIO {
  const int64_t* in_off;   // [n] element offset of signal i in the input buffer
  const int32_t* in_len;   // [n]
  const int64_t* out_off;  // [n] element offset of signal i in the output buffer
  const int32_t* out_len;  // [n]
};

// out[j] = np.interp(np.linspace(0, n_in - 1, n_out)[j], arange(n_in), x)  (fp64) -> fp32
// np.linspace: step = (n_in - 1) / (n_out - 1); pos_j = j * step; pos_{n_out-1} = n_in - 1 exactly.
// np.interp (compiled arr_interp): index i = floor(pos) (xp = 0, 1, 2, ...); i == n_in - 1 or pos == i -> x[i];
// else slope = (x[i+1] - x[i]) / (xp[i+1] - xp[i]); result = slope * (pos - xp[i]) + x[i], every operation rounded
// separately in fp64.
__global__ void __launch_bounds__(256) resample_linear_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                              RaggedIO io) {
  const int seg = blockIdx.y;
  const int n_in = io.in_len[seg], n_out = io.out_len[seg];
  const float* __restrict__ x = in + io.in_off[seg];
  float* __restrict__ y = out + io.out_off[seg];
  const double step = (n_out > 1) ? __ddiv_rn((double)(n_in - 1), (double)(n_out - 1)) : 0.0;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n_out; j += gridDim.x * blockDim.x) {
    double pos = __dmul_rn((double)j, step);
    if (j == n_out - 1) pos = (double)(n_in - 1);
    int i = (int)floor(pos);
    i = max(0, min(i, n_in - 1));
    double r;
    if (i >= n_in - 1 || pos == (double)i) {
      r = (double)x[min(i, n_in - 1)];
    } else {
      const double y0 = (double)x[i], y1 = (double)x[i + 1];
      const double slope = __ddiv_rn(__dsub_rn(y1, y0), 1.0);
      r = __dadd_rn(__dmul_rn(slope, __dsub_rn(pos, (double)i)), y0);
    }
    y[j] = __double2float_rn(r);
  }
}

// Polyphase windowed-sinc resampler (torchaudio _apply_sinc_resample_kernel):
//   padded = pad(x, (width, width + orig)); out[k * new + p] = sum_t kern[p][t] * padded[k * orig + t],
//   t in [0, 2 width + orig), cropped to ceil(new * n_in / orig) samples.  `kern` is [new][kw] fp32.
struct SincArgs {
  const float* kern;  // [new_freq][kw]
  int orig, neu, width, kw;
};
__global__ void __launch_bounds__(256) resample_sinc_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                            RaggedIO io, SincArgs a) {
  extern __shared__ float s_kern[];  // the whole filter bank when it fits (host decides), else read through L1
  const bool in_smem = (a.neu * a.kw * 4 <= 48 * 1024);
  if (in_smem) {
    for (int i = threadIdx.x; i < a.neu * a.kw; i += blockDim.x) s_kern[i] = __ldg(a.kern + i);
    __syncthreads();
  }
  const float* __restrict__ kern = in_smem ? s_kern : a.kern;
  const int seg = blockIdx.y;
  const int n_in = io.in_len[seg], n_out = io.out_len[seg];
  const float* __restrict__ x = in + io.in_off[seg];
  float* __restrict__ y = out + io.out_off[seg];
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n_out; j += gridDim.x * blockDim.x) {
    const int k = j / a.neu, p = j - k * a.neu;
    const float* __restrict__ kp = kern + (size_t)p * a.kw;
    const int base = k * a.orig - a.width;  // input index of tap 0 (negative / beyond the end = zero padding)
    float acc = 0.f;
    const int t_lo = max(0, -base), t_hi = min(a.kw, n_in - base);
    for (int t = t_lo; t < t_hi; ++t) acc = fmaf(kp[t], __ldg(x + base + t), acc);
    y[j] = acc;
  }
}

// frame_rms: frames start at 0, hop_s, ... while start < max(len - win_s + 1, 1); frame = x[start : start + win_s]
// (shorter than win_s only when len < win_s); rms = sqrt(mean(frame^2) + 1e-12).  The reference sums in fp32
// (NumPy pairwise); here the mean of squares is accumulated in fp64 and rounded once to fp32, then the two fp32
// operations (+ 1e-12f, sqrt) follow the reference's types.  One warp per frame.
struct RmsArgs {
  int win_s, hop_s;
};
__global__ void __launch_bounds__(256) frame_rms_kernel(const float* __restrict__ in, float* __restrict__ out, RaggedIO io,
                                                        RmsArgs a) {
  const int seg = blockIdx.y;
  const int n = io.in_len[seg], n_fr = io.out_len[seg];
  const float* __restrict__ x = in + io.in_off[seg];
  float* __restrict__ y = out + io.out_off[seg];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int f = blockIdx.x * wpb + warp; f < n_fr; f += gridDim.x * wpb) {
    const int start = f * a.hop_s;
    const int cnt = max(0, min(a.win_s, n - start));
    double acc = 0.0;
    for (int t = lane; t < cnt; t += 32) {
      const double v = (double)__ldg(x + start + t);
      acc = fma(v, v, acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) {
      float r = 0.f;
      if (cnt > 0) {
        const float m = __double2float_rn(acc / (double)cnt);
        r = __fsqrt_rn(__fadd_rn(m, 1e-12f));
      }
      y[f] = r;
    }
  }
}

// build_gt_overlap_mask: peak = max(max(rms1), max(rms2), 1e-9); mask = (rms1 > peak * thr) & (rms2 > peak * thr)
// with fp32 products, as NumPy evaluates it.  One CTA per signal pair; rms1 / rms2 share offsets and lengths.
__global__ void __launch_bounds__(256) gt_overlap_mask_kernel(const float* __restrict__ rms1, const float* __restrict__ rms2,
                                                              const int64_t* __restrict__ off, const int32_t* __restrict__ len,
                                                              float thr, uint8_t* __restrict__ mask) {
  __shared__ float s_max[8];
  const int seg = blockIdx.x;
  const int n = len[seg];
  const float* __restrict__ a = rms1 + off[seg];
  const float* __restrict__ b = rms2 + off[seg];
  uint8_t* __restrict__ m = mask + off[seg];
  float mx = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) mx = fmaxf(mx, fmaxf(a[i], b[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = s_max[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) mx = fmaxf(mx, s_max[w]);
  const float peak = fmaxf(mx, 1e-9f);
  const float lim = __fmul_rn(peak, thr);
  for (int i = threadIdx.x; i < n; i += blockDim.x) m[i] = (a[i] > lim && b[i] > lim) ? 1 : 0;
}

}  // namespace b200sep

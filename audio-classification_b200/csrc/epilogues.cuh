// GEMM epilogues of the 1x1 convolutions.  The same functors are used by the tcgen05 kernel (one
// thread = one frame = one TMEM lane, 32 consecutive output channels per call) and by the CUDA-core
// kernel (4 consecutive output channels per call).  All global accesses are of the form
// buf[m_tile][n][r] with r = the thread's frame -> coalesced across the lanes of a warp.
//
// gLN folding: for y = gamma*(h-mean)*rstd + beta followed by W*y + b,
//   W*y + b = rstd * ( (W.gamma) * h  -  mean * rowsum(W.gamma) ) + (W*beta + b)
// so the GEMM runs on the RAW h with pre-scaled weights Wg = W.gamma; `rs` = rowsum of the
// operand-rounded Wg and `cb` = W*beta + b are precomputed at model creation; the two per-segment
// scalars (mean, rstd) are reduced from the producer's partial sums at the start of every tile.
#pragma once
#include "common.cuh"

namespace b200sep {

struct EpiState {
  float mean, rstd;  // gLN statistics of the INPUT activation of this GEMM (if folded)
  float s, q;        // running (sum, sumsq) of this thread's outputs (if the epilogue emits stats)
  int F, t0;         // segment frame count, first segment-relative frame of the tile
};

__device__ __forceinline__ void epi_tile_ctx(EpiState& st, int m_tile, const Tables& tb, int& seg,
                                             int& tile0, int& ntiles) {
  seg = tb.tile_seg[m_tile];
  tile0 = tb.seg_tile0[seg];
  ntiles = tb.seg_ntiles[seg];
  st.F = tb.seg_frames[seg];
  st.t0 = (m_tile - tile0) * TILE_M;
  st.s = 0.f;
  st.q = 0.f;
}

// Writes this warp's (sum, sumsq) partial: part[m_tile * slots + slot].
__device__ __forceinline__ void epi_write_partial(EpiState& st, float2* part, int m_tile, int slots,
                                                  int slot) {
  const float s = warp_sum(st.s), q = warp_sum(st.q);
  if ((threadIdx.x & 31) == 0) part[(size_t)m_tile * slots + slot] = make_float2(s, q);
}

// -------------------------------------------------------------------------------------------------
// Bottleneck: x = conv1x1(gLN_N(tf_rep))  (masker.bottleneck).  Writes the fp32
// residual stream x and (fp16 modes) its operand-precision shadow xa.
template <typename T>
struct EpiBottleneck {
  const float2* part_in; int slots_in;
  const float* rs; const float* cb;  // [128]
  float* x;                          // [nt][128][128] fp32
  T* xa;                             // [nt][128][128] operand copy, or nullptr when T == float
  static constexpr bool kStats = false;

  __device__ __forceinline__ void tile_begin(EpiState& st, int m_tile, const Tables& tb) const {
    int seg, tile0, ntiles;
    epi_tile_ctx(st, m_tile, tb, seg, tile0, ntiles);
    warp_seg_stats(part_in, tile0, ntiles, slots_in, (double)C_ENC * (double)st.F, st.mean, st.rstd);
  }
  template <int W>
  __device__ __forceinline__ void apply(EpiState& st, int m_tile, int r, int n0, const float* acc) const {
    const bool valid = (st.t0 + r) < st.F;
    const float mr = st.mean;
#pragma unroll
    for (int j = 0; j < W; ++j) {
      const int n = n0 + j;
      float v = fmaf(st.rstd, fmaf(-mr, __ldg(rs + n), acc[j]), __ldg(cb + n));
      if (!valid) v = 0.f;
      const size_t idx = tt_index(m_tile, C_BN, n0, r) + (size_t)j * TILE_M;
      x[idx] = v;
      if (xa != nullptr) xa[idx] = from_f32<T>(v);
    }
  }
  __device__ __forceinline__ void tile_end(EpiState&, int, int, int) const {}
};

// -------------------------------------------------------------------------------------------------
// Block in-conv: h1 = PReLU(conv1x1(x) + b)   (shared_block.0/.1) + gLN1 partial statistics.
template <typename T>
struct EpiIn {
  const float* bias;  // [512]
  float alpha;
  T* h1;              // [nt][512][128]
  float2* part_out;   // [nt][slots]
  static constexpr bool kStats = true;

  __device__ __forceinline__ void tile_begin(EpiState& st, int m_tile, const Tables& tb) const {
    int seg, tile0, ntiles;
    epi_tile_ctx(st, m_tile, tb, seg, tile0, ntiles);
  }
  template <int W>
  __device__ __forceinline__ void apply(EpiState& st, int m_tile, int r, int n0, const float* acc) const {
    const bool valid = (st.t0 + r) < st.F;
#pragma unroll
    for (int j = 0; j < W; ++j) {
      const int n = n0 + j;
      float v = acc[j] + __ldg(bias + n);
      v = (v > 0.f) ? v : alpha * v;
      if (!valid) v = 0.f;
      h1[tt_index(m_tile, C_HID, n0, r) + (size_t)j * TILE_M] = from_f32<T>(v);
      st.s += v;
      st.q = fmaf(v, v, st.q);
    }
  }
  __device__ __forceinline__ void tile_end(EpiState& st, int m_tile, int slots, int slot) const {
    epi_write_partial(st, part_out, m_tile, slots, slot);
  }
};

// -------------------------------------------------------------------------------------------------
// Block out-convs: [res | skip] = conv1x1(gLN2(h2)) fused as one GEMM with N = 256 (128 for the last
// block, whose residual output is never read -- SURVEY.md A.2).  x += res ; skip_sum += skip.
template <typename T>
struct EpiResSkip {
  const float2* part_in; int slots_in;
  const float* rs; const float* cb;  // [256] (or [128] when !has_res)
  float* x;     // [nt][128][128] fp32 residual stream
  T* xa;        // operand shadow of x (nullptr when T == float)
  float* skip;  // [nt][128][128] fp32 skip accumulator
  T* skipa;     // PReLU(skip) in operand precision, written by the last block only
  float alpha_mask;
  int has_res, first, last;
  static constexpr bool kStats = false;

  __device__ __forceinline__ void tile_begin(EpiState& st, int m_tile, const Tables& tb) const {
    int seg, tile0, ntiles;
    epi_tile_ctx(st, m_tile, tb, seg, tile0, ntiles);
    warp_seg_stats(part_in, tile0, ntiles, slots_in, (double)C_HID * (double)st.F, st.mean, st.rstd);
  }
  template <int W>
  __device__ __forceinline__ void apply(EpiState& st, int m_tile, int r, int n0, const float* acc) const {
    const bool valid = (st.t0 + r) < st.F;
    const bool is_res = has_res && (n0 < C_BN);
    const int c0 = (has_res && !is_res) ? (n0 - C_BN) : n0;
    const size_t idx0 = tt_index(m_tile, C_BN, c0, r);
    // issue all loads of the chunk first: they are independent, and consecutive lanes (frames) hit
    // consecutive addresses, so W x 128 B per warp are in flight instead of one dependent load
    float old[W];
    const bool need_old = is_res || !first;
    const float* __restrict__ src = is_res ? x : skip;
    if (need_old) {
#pragma unroll
      for (int j = 0; j < W; ++j) old[j] = __ldcs(src + idx0 + (size_t)j * TILE_M);
    } else {
#pragma unroll
      for (int j = 0; j < W; ++j) old[j] = 0.f;
    }
    float out[W];
#pragma unroll
    for (int j = 0; j < W; ++j) {
      const int n = n0 + j;
      float v = fmaf(st.rstd, fmaf(-st.mean, __ldg(rs + n), acc[j]), __ldg(cb + n));
      if (!valid) v = 0.f;
      out[j] = old[j] + v;
    }
    if (is_res) {
#pragma unroll
      for (int j = 0; j < W; ++j) x[idx0 + (size_t)j * TILE_M] = out[j];
      if (xa != nullptr) {
#pragma unroll
        for (int j = 0; j < W; ++j) xa[idx0 + (size_t)j * TILE_M] = from_f32<T>(out[j]);
      }
    } else if (!last) {
#pragma unroll
      for (int j = 0; j < W; ++j) skip[idx0 + (size_t)j * TILE_M] = out[j];
    } else {
#pragma unroll
      for (int j = 0; j < W; ++j)
        skipa[idx0 + (size_t)j * TILE_M] = from_f32<T>((out[j] > 0.f) ? out[j] : alpha_mask * out[j]);
    }
  }
  __device__ __forceinline__ void tile_end(EpiState&, int, int, int) const {}
};

// -------------------------------------------------------------------------------------------------
// Mask net: masked = act(conv1x1(PReLU(skip_sum)) + b) * tf_rep   (mask_net + output_act + the
// `est_masks * tf_rep.unsqueeze(1)` of BaseEncoderMaskerDecoder.forward).  N = n_src * 512.
template <typename T, bool kFast>
struct EpiMask {
  const float* bias;  // [S*512]
  const T* enc;       // [nt][512][128]
  T* masked;          // [nt][S*512][128]
  int n_total;        // S*512
  int act;            // 0 sigmoid, 1 relu
  static constexpr bool kStats = false;

  __device__ __forceinline__ void tile_begin(EpiState& st, int m_tile, const Tables& tb) const {
    int seg, tile0, ntiles;
    epi_tile_ctx(st, m_tile, tb, seg, tile0, ntiles);
  }
  template <int W>
  __device__ __forceinline__ void apply(EpiState& st, int m_tile, int r, int n0, const float* acc) const {
    // a chunk never straddles a source boundary (W divides 512)
    const T* __restrict__ ep = enc + tt_index(m_tile, C_ENC, n0 & (C_ENC - 1), r);
    T* __restrict__ mp = masked + tt_index(m_tile, n_total, n0, r);
    float e[W];
#pragma unroll
    for (int j = 0; j < W; ++j) e[j] = to_f32(ep[(size_t)j * TILE_M]);
#pragma unroll
    for (int j = 0; j < W; ++j) {
      const float v = acc[j] + __ldg(bias + n0 + j);
      float m;
      if (act == 0) {
        if (kFast) m = __fdividef(1.f, 1.f + __expf(-v));
        else m = 1.f / (1.f + expf(-v));
      } else {
        m = fmaxf(v, 0.f);
      }
      mp[(size_t)j * TILE_M] = from_f32<T>(m * e[j]);
    }
  }
  __device__ __forceinline__ void tile_end(EpiState&, int, int, int) const {}
};

// Plain store (debug GEMM hook): C row-major [M][N] fp32.
struct EpiPlain {
  float* c; int N;
  static constexpr bool kStats = false;
  __device__ __forceinline__ void tile_begin(EpiState& st, int, const Tables&) const { st.s = st.q = 0.f; }
  template <int W>
  __device__ __forceinline__ void apply(EpiState&, int m_tile, int r, int n0, const float* acc) const {
#pragma unroll
    for (int j = 0; j < W; ++j) c[((size_t)m_tile * TILE_M + r) * N + n0 + j] = acc[j];
  }
  __device__ __forceinline__ void tile_end(EpiState&, int, int, int) const {}
};

}  // namespace b200sep

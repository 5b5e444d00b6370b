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
  double inv_count;  // 1 / (512 * F) of the tile's segment
  float s, q;        // running (sum, sumsq) of this thread's outputs (if the epilogue emits stats)
  int F, t0;         // segment frame count, first segment-relative frame of the tile
};

__device__ __forceinline__ void epi_tile_ctx(EpiState& st, int m_tile, const Tables& tb, int& seg,
                                             int& tile0, int& ntiles) {
  const TileInfo ti = load_tile_info(tb.tile_info + m_tile);
  seg = ti.seg;
  tile0 = ti.tile0;
  ntiles = ti.ntiles;
  st.F = ti.F;
  st.inv_count = ti.inv_count;
  st.t0 = ti.t0;
  st.s = 0.f;
  st.q = 0.f;
}

// Writes an already block-reduced (sum, sumsq) partial (single calling thread):
// part[m_tile * slots + slot].
__device__ __forceinline__ void epi_write_partial(EpiState& st, float2* part, int m_tile, int slots,
                                                  int slot) {
  part[(size_t)m_tile * slots + slot] = make_float2(st.s, st.q);
}

// 4 consecutive fp32 parameters with one 16-byte read-only load (n % 4 == 0).
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// Every epilogue provides
//   tile_begin(st, m_tile, tb)                 warp-collective: tile context (+ folded gLN statistics)
//   prefetch<W>(m_tile, r, n0, pre)            issue the global loads apply() will need (may be a no-op)
//   apply<W>(st, m_tile, r, n0, acc, pre)      consume W consecutive output channels of frame r
//   tile_end(st, m_tile, slots, slot)          single thread: write the block-reduced stat partial
// W is 32 (tcgen05 kernel) or 4 (CUDA-core kernel); n0 % W == 0.

// -------------------------------------------------------------------------------------------------
// Bottleneck: x = conv1x1(gLN_N(tf_rep))  (masker.bottleneck).  Writes the fp32
// residual stream x and (fp16 modes) its operand-precision shadow xa.
template <typename T>
struct EpiBottleneck {
  const float2* part_in; int slots_in;
  const float* rs; const float* cb;  // [128]
  float* x;                          // [nt][128][128] fp32
  T* xa;                             // [nt][128][128] operand copy, or nullptr when T == float
  float eps;                         // eps of the bottleneck gLN
  static constexpr bool kStats = false;

  __device__ __forceinline__ void tile_begin(EpiState& st, int m_tile, const Tables& tb) const {
    int seg, tile0, ntiles;
    epi_tile_ctx(st, m_tile, tb, seg, tile0, ntiles);
    warp_seg_stats(part_in, tile0, ntiles, slots_in, st.inv_count, eps, st.mean, st.rstd, tb.flags);
  }
  template <int W>
  __device__ __forceinline__ void prefetch(int, int, int, float (&)[W]) const {}
  template <int W>
  __device__ __forceinline__ void apply(EpiState& st, int m_tile, int r, int n0, const float* acc,
                                        const float (&)[W]) const {
    const float vm = ((st.t0 + r) < st.F) ? 1.f : 0.f;
    const float2 rstd2 = make_float2(st.rstd * vm, st.rstd * vm);
    const float2 nmr2 = make_float2(-st.mean * st.rstd * vm, -st.mean * st.rstd * vm);
    const float2 vm2 = make_float2(vm, vm);
    const size_t idx0 = tt_index(m_tile, C_BN, n0, r);
    float out[W];
#pragma unroll
    for (int q = 0; q < W / 4; ++q) {
      const float4 r4 = ldg4(rs + n0 + 4 * q), c4 = ldg4(cb + n0 + 4 * q);
      // out = vm * (rstd * (acc - mean * rs) + cb)
      const float2 t0 = fma2(make_float2(r4.x, r4.y), nmr2, mul2(make_float2(c4.x, c4.y), vm2));
      const float2 t1 = fma2(make_float2(r4.z, r4.w), nmr2, mul2(make_float2(c4.z, c4.w), vm2));
      const float2 o0 = fma2(make_float2(acc[4 * q + 0], acc[4 * q + 1]), rstd2, t0);
      const float2 o1 = fma2(make_float2(acc[4 * q + 2], acc[4 * q + 3]), rstd2, t1);
      out[4 * q + 0] = o0.x; out[4 * q + 1] = o0.y; out[4 * q + 2] = o1.x; out[4 * q + 3] = o1.y;
    }
    if (x != nullptr) {  // (nullptr: the fused pipeline keeps the residual stream in fp16 only)
#pragma unroll
      for (int j = 0; j < W; ++j) x[idx0 + (size_t)j * TILE_M] = out[j];
    }
    if (xa != nullptr) {
#pragma unroll
      for (int j = 0; j < W; ++j) xa[idx0 + (size_t)j * TILE_M] = from_f32<T>(out[j]);
    }
  }
  __device__ __forceinline__ void tile_end(EpiState&, int, int, int) const {}
};

// -------------------------------------------------------------------------------------------------
// Block in-conv: h1 = PReLU(conv1x1(x) + b)   (shared_block.0/.1) + gLN1 partial statistics.
template <typename T, bool A01>  // A01: 0 <= alpha <= 1, PReLU(v) == max(v, alpha*v)
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
  __device__ __forceinline__ void prefetch(int, int, int, float (&)[W]) const {}
  // Frames beyond the segment end keep whatever finite value falls out (their x is 0, so h1 =
  // PReLU(bias)); nothing downstream reads them: the depthwise conv masks its taps by frame index.
  // Only the statistics must exclude them (multiply by the 0/1 frame mask).
  template <int W>
  __device__ __forceinline__ void apply(EpiState& st, int m_tile, int r, int n0, const float* acc,
                                        const float (&)[W]) const {
    const float vm = ((st.t0 + r) < st.F) ? 1.f : 0.f;
    const float2 vm2 = make_float2(vm, vm);
    T* __restrict__ dst = h1 + tt_index(m_tile, C_HID, n0, r);
    float2 s2 = make_float2(0.f, 0.f), q2 = make_float2(0.f, 0.f);
#pragma unroll
    for (int q = 0; q < W / 4; ++q) {
      const float4 b4 = ldg4(bias + n0 + 4 * q);
      float2 v0 = add2(make_float2(acc[4 * q + 0], acc[4 * q + 1]), make_float2(b4.x, b4.y));
      float2 v1 = add2(make_float2(acc[4 * q + 2], acc[4 * q + 3]), make_float2(b4.z, b4.w));
      v0 = prelu2<A01>(v0, alpha);
      v1 = prelu2<A01>(v1, alpha);
      dst[(size_t)(4 * q + 0) * TILE_M] = from_f32<T>(v0.x);
      dst[(size_t)(4 * q + 1) * TILE_M] = from_f32<T>(v0.y);
      dst[(size_t)(4 * q + 2) * TILE_M] = from_f32<T>(v1.x);
      dst[(size_t)(4 * q + 3) * TILE_M] = from_f32<T>(v1.y);
      const float2 m0 = mul2(v0, vm2), m1 = mul2(v1, vm2);
      s2 = add2(s2, add2(m0, m1));
      q2 = fma2(m0, v0, fma2(m1, v1, q2));
    }
    st.s += s2.x + s2.y;
    st.q += q2.x + q2.y;
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
  float eps;    // eps of this block's gLN2
  static constexpr bool kStats = false;

  __device__ __forceinline__ void tile_begin(EpiState& st, int m_tile, const Tables& tb) const {
    int seg, tile0, ntiles;
    epi_tile_ctx(st, m_tile, tb, seg, tile0, ntiles);
    warp_seg_stats(part_in, tile0, ntiles, slots_in, st.inv_count, eps, st.mean, st.rstd, tb.flags);
  }
  // old x / skip values: independent loads, consecutive lanes (frames) hit consecutive addresses
  template <int W>
  __device__ __forceinline__ void prefetch(int m_tile, int r, int n0, float (&pre)[W]) const {
    const bool is_res = has_res && (n0 < C_BN);
    const int c0 = (has_res && !is_res) ? (n0 - C_BN) : n0;
    if (is_res || !first) {
      const float* __restrict__ src = (is_res ? x : skip) + tt_index(m_tile, C_BN, c0, r);
#pragma unroll
      for (int j = 0; j < W; ++j) pre[j] = __ldcs(src + (size_t)j * TILE_M);
    } else {
#pragma unroll
      for (int j = 0; j < W; ++j) pre[j] = 0.f;
    }
  }
  template <int W>
  __device__ __forceinline__ void apply(EpiState& st, int m_tile, int r, int n0, const float* acc,
                                        const float (&pre)[W]) const {
    const float vm = ((st.t0 + r) < st.F) ? 1.f : 0.f;
    const bool is_res = has_res && (n0 < C_BN);
    const int c0 = (has_res && !is_res) ? (n0 - C_BN) : n0;
    const size_t idx0 = tt_index(m_tile, C_BN, c0, r);
    const float2 rstd2 = make_float2(st.rstd * vm, st.rstd * vm);
    const float2 nmr2 = make_float2(-st.mean * st.rstd * vm, -st.mean * st.rstd * vm);
    const float2 vm2 = make_float2(vm, vm);
    float out[W];
#pragma unroll
    for (int q = 0; q < W / 4; ++q) {
      const float4 r4 = ldg4(rs + n0 + 4 * q), c4 = ldg4(cb + n0 + 4 * q);
      // out = old + vm * (rstd * (acc - mean * rs) + cb)
      const float2 t0 = fma2(make_float2(r4.x, r4.y), nmr2, fma2(make_float2(c4.x, c4.y), vm2, make_float2(pre[4 * q + 0], pre[4 * q + 1])));
      const float2 t1 = fma2(make_float2(r4.z, r4.w), nmr2, fma2(make_float2(c4.z, c4.w), vm2, make_float2(pre[4 * q + 2], pre[4 * q + 3])));
      const float2 o0 = fma2(make_float2(acc[4 * q + 0], acc[4 * q + 1]), rstd2, t0);
      const float2 o1 = fma2(make_float2(acc[4 * q + 2], acc[4 * q + 3]), rstd2, t1);
      out[4 * q + 0] = o0.x; out[4 * q + 1] = o0.y; out[4 * q + 2] = o1.x; out[4 * q + 3] = o1.y;
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
        skipa[idx0 + (size_t)j * TILE_M] = from_f32<T>(fmaxf(out[j], 0.f) + alpha_mask * fminf(out[j], 0.f));
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
  // encoder output of the same (frame, filter); a chunk never straddles a source boundary
  template <int W>
  __device__ __forceinline__ void prefetch(int m_tile, int r, int n0, float (&pre)[W]) const {
    const T* __restrict__ ep = enc + tt_index(m_tile, C_ENC, n0 & (C_ENC - 1), r);
#pragma unroll
    for (int j = 0; j < W; ++j) pre[j] = to_f32(ep[(size_t)j * TILE_M]);
  }
  template <int W>
  __device__ __forceinline__ void apply(EpiState&, int m_tile, int r, int n0, const float* acc,
                                        const float (&pre)[W]) const {
    T* __restrict__ mp = masked + tt_index(m_tile, n_total, n0, r);
    float v[W];
#pragma unroll
    for (int q = 0; q < W / 4; ++q) {
      const float4 b4 = ldg4(bias + n0 + 4 * q);
      const float2 v0 = add2(make_float2(acc[4 * q + 0], acc[4 * q + 1]), make_float2(b4.x, b4.y));
      const float2 v1 = add2(make_float2(acc[4 * q + 2], acc[4 * q + 3]), make_float2(b4.z, b4.w));
      v[4 * q + 0] = v0.x; v[4 * q + 1] = v0.y; v[4 * q + 2] = v1.x; v[4 * q + 3] = v1.y;
    }
    if (act == 0) {
      if (kFast) {
        // sigmoid(v) = 0.5 * tanh(0.5 v) + 0.5: one MUFU op per element (the exp + reciprocal form
        // needs two and makes this epilogue MUFU-bound); tanh.approx abs. error ~5e-4 (~ fp16 ulp)
        const float2 h2 = make_float2(0.5f, 0.5f);
#pragma unroll
        for (int j = 0; j < W; j += 2) {
          const float2 a = mul2(make_float2(v[j], v[j + 1]), h2);
          float t0, t1;
          asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(a.x));
          asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(a.y));
          const float2 m = mul2(fma2(make_float2(t0, t1), h2, h2), make_float2(pre[j], pre[j + 1]));
          mp[(size_t)j * TILE_M] = from_f32<T>(m.x);
          mp[(size_t)(j + 1) * TILE_M] = from_f32<T>(m.y);
        }
      } else {
#pragma unroll
        for (int j = 0; j < W; ++j) mp[(size_t)j * TILE_M] = from_f32<T>(pre[j] / (1.f + expf(-v[j])));
      }
    } else {
#pragma unroll
      for (int j = 0; j < W; ++j) mp[(size_t)j * TILE_M] = from_f32<T>(fmaxf(v[j], 0.f) * pre[j]);
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
  __device__ __forceinline__ void prefetch(int, int, int, float (&)[W]) const {}
  template <int W>
  __device__ __forceinline__ void apply(EpiState&, int m_tile, int r, int n0, const float* acc,
                                        const float (&)[W]) const {
#pragma unroll
    for (int j = 0; j < W; ++j) c[((size_t)m_tile * TILE_M + r) * N + n0 + j] = acc[j];
  }
  __device__ __forceinline__ void tile_end(EpiState&, int, int, int) const {}
};

}  // namespace b200sep

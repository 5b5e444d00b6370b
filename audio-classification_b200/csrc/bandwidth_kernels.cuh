// HBM-bound kernels of the Conv-TasNet forward: learned encoder (framing + 512xL filterbank),
// dilated depthwise conv fused with gLN-apply + PReLU + next-gLN statistics, and the transposed-
// conv decoder with overlap-add.  All operate on the tile-transposed layout of common.cuh.
#pragma once
#include "common.cuh"

namespace b200sep {

// =================================================================================================
// Encoder: tf_rep[c][f] = sum_l W[c][l] * wav[f*stride + l]     (asteroid Encoder(FreeFB), no bias,
// no activation).  stride == L/2.  grid = (n_tiles, 2 channel halves), 256 threads:
// thread = (frame r, channel quarter); the frame's L samples live in registers, the filter taps are
// broadcast from shared memory; stores are coalesced along frames.  Also emits the (sum, sumsq)
// partials of the bottleneck gLN (slots = 2 per m-tile).
// =================================================================================================
template <typename T, int L>
__global__ void __launch_bounds__(256) encoder_kernel(const float* __restrict__ mix,
                                                      const float* __restrict__ w_enc,  // [512][L]
                                                      T* __restrict__ encT,             // [nt][512][128]
                                                      float2* __restrict__ part,        // [nt][2]
                                                      Tables tb) {
  constexpr int S = L / 2;
  constexpr int CH_PER_CTA = 256;
  pdl_enter();
  __shared__ float4 s_w[CH_PER_CTA * L / 4];
  __shared__ float4 s_x4[(TILE_M * S + S) / 4 + 1];
  __shared__ float s_red[16];
  float* s_x = reinterpret_cast<float*>(s_x4);

  const int m_tile = blockIdx.x;
  const int chalf = blockIdx.y;
  const int tid = threadIdx.x;
  const TileInfo ti = load_tile_info(tb.tile_info + m_tile);
  const int seg = ti.seg, t0 = ti.t0, F = ti.F;
  const int T_len = tb.seg_len[seg];
  const float* x = mix + tb.seg_in_off[seg];

  // stage filter taps of this CTA's 256 channels and the tile's samples
  const float4* wg = reinterpret_cast<const float4*>(w_enc + (size_t)chalf * CH_PER_CTA * L);
  for (int i = tid; i < CH_PER_CTA * L / 4; i += 256) s_w[i] = wg[i];
  const int s0 = t0 * S;
  for (int i = tid; i < TILE_M * S + S; i += 256) {
    const int gi = s0 + i;
    s_x[i] = (gi < T_len) ? x[gi] : 0.f;
  }
  __syncthreads();

  const int r = tid & (TILE_M - 1);
  const int cq = tid >> 7;  // 0..1 -> 128 channels each
  const bool valid = (t0 + r) < F;
  float xv[L];
#pragma unroll
  for (int q = 0; q < L / 4; ++q) {
    const float4 v = s_x4[(r * S) / 4 + q];
    xv[4 * q + 0] = v.x; xv[4 * q + 1] = v.y; xv[4 * q + 2] = v.z; xv[4 * q + 3] = v.w;
  }
  float s_acc = 0.f, q_acc = 0.f;
  const int c_base = cq * 128;
  T* dst = encT + tt_index(m_tile, C_ENC, chalf * CH_PER_CTA + c_base, r);
#pragma unroll 4
  for (int j = 0; j < 128; ++j) {
    const float4* w4 = s_w + (c_base + j) * (L / 4);
    float acc = 0.f;
#pragma unroll
    for (int q = 0; q < L / 4; ++q) {
      const float4 w = w4[q];
      acc = fmaf(w.x, xv[4 * q + 0], acc);
      acc = fmaf(w.y, xv[4 * q + 1], acc);
      acc = fmaf(w.z, xv[4 * q + 2], acc);
      acc = fmaf(w.w, xv[4 * q + 3], acc);
    }
    if (!valid) acc = 0.f;
    dst[(size_t)j * TILE_M] = from_f32<T>(acc);
    s_acc += acc;
    q_acc = fmaf(acc, acc, q_acc);
  }
  s_acc = warp_sum(s_acc);
  q_acc = warp_sum(q_acc);
  const int warp = tid >> 5, lane = tid & 31;
  if (lane == 0) { s_red[warp] = s_acc; s_red[8 + warp] = q_acc; }
  __syncthreads();
  if (tid == 0) {
    float s = 0.f, q = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) { s += s_red[w]; q += s_red[8 + w]; }
    part[(size_t)m_tile * 2 + chalf] = make_float2(s, q);
  }
}

// =================================================================================================
// Depthwise dilated conv stage of a Conv1DBlock (ops 5-7 of section 2.2):
//   y   = gLN1(h1)            (gamma1/beta1, statistics from the producing GEMM's partials)
//   u   = dwconv_k3(y; dilation d, zero padding d) + bias
//   h2  = PReLU(u)            and (sum, sumsq) partials of h2 for gLN2 (slots = DW_SLOTS per m-tile)
// The zero padding applies to y (the normalised signal) at the TRUE segment ends, so taps that
// fall outside [0, F_i) contribute nothing -- not even beta.
//
// gLN1 is folded into the taps: with a = gamma*rstd, b = beta - a*mean,
//   u = bias + sum_k valid_k * (w_k*a * h_k + w_k*b)
// and for interior frames (all three taps inside the segment) u = c + wa0*h[t-d] + wa1*h[t] + wa2*h[t+d]
// -- 3 FMAs per output.  The per-channel constants are computed once per CTA into shared memory.
//
// grid = (512 / DW_CH channel groups, n_tiles), 256 threads.  A thread owns 8 consecutive frames
// (one 16-byte vector of fp16) of one channel per step; a warp covers 2 channels x 128 frames per
// step.  Taps at +-d are aligned vector loads from the shifted position when d % 8 == 0, and
// lane shuffles (+ one edge vector from the neighbouring m-tile) when d is 1, 2 or 4.
// =================================================================================================
struct DwParams {
  const float* gamma1;  // [512]
  const float* beta1;   // [512]
  const float* w;       // [512][3]
  const float* bias;    // [512]
  float alpha2;         // PReLU slope after the depthwise conv
  int dilation;
  float eps1;           // eps of gLN1
};

#ifndef B200SEP_DW_CH
#define B200SEP_DW_CH 256
#endif
#ifndef B200SEP_DW_UNROLL
#define B200SEP_DW_UNROLL 2
#endif
#ifndef B200SEP_DW_MINB
#define B200SEP_DW_MINB 4
#endif
constexpr int DW_CH = B200SEP_DW_CH;    // channels per CTA
constexpr int DW_SLOTS = C_HID / DW_CH; // stat partial slots per m-tile

// 8 consecutive frames of one channel
template <typename T> struct Vec8;
template <> struct Vec8<__half> {
  uint4 v;
  __device__ __forceinline__ void load(const __half* p) { v = *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ void store(__half* p) const { *reinterpret_cast<uint4*>(p) = v; }
  __device__ __forceinline__ void zero() { v = make_uint4(0, 0, 0, 0); }
  __device__ __forceinline__ void get(float (&o)[8]) const {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
      o[2 * i] = f.x; o[2 * i + 1] = f.y;
    }
  }
  __device__ __forceinline__ void set(const float (&o)[8]) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const __half2 h = __floats2half2_rn(o[2 * i], o[2 * i + 1]);
      w[i] = *reinterpret_cast<const uint32_t*>(&h);
    }
    v = make_uint4(w[0], w[1], w[2], w[3]);
  }
  __device__ __forceinline__ Vec8 shfl_up16() const {
    Vec8 r;
    r.v.x = __shfl_up_sync(0xffffffffu, v.x, 1, 16); r.v.y = __shfl_up_sync(0xffffffffu, v.y, 1, 16);
    r.v.z = __shfl_up_sync(0xffffffffu, v.z, 1, 16); r.v.w = __shfl_up_sync(0xffffffffu, v.w, 1, 16);
    return r;
  }
  __device__ __forceinline__ Vec8 shfl_down16() const {
    Vec8 r;
    r.v.x = __shfl_down_sync(0xffffffffu, v.x, 1, 16); r.v.y = __shfl_down_sync(0xffffffffu, v.y, 1, 16);
    r.v.z = __shfl_down_sync(0xffffffffu, v.z, 1, 16); r.v.w = __shfl_down_sync(0xffffffffu, v.w, 1, 16);
    return r;
  }
};
template <> struct Vec8<float> {
  float4 a, b;
  __device__ __forceinline__ void load(const float* p) {
    a = *reinterpret_cast<const float4*>(p); b = *reinterpret_cast<const float4*>(p + 4);
  }
  __device__ __forceinline__ void store(float* p) const {
    *reinterpret_cast<float4*>(p) = a; *reinterpret_cast<float4*>(p + 4) = b;
  }
  __device__ __forceinline__ void zero() { a = make_float4(0, 0, 0, 0); b = a; }
  __device__ __forceinline__ void get(float (&o)[8]) const {
    o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
  }
  __device__ __forceinline__ void set(const float (&o)[8]) {
    a = make_float4(o[0], o[1], o[2], o[3]); b = make_float4(o[4], o[5], o[6], o[7]);
  }
  __device__ __forceinline__ Vec8 shfl_up16() const {
    Vec8 r;
    r.a.x = __shfl_up_sync(0xffffffffu, a.x, 1, 16); r.a.y = __shfl_up_sync(0xffffffffu, a.y, 1, 16);
    r.a.z = __shfl_up_sync(0xffffffffu, a.z, 1, 16); r.a.w = __shfl_up_sync(0xffffffffu, a.w, 1, 16);
    r.b.x = __shfl_up_sync(0xffffffffu, b.x, 1, 16); r.b.y = __shfl_up_sync(0xffffffffu, b.y, 1, 16);
    r.b.z = __shfl_up_sync(0xffffffffu, b.z, 1, 16); r.b.w = __shfl_up_sync(0xffffffffu, b.w, 1, 16);
    return r;
  }
  __device__ __forceinline__ Vec8 shfl_down16() const {
    Vec8 r;
    r.a.x = __shfl_down_sync(0xffffffffu, a.x, 1, 16); r.a.y = __shfl_down_sync(0xffffffffu, a.y, 1, 16);
    r.a.z = __shfl_down_sync(0xffffffffu, a.z, 1, 16); r.a.w = __shfl_down_sync(0xffffffffu, a.w, 1, 16);
    r.b.x = __shfl_down_sync(0xffffffffu, b.x, 1, 16); r.b.y = __shfl_down_sync(0xffffffffu, b.y, 1, 16);
    r.b.z = __shfl_down_sync(0xffffffffu, b.z, 1, 16); r.b.w = __shfl_down_sync(0xffffffffu, b.w, 1, 16);
    return r;
  }
};

// The 8 frames [t, t+8) (t % 8 == 0, segment-relative) of channel c, or zeros outside the segment's
// m-tiles.
template <typename T>
__device__ __forceinline__ Vec8<T> load_group8(const T* __restrict__ buf, int tile0, int ntiles, int c, int t) {
  Vec8<T> v;
  if (t < 0 || t >= ntiles * TILE_M) v.zero();
  else v.load(buf + tt_index(tile0 + (t >> 7), C_HID, c, t & (TILE_M - 1)));
  return v;
}

// One step = one channel per half-warp (8 consecutive frames per thread).
// Interior tiles (every tap of every frame inside the segment): branch-free packed-FMA path with
// precomputed tap pointers; the channel step only adds a constant stride.
template <typename T, int D, bool A01>
__device__ __forceinline__ void dw_step_interior(const T* __restrict__ pc, const T* __restrict__ pl,
                                                 const T* __restrict__ pr, const T* __restrict__ pp,
                                                 const T* __restrict__ pn, T* __restrict__ po, const float4 pa,
                                                 int lane16, float alpha, float2& s2, float2& q2) {
  Vec8<T> vc;
  vc.load(pc);
  float ctr[8], lft[8], rgt[8];
  if (D == 0) {
    Vec8<T> vl, vr;
    vl.load(pl);
    vr.load(pr);
    vc.get(ctr); vl.get(lft); vr.get(rgt);
  } else {
    Vec8<T> vp = vc.shfl_up16(), vn = vc.shfl_down16();
    if (lane16 == 0) vp.load(pp);
    if (lane16 == 15) vn.load(pn);
    float pv[8], nv[8];
    vc.get(ctr); vp.get(pv); vn.get(nv);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      lft[i] = (i - D >= 0) ? ctr[(i - D) & 7] : pv[(8 + i - D) & 7];
      rgt[i] = (i + D < 8) ? ctr[(i + D) & 7] : nv[(i + D - 8) & 7];
    }
  }
  const float2 w0 = make_float2(pa.x, pa.x), w1 = make_float2(pa.y, pa.y), w2 = make_float2(pa.z, pa.z);
  const float2 cw = make_float2(pa.w, pa.w);
  float o[8];
#pragma unroll
  for (int i = 0; i < 8; i += 2) {
    float2 u = fma2(make_float2(lft[i], lft[i + 1]), w0,
                    fma2(make_float2(ctr[i], ctr[i + 1]), w1, fma2(make_float2(rgt[i], rgt[i + 1]), w2, cw)));
    u = prelu2<A01>(u, alpha);
    o[i] = u.x; o[i + 1] = u.y;
    s2 = add2(s2, u);
    q2 = fma2(u, u, q2);
  }
  Vec8<T> ov;
  ov.set(o);
  ov.store(po);
}

// Edge tiles: taps and frames are masked by segment-relative frame index.  The 0/1 masks depend only
// on the thread's frames (not on the channel) and are computed once per thread; the arithmetic is
// branch-free and packed:  u = bias + sum_k m_k * (wa_k * h_k + wb_k),  out = m_centre * PReLU(u).
// Masked-out taps read zeros or finite dead-frame values, so 0 * h is exactly 0.
struct DwEdgeMasks { float2 ml[4], mc[4], mr[4]; };

template <typename T, int D, bool A01>
__device__ __forceinline__ void dw_step_edge(const T* __restrict__ h1, T* __restrict__ h2, const float4 pa,
                                             const float4 pb, const DwEdgeMasks& mk, int tile0, int ntiles,
                                             int m_tile, int c, int t_base, int d, int lane16, float alpha,
                                             float2& s2, float2& q2) {
  const Vec8<T> vc = load_group8(h1, tile0, ntiles, c, t_base);
  float ctr[8], lft[8], rgt[8];
  if (D == 0) {
    const Vec8<T> vl = load_group8(h1, tile0, ntiles, c, t_base - d);
    const Vec8<T> vr = load_group8(h1, tile0, ntiles, c, t_base + d);
    vc.get(ctr); vl.get(lft); vr.get(rgt);
  } else {
    Vec8<T> vp = vc.shfl_up16(), vn = vc.shfl_down16();
    if (lane16 == 0) vp = load_group8(h1, tile0, ntiles, c, t_base - 8);
    if (lane16 == 15) vn = load_group8(h1, tile0, ntiles, c, t_base + 8);
    float pv[8], nv[8];
    vc.get(ctr); vp.get(pv); vn.get(nv);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      lft[i] = (i - D >= 0) ? ctr[(i - D) & 7] : pv[(8 + i - D) & 7];
      rgt[i] = (i + D < 8) ? ctr[(i + D) & 7] : nv[(i + D - 8) & 7];
    }
  }
  const float2 wa0 = make_float2(pa.x, pa.x), wa1 = make_float2(pa.y, pa.y), wa2 = make_float2(pa.z, pa.z);
  const float2 wb0 = make_float2(pb.x, pb.x), wb1 = make_float2(pb.y, pb.y), wb2 = make_float2(pb.z, pb.z);
  const float2 bias = make_float2(pb.w, pb.w);
  float o[8];
#pragma unroll
  for (int i = 0; i < 8; i += 2) {
    const int k = i >> 1;
    float2 u = fma2(fma2(make_float2(lft[i], lft[i + 1]), wa0, wb0), mk.ml[k], bias);
    u = fma2(fma2(make_float2(ctr[i], ctr[i + 1]), wa1, wb1), mk.mc[k], u);
    u = fma2(fma2(make_float2(rgt[i], rgt[i + 1]), wa2, wb2), mk.mr[k], u);
    u = mul2(prelu2<A01>(u, alpha), mk.mc[k]);
    o[i] = u.x; o[i + 1] = u.y;
    s2 = add2(s2, u);
    q2 = fma2(u, u, q2);
  }
  Vec8<T> ov;
  ov.set(o);
  ov.store(h2 + tt_index(m_tile, C_HID, c, lane16 * 8));
}

// D = compile-time dilation for the shuffle path (1, 2, 4); D = 0 selects the aligned-load path
// (runtime dilation d, d % 8 == 0).
template <typename T, int D, bool A01>
__global__ void __launch_bounds__(256, B200SEP_DW_MINB) dwconv_kernel(const T* __restrict__ h1,  // [nt][512][128]
                                                        T* __restrict__ h2,        // [nt][512][128]
                                                        const float2* __restrict__ part1, int slots1,
                                                        float2* __restrict__ part2,  // [nt][DW_SLOTS]
                                                        DwParams p, Tables tb) {
  __shared__ float s_red[16];
  __shared__ float4 s_par[DW_CH][2];  // {wa0, wa1, wa2, c_interior}, {wb0, wb1, wb2, bias}
  pdl_enter();
  const int cg = blockIdx.x;
  const int m_tile = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const TileInfo ti = load_tile_info(tb.tile_info + m_tile);
  const int tile0 = ti.tile0, ntiles = ti.ntiles, F = ti.F;
  const int sub = lane >> 4;                              // which of the warp's 2 channels
  const int lane16 = lane & 15;
  const int tile_t0 = ti.t0;
  const int t_base = tile_t0 + lane16 * 8;               // first of this thread's 8 frames
  const int d = (D > 0) ? D : p.dilation;

  // every warp reduces the (tiny) partial list itself: identical results, no extra barrier
  float mean, rstd;
  warp_seg_stats(part1, tile0, ntiles, slots1, ti.inv_count, p.eps1, mean, rstd, tb.flags);
  for (int i = tid; i < DW_CH; i += 256) {
    const int c = cg * DW_CH + i;
    const float a = p.gamma1[c] * rstd;
    const float b = p.beta1[c] - a * mean;  // y = a*h + b
    const float w0 = p.w[c * 3 + 0], w1 = p.w[c * 3 + 1], w2 = p.w[c * 3 + 2];
    const float bias = p.bias[c];
    s_par[i][0] = make_float4(w0 * a, w1 * a, w2 * a, bias + (w0 + w1 + w2) * b);
    s_par[i][1] = make_float4(w0 * b, w1 * b, w2 * b, bias);
  }
  __syncthreads();

  // CTA-uniform: every tap of every frame of this m-tile lies inside the segment (the shuffle path
  // needs all lanes of a warp in the same branch)
  const bool interior = (tile_t0 - d >= 0) && (tile_t0 + TILE_M - 1 + d < F);
  const float alpha = p.alpha2;
  float2 s2 = make_float2(0.f, 0.f), q2 = make_float2(0.f, 0.f);
  constexpr int STEPS = DW_CH / 16;  // 16 channels per CTA step (8 warps x 2)
  const int cl0 = warp * 2 + sub;
  const int c0 = cg * DW_CH + cl0;
  if (interior) {
    // tap pointers of channel c0; one channel step = +16 channels = 16 * 128 elements
    auto ptr_at = [&](int t) { return h1 + tt_index(tile0 + (t >> 7), C_HID, c0, t & (TILE_M - 1)); };
    const T* pc = ptr_at(t_base);
    const T* pl = ptr_at(t_base - d);
    const T* pr = ptr_at(t_base + d);
    const T* pp = (D > 0 && lane16 == 0) ? ptr_at(t_base - 8) : pc;
    const T* pn = (D > 0 && lane16 == 15) ? ptr_at(t_base + 8) : pc;
    T* po = h2 + tt_index(m_tile, C_HID, c0, lane16 * 8);
    constexpr size_t STRIDE = (size_t)16 * TILE_M;
    constexpr int kUnroll = B200SEP_DW_UNROLL;
#pragma unroll kUnroll
    for (int it = 0; it < STEPS; ++it)
      dw_step_interior<T, D, A01>(pc + it * STRIDE, pl + it * STRIDE, pr + it * STRIDE, pp + it * STRIDE,
                                  pn + it * STRIDE, po + it * STRIDE, s_par[it * 16 + cl0][0], lane16, alpha, s2, q2);
  } else {
    DwEdgeMasks mk;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int ta = t_base + 2 * k, tb2 = ta + 1;
      mk.ml[k] = make_float2((ta - d >= 0 && ta - d < F) ? 1.f : 0.f, (tb2 - d >= 0 && tb2 - d < F) ? 1.f : 0.f);
      mk.mc[k] = make_float2(ta < F ? 1.f : 0.f, tb2 < F ? 1.f : 0.f);
      mk.mr[k] = make_float2(ta + d < F ? 1.f : 0.f, tb2 + d < F ? 1.f : 0.f);
    }
#pragma unroll 1
    for (int it = 0; it < STEPS; ++it)
      dw_step_edge<T, D, A01>(h1, h2, s_par[it * 16 + cl0][0], s_par[it * 16 + cl0][1], mk, tile0, ntiles, m_tile,
                              c0 + it * 16, t_base, d, lane16, alpha, s2, q2);
  }
  float s_acc = warp_sum(s2.x + s2.y);
  float q_acc = warp_sum(q2.x + q2.y);
  if (lane == 0) { s_red[warp] = s_acc; s_red[8 + warp] = q_acc; }
  __syncthreads();
  if (tid == 0) {
    float s = 0.f, q = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) { s += s_red[w]; q += s_red[8 + w]; }
    part2[(size_t)m_tile * DW_SLOTS + cg] = make_float2(s, q);
  }
}

// =================================================================================================
// Decoder: dec[s][f*stride + l] += sum_n masked[s][n][f] * Wd[n][l]   (asteroid Decoder(FreeFB) =
// conv_transpose1d with stride L/2, no bias) followed by pad_x_to_y (zero tail).  d_out is zero-
// filled before this kernel.  grid = (n_tiles, n_src), 256 threads = (64 frame pairs) x (4 quarters of
// the 512 filters).  The filterbank sits in shared memory (broadcast reads); a thread accumulates the
// L taps of its two frames as packed fp32 pairs (taps l, l+1).  The four partial sums per frame are
// combined in shared memory, where the overlap-add of the two half-frames per hop also happens;
// stores are coalesced.  Only the two hops shared with the neighbouring m-tiles use atomicAdd; with a
// zeroed destination the sum of exactly two addends is order-independent, so the result is
// deterministic.
// =================================================================================================
template <int L>
constexpr int decoder_smem_bytes() {
  return (C_ENC * L * 4 > 4 * TILE_M * (L + 1) * 4) ? C_ENC * L * 4 : 4 * TILE_M * (L + 1) * 4;
}

template <typename T> struct Ld2;
template <> struct Ld2<__half> {
  __device__ __forceinline__ static float2 ld(const __half* p) {
    return __half22float2(*reinterpret_cast<const __half2*>(p));
  }
};
template <> struct Ld2<float> {
  __device__ __forceinline__ static float2 ld(const float* p) { return *reinterpret_cast<const float2*>(p); }
};

// Encoder, second version (L = 16): a thread owns FOUR consecutive frames and walks 32 filters; the taps of a filter are
// read once per thread as (w, w) pairs (one 16-byte shared-memory load = two taps) and applied to two frame pairs with
// packed FFMA2 -- a quarter of the shared-memory wavefronts and half of the FMA instructions per output of the
// one-frame-per-thread kernel above (which was L1-bound at 83 %).  Same tap order per output, so results are identical.
// grid = (n_tiles, 2): a CTA = 128 frames x 256 filters, warp w -> filters [32 w, 32 w + 32), lane -> frames 4 lane .. + 3.
template <typename T> __device__ __forceinline__ void enc_store4(T* p, float a, float b, float c, float d);
template <> __device__ __forceinline__ void enc_store4<float>(float* p, float a, float b, float c, float d) {
  *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d);
}
template <> __device__ __forceinline__ void enc_store4<__half>(__half* p, float a, float b, float c, float d) {
  const __half2 lo = __floats2half2_rn(a, b), hi = __floats2half2_rn(c, d);
  *reinterpret_cast<uint2*>(p) = make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
}
template <typename T>
__global__ void __launch_bounds__(256) encoder4_kernel(const float* __restrict__ mix,
                                                       const float* __restrict__ w_enc,  // [512][16]
                                                       T* __restrict__ encT,             // [nt][512][128]
                                                       float2* __restrict__ part,        // [nt][2]
                                                       Tables tb) {
  constexpr int L = 16, S = 8, CH_PER_CTA = 256;
  pdl_enter();
  __shared__ float4 s_w2[CH_PER_CTA * L / 2];               // (w_l, w_l, w_{l+1}, w_{l+1}): 32 KiB
  __shared__ float4 s_x4[(TILE_M * S + S) / 4 + 2];
  __shared__ float s_red[16];
  float* s_x = reinterpret_cast<float*>(s_x4);
  const int m_tile = blockIdx.x, chalf = blockIdx.y, tid = threadIdx.x;
  const TileInfo ti = load_tile_info(tb.tile_info + m_tile);
  const int seg = ti.seg, t0 = ti.t0, F = ti.F;
  const int T_len = tb.seg_len[seg];
  const float* x = mix + tb.seg_in_off[seg];
  const float2* wg = reinterpret_cast<const float2*>(w_enc + (size_t)chalf * CH_PER_CTA * L);
  for (int i = tid; i < CH_PER_CTA * L / 2; i += 256) {
    const float2 w = __ldg(wg + i);
    s_w2[i] = make_float4(w.x, w.x, w.y, w.y);
  }
  const int s0 = t0 * S;
  for (int i = tid; i < TILE_M * S + S; i += 256) {
    const int gi = s0 + i;
    s_x[i] = (gi < T_len) ? x[gi] : 0.f;
  }
  __syncthreads();

  const int warp = tid >> 5, lane = tid & 31;
  const int f0 = 4 * lane;
  // samples of the four frames: 40 consecutive floats; pairs (frame 0, frame 1) and (frame 2, frame 3) per tap
  float xv[40];
#pragma unroll
  for (int q = 0; q < 10; ++q) {
    const float4 v = s_x4[(f0 * S) / 4 + q];
    xv[4 * q] = v.x; xv[4 * q + 1] = v.y; xv[4 * q + 2] = v.z; xv[4 * q + 3] = v.w;
  }
  float2 p01[L], p23[L];
#pragma unroll
  for (int l = 0; l < L; ++l) {
    p01[l] = make_float2(xv[l], xv[S + l]);
    p23[l] = make_float2(xv[2 * S + l], xv[3 * S + l]);
  }
  const float2 vm01 = make_float2((t0 + f0) < F ? 1.f : 0.f, (t0 + f0 + 1) < F ? 1.f : 0.f);
  const float2 vm23 = make_float2((t0 + f0 + 2) < F ? 1.f : 0.f, (t0 + f0 + 3) < F ? 1.f : 0.f);
  float2 s2 = make_float2(0.f, 0.f), q2 = make_float2(0.f, 0.f);
  const int c_base = warp * 32;
  T* dst = encT + tt_index(m_tile, C_ENC, chalf * CH_PER_CTA + c_base, f0);
#pragma unroll 2
  for (int j = 0; j < 32; ++j) {
    const float4* w4 = s_w2 + (c_base + j) * (L / 2);
    float2 a01 = make_float2(0.f, 0.f), a23 = make_float2(0.f, 0.f);
#pragma unroll
    for (int q = 0; q < L / 2; ++q) {
      const float4 w = w4[q];
      a01 = fma2(make_float2(w.x, w.y), p01[2 * q], a01);
      a23 = fma2(make_float2(w.x, w.y), p23[2 * q], a23);
      a01 = fma2(make_float2(w.z, w.w), p01[2 * q + 1], a01);
      a23 = fma2(make_float2(w.z, w.w), p23[2 * q + 1], a23);
    }
    a01 = mul2(a01, vm01);
    a23 = mul2(a23, vm23);
    enc_store4<T>(dst + (size_t)j * TILE_M, a01.x, a01.y, a23.x, a23.y);
    s2 = add2(s2, add2(a01, a23));
    q2 = fma2(a01, a01, q2);
    q2 = fma2(a23, a23, q2);
  }
  const float s_acc = warp_sum(s2.x + s2.y), q_acc = warp_sum(q2.x + q2.y);
  if (lane == 0) { s_red[warp] = s_acc; s_red[8 + warp] = q_acc; }
  __syncthreads();
  if (tid == 0) {
    float s = 0.f, q = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) { s += s_red[w]; q += s_red[8 + w]; }
    part[(size_t)m_tile * 2 + chalf] = make_float2(s, q);
  }
}

template <typename T, int L>
__global__ void __launch_bounds__(256) decoder_kernel(const T* __restrict__ masked,  // [nt][S*512][128]
                                                      const float* __restrict__ w_dec,  // [512][L]
                                                      float* __restrict__ out, int n_src, Tables tb) {
  constexpr int S = L / 2;
  pdl_enter();
  extern __shared__ __align__(16) float s_dec[];
  float4* s_w4 = reinterpret_cast<float4*>(s_dec);  // [512][L/4] while accumulating
  const int m_tile = blockIdx.x;
  const int src = blockIdx.y;
  const int tid = threadIdx.x;
  const TileInfo ti = load_tile_info(tb.tile_info + m_tile);
  const int seg = ti.seg, t0 = ti.t0, F = ti.F;
  const int T_len = tb.seg_len[seg];
  const int fp = tid & 63;  // frames 2fp, 2fp+1
  const int nq = tid >> 6;  // filters [128 nq, 128 nq + 128)

  const float4* wg = reinterpret_cast<const float4*>(w_dec);
  for (int i = tid; i < C_ENC * L / 4; i += 256) s_w4[i] = __ldg(wg + i);
  __syncthreads();

  float2 acc[2][L / 2];
#pragma unroll
  for (int l = 0; l < L / 2; ++l) acc[0][l] = acc[1][l] = make_float2(0.f, 0.f);
  const T* mp = masked + tt_index(m_tile, n_src * C_ENC, src * C_ENC + nq * 128, 2 * fp);
  const float4* wp = s_w4 + (size_t)nq * 128 * (L / 4);
#pragma unroll 4
  for (int n = 0; n < 128; ++n) {
    const float2 m = Ld2<T>::ld(mp + (size_t)n * TILE_M);
    const float2 m0 = make_float2(m.x, m.x), m1 = make_float2(m.y, m.y);
#pragma unroll
    for (int q = 0; q < L / 4; ++q) {
      const float4 w = wp[n * (L / 4) + q];
      acc[0][2 * q + 0] = fma2(make_float2(w.x, w.y), m0, acc[0][2 * q + 0]);
      acc[0][2 * q + 1] = fma2(make_float2(w.z, w.w), m0, acc[0][2 * q + 1]);
      acc[1][2 * q + 0] = fma2(make_float2(w.x, w.y), m1, acc[1][2 * q + 0]);
      acc[1][2 * q + 1] = fma2(make_float2(w.z, w.w), m1, acc[1][2 * q + 1]);
    }
  }
  __syncthreads();  // filterbank no longer needed: reuse the buffer for the partial frames
  float (*s_d)[TILE_M][L + 1] = reinterpret_cast<float (*)[TILE_M][L + 1]>(s_dec);
#pragma unroll
  for (int f = 0; f < 2; ++f) {
    const bool valid = (t0 + 2 * fp + f) < F;
#pragma unroll
    for (int l = 0; l < L / 2; ++l) {
      s_d[nq][2 * fp + f][2 * l + 0] = valid ? acc[f][l].x : 0.f;
      s_d[nq][2 * fp + f][2 * l + 1] = valid ? acc[f][l].y : 0.f;
    }
  }
  __syncthreads();

  float* o = out + tb.seg_out_off[seg] + (size_t)src * T_len;
  // hops hh = 0..128 of this tile: sample (t0 + hh) * S + l gets frame hh's first half and frame
  // hh-1's second half.
  for (int i = tid; i < (TILE_M + 1) * S; i += 256) {
    const int hh = i / S, l = i % S;
    const int f_hi = t0 + hh;      // frame contributing taps [0, S)
    const int f_lo = t0 + hh - 1;  // frame contributing taps [S, L)
    float v = 0.f;
    bool any = false;
    if (hh < TILE_M && f_hi < F) {
      v += (s_d[0][hh][l] + s_d[1][hh][l]) + (s_d[2][hh][l] + s_d[3][hh][l]);
      any = true;
    }
    if (hh >= 1 && f_lo < F) {
      v += (s_d[0][hh - 1][S + l] + s_d[1][hh - 1][S + l]) + (s_d[2][hh - 1][S + l] + s_d[3][hh - 1][S + l]);
      any = true;
    }
    if (!any) continue;
    const size_t idx = (size_t)(t0 + hh) * S + l;  // < (F-1)*S + L <= T_len
    if (hh == 0 || hh == TILE_M) atomicAdd(o + idx, v);
    else o[idx] = v;
  }
}

// Converts a row-major fp32 matrix [M][K] (M % 128 == 0) into the tile-transposed layout of type T.
template <typename T>
__global__ void to_tile_transposed_kernel(const float* __restrict__ a, T* __restrict__ at, int M, int K) {
  const size_t n = (size_t)M * K;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    // i enumerates the destination (coalesced writes)
    const int r = (int)(i % TILE_M);
    const int k = (int)((i / TILE_M) % K);
    const int mt = (int)(i / ((size_t)TILE_M * K));
    at[i] = from_f32<T>(a[((size_t)mt * TILE_M + r) * K + k]);
  }
}

template <typename T>
__global__ void convert_kernel(const float* __restrict__ a, T* __restrict__ b, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    b[i] = from_f32<T>(a[i]);
}

}  // namespace b200sep

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
  __shared__ float4 s_w[CH_PER_CTA * L / 4];
  __shared__ float4 s_x4[(TILE_M * S + S) / 4 + 1];
  __shared__ float s_red[16];
  float* s_x = reinterpret_cast<float*>(s_x4);

  const int m_tile = blockIdx.x;
  const int chalf = blockIdx.y;
  const int tid = threadIdx.x;
  const int seg = tb.tile_seg[m_tile];
  const int t0 = (m_tile - tb.seg_tile0[seg]) * TILE_M;
  const int F = tb.seg_frames[seg];
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
//   h2  = PReLU(u)            and (sum, sumsq) partials of h2 for gLN2 (slots = 8 per m-tile)
// The zero padding applies to y (the normalised signal) at the TRUE segment ends, so taps that
// fall outside [0, F_i) contribute nothing -- not even beta.
// grid = (8 channel groups of 64, n_tiles), 256 threads: warp -> 8 channels, lane -> 4 consecutive
// frames (one 8/16-byte vector).  Taps at +-d are vector loads from the shifted position when
// d % 4 == 0, and lane shuffles (+ one edge vector) when d is 1 or 2.
// =================================================================================================
struct DwParams {
  const float* gamma1;  // [512]
  const float* beta1;   // [512]
  const float* w;       // [512][3]
  const float* bias;    // [512]
  float alpha2;         // PReLU slope after the depthwise conv
  int dilation;
};

template <typename T> struct Vec4;
template <> struct Vec4<float> {
  float4 v;
  __device__ __forceinline__ void load(const float* p) { v = *reinterpret_cast<const float4*>(p); }
  __device__ __forceinline__ void store(float* p) const { *reinterpret_cast<float4*>(p) = v; }
  __device__ __forceinline__ void get(float (&o)[4]) const { o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; }
  __device__ __forceinline__ void set(const float (&o)[4]) { v = make_float4(o[0], o[1], o[2], o[3]); }
};
template <> struct Vec4<__half> {
  uint2 v;
  __device__ __forceinline__ void load(const __half* p) { v = *reinterpret_cast<const uint2*>(p); }
  __device__ __forceinline__ void store(__half* p) const { *reinterpret_cast<uint2*>(p) = v; }
  __device__ __forceinline__ void get(float (&o)[4]) const {
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&v.x));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&v.y));
    o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y;
  }
  __device__ __forceinline__ void set(const float (&o)[4]) {
    const __half2 a = __floats2half2_rn(o[0], o[1]);
    const __half2 b = __floats2half2_rn(o[2], o[3]);
    v.x = *reinterpret_cast<const uint32_t*>(&a);
    v.y = *reinterpret_cast<const uint32_t*>(&b);
  }
};

// Loads the 4 frames [t, t+4) of channel c of segment-relative frame t (t % 4 == 0) as floats,
// or zeros when the group lies outside the segment's m-tiles.
template <typename T>
__device__ __forceinline__ void load_group(const T* __restrict__ buf, int tile0, int ntiles, int c,
                                           int t, float (&o)[4]) {
  if (t < 0 || t >= ntiles * TILE_M) {
    o[0] = o[1] = o[2] = o[3] = 0.f;
    return;
  }
  Vec4<T> v;
  v.load(buf + tt_index(tile0 + (t >> 7), C_HID, c, t & (TILE_M - 1)));
  v.get(o);
}

template <typename T>
__global__ void __launch_bounds__(256) dwconv_kernel(const T* __restrict__ h1,  // [nt][512][128]
                                                     T* __restrict__ h2,        // [nt][512][128]
                                                     const float2* __restrict__ part1, int slots1,
                                                     float2* __restrict__ part2,  // [nt][8]
                                                     DwParams p, Tables tb) {
  __shared__ float s_stat[2];
  __shared__ float s_red[16];
  const int cg = blockIdx.x;  // 64-channel group
  const int m_tile = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int seg = tb.tile_seg[m_tile];
  const int tile0 = tb.seg_tile0[seg];
  const int ntiles = tb.seg_ntiles[seg];
  const int F = tb.seg_frames[seg];
  const int t_base = (m_tile - tile0) * TILE_M + lane * 4;  // first of this lane's 4 frames
  const int d = p.dilation;

  if (warp == 0) {
    float mean, rstd;
    warp_seg_stats(part1, tile0, ntiles, slots1, (double)C_HID * (double)F, mean, rstd);
    if (lane == 0) { s_stat[0] = mean; s_stat[1] = rstd; }
  }
  __syncthreads();
  const float mean = s_stat[0], rstd = s_stat[1];

  float s_acc = 0.f, q_acc = 0.f;
#pragma unroll 2
  for (int j = 0; j < 8; ++j) {
    const int c = cg * 64 + warp * 8 + j;
    const float g = p.gamma1[c] * rstd;
    const float b = p.beta1[c] - g * mean;  // y = g*h + b
    const float w0 = p.w[c * 3 + 0], w1 = p.w[c * 3 + 1], w2 = p.w[c * 3 + 2];
    const float bias = p.bias[c];

    float ctr[4], lft[4], rgt[4];
    load_group(h1, tile0, ntiles, c, t_base, ctr);
    if ((d & 3) == 0) {
      load_group(h1, tile0, ntiles, c, t_base - d, lft);
      load_group(h1, tile0, ntiles, c, t_base + d, rgt);
    } else {
      // d in {1,2}: neighbours come from the adjacent lanes' vectors (edge lanes load one more)
      float pv[4], nv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        pv[i] = __shfl_up_sync(0xffffffffu, ctr[i], 1);
        nv[i] = __shfl_down_sync(0xffffffffu, ctr[i], 1);
      }
      if (lane == 0) load_group(h1, tile0, ntiles, c, t_base - 4, pv);
      if (lane == 31) load_group(h1, tile0, ntiles, c, t_base + 4, nv);
      if (d == 1) {
        lft[0] = pv[3]; lft[1] = ctr[0]; lft[2] = ctr[1]; lft[3] = ctr[2];
        rgt[0] = ctr[1]; rgt[1] = ctr[2]; rgt[2] = ctr[3]; rgt[3] = nv[0];
      } else {  // d == 2
        lft[0] = pv[2]; lft[1] = pv[3]; lft[2] = ctr[0]; lft[3] = ctr[1];
        rgt[0] = ctr[2]; rgt[1] = ctr[3]; rgt[2] = nv[0]; rgt[3] = nv[1];
      }
    }
    float o[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int t = t_base + i;
      float u = bias;
      if (t - d >= 0 && t - d < F) u = fmaf(w0, fmaf(g, lft[i], b), u);
      if (t < F) u = fmaf(w1, fmaf(g, ctr[i], b), u);
      if (t + d < F) u = fmaf(w2, fmaf(g, rgt[i], b), u);
      u = (u > 0.f) ? u : p.alpha2 * u;
      if (t >= F) u = 0.f;
      o[i] = u;
      s_acc += u;
      q_acc = fmaf(u, u, q_acc);
    }
    Vec4<T> ov;
    ov.set(o);
    ov.store(h2 + tt_index(m_tile, C_HID, c, lane * 4));
  }
  s_acc = warp_sum(s_acc);
  q_acc = warp_sum(q_acc);
  if (lane == 0) { s_red[warp] = s_acc; s_red[8 + warp] = q_acc; }
  __syncthreads();
  if (tid == 0) {
    float s = 0.f, q = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) { s += s_red[w]; q += s_red[8 + w]; }
    part2[(size_t)m_tile * 8 + cg] = make_float2(s, q);
  }
}

// =================================================================================================
// Decoder: dec[s][f*stride + l] += sum_n masked[s][n][f] * Wd[n][l]   (asteroid Decoder(FreeFB) =
// conv_transpose1d with stride L/2, no bias) followed by pad_x_to_y (zero tail).  d_out is zero-
// filled before this kernel.  grid = (n_tiles, n_src), 256 threads = (frame r, half of the 512
// channels); each thread accumulates the L taps of its frame, halves are combined in shared memory,
// then the overlap-add of the two half-frames per hop is done in shared memory and written with
// coalesced stores.  Only the two hops shared with the neighbouring m-tiles use atomicAdd; with a
// zeroed destination the sum of exactly two addends is order-independent, so the result is
// deterministic.
// =================================================================================================
template <typename T, int L>
__global__ void __launch_bounds__(256) decoder_kernel(const T* __restrict__ masked,  // [nt][S*512][128]
                                                      const float* __restrict__ w_dec,  // [512][L]
                                                      float* __restrict__ out, int n_src, Tables tb) {
  constexpr int S = L / 2;
  __shared__ float s_d[TILE_M][L + 1];
  const int m_tile = blockIdx.x;
  const int src = blockIdx.y;
  const int tid = threadIdx.x;
  const int seg = tb.tile_seg[m_tile];
  const int t0 = (m_tile - tb.seg_tile0[seg]) * TILE_M;
  const int F = tb.seg_frames[seg];
  const int T_len = tb.seg_len[seg];
  const int r = tid & (TILE_M - 1);
  const int nh = tid >> 7;

  float acc[L];
#pragma unroll
  for (int l = 0; l < L; ++l) acc[l] = 0.f;
  const T* mp = masked + tt_index(m_tile, n_src * C_ENC, src * C_ENC + nh * 256, r);
  const float4* wp = reinterpret_cast<const float4*>(w_dec + (size_t)nh * 256 * L);
#pragma unroll 2
  for (int n = 0; n < 256; ++n) {
    const float m = to_f32(mp[(size_t)n * TILE_M]);
#pragma unroll
    for (int q = 0; q < L / 4; ++q) {
      const float4 w = __ldg(wp + n * (L / 4) + q);
      acc[4 * q + 0] = fmaf(m, w.x, acc[4 * q + 0]);
      acc[4 * q + 1] = fmaf(m, w.y, acc[4 * q + 1]);
      acc[4 * q + 2] = fmaf(m, w.z, acc[4 * q + 2]);
      acc[4 * q + 3] = fmaf(m, w.w, acc[4 * q + 3]);
    }
  }
  if (nh == 1) {
#pragma unroll
    for (int l = 0; l < L; ++l) s_d[r][l] = acc[l];
  }
  __syncthreads();
  if (nh == 0) {
    const bool valid = (t0 + r) < F;
#pragma unroll
    for (int l = 0; l < L; ++l) s_d[r][l] = valid ? (acc[l] + s_d[r][l]) : 0.f;
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
    if (hh < TILE_M && f_hi < F) { v += s_d[hh][l]; any = true; }
    if (hh >= 1 && f_lo < F) { v += s_d[hh - 1][S + l]; any = true; }
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

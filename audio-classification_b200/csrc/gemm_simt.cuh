// CUDA-core GEMM for the 1x1 convolutions: out[frame][n] = sum_k A[frame][k] * B[n][k].
// Used by the fp32 "exact" mode (parity scaffold, >= 100 dB vs the oracle) and, with T = __half
// storage, to validate the tcgen05 path.  A is tile-transposed ([m_tile][k][128], frames fastest),
// B is [N][K] row-major.  CTA tile 128 frames x 128 channels, BK = 16, 256 threads, each thread
// 8 frames (r = tx + 16*i) x 8 channels (two groups of 4), fp32 accumulation in K order.
#pragma once
#include "common.cuh"
#include "epilogues.cuh"

namespace b200sep {

template <typename T> struct Ld8;
template <> struct Ld8<float> {
  __device__ __forceinline__ static void ld(const float* p, float (&o)[8]) {
    const float4 a = *reinterpret_cast<const float4*>(p);
    const float4 b = *reinterpret_cast<const float4*>(p + 4);
    o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
  }
};
template <> struct Ld8<__half> {
  __device__ __forceinline__ static void ld(const __half* p, float (&o)[8]) {
    const uint4 v = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
      o[2 * i] = f.x; o[2 * i + 1] = f.y;
    }
  }
};

template <typename T, int K, class Epi>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const T* __restrict__ A,  // [nt][K][128]
                                                        const T* __restrict__ B,  // [N][K]
                                                        Epi epi, Tables tb) {
  constexpr int BK = 16, BN = 128, PAD = 4;
  __shared__ __align__(16) float As[2][BK][TILE_M];
  __shared__ __align__(16) float Bs[2][BK][BN + PAD];
  pdl_enter();
  const int n_tile = blockIdx.x, m_tile = blockIdx.y;
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;

  const T* Ag = A + (size_t)m_tile * K * TILE_M;
  const T* Bg = B + (size_t)n_tile * BN * K;
  // global -> register staging assignments
  const int a_k = tid >> 4, a_m = (tid & 15) * 8;  // A: row k, 8 consecutive frames
  const int b_n = tid >> 1, b_k = (tid & 1) * 8;   // B: row n, 8 consecutive k

  float ra[8], rb[8];
  Ld8<T>::ld(Ag + (size_t)a_k * TILE_M + a_m, ra);
  Ld8<T>::ld(Bg + (size_t)b_n * K + b_k, rb);

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  constexpr int KT = K / BK;
  for (int kt = 0; kt < KT; ++kt) {
    const int buf = kt & 1;
    *reinterpret_cast<float4*>(&As[buf][a_k][a_m]) = make_float4(ra[0], ra[1], ra[2], ra[3]);
    *reinterpret_cast<float4*>(&As[buf][a_k][a_m + 4]) = make_float4(ra[4], ra[5], ra[6], ra[7]);
#pragma unroll
    for (int j = 0; j < 8; ++j) Bs[buf][b_k + j][b_n] = rb[j];
    __syncthreads();
    if (kt + 1 < KT) {
      Ld8<T>::ld(Ag + (size_t)((kt + 1) * BK + a_k) * TILE_M + a_m, ra);
      Ld8<T>::ld(Bg + (size_t)b_n * K + (kt + 1) * BK + b_k, rb);
    }
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[8], b[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = As[buf][k][tx + 16 * i];
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][ty * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + ty * 4]);
      b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
      b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    // the next iteration writes the other buffer; the barrier at its top orders reuse of this one
  }

  EpiState st;
  epi.tile_begin(st, m_tile, tb);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = tx + 16 * i;
    float pre0[4], pre1[4];
    epi.template prefetch<4>(m_tile, r, n_tile * BN + ty * 4, pre0);
    epi.template prefetch<4>(m_tile, r, n_tile * BN + 64 + ty * 4, pre1);
    epi.template apply<4>(st, m_tile, r, n_tile * BN + ty * 4, &acc[i][0], pre0);
    epi.template apply<4>(st, m_tile, r, n_tile * BN + 64 + ty * 4, &acc[i][4], pre1);
  }
  if (Epi::kStats) {
    // block-level reduction -> one partial per (m_tile, n_tile): slots = N / 128
    __shared__ float s_red[16];
    const float s = warp_sum(st.s), q = warp_sum(st.q);
    const int warp = tid >> 5, lane = tid & 31;
    if (lane == 0) { s_red[warp] = s; s_red[8 + warp] = q; }
    __syncthreads();
    if (tid == 0) {
      float ts = 0.f, tq = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) { ts += s_red[w]; tq += s_red[8 + w]; }
      st.s = ts;
      st.q = tq;
      epi.tile_end(st, m_tile, gridDim.x, n_tile);
    }
  }
}

}  // namespace b200sep

// Encoder on the tensor cores (sm_100a, fp16 tensor-core mode): tf_rep[filter][frame] = sum_k W_enc[filter][k] x[frame S + k]
// (asteroid Encoder(FreeFB): conv1d(1 -> 512, kernel L, stride L / 2, no bias, no activation).
//
// The CUDA-core kernel (bandwidth_kernels.cuh) is bound by its shared-memory / L1 traffic (l1tex 83 %, 94 us for
// 137 MB of output: 0.23 of HBM).  Here the L taps are the K dimension of one tcgen05 MMA per 128-filter chunk:
//   B operand  (frames x K, K-major): row f = [hi(x[f S .. f S + L)) | lo(...)], x = hi + lo split into two halves, so
//              the waveform enters the tensor core with ~22 mantissa bits (K = 2 L);
//   A operand  (filters x K, K-major): [W | W] rounded to fp16, resident in shared memory for the whole launch (64 KiB);
//   D[filter][frame] in tensor memory: four 128-column chunks, as in K_B;
//   epilogue   (8 warps, lane = filter): fp16 pack, bottleneck-gLN partial sums, swizzled staging boxes, TMA tile stores
//              -- K_B's epilogue without bias / PReLU.
// Persistent grid, one CTA per SM; frames beyond the segment end get an all-zero operand row, so their outputs are
// exact zeros (what the CUDA-core kernel writes) and need no masking in the statistics.
#pragma once
#include "fused_block.cuh"

namespace b200sep {
namespace enc {
using namespace tc;
using namespace fb;

template <int L>
struct EncCfg {
  static constexpr int K2 = 2 * L;                      // hi | lo
  static constexpr int W_BYTES = C_ENC * 128;           // 64 KiB: 512 rows x 128 B (K2 * 2 <= 128 bytes used per row)
  static constexpr int F_BYTES = TILE_M * 128;          // 16 KiB: 128 frame rows x 128 B
  static constexpr int ST_BYTES = TILE_M * 128 * 2;     // 32 KiB staging: 4 boxes of 64 filters x 64 frames
  static constexpr int EPI_WARPS = 8, FR_WARPS = 4;
  static constexpr int THREADS = (EPI_WARPS + 1 + FR_WARPS) * 32;  // 416
  static constexpr int TMEM_COLS = 512;
  static constexpr int OFF_F = W_BYTES;
  static constexpr int OFF_ST = OFF_F + 2 * F_BYTES;
  static constexpr int OFF_TI = OFF_ST + ST_BYTES;
  static constexpr int OFF_RED = OFF_TI + TI_CACHE * 32;
  static constexpr int OFF_CNT = OFF_RED + 2 * EPI_WARPS * 8;
  static constexpr int OFF_BAR = OFF_CNT + 16;
  static constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;
  static_assert(K2 * 2 <= 128, "one 128-byte operand row holds the hi | lo taps");
};

// x -> (hi, lo) halves with x ~= hi + lo
__device__ __forceinline__ void split_half(float x, __half& hi, __half& lo) {
  hi = __float2half_rn(x);
  lo = __float2half_rn(x - __half2float(hi));
}

template <int L>
__global__ void __launch_bounds__(EncCfg<L>::THREADS, 1)
encoder_tc_kernel(const __grid_constant__ CUtensorMap tm_enc, const float* __restrict__ mix,
                  const float* __restrict__ w_enc, float2* __restrict__ part, Tables tb) {
  using Cfg = EncCfg<L>;
  constexpr int S = L / 2;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t smem_w = smem_base;
  auto smem_f = [&](int b) { return smem_base + (uint32_t)(Cfg::OFF_F + b * Cfg::F_BYTES); };
  TileInfo* s_ti = reinterpret_cast<TileInfo*>(smem_gen + Cfg::OFF_TI);
  float2* s_red = reinterpret_cast<float2*>(smem_gen + Cfg::OFF_RED);  // [2][EPI_WARPS]
  int* s_cnt = reinterpret_cast<int*>(smem_gen + Cfg::OFF_CNT);          // [2]
  const uint32_t bar0 = smem_base + Cfg::OFF_BAR;
  auto fr_full = [&](int b) { return bar0 + 8u * (0 + b); };
  auto fr_empty = [&](int b) { return bar0 + 8u * (2 + b); };
  auto tfull_bar = [&](int q) { return bar0 + 8u * (4 + q); };
  auto tempty_bar = [&](int q) { return bar0 + 8u * (8 + q); };
  const uint32_t tmem_ptr_smem = bar0 + 8u * 12;
  const uint32_t done_bar = bar0 + 8u * 13;

  pdl_launch_dependents();
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int n_tiles = tb.n_tiles;
  constexpr int MMA_WARP = Cfg::EPI_WARPS;

  ti_cache_fill(s_ti, tb, threadIdx.x);
  if (threadIdx.x < 2) s_cnt[threadIdx.x] = 0;
  // resident A operand: row r = filter r, 16-byte chunk c (k = 8 c .. 8 c + 7) at chunk c ^ (r & 7); k < L: tap k,
  // L <= k < 2 L: tap k - L again (it multiplies the lo halves of the samples)
  for (int i = threadIdx.x; i < C_ENC * (Cfg::K2 / 8); i += Cfg::THREADS) {
    const int r = i / (Cfg::K2 / 8), c = i % (Cfg::K2 / 8);
    const float* wr = w_enc + (size_t)r * L + ((c * 8) % L);
    float o[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = __ldg(wr + k);
    st_shared_v4(smem_w + (uint32_t)(r * 128) + (uint32_t)((c ^ (r & 7)) << 4), pack8(o));
  }
  if (warp == MMA_WARP) {
    if (lane == 0) {
      tma_prefetch_desc(&tm_enc);
      for (int b = 0; b < 2; ++b) {
        mbar_init(fr_full(b), Cfg::FR_WARPS);
        mbar_init(fr_empty(b), 1);
      }
      for (int q = 0; q < 4; ++q) {
        mbar_init(tfull_bar(q), 1);
        mbar_init(tempty_bar(q), Cfg::EPI_WARPS);
      }
      mbar_init(done_bar, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_ptr_smem, Cfg::TMEM_COLS);
  }
  fence_proxy_async();   // the weight image was written with generic stores and is read by the tensor core
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  pdl_wait();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_smem));

  if (warp < MMA_WARP) {
    // ================================ epilogue: fp16 tf_rep + bottleneck-gLN partials =============
    const int ew = warp, quad = ew & 3, half = ew >> 2;   // half = frame half of the m-tile
    const int cl = quad * 32 + lane;                       // filter inside the chunk
    // every epilogue warp stages and stores its own 32 filters x 64 frames (a 4 KiB box per chunk), as in K_B
    const int row = lane;
    const uint32_t st_box = smem_base + (uint32_t)(Cfg::OFF_ST + ew * 4096);
    const uint32_t st_row = st_box + (uint32_t)(row * 128);
    const bool issuer = (lane == 0);                       // one thread per box issues its TMA store
    int it = 0;
    for (int w = blockIdx.x; w < n_tiles; w += gridDim.x, ++it) {
      float2 s2 = make_float2(0.f, 0.f), q2 = make_float2(0.f, 0.f);
      const uint32_t ph = (uint32_t)it & 1u;
      float v[2][32];
      mbar_wait_b(tb.flags, tfull_bar(0), ph);
      tcgen05_fence_after();
      tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(half * 64), v[0]);
#pragma unroll
      for (int pi = 0; pi < 8; ++pi) {
        const int q = pi >> 1, hh = pi & 1;
        tmem_ld_wait();  // piece pi (32 frames) is in v[pi & 1]
        if (hh == 1) {
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty_bar(q));  // the warp's part of chunk q is in registers
        }
        if (pi + 1 < 8) {
          const int qn = (pi + 1) >> 1;
          if (((pi + 1) & 1) == 0) {
            mbar_wait_b(tb.flags, tfull_bar(qn), ph);
            tcgen05_fence_after();
          }
          tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(qn * 128 + half * 64 + ((pi + 1) & 1) * 32),
                    v[(pi + 1) & 1]);
        }
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          const float2 u = make_float2(v[pi & 1][j], v[pi & 1][j + 1]);
          pk[j >> 1] = h2_as_u32(__floats2half2_rn(u.x, u.y));
          s2 = add2(s2, u);
          q2 = fma2(u, u, q2);
        }
        if (hh == 0) {
          // the previous staged chunk's TMA store must have finished READING this staging box before it is rewritten (waited
          // for after the piece's arithmetic: see K_B)
          if (issuer) bulk_wait_read0();
          __syncwarp();
        }
#pragma unroll
        for (int k = 0; k < 4; ++k)
          st_shared_v4(st_row + (uint32_t)((((hh * 4 + k) ^ (row & 7))) << 4),
                       make_uint4(pk[4 * k], pk[4 * k + 1], pk[4 * k + 2], pk[4 * k + 3]));
        if (hh == 1) {
          fence_proxy_async();
          __syncwarp();
          if (issuer) {
            tma_store_2d(&tm_enc, st_box, half * 64, w * C_ENC + q * 128 + quad * 32);
            bulk_commit();
          }
        }
      }
      // bottleneck-gLN partial of the m-tile: slot 0 carries the sum of the tile, slot 1 zeros (the consumer adds two
      // slots per tile: the layout the CUDA-core encoder writes)
      const float s = warp_sum(s2.x + s2.y), qq = warp_sum(q2.x + q2.y);
      tile_partial_reduce(s_red + (it & 1) * Cfg::EPI_WARPS, s_cnt + (it & 1), Cfg::EPI_WARPS, ew, s, qq, part + (size_t)w * 2);
      if (threadIdx.x == 0) part[(size_t)w * 2 + 1] = make_float2(0.f, 0.f);
    }
    if (issuer) bulk_wait0();  // all tile stores complete before the CTA exits
  } else if (warp == MMA_WARP) {
    // ================================ MMA issuer ==================================================
    if (lane == 0) {
      // A = W chunk (K-major), B = frame rows (K-major): D[filter][frame], M = 128, N = 128
      constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(TILE_M >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      int it = 0;
      for (int w = blockIdx.x; w < n_tiles; w += gridDim.x, ++it) {
        const int b = it & 1;
        mbar_wait_b(tb.flags, fr_full(b), (uint32_t)(it >> 1) & 1u);
        tcgen05_fence_after();
        for (int q = 0; q < 4; ++q) {
          mbar_wait_b(tb.flags, tempty_bar(q), ((uint32_t)it & 1u) ^ 1u);
          tcgen05_fence_after();
          const uint32_t tmem_d = tmem_base + (uint32_t)(q * 128);
#pragma unroll
          for (int j = 0; j < Cfg::K2 / 16; ++j) {
            const uint64_t dw = make_smem_desc(smem_w + (uint32_t)(q * 16384 + j * 32), 16u, 1024u);
            const uint64_t dx = make_smem_desc(smem_f(b) + (uint32_t)(j * 32), 16u, 1024u);
            umma_f16(tmem_d, dw, dx, idesc, (uint32_t)(j != 0));
          }
          umma_commit(tfull_bar(q));
        }
        umma_commit(fr_empty(b));
      }
      umma_commit(done_bar);
      mbar_wait_b(tb.flags, done_bar, 0);
    }
    __syncwarp();
  } else {
    // ================================ framing: waveform -> (hi | lo) operand rows ==================
    const int f = (warp - MMA_WARP - 1) * 32 + lane;   // frame of the m-tile == operand row
    int it = 0;
    for (int w = blockIdx.x; w < n_tiles; w += gridDim.x, ++it) {
      const int b = it & 1;
      const TileInfo ti = ti_get(s_ti, tb, it, w);
      const bool valid = (ti.t0 + f) < ti.F;
      const float* x = mix + tb.seg_in_off[ti.seg] + (size_t)(ti.t0 + f) * S;
      float xv[L];
#pragma unroll
      for (int k = 0; k < L; ++k) xv[k] = valid ? __ldg(x + k) : 0.f;   // (a valid frame lies inside the segment)
      mbar_wait_b(tb.flags, fr_empty(b), ((uint32_t)(it >> 1) & 1u) ^ 1u);
      const uint32_t row = smem_f(b) + (uint32_t)(f * 128);
#pragma unroll
      for (int c = 0; c < L / 8; ++c) {
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int k = 0; k < 8; k += 2) {
          __half h0, l0, h1, l1;
          split_half(xv[c * 8 + k], h0, l0);
          split_half(xv[c * 8 + k + 1], h1, l1);
          hi[k >> 1] = h2_as_u32(__halves2half2(h0, h1));
          lo[k >> 1] = h2_as_u32(__halves2half2(l0, l1));
        }
        st_shared_v4(row + (uint32_t)(((c) ^ (f & 7)) << 4), make_uint4(hi[0], hi[1], hi[2], hi[3]));
        st_shared_v4(row + (uint32_t)(((c + L / 8) ^ (f & 7)) << 4), make_uint4(lo[0], lo[1], lo[2], lo[3]));
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(fr_full(b));
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

}  // namespace enc
}  // namespace b200sep

// Fused TCN-block kernels (sm_100a, fp16 tensor-core mode).  One Conv1DBlock is two
// launches; the kernel boundaries are exactly the two segment-wide gLN reductions of the block:
//
//   K_B  apply_in_kernel :  x += gLN2-scaled res of the previous block, skip += ... (elementwise
//                           prologue on G), x -> fp16 operand in shared memory -> tcgen05 in-conv
//                           GEMM (128 -> 512, W_in resident in shared memory) -> bias + PReLU ->
//                           h1 (fp16) + gLN1 partial sums.
//   K_A  dw_rs_kernel    :  gLN1 folded into the dilated depthwise taps + PReLU computed by 16
//                           producer warps straight into the swizzled shared-memory A-operand ring
//                           (h2 never leaves the SM) -> tcgen05 [res | skip] GEMM (512 -> 256) on the
//                           RAW h2 with gamma-folded weights -> G (fp16, raw) + gLN2 partial sums.
//
// G is "raw": the two per-segment gLN2 scalars are only known once every tile of the segment has been
// through K_A, so they are applied by the next K_B's prologue:
//     x_new = x + vm * (rstd2 * (G_res - mean2 * rs) + cb)        (rs, cb: see epilogues.cuh)
// HBM bytes per frame and block: K_A 1024 (h1) + 512 (G); K_B 512 (G) + 2048 (x, skip r/w) + 1024 (h1)
// = 5.1 KB vs 6.6 KB for the three unfused kernels, and one launch less.
//
// Every mbarrier wait is bounded (g_abort): a protocol bug degrades into wrong numbers plus an error
// flag the host reads, never into a hung GPU.
#pragma once
#include "bandwidth_kernels.cuh"
#include "common.cuh"
#include "epilogues.cuh"
#include "gemm_tc.cuh"

namespace b200sep {

__device__ int g_abort = 0;  // set when a bounded wait ran out (read by b200sep_debug_abort_flag)

namespace fb {
using namespace tc;

__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(bar), "r"(parity), "r"(2000)
      : "memory");
  return ok != 0;
}
// Bounded wait: ~seconds in the worst case, then flags the abort and lets the caller run on (all
// addresses stay in bounds; the host reports the failure).
__device__ __forceinline__ void mbar_wait_b(uint32_t bar, uint32_t parity) {
  for (uint32_t i = 0; i < (1u << 21); ++i) {
    if (mbar_try(bar, parity)) return;
    if ((i & 255u) == 255u && *reinterpret_cast<volatile int*>(&g_abort) != 0) return;
  }
  *reinterpret_cast<volatile int*>(&g_abort) = 1;
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// byte offset of the 16-byte chunk (8 consecutive frames starting at frame f0, f0 % 8 == 0) of
// k-row `row` (0..63) inside one 64(k) x 128(m) MN-major SWIZZLE_128B operand stage: two 64x64 boxes
// 8 KiB apart (frames 0-63 / 64-127), 128-byte rows, chunk index xor (row & 7) -- the image a TMA
// box load with CU_TENSOR_MAP_SWIZZLE_128B produces (gemm_tc.cuh).
__device__ __forceinline__ uint32_t a_chunk_off(int row, int f0) {
  const int chunk = (f0 >> 3) & 7;
  return (uint32_t)((f0 >> 6) * 8192 + row * 128 + ((chunk ^ (row & 7)) << 4));
}

// =================================================================================================
// K_A: depthwise stage (producer warps) + [res | skip] GEMM -> raw G
// =================================================================================================
template <int BN>
struct KaCfg {
  static constexpr int STAGES = 4;
  static constexpr int A_STAGE = TILE_M * 64 * 2;  // 16 KiB
  static constexpr int B_STAGE = BN * 64 * 2;      // 32 / 16 KiB
  static constexpr int STAGE = A_STAGE + 32768;
  static constexpr int PAR_BYTES = C_HID * 32;     // per-channel folded taps (2 x float4)
  static constexpr int EPI_WARPS = 4, PROD_WARPS = 16;
  static constexpr int THREADS = (EPI_WARPS + 2 + PROD_WARPS) * 32;  // 704
  static constexpr int TMEM_COLS = 2 * BN;
  static constexpr int SMEM_BYTES = STAGES * STAGE + PAR_BYTES + 1024 /*align*/ + 256 /*barriers*/;
  static constexpr int SLOTS = PROD_WARPS;         // gLN2 partial slots per m-tile
};

// register image of one producer step: the centre vector and the two shifted / neighbouring vectors
struct DwLd { uint4 c, a, b; };

__device__ __forceinline__ uint4 ldg_v4(const __half* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ uint4 shfl_up16_v4(uint4 v) {
  uint4 r;
  r.x = __shfl_up_sync(0xffffffffu, v.x, 1, 16); r.y = __shfl_up_sync(0xffffffffu, v.y, 1, 16);
  r.z = __shfl_up_sync(0xffffffffu, v.z, 1, 16); r.w = __shfl_up_sync(0xffffffffu, v.w, 1, 16);
  return r;
}
__device__ __forceinline__ uint4 shfl_down16_v4(uint4 v) {
  uint4 r;
  r.x = __shfl_down_sync(0xffffffffu, v.x, 1, 16); r.y = __shfl_down_sync(0xffffffffu, v.y, 1, 16);
  r.z = __shfl_down_sync(0xffffffffu, v.z, 1, 16); r.w = __shfl_down_sync(0xffffffffu, v.w, 1, 16);
  return r;
}
__device__ __forceinline__ void unpack8(const uint4 v, float (&o)[8]) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
    o[2 * i] = f.x; o[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float (&o)[8]) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __half2 h = __floats2half2_rn(o[2 * i], o[2 * i + 1]);
    w[i] = *reinterpret_cast<const uint32_t*>(&h);
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}
// left / right tap values of the thread's 8 frames
template <int D>
__device__ __forceinline__ void dw_taps(const DwLd& in, int lane16, float (&ctr)[8], float (&lft)[8], float (&rgt)[8]) {
  unpack8(in.c, ctr);
  if (D == 0) {
    unpack8(in.a, lft);
    unpack8(in.b, rgt);
  } else {
    uint4 vp = shfl_up16_v4(in.c), vn = shfl_down16_v4(in.c);
    if (lane16 == 0) vp = in.a;
    if (lane16 == 15) vn = in.b;
    float pv[8], nv[8];
    unpack8(vp, pv);
    unpack8(vn, nv);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      lft[i] = (i - D >= 0) ? ctr[(i - D) & 7] : pv[(8 + i - D) & 7];
      rgt[i] = (i + D < 8) ? ctr[(i + D) & 7] : nv[(i + D - 8) & 7];
    }
  }
}
template <int D, bool A01>
__device__ __forceinline__ uint4 dw_compute_interior(const DwLd& in, const float4 pa, int lane16, float alpha,
                                                     float2& s2, float2& q2) {
  float ctr[8], lft[8], rgt[8];
  dw_taps<D>(in, lane16, ctr, lft, rgt);
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
  return pack8(o);
}
template <int D, bool A01>
__device__ __forceinline__ uint4 dw_compute_edge(const DwLd& in, const float4 pa, const float4 pb,
                                                 const DwEdgeMasks& mk, int lane16, float alpha, float2& s2,
                                                 float2& q2) {
  float ctr[8], lft[8], rgt[8];
  dw_taps<D>(in, lane16, ctr, lft, rgt);
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
  return pack8(o);
}
// 8 frames [t, t+8) of channel c of the segment, zeros outside the segment's m-tiles
__device__ __forceinline__ uint4 ld_group8(const __half* __restrict__ h1, int tile0, int ntiles, int c, int t) {
  if (t < 0 || t >= ntiles * TILE_M) return make_uint4(0, 0, 0, 0);
  return ldg_v4(h1 + tt_index(tile0 + (t >> 7), C_HID, c, t & (TILE_M - 1)));
}

template <int BN, int D, bool A01>
__global__ void __launch_bounds__(704, 1)
dw_rs_kernel(const __grid_constant__ CUtensorMap tm_b, const __half* __restrict__ h1, __half* __restrict__ g,
             const float2* __restrict__ part1, int slots1, float2* __restrict__ part2, DwParams p, Tables tb) {
  using Cfg = KaCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  auto stage_a = [&](int s) { return smem_base + (uint32_t)(s * Cfg::STAGE); };
  auto stage_b = [&](int s) { return smem_base + (uint32_t)(s * Cfg::STAGE + Cfg::A_STAGE); };
  const uint32_t par0 = smem_base + Cfg::STAGES * Cfg::STAGE;
  const uint32_t bar0 = par0 + Cfg::PAR_BYTES;
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (Cfg::STAGES + s); };
  auto tfull_bar = [&](int s) { return bar0 + 8u * (2 * Cfg::STAGES + s); };
  auto tempty_bar = [&](int s) { return bar0 + 8u * (2 * Cfg::STAGES + 2 + s); };
  const uint32_t tmem_ptr_smem = bar0 + 8u * (2 * Cfg::STAGES + 4);
  float4* s_par = reinterpret_cast<float4*>(smem_raw + (par0 - smem_u32(smem_raw)));  // [512][2]

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int n_tiles = tb.n_tiles;

  if (warp == 4 && lane == 0) tma_prefetch_desc(&tm_b);
  if (warp == 5) {
    if (lane == 0) {
      for (int s = 0; s < Cfg::STAGES; ++s) {
        mbar_init(full_bar(s), Cfg::PROD_WARPS * 32 + 1);
        mbar_init(empty_bar(s), 1);
      }
      for (int s = 0; s < 2; ++s) {
        mbar_init(tfull_bar(s), 1);
        mbar_init(tempty_bar(s), 32 * Cfg::EPI_WARPS);
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_ptr_smem, Cfg::TMEM_COLS);
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_smem));

  if (warp < 4) {
    // ================================ epilogue: TMEM -> fp16 G ====================================
    const int quad = warp;
    const int r = quad * 32 + lane;
    constexpr int NCH = BN / 32;
    int it = 0;
    for (int w = blockIdx.x; w < n_tiles; w += gridDim.x, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
      mbar_wait_b(tfull_bar(as), aphase);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * BN);
      __half* __restrict__ dst = g + tt_index(w, BN, 0, r);
      float v[2][32];
      tmem_ld32(taddr, v[0]);
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        tmem_ld_wait();
        if (c + 1 < NCH) tmem_ld32(taddr + (uint32_t)((c + 1) * 32), v[(c + 1) & 1]);
#pragma unroll
        for (int j = 0; j < 32; ++j) dst[(size_t)(c * 32 + j) * TILE_M] = __float2half_rn(v[c & 1][j]);
      }
      tcgen05_fence_before();
      mbar_arrive(tempty_bar(as));
    }
  } else if (warp == 4) {
    // ================================ TMA producer (weights) ======================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int w = blockIdx.x; w < n_tiles; w += gridDim.x) {
        for (int kb = 0; kb < 8; ++kb) {
          mbar_wait_b(empty_bar(stage), phase ^ 1u);
          mbar_expect_tx(full_bar(stage), Cfg::B_STAGE);
          tma_load_2d(stage_b(stage), &tm_b, full_bar(stage), kb * 64, 0);
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 5) {
    // ================================ MMA issuer ==================================================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BN);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int w = blockIdx.x; w < n_tiles; w += gridDim.x, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
        mbar_wait_b(tempty_bar(as), aphase ^ 1u);
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(as * BN);
        for (int kb = 0; kb < 8; ++kb) {
          mbar_wait_b(full_bar(stage), phase);
          tcgen05_fence_after();
          const uint32_t sa = stage_a(stage), sb = stage_b(stage);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint64_t da = make_smem_desc(sa + j * 2048u, 8192u, 1024u);
            const uint64_t db = make_smem_desc(sb + j * 32u, 16u, 1024u);
            umma_f16(tmem_d, da, db, idesc, (uint32_t)((kb | j) != 0));
          }
          umma_commit(empty_bar(stage));
          if (kb == 7) umma_commit(tfull_bar(as));
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else {
    // ================================ depthwise producers =========================================
    const int pw = warp - 6;              // 0..15
    const int ptid = pw * 32 + lane;      // 0..511
    const int sub = lane >> 4, lane16 = lane & 15;
    const int c0 = pw * 2 + sub;          // channel of step 0; step n handles channel c0 + 32 n
    const int d = (D > 0) ? D : p.dilation;
    const float alpha = p.alpha2;
    int stage = 0;
    uint32_t phase = 0;
    for (int w = blockIdx.x; w < n_tiles; w += gridDim.x) {
      const TileInfo ti = load_tile_info(tb.tile_info + w);
      const int tile0 = ti.tile0, ntiles = ti.ntiles, F = ti.F, tile_t0 = ti.t0;
      float mean, rstd;
      warp_seg_stats(part1, tile0, ntiles, slots1, ti.inv_count, mean, rstd);
      named_bar_sync(1, Cfg::PROD_WARPS * 32);  // nobody still reads the previous tile's table
      {
        const int c = ptid;
        const float a = __ldg(p.gamma1 + c) * rstd;
        const float b = __ldg(p.beta1 + c) - a * mean;  // y = a*h + b
        const float w0 = __ldg(p.w + c * 3 + 0), w1 = __ldg(p.w + c * 3 + 1), w2 = __ldg(p.w + c * 3 + 2);
        const float bias = __ldg(p.bias + c);
        s_par[c * 2 + 0] = make_float4(w0 * a, w1 * a, w2 * a, bias + (w0 + w1 + w2) * b);
        s_par[c * 2 + 1] = make_float4(w0 * b, w1 * b, w2 * b, bias);
      }
      named_bar_sync(1, Cfg::PROD_WARPS * 32);

      const int t_base = tile_t0 + lane16 * 8;
      const bool interior = (tile_t0 - d >= 0) && (tile_t0 + TILE_M - 1 + d < F);
      float2 s2 = make_float2(0.f, 0.f), q2 = make_float2(0.f, 0.f);
      if (interior) {
        auto ptr_at = [&](int t) { return h1 + tt_index(tile0 + (t >> 7), C_HID, c0, t & (TILE_M - 1)); };
        const __half* pc = ptr_at(t_base);
        const __half* pa_ = (D == 0) ? ptr_at(t_base - d) : ((lane16 == 0) ? ptr_at(t_base - 8) : pc);
        const __half* pb_ = (D == 0) ? ptr_at(t_base + d) : ((lane16 == 15) ? ptr_at(t_base + 8) : pc);
        constexpr size_t STRIDE = (size_t)32 * TILE_M;
        auto load = [&](int n) {
          DwLd r;
          r.c = ldg_v4(pc + n * STRIDE);
          if (D == 0 || lane16 == 0) r.a = ldg_v4(pa_ + n * STRIDE); else r.a = r.c;
          if (D == 0 || lane16 == 15) r.b = ldg_v4(pb_ + n * STRIDE); else r.b = r.c;
          return r;
        };
        DwLd cur = load(0);
#pragma unroll 2
        for (int n = 0; n < 16; ++n) {
          const DwLd nxt = load(n < 15 ? n + 1 : 15);
          if ((n & 1) == 0) mbar_wait_b(empty_bar(stage), phase ^ 1u);
          const int row = (n & 1) * 32 + c0;
          const float4 pa = s_par[(c0 + n * 32) * 2];
          const uint4 o = dw_compute_interior<D, A01>(cur, pa, lane16, alpha, s2, q2);
          st_shared_v4(stage_a(stage) + a_chunk_off(row, lane16 * 8), o);
          if (n & 1) {
            fence_proxy_async();
            mbar_arrive(full_bar(stage));
            if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
          }
          cur = nxt;
        }
      } else {
        DwEdgeMasks mk;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int ta = t_base + 2 * k, tb2 = ta + 1;
          mk.ml[k] = make_float2((ta - d >= 0 && ta - d < F) ? 1.f : 0.f, (tb2 - d >= 0 && tb2 - d < F) ? 1.f : 0.f);
          mk.mc[k] = make_float2(ta < F ? 1.f : 0.f, tb2 < F ? 1.f : 0.f);
          mk.mr[k] = make_float2(ta + d < F ? 1.f : 0.f, tb2 + d < F ? 1.f : 0.f);
        }
        auto load = [&](int n) {
          const int c = c0 + n * 32;
          DwLd r;
          r.c = ld_group8(h1, tile0, ntiles, c, t_base);
          if (D == 0) {
            r.a = ld_group8(h1, tile0, ntiles, c, t_base - d);
            r.b = ld_group8(h1, tile0, ntiles, c, t_base + d);
          } else {
            r.a = (lane16 == 0) ? ld_group8(h1, tile0, ntiles, c, t_base - 8) : r.c;
            r.b = (lane16 == 15) ? ld_group8(h1, tile0, ntiles, c, t_base + 8) : r.c;
          }
          return r;
        };
        DwLd cur = load(0);
#pragma unroll 1
        for (int n = 0; n < 16; ++n) {
          const DwLd nxt = load(n < 15 ? n + 1 : 15);
          if ((n & 1) == 0) mbar_wait_b(empty_bar(stage), phase ^ 1u);
          const int row = (n & 1) * 32 + c0;
          const float4 pa = s_par[(c0 + n * 32) * 2], pb = s_par[(c0 + n * 32) * 2 + 1];
          const uint4 o = dw_compute_edge<D, A01>(cur, pa, pb, mk, lane16, alpha, s2, q2);
          st_shared_v4(stage_a(stage) + a_chunk_off(row, lane16 * 8), o);
          if (n & 1) {
            fence_proxy_async();
            mbar_arrive(full_bar(stage));
            if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
          }
          cur = nxt;
        }
      }
      const float s_acc = warp_sum(s2.x + s2.y);
      const float q_acc = warp_sum(q2.x + q2.y);
      if (lane == 0) part2[(size_t)w * Cfg::SLOTS + pw] = make_float2(s_acc, q_acc);
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 5) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// =================================================================================================
// K_B: apply previous block's G (x += res, skip += skip) + in-conv GEMM -> h1
// =================================================================================================
struct KbCfg {
  static constexpr int W_BYTES = C_HID * C_BN * 2;      // 128 KiB: [2 k-blocks][512 rows][64 k]
  static constexpr int XN_BYTES = TILE_M * C_BN * 2;    // 32 KiB: [2 k-blocks][64 k][128 m]
  static constexpr int EPI_WARPS = 8, PRO_WARPS = 8;
  static constexpr int THREADS = (EPI_WARPS + 1 + PRO_WARPS) * 32;  // 544
  static constexpr int TMEM_COLS = 512;                 // 4 chunks of 128 output channels
  static constexpr int SMEM_BYTES = W_BYTES + 2 * XN_BYTES + 1024 + 256;
  static constexpr int SLOTS = EPI_WARPS;               // gLN1 partial slots per m-tile
};

struct ApplyParams {
  const __half* g;      // [nt][g_ch][128] raw [res | skip] of the previous block (nullptr: block 0)
  const float2* part2;  // gLN2 partials of the previous block
  int slots2;
  const float* rs;      // [256] rowsum(W . gamma2) of the previous block
  const float* cb;      // [256] W * beta2 + b
  float* x;             // [nt][128][128] fp32 residual stream (updated in place)
  float* skip;          // [nt][128][128] fp32 skip accumulator (updated in place)
  int first;            // previous block was block 0: skip has no old value
};

template <bool A01>
__global__ void __launch_bounds__(544, 1)
apply_in_kernel(const __grid_constant__ CUtensorMap tm_w, ApplyParams ap, EpiIn<__half, A01> epi, Tables tb) {
  using Cfg = KbCfg;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_w = smem_base;
  auto smem_xn = [&](int b) { return smem_base + (uint32_t)(Cfg::W_BYTES + b * Cfg::XN_BYTES); };
  const uint32_t bar0 = smem_base + Cfg::W_BYTES + 2 * Cfg::XN_BYTES;
  const uint32_t w_bar = bar0;
  auto xn_full = [&](int b) { return bar0 + 8u * (1 + b); };
  auto xn_empty = [&](int b) { return bar0 + 8u * (3 + b); };
  auto tfull_bar = [&](int q) { return bar0 + 8u * (5 + q); };
  auto tempty_bar = [&](int q) { return bar0 + 8u * (9 + q); };
  const uint32_t tmem_ptr_smem = bar0 + 8u * 13;

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int n_tiles = tb.n_tiles;

  if (warp == 8) {
    if (lane == 0) {
      tma_prefetch_desc(&tm_w);
      mbar_init(w_bar, 1);
      for (int b = 0; b < 2; ++b) {
        mbar_init(xn_full(b), Cfg::PRO_WARPS * 32);
        mbar_init(xn_empty(b), 1);
      }
      for (int q = 0; q < 4; ++q) {
        mbar_init(tfull_bar(q), 1);
        mbar_init(tempty_bar(q), Cfg::EPI_WARPS * 32);
      }
      fence_barrier_init();
      // resident weights: k-block kb at +64 KiB, rows 0-255 / 256-511 as two boxes of 64(k) x 256
      mbar_expect_tx(w_bar, Cfg::W_BYTES);
      for (int kb = 0; kb < 2; ++kb)
        for (int hb = 0; hb < 2; ++hb)
          tma_load_2d(smem_w + (uint32_t)(kb * 65536 + hb * 32768), &tm_w, w_bar, kb * 64, hb * 256);
    }
    __syncwarp();
    tmem_alloc(tmem_ptr_smem, Cfg::TMEM_COLS);
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_smem));

  if (warp < 8) {
    // ================================ epilogue: bias + PReLU -> h1, gLN1 partials ==================
    const int ew = warp, quad = ew & 3, half = ew >> 2;
    const int r = quad * 32 + lane;
    int it = 0;
    for (int w = blockIdx.x; w < n_tiles; w += gridDim.x, ++it) {
      EpiState st;
      epi.tile_begin(st, w, tb);
      const uint32_t ph = (uint32_t)it & 1u;
      float pre[32];  // unused by EpiIn
#pragma unroll 1
      for (int q = 0; q < 4; ++q) {
        mbar_wait_b(tfull_bar(q), ph);
        tcgen05_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(q * 128 + half * 64);
        const int n_base = q * 128 + half * 64;
        float v[2][32];
        tmem_ld32(taddr, v[0]);
        tmem_ld32(taddr + 32u, v[1]);
        tmem_ld_wait();
        tcgen05_fence_before();
        mbar_arrive(tempty_bar(q));  // accumulator chunk is in registers: the MMA may overwrite it
        epi.template apply<32>(st, w, r, n_base, v[0], pre);
        epi.template apply<32>(st, w, r, n_base + 32, v[1], pre);
      }
      const float s = warp_sum(st.s), qq = warp_sum(st.q);
      if (lane == 0) epi.part_out[(size_t)w * Cfg::SLOTS + ew] = make_float2(s, qq);
    }
  } else if (warp == 8) {
    // ================================ MMA issuer ==================================================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(128);
      mbar_wait_b(w_bar, 0);
      int it = 0;
      for (int w = blockIdx.x; w < n_tiles; w += gridDim.x, ++it) {
        const int b = it & 1;
        mbar_wait_b(xn_full(b), (uint32_t)(it >> 1) & 1u);
        tcgen05_fence_after();
        const uint32_t sa = smem_xn(b);
        for (int q = 0; q < 4; ++q) {
          mbar_wait_b(tempty_bar(q), ((uint32_t)it & 1u) ^ 1u);
          tcgen05_fence_after();
          const uint32_t tmem_d = tmem_base + (uint32_t)(q * 128);
#pragma unroll
          for (int kb = 0; kb < 2; ++kb) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint64_t da = make_smem_desc(sa + (uint32_t)(kb * 16384 + j * 2048), 8192u, 1024u);
              const uint64_t db = make_smem_desc(smem_w + (uint32_t)(kb * 65536 + q * 16384 + j * 32), 16u, 1024u);
              umma_f16(tmem_d, da, db, idesc, (uint32_t)((kb | j) != 0));
            }
          }
          umma_commit(tfull_bar(q));
        }
        umma_commit(xn_empty(b));
      }
    }
    __syncwarp();
  } else {
    // ================================ prologue: apply G, build the fp16 operand ===================
    const int pt = (warp - 9) * 32 + lane;  // 0..255
    const int lane16 = pt & 15, crow = pt >> 4;
    const int f0 = lane16 * 8;
    const bool has_g = ap.g != nullptr;
    int it = 0;
    for (int w = blockIdx.x; w < n_tiles; w += gridDim.x, ++it) {
      const int b = it & 1;
      const TileInfo ti = load_tile_info(tb.tile_info + w);
      float mean = 0.f, rstd = 0.f;
      if (has_g) warp_seg_stats(ap.part2, ti.tile0, ti.ntiles, ap.slots2, ti.inv_count, mean, rstd);
      float2 vm[4];
#pragma unroll
      for (int k = 0; k < 4; ++k)
        vm[k] = make_float2((ti.t0 + f0 + 2 * k) < ti.F ? 1.f : 0.f, (ti.t0 + f0 + 2 * k + 1) < ti.F ? 1.f : 0.f);
      const float nmr = -mean * rstd;
      const float2 rstd2 = make_float2(rstd, rstd);
      mbar_wait_b(xn_empty(b), ((uint32_t)(it >> 1) & 1u) ^ 1u);
      const uint32_t xn = smem_xn(b);
      float* __restrict__ xt = ap.x + tt_index(w, C_BN, 0, f0);
      float* __restrict__ st_ = ap.skip + tt_index(w, C_BN, 0, f0);
      const __half* __restrict__ gt = has_g ? ap.g + tt_index(w, 2 * C_BN, 0, f0) : nullptr;
#pragma unroll 2
      for (int ps = 0; ps < 8; ++ps) {
        const int c = ps * 16 + crow;
        const size_t off = (size_t)c * TILE_M;
        float4 xa = *reinterpret_cast<const float4*>(xt + off), xb = *reinterpret_cast<const float4*>(xt + off + 4);
        float xv[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
        if (has_g) {
          const uint4 gr = ldg_v4(gt + off);
          const uint4 gs = ldg_v4(gt + (size_t)C_BN * TILE_M + off);
          float sv[8];
          if (!ap.first) {
            const float4 sa4 = *reinterpret_cast<const float4*>(st_ + off), sb4 = *reinterpret_cast<const float4*>(st_ + off + 4);
            sv[0] = sa4.x; sv[1] = sa4.y; sv[2] = sa4.z; sv[3] = sa4.w; sv[4] = sb4.x; sv[5] = sb4.y; sv[6] = sb4.z; sv[7] = sb4.w;
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) sv[i] = 0.f;
          }
          const float kr = fmaf(nmr, __ldg(ap.rs + c), __ldg(ap.cb + c));
          const float ks = fmaf(nmr, __ldg(ap.rs + C_BN + c), __ldg(ap.cb + C_BN + c));
          float gv[8], hv[8];
          unpack8(gr, gv);
          unpack8(gs, hv);
#pragma unroll
          for (int i = 0; i < 8; i += 2) {
            const float2 dr = fma2(make_float2(gv[i], gv[i + 1]), rstd2, make_float2(kr, kr));
            const float2 ds = fma2(make_float2(hv[i], hv[i + 1]), rstd2, make_float2(ks, ks));
            const float2 nx = fma2(dr, vm[i >> 1], make_float2(xv[i], xv[i + 1]));
            const float2 ns = fma2(ds, vm[i >> 1], make_float2(sv[i], sv[i + 1]));
            xv[i] = nx.x; xv[i + 1] = nx.y;
            sv[i] = ns.x; sv[i + 1] = ns.y;
          }
          *reinterpret_cast<float4*>(xt + off) = make_float4(xv[0], xv[1], xv[2], xv[3]);
          *reinterpret_cast<float4*>(xt + off + 4) = make_float4(xv[4], xv[5], xv[6], xv[7]);
          *reinterpret_cast<float4*>(st_ + off) = make_float4(sv[0], sv[1], sv[2], sv[3]);
          *reinterpret_cast<float4*>(st_ + off + 4) = make_float4(sv[4], sv[5], sv[6], sv[7]);
        }
        st_shared_v4(xn + (uint32_t)((c >> 6) * 16384) + a_chunk_off(c & 63, f0), pack8(xv));
      }
      fence_proxy_async();
      mbar_arrive(xn_full(b));
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 8) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// =================================================================================================
// After the last block: skip_sum += gLN2-scaled skip of the last block, PReLU (mask_net.0), fp16
// operand of the mask GEMM.  grid = n_tiles, 256 threads: thread = (channel row, 8 frames).
// =================================================================================================
__global__ void __launch_bounds__(256) final_skip_kernel(const __half* __restrict__ g,  // [nt][128][128]
                                                         const float2* __restrict__ part2, int slots2,
                                                         const float* __restrict__ rs, const float* __restrict__ cb,
                                                         const float* __restrict__ skip, __half* __restrict__ skipa,
                                                         float alpha_mask, int first, Tables tb) {
  const int w = blockIdx.x;
  const TileInfo ti = load_tile_info(tb.tile_info + w);
  float mean, rstd;
  warp_seg_stats(part2, ti.tile0, ti.ntiles, slots2, ti.inv_count, mean, rstd);
  const int lane16 = threadIdx.x & 15, crow = threadIdx.x >> 4;
  const int f0 = lane16 * 8;
  const float nmr = -mean * rstd;
#pragma unroll 2
  for (int ps = 0; ps < 8; ++ps) {
    const int c = ps * 16 + crow;
    const size_t idx = tt_index(w, C_SKIP, c, f0);
    float gv[8], sv[8], o[8];
    unpack8(ldg_v4(g + idx), gv);
    if (!first) {
      const float4 a = *reinterpret_cast<const float4*>(skip + idx), b = *reinterpret_cast<const float4*>(skip + idx + 4);
      sv[0] = a.x; sv[1] = a.y; sv[2] = a.z; sv[3] = a.w; sv[4] = b.x; sv[5] = b.y; sv[6] = b.z; sv[7] = b.w;
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) sv[i] = 0.f;
    }
    const float ks = fmaf(nmr, __ldg(rs + c), __ldg(cb + c));
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float vmask = (ti.t0 + f0 + i) < ti.F ? 1.f : 0.f;
      const float v = fmaf(fmaf(gv[i], rstd, ks), vmask, sv[i]);
      o[i] = fmaxf(v, 0.f) + alpha_mask * fminf(v, 0.f);
    }
    *reinterpret_cast<uint4*>(skipa + idx) = pack8(o);
  }
}

}  // namespace fb
}  // namespace b200sep

// tcgen05 / TMEM / TMA GEMM for the 1x1 convolutions (sm_100a):
//     D[frame][n] = sum_k A[frame][k] * B[n][k],   fp16 operands, fp32 accumulation in TMEM.
//
//   A : activations, tile-transposed [m_tile][K][128] fp16  -> "MN-major" UMMA operand.  One
//       64(k) x 128(m) pipeline stage is fetched as two TMA boxes of 64(k) rows x 64(m) elements
//       (128 bytes, SWIZZLE_128B); the two boxes are the two M-atoms (LBO = 8 KiB apart), 8-row
//       K groups are SBO = 1 KiB apart.
//   B : weights [N][K] fp16, K-major; one stage = TMA box of BN rows x 64 k (SWIZZLE_128B),
//       SBO = 1 KiB between 8-row groups, the K=16 sub-steps advance the start address by 32 B.
//   D : 128 lanes x BN fp32 columns in TMEM, double-buffered (2*BN columns) so the epilogue of
//       work item i overlaps the MMAs of work item i+1.
//
// Warp roles (384 threads, 1 CTA / SM, persistent over (m_tile, n_tile) work items):
//   warp 0 : TMA producer (one elected lane)      warp 2 : TMEM allocator
//   warp 1 : tcgen05.mma issuer (one lane)        warps 4-11 : epilogue (lane = frame, tcgen05.ld);
//   two epilogue warps share each 32-lane TMEM quadrant and split the BN columns, so every SM
//   sub-partition has two warps to interleave; global operands of the epilogue (old x / skip, the
//   encoder output) are prefetched one 32-column chunk ahead, across work items.
// Pipelines: smem full/empty mbarriers (TMA <-> MMA), TMEM full/empty mbarriers (MMA <-> epilogue).
#pragma once
#include "common.cuh"
#include "epilogues.cuh"

namespace b200sep {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n\t"
      "@P1 bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(bar), "r"(parity), "r"(0x989680)
      : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const void* desc) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(desc)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const void* desc, uint32_t bar, int c0,
                                            int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(desc)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tcgen05_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]; kind::f16 covers fp16/bf16 inputs with fp32 accumulate.
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrive when all previously issued tcgen05.mma of this thread have completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread i of the warp receives lane (taddr.lane + i).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t* u = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
        "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]),
        "=r"(u[16]), "=r"(u[17]), "=r"(u[18]), "=r"(u[19]), "=r"(u[20]), "=r"(u[21]), "=r"(u[22]), "=r"(u[23]),
        "=r"(u[24]), "=r"(u[25]), "=r"(u[26]), "=r"(u[27]), "=r"(u[28]), "=r"(u[29]), "=r"(u[30]), "=r"(u[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory matrix descriptor (sm_100 format: version = 1, SWIZZLE_128B = 2 in [61,64)).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;
  d |= 2ull << 61;
  return d;
}

// Instruction descriptor for kind::f16: D = fp32, A = B = fp16, A MN-major, B K-major, M = 128.
__host__ __device__ constexpr uint32_t make_idesc(int bn) {
  return (1u << 4) | (0u << 7) | (0u << 10) | (1u << 15) | (0u << 16) | ((uint32_t)(bn >> 3) << 17) |
         ((uint32_t)(TILE_M >> 4) << 24);
}

template <int BN, int K>
struct TcConfig {
  static constexpr int KB = K / 64;                 // 64-wide k-blocks per work item
  static constexpr int A_STAGE = TILE_M * 64 * 2;   // 16 KiB
  static constexpr int B_STAGE = BN * 64 * 2;       // 16 / 32 KiB
  static constexpr int STAGES = (BN == 256) ? 4 : 6;
  static constexpr int TMEM_COLS = 2 * BN;          // double-buffered accumulator
  static constexpr int SMEM_BYTES = STAGES * (A_STAGE + B_STAGE) + 1024 /*align slack*/ + 512 /*barriers, stats*/;
  static constexpr int THREADS = 384;
  static constexpr int EPI_WARPS = 8;
  static constexpr int NCH = BN / 64;               // 32-column chunks per epilogue warp
};

template <int BN, int K, class Epi>
__global__ void __launch_bounds__(384, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
               int n_mtiles, int n_ntiles, uint32_t a_lbo, uint32_t a_sbo, Epi epi, Tables tb) {
  using Cfg = TcConfig<BN, K>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B needs 1 KiB alignment
  const uint32_t smem_a0 = smem_base;
  const uint32_t smem_b0 = smem_base + Cfg::STAGES * Cfg::A_STAGE;
  const uint32_t bar0 = smem_b0 + Cfg::STAGES * Cfg::B_STAGE;
  // barrier layout (8 B each): full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2], then tmem ptr
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (Cfg::STAGES + s); };
  auto tfull_bar = [&](int s) { return bar0 + 8u * (2 * Cfg::STAGES + s); };
  auto tempty_bar = [&](int s) { return bar0 + 8u * (2 * Cfg::STAGES + 2 + s); };
  const uint32_t tmem_ptr_smem = bar0 + 8u * (2 * Cfg::STAGES + 4);
  const uint32_t done_bar = bar0 + 8u * (2 * Cfg::STAGES + 5);  // final tcgen05.commit (see below)
  const uint32_t red_smem = bar0 + 256u;  // [2][8] float2 block-level stat reduction

  // warp index broadcast through a shuffle so the compiler can prove it warp-uniform (keeps the
  // role branches and the memory descriptors on the uniform datapath)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int total = n_mtiles * n_ntiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 32 * Cfg::EPI_WARPS);
    }
    mbar_init(done_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_ptr_smem, Cfg::TMEM_COLS);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  // the set-up above touched no global data; from here on we read what the previous kernel wrote
  pdl_enter();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_smem));

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int w = blockIdx.x; w < total; w += gridDim.x) {
        const int m_tile = w / n_ntiles, n_tile = w % n_ntiles;
        for (int kb = 0; kb < Cfg::KB; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          mbar_expect_tx(full_bar(stage), Cfg::A_STAGE + Cfg::B_STAGE);
          const uint32_t sa = smem_a0 + stage * Cfg::A_STAGE;
          const uint32_t sb = smem_b0 + stage * Cfg::B_STAGE;
          const int arow = m_tile * K + kb * 64;
          tma_load_2d(sa, &tm_a, full_bar(stage), 0, arow);
          tma_load_2d(sa + Cfg::A_STAGE / 2, &tm_a, full_bar(stage), 64, arow);
          tma_load_2d(sb, &tm_b, full_bar(stage), kb * 64, n_tile * BN);
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================================ MMA issuer ==================================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BN);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int w = blockIdx.x; w < total; w += gridDim.x, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
        mbar_wait(tempty_bar(as), aphase ^ 1u);  // epilogue has drained this accumulator
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(as * BN);
        for (int kb = 0; kb < Cfg::KB; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tcgen05_fence_after();
          const uint32_t sa = smem_a0 + stage * Cfg::A_STAGE;
          const uint32_t sb = smem_b0 + stage * Cfg::B_STAGE;
#pragma unroll
          for (int j = 0; j < 4; ++j) {  // UMMA_K = 16 -> 4 MMAs per 64-wide k-block
            const uint64_t da = make_smem_desc(sa + j * 2048u, a_lbo, a_sbo);  // 16 k-rows x 128 B
            const uint64_t db = make_smem_desc(sb + j * 32u, 16u, 1024u);      // 16 k x 2 B inside the atom
            umma_f16(tmem_d, da, db, idesc, (uint32_t)((kb | j) != 0));
          }
          umma_commit(empty_bar(stage));  // smem slot reusable once these MMAs retire
          if (kb == Cfg::KB - 1) umma_commit(tfull_bar(as));
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
        }
      }
      // The last empty-barrier arrivals have no waiter; a CTA must not exit while an asynchronous mbarrier
      // arrive of tcgen05.commit is still in flight (it would land in the shared memory of the next CTA).
      umma_commit(done_bar);
      mbar_wait(done_bar, 0);
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ================================ epilogue ====================================
    const int ew = warp - 4;       // 0..7
    const int quad = ew & 3;       // == warp % 4 -> TMEM lanes [32*quad, 32*quad + 32)
    const int half = ew >> 2;      // which half of the BN columns
    const int r = quad * 32 + lane;
    constexpr int NCH = Cfg::NCH;
    float pre[2][32];
    int it = 0;
    int w = blockIdx.x;
    if (w < total) {
      const int m_tile = w / n_ntiles, n_tile = w % n_ntiles;
      epi.template prefetch<32>(m_tile, r, n_tile * BN + half * (BN / 2), pre[0]);
    }
    for (; w < total; w += gridDim.x, ++it) {
      const int m_tile = w / n_ntiles, n_tile = w % n_ntiles;
      const int as = it & 1;
      const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
      EpiState st;
      epi.tile_begin(st, m_tile, tb);  // overlaps the MMAs of this work item
      mbar_wait(tfull_bar(as), aphase);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * BN + half * (BN / 2));
      const int n_base = n_tile * BN + half * (BN / 2);
      float v[2][32];
      tmem_ld32(taddr, v[0]);
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        if (c + 1 < NCH) {
          epi.template prefetch<32>(m_tile, r, n_base + (c + 1) * 32, pre[(c + 1) & 1]);
        } else {
          const int wn = w + gridDim.x;  // first chunk of the next work item
          if (wn < total) epi.template prefetch<32>(wn / n_ntiles, r, (wn % n_ntiles) * BN + half * (BN / 2), pre[0]);
        }
        tmem_ld_wait();  // chunk c has landed in v[c & 1]
        if (c + 1 < NCH) tmem_ld32(taddr + (uint32_t)((c + 1) * 32), v[(c + 1) & 1]);  // overlaps apply(c)
        epi.template apply<32>(st, m_tile, r, n_base + c * 32, v[c & 1], pre[c & 1]);
      }
      tcgen05_fence_before();
      mbar_arrive(tempty_bar(as));
      if (Epi::kStats) {
        // block-level reduction of the 8 epilogue warps -> one partial per (m_tile, n_tile)
        const float s = warp_sum(st.s), q = warp_sum(st.q);
        const uint32_t slot = red_smem + (uint32_t)(((it & 1) * 8 + ew) * 8);
        if (lane == 0) asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(slot), "f"(s), "f"(q) : "memory");
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (ew == 0 && lane == 0) {
          float ts = 0.f, tq = 0.f;
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            float a, b;
            asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(a), "=f"(b) : "r"(red_smem + (uint32_t)(((it & 1) * 8 + k) * 8)));
            ts += a; tq += b;
          }
          st.s = ts; st.q = tq;
          epi.tile_end(st, m_tile, n_ntiles, n_tile);
        }
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

}  // namespace tc
}  // namespace b200sep

// K_A as a CTA pair (tcgen05 cta_group::2) with the [res | skip] weights RESIDENT IN TENSOR MEMORY.
//
// Why (profiles/r01_ncu_fused_v5_summary.txt): the single-CTA K_A streams the whole 256 KiB weight
// matrix from L2 into shared memory for EVERY m-tile -- 256 KiB + 128..384 KiB of raw h1 per tile and SM is more than
// the ~43 B/cycle/SM the L2 can deliver next to the 4096-cycle MMA time of a tile, and the same bytes cross the shared
// memory twice more (stage write, operand read).  Here
//   * a CTA keeps its 128 rows of W (rank 0: residual conv, rank 1: skip conv; 128 x 512 fp16 = 256 tensor-memory
//     columns) for the whole launch and feeds them to the MMA as the A operand straight from tensor memory
//     (tcgen05.mma [d], [a_tmem], b_desc): zero weight traffic after the prologue;
//   * the depthwise producers of each CTA write their m-tile's h2 k-blocks as the MN-major B operand
//     (the same swizzled image as before); one MMA of the pair (M = 256 channels, N = 256 frames, K = 16) reads the two
//     CTAs' stages, so an SM's tensor core pulls 16 KiB of operand per k-block from shared memory instead of 48;
//   * D[channel][frame]: rank 0 ends up with the residual-conv rows of BOTH tiles, rank 1 with the skip-conv rows; the
//     epilogue (lane = channel) packs 8 frames per 16-byte st.shared into swizzled 64 x 64 boxes that leave as TMA tile
//     stores (the pattern of K_B's epilogue).
// Tensor memory: 256 columns of weights + one 256-column accumulator.  The accumulator is single-buffered, so the MMAs
// of a visit cannot start before the previous visit's accumulator has been read out; the kernel is built so that this
// costs little: every epilogue warp owns its 32 channels end to end (private 32 x 64 staging boxes and its own TMA tile
// stores -- no cross-warp barrier in the drain), the accumulator is released right after the last tcgen05.ld, and the
// producers keep filling the four-stage h2 ring meanwhile (the tensor core is twice as fast as the producers and
// catches up).
#pragma once
#include "pair_common.cuh"

namespace b200sep {
namespace pr2 {

struct KpCfg {
  static constexpr int S_STAGE = TILE_M * 64 * 2;   // 16 KiB: one h2 k-block of this CTA's m-tile (B operand)
  // h2 ring: 2 or 3 stages per producer group (chosen at run time from what the raw ring leaves).  The ring has to cover
  // the round trip of a stage -- remote arrive on the leader's barrier, the MMAs of the stages queued before it, the
  // multicast commit back to both CTAs: ~2 k cycles -- while a group produces a k-block in ~1.2 k.
  static constexpr int S_MAX = 8;
  static constexpr int ST_BOX = 32 * 128;           // 4 KiB: 32 channels x 64 frames store box (SWIZZLE_128B image)
  static constexpr int TAB_BYTES = C_HID * (2 * 8 + 16);  // per-visit tap table (two buffers, 4 halves per channel) + static t1 (edge tiles)
  static constexpr int EPI_WARPS = 4, PROD_WARPS = 16;
  static constexpr int THREADS = (EPI_WARPS + 3 + PROD_WARPS) * 32;  // 736
  // Warp roles.  The warp scheduler prefers the HIGHEST warp index among the eligible warps of a sub-partition, so the
  // roles on the tile-to-tile critical path (accumulator drain, MMA issue, statistics, TMA issue) sit above the 16
  // producer warps, which soak up every remaining issue slot.
  static constexpr int W_PROD0 = 0, W_TMA = 16, W_STAT = 17, W_MMA = 18, W_EPI0 = 19;
  static constexpr int TMEM_COLS = 512;
  static constexpr uint32_t D_COL0 = 0, A_COL0 = 256;
  static constexpr int OFF_TAB = 0;
  static constexpr int OFF_TI = OFF_TAB + TAB_BYTES;
  static constexpr int OFF_RED = OFF_TI + TI_CACHE * 32;
  static constexpr int OFF_CNT = OFF_RED + 2 * PROD_WARPS * 8;
  static constexpr int OFF_STAT = OFF_CNT + 16;      // [2] (mean, rstd) of the visit's segment, written by the statistics warp
  static constexpr int OFF_BAR = OFF_STAT + 16;
  static constexpr int OFF_S = (OFF_BAR + 512 + 1023) / 1024 * 1024;   // h2 ring, then the store staging, then the raw ring
  static constexpr int SMEM_BYTES = 232448;
  static constexpr int SMEM_USABLE = SMEM_BYTES - 1024;
};


__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t* u = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
        "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
      : "r"(taddr)
      : "memory");
}
// One pass of the accumulator drain of an epilogue warp: NCH chunks of 16 tensor-memory columns (= 16 frames of this
// lane's channel) starting at column c0 * 16 -> fp16 -> the warp's private staging boxes (box = four chunks = 64 frames,
// 128-byte rows, SWIZZLE_128B image); the tcgen05.ld of chunk c + 1 is in flight while chunk c is converted and stored.
// No fence, barrier or TMA issue inside: the accumulator stays locked for ~1.5 k cycles per visit.
// Returns after the LAST load has completed (the caller then releases the accumulator) with that chunk still to be
// stored: drain_tail finishes it.  (16-column chunks: 32-column ones do not fit the kernel's 80 registers per thread.)
__device__ __forceinline__ void drain_store(uint32_t st_box0, uint32_t st_row, uint32_t sw, int c, const float (&v)[16]) {
  uint32_t pk[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) pk[j] = h2_as_u32(__floats2half2_rn(v[2 * j], v[2 * j + 1]));
  const uint32_t box = st_box0 + (uint32_t)(c >> 2) * (uint32_t)KpCfg::ST_BOX + st_row;
#pragma unroll
  for (int k = 0; k < 2; ++k)
    st_shared_v4(box + ((((uint32_t)((c & 3) * 2 + k)) ^ sw) << 4),
                 make_uint4(pk[4 * k], pk[4 * k + 1], pk[4 * k + 2], pk[4 * k + 3]));
}
template <int NCH>
__device__ __forceinline__ void drain_pass(uint32_t t_lane, int c0, uint32_t st_box0, uint32_t st_row, uint32_t sw,
                                           float (&last)[16]) {
  float v[2][16];
  tmem_ld16(t_lane + (uint32_t)(c0 * 16), v[0]);
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    tmem_ld_wait();
    if (c + 1 < NCH) {
      tmem_ld16(t_lane + (uint32_t)((c0 + c + 1) * 16), v[(c + 1) & 1]);
      drain_store(st_box0, st_row, sw, c, v[c & 1]);
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) last[j] = v[c & 1][j];
    }
  }
}
template <int NCH>
__device__ __forceinline__ void drain_tail(uint32_t st_box0, uint32_t st_row, uint32_t sw, const float (&last)[16]) {
  drain_store(st_box0, st_row, sw, NCH - 1, last);
}

// ring position `by` k-blocks further on
__device__ __forceinline__ void ring_advance(int& stage, uint32_t& phase, int by, int n) {
  stage += by;
  while (stage >= n) { stage -= n; phase ^= 1u; }
}

// NG = number of producer groups (2 or 4): group g produces the k-blocks kb = g, g + NG, ... of every visit with its
// 16 / NG warps.  A warp's k-block is one long dependent chain (barrier waits, shared-memory round trips, fence, arrive:
// more than a thousand cycles of latency next to ~300 instructions), so what matters is how many k-blocks are in flight
// per SM: four groups need four raw stages (d <= 32), else two groups.
template <int D, bool A01, int NG>
__global__ void __launch_bounds__(KpCfg::THREADS, 1)
dw_rs_pair_kernel(const __grid_constant__ CUtensorMap tm_hc, const __grid_constant__ CUtensorMap tm_hh,
                  const __grid_constant__ CUtensorMap tm_res, const __grid_constant__ CUtensorMap tm_skip,
                  const __half* __restrict__ w256, int res_dead, int skip_row0, int skip_tile_rows,
                  const float2* __restrict__ part1, float2* __restrict__ part2, DwTab p, Tables tb) {
  using Cfg = KpCfg;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  auto stage_s = [&](int s) { return smem_base + (uint32_t)(Cfg::OFF_S + s * Cfg::S_STAGE); };
  // s_tap[b][c] = the channel's folded taps for the segment of visit it (b = it & 1), ready for HFMA2:
  //   { w0 g rstd, w1 g rstd, w2 g rstd, cb - (w0 + w1 + w2) g rstd mean } as four halves (the producers broadcast
  //   each into a half2 with one PRMT on the otherwise idle ALU pipe)
  // written by the statistics warp one visit ahead (the producers used to redo these 8 instructions per 8 outputs)
  uint2* s_tap = reinterpret_cast<uint2*>(smem_gen + Cfg::OFF_TAB);
  float4* s_t1 = reinterpret_cast<float4*>(s_tap + 2 * C_HID);
  TileInfo* s_ti = reinterpret_cast<TileInfo*>(smem_gen + Cfg::OFF_TI);
  float2* s_red = reinterpret_cast<float2*>(smem_gen + Cfg::OFF_RED);  // [2][16]
  int* s_cnt = reinterpret_cast<int*>(smem_gen + Cfg::OFF_CNT);          // [2]
  float2* s_stat = reinterpret_cast<float2*>(smem_gen + Cfg::OFF_STAT);   // [2]
  const uint32_t bar0 = smem_base + Cfg::OFF_BAR;
  auto raw_full = [&](int s) { return bar0 + 8u * (0 + s); };    // [<= 8] TMA -> producers (k-block n of the launch: stage n % raw_stages)
  auto raw_empty = [&](int s) { return bar0 + 8u * (8 + s); };   // [<= 8]
  auto s_full = [&](int s) { return bar0 + 8u * (16 + s); };     // [<= 8] producers of BOTH CTAs -> MMA (leader's copy is used)
  auto s_empty = [&](int s) { return bar0 + 8u * (24 + s); };    // [<= 8] MMA (multicast commit) -> producers
  const uint32_t tfull_bar = bar0 + 8u * 32;                     // MMA (multicast commit) -> epilogue
  const uint32_t tempty_bar = bar0 + 8u * 33;                    // epilogues of BOTH CTAs -> MMA (leader's copy)
  const uint32_t tmem_ptr_smem = bar0 + 8u * 34;
  const uint32_t done_bar = bar0 + 8u * 35;
  auto stat_full = [&](int s) { return bar0 + 8u * (36 + s); };  // [2] statistics warp -> producers
  auto stat_empty = [&](int s) { return bar0 + 8u * (38 + s); }; // [2] producers (one arrival per warp) -> statistics warp
  auto red_full = [&](int s) { return bar0 + 8u * (40 + s); };   // [2] producers' per-warp gLN2 partials -> statistics warp

  pdl_launch_dependents();
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int n_tiles = tb.n_tiles;
  // The pair (CTAs 2p, 2p + 1) visits the tile pairs p, p + n_pairs, ... (downwards, like the single-CTA K_A: it starts on
  // the h1 tiles K_B wrote last); rank r produces tile 2 * pair_tile + r.  Both CTAs make the same number of visits; with
  // an odd tile count rank 1's last visit re-processes the last tile and nothing of it is written.
  const int crank = (int)cluster_ctarank();
  const int n_pairs = (int)(gridDim.x >> 1), pair = (int)(blockIdx.x >> 1);
  const int n_pair_tiles = (n_tiles + 1) >> 1;
  const int n_visits = (pair < n_pair_tiles) ? (n_pair_tiles - pair + n_pairs - 1) / n_pairs : 0;
  auto pair_tile_at = [&](int it) { return n_pair_tiles - 1 - (pair + it * n_pairs); };
  auto tile_at = [&](int it, bool& valid) {
    const int wr = pair_tile_at(it) * 2 + crank;
    valid = wr < n_tiles;
    return valid ? wr : n_tiles - 1;
  };
  const int d = (D > 0) ? D : p.dilation;
  const int hw = (D > 0) ? 8 : d;  // halo width in frames (multiple of 8, <= 128)
  const uint32_t halo_bytes = (uint32_t)(64 * hw * 2);
  const uint32_t raw_stage_bytes = (16384u + 2u * halo_bytes + 1023u) & ~1023u;
  // Dynamic part of the shared memory: [h2 ring: s_stages x 16 KiB][store staging: st_pw boxes per epilogue warp][raw ring].
  // Priorities (measured): a six-stage h2 ring (see KpCfg), then four raw stages (two k-blocks in flight per group: the
  // L2 -> shared-memory latency of a stage is 1-1.8 k cycles), then four staging boxes per epilogue warp (a whole visit).
  const uint32_t dyn = (uint32_t)(Cfg::SMEM_USABLE - Cfg::OFF_S);
  const uint32_t S16 = (uint32_t)Cfg::S_STAGE, ST8 = 8u * (uint32_t)Cfg::ST_BOX;
  int s_stages, raw_stages, st_pw;
  if (6u * S16 + 4u * raw_stage_bytes + ST8 <= dyn) { s_stages = 6; raw_stages = 4; }
  else if (4u * S16 + 4u * raw_stage_bytes + ST8 <= dyn) { s_stages = 4; raw_stages = 4; }
  else if (6u * S16 + 2u * raw_stage_bytes + ST8 <= dyn) { s_stages = 6; raw_stages = 2; }
  else { s_stages = 4; raw_stages = 2; }
  st_pw = ((uint32_t)s_stages * S16 + (uint32_t)raw_stages * raw_stage_bytes + 2u * ST8 <= dyn) ? 4 : 2;
  const uint32_t off_st = (uint32_t)Cfg::OFF_S + (uint32_t)s_stages * S16;
  const uint32_t off_raw = off_st + (uint32_t)(4 * st_pw) * (uint32_t)Cfg::ST_BOX;
  auto raw_st = [&](int s) { return smem_base + off_raw + (uint32_t)s * raw_stage_bytes; };

  for (int i = threadIdx.x; i < C_HID; i += Cfg::THREADS) s_t1[i] = __ldg(p.t1 + i);
  if ((int)threadIdx.x < TI_CACHE && (int)threadIdx.x < n_visits) {
    bool v_;
    s_ti[threadIdx.x] = load_tile_info(tb.tile_info + tile_at(threadIdx.x, v_));
  }
  if (threadIdx.x < 2) s_cnt[threadIdx.x] = 0;
  if (warp == Cfg::W_TMA && lane == 0) { tma_prefetch_desc(&tm_hc); tma_prefetch_desc(&tm_hh); }
  if (warp == Cfg::W_STAT && lane == 0) { tma_prefetch_desc(&tm_res); tma_prefetch_desc(&tm_skip); }
  cluster_sync_all();  // both CTAs of the pair are resident before the pair-wide TMEM allocation
  if (warp == Cfg::W_MMA) {
    if (lane == 0) {
      for (int s = 0; s < 8; ++s) {
        mbar_init(raw_full(s), 1);
        mbar_init(raw_empty(s), Cfg::PROD_WARPS / NG);   // the warps of one producer group, one arrival each
      }
      for (int s = 0; s < Cfg::S_MAX; ++s) {
        mbar_init(s_full(s), 2 * Cfg::PROD_WARPS / NG);   // the warps of one group x 2 CTAs, one arrival per warp
        mbar_init(s_empty(s), 1);
      }
      mbar_init(tfull_bar, 1);
      mbar_init(tempty_bar, 2 * Cfg::EPI_WARPS);          // one arrival per epilogue warp of both CTAs
      mbar_init(done_bar, 1);
      for (int s = 0; s < 2; ++s) {
        mbar_init(stat_full(s), 1);
        mbar_init(stat_empty(s), Cfg::PROD_WARPS);
        mbar_init(red_full(s), Cfg::PROD_WARPS);
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc_pair(tmem_ptr_smem, Cfg::TMEM_COLS);       // issued by one warp of EACH CTA of the pair
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_smem));
  if (warp >= Cfg::W_EPI0 && warp < Cfg::W_EPI0 + 4) {
    // resident weights: thread = one of this CTA's 128 output channels, its 512 k-values -> 256 columns of its lane
    // (a warp reaches the tensor-memory lanes [32 (warp % 4), + 32))
    const int quad = warp & 3;
    const __half* wrow = w256 + (size_t)(crank * 128 + quad * 32 + lane) * C_HID;
#pragma unroll 1
    for (int c8 = 0; c8 < 8; ++c8)
      tmem_store_w64(tmem_base + ((uint32_t)(quad * 32) << 16) + Cfg::A_COL0 + (uint32_t)(c8 * 32), wrow + c8 * 64);
    tmem_st_wait();
  }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();  // both CTAs' barriers are initialised and both weight halves are in place
  tcgen05_fence_after();
  pdl_wait();

  if (warp >= Cfg::W_EPI0 && warp < Cfg::W_EPI0 + 4) {
    // ================================ epilogue: TMEM -> fp16 raw G (TMA tile stores) ===============
    // Warp q owns this CTA's output channels [32 q, 32 q + 32) (its tensor-memory lanes): per visit four boxes of
    // 32 channels x 64 frames (tile of the pair, frame half), each staged in a private swizzled buffer and stored by the
    // warp's own lane 0 -- no barrier with the other epilogue warps.
    const int quad = warp & 3;
    const bool tl_epi = (warp == Cfg::W_EPI0 && lane == 0);
    const CUtensorMap* tm_out = (crank == 0) ? &tm_res : &tm_skip;
    const bool store_on = !(crank == 0 && res_dead != 0);
    const uint32_t my_tempty = mapa_rank(tempty_bar, 0);
    const uint32_t st_warp = smem_base + off_st + (uint32_t)(quad * st_pw) * (uint32_t)Cfg::ST_BOX;
    const uint32_t st_row = (uint32_t)(lane * 128);
    const uint32_t sw = (uint32_t)(lane & 7);
    const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16) + Cfg::D_COL0;
    // issue the TMA tile stores of the staged boxes [b0, b0 + nb) of visit `pt` (one bulk group)
    auto store_boxes = [&](int pt, int b0, int nb) {
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        for (int b = 0; b < nb; ++b) {
          const int bx = b0 + b, t = bx >> 1, h = bx & 1;
          const int wt = pt * 2 + t;
          if (store_on && wt < n_tiles) {
            const int r0 = (crank == 0) ? wt * C_BN : wt * skip_tile_rows + skip_row0;
            tma_store_2d(tm_out, st_warp + (uint32_t)b * (uint32_t)Cfg::ST_BOX, h * 64, r0 + quad * 32);
          }
        }
        bulk_commit();
      }
    };
    auto release_acc = [&](int it) {
      // the accumulator is in registers / staging: the next visit's MMAs may overwrite it
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(my_tempty);
      if (tl_epi) tl_stamp(it, 82);
    };
    for (int it = 0; it < n_visits; ++it) {
      const int pt = pair_tile_at(it);
      mbar_wait_b(tb.flags, tfull_bar, (uint32_t)it & 1u);
      tcgen05_fence_after();
      if (tl_epi) tl_stamp(it, 80);
      // the previous visit's stores have finished READING the staging (they were issued a whole visit ago)
      if (lane == 0) bulk_wait_read0();
      __syncwarp();
      float last[16];
      if (st_pw == 4) {
        drain_pass<16>(t_lane, 0, st_warp, st_row, sw, last);
        release_acc(it);
        drain_tail<16>(st_warp, st_row, sw, last);
        store_boxes(pt, 0, 4);
      } else {
        drain_pass<8>(t_lane, 0, st_warp, st_row, sw, last);
        drain_tail<8>(st_warp, st_row, sw, last);
        store_boxes(pt, 0, 2);
        if (lane == 0) bulk_wait_read0();
        __syncwarp();
        drain_pass<8>(t_lane, 8, st_warp, st_row, sw, last);
        release_acc(it);
        drain_tail<8>(st_warp, st_row, sw, last);
        store_boxes(pt, 2, 2);
      }
      if (tl_epi) tl_stamp(it, 81);
    }
    if (lane == 0) bulk_wait0();  // every tile store has completed before the CTA exits
  } else if (warp == Cfg::W_TMA) {
    // ================================ TMA producer: raw h1 k-blocks + halos =======================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t bytes = (uint32_t)(64 * (TILE_M + 2 * hw) * 2);
      for (int it = 0; it < n_visits; ++it) {
        bool valid, nvalid = false;
        const int w = tile_at(it, valid);
        const int wn = (it + 1 < n_visits) ? tile_at(it + 1, nvalid) : 0;
        for (int kb = 0; kb < 8; ++kb) {
          // pull the same k-block of the NEXT tile this CTA visits into L2 (a whole tile ahead of its use)
          if (nvalid) l2_prefetch_tma_2d(&tm_hc, 0, wn * C_HID + kb * 64);
          mbar_wait_b(tb.flags, raw_empty(stage), phase ^ 1u);
          tl_stamp(it, 88 + kb);
          const uint32_t dst = raw_st(stage);
          mbar_expect_tx(raw_full(stage), bytes);
          tma_load_2d(dst, &tm_hc, raw_full(stage), 0, w * C_HID + kb * 64);
          // frames [t0 - hw, t0) live at the end of the previous m-tile, [t0 + 128, t0 + 128 + hw) at the
          // start of the next; rows outside the tensor are zero-filled by the TMA unit
          tma_load_2d(dst + 16384u, &tm_hh, raw_full(stage), TILE_M - hw, (w - 1) * C_HID + kb * 64);
          tma_load_2d(dst + 16384u + halo_bytes, &tm_hh, raw_full(stage), 0, (w + 1) * C_HID + kb * 64);
          if (++stage == raw_stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else if (warp == Cfg::W_STAT) {
    // ================================ statistics warp ==============================================
    // gLN1 (mean, rstd) of every visit's segment (fp64 reduction of the per-tile partials K_B left behind), one visit
    // ahead of the producers: 1.5-2.5 k cycles that every producer warp used to spend in the middle of each tile.
    // It also folds the statistics into the per-channel tap table of the visit, and finishes the gLN2 partial of the
    // previous visit (the 16 producer warps' (sum, sumsq) pairs, added in warp order: a fixed order, and no producer
    // warp stays behind to do it).
    auto reduce_visit = [&](int it) {
      bool valid;
      const int w = tile_at(it, valid);
      mbar_wait_b(tb.flags, red_full(it & 1), ((uint32_t)it >> 1) & 1u);
      const float2 v = (lane < Cfg::PROD_WARPS) ? s_red[(it & 1) * Cfg::PROD_WARPS + lane] : make_float2(0.f, 0.f);
      float ts = 0.f, tq = 0.f;
#pragma unroll
      for (int k = 0; k < Cfg::PROD_WARPS; ++k) {   // sequential in warp order (same sum as before)
        ts += __shfl_sync(0xffffffffu, v.x, k);
        tq += __shfl_sync(0xffffffffu, v.y, k);
      }
      if (lane == 0 && valid) part2[w] = make_float2(ts, tq);
    };
    for (int it = 0; it < n_visits; ++it) {
      bool v_;
      const TileInfo ti = ti_get(s_ti, tb, it, tile_at(it, v_));
      const float2 pre = part_prefetch(part1, ti);
      float mean, rstd;
      seg_stats_pre(part1, ti, pre, p.eps1, mean, rstd, tb.flags);
      const float rm = rstd * mean;
      mbar_wait_b(tb.flags, stat_empty(it & 1), (((uint32_t)it >> 1) & 1u) ^ 1u);
      uint2* tap = s_tap + (it & 1) * C_HID;
#pragma unroll 1
      for (int c0 = 0; c0 < C_HID; c0 += 128) {
        float4 t0[4];
        float cb[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          t0[j] = __ldg(p.t0 + c0 + j * 32 + lane);
          cb[j] = __ldg(p.cb + c0 + j * 32 + lane);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
          tap[c0 + j * 32 + lane] = make_uint2(h2_as_u32(__floats2half2_rn(t0[j].x * rstd, t0[j].y * rstd)),
                                               h2_as_u32(__floats2half2_rn(t0[j].z * rstd, fmaf(-t0[j].w, rm, cb[j]))));
      }
      if (lane == 0) s_stat[it & 1] = make_float2(mean, rstd);
      __syncwarp();
      if (lane == 0) mbar_arrive(stat_full(it & 1));  // (release at CTA scope orders the warp's stores before the arrival)
      if (it > 0) reduce_visit(it - 1);
    }
    if (n_visits > 0) reduce_visit(n_visits - 1);
  } else if (warp == Cfg::W_MMA) {
    // ================================ MMA issuer (leader CTA only) ================================
    if (lane == 0 && crank == 0) {
      constexpr uint32_t idesc = make_idesc_ts(256, 256);
      int sst = 0;
      uint32_t sph = 0;
      for (int it = 0; it < n_visits; ++it) {
        mbar_wait_b(tb.flags, tempty_bar, ((uint32_t)it & 1u) ^ 1u);  // both CTAs' epilogues drained the accumulator
        tcgen05_fence_after();
        tl_stamp(it, 64);
        const uint32_t tmem_d = tmem_base + Cfg::D_COL0;
        for (int kb = 0; kb < 8; ++kb) {
          const int g = sst;
          const uint32_t gphase = sph;
          ring_advance(sst, sph, 1, s_stages);
          mbar_wait_b(tb.flags, s_full(g), gphase);   // both CTAs' producer groups have written this stage
          tcgen05_fence_after();
          tl_stamp(it, 65 + kb);
          const uint32_t sb = stage_s(g);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint64_t db = make_smem_desc(sb + j * 2048u, 8192u, 1024u);  // 16 k-rows x 128 frames, MN-major
            umma_f16_ts_pair(tmem_d, tmem_base + Cfg::A_COL0 + (uint32_t)((kb * 4 + j) * 8), db, idesc,
                             (uint32_t)((kb | j) != 0));
          }
          umma_commit_pair(s_empty(g), (uint16_t)0x3);
          if (kb == 7) umma_commit_pair(tfull_bar, (uint16_t)0x3);
        }
      }
      umma_commit_pair(done_bar, (uint16_t)0x1);  // every asynchronous arrive has landed before the pair tears down
      mbar_wait_b(tb.flags, done_bar, 0);
    }
    __syncwarp();
  } else {
    // ================================ depthwise producers =========================================
    // NG groups of WPG warps; group g produces the k-blocks kb = g, g + NG, ..., so the groups run out of phase.
    // A warp-step = 2 channels x 128 frames; a group covers a k-block in NSTEP = 32 / WPG steps per warp.
    constexpr int WPG = Cfg::PROD_WARPS / NG, NSTEP = 32 / WPG;
    const int pw = warp - Cfg::W_PROD0;   // 0..15
    const int grp = pw / WPG, gw = pw % WPG;
    const int sub = lane >> 4, lane16 = lane & 15;
    const int row0 = gw * 2 + sub;        // k-row of step 0 inside a k-block; step j = row0 + 2 WPG j
    const int sv = (D > 0) ? 1 : (d >> 3);  // tap distance in 8-frame vectors
    const int hv = hw >> 3;                 // halo width in vectors
    const uint32_t c_pitch = 256u;
    uint32_t l_off, l_pitch, r_off, r_pitch;
    if (lane16 - sv >= 0) { l_off = (uint32_t)((lane16 - sv) * 16); l_pitch = 256u; }
    else { l_off = 16384u + (uint32_t)((lane16 - sv + hv) * 16); l_pitch = (uint32_t)(hw * 2); }
    if (lane16 + sv < 16) { r_off = (uint32_t)((lane16 + sv) * 16); r_pitch = 256u; }
    else { r_off = 16384u + halo_bytes + (uint32_t)((lane16 + sv - 16) * 16); r_pitch = (uint32_t)(hw * 2); }
    const uint32_t c_base = (uint32_t)(lane16 * 16) + (uint32_t)row0 * c_pitch;
    const uint32_t l_base = l_off + (uint32_t)row0 * l_pitch, r_base = r_off + (uint32_t)row0 * r_pitch;
    constexpr uint32_t RSTEP = 2u * WPG;   // rows per step (a multiple of 8: (row & 7) is the same for every step)
    const uint32_t c_step = RSTEP * c_pitch, l_step = RSTEP * l_pitch, r_step = RSTEP * r_pitch;
    const uint32_t a_off0 = a_chunk_off(row0, lane16 * 8);
    constexpr uint32_t A_STEP = RSTEP * 128u;   // bytes between a thread's stores of consecutive steps
    int sst = 0, rst = 0;             // h2 / raw stage of this group's next k-block
    uint32_t sphase = 0, rphase = 0;
    ring_advance(sst, sphase, grp, s_stages);
    ring_advance(rst, rphase, grp, raw_stages);
    const __half2 alpha_h2 = __float2half2_rn(p.alpha2);
    const bool tl_on = (gw == 0 && lane == 0 && grp < 2);
    const int tl_base = grp ? 100 : 40;
    for (int it = 0; it < n_visits; ++it) {
      bool valid;
      const int w = tile_at(it, valid);
      const TileInfo ti = ti_get(s_ti, tb, it, w);
      const int F = ti.F, tile_t0 = ti.t0;
      mbar_wait_b(tb.flags, stat_full(it & 1), ((uint32_t)it >> 1) & 1u);
      const float mean = s_stat[it & 1].x;   // (the tap table of the visit carries rstd)
      const uint2* tap = s_tap + (it & 1) * C_HID;
      if (tl_on) tl_stamp(it, 32 + grp);
      const int t_base = tile_t0 + lane16 * 8;
      const bool interior = (tile_t0 - d >= 0) && (tile_t0 + TILE_M - 1 + d < F);
      float2 s2 = make_float2(0.f, 0.f), q2 = make_float2(0.f, 0.f);
      DwEdgeH mk;
      if (!interior) dw_edge_masks(mk, t_base, d, F);  // taps / frames switched off by segment-relative frame index
#pragma unroll 1
      for (int kb = grp; kb < 8; kb += NG) {
        const int s_idx = sst;
        const uint32_t sa = stage_s(s_idx) + a_off0;
        const uint32_t my_s_full = mapa_rank(s_full(s_idx), 0);  // the LEADER's barrier (cluster address)
        const uint32_t my_s_empty = s_empty(s_idx);
        const int rs_ = rst;                                      // raw stage of this k-block
        mbar_wait_b(tb.flags, raw_full(rs_), rphase);
        if (tl_on) tl_stamp(it, tl_base + 3 * (kb / NG));
        const uint32_t raw = raw_st(rs_);
        const int ch0 = kb * 64 + row0;
        if (interior) {
#pragma unroll
          for (int jj = 0; jj < NSTEP; jj += 2) {
            // both steps' vectors and table entries first (independent shared-memory loads in flight) ...
            const uint4 vc0 = ld_shared_v4(raw + c_base + jj * c_step);
            const uint4 vc1 = ld_shared_v4(raw + c_base + (jj + 1) * c_step);
            const uint4 va0 = ld_shared_v4(raw + l_base + jj * l_step);
            const uint4 vb0 = ld_shared_v4(raw + r_base + jj * r_step);
            const uint4 va1 = ld_shared_v4(raw + l_base + (jj + 1) * l_step);
            const uint4 vb1 = ld_shared_v4(raw + r_base + (jj + 1) * r_step);
            const uint2 ta = tap[ch0 + jj * (int)RSTEP], tb_ = tap[ch0 + (jj + 1) * (int)RSTEP];
            // ... then the arithmetic; the stage is only needed for the stores
            __half2 sh = __float2half2_rn(0.f), qh = __float2half2_rn(0.f);
            const uint4 o0 = dw_step_h2p<D, A01>(vc0, va0, vb0, ta, alpha_h2, sh, qh);
            const uint4 o1 = dw_step_h2p<D, A01>(vc1, va1, vb1, tb_, alpha_h2, sh, qh);
            dw_stats_flush(sh, qh, s2, q2);
            if (jj == 0) {
              mbar_wait_b(tb.flags, my_s_empty, sphase ^ 1u);
              if (tl_on) tl_stamp(it, tl_base + 3 * (kb / NG) + 1);
            }
            st_shared_v4(sa + jj * A_STEP, o0);
            st_shared_v4(sa + (jj + 1) * A_STEP, o1);
          }
        } else {
#pragma unroll
          for (int jj = 0; jj < NSTEP; jj += 2) {
            const uint4 vc0 = ld_shared_v4(raw + c_base + jj * c_step);
            const uint4 vc1 = ld_shared_v4(raw + c_base + (jj + 1) * c_step);
            const uint4 va0 = ld_shared_v4(raw + l_base + jj * l_step);
            const uint4 vb0 = ld_shared_v4(raw + r_base + jj * r_step);
            const uint4 va1 = ld_shared_v4(raw + l_base + (jj + 1) * l_step);
            const uint4 vb1 = ld_shared_v4(raw + r_base + (jj + 1) * r_step);
            const uint2 ta = tap[ch0 + jj * (int)RSTEP], tb_ = tap[ch0 + (jj + 1) * (int)RSTEP];
            const float4 ua = s_t1[ch0 + jj * (int)RSTEP], ub = s_t1[ch0 + (jj + 1) * (int)RSTEP];
            __half2 sh = __float2half2_rn(0.f), qh = __float2half2_rn(0.f);
            const uint4 o0 = dw_step_h2p_edge<D, A01>(vc0, va0, vb0, ta, ua, mean, alpha_h2, mk, sh, qh);
            const uint4 o1 = dw_step_h2p_edge<D, A01>(vc1, va1, vb1, tb_, ub, mean, alpha_h2, mk, sh, qh);
            dw_stats_flush(sh, qh, s2, q2);
            if (jj == 0) mbar_wait_b(tb.flags, my_s_empty, sphase ^ 1u);
            st_shared_v4(sa + jj * A_STEP, o0);
            st_shared_v4(sa + (jj + 1) * A_STEP, o1);
          }
        }
        fence_proxy_async();
        __syncwarp();  // every lane has fenced its stage writes; lane 0 arrives (release.cluster is cumulative)
        if (lane == 0) {
          mbar_arrive_cluster(my_s_full);
          mbar_arrive(raw_empty(rs_));  // (the __syncwarp above also orders every lane's reads of the raw stage)
        }
        if (tl_on) tl_stamp(it, tl_base + 3 * (kb / NG) + 2);
        ring_advance(sst, sphase, NG, s_stages);
        ring_advance(rst, rphase, NG, raw_stages);
      }
      // gLN2 partial of this m-tile: every warp publishes its (sum, sumsq); the statistics warp adds the 16 pairs in
      // warp order and writes the tile's single partial.  The visit's tap table / statistics are free again.
      const float s_acc = warp_sum(s2.x + s2.y);
      const float q_acc = warp_sum(q2.x + q2.y);
      if (lane == 0) {
        s_red[(it & 1) * Cfg::PROD_WARPS + pw] = make_float2(s_acc, q_acc);
        mbar_arrive(red_full(it & 1));
        mbar_arrive(stat_empty(it & 1));
      }
      if (tl_on) tl_stamp(it, 34 + grp);
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer may still arrive on / read from this CTA's shared memory until it is done too
  if (warp == Cfg::W_MMA) {
    tcgen05_fence_after();
    tmem_dealloc_pair(tmem_base, Cfg::TMEM_COLS);
  }
}

}  // namespace pr2
}  // namespace b200sep

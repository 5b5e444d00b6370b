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
  static constexpr int S_PG = 2;                    // stages per producer group
  static constexpr int N_STAGES = 2 * S_PG;
  static constexpr int ST_BOX = 32 * 128;           // 4 KiB: 32 channels x 64 frames store box (SWIZZLE_128B image)
  static constexpr int TAB_BYTES = C_HID * (16 + 4);  // t0 + cb (t1 is only read on edge tiles: from global / L1)
  static constexpr int EPI_WARPS = 4, PROD_WARPS = 16;
  static constexpr int THREADS = (EPI_WARPS + 3 + PROD_WARPS) * 32;  // 736
  static constexpr int TMEM_COLS = 512;
  static constexpr uint32_t D_COL0 = 0, A_COL0 = 256;
  static constexpr int OFF_TAB = 0;
  static constexpr int OFF_TI = OFF_TAB + TAB_BYTES;
  static constexpr int OFF_RED = OFF_TI + TI_CACHE * 32;
  static constexpr int OFF_CNT = OFF_RED + 2 * PROD_WARPS * 8;
  static constexpr int OFF_BAR = OFF_CNT + 16;
  static constexpr int OFF_S = (OFF_BAR + 512 + 1023) / 1024 * 1024;   // h2 ring, then the store staging, then the raw ring
  static constexpr int OFF_ST = OFF_S + N_STAGES * S_STAGE;
  static constexpr int SMEM_BYTES = 232448;
  static constexpr int SMEM_USABLE = SMEM_BYTES - 1024;
};

__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read3() { asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory"); }

template <int D, bool A01>
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
  float4* s_t0 = reinterpret_cast<float4*>(smem_gen + Cfg::OFF_TAB);
  float* s_cb = reinterpret_cast<float*>(s_t0 + C_HID);
  TileInfo* s_ti = reinterpret_cast<TileInfo*>(smem_gen + Cfg::OFF_TI);
  float2* s_red = reinterpret_cast<float2*>(smem_gen + Cfg::OFF_RED);  // [2][16]
  int* s_cnt = reinterpret_cast<int*>(smem_gen + Cfg::OFF_CNT);          // [2]
  const uint32_t bar0 = smem_base + Cfg::OFF_BAR;
  auto raw_full = [&](int s) { return bar0 + 8u * (0 + s); };    // [4] TMA -> producers (stage s belongs to group s & 1)
  auto raw_empty = [&](int s) { return bar0 + 8u * (4 + s); };   // [4]
  auto s_full = [&](int s) { return bar0 + 8u * (8 + s); };      // [4] producers of BOTH CTAs -> MMA (leader's copy is used)
  auto s_empty = [&](int s) { return bar0 + 8u * (12 + s); };    // [4] MMA (multicast commit) -> producers
  const uint32_t tfull_bar = bar0 + 8u * 16;                     // MMA (multicast commit) -> epilogue
  const uint32_t tempty_bar = bar0 + 8u * 17;                    // epilogues of BOTH CTAs -> MMA (leader's copy)
  const uint32_t tmem_ptr_smem = bar0 + 8u * 18;
  const uint32_t done_bar = bar0 + 8u * 19;

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
  // what is left after the h2 ring goes to the raw ring (4 stages = two k-blocks in flight per producer group, when they
  // fit next to the small staging: the L2 -> shared-memory latency of a stage is about two k-block times) and to the
  // store staging (4 boxes per epilogue warp = a whole visit, so the drain never waits for a TMA store; else 2)
  const uint32_t dyn = (uint32_t)(Cfg::SMEM_USABLE - Cfg::OFF_ST);
  const int raw_stages = (4u * raw_stage_bytes + 8u * (uint32_t)Cfg::ST_BOX <= dyn) ? 4 : 2;
  const int st_pw = ((uint32_t)raw_stages * raw_stage_bytes + 16u * (uint32_t)Cfg::ST_BOX <= dyn) ? 4 : 2;
  const uint32_t off_raw = (uint32_t)Cfg::OFF_ST + (uint32_t)(4 * st_pw) * (uint32_t)Cfg::ST_BOX;
  auto raw_st = [&](int s) { return smem_base + off_raw + (uint32_t)s * raw_stage_bytes; };

  for (int i = threadIdx.x; i < C_HID; i += Cfg::THREADS) {
    s_t0[i] = __ldg(p.t0 + i);
    s_cb[i] = __ldg(p.cb + i);
  }
  if ((int)threadIdx.x < TI_CACHE && (int)threadIdx.x < n_visits) {
    bool v_;
    s_ti[threadIdx.x] = load_tile_info(tb.tile_info + tile_at(threadIdx.x, v_));
  }
  if (threadIdx.x < 2) s_cnt[threadIdx.x] = 0;
  if (warp == 4 && lane == 0) { tma_prefetch_desc(&tm_hc); tma_prefetch_desc(&tm_hh); }
  if (warp == 5 && lane == 0) { tma_prefetch_desc(&tm_res); tma_prefetch_desc(&tm_skip); }
  cluster_sync_all();  // both CTAs of the pair are resident before the pair-wide TMEM allocation
  if (warp == 6) {
    if (lane == 0) {
      for (int s = 0; s < 4; ++s) {
        mbar_init(raw_full(s), 1);
        mbar_init(raw_empty(s), Cfg::PROD_WARPS / 2);    // the 8 warps of the owning producer group, one arrival each
        mbar_init(s_full(s), Cfg::PROD_WARPS);            // 8 warps of the owning group x 2 CTAs, one arrival per warp
        mbar_init(s_empty(s), 1);
      }
      mbar_init(tfull_bar, 1);
      mbar_init(tempty_bar, 2 * Cfg::EPI_WARPS);          // one arrival per epilogue warp of both CTAs
      mbar_init(done_bar, 1);
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
  if (warp < 4) {
    // resident weights: thread = one of this CTA's 128 output channels, its 512 k-values -> 256 columns of its lane
    const __half* wrow = w256 + (size_t)(crank * 128 + warp * 32 + lane) * C_HID;
#pragma unroll 1
    for (int c8 = 0; c8 < 8; ++c8)
      tmem_store_w64(tmem_base + ((uint32_t)(warp * 32) << 16) + Cfg::A_COL0 + (uint32_t)(c8 * 32), wrow + c8 * 64);
    tmem_st_wait();
  }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();  // both CTAs' barriers are initialised and both weight halves are in place
  tcgen05_fence_after();
  pdl_wait();

  if (warp < 4) {
    // ================================ epilogue: TMEM -> fp16 raw G (TMA tile stores) ===============
    // Warp q owns this CTA's output channels [32 q, 32 q + 32) (its tensor-memory lanes): per visit four boxes of
    // 32 channels x 64 frames (tile of the pair, frame half), each staged in a private swizzled buffer and stored by the
    // warp's own lane 0 -- no barrier with the other epilogue warps.
    const int quad = warp;
    const CUtensorMap* tm_out = (crank == 0) ? &tm_res : &tm_skip;
    const bool store_on = !(crank == 0 && res_dead != 0);
    const uint32_t my_tempty = mapa_rank(tempty_bar, 0);
    const uint32_t st_warp = smem_base + (uint32_t)Cfg::OFF_ST + (uint32_t)(quad * st_pw) * (uint32_t)Cfg::ST_BOX;
    const uint32_t st_row = (uint32_t)(lane * 128);
    const uint32_t sw = (uint32_t)(lane & 7);
    const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16) + Cfg::D_COL0;
    int buf = 0;
    for (int it = 0; it < n_visits; ++it) {
      const int pt = pair_tile_at(it);
      mbar_wait_b(tb.flags, tfull_bar, (uint32_t)it & 1u);
      tcgen05_fence_after();
      if (threadIdx.x == 0) tl_stamp(it, 80);
#pragma unroll 1
      for (int bx = 0; bx < 4; ++bx) {          // (tile of the pair, frame half)
        const int t = bx >> 1, h = bx & 1;
        const int wt = pt * 2 + t;
        const uint32_t st_box = st_warp + (uint32_t)buf * (uint32_t)Cfg::ST_BOX;
        // the store issued from this staging buffer st_pw boxes ago must have finished READING it
        if (lane == 0) { if (st_pw == 4) bulk_wait_read3(); else bulk_wait_read1(); }
        __syncwarp();
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          float v[32];
          tmem_ld32(t_lane + (uint32_t)(t * 128 + h * 64 + hh * 32), v);
          tmem_ld_wait();
          if (bx == 3 && hh == 1) {
            // the accumulator is in registers: the next visit's MMAs may overwrite it
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(my_tempty);
            if (threadIdx.x == 0) tl_stamp(it, 82);
          }
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) pk[j] = h2_as_u32(__floats2half2_rn(v[2 * j], v[2 * j + 1]));
#pragma unroll
          for (int k = 0; k < 4; ++k)
            st_shared_v4(st_box + st_row + ((((uint32_t)(hh * 4 + k)) ^ sw) << 4),
                         make_uint4(pk[4 * k], pk[4 * k + 1], pk[4 * k + 2], pk[4 * k + 3]));
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          if (store_on && wt < n_tiles) {
            const int r0 = (crank == 0) ? wt * C_BN : wt * skip_tile_rows + skip_row0;
            tma_store_2d(tm_out, st_box, h * 64, r0 + quad * 32);
          }
          bulk_commit();  // (an empty group when nothing was stored: keeps the wait_group arithmetic uniform)
        }
        if (++buf == st_pw) buf = 0;
      }
      if (threadIdx.x == 0) tl_stamp(it, 81);
    }
    if (lane == 0) bulk_wait0();  // every tile store has completed before the CTA exits
  } else if (warp == 4) {
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
          mbar_expect_tx(raw_full(stage), bytes);
          const uint32_t dst = raw_st(stage);
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
  } else if (warp == 5) {
    // (spare warp)
  } else if (warp == 6) {
    // ================================ MMA issuer (leader CTA only) ================================
    if (lane == 0 && crank == 0) {
      constexpr uint32_t idesc = make_idesc_ts(256, 256);
      int nseq[2] = {0, 0};
      for (int it = 0; it < n_visits; ++it) {
        mbar_wait_b(tb.flags, tempty_bar, ((uint32_t)it & 1u) ^ 1u);  // both CTAs' epilogues drained the accumulator
        tcgen05_fence_after();
        tl_stamp(it, 64);
        const uint32_t tmem_d = tmem_base + Cfg::D_COL0;
        for (int kb = 0; kb < 8; ++kb) {
          const int grp = kb & 1;
          const int g = grp * Cfg::S_PG + (nseq[grp] & (Cfg::S_PG - 1));
          const uint32_t gphase = (uint32_t)(nseq[grp] / Cfg::S_PG) & 1u;
          ++nseq[grp];
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
    // Two groups of 8 warps; group g produces the k-blocks kb = g, g + 2, ... into its own two stages, so the groups run
    // out of phase.  A warp-step = 2 channels x 128 frames; a group covers a k-block in 4 steps per warp.
    const int pw = warp - 7;              // 0..15
    const int grp = pw >> 3, gw = pw & 7;
    const int sub = lane >> 4, lane16 = lane & 15;
    const int row0 = gw * 2 + sub;        // k-row of step 0 inside a k-block; step j = row0 + 16 j
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
    const uint32_t c_step = 16u * c_pitch, l_step = 16u * l_pitch, r_step = 16u * r_pitch;
    const uint32_t a_off0 = a_chunk_off(row0, lane16 * 8);  // (row & 7) is the same for every step; step j = + 2048 j
    int nseq = 0;  // k-blocks produced by this group so far -> stage / phase
    const __half2 alpha_h2 = __float2half2_rn(p.alpha2);
    const bool tl_on = (gw == 0 && lane == 0);
    const int tl_base = grp ? 100 : 40;

    const int rpg = raw_stages >> 1;  // raw stages per group
    float nmean = 0.f, nrstd = 0.f;  // statistics of the NEXT tile's segment (computed mid-tile)
    if (n_visits > 0) {
      bool v_;
      const TileInfo t0i = ti_get(s_ti, tb, 0, tile_at(0, v_));
      seg_stats_pre(part1, t0i, part_prefetch(part1, t0i), p.eps1, nmean, nrstd, tb.flags);
    }
    for (int it = 0; it < n_visits; ++it) {
      bool valid;
      const int w = tile_at(it, valid);
      const TileInfo ti = ti_get(s_ti, tb, it, w);
      const int F = ti.F, tile_t0 = ti.t0;
      const float mean = nmean, rstd = nrstd;
      const bool has_next = it + 1 < n_visits;
      TileInfo tn = ti;
      float2 pre = make_float2(0.f, 0.f);
      if (has_next) {
        bool v_;
        tn = ti_get(s_ti, tb, it + 1, tile_at(it + 1, v_));
        pre = part_prefetch(part1, tn);
      }
      const float rm = rstd * mean;
      if (tl_on) tl_stamp(it, 32 + grp);
      const int t_base = tile_t0 + lane16 * 8;
      const bool interior = (tile_t0 - d >= 0) && (tile_t0 + TILE_M - 1 + d < F);
      float2 s2 = make_float2(0.f, 0.f), q2 = make_float2(0.f, 0.f);
      DwEdgeH mk;
      if (!interior) dw_edge_masks(mk, t_base, d, F);  // taps / frames switched off by segment-relative frame index
#pragma unroll 1
      for (int kb = grp; kb < 8; kb += 2) {
        const int s_idx = grp * Cfg::S_PG + (nseq & (Cfg::S_PG - 1));
        const uint32_t sphase = (uint32_t)(nseq / Cfg::S_PG) & 1u;
        const uint32_t sa = stage_s(s_idx) + a_off0;
        const uint32_t my_s_full = mapa_rank(s_full(s_idx), 0);  // the LEADER's barrier (cluster address)
        const uint32_t my_s_empty = s_empty(s_idx);
        const int rs_ = grp + 2 * (nseq & (rpg - 1));             // raw stage of this k-block
        const uint32_t rphase = (uint32_t)(nseq / rpg) & 1u;
        mbar_wait_b(tb.flags, raw_full(rs_), rphase);
        if (tl_on) tl_stamp(it, tl_base + 3 * (kb >> 1));
        const uint32_t raw = raw_st(rs_);
        const int ch0 = kb * 64 + row0;
        if (interior) {
#pragma unroll
          for (int jj = 0; jj < 4; jj += 2) {
            // both steps' vectors and table entries first (independent shared-memory loads in flight) ...
            const uint4 vc0 = ld_shared_v4(raw + c_base + jj * c_step);
            const uint4 vc1 = ld_shared_v4(raw + c_base + (jj + 1) * c_step);
            const uint4 va0 = ld_shared_v4(raw + l_base + jj * l_step);
            const uint4 vb0 = ld_shared_v4(raw + r_base + jj * r_step);
            const uint4 va1 = ld_shared_v4(raw + l_base + (jj + 1) * l_step);
            const uint4 vb1 = ld_shared_v4(raw + r_base + (jj + 1) * r_step);
            const float4 ta = s_t0[ch0 + jj * 16], tb_ = s_t0[ch0 + jj * 16 + 16];
            const float cba = s_cb[ch0 + jj * 16], cbb = s_cb[ch0 + jj * 16 + 16];
            // ... then the arithmetic; the stage is only needed for the stores
            __half2 sh = __float2half2_rn(0.f), qh = __float2half2_rn(0.f);
            const uint4 o0 = dw_step_h2<D, A01>(vc0, va0, vb0, ta, cba, rstd, rm, alpha_h2, sh, qh);
            const uint4 o1 = dw_step_h2<D, A01>(vc1, va1, vb1, tb_, cbb, rstd, rm, alpha_h2, sh, qh);
            dw_stats_flush(sh, qh, s2, q2);
            if (jj == 0) {
              mbar_wait_b(tb.flags, my_s_empty, sphase ^ 1u);
              if (tl_on) tl_stamp(it, tl_base + 3 * (kb >> 1) + 1);
            }
            st_shared_v4(sa + jj * 2048u, o0);
            st_shared_v4(sa + (jj + 1) * 2048u, o1);
          }
        } else {
#pragma unroll
          for (int jj = 0; jj < 4; jj += 2) {
            const uint4 vc0 = ld_shared_v4(raw + c_base + jj * c_step);
            const uint4 vc1 = ld_shared_v4(raw + c_base + (jj + 1) * c_step);
            const uint4 va0 = ld_shared_v4(raw + l_base + jj * l_step);
            const uint4 vb0 = ld_shared_v4(raw + r_base + jj * r_step);
            const uint4 va1 = ld_shared_v4(raw + l_base + (jj + 1) * l_step);
            const uint4 vb1 = ld_shared_v4(raw + r_base + (jj + 1) * r_step);
            const float4 ta = s_t0[ch0 + jj * 16], tb_ = s_t0[ch0 + jj * 16 + 16];
            const float4 ua = __ldg(p.t1 + ch0 + jj * 16), ub = __ldg(p.t1 + ch0 + jj * 16 + 16);
            __half2 sh = __float2half2_rn(0.f), qh = __float2half2_rn(0.f);
            const uint4 o0 = dw_step_h2_edge<D, A01>(vc0, va0, vb0, ta, ua, rstd, rm, alpha_h2, mk, sh, qh);
            const uint4 o1 = dw_step_h2_edge<D, A01>(vc1, va1, vb1, tb_, ub, rstd, rm, alpha_h2, mk, sh, qh);
            dw_stats_flush(sh, qh, s2, q2);
            if (jj == 0) mbar_wait_b(tb.flags, my_s_empty, sphase ^ 1u);
            st_shared_v4(sa + jj * 2048u, o0);
            st_shared_v4(sa + (jj + 1) * 2048u, o1);
          }
        }
        fence_proxy_async();
        __syncwarp();  // every lane has fenced its stage writes; lane 0 arrives (release.cluster is cumulative)
        if (lane == 0) {
          mbar_arrive_cluster(my_s_full);
          mbar_arrive(raw_empty(rs_));  // (the __syncwarp above also orders every lane's reads of the raw stage)
        }
        if (tl_on) tl_stamp(it, tl_base + 3 * (kb >> 1) + 2);
        ++nseq;
        // statistics of the next tile's segment, off the tile-boundary critical path
        if (kb == 4 + grp && has_next) seg_stats_pre(part1, tn, pre, p.eps1, nmean, nrstd, tb.flags);
      }
      // gLN2 partial of this m-tile: fixed-order block reduction -> one slot per tile
      const float s_acc = warp_sum(s2.x + s2.y);
      const float q_acc = warp_sum(q2.x + q2.y);
      tile_partial_reduce(s_red + (it & 1) * Cfg::PROD_WARPS, s_cnt + (it & 1), Cfg::PROD_WARPS, pw, s_acc, q_acc,
                          valid ? part2 + w : nullptr);
      if (tl_on) tl_stamp(it, 34 + grp);
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer may still arrive on / read from this CTA's shared memory until it is done too
  if (warp == 6) {
    tcgen05_fence_after();
    tmem_dealloc_pair(tmem_base, Cfg::TMEM_COLS);
  }
}

}  // namespace pr2
}  // namespace b200sep

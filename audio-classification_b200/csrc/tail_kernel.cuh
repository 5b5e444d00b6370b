// Fused mask net + mask application + decoder (sm_100a, fp16 tensor-core mode): the `masked` tensor
// (n_src x 512 x F, the largest activation of the forward) never leaves the SM.
//
//   per m-tile (128 frames), for every source s and every 128-filter chunk c of its 512 filters:
//     MMA 1 : score[n][frame] = sum_k W_mask[s*512 + c*128 + n][k] * PReLU(skip_sum)[frame][k]   (K = 128; TRANSPOSED:
//             TMEM lane = filter, column = frame, as in K_B)
//     epi 1 : masked = act(score + b) * tf_rep[c*128 + n][frame]  -> fp16 -> swizzled MN-major smem operand.  A thread
//             owns one filter x 32 consecutive frames: the bias is a per-lane constant and the encoder output of the
//             filter chunk is staged by TMA (four swizzled 64 x 64 boxes) and read as four 16-byte ld.shared per thread
//             (the lane = frame mapping needed one 2-byte load per value and re-read the bias row per thread; direct
//             16-byte global loads of a thread's row piece touch 32 lines per warp instruction and serialised in the LSU)
//     MMA 2 : dec[frame][l] += sum_n masked[frame][n] * W_dec[c*128 + n][l]      (N = L taps, K = 128 / chunk)
//   Chunk order: filter chunk outermost, source innermost -- the encoder-output rows of a chunk are loaded ONCE and
//   multiply all n_src masks (source-outermost re-read them per source: 3 x 128 KiB per m-tile from L2).
//   After the last chunk of a source:
//     epi 2 : overlap-add of the L taps of the 128 frames (stride L/2) + zero tail (pad_x_to_y) -> d_out
//
// (asteroid: mask_net + output_act, `est_masks * tf_rep.unsqueeze(1)`, Decoder(FreeFB) = conv_transpose1d;
//  SURVEY.md A.2.)  Warp roles: warp 0 TMA (skip operand, W_mask chunks, decoder filterbank), warp 1 MMA issuer (both
// MMAs), warp 2 TMEM allocator, warp 3 TMA (encoder rows of the filter chunks), warps 4-19 epilogue (lane = filter; the
// overlap-add runs on all 16 warps with lane = frame, L / 4 taps each, one chunk after the source's last MMA).  TMEM: 2 x 128 score columns (double-buffered) + n_src x L
// decoder columns.  d_out must be zero-filled by the caller (the two hops shared with the neighbouring
// m-tiles are accumulated with atomicAdd: two addends onto zero, order-independent).
#pragma once
#include "common.cuh"
#include "fused_block.cuh"
#include "gemm_tc.cuh"

namespace b200sep {
namespace tail {
using namespace tc;
using fb::mbar_wait_b;
using fb::st_shared_v4;

// 16 fp32 columns of 32 lanes
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t* u = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
        "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
      : "r"(taddr)
      : "memory");
}

// 4 / 8 fp32 columns of 32 lanes
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float (&v)[4]) {
  uint32_t* u = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t* u = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7])
               : "r"(taddr)
               : "memory");
}

template <int L>
struct TailCfg {
  static constexpr int A_BYTES = TILE_M * C_SKIP * 2;     // 32 KiB: PReLU(skip_sum) tile, 2 k-blocks
  static constexpr int B_STAGE = 128 * C_SKIP * 2;        // 32 KiB: 128 mask-conv rows x 128 k
  // weight ring: as many stages as fit next to the encoder-row buffer (three; two measured 5 us slower even with a second
  // masked-chunk staging buffer in exchange); the masked-chunk staging needs only one (the decoder MMAs that read it are
  // issued the moment it is complete and take a few hundred cycles).
  static constexpr int B_STAGES = (L == 16) ? 3 : 2;
  static constexpr int WD_BYTES = L * C_ENC * 2;          // decoder filterbank [L][512] fp16, 8 k-blocks
  static constexpr int ST_BYTES = TILE_M * 128 * 2;       // 32 KiB: masked chunk [128 frames][128 filters]
  static constexpr int ST_STAGES = 1;
  static constexpr int OLA_BYTES = 2 * TILE_M * (L + 1) * 4;   // two buffers (sources alternate)
  static constexpr int OFF_B = A_BYTES;
  static constexpr int OFF_WD = OFF_B + B_STAGES * B_STAGE;
  static constexpr int OFF_ST = OFF_WD + WD_BYTES;
  // encoder-output rows of the current filter chunk (128 filters x 128 frames fp16 as four swizzled 64 x 64 TMA boxes):
  // each epilogue thread needs 64 contiguous bytes of ITS filter row, i.e. a warp-level global load touches 32 different
  // lines -- 64 such loads per filter chunk issued by all 16 warps at once kept the load/store unit busy for ~2.5 k
  // cycles on the chunk-to-chunk critical path (tools/timeline_tail.py); the TMA unit fetches whole rows instead and
  // the threads read their pieces from shared memory (4 conflict-free ld.shared.v4 each)
  static constexpr int ENC_BYTES = 128 * TILE_M * 2;      // 32 KiB
  static constexpr int OFF_ENC = OFF_ST + ST_STAGES * ST_BYTES;
  static constexpr int OFF_OLA = OFF_ENC + ENC_BYTES;
  static constexpr int OFF_BAR = (OFF_OLA + OLA_BYTES + 15) / 16 * 16;
  static constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;
  static constexpr int EPI_WARPS = 16;  // 4 per TMEM lane quadrant, 32 filters each
  static constexpr int THREADS = (4 + EPI_WARPS) * 32;
  static constexpr int TMEM_COLS = 512;
  static constexpr int DEC_COL0 = 256;                    // decoder accumulators start here
  static_assert(SMEM_BYTES <= 232448, "tail kernel shared memory over the 227 KiB limit");
  static_assert(L == 16 || L == 32, "overlap-add mapping: L / 4 taps per warp");
};

struct TailParams {
  const float* bias;   // [S*512] mask conv bias
  const __half* enc;   // [nt][512][128] encoder output
  float* out;          // separated waveforms (zero-filled)
  int n_src;
  int act;             // 0 sigmoid, 1 relu
};

template <int L>
__global__ void __launch_bounds__(TailCfg<L>::THREADS, 1)
mask_dec_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                const __grid_constant__ CUtensorMap tm_wd, const __grid_constant__ CUtensorMap tm_enc, TailParams p,
                Tables tb) {
  using Cfg = TailCfg<L>;
  constexpr int S = L / 2;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t smem_a = smem_base;
  auto smem_b = [&](int s) { return smem_base + (uint32_t)(Cfg::OFF_B + s * Cfg::B_STAGE); };
  const uint32_t smem_wd = smem_base + Cfg::OFF_WD;
  const uint32_t smem_enc = smem_base + Cfg::OFF_ENC;
  auto smem_st = [&](int s) { return smem_base + (uint32_t)(Cfg::OFF_ST + s * Cfg::ST_BYTES); };
  float (*s_ola)[TILE_M][L + 1] = reinterpret_cast<float (*)[TILE_M][L + 1]>(smem_gen + Cfg::OFF_OLA);   // [2]
  const uint32_t bar0 = smem_base + Cfg::OFF_BAR;
  const uint32_t wd_bar = bar0;                                  // decoder filterbank landed (once)
  const uint32_t a_full = bar0 + 8u, a_empty = bar0 + 16u;       // skip operand of the m-tile
  auto b_full = [&](int s) { return bar0 + 8u * (20 + s); };     // [B_STAGES <= 4]
  auto b_empty = [&](int s) { return bar0 + 8u * (24 + s); };
  auto acc_full = [&](int s) { return bar0 + 8u * (7 + s); };
  auto acc_empty = [&](int s) { return bar0 + 8u * (9 + s); };
  auto st_full = [&](int s) { return bar0 + 8u * (11 + s); };
  auto st_empty = [&](int s) { return bar0 + 8u * (13 + s); };
  // decoder accumulator of a source complete: two barriers, sources alternate (the overlap-add of source s runs one
  // chunk late, so a single barrier could run two phases ahead of a waiter and alias its parity)
  auto dec_full = [&](int s) { return bar0 + 8u * (28 + s); };
  const uint32_t dec_empty = bar0 + 8u * 16;                     // ... drained by the overlap-add
  const uint32_t tmem_ptr_smem = bar0 + 8u * 17;
  const uint32_t enc_full = bar0 + 8u * 19, enc_empty = bar0 + 8u * 30;   // encoder rows of a filter chunk
  const uint32_t done_bar = bar0 + 8u * 18;  // final tcgen05.commit: no async mbarrier arrive may outlive the CTA

  pdl_launch_dependents();  // PDL: the prologue reads constants only (see fused_block.cuh)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int n_tiles = tb.n_tiles;
  const int n_chunks = p.n_src * 4;  // 128-filter chunks per m-tile

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_b);
    tma_prefetch_desc(&tm_wd);
    tma_prefetch_desc(&tm_enc);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(wd_bar, 1);
    mbar_init(a_full, 1);
    mbar_init(a_empty, 1);
    for (int s = 0; s < Cfg::B_STAGES; ++s) {
      mbar_init(b_full(s), 1);
      mbar_init(b_empty(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(acc_full(s), 1);
      mbar_init(acc_empty(s), Cfg::EPI_WARPS);   // one arrival per warp
      mbar_init(st_full(s), Cfg::EPI_WARPS);
      mbar_init(st_empty(s), 1);
    }
    mbar_init(dec_full(0), 1);
    mbar_init(dec_full(1), 1);
    mbar_init(dec_empty, Cfg::EPI_WARPS);
    mbar_init(done_bar, 1);
    mbar_init(enc_full, 1);
    mbar_init(enc_empty, Cfg::EPI_WARPS);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_ptr_smem, Cfg::TMEM_COLS);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  pdl_wait();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_smem));

  if (warp == 0) {
    // ================================ TMA producer ================================================
    if (lane == 0) {
      mbar_expect_tx(wd_bar, Cfg::WD_BYTES);
      for (int kb = 0; kb < 8; ++kb) tma_load_2d(smem_wd + (uint32_t)(kb * L * 128), &tm_wd, wd_bar, kb * 64, 0);
      int bs = 0;
      uint32_t bphase = 0;
      int it = 0;
      for (int w = blockIdx.x; w < n_tiles; w += gridDim.x, ++it) {
        const int rot = (load_tile_info(tb.tile_info + w).t0 >> 7) & 3;
        mbar_wait_b(tb.flags, a_empty, ((uint32_t)it & 1u) ^ 1u);
        mbar_expect_tx(a_full, Cfg::A_BYTES);
        for (int kb = 0; kb < 2; ++kb) {
          tma_load_2d(smem_a + (uint32_t)(kb * 16384), &tm_a, a_full, 0, w * C_SKIP + kb * 64);
          tma_load_2d(smem_a + (uint32_t)(kb * 16384 + 8192), &tm_a, a_full, 64, w * C_SKIP + kb * 64);
        }
        for (int c = 0; c < n_chunks; ++c) {
          mbar_wait_b(tb.flags, b_empty(bs), bphase ^ 1u);
          tl_stamp(it, 80 + c);
          mbar_expect_tx(b_full(bs), Cfg::B_STAGE);
          // chunk c = (filter chunk (c / n_src + rot) % 4, source c % n_src); rot staggers the filter-chunk order by the
          // tile's position in its segment, so that the CTAs do not all pull the same 32 KiB of W_mask at the same time
          const int row0 = ((c % p.n_src) * 4 + ((c / p.n_src + rot) & 3)) * 128;
          tma_load_2d(smem_b(bs), &tm_b, b_full(bs), 0, row0);
          tma_load_2d(smem_b(bs) + 16384u, &tm_b, b_full(bs), 64, row0);
          if (++bs == Cfg::B_STAGES) { bs = 0; bphase ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================================ MMA issuer (mask GEMM + decoder GEMM) =======================
    if (lane == 0) {
      // mask GEMM: A = W_mask chunk (K-major), B = skip operand (MN-major, frames contiguous): D[filter][frame]
      constexpr uint32_t idesc_mask = (1u << 4) | (1u << 16) | ((uint32_t)(TILE_M >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      constexpr uint32_t idesc_dec = make_idesc(L);        // A MN-major (masked chunk written by the epilogue), B K-major
      mbar_wait_b(tb.flags, wd_bar, 0);
      int bs = 0;
      uint32_t bphase = 0;
      int n_acc = 0, n_st = 0, n_dec = 0;  // running counters -> stage / phase
      int tl_it = 0;
      auto issue_mask = [&](int c) {
        const int as = n_acc & 1;
        mbar_wait_b(tb.flags, acc_empty(as), (((uint32_t)n_acc >> 1) & 1u) ^ 1u);
        mbar_wait_b(tb.flags, b_full(bs), bphase);
        tcgen05_fence_after();
        tl_stamp(tl_it, c);
        const uint32_t tmem_d = tmem_base + (uint32_t)(as * 128);
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint64_t dx = make_smem_desc(smem_a + (uint32_t)(kb * 16384 + j * 2048), 8192u, 1024u);     // skip, MN-major
            const uint64_t dw = make_smem_desc(smem_b(bs) + (uint32_t)(kb * 16384 + j * 32), 16u, 1024u);     // W_mask, K-major
            umma_f16(tmem_d, dw, dx, idesc_mask, (uint32_t)((kb | j) != 0));
          }
        }
        umma_commit(b_empty(bs));
        umma_commit(acc_full(as));
        if (++bs == Cfg::B_STAGES) { bs = 0; bphase ^= 1u; }
        ++n_acc;
      };
      int it = 0;
      for (int w = blockIdx.x; w < n_tiles; w += gridDim.x, ++it) {
        const int rot = (load_tile_info(tb.tile_info + w).t0 >> 7) & 3;
        mbar_wait_b(tb.flags, a_full, (uint32_t)it & 1u);
        tl_it = it;
        issue_mask(0);
        for (int c = 0; c < n_chunks; ++c) {
          if (c + 1 < n_chunks) issue_mask(c + 1);  // keeps the tensor core busy while chunk c is in its epilogue
          else umma_commit(a_empty);                // all mask MMAs of this m-tile issued: operand reusable when done
          const int src = c % p.n_src, cc = c / p.n_src;
          const int ss = 0;
          // (cc == 0: the previous m-tile's overlap-adds -- all of them, they finish in order -- drained the accumulators)
          if (cc == 0) mbar_wait_b(tb.flags, dec_empty, ((uint32_t)n_dec & 1u) ^ 1u);
          mbar_wait_b(tb.flags, st_full(ss), (uint32_t)n_st & 1u);
          tcgen05_fence_after();
          tl_stamp(it, 16 + c);
          const uint32_t tmem_dec = tmem_base + (uint32_t)(Cfg::DEC_COL0 + src * L);
#pragma unroll
          for (int kb = 0; kb < 2; ++kb) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint64_t da = make_smem_desc(smem_st(ss) + (uint32_t)(kb * 16384 + j * 2048), 8192u, 1024u);  // MN-major
              const uint64_t db = make_smem_desc(smem_wd + (uint32_t)((((cc + rot) & 3) * 2 + kb) * L * 128 + j * 32), 16u, 1024u);
              umma_f16(tmem_dec, da, db, idesc_dec, (uint32_t)((cc | kb | j) != 0));
            }
          }
          umma_commit(st_empty(ss));
          if (cc == 3) { umma_commit(dec_full(n_dec & 1)); ++n_dec; }
          ++n_st;
        }
      }
      umma_commit(done_bar);  // the last b_empty / st_empty / a_empty arrivals have no waiter
      mbar_wait_b(tb.flags, done_bar, 0);
    }
    __syncwarp();
  } else if (warp == 3) {
    // ================================ TMA producer: encoder rows of the filter chunks =============
    if (lane == 0) {
      int n_enc = 0;
      for (int w = blockIdx.x; w < n_tiles; w += gridDim.x) {
        const int rot = (load_tile_info(tb.tile_info + w).t0 >> 7) & 3;
        for (int cc = 0; cc < 4; ++cc, ++n_enc) {
          mbar_wait_b(tb.flags, enc_empty, ((uint32_t)n_enc & 1u) ^ 1u);  // every epilogue warp has converted the last chunk
          mbar_expect_tx(enc_full, Cfg::ENC_BYTES);
          const int row0 = w * C_ENC + ((cc + rot) & 3) * 128;
          for (int rb = 0; rb < 2; ++rb)
            for (int fbx = 0; fbx < 2; ++fbx)
              tma_load_2d(smem_enc + (uint32_t)((rb * 2 + fbx) * 8192), &tm_enc, enc_full, fbx * 64, row0 + rb * 64);
        }
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ================================ epilogue (16 warps) =========================================
    // warp = (TMEM lane quadrant = 32 filters of the chunk, 32-frame quarter of the m-tile); a thread owns one filter x
    // 32 frames, processed as two 16-column pieces (small live register set: 640 threads share the register file).
    const int ew = warp - 4, quad = ew & 3, part = ew >> 2;
    const int fl = quad * 32 + lane;             // filter within the 128-filter chunk == TMEM lane
    const int r = fl;                            // (the overlap-add below runs with lane = frame)
    int n_acc = 0, n_st = 0, n_dec = 0, n_enc = 0;
    const bool tl_on = (ew == 0 && lane == 0);
    // this thread's 64 bytes (32 frames) of its filter row inside the staged chunk: box (filter / 64, frame half), row
    // filter % 64, four 16-byte pieces at the swizzled positions
    const uint32_t enc_row = smem_enc + (uint32_t)(((fl >> 6) * 2 + (part >> 1)) * 8192 + (fl & 63) * 128);
    int tl_it = -1;
    for (int w = blockIdx.x; w < n_tiles; w += gridDim.x) {
      ++tl_it;
      const TileInfo ti = load_tile_info(tb.tile_info + w);
      const int seg = ti.seg, t0 = ti.t0, F = ti.F;
      const int T_len = tb.seg_len[seg];
      const int nvalid = F - t0 - part * 32;       // frames of this thread's 32 that lie inside the segment
      const int rot = (t0 >> 7) & 3;               // filter-chunk order of this tile (see the TMA producer)
      // encoder output of this thread's filter x 32 frames in fp32, pre-scaled by 0.5 for the sigmoid form below; loaded
      // and converted once per filter chunk, used by every source
      float2 ef[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) ef[j] = make_float2(0.f, 0.f);
      const float esc = (p.act == 0) ? 0.5f : 1.0f;
      float bias_raw = __ldg(p.bias + (rot & 3) * 128 + fl);   // chunk 0: source 0, filter chunk rot
      float* const o_seg = p.out + tb.seg_out_off[seg];        // (loaded per tile, not behind the overlap-add's barrier)
      auto ola = [&](int src) {
        constexpr int LP = L / 4;                  // taps per warp of a quadrant
        float (*buf)[L + 1] = s_ola[n_dec & 1];
        mbar_wait_b(tb.flags, dec_full(n_dec & 1), ((uint32_t)n_dec >> 1) & 1u);
        tcgen05_fence_after();
        float taps[LP];
        const uint32_t daddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(Cfg::DEC_COL0 + src * L + part * LP);
        if (LP == 4) tmem_ld4(daddr, reinterpret_cast<float (&)[4]>(taps));
        else tmem_ld8(daddr, reinterpret_cast<float (&)[8]>(taps));
        tmem_ld_wait();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(dec_empty);
#pragma unroll
        for (int l = 0; l < LP; ++l) buf[r][part * LP + l] = taps[l];
        // (one barrier per source: the buffer written here was last read two sources ago, and every thread has passed
        // the barrier of the source in between since)
        fb::named_bar_sync(2, Cfg::EPI_WARPS * 32);
        float* o = o_seg + (size_t)src * T_len;
        const int tid = ew * 32 + lane;
        // hops hh = 0..128 of this tile: sample (t0 + hh) * S + l gets frame hh's first half and frame
        // hh-1's second half
        for (int i = tid; i < (TILE_M + 1) * S; i += Cfg::EPI_WARPS * 32) {
          const int hh = i / S, l = i % S;
          const int f_hi = t0 + hh, f_lo = t0 + hh - 1;
          float val = 0.f;
          bool any = false;
          if (hh < TILE_M && f_hi < F) { val += buf[hh][l]; any = true; }
          if (hh >= 1 && f_lo < F) { val += buf[hh - 1][S + l]; any = true; }
          if (!any) continue;
          if (!(fabsf(val) <= 3.0e38f)) atomicOr(tb.flags, 2);  // inf / NaN samples: flagged for the host
          const size_t idx = (size_t)(t0 + hh) * S + l;
          if (hh == 0 || hh == TILE_M) atomicAdd(o + idx, val);
          else o[idx] = val;
        }
        ++n_dec;
      };
      for (int c = 0; c < n_chunks; ++c) {
        const int src = c % p.n_src, cc = c / p.n_src;
        const int as = n_acc & 1, ss = 0;
        // (this chunk's bias was requested one chunk ago: read at its first use it was a long-scoreboard stall per chunk)
        const float bias = bias_raw * esc;             // (0.5 b for the sigmoid form)
        if (c + 1 < n_chunks)
          bias_raw = __ldg(p.bias + ((c + 1) % p.n_src) * C_ENC + (((c + 1) / p.n_src + rot) & 3) * 128 + fl);
        // encoder output of this filter for the 32 frames, once per filter chunk (it serves all sources)
        if (src == 0) {
          mbar_wait_b(tb.flags, enc_full, (uint32_t)n_enc & 1u);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint4 e4 = fb::ld_shared_v4(enc_row + (uint32_t)(((((part & 1) * 4 + j) ^ (fl & 7))) << 4));
            const uint32_t ew4[4] = {e4.x, e4.y, e4.z, e4.w};
#pragma unroll
            for (int k = 0; k < 4; ++k)
              ef[j * 4 + k] = mul2(__half22float2(*reinterpret_cast<const __half2*>(&ew4[k])), make_float2(esc, esc));
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(enc_empty);   // the next filter chunk's rows may land
          ++n_enc;
        }
        if (tl_on) tl_stamp(tl_it, 32 + c);
        mbar_wait_b(tb.flags, acc_full(as), ((uint32_t)n_acc >> 1) & 1u);
        tcgen05_fence_after();
        if (tl_on) tl_stamp(tl_it, 48 + c);
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * 128 + part * 32);
        uint32_t pk[16];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float v[16];
          tmem_ld16(taddr + (uint32_t)(h * 16), v);
          tmem_ld_wait();
          if (h == 1) {
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty(as));  // the warp's columns are in registers: the next mask MMA may overwrite them
            ++n_acc;
          }
          // sigmoid(a) e = (0.5 tanh(0.5 a) + 0.5) e = t e' + e'  with  t = tanh(0.5 v + 0.5 b),  e' = 0.5 e:
          // one packed FMA, one MUFU per element, one packed FMA per pair (was ~10 scalar instructions per element)
          const float2 b2 = make_float2(bias, bias);
#pragma unroll
          for (int j = 0; j < 16; j += 2) {
            const float2 e2 = ef[h * 8 + (j >> 1)];
            float2 m2;
            if (p.act == 0) {
              const float2 a2 = fma2(make_float2(v[j], v[j + 1]), make_float2(0.5f, 0.5f), b2);
              float2 t2;
              asm("tanh.approx.f32 %0, %1;" : "=f"(t2.x) : "f"(a2.x));
              asm("tanh.approx.f32 %0, %1;" : "=f"(t2.y) : "f"(a2.y));
              m2 = fma2(t2, e2, e2);
            } else {
              const float2 a2 = add2(make_float2(v[j], v[j + 1]), b2);
              m2 = mul2(make_float2(fmaxf(a2.x, 0.f), fmaxf(a2.y, 0.f)), e2);
            }
            const __half2 hh2 = __floats2half2_rn(m2.x, m2.y);
            pk[h * 8 + (j >> 1)] = *reinterpret_cast<const uint32_t*>(&hh2);
          }
        }
        if (nvalid < 32) {   // (warp-uniform, last tile of a segment only) frames beyond the segment end contribute nothing
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            if (2 * j >= nvalid) pk[j] = 0u;
            else if (2 * j + 1 >= nvalid) pk[j] &= 0xffffu;
          }
        }
        // masked row -> MN-major SWIZZLE_128B operand (the image of K_A's h2 stage): k-block = filter / 64, row =
        // filter % 64, 8 frames per 16-byte chunk
        mbar_wait_b(tb.flags, st_empty(ss), ((uint32_t)n_st & 1u) ^ 1u);
        const uint32_t row_base = smem_st(ss) + (uint32_t)((fl >> 6) * 16384);
#pragma unroll
        for (int q = 0; q < 4; ++q)
          st_shared_v4(row_base + fb::a_chunk_off(fl & 63, part * 32 + q * 8),
                       make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]));
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(st_full(ss));
        if (tl_on) tl_stamp(tl_it, 64 + c);
        ++n_st;

        if (cc == 3) {   // (the last filter chunk: the sources complete one after the other)
          // decoder output of a source: taps -> overlap-add -> global, by ALL 16 epilogue warps (warp (quad, part) drains
          // taps [part L/4, +L/4) of its quadrant's 32 frames; the 129 x S output samples are dealt over the 512 threads).
          // Source s is drained after chunk s + 1 has been staged (its decoder MMAs have long completed by then); only the
          // last source waits for its own MMAs.  The 4 warps of part 0 used to do all of this while the other 12 waited
          // for them at the next staging barrier (~3.5 k cycles per source on the chunk-to-chunk critical path).
          if (src > 0) ola(src - 1);
          if (src == p.n_src - 1) ola(src);
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

}  // namespace tail
}  // namespace b200sep

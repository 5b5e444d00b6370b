// Fused TCN-block kernels (sm_100a, fp16 tensor-core mode).  One Conv1DBlock is two
// launches; the kernel boundaries are exactly the two segment-wide gLN reductions of the block:
//
//   K_B  apply_in_kernel :  x += gLN2-scaled residual output of the previous block (elementwise
//                           prologue on the raw G_res), x -> fp16 operand in shared memory -> tcgen05
//                           in-conv GEMM (128 -> 512, W_in resident in shared memory) -> bias + PReLU
//                           -> h1 (fp16) + gLN1 partial sums.
//   K_A  dw_rs_kernel    :  raw h1 k-blocks (+ dilation halos) staged into shared memory by TMA; 16
//                           producer warps fold gLN1 into the dilated depthwise taps, apply PReLU and
//                           write the result straight into the swizzled shared-memory A-operand ring
//                           (h2 never leaves the SM) -> tcgen05 [res | skip] GEMM (512 -> 256) on the
//                           RAW h2 with gamma-folded weights -> raw G_res (fp16) + raw G_skip (fp16,
//                           one stash slot per block) + gLN2 partial sums.
//
// G is "raw": the two per-segment gLN2 scalars are only known once every tile of the segment has been
// through K_A, so they are applied by the consumer:
//     x_new    = x + vm * (rstd2 * (G_res - mean2 * rs) + cb)              (next K_B prologue)
//     skip_sum = sum_b vm * (rstd2_b * (G_skip_b - mean2_b * rs_b) + cb_b) (skip_merge_kernel, once)
// HBM bytes per frame and block: K_A 1024 (h1) + 512 (G); K_B 256 (G_res) + 1024 (x r/w) + 1024 (h1);
// + 256 per block read once by the skip merge = 4.1 KB vs 6.6 KB for the three unfused kernels.
//
// Every mbarrier wait is bounded (Tables::flags): a protocol bug degrades into wrong numbers plus an error
// flag the host reads, never into a hung GPU.
#pragma once
#include <type_traits>
#include "bandwidth_kernels.cuh"
#include "common.cuh"
#include "epilogues.cuh"
#include "gemm_tc.cuh"

namespace b200sep {

// Device flags of ONE forward call live in the call's table block (Tables::flags, zeroed by the table upload):
// bit 0: a bounded wait ran out; bit 1: non-finite output samples.  The host gets them through the `h_flags` argument
// of b200sep_separate_batch[_host].
constexpr int FLAG_WAIT_TIMEOUT = 1, FLAG_NONFINITE = 2;  // (bit 2, eps-dominated gLN: common.cuh flag_eps_dominated)

// Optional in-kernel timeline (b200sep_debug_timeline): clock64 stamps per (CTA, tile visit, slot).
constexpr int TL_TILES = 8, TL_SLOTS = 128;
__constant__ long long* c_timeline = nullptr;
__device__ __forceinline__ void tl_stamp(int it, int slot) {
  long long* tl = c_timeline;
  if (tl != nullptr && it < TL_TILES) tl[((size_t)blockIdx.x * TL_TILES + it) * TL_SLOTS + slot] = clock64();
}
// kernel-level stamps (entry / dependency wait returned / exit) in globaltimer nanoseconds AND clock64, in the row of the
// last visit slot (single-visit launches -- the one-call latency path -- leave it free): tools/single_call_timeline.py
__device__ __forceinline__ void tl_stamp_ns(int slot) {
  long long* tl = c_timeline;
  if (tl != nullptr) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    long long* row = tl + ((size_t)blockIdx.x * TL_TILES + (TL_TILES - 1)) * TL_SLOTS;
    row[slot] = (long long)t;
    row[slot + 4] = clock64();
  }
}

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
      : "r"(bar), "r"(parity), "r"(20000)
      : "memory");
  return ok != 0;
}
// Bounded wait: ~seconds in the worst case, then flags the abort and lets the caller run on (all
// addresses stay in bounds; the host reports the failure).
__device__ __forceinline__ void mbar_wait_b(int* flags, uint32_t bar, uint32_t parity) {
#pragma unroll 1
  for (uint32_t i = 0; i < (1u << 18); ++i) {
    if (mbar_try(bar, parity)) return;
    if ((i & 31u) == 31u && flags != nullptr && (*reinterpret_cast<volatile int*>(flags) & FLAG_WAIT_TIMEOUT) != 0) return;
  }
  if (flags != nullptr) atomicOr(flags, FLAG_WAIT_TIMEOUT);
}
// L2 prefetch of a contiguous global range (bytes % 16 == 0) / of a TMA box
__device__ __forceinline__ void l2_prefetch_bulk(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
__device__ __forceinline__ void l2_prefetch_tma_2d(const void* desc, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(desc)),
               "r"(c0), "r"(c1)
               : "memory");
}
// Wait-free block reduction of per-warp (sum, sumsq) partials of one m-tile: every warp publishes its pair
// and bumps a shared counter; whichever warp arrives last adds the n_warps pairs in index order (so the
// result does not depend on who that is) and writes the tile's single partial.  `red` / `cnt` are double-
// buffered by tile parity by the caller; warps of one role are never two tiles apart (they meet at the
// pipeline mbarriers every chunk), so a buffer is free again when it is reused.  (A bar.arrive / bar.sync
// pair is NOT safe here: a fast warp can arrive twice in one barrier phase and hang the waiting warp.)
__device__ __forceinline__ void tile_partial_reduce(float2* red, int* cnt, int n_warps, int my_warp, float s, float q,
                                                    float2* out) {
  if ((threadIdx.x & 31) == 0) {
    red[my_warp] = make_float2(s, q);
    __threadfence_block();
    const int prev = atomicAdd(cnt, 1);
    if (prev == n_warps - 1) {
      __threadfence_block();
      float ts = 0.f, tq = 0.f;
      for (int k = 0; k < n_warps; ++k) {
        const volatile float* v = reinterpret_cast<const volatile float*>(red + k);
        ts += v[0]; tq += v[1];
      }
      *cnt = 0;
      if (out != nullptr) *out = make_float2(ts, tq);
    }
  }
}
// ---- 2-CTA cluster helpers (used by the CTA-pair K_A, fused_pair.cuh) --------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// ---- cta_group::2 (CTA pair) helpers ---------------------------------------------------------------------------
// shared::cluster address of `addr` (a shared::cta address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_rank(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// arrive on an mbarrier anywhere in the cluster.  Default semantics (release at CTA scope), as CUTLASS's
// ClusterBarrier::arrive(cta_id) does: the data the peer's tensor core reads afterwards was made visible to the async
// proxy by the writers' fence.proxy.async.  (`.release.cluster` here compiles to MEMBAR.ALL.GPU + CGAERRBAR in front of
// every arrive -- measured: the CTA-pair kernel ran 105 instead of 92 us with it.)
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// commit of the pair's MMAs: one arrival on the barrier at this offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_pair(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"(mask)
               : "memory");
}
// TMA tile store shared -> global (bulk async group of the issuing thread)
__device__ __forceinline__ void tma_store_2d(const void* desc, uint32_t smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(desc)),
               "r"(smem_src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void st_shared_u16(uint32_t addr, uint16_t v) {
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(v) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
// byte offset of the 16-byte chunk (8 consecutive frames starting at frame f0, f0 % 8 == 0) of
// k-row `row` (0..63) inside one 64(k) x 128(m) MN-major SWIZZLE_128B operand stage: two 64x64 boxes
// 8 KiB apart (frames 0-63 / 64-127), 128-byte rows, chunk index xor (row & 7) -- the image a TMA
// box load with CU_TENSOR_MAP_SWIZZLE_128B produces (gemm_tc.cuh).
__device__ __forceinline__ uint32_t a_chunk_off(int row, int f0) {
  const int chunk = (f0 >> 3) & 7;
  return (uint32_t)((f0 >> 6) * 8192 + row * 128 + ((chunk ^ (row & 7)) << 4));
}
__device__ __forceinline__ uint4 ldg_v4(const __half* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
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

// Per-CTA copy of the TileInfo records of the m-tiles this CTA will visit (tile it = blockIdx.x +
// it * gridDim.x), so no tile starts with a dependent global load chain.
constexpr int TI_CACHE = 32;
__device__ __forceinline__ void ti_cache_fill(TileInfo* s_ti, const Tables& tb, int tid) {
  if (tid < TI_CACHE) {
    const int w = blockIdx.x + tid * gridDim.x;
    if (w < tb.n_tiles) s_ti[tid] = load_tile_info(tb.tile_info + w);
  }
}
__device__ __forceinline__ TileInfo ti_get(const TileInfo* s_ti, const Tables& tb, int it, int w) {
  if (it < TI_CACHE) return s_ti[it];
  return load_tile_info(tb.tile_info + w);
}
// Segment gLN statistics from single-slot per-tile partials; `pre` is this lane's (pre-fetched) first
// element part[tile0 + lane] (zeros when lane >= ntiles).  Same arithmetic as warp_seg_stats.
__device__ __forceinline__ void seg_stats_pre(const float2* __restrict__ part, const TileInfo& ti, float2 pre,
                                              float eps, float& mean, float& rstd, int* flags = nullptr) {
  const int lane = threadIdx.x & 31;
  double s = (double)pre.x, q = (double)pre.y;
  for (int i = lane + 32; i < ti.ntiles; i += 32) {
    const float2 v = part[ti.tile0 + i];
    s += (double)v.x;
    q += (double)v.y;
  }
  s = warp_sum(s);
  q = warp_sum(q);
  const double m = s * ti.inv_count;
  double var = fma(-m, m, q * ti.inv_count);
  if (var < 0.0) var = 0.0;
  mean = (float)m;
  const float ve = (float)var + eps;
  float y = rsqrtf(ve);
  y = y * (1.5f - 0.5f * ve * y * y);
  rstd = y;
  flag_eps_dominated(flags, (float)var, eps);
}
__device__ __forceinline__ float2 part_prefetch(const float2* __restrict__ part, const TileInfo& ti) {
  const int lane = threadIdx.x & 31;
  return (lane < ti.ntiles) ? part[ti.tile0 + lane] : make_float2(0.f, 0.f);
}

// =================================================================================================
// K_A: depthwise stage (producer warps) + [res | skip] GEMM -> raw G
// =================================================================================================
// Static per-block tap table (built on the host, fp64): for channel c with depthwise taps w0..w2,
// gLN1 gamma/beta and depthwise bias,
//   t0[c] = { w0*gamma, w1*gamma, w2*gamma, (w0+w1+w2)*gamma }
//   t1[c] = { w0*beta,  w1*beta,  w2*beta,  bias }
//   cb[c] = (w0+w1+w2)*beta + bias
// so that with the segment's (mean, rstd):  wa_k = t0_k * rstd,  wb_k = t1_k - t0_k * rstd*mean, and for
// interior frames  u = cb - t0.w * rstd*mean + sum_k wa_k * h[t + (k-1) d].
struct DwTab {
  const float4* t0;  // [512]
  const float4* t1;  // [512]
  const float* cb;   // [512]
  float alpha2;      // PReLU slope after the depthwise conv
  int dilation;
  float eps1;        // eps of gLN1 (shared_block.2)
};

template <int BN>
struct KaCfg {
  // Shared-memory plan (dynamic, always the full 227 KiB; offsets from the 1 KiB-aligned base):
  //   [A ring: 2 x 16 KiB, one stage per producer group][tap tables 18 KiB][TileInfo cache][reduction scratch]
  //   [barriers] then [raw h1 ring: 2 stages of 64 ch x (128 + 2 hw) frames][B ring: as many 64-k weight stages
  //   as still fit, 2..4].
  // The B ring is decoupled from the A ring and as deep as possible: a weight stage can only be re-requested
  // when the MMA that read it has completed, and its L2 -> shared-memory latency (~1.7 k cycles) is three times
  // the MMA time of a k-block (512 cycles); with 2 stages that latency, not the work, paced the kernel.
  static constexpr int A_STAGE = TILE_M * 64 * 2;    // 16 KiB
  static constexpr int B_STAGE = BN * 64 * 2;  // 32 KiB
  static constexpr int B_MAX_STAGES = 4;
  static constexpr int TAB_BYTES = C_HID * (2 * 8 + 16);  // per-visit folded tap table (two buffers) + static t1: 16 KiB
  static constexpr int EPI_WARPS = 4, PROD_WARPS = 16;
  static constexpr int W_STAT = EPI_WARPS + 3 + PROD_WARPS;          // statistics warp (the 24th warp costs no registers:
  static constexpr int THREADS = (W_STAT + 1) * 32;                  // 768     23 and 24 warps get the same 80 per thread)
  static constexpr int TMEM_COLS = 2 * BN;
  // A stages per producer group: with 2 the producers of a group write k-block n + 1 while the MMA of k-block n (and, in
  // pair mode, the cross-SM arrive / commit round trip) is still in flight
  static constexpr int A_PG = 1;
  static_assert(A_PG == 1 || A_PG == 2, "A stages per group");
  static constexpr int OFF_A = 0;
  static constexpr int OFF_TAB = OFF_A + 2 * A_PG * A_STAGE;
  static constexpr int OFF_TI = OFF_TAB + TAB_BYTES;
  static constexpr int OFF_RED = OFF_TI + TI_CACHE * 32;
  static constexpr int OFF_CNT = OFF_RED + 2 * PROD_WARPS * 8;
  static constexpr int OFF_STAT = OFF_CNT + 16;  // [2] (mean, rstd) of the visit's segment
  static constexpr int OFF_BAR = OFF_STAT + 16;
  static constexpr int OFF_RAW = (OFF_BAR + 512 + 1023) / 1024 * 1024;  // raw ring, then B ring
  static constexpr int SMEM_BYTES = 232448;          // everything an SM offers to one CTA
  static constexpr int SMEM_USABLE = SMEM_BYTES - 1024 /*alignment slack*/;
};

// Interior-tile producer step in packed half arithmetic: 8 outputs of one channel from the three 8-frame
// vectors.  Taps, PReLU and the operand rounding happen in fp16 (HFMA2 / HMUL2 / HMNMX2: no fp16 -> fp32
// conversions on the tap path); only the gLN2 statistics are accumulated in fp32.
// d = a * b + c with fp16 a, b and fp32 c, d (sm_100 mixed-precision FMA: no conversion instruction)
__device__ __forceinline__ float fma_f32_f16(uint16_t a, uint16_t b, float c) {
  float d;
  asm("fma.rn.f32.f16 %0, %1, %2, %3;" : "=f"(d) : "h"(a), "h"(b), "f"(c));
  return d;
}
__device__ __forceinline__ uint32_t h2_as_u32(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }
__device__ __forceinline__ __half2 u32_as_h2(uint32_t u) { return *reinterpret_cast<__half2*>(&u); }
// left / right neighbour pairs of the thread's four fp16 pairs (8 frames) from the centre vector and the two neighbouring
// vectors: D == 0 -> the neighbour vectors ARE the taps (d % 8 == 0); D in {1, 2, 4} -> picked across the seams.
template <int D>
__device__ __forceinline__ void dw_lr(const uint4 vc, const uint4 va, const uint4 vb, uint32_t (&c)[4], uint32_t (&l)[4],
                                      uint32_t (&r)[4]) {
  c[0] = vc.x; c[1] = vc.y; c[2] = vc.z; c[3] = vc.w;
  const uint32_t pv[4] = {va.x, va.y, va.z, va.w};
  const uint32_t nv[4] = {vb.x, vb.y, vb.z, vb.w};
  if (D == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) { l[i] = pv[i]; r[i] = nv[i]; }
  } else if (D == 1) {
    // pairs shifted by one frame: (hi of word i-1, lo of word i)
    const uint32_t m0 = __byte_perm(pv[3], c[0], 0x5432), m1 = __byte_perm(c[0], c[1], 0x5432);
    const uint32_t m2 = __byte_perm(c[1], c[2], 0x5432), m3 = __byte_perm(c[2], c[3], 0x5432);
    const uint32_t m4 = __byte_perm(c[3], nv[0], 0x5432);
    l[0] = m0; l[1] = m1; l[2] = m2; l[3] = m3;
    r[0] = m1; r[1] = m2; r[2] = m3; r[3] = m4;
  } else if (D == 2) {
    l[0] = pv[3]; l[1] = c[0]; l[2] = c[1]; l[3] = c[2];
    r[0] = c[1]; r[1] = c[2]; r[2] = c[3]; r[3] = nv[0];
  } else {  // D == 4
    l[0] = pv[2]; l[1] = pv[3]; l[2] = c[0]; l[3] = c[1];
    r[0] = c[2]; r[1] = c[3]; r[2] = nv[0]; r[3] = nv[1];
  }
}
template <bool A01>
__device__ __forceinline__ __half2 prelu_h2(__half2 u, const __half2 alpha2) {
  if (A01) return __hmax2(u, __hmul2(u, alpha2));
  return __hadd2(__hmax2(u, __float2half2_rn(0.f)), __hmul2(alpha2, __hmin2(u, __float2half2_rn(0.f))));
}
// Edge-tile producer step (a tap or a frame of the tile falls outside the segment), still in packed halves.  Zero
// padding applies to the NORMALISED signal at the true segment ends, so every tap carries its own constant term
// (a_k h + b_k) and is switched off as a whole by a bit mask (0xffff / 0 per fp16 lane: exact zero whatever the
// masked-out raw value was -- the neighbouring segment's data or the TMA zero fill); frames >= F produce zeros.
//   t0 = {w0 g, w1 g, w2 g, .}, t1 = {w0 beta, w1 beta, w2 beta, bias}:  a_k = fp16(t0_k rstd),  b_k = t1_k - a_k mean
// (b_k from the ROUNDED a_k: the per-visit table only holds the rounded taps)
struct DwEdgeH { uint32_t ml[4], mc[4], mr[4]; };
// The producer steps read the per-visit tap table: tap = {a0, a1, a2, cw} as four halves, prepared once per visit and
// channel by the statistics warp; broadcast to half2 pairs here (PRMT, on the otherwise idle ALU pipe).
__device__ __forceinline__ __half2 bcast_lo(uint32_t u) { return u32_as_h2(__byte_perm(u, u, 0x1010)); }
__device__ __forceinline__ __half2 bcast_hi(uint32_t u) { return u32_as_h2(__byte_perm(u, u, 0x3232)); }
template <int D, bool A01>
__device__ __forceinline__ uint4 dw_step_h2p(const uint4 vc, const uint4 va, const uint4 vb, const uint2 tap,
                                             const __half2 alpha2, __half2& sh, __half2& qh) {
  uint32_t c[4], l[4], r[4];
  dw_lr<D>(vc, va, vb, c, l, r);
  const __half2 w0 = bcast_lo(tap.x), w1 = bcast_hi(tap.x), w2 = bcast_lo(tap.y), cw = bcast_hi(tap.y);
  uint32_t o[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __half2 u = __hfma2(u32_as_h2(l[i]), w0, __hfma2(u32_as_h2(c[i]), w1, __hfma2(u32_as_h2(r[i]), w2, cw)));
    u = prelu_h2<A01>(u, alpha2);
    o[i] = h2_as_u32(u);
    sh = __hadd2(sh, u);
    qh = __hfma2(u, u, qh);
  }
  return make_uint4(o[0], o[1], o[2], o[3]);
}
// edge tiles: b_k = t1_k - a_k mean with the table's (rounded) a_k = fp16(t0_k rstd)
template <int D, bool A01>
__device__ __forceinline__ uint4 dw_step_h2p_edge(const uint4 vc, const uint4 va, const uint4 vb, const uint2 tap,
                                                  const float4 t1, const float mean, const __half2 alpha2,
                                                  const DwEdgeH& mk, __half2& sh, __half2& qh) {
  uint32_t c[4], l[4], r[4];
  dw_lr<D>(vc, va, vb, c, l, r);
  const __half2 w0 = bcast_lo(tap.x), w1 = bcast_hi(tap.x), w2 = bcast_lo(tap.y);
  const __half2 b0 = __float2half2_rn(fmaf(-__low2float(w0), mean, t1.x));
  const __half2 b1 = __float2half2_rn(fmaf(-__low2float(w1), mean, t1.y));
  const __half2 b2 = __float2half2_rn(fmaf(-__low2float(w2), mean, t1.z));
  const __half2 bias = __float2half2_rn(t1.w);
  uint32_t o[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t ta = h2_as_u32(__hfma2(u32_as_h2(l[i]), w0, b0)) & mk.ml[i];
    const uint32_t tc = h2_as_u32(__hfma2(u32_as_h2(c[i]), w1, b1)) & mk.mc[i];
    const uint32_t tr = h2_as_u32(__hfma2(u32_as_h2(r[i]), w2, b2)) & mk.mr[i];
    __half2 u = __hadd2(__hadd2(u32_as_h2(ta), u32_as_h2(tc)), __hadd2(u32_as_h2(tr), bias));
    u = u32_as_h2(h2_as_u32(prelu_h2<A01>(u, alpha2)) & mk.mc[i]);
    o[i] = h2_as_u32(u);
    sh = __hadd2(sh, u);
    qh = __hfma2(u, u, qh);
  }
  return make_uint4(o[0], o[1], o[2], o[3]);
}
// bit masks of the thread's 8 frames (segment-relative t_base .. t_base + 7) for dilation d in a segment of F frames
__device__ __forceinline__ void dw_edge_masks(DwEdgeH& mk, int t_base, int d, int F) {
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int ta = t_base + 2 * k, tb2 = ta + 1;
    mk.ml[k] = ((ta - d >= 0 && ta - d < F) ? 0xffffu : 0u) | ((tb2 - d >= 0 && tb2 - d < F) ? 0xffff0000u : 0u);
    mk.mc[k] = (ta < F ? 0xffffu : 0u) | (tb2 < F ? 0xffff0000u : 0u);
    mk.mr[k] = (ta + d < F ? 0xffffu : 0u) | (tb2 + d < F ? 0xffff0000u : 0u);
  }
}

// fp16 partial statistics of (at most) two producer steps -> the thread's fp32 accumulators
__device__ __forceinline__ void dw_stats_flush(__half2& sh, __half2& qh, float2& s2, float2& q2) {
  const uint32_t su = h2_as_u32(sh), qu = h2_as_u32(qh);
  s2.x = fma_f32_f16((uint16_t)(su & 0xffffu), (uint16_t)0x3c00u, s2.x);  // + lo * 1.0
  s2.y = fma_f32_f16((uint16_t)(su >> 16), (uint16_t)0x3c00u, s2.y);
  q2.x = fma_f32_f16((uint16_t)(qu & 0xffffu), (uint16_t)0x3c00u, q2.x);
  q2.y = fma_f32_f16((uint16_t)(qu >> 16), (uint16_t)0x3c00u, q2.y);
  sh = __float2half2_rn(0.f);
  qh = __float2half2_rn(0.f);
}

template <int BN, int D, bool A01>
__global__ void __launch_bounds__(KaCfg<256>::THREADS, 1)
dw_rs_kernel(const __grid_constant__ CUtensorMap tm_b, const __grid_constant__ CUtensorMap tm_hc,
             const __grid_constant__ CUtensorMap tm_hh, __half* __restrict__ g_res, __half* __restrict__ g_skip,
             size_t skip_tile_stride, const float2* __restrict__ part1, float2* __restrict__ part2, DwTab p,
             Tables tb) {
  using Cfg = KaCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  auto stage_a = [&](int s) { return smem_base + (uint32_t)(Cfg::OFF_A + s * Cfg::A_STAGE); };
  // s_tap[b][c]: the channel's folded taps for the segment of visit it (b = it & 1) as four halves
  //   { w0 g rstd, w1 g rstd, w2 g rstd, cb - (w0 + w1 + w2) g rstd mean }
  // written by the statistics warp one visit ahead (the producers used to redo these 8 instructions per 8 outputs and
  // the segment statistics in the middle of every tile)
  uint2* s_tap = reinterpret_cast<uint2*>(smem_gen + Cfg::OFF_TAB);
  float4* s_t1 = reinterpret_cast<float4*>(s_tap + 2 * C_HID);
  float2* s_stat = reinterpret_cast<float2*>(smem_gen + Cfg::OFF_STAT);   // [2]
  TileInfo* s_ti = reinterpret_cast<TileInfo*>(smem_gen + Cfg::OFF_TI);
  float2* s_red = reinterpret_cast<float2*>(smem_gen + Cfg::OFF_RED);  // [2][16]
  int* s_cnt = reinterpret_cast<int*>(smem_gen + Cfg::OFF_CNT);          // [2]
  const uint32_t bar0 = smem_base + Cfg::OFF_BAR;
  auto raw_full = [&](int s) { return bar0 + 8u * (30 + s); };     // [2 or 4]  TMA -> producers (stage s belongs to group s & 1)
  auto raw_empty = [&](int s) { return bar0 + 8u * (34 + s); };    // [2 or 4]
  auto a_full = [&](int s) { return bar0 + 8u * (22 + s); };       // [2 * A_PG]  producers -> MMA (stage = group * A_PG + n % A_PG)
  auto a_empty = [&](int s) { return bar0 + 8u * (26 + s); };      // [2 * A_PG]
  auto b_full = [&](int s) { return bar0 + 8u * (8 + s); };        // [4]  TMA (weights) -> MMA
  auto b_empty = [&](int s) { return bar0 + 8u * (12 + s); };      // [4]
  auto tfull_bar = [&](int s) { return bar0 + 8u * (16 + s); };
  auto tempty_bar = [&](int s) { return bar0 + 8u * (18 + s); };
  const uint32_t tmem_ptr_smem = bar0 + 8u * 20;
  const uint32_t done_bar = bar0 + 8u * 21;  // final tcgen05.commit: no async mbarrier arrive may outlive the CTA
  auto stat_full = [&](int s) { return bar0 + 8u * (38 + s); };  // [2] statistics warp -> producers
  auto stat_empty = [&](int s) { return bar0 + 8u * (40 + s); }; // [2] producers (one arrival per warp) -> statistics warp
  auto red_full = [&](int s) { return bar0 + 8u * (42 + s); };   // [2] producers' per-warp gLN2 partials -> statistics warp

  // PDL: everything up to the prologue barrier touches only constants (weights, tap tables, per-call layout
  // tables), so it may overlap the previous kernel's tail; activations are read after pdl_wait().
  pdl_launch_dependents();
  if (threadIdx.x == 0) tl_stamp_ns(112);
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int n_tiles = tb.n_tiles;
  // 2-CTA clusters: the pair (2p, 2p+1) visits the tile pairs p, p + n_pairs, ...; rank r takes tile 2*pair_tile + r.
  // Both CTAs of a pair make the same number of visits (each visit consumes 8 weight stages that the two CTAs load
  // for each other); with an odd tile count the last visit of rank 1 re-processes the last tile without writing.
  const int n_pairs = (int)gridDim.x, pair = (int)blockIdx.x;
  const int n_pair_tiles = n_tiles;
  const int n_visits = (pair < n_pair_tiles) ? (n_pair_tiles - pair + n_pairs - 1) / n_pairs : 0;
  auto tile_at = [&](int it, bool& valid) {
    // K_B walks the m-tiles upwards, K_A downwards: K_A starts on the h1 tiles K_B wrote LAST (still in the 126 MB
    // L2) and finishes on the tiles the next K_B reads FIRST (their G_res lines are the freshest)
    const int wr = n_tiles - 1 - (pair + it * n_pairs);
    valid = wr < n_tiles;
    return valid ? wr : n_tiles - 1;
  };
  const int d = (D > 0) ? D : p.dilation;
  const int hw = (D > 0) ? 8 : d;  // halo width in frames (multiple of 8, <= 128)
  const uint32_t halo_bytes = (uint32_t)(64 * hw * 2);
  const uint32_t raw_stage_bytes = (16384u + 2u * halo_bytes + 1023u) & ~1023u;
  // raw ring: k-block kb lives in stage kb % raw_stages; with an EVEN ring every stage and its mbarriers belong to ONE
  // producer group (kb & 1), so a parity wait is never more than one phase ahead.  Four stages (two per group: the next
  // k-block's L2 -> shared-memory latency, ~2.3 k cycles, hides behind the current one's arithmetic) when they fit next to
  // two weight stages, else two; the rest of the shared memory is the weight ring.
  const int raw_stages = ((uint32_t)Cfg::OFF_RAW + 4u * raw_stage_bytes + 2u * (uint32_t)Cfg::B_STAGE <=
                                                        (uint32_t)Cfg::SMEM_USABLE) ? 4 : 2;
  const uint32_t off_b = (uint32_t)Cfg::OFF_RAW + (uint32_t)raw_stages * raw_stage_bytes;
  const int b_stages = min((int)(((uint32_t)Cfg::SMEM_USABLE - off_b) / (uint32_t)Cfg::B_STAGE), Cfg::B_MAX_STAGES);
  auto stage_b = [&](int s) { return smem_base + off_b + (uint32_t)(s * Cfg::B_STAGE); };
  auto raw_st = [&](int s) { return smem_base + (uint32_t)Cfg::OFF_RAW + (uint32_t)s * raw_stage_bytes; };

  // static tables -> shared memory
  for (int i = threadIdx.x; i < C_HID; i += Cfg::THREADS) s_t1[i] = __ldg(p.t1 + i);
  if ((int)threadIdx.x < TI_CACHE && (int)threadIdx.x < n_visits) {
    bool v_;
    s_ti[threadIdx.x] = load_tile_info(tb.tile_info + tile_at(threadIdx.x, v_));
  }
  if (threadIdx.x < 2) s_cnt[threadIdx.x] = 0;
  if (warp == 4 && lane == 0) { tma_prefetch_desc(&tm_hc); tma_prefetch_desc(&tm_hh); }
  if (warp == 5 && lane == 0) tma_prefetch_desc(&tm_b);
  if (warp == 6) {
    if (lane == 0) {
      for (int s = 0; s < 4; ++s) {
        mbar_init(raw_full(s), 1);
        mbar_init(raw_empty(s), Cfg::PROD_WARPS / 2);   // one arrival per warp of the owning group
      }
      for (int s = 0; s < 2; ++s) {
        // pair mode: the producers / epilogue threads of BOTH CTAs arrive on the leader's a_full / tempty
        mbar_init(tfull_bar(s), 1);
        mbar_init(tempty_bar(s), Cfg::EPI_WARPS);
      }
      for (int s = 0; s < 2 * Cfg::A_PG; ++s) {
        mbar_init(a_full(s), Cfg::PROD_WARPS / 2);
        mbar_init(a_empty(s), 1);
      }
      for (int s = 0; s < Cfg::B_MAX_STAGES; ++s) {
        mbar_init(b_full(s), 1);
        mbar_init(b_empty(s), 1);
      }
      mbar_init(done_bar, 1);
      for (int s = 0; s < 2; ++s) {
        mbar_init(stat_full(s), 1);
        mbar_init(stat_empty(s), Cfg::PROD_WARPS);
        mbar_init(red_full(s), Cfg::PROD_WARPS);
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_ptr_smem, Cfg::TMEM_COLS);
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  pdl_wait();
  if (threadIdx.x == 0) tl_stamp_ns(113);
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_smem));

  if (warp < 4) {
    // ================================ epilogue: TMEM -> fp16 G ====================================
    const int quad = warp;
    const int r = quad * 32 + lane;
    constexpr int NCH = BN / 32;
    for (int it = 0; it < n_visits; ++it) {
      bool valid;
      const int w = tile_at(it, valid);
      const int as = it & 1;
      const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
      mbar_wait_b(tb.flags, tfull_bar(as), aphase);
      tcgen05_fence_after();
      if (threadIdx.x == 0) tl_stamp(it, 80);
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * BN);
      __half* __restrict__ d_res = g_res + tt_index(w, C_BN, 0, r);
      __half* __restrict__ d_skip = g_skip + (size_t)w * skip_tile_stride + r;  // stash: [tile][block][128][128]
      float v[2][32];
      tmem_ld32(taddr, v[0]);
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        tmem_ld_wait();
        if (c + 1 < NCH) tmem_ld32(taddr + (uint32_t)((c + 1) * 32), v[(c + 1) & 1]);
        // BN == 256: columns 0-127 = residual conv, 128-255 = skip conv; BN == 128 (last block): skip only
        __half* __restrict__ dst = (BN == 256 && c < 4) ? d_res + (size_t)(c * 32) * TILE_M
                                                        : d_skip + (size_t)((BN == 256 ? c - 4 : c) * 32) * TILE_M;
        if (valid)
#pragma unroll
        for (int j = 0; j < 32; j += 2) {  // one packed conversion per two channels
          const __half2 h = __floats2half2_rn(v[c & 1][j], v[c & 1][j + 1]);
          dst[(size_t)j * TILE_M] = __low2half(h);
          dst[(size_t)(j + 1) * TILE_M] = __high2half(h);
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(as));
      if (threadIdx.x == 0) tl_stamp(it, 81);
    }
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
    // ================================ TMA producer: weights =======================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < n_visits; ++it) {
        for (int kb = 0; kb < 8; ++kb) {
          mbar_wait_b(tb.flags, b_empty(stage), phase ^ 1u);  // BOTH CTAs' MMAs are done with this stage
          // the whole stage lands here (own half + the peer's half); this CTA loads rows [crank * BN/2, +BN/2) of the
          // k-block and multicasts them to the same offset of both CTAs
          mbar_expect_tx(b_full(stage), Cfg::B_STAGE);
          tma_load_2d(stage_b(stage), &tm_b, b_full(stage), kb * 64, 0);
          if (++stage == b_stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 6) {
    // ================================ MMA issuer ==================================================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BN);
      int bs = 0;
      uint32_t bphase = 0;
      for (int it = 0; it < n_visits; ++it) {
        const int as = it & 1;
        const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
        mbar_wait_b(tb.flags, tempty_bar(as), aphase ^ 1u);
        tcgen05_fence_after();
        tl_stamp(it, 64);
        const uint32_t tmem_d = tmem_base + (uint32_t)(as * BN);
        for (int kb = 0; kb < 8; ++kb) {
          const int nseq = it * 4 + (kb >> 1);  // k-blocks this group has produced before this one
          const int g = (kb & 1) * Cfg::A_PG + (nseq & (Cfg::A_PG - 1));
          const uint32_t gphase = (uint32_t)(nseq / Cfg::A_PG) & 1u;
          mbar_wait_b(tb.flags, a_full(g), gphase);
          mbar_wait_b(tb.flags, b_full(bs), bphase);
          tcgen05_fence_after();
          tl_stamp(it, 65 + kb);
          const uint32_t sa = stage_a(g), sb = stage_b(bs);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint64_t da = make_smem_desc(sa + j * 2048u, 8192u, 1024u);
            const uint64_t db = make_smem_desc(sb + j * 32u, 16u, 1024u);
            umma_f16(tmem_d, da, db, idesc, (uint32_t)((kb | j) != 0));
          }
          umma_commit(a_empty(g));
          umma_commit(b_empty(bs));
          if (kb == 7) umma_commit(tfull_bar(as));
          if (++bs == b_stages) { bs = 0; bphase ^= 1u; }
        }
      }
      // the last ab_empty arrivals have no waiter: make sure every asynchronous arrive has landed before exit
      umma_commit(done_bar);
      mbar_wait_b(tb.flags, done_bar, 0);
    }
    __syncwarp();
  } else if (warp == Cfg::W_STAT) {
    // ================================ statistics warp ==============================================
    // gLN1 (mean, rstd) of every visit's segment (fp64 reduction of the per-tile partials K_B left behind) folded into
    // the visit's per-channel tap table, one visit ahead of the producers; and the gLN2 partial of the previous visit
    // (the 16 producer warps' (sum, sumsq) pairs added in warp order: a fixed order, no producer warp stays behind).
    auto reduce_visit = [&](int it) {
      bool valid;
      const int w = tile_at(it, valid);
      mbar_wait_b(tb.flags, red_full(it & 1), ((uint32_t)it >> 1) & 1u);
      const float2 v = (lane < Cfg::PROD_WARPS) ? s_red[(it & 1) * Cfg::PROD_WARPS + lane] : make_float2(0.f, 0.f);
      float ts = 0.f, tq = 0.f;
#pragma unroll
      for (int k = 0; k < Cfg::PROD_WARPS; ++k) {
        ts += __shfl_sync(0xffffffffu, v.x, k);
        tq += __shfl_sync(0xffffffffu, v.y, k);
      }
      if (lane == 0 && valid) part2[w] = make_float2(ts, tq);
    };
    for (int it = 0; it < n_visits; ++it) {
      bool v_;
      const TileInfo ti = ti_get(s_ti, tb, it, tile_at(it, v_));
      const float2 pre = part_prefetch(part1, ti);
      // (the constants of the first 128 channels are requested before the statistics, those of the next 128 before this
      // group's table entries are computed: four back-to-back load latencies behind the fp64 reduction were ~2 k cycles
      // in front of the FIRST k-block of a launch, where nothing hides them)
      float4 t0[2][4];
      float cb[2][4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        t0[0][j] = __ldg(p.t0 + j * 32 + lane);
        cb[0][j] = __ldg(p.cb + j * 32 + lane);
      }
      float mean, rstd;
      seg_stats_pre(part1, ti, pre, p.eps1, mean, rstd, tb.flags);
      const float rm = rstd * mean;
      mbar_wait_b(tb.flags, stat_empty(it & 1), (((uint32_t)it >> 1) & 1u) ^ 1u);
      uint2* tap = s_tap + (it & 1) * C_HID;
#pragma unroll
      for (int g = 0; g < C_HID / 128; ++g) {
        const int c0 = g * 128;
        if (g + 1 < C_HID / 128) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            t0[(g + 1) & 1][j] = __ldg(p.t0 + c0 + 128 + j * 32 + lane);
            cb[(g + 1) & 1][j] = __ldg(p.cb + c0 + 128 + j * 32 + lane);
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 t = t0[g & 1][j];
          tap[c0 + j * 32 + lane] = make_uint2(h2_as_u32(__floats2half2_rn(t.x * rstd, t.y * rstd)),
                                               h2_as_u32(__floats2half2_rn(t.z * rstd, fmaf(-t.w, rm, cb[g & 1][j]))));
        }
      }
      if (lane == 0) s_stat[it & 1] = make_float2(mean, rstd);
      __syncwarp();
      if (lane == 0) mbar_arrive(stat_full(it & 1));  // (release at CTA scope orders the warp's stores before the arrival)
      if (it > 0) reduce_visit(it - 1);
    }
    if (n_visits > 0) reduce_visit(n_visits - 1);
  } else {
    // ================================ depthwise producers =========================================
    // Two groups of 8 warps; group g produces the k-blocks kb = g, g+2, ... (and owns A/B stage g), so
    // the groups run out of phase: while one waits on its barriers the other keeps the issue slots busy.
    // A warp-step = 2 channels x 128 frames; a group covers a k-block in 4 steps per warp.
    const int pw = warp - 7;              // 0..15
    const int grp = pw >> 3, gw = pw & 7;
    const int sub = lane >> 4, lane16 = lane & 15;
    const int row0 = gw * 2 + sub;        // k-row of step 0 inside a k-block; step j = row0 + 16 j
    // shared-memory offsets (inside a raw stage) of this thread's centre / left / right vectors
    const int sv = (D > 0) ? 1 : (d >> 3);  // tap distance in 8-frame vectors
    const int hv = hw >> 3;                 // halo width in vectors
    const uint32_t c_pitch = 256u;
    uint32_t l_off, l_pitch, r_off, r_pitch;
    if (lane16 - sv >= 0) { l_off = (uint32_t)((lane16 - sv) * 16); l_pitch = 256u; }
    else { l_off = 16384u + (uint32_t)((lane16 - sv + hv) * 16); l_pitch = (uint32_t)(hw * 2); }
    if (lane16 + sv < 16) { r_off = (uint32_t)((lane16 + sv) * 16); r_pitch = 256u; }
    else { r_off = 16384u + halo_bytes + (uint32_t)((lane16 + sv - 16) * 16); r_pitch = (uint32_t)(hw * 2); }
    // step 0 addresses of this thread (step j = + 16 j rows)
    const uint32_t c_base = (uint32_t)(lane16 * 16) + (uint32_t)row0 * c_pitch;
    const uint32_t l_base = l_off + (uint32_t)row0 * l_pitch, r_base = r_off + (uint32_t)row0 * r_pitch;
    const uint32_t c_step = 16u * c_pitch, l_step = 16u * l_pitch, r_step = 16u * r_pitch;
    const uint32_t a_off0 = a_chunk_off(row0, lane16 * 8);  // (row & 7) is the same for every step; step j = + 2048 j
    int nseq = 0;  // k-blocks produced by this group so far -> A stage / phase
    const __half2 alpha_h2 = __float2half2_rn(p.alpha2);
    const bool tl_on = (gw == 0 && lane == 0);
    const int tl_base = grp ? 100 : 40;  // timeline slots: group 0 40.., group 1 100..

    const int rpg = raw_stages >> 1;  // raw stages per group
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
      for (int kb = grp; kb < 8; kb += 2) {
        const int a_idx = grp * Cfg::A_PG + (nseq & (Cfg::A_PG - 1));
        const uint32_t aphase = (uint32_t)(nseq / Cfg::A_PG) & 1u;
        const uint32_t sa = stage_a(a_idx) + a_off0;
        const uint32_t my_ab_full = a_full(a_idx);
        const uint32_t my_ab_empty = a_empty(a_idx);
        const int rs_ = grp + 2 * (nseq & (rpg - 1));             // raw stage of this k-block (= kb % raw_stages)
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
            const uint2 ta = tap[ch0 + jj * 16], tb_ = tap[ch0 + jj * 16 + 16];
            // ... then the arithmetic; the (single) A stage of this group is only needed for the stores, so the
            // tensor core drains the previous k-block of the group while the first two steps are computed
            __half2 sh = __float2half2_rn(0.f), qh = __float2half2_rn(0.f);
            const uint4 o0 = dw_step_h2p<D, A01>(vc0, va0, vb0, ta, alpha_h2, sh, qh);
            const uint4 o1 = dw_step_h2p<D, A01>(vc1, va1, vb1, tb_, alpha_h2, sh, qh);
            dw_stats_flush(sh, qh, s2, q2);
            if (jj == 0) {
              mbar_wait_b(tb.flags, my_ab_empty, aphase ^ 1u);
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
            const uint2 ta = tap[ch0 + jj * 16], tb_ = tap[ch0 + jj * 16 + 16];
            const float4 ua = s_t1[ch0 + jj * 16], ub = s_t1[ch0 + jj * 16 + 16];
            __half2 sh = __float2half2_rn(0.f), qh = __float2half2_rn(0.f);
            const uint4 o0 = dw_step_h2p_edge<D, A01>(vc0, va0, vb0, ta, ua, mean, alpha_h2, mk, sh, qh);
            const uint4 o1 = dw_step_h2p_edge<D, A01>(vc1, va1, vb1, tb_, ub, mean, alpha_h2, mk, sh, qh);
            dw_stats_flush(sh, qh, s2, q2);
            if (jj == 0) mbar_wait_b(tb.flags, my_ab_empty, aphase ^ 1u);
            st_shared_v4(sa + jj * 2048u, o0);
            st_shared_v4(sa + (jj + 1) * 2048u, o1);
          }
        }
        fence_proxy_async();
        __syncwarp();  // every lane has fenced its stage writes and finished its reads of the raw stage
        if (lane == 0) {
          mbar_arrive(my_ab_full);
          mbar_arrive(raw_empty(rs_));
        }
        if (tl_on) tl_stamp(it, tl_base + 3 * (kb >> 1) + 2);
        ++nseq;
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
  if (threadIdx.x == 0) tl_stamp_ns(114);
  if (warp == 6) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// =================================================================================================
// K_B: apply previous block's G_res (x += res) + in-conv GEMM -> h1
// =================================================================================================
struct KbCfg {
  static constexpr int W_BYTES = C_HID * C_BN * 2;      // 128 KiB: [2 k-blocks][512 rows][64 k]
  static constexpr int XN_BYTES = TILE_M * C_BN * 2;    // 32 KiB: [2 k-blocks][64 k][128 m]
  static constexpr int EPI_WARPS = 8;   // two per TMEM lane quadrant
  static constexpr int PRO_WARPS = 8;
  static constexpr int PRO_RND = 4;     // prologue passes per round (global loads in flight per thread: 2 x PRO_RND vectors)
  static constexpr int THREADS = (EPI_WARPS + 1 + PRO_WARPS) * 32;  // 544
  static constexpr int TMEM_COLS = 512;                 // 4 chunks of 128 output channels
  static constexpr int ST_BYTES = TILE_M * 128 * 2;  // h1 staging
  static constexpr int OFF_XN = W_BYTES;
  static constexpr int OFF_ST = OFF_XN + 2 * XN_BYTES;
  static constexpr int OFF_TI = OFF_ST + ST_BYTES;
  static constexpr int OFF_RED = OFF_TI + TI_CACHE * 32;
  static constexpr int OFF_CNT = OFF_RED + 2 * EPI_WARPS * 8;
  static constexpr int OFF_BAR = OFF_CNT + 16;
  static constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;
  static_assert(SMEM_BYTES <= 232448, "K_B shared memory over the 227 KiB limit");
};

struct ApplyParams {
  const __half* g;      // [nt][128][128] raw residual-conv output of the previous block (nullptr: block 0)
  const float2* part2;  // [nt] gLN2 partials of the previous block
  const float* rs;      // [>=128] rowsum(W_res . gamma2) of the previous block
  const float* cb;      // [>=128] W_res * beta2 + b_res
  float* x;             // (unused by the fused pipeline: the unfused scaffold keeps an fp32 stream)
  __half* x16;          // [nt][128][128] fp16 residual stream (updated in place; every update is computed in fp32)
  float eps2;           // eps of the previous block's gLN2 (shared_block.5)
};

template <bool A01>
__global__ void __launch_bounds__(KbCfg::THREADS, 1)
apply_in_kernel(const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ CUtensorMap tm_h1, ApplyParams ap,
                EpiIn<__half, A01> epi, Tables tb) {
  using Cfg = KbCfg;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t smem_w = smem_base;
  auto smem_xn = [&](int b) { return smem_base + (uint32_t)(Cfg::OFF_XN + b * Cfg::XN_BYTES); };
  TileInfo* s_ti = reinterpret_cast<TileInfo*>(smem_gen + Cfg::OFF_TI);
  float2* s_red = reinterpret_cast<float2*>(smem_gen + Cfg::OFF_RED);  // [2][EPI_WARPS]
  int* s_cnt = reinterpret_cast<int*>(smem_gen + Cfg::OFF_CNT);          // [2]
  const uint32_t bar0 = smem_base + Cfg::OFF_BAR;
  const uint32_t w_bar = bar0;
  auto xn_full = [&](int b) { return bar0 + 8u * (1 + b); };
  auto xn_empty = [&](int b) { return bar0 + 8u * (3 + b); };
  auto tfull_bar = [&](int q) { return bar0 + 8u * (5 + q); };
  auto tempty_bar = [&](int q) { return bar0 + 8u * (9 + q); };
  const uint32_t tmem_ptr_smem = bar0 + 8u * 13;
  const uint32_t done_bar = bar0 + 8u * 14;  // final tcgen05.commit: no async mbarrier arrive may outlive the CTA

  pdl_launch_dependents();  // PDL: the prologue below (barriers, TMEM, resident W_in, layout tables) reads constants only
  if (threadIdx.x == 0) tl_stamp_ns(24);
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int n_tiles = tb.n_tiles;

  constexpr int MMA_WARP = Cfg::EPI_WARPS;
  ti_cache_fill(s_ti, tb, threadIdx.x);
  if (threadIdx.x < 2) s_cnt[threadIdx.x] = 0;
  if (warp == MMA_WARP) {
    if (lane == 0) {
      tma_prefetch_desc(&tm_w);
      tma_prefetch_desc(&tm_h1);
      mbar_init(w_bar, 1);
      for (int b = 0; b < 2; ++b) {
        mbar_init(xn_full(b), Cfg::PRO_WARPS);           // one arrival per warp
        mbar_init(xn_empty(b), 1);
      }
      for (int q = 0; q < 4; ++q) {
        mbar_init(tfull_bar(q), 1);
        mbar_init(tempty_bar(q), Cfg::EPI_WARPS);
      }
      mbar_init(done_bar, 1);
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
  pdl_wait();
  if (threadIdx.x == 0) tl_stamp_ns(25);
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_smem));

  if (warp < MMA_WARP) {
    // ================================ epilogue: bias + PReLU -> h1, gLN1 partials ==================
    // D[channel][frame]: TMEM lane = output channel of the 128-channel chunk, column = frame.  A thread owns 32
    // consecutive frames of one channel per tcgen05.ld: the bias is a per-lane constant, eight frames pack into one
    // 16-byte st.shared (swizzled like a TMA box, conflict-free), and h1 leaves the SM as 64-frame x 32-channel TMA tile stores --
    // 1/8 of the LSU instructions of the lane = frame mapping (which needs one 2-byte store per value).
    static_assert(Cfg::EPI_WARPS == 8, "epilogue mapping assumes two warps per TMEM lane quadrant");
    const int ew = warp, quad = ew & 3, half = ew >> 2;   // half = frame half of the m-tile
    const int cl = quad * 32 + lane;                       // channel inside the chunk
    const int row = lane;                                  // row inside the warp's staging box
    // every epilogue warp stages and stores its own 32 channels x 64 frames (a 4 KiB box per chunk): no barrier between
    // warps, only __syncwarp (two warps sharing a 64 x 64 box met at two named barriers per chunk: 56.0 vs 55.4 us)
    const uint32_t st_box = smem_base + (uint32_t)(Cfg::OFF_ST + ew * 4096);
    const uint32_t st_row = st_box + (uint32_t)(lane * 128);
    const bool issuer = (lane == 0);                       // one thread per box issues its TMA store
    const float alpha = epi.alpha;
    int it = 0;
    for (int w = blockIdx.x; w < n_tiles; w += gridDim.x, ++it) {
      const TileInfo ti = ti_get(s_ti, tb, it, w);
      const int nvalid = ti.F - ti.t0;                     // frames >= nvalid lie beyond the segment end
      const bool full = nvalid >= TILE_M;
      float2 s2 = make_float2(0.f, 0.f), q2 = make_float2(0.f, 0.f);
      const uint32_t ph = (uint32_t)it & 1u;
      float v[2][32];
      mbar_wait_b(tb.flags, tfull_bar(0), ph);
      tcgen05_fence_after();
      if (threadIdx.x == 0) tl_stamp(it, 16);
      tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(half * 64), v[0]);
      float bch = __ldg(epi.bias + cl);
#pragma unroll
      for (int pi = 0; pi < 8; ++pi) {
        const int q = pi >> 1, hh = pi & 1;
        tmem_ld_wait();  // piece pi (32 frames) is in v[pi & 1]
        if (hh == 1) {
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty_bar(q));  // the warp's part of chunk q is in registers: the MMA may overwrite it
        }
        float bnext = bch;
        if (pi + 1 < 8) {
          const int qn = (pi + 1) >> 1;
          if (((pi + 1) & 1) == 0) {
            mbar_wait_b(tb.flags, tfull_bar(qn), ph);
            tcgen05_fence_after();
            if (threadIdx.x == 0) tl_stamp(it, 16 + qn);
            bnext = __ldg(epi.bias + qn * 128 + cl);
          }
          tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(qn * 128 + half * 64 + ((pi + 1) & 1) * 32),
                    v[(pi + 1) & 1]);
        }
        const float2 b2 = make_float2(bch, bch);
        const int f0 = half * 64 + hh * 32;  // first frame of the piece inside the m-tile
        uint32_t pk[16];
        // two copies of the piece arithmetic behind one uniform branch: the per-element frame masks of a segment's last
        // tile cost ~60 of the ~270 instructions of a piece when they are if-converted into the common path
        // two copies of the piece arithmetic behind one uniform branch: the per-element frame masks of a segment's last
        // tile cost ~60 of the ~270 instructions of a piece when they are if-converted into the common path.  Everything
        // in fp32: h1 = W_in x of the UN-normalised residual stream is only O(1) as far as the load-time conditioning's
        // white model of x holds, and a packed-half first level of its sum of squares (what K_A uses for h2, the image of
        // a NORMALISED tensor) overflows beyond |h1| ~ 90 -- measured 0.6 us per launch faster, not kept.
        auto piece_math = [&](auto full_c) {
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            float2 u = add2(make_float2(v[pi & 1][j], v[pi & 1][j + 1]), b2);
            u = prelu2<A01>(u, alpha);
            pk[j >> 1] = h2_as_u32(__floats2half2_rn(u.x, u.y));
            if constexpr (!decltype(full_c)::value) {  // statistics exclude frames beyond the segment end
              if (f0 + j >= nvalid) u.x = 0.f;
              if (f0 + j + 1 >= nvalid) u.y = 0.f;
            }
            s2 = add2(s2, u);
            q2 = fma2(u, u, q2);
          }
        };
        if (full) piece_math(std::true_type{});
        else piece_math(std::false_type{});
        if (hh == 0) {
          // the previous staged chunk's TMA store must have finished READING this staging box before it is rewritten: waited
          // for AFTER this piece's arithmetic (the store was issued only a few hundred cycles ago, at the end of the
          // previous piece; in front of the arithmetic the issuer -- and its partner warp at the barrier -- sat out the TMA
          // unit's latency once per chunk)
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
            tma_store_2d(&tm_h1, st_box, half * 64, w * C_HID + q * 128 + quad * 32);
            bulk_commit();
          }
          if (threadIdx.x == 0) tl_stamp(it, 20 + q);
        }
        bch = bnext;
      }
      const float s = warp_sum(s2.x + s2.y), qq = warp_sum(q2.x + q2.y);
      tile_partial_reduce(s_red + (it & 1) * Cfg::EPI_WARPS, s_cnt + (it & 1), Cfg::EPI_WARPS, ew, s, qq,
                          epi.part_out + w);
    }
    if (issuer) bulk_wait0();  // all h1 tile stores complete before the CTA exits
  } else if (warp == MMA_WARP) {
    // ================================ MMA issuer ==================================================
    if (lane == 0) {
      // A = W_in chunk (K-major), B = x operand (MN-major): bit 15 (A major) = 0, bit 16 (B major) = 1
      constexpr uint32_t idesc = (1u << 4) | (1u << 16) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
      mbar_wait_b(tb.flags, w_bar, 0);
      int it = 0;
      for (int w = blockIdx.x; w < n_tiles; w += gridDim.x, ++it) {
        const int b = it & 1;
        mbar_wait_b(tb.flags, xn_full(b), (uint32_t)(it >> 1) & 1u);
        tcgen05_fence_after();
        tl_stamp(it, 8);
        const uint32_t sa = smem_xn(b);
        for (int q = 0; q < 4; ++q) {
          mbar_wait_b(tb.flags, tempty_bar(q), ((uint32_t)it & 1u) ^ 1u);
          tcgen05_fence_after();
          tl_stamp(it, 9 + q);
          const uint32_t tmem_d = tmem_base + (uint32_t)(q * 128);
#pragma unroll
          for (int kb = 0; kb < 2; ++kb) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint64_t dx = make_smem_desc(sa + (uint32_t)(kb * 16384 + j * 2048), 8192u, 1024u);  // MN-major
              const uint64_t dw = make_smem_desc(smem_w + (uint32_t)(kb * 65536 + q * 16384 + j * 32), 16u, 1024u);  // K-major
              umma_f16(tmem_d, dw, dx, idesc, (uint32_t)((kb | j) != 0));  // D[channel][frame] = W . x^T
            }
          }
          umma_commit(tfull_bar(q));
        }
        umma_commit(xn_empty(b));
      }
      umma_commit(done_bar);  // the last xn_empty arrivals have no waiter
      mbar_wait_b(tb.flags, done_bar, 0);
    }
    __syncwarp();
  } else {
    // ================================ prologue: apply G_res, build the fp16 operand ===============
    const int pt = (warp - MMA_WARP - 1) * 32 + lane;  // 0..PRO_WARPS*32-1
    constexpr int ROWS = Cfg::PRO_WARPS * 2;  // channel rows covered per pass
    const int lane16 = pt & 15, crow = pt >> 4;
    const int f0 = lane16 * 8;
    const bool has_g = ap.g != nullptr;
    int it = 0;
    float2 pre = make_float2(0.f, 0.f);
    if (has_g && (int)blockIdx.x < n_tiles) pre = part_prefetch(ap.part2, ti_get(s_ti, tb, 0, blockIdx.x));
    for (int w = blockIdx.x; w < n_tiles; w += gridDim.x, ++it) {
      const int b = it & 1;
      const TileInfo ti = ti_get(s_ti, tb, it, w);
      __half* __restrict__ xt = ap.x16 + tt_index(w, C_BN, 0, f0);
      const __half* __restrict__ gt = has_g ? ap.g + tt_index(w, C_BN, 0, f0) : nullptr;
      constexpr int PASSES = C_BN / ROWS;
      constexpr int RND = Cfg::PRO_RND;
      static_assert(PASSES % RND == 0, "prologue rounds");
      uint4 xh[RND];
      uint4 gr[RND];
      float rsv[RND], cbv[RND];   // per-channel constants of the round: requested with the vectors, not at their first use
      auto load_round = [&](int p0) {
#pragma unroll
        for (int j = 0; j < RND; ++j) {
          const size_t off = (size_t)((p0 + j) * ROWS + crow) * TILE_M;
          xh[j] = *reinterpret_cast<const uint4*>(xt + off);
          if (has_g) {
            gr[j] = ldg_v4(gt + off);
            rsv[j] = __ldg(ap.rs + (p0 + j) * ROWS + crow);
            cbv[j] = __ldg(ap.cb + (p0 + j) * ROWS + crow);
          }
        }
      };
      // the first round's global loads are in flight while the segment statistics are reduced (fp64, ~500 cycles), all
      // of them are issued before the (possibly blocking) operand-buffer wait
      load_round(0);
      float mean = 0.f, rstd = 0.f;
      if (has_g) {
        seg_stats_pre(ap.part2, ti, pre, ap.eps2, mean, rstd, tb.flags);
        if (w + (int)gridDim.x < n_tiles) pre = part_prefetch(ap.part2, ti_get(s_ti, tb, it + 1, w + gridDim.x));
      }
      float2 vm[4];
#pragma unroll
      for (int k = 0; k < 4; ++k)
        vm[k] = make_float2((ti.t0 + f0 + 2 * k) < ti.F ? 1.f : 0.f, (ti.t0 + f0 + 2 * k + 1) < ti.F ? 1.f : 0.f);
      // rstd2 rounded to fp16: multiplier of the mixed-precision FMA on the fp16 G_res (the constant term uses the
      // same rounded value)
      const __half rstd_hh = __float2half_rn(rstd);
      const uint16_t rstd_h = __half_as_ushort(rstd_hh);
      const float nmr = -mean * __half2float(rstd_hh);
      if (pt == 0) tl_stamp(it, 0);
      if (pt == 0 && w + (int)gridDim.x < n_tiles) {  // next tile's x / G_res -> L2 while this tile is processed
        const size_t nidx = tt_index(w + gridDim.x, C_BN, 0, 0);
        l2_prefetch_bulk(ap.x16 + nidx, TILE_M * C_BN * 2);
        if (has_g) l2_prefetch_bulk(ap.g + nidx, TILE_M * C_BN * 2);
      }
      uint32_t xn = 0;
      auto compute_round = [&](int p0) {
#pragma unroll
        for (int j = 0; j < RND; ++j) {
          const int c = (p0 + j) * ROWS + crow;
          const size_t off = (size_t)c * TILE_M;
          if (has_g) {
            float xv[8];
            const uint32_t xw[4] = {xh[j].x, xh[j].y, xh[j].z, xh[j].w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float2 f = __half22float2(u32_as_h2(xw[i]));
              xv[2 * i] = f.x; xv[2 * i + 1] = f.y;
            }
            const float kr = fmaf(nmr, rsv[j], cbv[j]);
            const uint32_t gw[4] = {gr[j].x, gr[j].y, gr[j].z, gr[j].w};
#pragma unroll
            for (int i = 0; i < 8; i += 2) {
              const float2 dr = make_float2(fma_f32_f16((uint16_t)(gw[i >> 1] & 0xffffu), rstd_h, kr),
                                            fma_f32_f16((uint16_t)(gw[i >> 1] >> 16), rstd_h, kr));
              const float2 nx = fma2(dr, vm[i >> 1], make_float2(xv[i], xv[i + 1]));
              xv[i] = nx.x; xv[i + 1] = nx.y;
            }
            const uint4 pk = pack8(xv);  // the rounded stream IS the GEMM operand: one rounding, stored twice
            *reinterpret_cast<uint4*>(xt + off) = pk;
            st_shared_v4(xn + (uint32_t)((c >> 6) * 16384) + a_chunk_off(c & 63, f0), pk);
          } else {
            st_shared_v4(xn + (uint32_t)((c >> 6) * 16384) + a_chunk_off(c & 63, f0), xh[j]);  // block 0: x as stored
          }
        }
      };
      mbar_wait_b(tb.flags, xn_empty(b), ((uint32_t)(it >> 1) & 1u) ^ 1u);
      xn = smem_xn(b);
      if (pt == 0) tl_stamp(it, 1);
      compute_round(0);
#pragma unroll 1
      for (int p0 = RND; p0 < PASSES; p0 += RND) {
        load_round(p0);
        compute_round(p0);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(xn_full(b));
      if (pt == 0) tl_stamp(it, 2);
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) tl_stamp_ns(26);
  if (warp == MMA_WARP) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// =================================================================================================
// After the last block: skip_sum = sum over blocks of the gLN2-scaled raw skip outputs, then PReLU
// (mask_net.0) -> fp16 operand of the mask GEMM.  grid = n_tiles, 256 threads: thread = (channel row,
// 8 frames).  stash[b] = raw G_skip of block b, part2[b] its gLN2 partials, rs/cb [n_layers][128].
// =================================================================================================
constexpr int MAX_LAYERS = 64;
__global__ void __launch_bounds__(256) skip_merge_kernel(const __half* __restrict__ stash,
                                                         const float2* __restrict__ part2, size_t part_stride,
                                                         const float* __restrict__ rs, const float* __restrict__ cb,
                                                         const float* __restrict__ eps2, __half* __restrict__ skipa, float alpha_mask, int n_layers,
                                                         Tables tb) {
  __shared__ float2 s_stat[MAX_LAYERS];    // (rstd, -mean*rstd) per block
  __shared__ uint16_t s_rstd_h[MAX_LAYERS];  // rstd rounded to fp16: multiplier of the mixed-precision FMA
  pdl_enter();
  const int w = blockIdx.x;
  const TileInfo ti = load_tile_info(tb.tile_info + w);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int b = warp; b < n_layers; b += 8) {
    float mean, rstd;
    warp_seg_stats(part2 + (size_t)b * part_stride, ti.tile0, ti.ntiles, 1, ti.inv_count, __ldg(eps2 + b), mean, rstd, tb.flags);
    if (lane == 0) {
      // the constant term uses the SAME rounded rstd, so sum_b rstd_b * (G_b - mean_b * rs_b) stays consistent
      const __half rh = __float2half_rn(rstd);
      s_rstd_h[b] = __half_as_ushort(rh);
      const float rr = __half2float(rh);
      s_stat[b] = make_float2(rr, -mean * rr);
    }
  }
  __syncthreads();
  const int lane16 = threadIdx.x & 15, crow = threadIdx.x >> 4;
  const int f0 = lane16 * 8;
#pragma unroll 1
  // grid = (n_tiles, SKIP_MERGE_SPLIT): a CTA covers 128 / SKIP_MERGE_SPLIT channel rows of its m-tile, so that the launch
  // is many short CTAs (1050 whole-tile CTAs on 148 x 5 resident slots were 1.4 waves: the second wave ran at 42 %)
  const int ps_per = 8 / (int)gridDim.y;
  for (int ps = (int)blockIdx.y * ps_per; ps < ((int)blockIdx.y + 1) * ps_per; ++ps) {
    const int c = ps * 16 + crow;
    const size_t idx = tt_index(w, C_SKIP, c, f0);
    // stash layout [tile][block][128][128]: the n_layers slots of an m-tile are contiguous
    const __half* __restrict__ sp = stash + ((size_t)w * n_layers * C_SKIP + c) * TILE_M + f0;
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    float kc = 0.f;
#pragma unroll 8
    for (int b = 0; b < n_layers; ++b) {
      const uint4 gv4 = ldg_v4(sp + (size_t)b * (C_SKIP * TILE_M));
      const float2 sb = s_stat[b];
      const uint16_t rh = s_rstd_h[b];
      kc += fmaf(sb.y, __ldg(rs + b * C_SKIP + c), __ldg(cb + b * C_SKIP + c));
      const uint32_t gw[4] = {gv4.x, gv4.y, gv4.z, gv4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {  // fp16 x fp16 + fp32 (no conversion instructions)
        acc[2 * i] = fma_f32_f16((uint16_t)(gw[i] & 0xffffu), rh, acc[2 * i]);
        acc[2 * i + 1] = fma_f32_f16((uint16_t)(gw[i] >> 16), rh, acc[2 * i + 1]);
      }
    }
    float o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float v = ((ti.t0 + f0 + i) < ti.F) ? acc[i] + kc : 0.f;
      o[i] = fmaxf(v, 0.f) + alpha_mask * fminf(v, 0.f);
    }
    *reinterpret_cast<uint4*>(skipa + idx) = pack8(o);
  }
}

}  // namespace fb
}  // namespace b200sep

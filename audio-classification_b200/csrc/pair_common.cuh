// CTA-pair (tcgen05 cta_group::2) building blocks of the K_A pair kernel (fused_pair.cuh):
//   * the A operand of the MMA lives in TENSOR MEMORY (the `[a_tmem]` form of tcgen05.mma): a CTA's 128 rows of the
//     [res | skip] weight matrix stay resident in 256 TMEM columns for the whole launch -- no weight byte is re-read
//     from L2 or shared memory per m-tile;
//   * B (the producers' h2 stage, MN-major) comes from the shared memory of BOTH CTAs (N = 256: 128 frames each);
//   * D: each CTA's tensor memory receives [its 128 output channels] x [256 frames of the tile pair].
// Plus a probe kernel (b200sep_debug_ts_probe) that checks exactly these semantics against a plain GEMM.
#pragma once
#include "fused_block.cuh"

namespace b200sep {
namespace pr2 {
using namespace tc;
using namespace fb;

// 32 lanes x 32 consecutive 32-bit columns: thread i of the warp writes lane (taddr.lane + i)
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&u)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(u[0]), "r"(u[1]), "r"(u[2]), "r"(u[3]), "r"(u[4]), "r"(u[5]), "r"(u[6]), "r"(u[7]), "r"(u[8]), "r"(u[9]),
      "r"(u[10]), "r"(u[11]), "r"(u[12]), "r"(u[13]), "r"(u[14]), "r"(u[15]), "r"(u[16]), "r"(u[17]), "r"(u[18]),
      "r"(u[19]), "r"(u[20]), "r"(u[21]), "r"(u[22]), "r"(u[23]), "r"(u[24]), "r"(u[25]), "r"(u[26]), "r"(u[27]),
      "r"(u[28]), "r"(u[29]), "r"(u[30]), "r"(u[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// D[tmem] (+)= A[tmem] * B[smem desc]   (kind::f16, one CTA)
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// ... CTA pair: A = 256 rows (128 from each CTA's tensor memory), B = N rows (N/2 from each CTA's shared memory),
// D -> both CTAs' tensor memory.  Issued by the leader CTA only.
__device__ __forceinline__ void umma_f16_ts_pair(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                                 uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// kind::f16 instruction descriptor: D fp32, A / B fp16, A K-major (tensor memory operands cannot be transposed),
// B MN-major (frames contiguous), M x N
__host__ __device__ constexpr uint32_t make_idesc_ts(int m, int n) {
  return (1u << 4) | (0u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// A thread's weight row (k0 .. k0 + 63 of one output channel, 128 bytes) -> 32 tensor-memory columns of its lane:
// column j holds the fp16 pair (k0 + 2 j, k0 + 2 j + 1), low half first -- the K-major operand image the MMA expects.
__device__ __forceinline__ void tmem_store_w64(uint32_t taddr, const __half* __restrict__ wrow) {
  uint32_t u[32];
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(wrow) + q);
    u[4 * q] = v.x; u[4 * q + 1] = v.y; u[4 * q + 2] = v.z; u[4 * q + 3] = v.w;
  }
  tmem_st32(taddr, u);
}

// ---------------------------------------------------------------------------------------------------------------
// Probe: out[c][f] = sum_k W[c][k] * H[f][k]  with K = 64, through the tensor-memory-A MMA.
//   pair == 0: one CTA,  M = 128 channels (rows 0..127 of W), N = 128 frames (0..127);
//   pair == 1: a CTA pair, M = 256 (128 rows per CTA), N = 256 (128 frames per CTA), leader issues.
// w16: [256][64] fp16, h16: tile-transposed [2][64 k][128 frames] fp16 (the layout of every activation here),
// out: [256][256] fp32 (pair == 0 fills the top-left 128 x 128 block).  128 threads per CTA.
// ---------------------------------------------------------------------------------------------------------------
template <int PAIR>
__global__ void __launch_bounds__(128, 1) ts_probe_kernel(const __half* __restrict__ w16, const __half* __restrict__ h16,
                                                          float* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t stage = smem_base;             // 16 KiB B-operand stage (64 k x 128 frames, two swizzled 64 x 64 boxes)
  const uint32_t bar = smem_base + 16384u;      // done barrier
  const uint32_t tptr = smem_base + 16384u + 8u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = PAIR ? (int)cluster_ctarank() : 0;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  if (PAIR) cluster_sync_all();
  if (warp == 0) {
    if (PAIR) tmem_alloc_pair(tptr, 512);
    else tmem_alloc(tptr, 512);
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tptr));
  // weights -> tensor memory columns [256, 288): lane = channel row of this CTA
  const int row = warp * 32 + lane;
  tmem_store_w64(tmem_base + ((uint32_t)(warp * 32) << 16) + 256u, w16 + (size_t)(rank * 128 + row) * 64);
  tmem_st_wait();
  // this CTA's 128 frames of H -> the swizzled MN-major stage (what the depthwise producers write)
  const __half* hsrc = h16 + (size_t)rank * 64 * 128;
  for (int i = threadIdx.x; i < 64 * 16; i += 128) {
    const int k = i >> 4, f0 = (i & 15) * 8;
    st_shared_v4(stage + a_chunk_off(k, f0), ldg_v4(hsrc + (size_t)k * 128 + f0));
  }
  fence_proxy_async();
  tcgen05_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();
  tcgen05_fence_after();
  if (threadIdx.x == 0 && rank == 0) {
    constexpr uint32_t idesc = make_idesc_ts(PAIR ? 256 : 128, PAIR ? 256 : 128);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint64_t db = make_smem_desc(stage + j * 2048u, 8192u, 1024u);
      if (PAIR) umma_f16_ts_pair(tmem_base, tmem_base + 256u + (uint32_t)(j * 8), db, idesc, (uint32_t)(j != 0));
      else umma_f16_ts(tmem_base, tmem_base + 256u + (uint32_t)(j * 8), db, idesc, (uint32_t)(j != 0));
    }
    if (PAIR) umma_commit_pair(bar, (uint16_t)0x3);
    else umma_commit(bar);
  }
  mbar_wait_b(nullptr, bar, 0);
  tcgen05_fence_after();
  const int ncol = PAIR ? 256 : 128;
  for (int c0 = 0; c0 < ncol; c0 += 32) {
    float v[32];
    tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[(size_t)(rank * 128 + row) * 256 + c0 + j] = v[j];
  }
  tcgen05_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();
  if (warp == 0) {
    tcgen05_fence_after();
    if (PAIR) tmem_dealloc_pair(tmem_base, 512);
    else tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace pr2
}  // namespace b200sep

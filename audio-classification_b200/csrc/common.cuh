// Shared definitions for the b200sep kernels (sm_100a only).
//
// Activation layout ("tile-transposed"): every per-frame activation tensor with C channels is
// stored as buf[m_tile][c][128] -- frames are grouped in m-tiles of 128, and inside an m-tile the
// frame index is the fastest dimension.  Each segment is padded to a whole number of m-tiles, so
// an m-tile never straddles two segments.  Consequences:
//   * a GEMM-epilogue thread owns one frame (= one TMEM lane); consecutive lanes touch consecutive
//     addresses for every channel -> all epilogue loads/stores are coalesced without staging;
//   * the tensor-core A operand (frames x channels) is "MN-major" and is fetched by TMA as
//     64(k) x 64(m) boxes with the 128-byte swizzle;
//   * the depthwise conv is a 1-D stencil along the contiguous dimension.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200sep {

constexpr int TILE_M = 128;  // frames per m-tile
constexpr int C_ENC = 512;   // n_filters (N)
constexpr int C_BN = 128;    // bn_chan (B)
constexpr int C_HID = 512;   // hid_chan (H)
constexpr int C_SKIP = 128;  // skip_chan (Sc)
constexpr float GLN_EPS = 1e-8f;  // asteroid GlobLN default; the per-gLN value actually used comes from the model (b200sep_weights.gln_eps)
static_assert(C_ENC == C_HID, "one inverse element count per segment serves every gLN (all normalise 512 channels)");

// Everything a kernel needs to know about an m-tile, in one 32-byte record (one load, no dependent
// tile -> segment -> fields chain).
struct __align__(16) TileInfo {
  int32_t seg;      // segment index
  int32_t t0;       // first segment-relative frame of the tile
  int32_t F;        // valid frames of the segment
  int32_t tile0;    // first m-tile of the segment
  int32_t ntiles;   // m-tiles of the segment
  int32_t pad;
  double inv_count; // 1 / (512 * F)
};
static_assert(sizeof(TileInfo) == 32, "TileInfo must be 32 bytes");

__device__ __forceinline__ TileInfo load_tile_info(const TileInfo* p) {
  const uint4 a = __ldg(reinterpret_cast<const uint4*>(p));
  const uint4 b = __ldg(reinterpret_cast<const uint4*>(p) + 1);
  TileInfo ti;
  ti.seg = (int)a.x; ti.t0 = (int)a.y; ti.F = (int)a.z; ti.tile0 = (int)a.w;
  ti.ntiles = (int)b.x; ti.pad = 0;
  ti.inv_count = __hiloint2double((int)b.w, (int)b.z);
  return ti;
}

// Per-batch tables (device pointers), built on the host for every call.
struct Tables {
  const TileInfo* tile_info;   // [n_tiles]
  const int64_t* seg_in_off;   // [n_seg] sample offset of segment in d_mix
  const int64_t* seg_out_off;  // [n_seg] float offset of segment in d_out (= n_src * in_off)
  const int32_t* seg_len;      // [n_seg] T_i samples
  const int32_t* seg_frames;   // [n_seg] F_i = (T_i - L) / stride + 1
  const int32_t* seg_tile0;    // [n_seg] first m-tile of the segment
  const int32_t* seg_ntiles;   // [n_seg] ceil(F_i / 128)
  const int32_t* tile_seg;     // [n_tiles] m-tile -> segment
  const double* seg_inv_count; // [n_seg] 1 / (512 * F_i): every gLN here normalises 512 channels
  int* flags;                  // device flag word of THIS call (zeroed with the tables): bit 0 wait timeout, bit 1 non-finite output
  int32_t n_seg;
  int32_t n_tiles;
};

__host__ __device__ __forceinline__ size_t tt_index(int m_tile, int C, int c, int r) {
  return ((size_t)m_tile * C + c) * TILE_M + r;
}

// ---- programmatic dependent launch (PDL) ------------------------------------------------------
// Every forward kernel starts with pdl_enter(): it lets the NEXT kernel of the stream begin its launch
// processing / prologue right away, then blocks until the PREVIOUS kernel has fully completed and
// flushed its memory -- same ordering as plain stream serialisation, without the launch bubble.
// (No-ops when the kernel was launched without the programmatic-serialisation attribute.)
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_enter() { pdl_launch_dependents(); pdl_wait(); }

// ---- element conversion helpers -------------------------------------------------------------
__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__half v) { return __half2float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }

// ---- packed fp32x2 math (sm_100 FADD2 / FMUL2 / FFMA2: two fp32 lanes per issue slot) ----------
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }

// PReLU.  A01 (0 <= alpha <= 1, the practical case): max(v, alpha*v) is exact and costs 1.5 issue
// slots per element with the packed multiply; otherwise the general max(v,0) + alpha*min(v,0).
template <bool A01>
__device__ __forceinline__ float2 prelu2(float2 v, float alpha) {
  if (A01) {
    const float2 t = mul2(v, make_float2(alpha, alpha));
    return make_float2(fmaxf(v.x, t.x), fmaxf(v.y, t.y));
  }
  return make_float2(fmaxf(v.x, 0.f) + alpha * fminf(v.x, 0.f), fmaxf(v.y, 0.f) + alpha * fminf(v.y, 0.f));
}

// ---- reductions ------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// A gLN whose input variance is far below its eps divides by ~sqrt(eps) instead of the signal's own scale: the
// normalised signal becomes a small ripple on the bias terms of the following layers, which fp16 activation storage
// cannot carry (measured: 5 dB at var / eps = 1e-4).  The half-precision pipelines report it (flag bit 2,
// B200SEP_FLAG_EPS_DOMINATED) and the host re-runs the batch on the fp32 pipeline.
//
// The same place catches fp16 OVERFLOW anywhere upstream: an inf / NaN in a stored activation makes the next gLN's
// variance non-finite (flag bit 1).  The check at the output alone is not enough -- the sigmoid mask maps +-inf
// scores to a clean 0 / 1, so an overflowed network can emit finite garbage (measured with weights scaled by 1e6).
constexpr float EPS_DOMINATED_RATIO = 1.0f / 16.0f;
__device__ __forceinline__ void flag_eps_dominated(int* flags, float var, float eps) {
  if (flags == nullptr || (threadIdx.x & 31) != 0) return;
  if (!(var <= 3.0e38f)) atomicOr(flags, 2);
  else if (var < eps * EPS_DOMINATED_RATIO) atomicOr(flags, 4);
}

// gLN statistics of one segment from the per-tile partial sums a producer kernel left behind.
// part[(tile0 + i) * slots + s] = (sum, sum of squares) over the valid frames of that slot.
// Must be called by all 32 lanes of a warp; every lane (and every warp that repeats the call)
// gets bit-identical results: fixed traversal order, fp64 accumulation, xor-butterfly.
// inv_count = 1 / (C * F_i) (host-computed).  var = E[x^2] - mean^2 in fp64 (biased, like asteroid
// GlobLN); the final reciprocal square root is taken in fp32 (rsqrtf + one Newton step).
__device__ __forceinline__ void warp_seg_stats(const float2* __restrict__ part, int tile0,
                                               int ntiles, int slots, double inv_count, float eps,
                                               float& mean, float& rstd, int* flags = nullptr) {
  const int lane = threadIdx.x & 31;
  const int n = ntiles * slots;
  const float2* p = part + (size_t)tile0 * slots;
  double s = 0.0, q = 0.0;
  for (int i = lane; i < n; i += 32) {
    const float2 v = p[i];
    s += (double)v.x;
    q += (double)v.y;
  }
  s = warp_sum(s);
  q = warp_sum(q);
  const double m = s * inv_count;
  double var = fma(-m, m, q * inv_count);
  if (var < 0.0) var = 0.0;
  mean = (float)m;
  const float ve = (float)var + eps;
  float y = rsqrtf(ve);
  y = y * (1.5f - 0.5f * ve * y * y);  // Newton step: ~0.5 ulp
  rstd = y;
  flag_eps_dominated(flags, (float)var, eps);
}

}  // namespace b200sep

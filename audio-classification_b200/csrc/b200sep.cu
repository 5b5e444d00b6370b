// b200sep: host side of the C ABI declared in include/b200sep.h -- model creation (weight folding,
// operand conversion, TMA descriptors), per-batch layout tables, and the launch sequence of one
// Conv-TasNet forward over a ragged batch of segments; plus the batched PIT entry point.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <tuple>
#include <type_traits>
#include <queue>
#include <vector>

#include "../../include/b200sep.h"
#include "bandwidth_kernels.cuh"
#include "common.cuh"
#include "epilogues.cuh"
#include "gemm_simt.cuh"
#include "gemm_tc.cuh"
#include "fused_block.cuh"
#include "pair_common.cuh"
#include "fused_pair.cuh"
#include "encoder_tc.cuh"
#include "tail_kernel.cuh"
#include "pit.cuh"
#include "aux_kernels.cuh"

using namespace b200sep;

// ------------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_err;

static int fail(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return 1;
}
#define CK(expr)                                                                            \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) return fail("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                                       __FILE__, __LINE__);                                 \
  } while (0)

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Function attributes (opt-in shared-memory size) are per DEVICE: one flag per (kernel instantiation, device).
constexpr int kMaxDev = 64;
struct DevOnce {
  std::atomic<uint64_t> mask{0};
  bool need(int dev) const { return dev < 0 || dev >= kMaxDev || ((mask.load(std::memory_order_acquire) >> dev) & 1ull) == 0; }
  void done(int dev) { if (dev >= 0 && dev < kMaxDev) mask.fetch_or(1ull << dev, std::memory_order_release); }
};
static int current_device_index() {
  int d = 0;
  if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= kMaxDev) return 0;
  return d;
}

// ------------------------------------------------------------------------------------------------
// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time libcuda dependency)
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// 2-D fp16 tensor map: inner dimension `inner` elements (contiguous), `outer` rows of pitch
// `pitch_bytes`; box = box_inner x box_outer; 128-byte swizzle (box_inner * 2 bytes must be 128).
static int make_tmap_2d(CUtensorMap* tm, const void* base, uint64_t inner, uint64_t outer,
                        uint64_t pitch_bytes, uint32_t box_inner, uint32_t box_outer,
                        CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return fail("cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {pitch_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return 0;
}
// activations [n_tiles][C][128] fp16 as a 2-D tensor (128 x n_tiles*C), box 64 x 64
static int make_tmap_act(CUtensorMap* tm, const void* base, int n_tiles, int C) {
  return make_tmap_2d(tm, base, TILE_M, (uint64_t)n_tiles * C, TILE_M * 2, 64, 64);
}
// raw (unswizzled) view of an activation buffer for the depthwise producers: box = `frames` x 64 channels
static int make_tmap_act_raw(CUtensorMap* tm, const void* base, int n_tiles, int C, int frames) {
  return make_tmap_2d(tm, base, TILE_M, (uint64_t)n_tiles * C, TILE_M * 2, (uint32_t)frames, 64,
                      CU_TENSOR_MAP_SWIZZLE_NONE);
}
// weights [N][K] fp16, box 64(k) x bn rows
static int make_tmap_w(CUtensorMap* tm, const void* base, int N, int K, int bn) {
  return make_tmap_2d(tm, base, K, N, (uint64_t)K * 2, 64, bn);
}

// ------------------------------------------------------------------------------------------------
// pinned staging ring for per-call tables (H2D without pageable-copy synchronisation)
// ------------------------------------------------------------------------------------------------
struct StagingSlot {
  void* host = nullptr;
  size_t cap = 0;
  cudaEvent_t ev = nullptr;
};
struct Staging {
  static constexpr int kSlots = 8;
  StagingSlot slots[kSlots];
  int next = 0;
  std::mutex mu;
  // Returns a pinned buffer of at least `bytes`, safe to overwrite (its previous use has completed).
  int acquire(size_t bytes, StagingSlot** out) {
    StagingSlot& s = slots[next];
    next = (next + 1) % kSlots;
    if (!s.ev) CK(cudaEventCreateWithFlags(&s.ev, cudaEventDisableTiming));
    else CK(cudaEventSynchronize(s.ev));
    if (s.cap < bytes) {
      if (s.host) CK(cudaFreeHost(s.host));
      s.host = nullptr;
      s.cap = 0;
      const size_t cap = align_up(bytes * 2, 4096);
      CK(cudaHostAlloc(&s.host, cap, cudaHostAllocDefault));
      s.cap = cap;
    }
    *out = &s;
    return 0;
  }
  ~Staging() {
    for (auto& s : slots) {
      if (s.host) cudaFreeHost(s.host);
      if (s.ev) cudaEventDestroy(s.ev);
    }
  }
};

// ------------------------------------------------------------------------------------------------
// model
// ------------------------------------------------------------------------------------------------
struct BlockW {
  void* w_in = nullptr;   // [512][128] T
  float* b_in = nullptr;  // [512]
  float a1 = 0.25f;
  float *g1 = nullptr, *be1 = nullptr, *wd = nullptr, *bd = nullptr;  // [512],[512],[512][3],[512]
  float a2 = 0.25f;
  void* w_rs = nullptr;  // [256][512] T ([128][512] for the last block): (W . gamma2)
  void* w_rs_pair = nullptr;  // [256][512] fp16 for the pair kernel: w_rs, or (last block) [128 zero rows | skip rows]
  float *rs = nullptr, *cb = nullptr;
  int dilation = 1;
  int n_out = 256;
  float eps1 = GLN_EPS, eps2 = GLN_EPS;  // eps of gLN1 / gLN2 of this block (b200sep_weights.gln_eps)
  CUtensorMap tm_in, tm_rs;
  float4 *dw_t0 = nullptr, *dw_t1 = nullptr;  // folded depthwise tap tables (fused_block.cuh DwTab)
  float* dw_cb = nullptr;
};

struct b200sep_model {
  b200sep_config cfg;
  int device = 0;
  int n_sm = 148;
  int n_layers = 0;
  bool half = true, use_tc = true;
  bool fused = true;  // fp16 tensor-core mode: two fused kernels per TCN block (fused_block.cuh)
  bool fused_tail = true;  // ... and mask GEMM + mask application + decoder in one kernel (tail_kernel.cuh)
  bool ka_pair = false;    // K_A as a CTA pair with the [res | skip] weights resident in tensor memory (fused_pair.cuh)
  std::atomic<int> max_split{0};  // b200sep_set_max_split: cap on the concurrent sub-batches of a call (0: automatic)
  void* dec_w16 = nullptr;  // [L][512] fp16 decoder filterbank (B operand of the decoder MMA)
  CUtensorMap tm_mask128, tm_wd;
  std::vector<void*> allocs;
  float *enc_w = nullptr, *dec_w = nullptr;
  void* bn_w = nullptr;
  float *bn_rs = nullptr, *bn_cb = nullptr;
  std::vector<BlockW> blocks;
  void* mask_w = nullptr;
  float* mask_b = nullptr;
  float alpha_mask = 0.25f;
  CUtensorMap tm_bn, tm_mask;
  float *skip_rs = nullptr, *skip_cb = nullptr;  // [n_layers][128] skip halves of rs / cb (skip_merge_kernel)
  float* skip_eps = nullptr;                     // [n_layers] eps of every block's gLN2 (skip_merge_kernel)
  float eps_bn = GLN_EPS;                        // eps of the bottleneck gLN
  Staging staging;
  std::mutex mu;
  int debug_stop_after = 0;  // 0 = run everything; k = stop after k-th launch (tests)
  // optional per-launch CUDA-event timing (bench.py roofline pass)
  // concurrent sub-batches (b200sep_separate_batch): auxiliary streams and fork / join events
  // (two sets: calls of the host entry point alternate between them, see hin below)
  cudaStream_t aux_st2[2][4] = {};
  cudaEvent_t ev_fork2[2] = {}, ev_join2[2][4] = {};
  // CUDA graphs of the launch sequence for small batches (one per (n_seg, n_tiles, buffer addresses)): a single short
  // segment is launch-bound (52 launches, ~7 us of host time each), the reference's per-segment separate() loop lives there
  struct GraphKey {
    int n_seg, n_tiles;
    const void *mix, *out, *ws;
    bool operator<(const GraphKey& o) const {
      return std::tie(n_seg, n_tiles, mix, out, ws) < std::tie(o.n_seg, o.n_tiles, o.mix, o.out, o.ws);
    }
  };
  cudaStream_t own_st = nullptr;  // stands in for the legacy default stream (StreamHop)
  cudaEvent_t own_ev = nullptr;
  std::map<GraphKey, cudaGraphExec_t> graphs;
  std::mutex graph_mu;
  long long graph_hits = 0, graph_misses = 0;  // a caller that passes fresh buffers every call never hits: give up then
  // pinned input staging of b200sep_separate_batch_host: two slots used alternately, so that a caller may queue the
  // next call (on another stream, with another workspace) while the previous one is still copying / computing
  static constexpr int kHin = 2;
  float* hin[kHin] = {nullptr, nullptr};
  size_t hin_cap[kHin] = {0, 0};  // floats
  cudaEvent_t hin_ev[kHin] = {nullptr, nullptr};
  int hin_next = 0;
  bool prof_on = false;
  std::vector<cudaEvent_t> prof_ev;
  int prof_n = 0;
};

static Staging g_pit_staging[kMaxDev];  // per device (events and pinned slots belong to a device context)
static std::mutex g_pit_mu;
// optional per-launch timing of the PIT kernels (b200sep_pit_profile_enable / _read): three events per device
struct PitProf { cudaEvent_t ev[3] = {nullptr, nullptr, nullptr}; bool valid = false; };
static PitProf g_pit_prof[kMaxDev];
static std::atomic<int> g_pit_prof_on{0};

template <typename T> static T host_cvt(float v);
template <> float host_cvt<float>(float v) { return v; }
template <> __half host_cvt<__half>(float v) { return __float2half_rn(v); }
static inline float host_back(float v) { return v; }
static inline float host_back(__half v) { return __half2float(v); }

static int dev_upload(b200sep_model* m, const void* src, size_t bytes, void** out) {
  void* d = nullptr;
  CK(cudaMalloc(&d, std::max<size_t>(bytes, 16)));
  m->allocs.push_back(d);
  CK(cudaMemcpy(d, src, bytes, cudaMemcpyHostToDevice));
  *out = d;
  return 0;
}
static int dev_upload_f(b200sep_model* m, const std::vector<float>& v, float** out) {
  return dev_upload(m, v.data(), v.size() * sizeof(float), reinterpret_cast<void**>(out));
}

// Folds gamma into W (operand type T), returns rowsum of the rounded operand and W*beta + b.
template <typename T>
static int fold_and_upload(b200sep_model* m, const std::vector<const float*>& Ws, const std::vector<const float*>& bs,
                           int rows_each, int K, const float* gamma, const float* beta, void** d_w, float** d_rs,
                           float** d_cb, std::vector<float>* h_rs = nullptr, std::vector<float>* h_cb = nullptr) {
  const int N = rows_each * (int)Ws.size();
  std::vector<T> w((size_t)N * K);
  std::vector<float> rs(N), cb(N);
  for (size_t g = 0; g < Ws.size(); ++g)
    for (int r = 0; r < rows_each; ++r) {
      const int n = (int)g * rows_each + r;
      double srs = 0.0, scb = 0.0;
      for (int k = 0; k < K; ++k) {
        const float wv = Ws[g][(size_t)r * K + k];
        const T q = host_cvt<T>(wv * gamma[k]);
        w[(size_t)n * K + k] = q;
        srs += (double)host_back(q);
        scb += (double)wv * (double)beta[k];
      }
      rs[n] = (float)srs;
      cb[n] = (float)(scb + (double)bs[g][r]);
    }
  if (dev_upload(m, w.data(), w.size() * sizeof(T), d_w)) return 1;
  if (dev_upload_f(m, rs, d_rs)) return 1;
  if (dev_upload_f(m, cb, d_cb)) return 1;
  if (h_rs) *h_rs = rs;
  if (h_cb) *h_cb = cb;
  return 0;
}

template <typename T>
static int upload_cvt(b200sep_model* m, const float* src, size_t n, void** out) {
  std::vector<T> w(n);
  for (size_t i = 0; i < n; ++i) w[i] = host_cvt<T>(src[i]);
  return dev_upload(m, w.data(), n * sizeof(T), out);
}

template <typename T>
static int build_weights(b200sep_model* m, const b200sep_weights* w) {
  const b200sep_config& c = m->cfg;
  const int L = c.kernel_size, S = c.n_src;
  if (dev_upload(m, w->enc_filters, (size_t)C_ENC * L * 4, reinterpret_cast<void**>(&m->enc_w))) return 1;
  if (dev_upload(m, w->dec_filters, (size_t)C_ENC * L * 4, reinterpret_cast<void**>(&m->dec_w))) return 1;
  if (fold_and_upload<T>(m, {w->bn_weight}, {w->bn_bias}, C_BN, C_ENC, w->bn_gamma, w->bn_beta, &m->bn_w,
                         &m->bn_rs, &m->bn_cb))
    return 1;
  m->blocks.resize(m->n_layers);
  std::vector<float> skip_rs, skip_cb, skip_eps;
  auto eps_at = [&](int k) -> float {
    const float e = w->gln_eps ? w->gln_eps[k] : GLN_EPS;
    return e;
  };
  m->eps_bn = eps_at(0);
  for (int i = 0; i < m->n_layers; ++i) {
    BlockW& b = m->blocks[i];
    const bool last = (i == m->n_layers - 1);
    b.dilation = 1 << (i % c.n_blocks);
    b.eps1 = eps_at(1 + 2 * i);
    b.eps2 = eps_at(2 + 2 * i);
    skip_eps.push_back(b.eps2);
    if (upload_cvt<T>(m, w->in_weight[i], (size_t)C_HID * C_BN, &b.w_in)) return 1;
    if (dev_upload(m, w->in_bias[i], C_HID * 4, reinterpret_cast<void**>(&b.b_in))) return 1;
    b.a1 = w->prelu1[i][0];
    if (dev_upload(m, w->norm1_gamma[i], C_HID * 4, reinterpret_cast<void**>(&b.g1))) return 1;
    if (dev_upload(m, w->norm1_beta[i], C_HID * 4, reinterpret_cast<void**>(&b.be1))) return 1;
    if (dev_upload(m, w->dw_weight[i], C_HID * 3 * 4, reinterpret_cast<void**>(&b.wd))) return 1;
    if (dev_upload(m, w->dw_bias[i], C_HID * 4, reinterpret_cast<void**>(&b.bd))) return 1;
    b.a2 = w->prelu2[i][0];
    {  // folded depthwise tap tables (fp64 products, rounded once)
      std::vector<float> t0(C_HID * 4), t1(C_HID * 4), tcb(C_HID);
      for (int ch = 0; ch < C_HID; ++ch) {
        const double g = w->norm1_gamma[i][ch], be = w->norm1_beta[i][ch], bias = w->dw_bias[i][ch];
        const double w0 = w->dw_weight[i][ch * 3 + 0], w1 = w->dw_weight[i][ch * 3 + 1], w2 = w->dw_weight[i][ch * 3 + 2];
        t0[ch * 4 + 0] = (float)(w0 * g); t0[ch * 4 + 1] = (float)(w1 * g); t0[ch * 4 + 2] = (float)(w2 * g);
        t0[ch * 4 + 3] = (float)((w0 + w1 + w2) * g);
        t1[ch * 4 + 0] = (float)(w0 * be); t1[ch * 4 + 1] = (float)(w1 * be); t1[ch * 4 + 2] = (float)(w2 * be);
        t1[ch * 4 + 3] = (float)bias;
        tcb[ch] = (float)((w0 + w1 + w2) * be + bias);
      }
      if (dev_upload_f(m, t0, reinterpret_cast<float**>(&b.dw_t0))) return 1;
      if (dev_upload_f(m, t1, reinterpret_cast<float**>(&b.dw_t1))) return 1;
      if (dev_upload_f(m, tcb, &b.dw_cb)) return 1;
    }
    std::vector<float> h_rs, h_cb;
    if (!last) {
      b.n_out = 2 * C_BN;
      if (fold_and_upload<T>(m, {w->res_weight[i], w->skip_weight[i]}, {w->res_bias[i], w->skip_bias[i]}, C_BN,
                             C_HID, w->norm2_gamma[i], w->norm2_beta[i], &b.w_rs, &b.rs, &b.cb, &h_rs, &h_cb))
        return 1;
      skip_rs.insert(skip_rs.end(), h_rs.begin() + C_BN, h_rs.end());
      skip_cb.insert(skip_cb.end(), h_cb.begin() + C_BN, h_cb.end());
    } else {  // the last block's residual output is never read: skip half only
      b.n_out = C_SKIP;
      if (fold_and_upload<T>(m, {w->skip_weight[i]}, {w->skip_bias[i]}, C_SKIP, C_HID, w->norm2_gamma[i],
                             w->norm2_beta[i], &b.w_rs, &b.rs, &b.cb, &h_rs, &h_cb))
        return 1;
      skip_rs.insert(skip_rs.end(), h_rs.begin(), h_rs.end());
      skip_cb.insert(skip_cb.end(), h_cb.begin(), h_cb.end());
    }
    b.w_rs_pair = b.w_rs;
    if (last && std::is_same<T, __half>::value) {
      // the pair kernel always runs M = 256: rank 0's (dead) residual rows are zeros, rank 1 holds the skip rows
      void* padded = nullptr;
      CK(cudaMalloc(&padded, (size_t)2 * C_BN * C_HID * 2));
      m->allocs.push_back(padded);
      CK(cudaMemset(padded, 0, (size_t)C_BN * C_HID * 2));
      CK(cudaMemcpy(static_cast<uint8_t*>(padded) + (size_t)C_BN * C_HID * 2, b.w_rs, (size_t)C_SKIP * C_HID * 2,
                    cudaMemcpyDeviceToDevice));
      b.w_rs_pair = padded;
    }
  }
  if (dev_upload_f(m, skip_rs, &m->skip_rs)) return 1;
  if (dev_upload_f(m, skip_cb, &m->skip_cb)) return 1;
  if (dev_upload_f(m, skip_eps, &m->skip_eps)) return 1;
  {  // decoder filterbank as a K-major fp16 B operand: [L][512]
    std::vector<__half> wd((size_t)L * C_ENC);
    for (int n = 0; n < C_ENC; ++n)
      for (int l = 0; l < L; ++l) wd[(size_t)l * C_ENC + n] = __float2half_rn(w->dec_filters[(size_t)n * L + l]);
    if (dev_upload(m, wd.data(), wd.size() * sizeof(__half), &m->dec_w16)) return 1;
  }
  m->alpha_mask = w->mask_prelu[0];
  if (upload_cvt<T>(m, w->mask_weight, (size_t)S * C_ENC * C_SKIP, &m->mask_w)) return 1;
  if (dev_upload(m, w->mask_bias, (size_t)S * C_ENC * 4, reinterpret_cast<void**>(&m->mask_b))) return 1;
  return 0;
}

// ------------------------------------------------------------------------------------------------
// batch layout
// ------------------------------------------------------------------------------------------------
struct Layout {
  int n_seg = 0, n_tiles = 0;
  size_t tables_bytes = 0;
  // byte offsets inside the workspace
  size_t o_tables, o_enc, o_x, o_xa, o_skip, o_skipa, o_h1, o_h2, o_g, o_stash, o_masked, o_p0, o_p1, o_p2, total;
  size_t stash_stride = 0, part2_stride = 0;  // elements per block slot
  // offsets inside the table block
  size_t t_info, t_in_off, t_out_off, t_inv, t_len, t_frames, t_tile0, t_ntiles, t_tile_seg, t_flags;
};

static int compute_layout(const b200sep_config& c, bool half, bool fused, const int32_t* lens, int n_seg, Layout* lo) {
  if (n_seg <= 0 || !lens) return fail("empty batch");
  long long tiles = 0;
  for (int i = 0; i < n_seg; ++i) {
    if (lens[i] < c.kernel_size)
      return fail("segment %d has %d samples < kernel_size %d (conv1d would fail in the reference too)", i,
                  lens[i], c.kernel_size);
    const long long F = (lens[i] - c.kernel_size) / c.stride + 1;
    tiles += (F + TILE_M - 1) / TILE_M;
  }
  if (tiles > (1ll << 22)) return fail("batch too large (%lld m-tiles)", tiles);
  lo->n_seg = n_seg;
  lo->n_tiles = (int)tiles;
  size_t t = 0;
  lo->t_info = t; t += align_up((size_t)tiles * sizeof(TileInfo), 16);
  lo->t_in_off = t; t += align_up((size_t)n_seg * 8, 16);
  lo->t_out_off = t; t += align_up((size_t)n_seg * 8, 16);
  lo->t_inv = t; t += align_up((size_t)n_seg * 8, 16);
  lo->t_len = t; t += align_up((size_t)n_seg * 4, 16);
  lo->t_frames = t; t += align_up((size_t)n_seg * 4, 16);
  lo->t_tile0 = t; t += align_up((size_t)n_seg * 4, 16);
  lo->t_ntiles = t; t += align_up((size_t)n_seg * 4, 16);
  lo->t_tile_seg = t; t += align_up((size_t)tiles * 4, 16);
  lo->t_flags = t; t += 16;  // device flag word of the call (the zero-filled staging image clears it)
  lo->tables_bytes = t;
  const size_t es = half ? 2 : 4;  // activation element size
  const size_t rows = (size_t)tiles * TILE_M;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o = align_up(o + bytes, 1024); return r; };
  lo->o_tables = take(lo->tables_bytes);
  lo->o_enc = take(rows * C_ENC * es);
  lo->o_x = take(rows * C_BN * 4);
  lo->o_xa = half ? take(rows * C_BN * es) : lo->o_x;  // fp16 operand shadow (unfused) / fp16 residual stream (fused)
  lo->o_skip = take(fused ? 0 : rows * C_SKIP * 4);
  lo->o_skipa = take(rows * C_SKIP * es);
  lo->o_h1 = take(rows * C_HID * es);
  lo->o_h2 = take(fused ? 0 : rows * C_HID * es);
  // fused path: raw residual-conv output (fp16) and one raw skip-conv stash slot per block; the unfused
  // path's h2 and skip accumulators are not needed then
  const int n_layers = c.n_blocks * c.n_repeats;
  lo->o_g = take(fused ? rows * C_BN * 2 : 0);
  lo->stash_stride = rows * C_SKIP;
  lo->o_stash = take(fused ? (size_t)n_layers * lo->stash_stride * 2 : 0);
  lo->o_masked = take(rows * (size_t)c.n_src * C_ENC * es);
  lo->o_p0 = take((size_t)tiles * 2 * 8);
  lo->o_p1 = take((size_t)tiles * 8 * 8);
  lo->part2_stride = (size_t)tiles;
  lo->o_p2 = take(fused ? (size_t)tiles * n_layers * 8 : (size_t)tiles * 8 * 8);
  lo->total = o;
  return 0;
}

// ------------------------------------------------------------------------------------------------
// launch helpers
// ------------------------------------------------------------------------------------------------
// Programmatic dependent launch: optional (B200SEP_PDL=1 enables it for the fused fp16 pipeline only).  Measured
// on B200: within +-0.5 % of plain stream order, so it is off by default.  (An intermittent launch failure first
// attributed to it was a parity-aliasing bug of the odd-length raw ring in K_A -- see fused_block.cuh -- and
// did not recur in 26 stress runs after that fix.)  The unfused pipelines keep plain stream order: with
// overlapping launches they showed stale per-SM L1 lines of buffers rewritten in place.
static bool pdl_env() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("B200SEP_PDL");
    v = (e && atoi(e) == 1) ? 1 : 0;
  }
  return v == 1;
}
static thread_local bool g_pdl_on = false;  // set per forward() call
static bool pdl_enabled() { return g_pdl_on; }

// Launch with programmatic stream serialisation (PDL): the kernel may start while its predecessor
// drains; every kernel orders itself with griddepcontrol.wait (pdl_enter) before touching global data.
template <typename... KArgs, typename... Args>
static cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                            Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
template <int BN, int K, class Epi>
static int launch_tc(const b200sep_model* m, const CUtensorMap& ta, const CUtensorMap& tbm, int n_mtiles, int N,
                     const Epi& epi, const Tables& tb, cudaStream_t st, uint32_t a_lbo = 8192, uint32_t a_sbo = 1024) {
  using Cfg = tc::TcConfig<BN, K>;
  auto kern = tc::gemm_tc_kernel<BN, K, Epi>;
  { static DevOnce once; if (once.need(m->device)) { CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES)); once.done(m->device); } }
  const int n_ntiles = N / BN;
  const int total = n_mtiles * n_ntiles;
  const int grid = std::min(total, m->n_sm);
  CK(launch_k(kern, dim3(grid), dim3(Cfg::THREADS), Cfg::SMEM_BYTES, st, ta, tbm, n_mtiles, n_ntiles, a_lbo, a_sbo, epi, tb));
  return 0;
}

template <typename T, int K, class Epi>
static int launch_simt(const T* A, const T* B, int n_mtiles, int N, const Epi& epi, const Tables& tb,
                       cudaStream_t st) {
  dim3 grid(N / 128, n_mtiles);
  CK(launch_k(gemm_simt_kernel<T, K, Epi>, grid, dim3(256), 0, st, A, B, epi, tb));
  return 0;
}

// One GEMM of the model in the configured engine.  BN is the tcgen05 N-tile.
template <typename T, int BN, int K, class Epi>
static int run_gemm(const b200sep_model* m, const T* A, const CUtensorMap* tmA, const T* B, const CUtensorMap* tmB,
                    int n_mtiles, int N, const Epi& epi, const Tables& tb, cudaStream_t st) {
  if constexpr (std::is_same<T, __half>::value) {
    if (m->use_tc) return launch_tc<BN, K, Epi>(m, *tmA, *tmB, n_mtiles, N, epi, tb, st);
  }
  return launch_simt<T, K, Epi>(A, B, n_mtiles, N, epi, tb, st);
}


// ------------------------------------------------------------------------------------------------
// fused TCN block (fp16 tensor-core mode): launches of fused_block.cuh
// ------------------------------------------------------------------------------------------------
struct KaArgs {
  const CUtensorMap* tm_b;
  const CUtensorMap* tm_hc;
  const CUtensorMap* tm_hh;
  __half* g_res;
  __half* g_skip;
  size_t skip_tile_stride;
  const float2* p1;
  float2* p2;
  fb::DwTab tab;
};
template <int BN, int D, bool A01>
static int launch_dw_rs(const b200sep_model* m, const KaArgs& a, const Tables& tb, cudaStream_t st) {
  using Cfg = fb::KaCfg<BN>;
  auto kern = fb::dw_rs_kernel<BN, D, A01>;
  { static DevOnce once; if (once.need(m->device)) { CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES)); once.done(m->device); } }
  const int grid = std::min(tb.n_tiles, m->n_sm);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(Cfg::THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  CK(cudaLaunchKernelEx(&cfg, kern, *a.tm_b, *a.tm_hc, *a.tm_hh, a.g_res, a.g_skip, a.skip_tile_stride, a.p1, a.p2, a.tab,
                        tb));
  return 0;
}
template <int BN, bool A01>
static int launch_dw_rs_d(const b200sep_model* m, const KaArgs& a, const Tables& tb, cudaStream_t st) {
  switch (a.tab.dilation) {
    case 1: return launch_dw_rs<BN, 1, A01>(m, a, tb, st);
    case 2: return launch_dw_rs<BN, 2, A01>(m, a, tb, st);
    case 4: return launch_dw_rs<BN, 4, A01>(m, a, tb, st);
    default: return launch_dw_rs<BN, 0, A01>(m, a, tb, st);
  }
}
struct KpArgs {
  const CUtensorMap* tm_hc;
  const CUtensorMap* tm_hh;
  const CUtensorMap* tm_res;
  const CUtensorMap* tm_skip;
  const __half* w256;
  int res_dead, skip_row0, skip_tile_rows;
  const float2* p1;
  float2* p2;
  fb::DwTab tab;
};
template <int D, bool A01, int NG>
static int launch_dw_rs_pair(const b200sep_model* m, const KpArgs& a, const Tables& tb, cudaStream_t st) {
  using Cfg = pr2::KpCfg;
  auto kern = pr2::dw_rs_pair_kernel<D, A01, NG>;
  { static DevOnce once; if (once.need(m->device)) { CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES)); once.done(m->device); } }
  const int n_pairs = std::max(1, std::min((tb.n_tiles + 1) / 2, m->n_sm / 2));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * n_pairs);
  cfg.blockDim = dim3(Cfg::THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  CK(cudaLaunchKernelEx(&cfg, kern, *a.tm_hc, *a.tm_hh, *a.tm_res, *a.tm_skip, a.w256, a.res_dead, a.skip_row0,
                        a.skip_tile_rows, a.p1, a.p2, a.tab, tb));
  return 0;
}
template <bool A01>
static int launch_dw_rs_pair_d(const b200sep_model* m, const KpArgs& a, const Tables& tb, cudaStream_t st) {
  // two producer groups (NG = 2).  Four groups of four warps (the kernel supports NG = 4 where four raw stages fit) were
  // measured equal: 76.9 vs 76.4 us per launch -- the tensor core's own k-block time grows from ~550 to ~890 cycles when
  // all producers stream at once (shared-memory bandwidth), see DESIGN.md.
  switch (a.tab.dilation) {
    case 1: return launch_dw_rs_pair<1, A01, 2>(m, a, tb, st);
    case 2: return launch_dw_rs_pair<2, A01, 2>(m, a, tb, st);
    case 4: return launch_dw_rs_pair<4, A01, 2>(m, a, tb, st);
    default: return launch_dw_rs_pair<0, A01, 2>(m, a, tb, st);
  }
}
template <bool A01>
static int launch_apply_in(const b200sep_model* m, const CUtensorMap& tmw, const CUtensorMap& tmh1, const fb::ApplyParams& ap,
                           const EpiIn<__half, A01>& epi, const Tables& tb, cudaStream_t st) {
  using Cfg = fb::KbCfg;
  auto kern = fb::apply_in_kernel<A01>;
  { static DevOnce once; if (once.need(m->device)) { CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES)); once.done(m->device); } }
  const int grid = std::min(tb.n_tiles, m->n_sm);
  CK(launch_k(kern, dim3(grid), dim3(Cfg::THREADS), Cfg::SMEM_BYTES, st, tmw, tmh1, ap, epi, tb));
  return 0;
}

template <typename T, int L>
static int forward(b200sep_model* m, const Layout& lo, const Tables& tb, const float* d_mix, float* d_out,
                   uint8_t* ws, cudaStream_t st) {
  const b200sep_config& c = m->cfg;
  const int nt = lo.n_tiles;
  T* enc = reinterpret_cast<T*>(ws + lo.o_enc);
  float* x = reinterpret_cast<float*>(ws + lo.o_x);
  T* xa = reinterpret_cast<T*>(ws + lo.o_xa);
  float* skip = reinterpret_cast<float*>(ws + lo.o_skip);
  T* skipa = reinterpret_cast<T*>(ws + lo.o_skipa);
  T* h1 = reinterpret_cast<T*>(ws + lo.o_h1);
  T* h2 = reinterpret_cast<T*>(ws + lo.o_h2);
  __half* gbuf = reinterpret_cast<__half*>(ws + lo.o_g);
  T* masked = reinterpret_cast<T*>(ws + lo.o_masked);
  float2* p0 = reinterpret_cast<float2*>(ws + lo.o_p0);
  float2* p1 = reinterpret_cast<float2*>(ws + lo.o_p1);
  float2* p2 = reinterpret_cast<float2*>(ws + lo.o_p2);
  constexpr bool kHalf = std::is_same<T, __half>::value;
  const bool fused_mode = kHalf && m->use_tc && m->fused;
  g_pdl_on = fused_mode && pdl_env();
  T* xa_or_null = (kHalf && !fused_mode) ? xa : nullptr;
  const T* a_x = kHalf ? xa : reinterpret_cast<const T*>(x);  // A operand of the in-conv

  CUtensorMap tm_enc, tm_xa, tm_h2, tm_skipa;
  if (kHalf && m->use_tc) {
    if (make_tmap_act(&tm_enc, enc, nt, C_ENC)) return 1;
    if (make_tmap_act(&tm_xa, xa, nt, C_BN)) return 1;
    if (make_tmap_act(&tm_h2, h2, nt, C_HID)) return 1;
    if (make_tmap_act(&tm_skipa, skipa, nt, C_SKIP)) return 1;
  }
  int launches = 0;
  m->prof_n = 0;
  auto prof_mark = [&]() {
    if (!m->prof_on) return;
    if ((int)m->prof_ev.size() <= m->prof_n) {
      cudaEvent_t e;
      if (cudaEventCreate(&e) != cudaSuccess) return;
      m->prof_ev.push_back(e);
    }
    cudaEventRecord(m->prof_ev[m->prof_n++], st);
  };
  prof_mark();
  static const bool sync_debug = getenv("B200SEP_SYNC_DEBUG") != nullptr;  // debug: localise a faulting launch
  int launch_no = 0;
  bool sync_failed = false;
  auto stop = [&]() {
    prof_mark();
    ++launch_no;
    if (sync_debug) {
      const cudaError_t e = cudaStreamSynchronize(st);
      if (e != cudaSuccess) {
        fail("launch #%d of the forward failed: %s", launch_no, cudaGetErrorString(e));
        fprintf(stderr, "[b200sep] launch #%d of the forward failed: %s\n", launch_no, cudaGetErrorString(e));
        sync_failed = true;
        return true;
      }
    }
    return m->debug_stop_after > 0 && ++launches >= m->debug_stop_after;
  };
  const int slots1 = (kHalf && m->use_tc) ? C_HID / 256 : C_HID / 128;  // partial slots the in-conv GEMM writes

  // 1. encoder (+ bottleneck-gLN partials)
  static const bool enc_tc = [] { const char* e = getenv("B200SEP_ENCODER"); return !(e && strcmp(e, "simt") == 0); }();
  bool enc_done = false;
  if constexpr (kHalf) {
    if (m->use_tc && enc_tc) {
      // tensor-core encoder (encoder_tc.cuh): taps as the K dimension of one MMA per 128-filter chunk, the waveform
      // split into hi + lo halves; B200SEP_ENCODER=simt selects the CUDA-core kernel
      using Cfg = enc::EncCfg<L>;
      auto kern = enc::encoder_tc_kernel<L>;
      { static DevOnce once; if (once.need(m->device)) { CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES)); once.done(m->device); } }
      CUtensorMap tm_enc_st;   // store boxes of 64 frames x 32 filters: one per epilogue warp and chunk
      if (make_tmap_2d(&tm_enc_st, enc, TILE_M, (uint64_t)nt * C_ENC, TILE_M * 2, 64, 32)) return 1;
      CK(launch_k(kern, dim3(std::min(nt, m->n_sm)), dim3(Cfg::THREADS), Cfg::SMEM_BYTES, st, tm_enc_st, d_mix,
                  (const float*)m->enc_w, p0, tb));
      enc_done = true;
    }
  }
  if (enc_done) {
  } else if constexpr (L == 16) {
    // B200SEP_ENCODER_V2=1: four frames per thread with packed FFMA2 (bandwidth_kernels.cuh) -- half the instructions and a
    // quarter of the shared-memory wavefronts per output, but measured SLOWER (102 vs 94 us: 88 registers, 37 KB of
    // shared memory per CTA, and FFMA2 issues at half rate), so the one-frame-per-thread kernel stays the default
    static const bool enc_v1 = getenv("B200SEP_ENCODER_V2") == nullptr;
    if (!enc_v1) CK(launch_k(encoder4_kernel<T>, dim3(nt, 2), dim3(256), 0, st, d_mix, (const float*)m->enc_w, enc, p0, tb));
    else CK(launch_k(encoder_kernel<T, L>, dim3(nt, 2), dim3(256), 0, st, d_mix, (const float*)m->enc_w, enc, p0, tb));
  } else {
    CK(launch_k(encoder_kernel<T, L>, dim3(nt, 2), dim3(256), 0, st, d_mix, (const float*)m->enc_w, enc, p0, tb));
  }
  if (stop()) return sync_failed ? 1 : 0;

  // 2. bottleneck: x = Wb . gLN(tf_rep) + bb
  {
    // (fused path: the fp16 copy IS the residual stream the TCN blocks update)
    const bool x16 = kHalf && fused_mode;
    EpiBottleneck<T> e{p0, 2, m->bn_rs, m->bn_cb, x16 ? nullptr : x, x16 ? xa : xa_or_null, m->eps_bn};
    if (run_gemm<T, 128, C_ENC>(m, enc, &tm_enc, reinterpret_cast<const T*>(m->bn_w), &m->tm_bn, nt, C_BN, e, tb, st))
      return 1;
  }
  if (stop()) return sync_failed ? 1 : 0;

  // 3. TCN blocks
  bool fused_done = false;
  if constexpr (kHalf) {
    if (fused_mode) {
      // two launches per block: K_B (apply previous G_res + in-conv GEMM -> h1), K_A (depthwise + [res|skip]
      // GEMM -> raw G_res + skip stash slot); after the last block the stash is merged into the mask operand
      __half* stash = reinterpret_cast<__half*>(ws + lo.o_stash);
      CUtensorMap tm_hc, tm_hh[5];  // raw h1 views: centre box and halo boxes of 8..128 frames
      CUtensorMap tm_h1s;  // swizzled store boxes of h1 (K_B epilogue)
      // (64 frames x 32 channels per store: one box per epilogue warp and chunk)
      if (make_tmap_2d(&tm_h1s, h1, TILE_M, (uint64_t)nt * C_HID, TILE_M * 2, 64, 32)) return 1;
      // K_A as a CTA pair with tensor-memory-resident weights (fused_pair.cuh; model created under B200SEP_KA=pair): raw
      // G_res / skip-stash store boxes.  Default: the one-CTA kernel (86 vs 92 us per launch on the configs[1] batch)
      const bool ka_pair = m->ka_pair;
      CUtensorMap tm_gres, tm_stash;
      // (boxes of 32 channels x 64 frames: one per epilogue warp of the pair kernel)
      if (make_tmap_2d(&tm_gres, gbuf, TILE_M, (uint64_t)nt * C_BN, TILE_M * 2, 64, 32)) return 1;
      if (make_tmap_2d(&tm_stash, stash, TILE_M, (uint64_t)nt * m->n_layers * C_SKIP, TILE_M * 2, 64, 32)) return 1;
      if (make_tmap_act_raw(&tm_hc, h1, nt, C_HID, TILE_M)) return 1;
      for (int k = 0; k < 5; ++k)
        if (make_tmap_act_raw(&tm_hh[k], h1, nt, C_HID, 8 << k)) return 1;
      for (int i = 0; i < m->n_layers; ++i) {
        BlockW& b = m->blocks[i];
        const bool last = (i == m->n_layers - 1);
        if (b.dilation > 128) return fail("fused path supports dilations up to 128 (got %d)", b.dilation);
        fb::ApplyParams ap{};
        ap.x = x;
        ap.x16 = reinterpret_cast<__half*>(xa);
        if (i > 0) {
          ap.g = gbuf;
          ap.part2 = p2 + (size_t)(i - 1) * lo.part2_stride;
          ap.rs = m->blocks[i - 1].rs;
          ap.cb = m->blocks[i - 1].cb;
          ap.eps2 = m->blocks[i - 1].eps2;
        }
        int rc;
        if (b.a1 >= 0.f && b.a1 <= 1.f) {
          EpiIn<__half, true> e{b.b_in, b.a1, h1, p1};
          rc = launch_apply_in<true>(m, b.tm_in, tm_h1s, ap, e, tb, st);
        } else {
          EpiIn<__half, false> e{b.b_in, b.a1, h1, p1};
          rc = launch_apply_in<false>(m, b.tm_in, tm_h1s, ap, e, tb, st);
        }
        if (rc) return 1;
        if (stop()) return sync_failed ? 1 : 0;
        int hk = 0;  // halo box: 8 frames for d < 8, else d frames
        while ((8 << hk) < b.dilation) ++hk;
        // stash layout [tile][block][128][128]
        KaArgs ka{&b.tm_rs, &tm_hc, &tm_hh[hk], gbuf, stash + (size_t)i * (C_SKIP * TILE_M),
                  (size_t)m->n_layers * C_SKIP * TILE_M, p1,
                  p2 + (size_t)i * lo.part2_stride, fb::DwTab{b.dw_t0, b.dw_t1, b.dw_cb, b.a2, b.dilation, b.eps1}};
        const bool a01 = b.a2 >= 0.f && b.a2 <= 1.f;
        if (ka_pair) {
          KpArgs kp{&tm_hc, &tm_hh[hk], &tm_gres, &tm_stash, reinterpret_cast<const __half*>(b.w_rs_pair), last ? 1 : 0,
                    i * C_SKIP, m->n_layers * C_SKIP, p1, p2 + (size_t)i * lo.part2_stride,
                    fb::DwTab{b.dw_t0, b.dw_t1, b.dw_cb, b.a2, b.dilation, b.eps1}};
          rc = a01 ? launch_dw_rs_pair_d<true>(m, kp, tb, st) : launch_dw_rs_pair_d<false>(m, kp, tb, st);
        } else if (!last) rc = a01 ? launch_dw_rs_d<256, true>(m, ka, tb, st) : launch_dw_rs_d<256, false>(m, ka, tb, st);
        else rc = a01 ? launch_dw_rs_d<128, true>(m, ka, tb, st) : launch_dw_rs_d<128, false>(m, ka, tb, st);
        if (rc) return 1;
        if (stop()) return sync_failed ? 1 : 0;
      }
      if (m->n_layers > fb::MAX_LAYERS) return fail("fused path supports up to %d TCN blocks", fb::MAX_LAYERS);
      static const int merge_split = [] { const char* e = getenv("B200SEP_MERGE_SPLIT"); const int v = e ? atoi(e) : 2;
                                          return (v == 1 || v == 2 || v == 4 || v == 8) ? v : 2; }();
      CK(launch_k(fb::skip_merge_kernel, dim3(nt, merge_split), dim3(256), 0, st, (const __half*)stash,
                  (const float2*)p2, lo.part2_stride, (const float*)m->skip_rs, (const float*)m->skip_cb, (const float*)m->skip_eps, skipa,
                  m->alpha_mask, m->n_layers, tb));
      if (stop()) return sync_failed ? 1 : 0;
      fused_done = true;
    }
  }
  for (int i = 0; i < m->n_layers && !fused_done; ++i) {
    BlockW& b = m->blocks[i];
    const bool first = (i == 0), last = (i == m->n_layers - 1);
    {
      int rc;
      if (b.a1 >= 0.f && b.a1 <= 1.f) {
        EpiIn<T, true> e{b.b_in, b.a1, h1, p1};
        rc = run_gemm<T, 256, C_BN>(m, a_x, &tm_xa, reinterpret_cast<const T*>(b.w_in), &b.tm_in, nt, C_HID, e, tb, st);
      } else {
        EpiIn<T, false> e{b.b_in, b.a1, h1, p1};
        rc = run_gemm<T, 256, C_BN>(m, a_x, &tm_xa, reinterpret_cast<const T*>(b.w_in), &b.tm_in, nt, C_HID, e, tb, st);
      }
      if (rc) return 1;
    }
    if (stop()) return sync_failed ? 1 : 0;
    {
      DwParams p{b.g1, b.be1, b.wd, b.bd, b.a2, b.dilation, b.eps1};
      const dim3 g(DW_SLOTS, nt);
      const bool a01 = b.a2 >= 0.f && b.a2 <= 1.f;
#define DW_LAUNCH(D_)                                                                         \
  do {                                                                                        \
    if (a01) CK(launch_k(dwconv_kernel<T, D_, true>, g, dim3(256), 0, st, (const T*)h1, h2, (const float2*)p1, slots1, p2, p, tb)); \
    else CK(launch_k(dwconv_kernel<T, D_, false>, g, dim3(256), 0, st, (const T*)h1, h2, (const float2*)p1, slots1, p2, p, tb));     \
  } while (0)
      if (b.dilation == 1) DW_LAUNCH(1);
      else if (b.dilation == 2) DW_LAUNCH(2);
      else if (b.dilation == 4) DW_LAUNCH(4);
      else DW_LAUNCH(0);
#undef DW_LAUNCH
    }
    if (stop()) return sync_failed ? 1 : 0;
    {
      EpiResSkip<T> e{p2, DW_SLOTS, b.rs, b.cb, x, xa_or_null, skip, skipa, m->alpha_mask, last ? 0 : 1, first ? 1 : 0,
                      last ? 1 : 0, b.eps2};
      int rc;
      if (!last)
        rc = run_gemm<T, 256, C_HID>(m, h2, &tm_h2, reinterpret_cast<const T*>(b.w_rs), &b.tm_rs, nt, 2 * C_BN, e, tb, st);
      else
        rc = run_gemm<T, 128, C_HID>(m, h2, &tm_h2, reinterpret_cast<const T*>(b.w_rs), &b.tm_rs, nt, C_SKIP, e, tb, st);
      if (rc) return 1;
    }
    if (stop()) return sync_failed ? 1 : 0;
  }

  // 4+5 fused: mask net + mask application + decoder + overlap-add (d_out was zero-filled by the caller)
  if constexpr (kHalf) {
    if (fused_mode && m->fused_tail) {
      using Cfg = tail::TailCfg<L>;
      auto kern = tail::mask_dec_kernel<L>;
          { static DevOnce once; if (once.need(m->device)) { CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES)); once.done(m->device); } }
      tail::TailParams tp{m->mask_b, enc, d_out, c.n_src, c.mask_act};
      CK(launch_k(kern, dim3(std::min(nt, m->n_sm)), dim3(Cfg::THREADS), Cfg::SMEM_BYTES, st, tm_skipa, m->tm_mask128,
                  m->tm_wd, tm_enc, tp, tb));
      stop();
      return sync_failed ? 1 : 0;
    }
  }
  // 4. mask net + mask application
  {
    const int n_total = c.n_src * C_ENC;
    if (kHalf) {
      EpiMask<T, true> e{m->mask_b, enc, masked, n_total, c.mask_act};
      if (run_gemm<T, 256, C_SKIP>(m, skipa, &tm_skipa, reinterpret_cast<const T*>(m->mask_w), &m->tm_mask, nt, n_total,
                                   e, tb, st))
        return 1;
    } else {
      EpiMask<T, false> e{m->mask_b, enc, masked, n_total, c.mask_act};
      if (run_gemm<T, 256, C_SKIP>(m, skipa, &tm_skipa, reinterpret_cast<const T*>(m->mask_w), &m->tm_mask, nt, n_total,
                                   e, tb, st))
        return 1;
    }
  }
  if (stop()) return sync_failed ? 1 : 0;

  // 5. decoder + overlap-add (d_out was zero-filled by the caller of forward())
  {
    constexpr int kSmem = decoder_smem_bytes<L>();
    auto kern = decoder_kernel<T, L>;
      { static DevOnce once; if (once.need(m->device)) { CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem)); once.done(m->device); } }
    CK(launch_k(kern, dim3(nt, c.n_src), dim3(256), kSmem, st, (const T*)masked, (const float*)m->dec_w, d_out, c.n_src, tb));
  }
  prof_mark();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

int b200sep_version(void) { return B200SEP_VERSION; }

// Deterministic longest-processing-time-first partition of n independent segments over `world` ranks (host only; the
// evaluation sweep runs it on every rank -- 11 k segments took 13.6 ms in Python, a fixed cost on an 85 ms shard):
// segments in descending cost (ties: ascending index) each go to the least loaded rank (ties: lowest rank).
int b200sep_lpt_partition(const int64_t* costs, int32_t n, int32_t world, int32_t* owner) {
  if (n < 0 || world < 1 || (n > 0 && (!costs || !owner))) return fail("bad lpt_partition arguments");
  std::vector<int32_t> order((size_t)n);
  for (int32_t i = 0; i < n; ++i) order[(size_t)i] = i;
  std::sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return costs[a] != costs[b] ? costs[a] > costs[b] : a < b; });
  using Slot = std::pair<int64_t, int32_t>;  // (load, rank): smallest first
  std::priority_queue<Slot, std::vector<Slot>, std::greater<Slot>> heap;
  for (int32_t r = 0; r < world; ++r) heap.push({0, r});
  for (int32_t i : order) {
    Slot s = heap.top();
    heap.pop();
    owner[i] = s.second;
    heap.push({s.first + costs[i], s.second});
  }
  return 0;
}
const char* b200sep_last_error(void) { return g_err.c_str(); }

int b200sep_model_create(const b200sep_config* cfg, const b200sep_weights* w, int device, b200sep_model** out) {
  if (!cfg || !w || !out) return fail("null argument");
  *out = nullptr;
  const b200sep_config& c = *cfg;
  if (c.n_src != 2 && c.n_src != 3) return fail("n_src must be 2 or 3 (got %d)", c.n_src);
  if (c.n_filters != C_ENC || c.bn_chan != C_BN || c.hid_chan != C_HID || c.skip_chan != C_SKIP ||
      c.conv_kernel_size != 3)
    return fail("unsupported channel configuration N=%d B=%d H=%d Sc=%d P=%d (supported: 512/128/512/128/3)",
                c.n_filters, c.bn_chan, c.hid_chan, c.skip_chan, c.conv_kernel_size);
  if (!((c.kernel_size == 16 && c.stride == 8) || (c.kernel_size == 32 && c.stride == 16)))
    return fail("unsupported filterbank (kernel_size=%d, stride=%d); supported (16,8) and (32,16)", c.kernel_size,
                c.stride);
  if (c.n_blocks < 1 || c.n_blocks > 10 || c.n_repeats < 1) return fail("bad n_blocks/n_repeats");
  if (c.mask_act != B200SEP_MASK_SIGMOID && c.mask_act != B200SEP_MASK_RELU) return fail("bad mask_act");
  if (c.precision < 0 || c.precision > 2) return fail("bad precision");

  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail("no CUDA device: the b200 separation backend has no CPU fallback");
  if (device < 0 || device >= ndev) return fail("device %d out of range (%d devices)", device, ndev);
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail("device %d is sm_%d%d; this backend is built for sm_100a (B200) only", device, prop.major,
                prop.minor);
  CK(cudaSetDevice(device));

  b200sep_model* m = new b200sep_model();
  m->cfg = c;
  m->device = device;
  m->n_sm = prop.multiProcessorCount;
  m->n_layers = c.n_blocks * c.n_repeats;
  m->half = (c.precision != B200SEP_PREC_FP32);
  m->use_tc = (c.precision == B200SEP_PREC_FP16_TC);
  if (const char* f = getenv("B200SEP_FUSED")) m->fused = atoi(f) != 0;
  if (const char* f = getenv("B200SEP_FUSED_TAIL")) m->fused_tail = atoi(f) != 0;
  if (const char* f = getenv("B200SEP_KA")) m->ka_pair = strcmp(f, "pair") == 0;
  const char* dbg = getenv("B200SEP_DEBUG_STOP_AFTER");
  if (dbg) m->debug_stop_after = atoi(dbg);
  int rc = m->half ? build_weights<__half>(m, w) : build_weights<float>(m, w);
  if (!rc && m->use_tc) {
    rc = make_tmap_w(&m->tm_bn, m->bn_w, C_BN, C_ENC, 128);
    for (int i = 0; i < m->n_layers && !rc; ++i) {
      BlockW& b = m->blocks[i];
      rc = make_tmap_w(&b.tm_in, b.w_in, C_HID, C_BN, 256);
      if (!rc) rc = make_tmap_w(&b.tm_rs, b.w_rs, b.n_out, C_HID, b.n_out == 256 ? 256 : 128);
    }
    if (!rc) rc = make_tmap_w(&m->tm_mask, m->mask_w, c.n_src * C_ENC, C_SKIP, 256);
    if (!rc) rc = make_tmap_w(&m->tm_mask128, m->mask_w, c.n_src * C_ENC, C_SKIP, 128);
    if (!rc) rc = make_tmap_w(&m->tm_wd, m->dec_w16, c.kernel_size, C_ENC, c.kernel_size);
  }
  if (rc) {
    std::string keep = g_err;
    b200sep_model_destroy(m);
    g_err = keep;
    return 1;
  }
  *out = m;
  return 0;
}

void b200sep_model_destroy(b200sep_model* m) {
  if (!m) return;
  cudaSetDevice(m->device);
  for (void* p : m->allocs) cudaFree(p);
  for (cudaEvent_t e : m->prof_ev) cudaEventDestroy(e);
  for (int k = 0; k < 2; ++k) {
    for (int g = 0; g < 4; ++g) {
      if (m->aux_st2[k][g]) cudaStreamDestroy(m->aux_st2[k][g]);
      if (m->ev_join2[k][g]) cudaEventDestroy(m->ev_join2[k][g]);
    }
    if (m->ev_fork2[k]) cudaEventDestroy(m->ev_fork2[k]);
  }
  for (auto& kv : m->graphs) cudaGraphExecDestroy(kv.second);
  if (m->own_st) cudaStreamDestroy(m->own_st);
  if (m->own_ev) cudaEventDestroy(m->own_ev);
  for (int k = 0; k < b200sep_model::kHin; ++k) {
    if (m->hin_ev[k]) cudaEventDestroy(m->hin_ev[k]);
    if (m->hin[k]) cudaFreeHost(m->hin[k]);
  }
  delete m;
}

int b200sep_debug_set_stop(b200sep_model* m, int stop_after) {
  if (!m) return fail("null model");
  m->debug_stop_after = stop_after;
  return 0;
}

// Registers (or, with nullptr, removes) a device buffer of 148 * 8 * 128 int64 the fused kernels stamp with
// clock64 values per (CTA, tile visit, slot) -- see tl_stamp in fused_block.cuh.  Debug only.
int b200sep_debug_timeline(void* d_buf) {
  long long* p = static_cast<long long*>(d_buf);
  CK(cudaMemcpyToSymbol(b200sep::c_timeline, &p, sizeof(p)));
  return 0;
}

// Launch sequence the model runs: 0 = unfused (encoder, bottleneck, 3 kernels per block, mask, decoder),
// 1 = fused (encoder, bottleneck, 2 kernels per block, skip merge, mask, decoder),
// 2 = fused + fused tail (encoder, bottleneck, 2 kernels per block, skip merge, mask+decoder).
int b200sep_pipeline(const b200sep_model* m) {
  if (!(m && m->half && m->use_tc && m->fused)) return 0;
  return m->fused_tail ? 2 : 1;
}

int b200sep_profile_enable(b200sep_model* m, int on) {
  if (!m) return fail("null model");
  m->prof_on = on != 0;
  return 0;
}

// Durations (ms) of the launches of the most recent b200sep_separate_batch call made while
// profiling was enabled, in launch order: encoder, bottleneck, then per block (in-conv GEMM,
// depthwise, res/skip GEMM), mask GEMM, decoder.  Synchronises on the last launch.
int b200sep_profile_read(b200sep_model* m, float* ms_out, int32_t cap, int32_t* n_out) {
  if (!m || !ms_out || !n_out) return fail("null argument");
  std::lock_guard<std::mutex> lock(m->mu);
  const int n = m->prof_n - 1;
  *n_out = n > 0 ? n : 0;
  if (n <= 0) return 0;
  CK(cudaEventSynchronize(m->prof_ev[m->prof_n - 1]));
  for (int i = 0; i < n && i < cap; ++i) CK(cudaEventElapsedTime(&ms_out[i], m->prof_ev[i], m->prof_ev[i + 1]));
  return 0;
}

// Byte offsets of the intermediate buffers inside the workspace (tests read them back):
// out[0..9] = enc, x, xa, skip, skipa, h1, h2, masked, n_tiles, total
int b200sep_debug_layout(const b200sep_model* m, const int32_t* lens, int32_t n_seg, uint64_t* out) {
  if (!m || !out) return fail("null argument");
  Layout lo;
  if (compute_layout(m->cfg, m->half, m->half && m->use_tc && m->fused, lens, n_seg, &lo)) return 1;
  out[0] = lo.o_enc; out[1] = lo.o_x; out[2] = lo.o_xa; out[3] = lo.o_skip; out[4] = lo.o_skipa;
  out[5] = lo.o_h1; out[6] = lo.o_h2; out[7] = lo.o_masked; out[8] = (uint64_t)lo.n_tiles; out[9] = lo.total;
  out[10] = lo.o_g;
  return 0;
}

// ---- concurrent sub-batches -------------------------------------------------------------------------------------
// Every kernel of the forward is a persistent grid of one CTA per SM, so a batch of n m-tiles costs
// ceil(n / n_sm) tile-times per launch (1050 tiles on 148 SMs: 8 rounds for 7.09 rounds of work), and its two
// block kernels are bound by different units (K_B: HBM, K_A: shared memory).  Segments are independent, so a large
// batch is cut into `split` tile-balanced sub-batches that run the whole forward on separate streams: the CTAs of
// one sub-batch's launch fill the SMs the other's launch has already drained, and K_A / K_B of different
// sub-batches overlap.  Results do not change (a segment's arithmetic never depends on its batch neighbours).
static int split_env() {
  static int v = -2;
  if (v == -2) {
    const char* e = getenv("B200SEP_SPLIT");
    v = e ? atoi(e) : -1;
  }
  return v;
}
static constexpr int MAX_SPLIT = B200SEP_MAX_SPLIT;
// Deterministic LPT partition of the segments into tile-balanced groups (original order kept inside a group).
static int plan_split(const b200sep_model* m, const int32_t* lens, int n_seg, std::vector<std::vector<int>>* groups) {
  long long tiles = 0;
  std::vector<int> nt(n_seg);
  for (int i = 0; i < n_seg; ++i) {
    if (lens[i] < m->cfg.kernel_size)
      return fail("segment %d has %d samples < kernel_size %d (conv1d would fail in the reference too)", i, lens[i],
                  m->cfg.kernel_size);
    const int F = (lens[i] - m->cfg.kernel_size) / m->cfg.stride + 1;
    nt[i] = (F + TILE_M - 1) / TILE_M;
    tiles += nt[i];
  }
  int S = 1;
  const bool eligible = m->half && m->use_tc && m->fused && !m->prof_on && m->debug_stop_after == 0 &&
                        getenv("B200SEP_SYNC_DEBUG") == nullptr;
  if (eligible) {
    const int e = split_env();
    if (e >= 1) S = std::min(e, MAX_SPLIT);
    else S = (tiles >= 3ll * m->n_sm) ? 2 : 1;  // default: two sub-batches once each still fills the GPU
    const int cap = m->max_split.load();
    if (cap >= 1) S = std::min(S, cap);
    S = std::max(1, std::min(S, n_seg));
  }
  groups->assign(S, {});
  if (S == 1) {
    (*groups)[0].resize(n_seg);
    for (int i = 0; i < n_seg; ++i) (*groups)[0][i] = i;
    return 0;
  }
  std::vector<int> order(n_seg);
  for (int i = 0; i < n_seg; ++i) order[i] = i;
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return nt[a] > nt[b]; });
  // target shares: equal, unless B200SEP_SPLIT_RATIO=<share of group 0> (experiments: a smaller first group finishes
  // earlier, so its device -> host copy overlaps the other group's last kernels)
  static const double ratio0 = [] { const char* e = getenv("B200SEP_SPLIT_RATIO"); return e ? atof(e) : 0.0; }();
  std::vector<double> share(S, 1.0 / S);
  if (S == 2 && ratio0 > 0.05 && ratio0 < 0.95) { share[0] = ratio0; share[1] = 1.0 - ratio0; }
  std::vector<long long> load(S, 0);
  for (int i : order) {
    int best = 0;
    for (int g = 1; g < S; ++g)
      if ((double)load[g] / share[g] < (double)load[best] / share[best]) best = g;
    (*groups)[best].push_back(i);
    load[best] += nt[i];
  }
  for (auto& g : *groups) std::sort(g.begin(), g.end());
  return 0;
}

size_t b200sep_workspace_bytes(const b200sep_model* m, const int32_t* lens, int32_t n_seg) {
  if (!m) { fail("null model"); return 0; }
  if (n_seg <= 0 || !lens) { fail("empty batch"); return 0; }
  // the larger of the unsplit and the split plan (the split decision may differ between this call and the run,
  // e.g. when profiling is switched on in between)
  Layout lo;
  if (compute_layout(m->cfg, m->half, m->half && m->use_tc && m->fused, lens, n_seg, &lo)) return 0;
  size_t total = lo.total;
  std::vector<std::vector<int>> groups;
  if (plan_split(m, lens, n_seg, &groups)) return 0;
  if (groups.size() > 1) {
    size_t sum = 0;
    for (auto& g : groups) {
      std::vector<int32_t> sl(g.size());
      for (size_t k = 0; k < g.size(); ++k) sl[k] = lens[g[k]];
      if (compute_layout(m->cfg, m->half, true, sl.data(), (int)sl.size(), &lo)) return 0;
      sum += align_up(lo.total, 1024);
    }
    total = std::max(total, sum);
  }
  return total;
}

// Cap on the concurrent sub-batches a call is cut into (0 restores the automatic choice).  A caller that already keeps
// several batches in flight on streams of its own (Separator.separate_stream / score_stream) sets 1: the batches fill
// each other's launch tails, and cutting each of them again only adds launches, host threads and pipeline fills
// (measured end to end, two batches in flight: 38.8 k audio-s/s unsplit against 36.3 k split).
int b200sep_set_max_split(b200sep_model* m, int32_t n) {
  if (!m || n < 0) return fail("bad set_max_split argument");
  m->max_split.store((int)n);
  return 0;
}

int b200sep_split_count(const b200sep_model* m, const int32_t* lens, int32_t n_seg) {
  if (!m || !lens || n_seg <= 0) { fail("null / empty argument"); return -1; }
  std::vector<std::vector<int>> groups;
  if (plan_split(m, lens, n_seg, &groups)) return -1;
  return (int)groups.size();
}

// One forward over the segments (offsets[i], lens[i]) on stream `st`, workspace `ws`.  The caller has zero-filled the
// output span.  Error text is left in this thread's g_err.
static int run_sub_launch(b200sep_model* m, const Layout& lo, const Tables& tb, const float* d_mix, float* d_out, uint8_t* ws,
                          int n_seg, cudaStream_t st);

static int run_sub(b200sep_model* m, const float* d_mix, const int64_t* offsets, const int32_t* lens, int n_seg,
                   float* d_out, uint8_t* ws, cudaStream_t st, int32_t* h_flag) {
  Layout lo;
  if (compute_layout(m->cfg, m->half, m->half && m->use_tc && m->fused, lens, n_seg, &lo)) return 1;
  CK(cudaSetDevice(m->device));
  // ---- tables -> pinned staging -> workspace head
  StagingSlot* slot = nullptr;
  {
    std::lock_guard<std::mutex> l2(m->staging.mu);
    if (m->staging.acquire(lo.tables_bytes, &slot)) return 1;
  }
  uint8_t* h = static_cast<uint8_t*>(slot->host);
  memset(h, 0, lo.tables_bytes);
  int64_t* h_in = reinterpret_cast<int64_t*>(h + lo.t_in_off);
  int64_t* h_out = reinterpret_cast<int64_t*>(h + lo.t_out_off);
  double* h_inv = reinterpret_cast<double*>(h + lo.t_inv);
  int32_t* h_len = reinterpret_cast<int32_t*>(h + lo.t_len);
  int32_t* h_fr = reinterpret_cast<int32_t*>(h + lo.t_frames);
  int32_t* h_t0 = reinterpret_cast<int32_t*>(h + lo.t_tile0);
  int32_t* h_nt = reinterpret_cast<int32_t*>(h + lo.t_ntiles);
  int32_t* h_ts = reinterpret_cast<int32_t*>(h + lo.t_tile_seg);
  TileInfo* h_ti = reinterpret_cast<TileInfo*>(h + lo.t_info);
  int tile = 0;
  for (int i = 0; i < n_seg; ++i) {
    h_in[i] = offsets[i];
    h_out[i] = offsets[i] * m->cfg.n_src;
    h_len[i] = lens[i];
    const int F = (lens[i] - m->cfg.kernel_size) / m->cfg.stride + 1;
    h_fr[i] = F;
    h_inv[i] = 1.0 / ((double)C_HID * (double)F);
    h_t0[i] = tile;
    const int ntile = (F + TILE_M - 1) / TILE_M;
    h_nt[i] = ntile;
    for (int t = 0; t < ntile; ++t) {
      h_ts[tile + t] = i;
      TileInfo& ti = h_ti[tile + t];
      ti.seg = i; ti.t0 = t * TILE_M; ti.F = F; ti.tile0 = tile; ti.ntiles = ntile; ti.pad = 0;
      ti.inv_count = h_inv[i];
    }
    tile += ntile;
  }
  CK(cudaMemcpyAsync(ws + lo.o_tables, h, lo.tables_bytes, cudaMemcpyHostToDevice, st));
  CK(cudaEventRecord(slot->ev, st));

  Tables tb;
  uint8_t* t = ws + lo.o_tables;
  tb.tile_info = reinterpret_cast<const TileInfo*>(t + lo.t_info);
  tb.seg_in_off = reinterpret_cast<const int64_t*>(t + lo.t_in_off);
  tb.seg_out_off = reinterpret_cast<const int64_t*>(t + lo.t_out_off);
  tb.seg_inv_count = reinterpret_cast<const double*>(t + lo.t_inv);
  tb.seg_len = reinterpret_cast<const int32_t*>(t + lo.t_len);
  tb.seg_frames = reinterpret_cast<const int32_t*>(t + lo.t_frames);
  tb.seg_tile0 = reinterpret_cast<const int32_t*>(t + lo.t_tile0);
  tb.seg_ntiles = reinterpret_cast<const int32_t*>(t + lo.t_ntiles);
  tb.tile_seg = reinterpret_cast<const int32_t*>(t + lo.t_tile_seg);
  tb.flags = reinterpret_cast<int*>(t + lo.t_flags);
  tb.n_seg = n_seg;
  tb.n_tiles = lo.n_tiles;

  if (run_sub_launch(m, lo, tb, d_mix, d_out, ws, n_seg, st)) return 1;
  // the call's device flags -> the caller's (pinned) host word, readable once the stream has been synchronised
  if (h_flag) CK(cudaMemcpyAsync(h_flag, tb.flags, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  return 0;
}

static int run_sub_launch(b200sep_model* m, const Layout& lo, const Tables& tb, const float* d_mix, float* d_out, uint8_t* ws,
                          int n_seg, cudaStream_t st) {
  const bool l16 = (m->cfg.kernel_size == 16);
  auto run_forward = [&]() -> int {
    if (m->half)
      return l16 ? forward<__half, 16>(m, lo, tb, d_mix, d_out, ws, st) : forward<__half, 32>(m, lo, tb, d_mix, d_out, ws, st);
    return l16 ? forward<float, 16>(m, lo, tb, d_mix, d_out, ws, st) : forward<float, 32>(m, lo, tb, d_mix, d_out, ws, st);
  };
  // Small batches (at most one m-tile per SM) replay a captured graph of the launch sequence: every kernel argument is
  // a function of (n_seg, n_tiles, buffer addresses) -- the per-call data (segment lengths, offsets) lives in the layout
  // tables uploaded above, outside the graph.
  static const bool graphs_on = [] { const char* e = getenv("B200SEP_GRAPH"); return !(e && atoi(e) == 0); }();
  const bool use_graph = graphs_on && st != nullptr && !m->prof_on && m->debug_stop_after == 0 && lo.n_tiles <= m->n_sm &&
                         getenv("B200SEP_SYNC_DEBUG") == nullptr;
  if (!use_graph || (m->graph_misses > 64 && m->graph_hits < m->graph_misses)) return run_forward();
  {  // a caller that is itself capturing this stream gets the plain launch sequence recorded into ITS graph
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) {
      (void)cudaGetLastError();
      return run_forward();
    }
  }
  std::lock_guard<std::mutex> glock(m->graph_mu);  // sub-batch worker threads share the cache
  const b200sep_model::GraphKey key{n_seg, lo.n_tiles, d_mix, d_out, ws};
  auto it = m->graphs.find(key);
  if (it != m->graphs.end()) ++m->graph_hits;
  else {
    ++m->graph_misses;
    CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    const int rc = run_forward();
    const std::string keep = g_err;
    cudaGraph_t g = nullptr;
    const cudaError_t e = cudaStreamEndCapture(st, &g);
    if (rc != 0 || e != cudaSuccess || g == nullptr) {
      if (g) cudaGraphDestroy(g);
      (void)cudaGetLastError();
      if (rc != 0) return fail("%s", keep.c_str());
      return fail("graph capture of the launch sequence failed: %s", cudaGetErrorString(e));
    }
    cudaGraphExec_t ex = nullptr;
    const cudaError_t ie = cudaGraphInstantiate(&ex, g, 0);
    cudaGraphDestroy(g);
    if (ie != cudaSuccess) return fail("cudaGraphInstantiate failed: %s", cudaGetErrorString(ie));
    if (m->graphs.size() >= 256) {  // bounded cache: drop everything rather than track recency
      for (auto& kv : m->graphs) cudaGraphExecDestroy(kv.second);
      m->graphs.clear();
    }
    it = m->graphs.emplace(key, ex).first;
  }
  CK(cudaGraphLaunch(it->second, st));
  return 0;
}

// A caller on the legacy default stream (NULL) is moved onto a stream of the model for the duration of the call,
// ordered after / before the default stream by events: stream capture (the graphs of run_sub) is not possible on the
// legacy stream.
struct StreamHop {
  b200sep_model* m;
  cudaStream_t user, st;
  int enter() {
    st = user;
    if (user != nullptr) return 0;
    if (!m->own_st) CK(cudaStreamCreateWithFlags(&m->own_st, cudaStreamNonBlocking));
    if (!m->own_ev) CK(cudaEventCreateWithFlags(&m->own_ev, cudaEventDisableTiming));
    CK(cudaEventRecord(m->own_ev, user));
    CK(cudaStreamWaitEvent(m->own_st, m->own_ev, 0));
    st = m->own_st;
    return 0;
  }
  int leave() {
    if (user != nullptr) return 0;
    CK(cudaEventRecord(m->own_ev, m->own_st));
    CK(cudaStreamWaitEvent(user, m->own_ev, 0));
    return 0;
  }
};

static int separate_batch_locked(b200sep_model* m, const float* d_mix, const int64_t* offsets, const int32_t* lens,
                                 int32_t n_seg, float* d_out, void* d_ws, size_t ws_bytes, cudaStream_t st,
                                 int32_t* h_flags);

int b200sep_separate_batch(b200sep_model* m, const float* d_mix, const int64_t* offsets, const int32_t* lens,
                           int32_t n_seg, float* d_out, void* d_ws, size_t ws_bytes, void* stream, int32_t* h_flags) {
  if (!m || !d_mix || !lens || !d_out || !d_ws) return fail("null argument");
  if ((reinterpret_cast<uintptr_t>(d_ws) & 255u) != 0) return fail("d_ws must be 256-byte aligned");
  if (n_seg <= 0) return fail("empty batch");
  std::lock_guard<std::mutex> lock(m->mu);
  CK(cudaSetDevice(m->device));
  StreamHop hop{m, reinterpret_cast<cudaStream_t>(stream), nullptr};
  if (hop.enter()) return 1;
  if (h_flags) for (int g = 0; g < B200SEP_MAX_SPLIT; ++g) h_flags[g] = 0;
  const int rc = separate_batch_locked(m, d_mix, offsets, lens, n_seg, d_out, d_ws, ws_bytes, hop.st, h_flags);
  const std::string keep = g_err;
  if (hop.leave()) return 1;
  if (rc) return fail("%s", keep.c_str());
  return 0;
}

static int separate_batch_locked(b200sep_model* m, const float* d_mix, const int64_t* offsets, const int32_t* lens,
                                 int32_t n_seg, float* d_out, void* d_ws, size_t ws_bytes, cudaStream_t st,
                                 int32_t* h_flags) {
  cudaStream_t* aux_st = m->aux_st2[0];
  cudaEvent_t* ev_join = m->ev_join2[0];
  cudaEvent_t& ev_fork = m->ev_fork2[0];
  std::vector<std::vector<int>> groups;
  if (plan_split(m, lens, n_seg, &groups)) return 1;
  const int S = (int)groups.size();

  // absolute offsets + the output span
  std::vector<int64_t> off(n_seg);
  int64_t run = 0, out_lo = INT64_MAX, out_hi = 0;
  for (int i = 0; i < n_seg; ++i) {
    off[i] = offsets ? offsets[i] : run;
    if (off[i] < 0) return fail("negative offset for segment %d", i);
    run += lens[i];
    out_lo = std::min(out_lo, off[i] * m->cfg.n_src);
    out_hi = std::max(out_hi, (off[i] + (int64_t)lens[i]) * m->cfg.n_src);
  }
  // per-group segment lists and workspace slices
  std::vector<std::vector<int32_t>> g_len(S);
  std::vector<std::vector<int64_t>> g_off(S);
  std::vector<size_t> g_ws(S);
  size_t need = 0;
  for (int g = 0; g < S; ++g) {
    for (int i : groups[g]) { g_len[g].push_back(lens[i]); g_off[g].push_back(off[i]); }
    Layout lo;
    if (compute_layout(m->cfg, m->half, m->half && m->use_tc && m->fused, g_len[g].data(), (int)g_len[g].size(), &lo))
      return 1;
    g_ws[g] = need;
    need += align_up(lo.total, 1024);
  }
  if (ws_bytes < need) return fail("workspace too small: %zu < %zu bytes", ws_bytes, need);
  uint8_t* ws = static_cast<uint8_t*>(d_ws);
  // zero-fill the output span (overlap-add boundary hops accumulate atomically; pad_x_to_y tail)
  CK(cudaMemsetAsync(d_out + out_lo, 0, (size_t)(out_hi - out_lo) * sizeof(float), st));
  if (S == 1) return run_sub(m, d_mix, g_off[0].data(), g_len[0].data(), n_seg, d_out, ws, st, h_flags);

  // fork: sub-batch 0 on the caller's stream, the others on the model's auxiliary streams, launches issued by one
  // host thread per sub-batch so that the streams fill at the same pace
  if (!ev_fork) CK(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
  for (int g = 1; g < S; ++g) {
    if (!aux_st[g]) CK(cudaStreamCreateWithFlags(&aux_st[g], cudaStreamNonBlocking));
    if (!ev_join[g]) CK(cudaEventCreateWithFlags(&ev_join[g], cudaEventDisableTiming));
  }
  CK(cudaEventRecord(ev_fork, st));
  std::vector<int> rc(S, 0);
  std::vector<std::string> err(S);
  std::vector<std::thread> workers;
  for (int g = 1; g < S; ++g) {
    CK(cudaStreamWaitEvent(aux_st[g], ev_fork, 0));
    workers.emplace_back([&, g]() {
      rc[g] = run_sub(m, d_mix, g_off[g].data(), g_len[g].data(), (int)g_len[g].size(), d_out, ws + g_ws[g], aux_st[g],
                      h_flags ? h_flags + g : nullptr);
      if (rc[g]) err[g] = g_err;
    });
  }
  rc[0] = run_sub(m, d_mix, g_off[0].data(), g_len[0].data(), (int)g_len[0].size(), d_out, ws + g_ws[0], st, h_flags);
  if (rc[0]) err[0] = g_err;
  for (auto& w : workers) w.join();
  // join (also on failure: the caller's stream must not run ahead of work already queued on the auxiliary streams)
  for (int g = 1; g < S; ++g) {
    CK(cudaEventRecord(ev_join[g], aux_st[g]));
    CK(cudaStreamWaitEvent(st, ev_join[g], 0));
  }
  for (int g = 0; g < S; ++g)
    if (rc[g]) return fail("%s", err[g].c_str());
  return 0;
}

// ---- host boundary: pageable / pinned host segments in, host waveforms out -------------------------------------
// Device-side extras of the host entry point: the packed mixtures and the packed outputs live behind the forward's
// own scratch.
static size_t host_extra_bytes(const b200sep_model* m, const int32_t* lens, int n_seg, size_t* o_mix, size_t* o_out) {
  size_t total = 0;
  for (int i = 0; i < n_seg; ++i) total += (size_t)lens[i];
  const size_t mix_b = align_up(total * 4, 1024), out_b = align_up(total * 4 * m->cfg.n_src, 1024);
  if (o_mix) *o_mix = 0;
  if (o_out) *o_out = mix_b;
  return mix_b + out_b;
}
size_t b200sep_host_workspace_bytes(const b200sep_model* m, const int32_t* lens, int32_t n_seg) {
  const size_t fw = b200sep_workspace_bytes(m, lens, n_seg);
  if (fw == 0) return 0;
  return align_up(fw, 1024) + host_extra_bytes(m, lens, n_seg, nullptr, nullptr);
}

static int separate_batch_host_locked(b200sep_model* m, const float* const* h_segs, const int32_t* lens, int32_t n_seg,
                                      float* h_out, void* d_ws, size_t ws_bytes, cudaStream_t st, int32_t* h_flags);

int b200sep_separate_batch_host(b200sep_model* m, const float* const* h_segs, const int32_t* lens, int32_t n_seg,
                                float* h_out, void* d_ws, size_t ws_bytes, void* stream, int32_t* h_flags) {
  if (!m || !h_segs || !lens || !h_out || !d_ws) return fail("null argument");
  if ((reinterpret_cast<uintptr_t>(d_ws) & 255u) != 0) return fail("d_ws must be 256-byte aligned");
  if (n_seg <= 0) return fail("empty batch");
  std::lock_guard<std::mutex> lock(m->mu);
  CK(cudaSetDevice(m->device));
  StreamHop hop{m, reinterpret_cast<cudaStream_t>(stream), nullptr};
  if (hop.enter()) return 1;
  if (h_flags) for (int g = 0; g < B200SEP_MAX_SPLIT; ++g) h_flags[g] = 0;
  const int rc = separate_batch_host_locked(m, h_segs, lens, n_seg, h_out, d_ws, ws_bytes, hop.st, h_flags);
  const std::string keep = g_err;
  if (hop.leave()) return 1;
  if (rc) return fail("%s", keep.c_str());
  return 0;
}

static int separate_batch_host_locked(b200sep_model* m, const float* const* h_segs, const int32_t* lens, int32_t n_seg,
                                      float* h_out, void* d_ws, size_t ws_bytes, cudaStream_t st, int32_t* h_flags) {
  std::vector<std::vector<int>> groups;
  if (plan_split(m, lens, n_seg, &groups)) return 1;
  const int S = (int)groups.size(), NS = m->cfg.n_src;

  // device packing: group-contiguous (each group's H2D is one copy); the host output keeps the caller's order
  std::vector<int64_t> h_off(n_seg), d_off(n_seg);
  int64_t total = 0;
  for (int i = 0; i < n_seg; ++i) { h_off[i] = total; total += lens[i]; }
  std::vector<std::vector<int32_t>> g_len(S);
  std::vector<std::vector<int64_t>> g_off(S);
  std::vector<int64_t> g_begin(S), g_end(S);
  std::vector<size_t> g_ws(S);
  size_t fw_need = 0;
  int64_t run = 0;
  for (int g = 0; g < S; ++g) {
    g_begin[g] = run;
    for (int i : groups[g]) {
      d_off[i] = run;
      g_len[g].push_back(lens[i]);
      g_off[g].push_back(run);
      run += lens[i];
    }
    g_end[g] = run;
    Layout lo;
    if (compute_layout(m->cfg, m->half, m->half && m->use_tc && m->fused, g_len[g].data(), (int)g_len[g].size(), &lo))
      return 1;
    g_ws[g] = fw_need;
    fw_need += align_up(lo.total, 1024);
  }
  size_t o_mix, o_out;
  const size_t extra = host_extra_bytes(m, lens, n_seg, &o_mix, &o_out);
  if (ws_bytes < fw_need + extra) return fail("workspace too small: %zu < %zu bytes", ws_bytes, fw_need + extra);
  uint8_t* ws = static_cast<uint8_t*>(d_ws);
  float* d_mix = reinterpret_cast<float*>(ws + fw_need + o_mix);
  float* d_out = reinterpret_cast<float*>(ws + fw_need + o_out);

  // pinned input staging (grown on demand); its previous H2D copies must have completed before it is rewritten
  const int hs = m->hin_next;
  m->hin_next = (hs + 1) % b200sep_model::kHin;
  if (m->hin_cap[hs] < (size_t)total) {
    if (m->hin_ev[hs]) CK(cudaEventSynchronize(m->hin_ev[hs]));
    if (m->hin[hs]) CK(cudaFreeHost(m->hin[hs]));
    m->hin[hs] = nullptr; m->hin_cap[hs] = 0;
    const size_t cap = std::max<size_t>((size_t)total * 2, 1 << 16);
    CK(cudaHostAlloc(reinterpret_cast<void**>(&m->hin[hs]), cap * sizeof(float), cudaHostAllocDefault));
    m->hin_cap[hs] = cap;
  }
  if (!m->hin_ev[hs]) CK(cudaEventCreateWithFlags(&m->hin_ev[hs], cudaEventDisableTiming));
  else CK(cudaEventSynchronize(m->hin_ev[hs]));  // the call that used this slot two calls ago has copied its input
  float* const hin = m->hin[hs];
  cudaStream_t* aux_st = m->aux_st2[hs];
  cudaEvent_t* ev_join = m->ev_join2[hs];
  cudaEvent_t& ev_fork = m->ev_fork2[hs];

  CK(cudaMemsetAsync(d_out, 0, (size_t)total * NS * sizeof(float), st));
  if (!ev_fork) CK(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
  for (int g = 1; g < S; ++g) {
    if (!aux_st[g]) CK(cudaStreamCreateWithFlags(&aux_st[g], cudaStreamNonBlocking));
    if (!ev_join[g]) CK(cudaEventCreateWithFlags(&ev_join[g], cudaEventDisableTiming));
  }
  CK(cudaEventRecord(ev_fork, st));
  for (int g = 1; g < S; ++g) CK(cudaStreamWaitEvent(aux_st[g], ev_fork, 0));

  // per group: gather -> H2D -> forward -> D2H, everything of a group on its own stream and host thread, so the
  // gather / copies of one group overlap the kernels of the other
  auto run_group = [&](int g, cudaStream_t gs) -> int {
    CK(cudaSetDevice(m->device));
    for (int i : groups[g]) memcpy(hin + d_off[i], h_segs[i], (size_t)lens[i] * sizeof(float));
    CK(cudaMemcpyAsync(d_mix + g_begin[g], hin + g_begin[g], (size_t)(g_end[g] - g_begin[g]) * sizeof(float),
                       cudaMemcpyHostToDevice, gs));
    if (run_sub(m, d_mix, g_off[g].data(), g_len[g].data(), (int)g_len[g].size(), d_out, ws + g_ws[g], gs,
                h_flags ? h_flags + g : nullptr))
      return 1;
    // D2H: maximal runs of segments that are adjacent on both sides
    const std::vector<int>& idx = groups[g];
    size_t k = 0;
    while (k < idx.size()) {
      size_t e = k + 1;
      while (e < idx.size() && idx[e] == idx[e - 1] + 1) ++e;
      const int i0 = idx[k];
      int64_t n = 0;
      for (size_t q = k; q < e; ++q) n += lens[idx[q]];
      CK(cudaMemcpyAsync(h_out + (size_t)NS * h_off[i0], d_out + (size_t)NS * d_off[i0], (size_t)n * NS * sizeof(float),
                         cudaMemcpyDeviceToHost, gs));
      k = e;
    }
    return 0;
  };
  std::vector<int> rc(S, 0);
  std::vector<std::string> err(S);
  std::vector<std::thread> workers;
  for (int g = 1; g < S; ++g)
    workers.emplace_back([&, g]() {
      rc[g] = run_group(g, aux_st[g]);
      if (rc[g]) err[g] = g_err;
    });
  rc[0] = run_group(0, st);
  if (rc[0]) err[0] = g_err;
  for (auto& w : workers) w.join();
  for (int g = 1; g < S; ++g) {
    CK(cudaEventRecord(ev_join[g], aux_st[g]));
    CK(cudaStreamWaitEvent(st, ev_join[g], 0));
  }
  CK(cudaEventRecord(m->hin_ev[hs], st));
  for (int g = 0; g < S; ++g)
    if (rc[g]) return fail("%s", err[g].c_str());
  return 0;
}

// ---------------------------------------------------------------------------------------------- PIT
static int pit_geometry(const int32_t* lens, int n_seg, int* chunk, int* max_chunks) {
  if (n_seg <= 0 || !lens) return fail("empty batch");
  int max_len = 0;
  for (int i = 0; i < n_seg; ++i) {
    if (lens[i] < 0) return fail("negative length");
    max_len = std::max(max_len, lens[i]);
  }
  // chunk = samples per CTA of the moments kernel: small for small batches (parallelism), large for big sweeps (the
  // 26-value fp64 block reduction at the end of a CTA costs as much as streaming ~10 k samples)
  int ch = 4096;
  while ((max_len + ch - 1) / ch > 64) ch *= 2;
  while ((long long)n_seg * ((max_len + ch - 1) / ch) > 4096 && ch < 65536 && ch < max_len) ch *= 2;
  *chunk = ch;
  *max_chunks = std::max(1, (max_len + ch - 1) / ch);
  return 0;
}

size_t b200sep_pit_workspace_bytes_n(const int32_t* lens, int32_t n_seg, int32_t N) {
  int chunk, max_chunks;
  if (pit_geometry(lens, n_seg, &chunk, &max_chunks)) return 0;
  if (N < 1 || N > PIT_MAX_N) { fail("PIT supports at most %d estimates per segment (got %d)", PIT_MAX_N, N); return 0; }
  const size_t tables = align_up((size_t)n_seg * 8, 256) + align_up((size_t)n_seg * 4, 256);
  const size_t groups = (size_t)((N + PIT_GROUP - 1) / PIT_GROUP);
  return tables + groups * (size_t)n_seg * max_chunks * PIT_NACC * sizeof(double);
}
size_t b200sep_pit_workspace_bytes(const int32_t* lens, int32_t n_seg) {
  return b200sep_pit_workspace_bytes_n(lens, n_seg, 3);
}

int b200sep_pit_sisdr_batch(const float* d_refs, const float* d_ests, const float* d_mix, const int64_t* offsets,
                            const int32_t* lens, int32_t n_seg, int32_t K, int32_t N, int32_t variant, double* d_pair,
                            double* d_sisdr, double* d_sisdri, int32_t* d_perm, double* d_gap, void* d_ws,
                            size_t ws_bytes, void* stream) {
  if (!d_refs || !d_ests || !d_mix || !lens || !d_pair || !d_sisdr || !d_sisdri || !d_perm || !d_ws)
    return fail("null argument");
  if (!(K == 2 || K == 3) || N < K || N > PIT_MAX_N)
    return fail("PIT supports K in {2,3}, K <= N <= %d (got K=%d N=%d)", PIT_MAX_N, K, N);
  if (variant != B200SEP_PIT_OVERLAP3 && variant != B200SEP_PIT_EVALSRC) return fail("bad PIT variant");
  if ((reinterpret_cast<uintptr_t>(d_ws) & 255u) != 0) return fail("d_ws must be 256-byte aligned");
  int chunk, max_chunks;
  if (pit_geometry(lens, n_seg, &chunk, &max_chunks)) return 1;
  const size_t need = b200sep_pit_workspace_bytes_n(lens, n_seg, N);
  if (ws_bytes < need) return fail("PIT workspace too small: %zu < %zu", ws_bytes, need);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  std::lock_guard<std::mutex> lock(g_pit_mu);

  const size_t off_bytes = align_up((size_t)n_seg * 8, 256), len_bytes = align_up((size_t)n_seg * 4, 256);
  StagingSlot* slot = nullptr;
  if (g_pit_staging[current_device_index()].acquire(off_bytes + len_bytes, &slot)) return 1;
  uint8_t* h = static_cast<uint8_t*>(slot->host);
  int64_t* h_off = reinterpret_cast<int64_t*>(h);
  int32_t* h_len = reinterpret_cast<int32_t*>(h + off_bytes);
  int64_t run = 0;
  for (int i = 0; i < n_seg; ++i) {
    h_off[i] = offsets ? offsets[i] : run;
    run += lens[i];
    h_len[i] = lens[i];
  }
  uint8_t* ws = static_cast<uint8_t*>(d_ws);
  CK(cudaMemcpyAsync(ws, h, off_bytes + len_bytes, cudaMemcpyHostToDevice, st));
  CK(cudaEventRecord(slot->ev, st));
  double* partial = reinterpret_cast<double*>(ws + off_bytes + len_bytes);

  PitArgs a{d_refs, d_ests, d_mix, reinterpret_cast<const int64_t*>(ws), reinterpret_cast<const int32_t*>(ws + off_bytes),
            K, N, chunk, max_chunks, N, 0};
  dim3 grid(max_chunks, n_seg);
  PitProf* prof = nullptr;
  if (g_pit_prof_on.load() != 0) {
    prof = &g_pit_prof[current_device_index()];
    for (auto& e : prof->ev)
      if (e == nullptr) CK(cudaEventCreate(&e));
    CK(cudaEventRecord(prof->ev[0], st));
  }
  // one pass per group of <= 3 estimates (N <= 3, every reference call site: a single pass, each sample read once)
  const size_t group_stride = (size_t)n_seg * max_chunks * PIT_NACC;
  for (int g = 0; g * PIT_GROUP < N; ++g) {
    PitArgs ag = a;
    ag.j0 = g * PIT_GROUP;
    ag.N = std::min(PIT_GROUP, N - ag.j0);
    double* pg = partial + (size_t)g * group_stride;
    const int key = K * 10 + ag.N;
    switch (key) {
      case 21: pit_moments_kernel<2, 1><<<grid, 256, 0, st>>>(ag, pg); break;
      case 22: pit_moments_kernel<2, 2><<<grid, 256, 0, st>>>(ag, pg); break;
      case 23: pit_moments_kernel<2, 3><<<grid, 256, 0, st>>>(ag, pg); break;
      case 31: pit_moments_kernel<3, 1><<<grid, 256, 0, st>>>(ag, pg); break;
      case 32: pit_moments_kernel<3, 2><<<grid, 256, 0, st>>>(ag, pg); break;
      default: pit_moments_kernel<3, 3><<<grid, 256, 0, st>>>(ag, pg); break;
    }
    CK(cudaGetLastError());
  }
  if (prof) CK(cudaEventRecord(prof->ev[1], st));
  pit_finalize_kernel<<<(n_seg + 127) / 128, 128, 0, st>>>(partial, a.len, n_seg, max_chunks, K, N, variant, d_pair,
                                                         d_sisdr, d_sisdri, d_perm, d_gap);
  CK(cudaGetLastError());
  if (prof) {
    CK(cudaEventRecord(prof->ev[2], st));
    prof->valid = true;
  }
  return 0;
}

int b200sep_pit_profile_enable(int on) {
  g_pit_prof_on.store(on ? 1 : 0);
  return 0;
}

int b200sep_pit_profile_read(float* ms_out) {
  if (!ms_out) return fail("null argument");
  std::lock_guard<std::mutex> lock(g_pit_mu);
  PitProf& p = g_pit_prof[current_device_index()];
  if (!p.valid) return fail("no profiled b200sep_pit_sisdr_batch call on this device");
  CK(cudaEventSynchronize(p.ev[2]));
  CK(cudaEventElapsedTime(&ms_out[0], p.ev[0], p.ev[1]));
  CK(cudaEventElapsedTime(&ms_out[1], p.ev[1], p.ev[2]));
  return 0;
}

// ---------------------------------------------------------------------------------------------- aux ops
// Ragged-batch tables of the auxiliary kernels: pinned staging slot + a device mirror of the same size, recycled
// through an 8-deep ring (a slot is reused once the copy / kernels recorded on its event have completed).
struct AuxSlot {
  void* host = nullptr;
  void* dev = nullptr;
  size_t cap = 0;
  cudaEvent_t ev = nullptr;
};
static AuxSlot g_aux[kMaxDev][8];  // per device: the device mirror and the event live on one device
static int g_aux_next[kMaxDev] = {0};
static std::mutex g_aux_mu;
static int aux_acquire(size_t bytes, AuxSlot** out) {
  const int dev = current_device_index();
  AuxSlot& s = g_aux[dev][g_aux_next[dev]];
  g_aux_next[dev] = (g_aux_next[dev] + 1) % 8;
  if (!s.ev) CK(cudaEventCreateWithFlags(&s.ev, cudaEventDisableTiming));
  else CK(cudaEventSynchronize(s.ev));
  if (s.cap < bytes) {
    if (s.host) CK(cudaFreeHost(s.host));
    if (s.dev) CK(cudaFree(s.dev));
    s.host = s.dev = nullptr;
    s.cap = 0;
    const size_t cap = align_up(bytes * 2, 4096);
    CK(cudaHostAlloc(&s.host, cap, cudaHostAllocDefault));
    CK(cudaMalloc(&s.dev, cap));
    s.cap = cap;
  }
  *out = &s;
  return 0;
}
// Packs (in_len, out_len) into tightly packed offsets and uploads the four tables; fills `io` with device pointers.
static int aux_upload_io(const int32_t* in_len, const int32_t* out_len, int n, cudaStream_t st, RaggedIO* io, AuxSlot** slot_out) {
  const size_t b64 = align_up((size_t)n * 8, 256), b32 = align_up((size_t)n * 4, 256);
  AuxSlot* slot = nullptr;
  if (aux_acquire(2 * b64 + 2 * b32, &slot)) return 1;
  uint8_t* h = static_cast<uint8_t*>(slot->host);
  int64_t* h_io = reinterpret_cast<int64_t*>(h);
  int64_t* h_oo = reinterpret_cast<int64_t*>(h + b64);
  int32_t* h_il = reinterpret_cast<int32_t*>(h + 2 * b64);
  int32_t* h_ol = reinterpret_cast<int32_t*>(h + 2 * b64 + b32);
  int64_t ri = 0, ro = 0;
  for (int i = 0; i < n; ++i) {
    h_io[i] = ri; h_oo[i] = ro; h_il[i] = in_len[i]; h_ol[i] = out_len[i];
    ri += in_len[i]; ro += out_len[i];
  }
  CK(cudaMemcpyAsync(slot->dev, h, 2 * b64 + 2 * b32, cudaMemcpyHostToDevice, st));
  uint8_t* d = static_cast<uint8_t*>(slot->dev);
  io->in_off = reinterpret_cast<const int64_t*>(d);
  io->out_off = reinterpret_cast<const int64_t*>(d + b64);
  io->in_len = reinterpret_cast<const int32_t*>(d + 2 * b64);
  io->out_len = reinterpret_cast<const int32_t*>(d + 2 * b64 + b32);
  *slot_out = slot;
  return 0;
}

static int gcd_i(int a, int b) { while (b) { const int t = a % b; a = b; b = t; } return a; }

// torchaudio _get_sinc_resample_kernel (sinc_interp_hann, lowpass_filter_width 6, rolloff 0.99) evaluated in fp32 in
// the same operation order (functional.resample passes the waveform's dtype): bank [new][2 width + orig].
struct SincBank {
  float* dev = nullptr;
  int orig = 0, neu = 0, width = 0, kw = 0;
};
static std::map<std::tuple<int, int, int>, SincBank> g_sinc;  // (device, sr_in, sr_out) -> filter bank on that device
static int sinc_bank(int sr_in, int sr_out, SincBank* out) {
  const auto key = std::make_tuple(current_device_index(), sr_in, sr_out);
  auto it = g_sinc.find(key);
  if (it != g_sinc.end()) { *out = it->second; return 0; }
  const int g = gcd_i(sr_in, sr_out);
  const int orig = sr_in / g, neu = sr_out / g;
  const int lpw = 6;
  const double base_freq = (double)std::min(orig, neu) * 0.99;  // Python float arithmetic
  const int width = (int)std::ceil((double)lpw * (double)orig / base_freq);
  const int kw = 2 * width + orig;
  if ((long long)neu * kw > (1ll << 24)) return fail("resample %d -> %d: filter bank too large", sr_in, sr_out);
  std::vector<float> bank((size_t)neu * kw);
  const float bf = (float)base_freq, pi_over = (float)(M_PI / lpw / 2.0), pif = (float)M_PI, scale = (float)(base_freq / orig);
  for (int p = 0; p < neu; ++p)
    for (int k = 0; k < kw; ++k) {
      const float idx = (float)(k - width) / (float)orig;
      float t = (float)(-p) / (float)neu + idx;
      t *= bf;
      t = std::min(std::max(t, (float)-lpw), (float)lpw);
      const float c = std::cos(t * pif / (float)lpw / 2.0f);
      (void)pi_over;
      const float window = c * c;
      t *= pif;
      const float kern = (t == 0.f) ? 1.0f : std::sin(t) / t;
      bank[(size_t)p * kw + k] = kern * (window * scale);
    }
  SincBank b;
  b.orig = orig; b.neu = neu; b.width = width; b.kw = kw;
  CK(cudaMalloc(reinterpret_cast<void**>(&b.dev), bank.size() * sizeof(float)));
  CK(cudaMemcpy(b.dev, bank.data(), bank.size() * sizeof(float), cudaMemcpyHostToDevice));
  g_sinc[key] = b;
  *out = b;
  return 0;
}

int b200sep_resample_out_lens(const int32_t* lens, int32_t n_seg, int32_t sr_in, int32_t sr_out, int32_t method,
                              int32_t* out_lens) {
  if (!lens || !out_lens || n_seg <= 0) return fail("null / empty argument");
  if (sr_in <= 0 || sr_out <= 0) return fail("sample rates must be positive");
  if (method != B200SEP_RESAMPLE_LINEAR && method != B200SEP_RESAMPLE_SINC) return fail("bad resample method");
  const int g = gcd_i(sr_in, sr_out);
  for (int i = 0; i < n_seg; ++i) {
    const int n = lens[i];
    if (n < 0) return fail("negative length");
    int o = n;
    if (sr_in != sr_out && n > 1) {
      if (method == B200SEP_RESAMPLE_LINEAR) {
        // separation.py:98: int(round(len * sample_rate / sr)) (round-half-even on the correctly rounded quotient)
        const double q = (double)((long long)n * sr_out) / (double)sr_in;
        const long long t = (long long)std::nearbyint(q);
        o = (t <= 1) ? n : (int)t;  // :99-100: tgt_n <= 1 -> unchanged
      } else {
        o = (int)std::ceil((double)((long long)(sr_out / g) * n) / (double)(sr_in / g));
      }
    }
    out_lens[i] = o;
  }
  return 0;
}

int b200sep_resample_batch(const float* d_in, const int32_t* lens, int32_t n_seg, int32_t sr_in, int32_t sr_out,
                           int32_t method, float* d_out, void* stream) {
  if (!d_in || !d_out || !lens) return fail("null argument");
  std::vector<int32_t> ol(std::max(n_seg, 1));
  if (b200sep_resample_out_lens(lens, n_seg, sr_in, sr_out, method, ol.data())) return 1;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  std::lock_guard<std::mutex> lock(g_aux_mu);
  RaggedIO io;
  AuxSlot* slot = nullptr;
  if (aux_upload_io(lens, ol.data(), n_seg, st, &io, &slot)) return 1;
  int max_out = 1;
  bool any_change = false;
  for (int i = 0; i < n_seg; ++i) { max_out = std::max(max_out, ol[i]); any_change |= (ol[i] != lens[i]); }
  const dim3 grid(std::min((max_out + 255) / 256, 1024), n_seg);
  if (method == B200SEP_RESAMPLE_LINEAR || sr_in == sr_out || !any_change) {
    // (a pass-through signal has n_out == n_in: positions are the integers, the kernel copies)
    resample_linear_kernel<<<grid, 256, 0, st>>>(d_in, d_out, io);
  } else {
    SincBank b;
    if (sinc_bank(sr_in, sr_out, &b)) return 1;
    // signals of <= 1 sample pass through unchanged (overlap3_core.py:139): out_len == in_len there and the sinc
    // kernel would scale them, so they are routed through the copy path of the linear kernel first
    bool tiny = false;
    for (int i = 0; i < n_seg; ++i) tiny |= (lens[i] <= 1);
    if (tiny) return fail("sinc resample of signals with <= 1 sample is a pass-through; filter them on the host");
    SincArgs a{b.dev, b.orig, b.neu, b.width, b.kw};
    const size_t smem = ((size_t)b.neu * b.kw * 4 <= 48 * 1024) ? (size_t)b.neu * b.kw * 4 : 0;
    resample_sinc_kernel<<<grid, 256, smem, st>>>(d_in, d_out, io, a);
  }
  CK(cudaGetLastError());
  CK(cudaEventRecord(slot->ev, st));
  return 0;
}

int b200sep_frame_rms_out_lens(const int32_t* lens, int32_t n_seg, int32_t win_s, int32_t hop_s, int32_t* out_lens) {
  if (!lens || !out_lens || n_seg <= 0) return fail("null / empty argument");
  if (win_s <= 0) return fail("win too small");  // evaluate_with_sources.py:185-186
  if (hop_s <= 0) return fail("hop too small");  // :187-188
  for (int i = 0; i < n_seg; ++i) {
    if (lens[i] < 0) return fail("negative length");
    const int stop = std::max(lens[i] - win_s + 1, 1);  // range(0, max(len - win_s + 1, 1), hop_s)
    out_lens[i] = (stop + hop_s - 1) / hop_s;
  }
  return 0;
}

int b200sep_frame_rms_batch(const float* d_wav, const int32_t* lens, int32_t n_seg, int32_t win_s, int32_t hop_s,
                            float* d_rms, void* stream) {
  if (!d_wav || !d_rms || !lens) return fail("null argument");
  std::vector<int32_t> ol(std::max(n_seg, 1));
  if (b200sep_frame_rms_out_lens(lens, n_seg, win_s, hop_s, ol.data())) return 1;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  std::lock_guard<std::mutex> lock(g_aux_mu);
  RaggedIO io;
  AuxSlot* slot = nullptr;
  if (aux_upload_io(lens, ol.data(), n_seg, st, &io, &slot)) return 1;
  int max_fr = 1;
  for (int i = 0; i < n_seg; ++i) max_fr = std::max(max_fr, ol[i]);
  const dim3 grid(std::min((max_fr + 7) / 8, 2048), n_seg);
  frame_rms_kernel<<<grid, 256, 0, st>>>(d_wav, d_rms, io, RmsArgs{win_s, hop_s});
  CK(cudaGetLastError());
  CK(cudaEventRecord(slot->ev, st));
  return 0;
}

int b200sep_gt_overlap_mask_batch(const float* d_rms1, const float* d_rms2, const int32_t* n_frames, int32_t n_seg,
                                  float thr_ratio, uint8_t* d_mask, void* stream) {
  if (!d_rms1 || !d_rms2 || !n_frames || !d_mask || n_seg <= 0) return fail("null / empty argument");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  std::lock_guard<std::mutex> lock(g_aux_mu);
  RaggedIO io;
  AuxSlot* slot = nullptr;
  if (aux_upload_io(n_frames, n_frames, n_seg, st, &io, &slot)) return 1;
  gt_overlap_mask_kernel<<<n_seg, 256, 0, st>>>(d_rms1, d_rms2, io.in_off, io.in_len, thr_ratio, d_mask);
  CK(cudaGetLastError());
  CK(cudaEventRecord(slot->ev, st));
  return 0;
}

// ---------------------------------------------------------------------------------------------- debug GEMM
int b200sep_debug_gemm(const float* d_a, const float* d_b, float* d_c, int32_t M, int32_t N, int32_t K, int32_t bn,
                       int32_t use_tc, void* stream) {
  if (!d_a || !d_b || !d_c) return fail("null argument");
  if (M % 128 || !(K == 128 || K == 512) || !(bn == 128 || bn == 256) || N % bn)
    return fail("debug gemm: M%%128==0, K in {128,512}, bn in {128,256}, N%%bn==0 required");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int dev = 0;
  CK(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10) return fail("sm_100 device required");
  b200sep_model fake;
  fake.device = dev;
  fake.n_sm = prop.multiProcessorCount;
  const int nt = M / 128;
  __half *a16 = nullptr, *b16 = nullptr;
  CK(cudaMalloc(&a16, (size_t)M * K * 2));
  CK(cudaMalloc(&b16, (size_t)N * K * 2));
  to_tile_transposed_kernel<__half><<<1024, 256, 0, st>>>(d_a, a16, M, K);
  convert_kernel<__half><<<1024, 256, 0, st>>>(d_b, b16, (size_t)N * K);
  CK(cudaGetLastError());
  Tables tb{};
  tb.n_tiles = nt;
  EpiPlain e{d_c, N};
  int rc = 0;
  if (use_tc) {
    CUtensorMap ta, tbm;
    rc = make_tmap_act(&ta, a16, nt, K);
    if (!rc) rc = make_tmap_w(&tbm, b16, N, K, bn);
    uint32_t lbo = 8192, sbo = 1024;
    if (const char* s = getenv("B200SEP_DBG_A_LBO")) lbo = (uint32_t)atoi(s);
    if (const char* s = getenv("B200SEP_DBG_A_SBO")) sbo = (uint32_t)atoi(s);
    if (!rc) {
      if (bn == 256 && K == 128) rc = launch_tc<256, 128, EpiPlain>(&fake, ta, tbm, nt, N, e, tb, st, lbo, sbo);
      else if (bn == 256 && K == 512) rc = launch_tc<256, 512, EpiPlain>(&fake, ta, tbm, nt, N, e, tb, st, lbo, sbo);
      else if (bn == 128 && K == 128) rc = launch_tc<128, 128, EpiPlain>(&fake, ta, tbm, nt, N, e, tb, st, lbo, sbo);
      else rc = launch_tc<128, 512, EpiPlain>(&fake, ta, tbm, nt, N, e, tb, st, lbo, sbo);
    }
  } else {
    if (N % 128) rc = fail("simt debug gemm needs N%%128==0");
    else if (K == 128) rc = launch_simt<__half, 128, EpiPlain>(a16, b16, nt, N, e, tb, st);
    else rc = launch_simt<__half, 512, EpiPlain>(a16, b16, nt, N, e, tb, st);
  }
  cudaError_t se = cudaStreamSynchronize(st);
  cudaFree(a16);
  cudaFree(b16);
  if (rc) return rc;
  if (se != cudaSuccess) return fail("debug gemm execution failed: %s", cudaGetErrorString(se));
  return 0;
}

// Validation hook of the tensor-memory-A / CTA-pair MMA the K_A pair kernel is built on (pair_common.cuh):
// out[c][f] = sum_k W[c][k] * H[f][k], K = 64; W [256][64], H [256][64], out [256][256] fp32 device arrays (operands
// rounded to fp16).  pair = 0: one CTA, the top-left 128 x 128 block; pair = 1: a CTA pair, everything.
int b200sep_debug_ts_probe(const float* d_w, const float* d_h, float* d_out, int32_t pair, void* stream) {
  if (!d_w || !d_h || !d_out) return fail("null argument");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  __half *w16 = nullptr, *h16 = nullptr;
  CK(cudaMalloc(&w16, 256 * 64 * 2));
  CK(cudaMalloc(&h16, 256 * 64 * 2));
  convert_kernel<__half><<<64, 256, 0, st>>>(d_w, w16, (size_t)256 * 64);
  to_tile_transposed_kernel<__half><<<64, 256, 0, st>>>(d_h, h16, 256, 64);
  CK(cudaGetLastError());
  CK(cudaMemsetAsync(d_out, 0, 256 * 256 * 4, st));
  const size_t smem = 16384 + 64 + 1024;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(pair ? 2 : 1);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = pair ? 2 : 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t le;
  if (pair) le = cudaLaunchKernelEx(&cfg, pr2::ts_probe_kernel<1>, (const __half*)w16, (const __half*)h16, d_out);
  else le = cudaLaunchKernelEx(&cfg, pr2::ts_probe_kernel<0>, (const __half*)w16, (const __half*)h16, d_out);
  cudaError_t se = cudaStreamSynchronize(st);
  cudaFree(w16);
  cudaFree(h16);
  if (le != cudaSuccess) return fail("ts probe launch failed: %s", cudaGetErrorString(le));
  if (se != cudaSuccess) return fail("ts probe execution failed: %s", cudaGetErrorString(se));
  return 0;
}

}  // extern "C"

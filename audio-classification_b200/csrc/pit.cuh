// Batched PIT SI-SDR / SI-SDRi (reference: scripts/osd/overlap3_core.py:48-117 and
// scripts/osd/evaluate_with_sources.py:286-441).
//
// Pass 1 (pit_moments_kernel, HBM-bound, each sample read exactly once): per segment accumulate in
// fp64 the first and second moments of the K references, N estimates and the mixture:
//   sum r_i, sum e_j, sum m, sum r_i^2, sum e_j^2, sum m^2, sum r_i e_j, sum r_i m      (26 values)
// Pass 2 (pit_finalize_kernel): centred inner products <x,y> = sum xy - sum x sum y / n, every pair
// SI-SDR, the mixture baseline and all permutations, replaying the reference's enumeration order so
// the selected permutation is identical given the same pair matrix.
#pragma once
#include "common.cuh"

namespace b200sep {

constexpr int PIT_NACC = 26;
// accumulator layout
constexpr int PA_R = 0;    // [3] sum r_i
constexpr int PA_E = 3;    // [3] sum e_j
constexpr int PA_M = 6;    //     sum m
constexpr int PA_RR = 7;   // [3]
constexpr int PA_EE = 10;  // [3]
constexpr int PA_MM = 13;
constexpr int PA_RE = 14;  // [3][3]
constexpr int PA_RM = 23;  // [3]

constexpr int PIT_MAX_N = 8;       // estimates per segment (the reference's combinations(range(N), K) takes any N >= K)
constexpr int PIT_GROUP = 3;       // estimates per moments pass: N > 3 runs ceil(N / 3) passes over groups of <= 3
struct PitArgs {
  const float* refs; const float* ests; const float* mix;
  const int64_t* off;  // [n_seg] device
  const int32_t* len;  // [n_seg] device
  int K, N, chunk, max_chunks;
  int n_total;         // estimates per segment in the ragged layout (N = the estimates of THIS pass, first one j0)
  int j0;
};

template <int K, int N>
__global__ void __launch_bounds__(256) pit_moments_kernel(PitArgs a, double* __restrict__ partial) {
  __shared__ double s_red[8][PIT_NACC];
  const int seg = blockIdx.y, ch = blockIdx.x;
  const int T = a.len[seg];
  const int beg = ch * a.chunk;
  double* dst = partial + ((size_t)seg * a.max_chunks + ch) * PIT_NACC;
  if (beg >= T) {  // chunk beyond this segment: contributes zeros
    if (threadIdx.x < PIT_NACC) dst[threadIdx.x] = 0.0;
    return;
  }
  const int end = min(T, beg + a.chunk);
  const int64_t off = a.off[seg];
  const float* r = a.refs + (size_t)K * off;
  const float* e = a.ests + (size_t)a.n_total * off + (size_t)a.j0 * T;
  const float* m = a.mix + off;
  double acc[PIT_NACC];
#pragma unroll
  for (int i = 0; i < PIT_NACC; ++i) acc[i] = 0.0;
  // U samples per thread and trip: all (K + N + 1) * U loads are issued before the first fp64 FMA consumes one (the
  // kernel is a pure stream: memory-level parallelism, not arithmetic, sets its rate); the accumulation order of a
  // thread is unchanged (t ascending), so results are bit-identical to the one-sample-per-trip loop
  constexpr int U = 4;
  for (int t0 = beg + threadIdx.x; t0 < end; t0 += 256 * U) {
    float rl[U][3], el[U][3], ml[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int t = t0 + u * 256;
      const bool ok = t < end;
#pragma unroll
      for (int i = 0; i < K; ++i) rl[u][i] = ok ? __ldg(r + (size_t)i * T + t) : 0.f;
#pragma unroll
      for (int j = 0; j < N; ++j) el[u][j] = ok ? __ldg(e + (size_t)j * T + t) : 0.f;
      ml[u] = ok ? __ldg(m + t) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
    if (t0 + u * 256 >= end) break;
    double rv[3] = {0, 0, 0}, ev[3] = {0, 0, 0};
#pragma unroll
    for (int i = 0; i < K; ++i) rv[i] = (double)rl[u][i];
#pragma unroll
    for (int j = 0; j < N; ++j) ev[j] = (double)el[u][j];
    const double mv = (double)ml[u];
    acc[PA_M] += mv;
    acc[PA_MM] = fma(mv, mv, acc[PA_MM]);
#pragma unroll
    for (int i = 0; i < K; ++i) {
      acc[PA_R + i] += rv[i];
      acc[PA_RR + i] = fma(rv[i], rv[i], acc[PA_RR + i]);
      acc[PA_RM + i] = fma(rv[i], mv, acc[PA_RM + i]);
#pragma unroll
      for (int j = 0; j < N; ++j) acc[PA_RE + i * 3 + j] = fma(rv[i], ev[j], acc[PA_RE + i * 3 + j]);
    }
#pragma unroll
    for (int j = 0; j < N; ++j) {
      acc[PA_E + j] += ev[j];
      acc[PA_EE + j] = fma(ev[j], ev[j], acc[PA_EE + j]);
    }
    }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < PIT_NACC; ++i) {
    const double v = warp_sum(acc[i]);
    if (lane == 0) s_red[warp][i] = v;
  }
  __syncthreads();
  if (threadIdx.x < PIT_NACC) {
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) v += s_red[w][threadIdx.x];
    dst[threadIdx.x] = v;
  }
}

// SI-SDR (dB) of `est` against `ref` from centred second moments (Srr, See, Sre), following the
// formula order of overlap3_core.py:53-63 with the reference's +1e-12 guards.
__device__ __forceinline__ double si_sdr_from_moments(double Srr, double See, double Sre) {
  const double ref_energy = Srr + 1e-12;
  const double scale = Sre / ref_energy;
  const double num = scale * scale * Srr + 1e-12;
  double noise = See - 2.0 * scale * Sre + scale * scale * Srr;
  if (noise < 0.0) noise = 0.0;
  const double den = noise + 1e-12;
  return 10.0 * log10(num / den);
}

// `partial` holds one [n_seg][max_chunks][PIT_NACC] block per estimate group (group g = estimates 3g .. 3g+2); the
// reference / mixture moments are the same in every block (group 0's are used).  pair is [n_seg][3][W], W =
// max(N, 3) + 1, column W - 1 = SI-SDR against the mixture (N <= 3: the [3][4] layout).
__global__ void pit_finalize_kernel(const double* __restrict__ partial, const int32_t* __restrict__ len,
                                    int n_seg, int max_chunks, int K, int N, int variant,
                                    double* __restrict__ pair, double* __restrict__ sisdr,
                                    double* __restrict__ sisdri, int32_t* __restrict__ perm_out,
                                    double* __restrict__ gap_out) {
  const int seg = blockIdx.x * blockDim.x + threadIdx.x;
  if (seg >= n_seg) return;
  const double n = (double)len[seg];
  const double inv_n = (n > 0.0) ? 1.0 / n : 0.0;
  const int W = (N > 3 ? N : 3) + 1, MIXC = W - 1;
  const int n_groups = (N + PIT_GROUP - 1) / PIT_GROUP;
  const size_t group_stride = (size_t)n_seg * max_chunks * PIT_NACC;

  double M[3][PIT_MAX_N + 1];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j <= PIT_MAX_N; ++j) M[i][j] = nan("");
  for (int g = 0; g < n_groups; ++g) {
    double a[PIT_NACC];
    for (int i = 0; i < PIT_NACC; ++i) a[i] = 0.0;
    const double* p = partial + g * group_stride + (size_t)seg * max_chunks * PIT_NACC;
    for (int c = 0; c < max_chunks; ++c)
      for (int i = 0; i < PIT_NACC; ++i) a[i] += p[(size_t)c * PIT_NACC + i];
    const int nb = min(PIT_GROUP, N - g * PIT_GROUP);
    for (int i = 0; i < K; ++i) {
      const double Srr = a[PA_RR + i] - a[PA_R + i] * a[PA_R + i] * inv_n;
      for (int j = 0; j < nb; ++j) {
        const double See = a[PA_EE + j] - a[PA_E + j] * a[PA_E + j] * inv_n;
        const double Sre = a[PA_RE + i * 3 + j] - a[PA_R + i] * a[PA_E + j] * inv_n;
        M[i][g * PIT_GROUP + j] = si_sdr_from_moments(Srr, See, Sre);
      }
      if (g == 0) {
        const double Smm = a[PA_MM] - a[PA_M] * a[PA_M] * inv_n;
        const double Srm = a[PA_RM + i] - a[PA_R + i] * a[PA_M] * inv_n;
        M[i][MIXC] = si_sdr_from_moments(Srr, Smm, Srm);
      }
    }
  }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < W; ++j) pair[((size_t)seg * 3 + i) * W + j] = M[i][j];

  // baseline: mean_i SI-SDR(ref_i, mix)
  double base;
  if (variant == 1 && K == 2) {
    base = 0.5 * (M[0][MIXC] + M[1][MIXC]);  // evaluate_with_sources.py:389
  } else {
    double s = 0.0;  // np.mean over a K-list == pairwise sum (K <= 3: left to right) / K
    for (int i = 0; i < K; ++i) s += M[i][MIXC];
    base = s / (double)K;
  }

  double best = -1e9;
  // top-2 gap: best minus the second-best mean over the DISTINCT assignments (how far the choice is from flipping)
  double b1 = -INFINITY, b2 = -INFINITY;
  int n_cand = 0;
  auto candidate = [&](double v) {
    ++n_cand;
    if (v > b1) { b2 = b1; b1 = v; }
    else if (v > b2) b2 = v;
  };
  int bi[3] = {-1, -1, -1};
  int swapped = 0;
  bool found = false;
  const int perms3[6][3] = {{0, 1, 2}, {0, 2, 1}, {1, 0, 2}, {1, 2, 0}, {2, 0, 1}, {2, 1, 0}};
  const int perms2[2][2] = {{0, 1}, {1, 0}};
  if (variant == 1 && K == 2) {
    // evaluate_with_sources.py:352-373 -- ordered pairs j != k, unswapped preferred on >=
    for (int j = 0; j < N; ++j)
      for (int k = 0; k < N; ++k) {
        if (j == k) continue;
        const double s12 = 0.5 * (M[0][j] + M[1][k]);
        const double s21 = 0.5 * (M[0][k] + M[1][j]);
        double cur;
        int sw;
        if (!isnan(s12)) candidate(s12);  // (j, k) unswapped; its swap is pair (k, j)'s unswapped
        if (s12 >= s21) { cur = s12; sw = 0; } else { cur = s21; sw = 1; }
        if (cur > best) { best = cur; bi[0] = j; bi[1] = k; swapped = sw; found = true; }
      }
  } else {
    // cols in itertools.combinations(range(N), K), lexicographic (c0 < c1 < c2)
    for (int c0 = 0; c0 < N; ++c0)
      for (int c1 = c0 + 1; c1 < N; ++c1)
        for (int c2 = (K == 3) ? c1 + 1 : N; c2 < ((K == 3) ? N : N + 1); ++c2) {
          const int cols[3] = {c0, c1, c2};
          const int n_perm = (K == 3) ? 6 : 2;
          for (int pi = 0; pi < n_perm; ++pi) {
            int pm[3];
            for (int q = 0; q < K; ++q) pm[q] = (K == 3) ? perms3[pi][q] : perms2[pi][q];
            if (variant == 0) {
              // overlap3_core.py:78-97 -- perm in permutations(range(K)): s = sum_r M[perm[r]][cols[r]], mean = s / K,
              // NaN assignments skipped, strict >, assigned[i] = cols[perm.index(i)]
              double s = 0.0;
              bool valid = true;
              for (int q = 0; q < K; ++q) {
                const double v = M[pm[q]][cols[q]];
                if (isnan(v)) { valid = false; break; }
                s += v;
              }
              if (!valid) continue;
              const double mean_sdr = s / (double)K;
              candidate(mean_sdr);
              if (mean_sdr > best) {
                best = mean_sdr;
                found = true;
                for (int i = 0; i < K; ++i)
                  for (int q = 0; q < K; ++q)
                    if (pm[q] == i) bi[i] = cols[q];
              }
            } else {
              // evaluate_with_sources.py:418-427 (K == 3) -- perm in permutations(cols, 3), perm[i] = estimate for ref i
              const double cur = (M[0][cols[pm[0]]] + M[1][cols[pm[1]]] + M[2][cols[pm[2]]]) / 3.0;
              if (!isnan(cur)) candidate(cur);
              if (cur > best) { best = cur; found = true; for (int i = 0; i < 3; ++i) bi[i] = cols[pm[i]]; }
            }
          }
        }
  }
  if (gap_out != nullptr)
    gap_out[seg] = (found && n_cand > 1) ? b1 - b2 : ((found && n_cand == 1) ? INFINITY : nan(""));
  if (!found || isnan(base)) {
    sisdr[seg] = nan("");
    sisdri[seg] = nan("");
    for (int i = 0; i < 4; ++i) perm_out[seg * 4 + i] = -1;
    return;
  }
  sisdr[seg] = best;
  sisdri[seg] = best - base;
  for (int i = 0; i < 3; ++i) perm_out[seg * 4 + i] = (i < K) ? bi[i] : -1;
  perm_out[seg * 4 + 3] = swapped;
}

}  // namespace b200sep

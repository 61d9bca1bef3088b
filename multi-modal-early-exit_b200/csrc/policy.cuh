// Post-hoc exit policy over stored per-exit logits, on the device.
//
// Replaces the per-sample Python double loop of Policy.max_confidence_global_thresholding_policy and
// Policy.accuracy_calibration_heuristic (EE/policy.py:12-53, 55-111: for every sample, the first exit whose
// max softmax exceeds its threshold, else the last exit), the threshold sweeps that re-run that loop once per
// threshold (EE/eval.py:227-274 full_test_iteration; EE/thresh.py:106-132) and the per-exit threshold-vector
// ("mixture") sweeps of EE/thresh.py:184-215 / EE/large_scale.py:42-84 (check_2D_threshold over 1.5 M mixtures).
// The reference computes the softmax in fp64 (scipy.special.softmax on the f64 logits store, EE/utils.py:160-164),
// so this does too.
//
//   policy_crit_kernel : crit[e][s] = max softmax(logits[e][s] / T_e)  |  entropy(...), the arg-max class and a
//                        per-sample bit mask "exit e predicts the label"
//   policy_scan_kernel : thread = sample, grid.y = sweep point: exit[t][s] = first e whose criterion fires against
//                        thr[t][e]; per-(t, e) histogram and per-t correct count (shared-memory atomics).  Used when
//                        the caller wants the exit indices themselves, or for a handful of sweep points.
//   policy_hist_kernel : thread = sweep point, loop over ALL samples (criteria staged through shared memory and
//                        broadcast to the warp): histogram and correct count stay in registers, no atomics and no
//                        [n_thr, N] index matrix (240 GB for large_scale.py's 1.5 M x 40 k) is ever written.
//
// Comparison modes (PolicyCmp): the policy of EE/policy.py:33 is strict (`>`; entropy `<`, EE_modules.py:142-143) and
// the last exit fires unconditionally (:40-45).  check_2D_threshold (EE/thresh.py:184-185 = EE/large_scale.py:42-43)
// is `(CSF >= thr[:, None]).argmax(0)`: non-strict, EVERY exit's threshold is tested (the last one too) and a sample
// that fires nowhere gets exit 0 (argmax of an all-False column); its entropy CSF is the negated entropy
// (large_scale.py:15), i.e. -H >= thr  <=>  H <= -thr (the host negates the thresholds).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mmee {

enum PolicyCmp : int { POLICY_GT = 0, POLICY_LT = 1, POLICY_GE = 2, POLICY_LE = 3 };

__device__ __forceinline__ bool policy_fires(double c, double thr, int cmp) {
  return cmp == POLICY_GT ? (c > thr) : cmp == POLICY_LT ? (c < thr) : cmp == POLICY_GE ? (c >= thr) : (c <= thr);
}

// one thread per (exit e, sample s); logits [E1][N][K] fp64
__global__ void policy_crit_kernel(const double* __restrict__ logits, const double* __restrict__ temps, int E1, int64_t N,
                                   int K, int criterion, double* __restrict__ crit, int* __restrict__ argmax) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= static_cast<int64_t>(E1) * N) return;
  const int e = static_cast<int>(i / N);
  const double* x = logits + i * K;
  // the reference divides the stored logits by T_e first (EE/generic_scaling.py:60: logits / T), then softmax
  double m = temps ? x[0] / temps[e] : x[0];
  int am = 0;
  for (int k = 1; k < K; ++k) {
    const double v = temps ? x[k] / temps[e] : x[k];
    if (v > m) { m = v; am = k; }
  }
  double c;
  if (criterion == 2) {
    // CSF "margin" exactly as the reference computes it (EE/thresh.py:48-52 = EE/large_scale.py:28-32,
    // `top12_margin_np`): np.sort ascending, values[0] - values[1], i.e. the SMALLEST minus the second smallest scaled
    // logit (<= 0) — not a top-1 / top-2 margin; kept bug-compatible so stores swept with it give the reference's numbers
    double lo1 = INFINITY, lo2 = INFINITY;
    for (int k = 0; k < K; ++k) {
      const double v = temps ? x[k] / temps[e] : x[k];
      if (v < lo1) { lo2 = lo1; lo1 = v; } else if (v < lo2) { lo2 = v; }
    }
    c = lo1 - lo2;
  } else if (criterion == 0) {
    // scipy.special.softmax: exp(x - max) / sum; its maximum is exp(0) / sum
    double s = 0.0;
    for (int k = 0; k < K; ++k) s += exp((temps ? x[k] / temps[e] : x[k]) - m);
    c = 1.0 / s;
  } else {
    // EE/models/EE_modules.py:149-154: log(sum e^x) - sum(x e^x) / sum(e^x), evaluated max-shifted: with y = x - m,
    // log(sum e^x) - sum(x e^x)/sum(e^x) = log(sum e^y) - sum(y e^y)/sum(e^y) exactly (the shift cancels), and no
    // exp() overflows at small temperatures (the un-shifted form gives inf/inf = NaN once max|x|/T > 709, and a NaN
    // criterion never fires).  Same form as the engine's own exit kernel.
    double a = 0.0, b = 0.0;
    for (int k = 0; k < K; ++k) {
      const double v = (temps ? x[k] / temps[e] : x[k]) - m;
      const double ev = exp(v);
      a += ev;
      b += v * ev;
    }
    c = log(a) - b / a;
  }
  crit[i] = c;
  argmax[i] = am;
}

// cmask[s] bit e = (argmax[e][s] == labels[s]); E1 <= 64
__global__ void policy_cmask_kernel(const int* __restrict__ argmax, const int64_t* __restrict__ labels, int E1, int64_t N,
                                    unsigned long long* __restrict__ cmask) {
  const int64_t s = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (s >= N) return;
  unsigned long long m = 0ull;
  const int lab = static_cast<int>(labels[s]);
  for (int e = 0; e < E1 && e < 64; ++e)
    if (argmax[static_cast<size_t>(e) * N + s] == lab) m |= 1ull << e;
  cmask[s] = m;
}

// grid (ceil(N / blockDim), n_thr <= 65535 per launch); shared-memory histogram per block.
// n_test = number of exits whose threshold is tested (E1 - 1: the last exit fires unconditionally; E1 for
// check_2D_threshold), fallback = exit of a sample that fires nowhere (E1 - 1, or 0 for check_2D_threshold).
__global__ void policy_scan_kernel(const double* __restrict__ crit, const int* __restrict__ argmax,
                                   const double* __restrict__ thr, const int64_t* __restrict__ labels, int E1, int64_t N,
                                   int cmp, int n_test, int fallback, int32_t* __restrict__ exits,
                                   unsigned long long* __restrict__ hist, unsigned long long* __restrict__ correct) {
  extern __shared__ unsigned int s_hist[];      // [E1] + [1]
  const int t = blockIdx.y;
  for (int i = threadIdx.x; i <= E1; i += blockDim.x) s_hist[i] = 0u;
  __syncthreads();
  const int64_t s = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (s < N) {
    const double* th = thr + static_cast<size_t>(t) * E1;
    int ex = fallback;
    for (int e = 0; e < n_test; ++e) {
      if (policy_fires(crit[static_cast<size_t>(e) * N + s], th[e], cmp)) { ex = e; break; }
    }
    if (exits) exits[static_cast<size_t>(t) * N + s] = ex;
    atomicAdd(&s_hist[ex], 1u);
    if (labels && argmax[static_cast<size_t>(ex) * N + s] == static_cast<int>(labels[s])) atomicAdd(&s_hist[E1], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < E1; i += blockDim.x)
    if (s_hist[i]) atomicAdd(hist + static_cast<size_t>(t) * E1 + i, static_cast<unsigned long long>(s_hist[i]));
  if (threadIdx.x == 0 && s_hist[E1] && correct) atomicAdd(correct + t, static_cast<unsigned long long>(s_hist[E1]));
}

// thread = sweep point t (its E1 thresholds in registers), loop over all samples in chunks of POLICY_HIST_CHUNK
// staged in shared memory ([chunk][E1] criteria + the label masks): every lane reads the same address (broadcast).
// The exit is found with a branch-free reverse select chain; counts are byte-packed (one 64-bit word per 8 exits,
// flushed to 32-bit counters every 255 samples).  NE = compile-time bound on E1 (8 / 16 / 32 / 64).  The `<` / `<=`
// modes run as `>` / `>=` on negated criteria and thresholds (exact in IEEE arithmetic), so the inner loop has one
// compare per exit: STRICT selects `>` or `>=`.
constexpr int POLICY_HIST_THREADS = 128;
constexpr int POLICY_HIST_CHUNK = 255;          // samples per shared-memory stage = flush period of the byte counters

template <int NE, bool STRICT>
__global__ void __launch_bounds__(POLICY_HIST_THREADS)
policy_hist_kernel(const double* __restrict__ crit, const unsigned long long* __restrict__ cmask,
                   const double* __restrict__ thr, int E1, int64_t N, int64_t n_thr, int cmp, int n_test, int fallback,
                   long long* __restrict__ hist, long long* __restrict__ correct) {
  extern __shared__ __align__(16) unsigned char ph_smem[];
  double* s_crit = reinterpret_cast<double*>(ph_smem);                                      // [CHUNK][E1]
  unsigned long long* s_mask = reinterpret_cast<unsigned long long*>(s_crit + POLICY_HIST_CHUNK * E1);
  const int64_t t = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  const bool live = t < n_thr;
  const double sign = (cmp == POLICY_LT || cmp == POLICY_LE) ? -1.0 : 1.0;
  double th[NE];
#pragma unroll
  for (int e = 0; e < NE; ++e) th[e] = (live && e < E1) ? sign * thr[t * E1 + e] : 0.0;
  unsigned int cnt[NE];
#pragma unroll
  for (int e = 0; e < NE; ++e) cnt[e] = 0u;
  unsigned int n_correct = 0u;
  for (int64_t s0 = 0; s0 < N; s0 += POLICY_HIST_CHUNK) {
    const int n = static_cast<int>(min(static_cast<int64_t>(POLICY_HIST_CHUNK), N - s0));
    __syncthreads();
    for (int i = threadIdx.x; i < n * E1; i += blockDim.x) {
      const int e = i / n, s = i - e * n;                        // coalesced over s for each exit row
      s_crit[s * E1 + e] = sign * crit[static_cast<size_t>(e) * N + s0 + s];
    }
    for (int i = threadIdx.x; i < n; i += blockDim.x) s_mask[i] = cmask ? cmask[s0 + i] : 0ull;
    __syncthreads();
    unsigned long long packed[NE / 8];
#pragma unroll
    for (int w = 0; w < NE / 8; ++w) packed[w] = 0ull;
    for (int s = 0; s < n; ++s) {
      const double* c = s_crit + s * E1;
      int ex = fallback;
#pragma unroll
      for (int e = NE - 1; e >= 0; --e)
        if (e < n_test && (STRICT ? (c[e] > th[e]) : (c[e] >= th[e]))) ex = e;
      const unsigned long long one = 1ull << ((ex & 7) * 8);
#pragma unroll
      for (int w = 0; w < NE / 8; ++w) packed[w] += ((ex >> 3) == w) ? one : 0ull;
      n_correct += static_cast<unsigned int>((s_mask[s] >> ex) & 1ull);
    }
#pragma unroll
    for (int e = 0; e < NE; ++e) cnt[e] += static_cast<unsigned int>((packed[e >> 3] >> ((e & 7) * 8)) & 0xFFull);
  }
  if (!live) return;
#pragma unroll
  for (int e = 0; e < NE; ++e)
    if (e < E1) hist[t * E1 + e] = static_cast<long long>(cnt[e]);
  if (correct) correct[t] = static_cast<long long>(n_correct);
}

}  // namespace mmee

// Post-hoc exit policy over stored per-exit logits, on the device.
//
// Replaces the per-sample Python double loop of Policy.max_confidence_global_thresholding_policy and
// Policy.accuracy_calibration_heuristic (EE/policy.py:12-53, 55-111: for every sample, the first exit whose
// max softmax exceeds its threshold, else the last exit) and the threshold sweeps that re-run that loop once per
// threshold (EE/eval.py:227-274 full_test_iteration; EE/thresh.py:106-132, 184-215).  The reference computes the
// softmax in fp64 (scipy.special.softmax on the f64 logits store, EE/utils.py:160-164), so this does too.
//
//   policy_crit_kernel : crit[e][s] = max softmax(logits[e][s] / T_e)  |  entropy(...)   and the arg-max class
//   policy_scan_kernel : for every sweep point t and sample s: exit[t][s] = first e with crit "fires" against
//                        thr[t][e] (strict > for max-confidence, strict < for entropy, EE/models/EE_modules.py:139-143);
//                        the last exit always fires; per-(t, e) histogram and per-t correct-prediction count
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mmee {

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
  if (criterion == 0) {
    // scipy.special.softmax: exp(x - max) / sum; its maximum is exp(0) / sum
    double s = 0.0;
    for (int k = 0; k < K; ++k) s += exp((temps ? x[k] / temps[e] : x[k]) - m);
    c = 1.0 / s;
  } else {
    // EE/models/EE_modules.py:149-154: log(sum e^x) - sum(x e^x) / sum(e^x)   (un-shifted, as the reference)
    double a = 0.0, b = 0.0;
    for (int k = 0; k < K; ++k) {
      const double v = temps ? x[k] / temps[e] : x[k];
      const double ev = exp(v);
      a += ev;
      b += v * ev;
    }
    c = log(a) - b / a;
  }
  crit[i] = c;
  argmax[i] = am;
}

// grid (ceil(N / blockDim), n_thr); shared-memory histogram per block
__global__ void policy_scan_kernel(const double* __restrict__ crit, const int* __restrict__ argmax,
                                   const double* __restrict__ thr, const int64_t* __restrict__ labels, int E1, int64_t N,
                                   int criterion, int32_t* __restrict__ exits, unsigned long long* __restrict__ hist,
                                   unsigned long long* __restrict__ correct) {
  extern __shared__ unsigned int s_hist[];      // [E1] + [1]
  const int t = blockIdx.y;
  for (int i = threadIdx.x; i <= E1; i += blockDim.x) s_hist[i] = 0u;
  __syncthreads();
  const int64_t s = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (s < N) {
    const double* th = thr + static_cast<size_t>(t) * E1;
    int ex = E1 - 1;
    for (int e = 0; e < E1 - 1; ++e) {
      const double c = crit[static_cast<size_t>(e) * N + s];
      const bool fire = criterion == 0 ? (c > th[e]) : (c < th[e]);
      if (fire) { ex = e; break; }
    }
    exits[static_cast<size_t>(t) * N + s] = ex;
    atomicAdd(&s_hist[ex], 1u);
    if (labels && argmax[static_cast<size_t>(ex) * N + s] == static_cast<int>(labels[s])) atomicAdd(&s_hist[E1], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < E1; i += blockDim.x)
    if (s_hist[i]) atomicAdd(hist + static_cast<size_t>(t) * E1 + i, static_cast<unsigned long long>(s_hist[i]));
  if (threadIdx.x == 0 && s_hist[E1]) atomicAdd(correct + t, static_cast<unsigned long long>(s_hist[E1]));
}

}  // namespace mmee

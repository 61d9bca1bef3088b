// Post-hoc temperature calibration of the per-exit logits, on the device.
//
// Replaces TemperatureScaler.set_temperature (EE/generic_scaling.py:64-111: scipy L-BFGS-B with finite-difference
// gradients on log_loss(labels, softmax(logits / T))) as EE/eval.py:311-335 runs it once per exit, and the
// per-exit statistics that loop collects (accuracy, average max-softmax confidence after scaling).
//
// The objective in beta = 1 / T is   nll(beta) = mean_i [ logsumexp(beta z_i) - beta z_i[y_i] ]
// with   d/dbeta  = mean_i [ E_p z - z[y] ]   and   d2/dbeta2 = mean_i Var_p z >= 0   (p = softmax(beta z)),
// i.e. a one-dimensional convex problem: safeguarded Newton steps reach the minimiser to machine precision in a
// handful of iterations, every exit in parallel, fp64 like the reference.  The reference's own answer sits within
// ~1e-4 (relative) of that minimiser (L-BFGS-B's default tolerances; measured in tests/golden/make_calibration_golden.py).
//
//   calib_stats_kernel  : per (exit, block) partial sums over samples of {nll, g, h, max-softmax, arg-max == label}
//   calib_update_kernel : one thread per exit: fold the partials in a fixed order (deterministic), accept / backtrack,
//                         next beta
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mmee {

constexpr int CALIB_THREADS = 256;
constexpr int CALIB_NSTAT = 5;

// grid (blocks, E1); logits [E1][N][K] fp64; beta [E1]; partial [E1][blocks][CALIB_NSTAT]
__global__ void __launch_bounds__(CALIB_THREADS) calib_stats_kernel(const double* __restrict__ logits,
                                                                    const int64_t* __restrict__ labels,
                                                                    const double* __restrict__ beta, int64_t N, int K,
                                                                    double* __restrict__ partial) {
  const int e = blockIdx.y;
  const double b = beta[e];
  double acc[CALIB_NSTAT] = {0.0, 0.0, 0.0, 0.0, 0.0};
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < N;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const double* z = logits + (static_cast<int64_t>(e) * N + i) * K;
    const int y = static_cast<int>(labels[i]);
    double m = z[0];
    int am = 0;
    for (int k = 1; k < K; ++k)
      if (z[k] > m) { m = z[k]; am = k; }               // first maximum, like numpy argmax (beta > 0 keeps the order)
    double s = 0.0, sz = 0.0, sz2 = 0.0;
    for (int k = 0; k < K; ++k) {
      const double w = exp(b * (z[k] - m));
      s += w;
      sz += w * z[k];
      sz2 += w * z[k] * z[k];
    }
    const double ez = sz / s;
    acc[0] += b * m + log(s) - b * z[y];                // -log softmax(beta z)[y]
    acc[1] += ez - z[y];
    acc[2] += sz2 / s - ez * ez;
    acc[3] += 1.0 / s;                                  // max softmax(beta z)
    acc[4] += (am == y) ? 1.0 : 0.0;
  }
  __shared__ double red[CALIB_NSTAT][CALIB_THREADS / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < CALIB_NSTAT; ++q) {
    double v = acc[q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) red[q][warp] = v;
  }
  __syncthreads();
  if (threadIdx.x < CALIB_NSTAT) {
    double v = 0.0;
    for (int w = 0; w < CALIB_THREADS / 32; ++w) v += red[threadIdx.x][w];
    partial[(static_cast<size_t>(e) * gridDim.x + blockIdx.x) * CALIB_NSTAT + threadIdx.x] = v;
  }
}

struct CalibState {
  double beta_cur, beta_ok, nll_ok, step, nll_first;
  int have_ok;
};

// one thread per exit.  stats_out [E1][CALIB_NSTAT] = the folded means of this evaluation (used by the caller for the
// final statistics pass).
__global__ void calib_update_kernel(const double* __restrict__ partial, int blocks, int64_t N, int E1,
                                    CalibState* __restrict__ st, double* __restrict__ beta,
                                    double* __restrict__ stats_out, int advance) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E1) return;
  double v[CALIB_NSTAT];
  for (int q = 0; q < CALIB_NSTAT; ++q) {
    double s = 0.0;
    for (int bk = 0; bk < blocks; ++bk) s += partial[(static_cast<size_t>(e) * blocks + bk) * CALIB_NSTAT + q];
    v[q] = s / static_cast<double>(N);
    stats_out[e * CALIB_NSTAT + q] = v[q];
  }
  if (!advance) return;
  CalibState s = st[e];
  const double nll = v[0], g = v[1], h = v[2];
  if (!s.have_ok || nll <= s.nll_ok) {
    // accept this point and take a Newton step from it (bounded to a factor 4 per step; beta stays positive)
    if (!s.have_ok) s.nll_first = nll;
    s.have_ok = 1;
    s.beta_ok = s.beta_cur;
    s.nll_ok = nll;
    double nb = (h > 1e-300) ? s.beta_cur - g / h : (g < 0.0 ? 4.0 * s.beta_cur : 0.25 * s.beta_cur);
    nb = fmin(fmax(nb, 0.25 * s.beta_cur), 4.0 * s.beta_cur);
    nb = fmin(fmax(nb, 1e-12), 1e12);
    s.step = nb - s.beta_cur;
  } else {
    s.step *= 0.5;                                      // the step overshot: backtrack towards the accepted point
  }
  s.beta_cur = s.beta_ok + s.step;
  st[e] = s;
  beta[e] = s.beta_cur;
}

}  // namespace mmee

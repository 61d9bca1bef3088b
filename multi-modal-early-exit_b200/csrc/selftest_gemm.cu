// Developer self-test for the tcgen05 GEMM (no torch): compares against a naive device GEMM and times it.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo selftest_gemm.cu -o selftest_gemm
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "gemm.cuh"
#include "tmap.h"

using namespace mmee;

#define CK(x)                                                                                  \
  do {                                                                                         \
    cudaError_t e_ = (x);                                                                      \
    if (e_ != cudaSuccess) {                                                                   \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);          \
      exit(2);                                                                                 \
    }                                                                                          \
  } while (0)

__global__ void fill_bf16(__nv_bfloat16* p, size_t n, uint32_t seed, float scale) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t h = (uint32_t)i * 2654435761u + seed;
  h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
  p[i] = __float2bfloat16_rn(((h & 0xFFFF) / 65536.0f - 0.5f) * 2.f * scale);
}
__global__ void fill_f32(float* p, size_t n, uint32_t seed, float scale) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t h = (uint32_t)i * 2246822519u + seed;
  h ^= h >> 15; h *= 0x85ebca6bu; h ^= h >> 13;
  p[i] = ((h & 0xFFFF) / 65536.0f - 0.5f) * 2.f * scale;
}
// naive reference: one thread per output
__global__ void ref_gemm(const __nv_bfloat16* A, const __nv_bfloat16* W, const float* bias, float* C, int M, int N,
                         int K) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  int m = blockIdx.y;
  if (n >= N || m >= M) return;
  float acc = 0.f;
  for (int k = 0; k < K; ++k) acc += __bfloat162float(A[(size_t)m * K + k]) * __bfloat162float(W[(size_t)n * K + k]);
  C[(size_t)m * N + n] = acc + bias[n];
}

static double gelu_ref(double x) { return 0.5 * x * (1.0 + erf(x / sqrt(2.0))); }

static bool g_pair = false;      // use the CTA-pair kernel (BN must be 256; tb then has 128-row boxes)
template <int BN, int EPI, bool SPLIT = false>
static void launch(const CUtensorMap& ta, const CUtensorMap& tb, const GemmArgs& a, int sms, cudaStream_t st = 0,
                   const CUtensorMap* ta_lo = nullptr, const CUtensorMap* tb_lo = nullptr) {
  if (!ta_lo) ta_lo = &ta;
  if (!tb_lo) tb_lo = &tb;
  if (g_pair && BN == 256) {
    auto kp = gemm_tc_pair_kernel<EPI, SPLIT>;
    using PS = GemmPairSmemT<gemm_pair_epi_warps<EPI>()>;
    CK(cudaFuncSetAttribute(kp, cudaFuncAttributeMaxDynamicSharedMemorySize, PS::DYN_BYTES));
    kp<<<sms & ~1, PS::THREADS, PS::DYN_BYTES, st>>>(ta, tb, *ta_lo, *tb_lo, a);
    return;
  }
  auto kern = gemm_tc_kernel<BN, EPI, SPLIT>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmSmem<BN>::DYN_BYTES));
  kern<<<sms, GEMM_THREADS, GemmSmem<BN>::DYN_BYTES, st>>>(ta, tb, *ta_lo, *tb_lo, a);
}

// ---- fp32 engine mode (SPLIT): fp32 operands as split-bf16 pairs, fp64 reference on the fp32 values
__global__ void split_f32(const float* src, __nv_bfloat16* hi, __nv_bfloat16* lo, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const __nv_bfloat16 h = __float2bfloat16_rn(src[i]);
  hi[i] = h;
  lo[i] = __float2bfloat16_rn(src[i] - __bfloat162float(h));
}
__global__ void ref_gemm_f64(const float* A, const float* W, const float* bias, double* C, int M, int N, int K) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  int m = blockIdx.y;
  if (n >= N || m >= M) return;
  double acc = 0.0;
  for (int k = 0; k < K; ++k) acc += (double)A[(size_t)m * K + k] * (double)W[(size_t)n * K + k];
  C[(size_t)m * N + n] = acc + bias[n];
}

template <int BN, int EPI>
static int run_split_case(const char* name, int M, int N, int K, int sms, bool timing) {
  int Mmax = ((M + 255) / 256) * 256 + 128;
  float *A32, *W32, *bias, *O32;
  __nv_bfloat16 *A, *Al, *W, *Wl, *R, *Rl, *O16, *O16l;
  double* Cref;
  CK(cudaMalloc(&A32, (size_t)Mmax * K * 4)); CK(cudaMalloc(&W32, (size_t)N * K * 4));
  CK(cudaMalloc(&A, (size_t)Mmax * K * 2)); CK(cudaMalloc(&Al, (size_t)Mmax * K * 2));
  CK(cudaMalloc(&W, (size_t)N * K * 2)); CK(cudaMalloc(&Wl, (size_t)N * K * 2));
  CK(cudaMalloc(&R, (size_t)Mmax * N * 2)); CK(cudaMalloc(&Rl, (size_t)Mmax * N * 2));
  CK(cudaMalloc(&O16, (size_t)Mmax * N * 2)); CK(cudaMalloc(&O16l, (size_t)Mmax * N * 2));
  CK(cudaMalloc(&O32, (size_t)Mmax * N * 4));
  CK(cudaMalloc(&Cref, (size_t)M * N * 8));
  CK(cudaMalloc(&bias, N * 4));
  int* m_dev;
  CK(cudaMalloc(&m_dev, 4));
  CK(cudaMemcpy(m_dev, &M, 4, cudaMemcpyHostToDevice));
  fill_f32<<<((size_t)Mmax * K + 255) / 256, 256>>>(A32, (size_t)Mmax * K, 11, 1.0f);
  fill_f32<<<((size_t)N * K + 255) / 256, 256>>>(W32, (size_t)N * K, 12, 0.05f);
  fill_f32<<<(N + 255) / 256, 256>>>(bias, N, 4, 0.5f);
  split_f32<<<((size_t)Mmax * K + 255) / 256, 256>>>(A32, A, Al, (size_t)Mmax * K);
  split_f32<<<((size_t)N * K + 255) / 256, 256>>>(W32, W, Wl, (size_t)N * K);
  CK(cudaMemset(R, 0, (size_t)Mmax * N * 2)); CK(cudaMemset(Rl, 0, (size_t)Mmax * N * 2));
  CK(cudaMemset(O16, 0xFF, (size_t)Mmax * N * 2)); CK(cudaMemset(O16l, 0xFF, (size_t)Mmax * N * 2));
  CK(cudaMemset(O32, 0xFF, (size_t)Mmax * N * 4));
  const int wb = (g_pair && BN == 256) ? 128 : BN;
  CUtensorMap ta = make_tmap_2d_sw128(A, Mmax, K, K, 128), tal = make_tmap_2d_sw128(Al, Mmax, K, K, 128);
  CUtensorMap tb = make_tmap_2d_sw128(W, N, K, K, wb), tbl = make_tmap_2d_sw128(Wl, N, K, K, wb);
  GemmArgs a{};
  a.m_dev = m_dev; a.m_static = M; a.N = N; a.K = K; a.bias = bias; a.ld_out = N; a.resid = R; a.resid_lo = Rl;
  a.out = (EPI == EPI_RESID_F32) ? (void*)O32 : (void*)O16;
  a.out_lo = O16l;
  launch<BN, EPI, true>(ta, tb, a, sms, 0, &tal, &tbl);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  ref_gemm_f64<<<dim3((N + 127) / 128, M), 128>>>(A32, W32, bias, Cref, M, N, K);
  CK(cudaDeviceSynchronize());
  std::vector<double> ref((size_t)M * N);
  CK(cudaMemcpy(ref.data(), Cref, ref.size() * 8, cudaMemcpyDeviceToHost));
  std::vector<float> o32;
  std::vector<__nv_bfloat16> oh, ol;
  if (EPI == EPI_RESID_F32) {
    o32.resize((size_t)M * N);
    CK(cudaMemcpy(o32.data(), O32, o32.size() * 4, cudaMemcpyDeviceToHost));
  } else {
    oh.resize((size_t)M * N); ol.resize((size_t)M * N);
    CK(cudaMemcpy(oh.data(), O16, oh.size() * 2, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(ol.data(), O16l, ol.size() * 2, cudaMemcpyDeviceToHost));
  }
  int bad = 0;
  double maxerr = 0, scale = 0;
  for (size_t i = 0; i < ref.size(); ++i) scale = fmax(scale, fabs(ref[i]));
  for (size_t i = 0; i < ref.size(); ++i) {
    double want = ref[i];
    if (EPI == EPI_GELU_BF16) want = gelu_ref(want);
    const double got = (EPI == EPI_RESID_F32) ? o32[i] : (double)__bfloat162float(oh[i]) + (double)__bfloat162float(ol[i]);
    const double err = fabs(got - want);
    // products carry ~2^-17 relative error each (dropped lo x lo, operand residuals); outputs as hi + lo another 2^-17
    const double tol = 3e-5 * scale;
    if (!(err <= tol)) { if (bad < 5) printf("  mismatch @%zu: got %.9g want %.9g\n", i, got, want); ++bad; }
    if (err > maxerr) maxerr = err;
  }
  printf("%-28s M=%d N=%d K=%d  maxerr=%.4g (|C| up to %.3g)  %s\n", name, M, N, K, maxerr, scale, bad ? "FAIL" : "ok");
  if (timing) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) launch<BN, EPI, true>(ta, tb, a, sms, 0, &tal, &tbl);
    cudaEventRecord(e0);
    const int it = 10;
    for (int i = 0; i < it; ++i) launch<BN, EPI, true>(ta, tb, a, sms, 0, &tal, &tbl);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= it;
    printf("    time %.3f ms  %.1f TFLOP/s of fp32-equivalent work (%.1f TFLOP/s of bf16 MMAs)\n", ms,
           2.0 * M * N * K / ms / 1e9, 6.0 * M * N * K / ms / 1e9);
  }
  cudaFree(A32); cudaFree(W32); cudaFree(A); cudaFree(Al); cudaFree(W); cudaFree(Wl); cudaFree(R); cudaFree(Rl);
  cudaFree(O16); cudaFree(O16l); cudaFree(O32); cudaFree(Cref); cudaFree(bias); cudaFree(m_dev);
  return bad;
}

template <int BN, int EPI>
static int run_case(const char* name, int M, int N, int K, int sms, bool timing) {
  int Mmax = ((M + 255) / 256) * 256 + 128;
  __nv_bfloat16 *A, *W, *R, *O16;
  float *bias, *Cref, *O32;
  CK(cudaMalloc(&A, (size_t)Mmax * K * 2));
  CK(cudaMalloc(&W, (size_t)N * K * 2));
  CK(cudaMalloc(&R, (size_t)Mmax * N * 2));
  CK(cudaMalloc(&O16, (size_t)Mmax * N * 2));
  CK(cudaMalloc(&O32, (size_t)Mmax * N * 4));
  CK(cudaMalloc(&Cref, (size_t)M * N * 4));
  CK(cudaMalloc(&bias, N * 4));
  int* m_dev;
  CK(cudaMalloc(&m_dev, 4));
  CK(cudaMemcpy(m_dev, &M, 4, cudaMemcpyHostToDevice));
  fill_bf16<<<((size_t)Mmax * K + 255) / 256, 256>>>(A, (size_t)Mmax * K, 1, 1.0f);
  fill_bf16<<<((size_t)N * K + 255) / 256, 256>>>(W, (size_t)N * K, 2, 0.05f);
  fill_bf16<<<((size_t)Mmax * N + 255) / 256, 256>>>(R, (size_t)Mmax * N, 3, 1.0f);
  fill_f32<<<(N + 255) / 256, 256>>>(bias, N, 4, 0.5f);
  CK(cudaMemset(O16, 0xFF, (size_t)Mmax * N * 2));
  CK(cudaMemset(O32, 0xFF, (size_t)Mmax * N * 4));
  CUtensorMap ta = make_tmap_2d_sw128(A, Mmax, K, K, 128);
  CUtensorMap tb = make_tmap_2d_sw128(W, N, K, K, (g_pair && BN == 256) ? 128 : BN);
  GemmArgs a{};
  a.m_dev = m_dev; a.m_static = M; a.N = N; a.K = K; a.bias = bias; a.ld_out = N; a.resid = R;
  a.out = (EPI == EPI_RESID_F32) ? (void*)O32 : (void*)O16;
  launch<BN, EPI>(ta, tb, a, sms);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  int bad = 0;
  double maxerr = 0;
  if (!timing || M <= 4096) {
    ref_gemm<<<dim3((N + 127) / 128, M), 128>>>(A, W, bias, Cref, M, N, K);
    CK(cudaDeviceSynchronize());
    std::vector<float> ref((size_t)M * N), o32;
    std::vector<__nv_bfloat16> o16, r16;
    CK(cudaMemcpy(ref.data(), Cref, ref.size() * 4, cudaMemcpyDeviceToHost));
    if (EPI == EPI_RESID_F32) {
      o32.resize((size_t)M * N); r16.resize((size_t)M * N);
      CK(cudaMemcpy(o32.data(), O32, o32.size() * 4, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(r16.data(), R, r16.size() * 2, cudaMemcpyDeviceToHost));
    } else {
      o16.resize((size_t)M * N);
      CK(cudaMemcpy(o16.data(), O16, o16.size() * 2, cudaMemcpyDeviceToHost));
    }
    for (size_t i = 0; i < ref.size(); ++i) {
      double want = ref[i], got, tol;
      if (EPI == EPI_RESID_F32) { want += __bfloat162float(r16[i]); got = o32[i]; tol = 2e-3 + 1e-4 * fabs(want); }
      else { if (EPI == EPI_GELU_BF16) want = gelu_ref(want); got = __bfloat162float(o16[i]); tol = 1e-2 + 8e-3 * fabs(want); }
      double err = fabs(got - want);
      if (!(err <= tol)) { if (bad < 5) printf("  mismatch @%zu (m=%zu n=%zu): got %f want %f\n", i, i / N, i % N, got, want); ++bad; }
      if (err > maxerr) maxerr = err;
    }
    // rows >= M must be untouched (0xFF pattern)
    std::vector<uint16_t> tail(N);
    if (EPI != EPI_RESID_F32) {
      CK(cudaMemcpy(tail.data(), O16 + (size_t)M * N, N * 2, cudaMemcpyDeviceToHost));
      for (int i = 0; i < N; ++i) if (tail[i] != 0xFFFF) { ++bad; if (bad < 8) printf("  row M written!\n"); break; }
    }
  }
  printf("%-28s M=%d N=%d K=%d  maxerr=%.4g  %s\n", name, M, N, K, maxerr, bad ? "FAIL" : "ok");
  if (timing) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) launch<BN, EPI>(ta, tb, a, sms);
    cudaEventRecord(e0);
    const int it = 20;
    for (int i = 0; i < it; ++i) launch<BN, EPI>(ta, tb, a, sms);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= it;
    printf("    time %.3f ms  %.1f TFLOP/s\n", ms, 2.0 * M * N * K / ms / 1e9);
  }
  cudaFree(A); cudaFree(W); cudaFree(R); cudaFree(O16); cudaFree(O32); cudaFree(Cref); cudaFree(bias); cudaFree(m_dev);
  return bad;
}

int main(int argc, char** argv) {
  int dev = 0; cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, dev));
  int sms = prop.multiProcessorCount;
  printf("device %s, %d SMs\n", prop.name, sms);
  int bad = 0;
  bad += run_case<128, EPI_BIAS_BF16>("bias bn128 small", 200, 128, 128, sms, false);
  bad += run_case<128, EPI_BIAS_BF16>("bias bn128", 2127, 384, 128, sms, false);
  bad += run_case<256, EPI_BIAS_BF16>("bias bn256", 2127, 2304, 768, sms, false);
  bad += run_case<256, EPI_GELU_BF16>("gelu bn256", 2127, 3072, 768, sms, false);
  bad += run_case<256, EPI_RESID_F32>("resid bn256", 2127, 768, 3072, sms, false);
  bad += run_case<128, EPI_RESID_F32>("resid bn128", 1418, 128, 256, sms, false);
  g_pair = true;
  printf("-- CTA-pair (cta_group::2) kernel\n");
  bad += run_case<256, EPI_BIAS_BF16>("pair bias", 2127, 2304, 768, sms, false);
  bad += run_case<256, EPI_GELU_BF16>("pair gelu", 2127, 3072, 768, sms, false);
  bad += run_case<256, EPI_RESID_F32>("pair resid", 2127, 768, 3072, sms, false);
  bad += run_case<256, EPI_RESID_F32>("pair resid small", 100, 768, 768, sms, false);
  printf("-- fp32 engine mode (split-bf16 operands, three k segments)\n");
  g_pair = false;
  bad += run_split_case<128, EPI_RESID_F32>("split resid bn128", 1418, 128, 256, sms, false);
  bad += run_split_case<128, EPI_BIAS_BF16>("split bias bn128", 300, 384, 128, sms, false);
  g_pair = true;
  bad += run_split_case<256, EPI_RESID_F32>("split pair resid", 2127, 768, 3072, sms, false);
  bad += run_split_case<256, EPI_BIAS_BF16>("split pair bias", 2127, 2304, 768, sms, false);
  bad += run_split_case<256, EPI_GELU_BF16>("split pair gelu", 2127, 3072, 768, sms, false);
  if (argc > 1) bad += run_split_case<256, EPI_RESID_F32>("T split mlp-down", 64 * 709, 768, 3072, sms, true);
  for (int pass = 0; pass < 2 && argc > 1; ++pass) {
    g_pair = pass == 1;
    printf(g_pair ? "-- timing, CTA-pair kernel\n" : "-- timing, single-CTA kernel\n");
    int M = 256 * 709;
    bad += run_case<256, EPI_BIAS_BF16>("T qkv", M, 2304, 768, sms, true);
    bad += run_case<256, EPI_GELU_BF16>("T mlp-up gelu", M, 3072, 768, sms, true);
    bad += run_case<256, EPI_RESID_F32>("T mlp-down resid", M, 768, 3072, sms, true);
    bad += run_case<256, EPI_RESID_F32>("T out-proj resid", M, 768, 768, sms, true);
  }
  printf(bad ? "SELFTEST FAILED (%d)\n" : "SELFTEST PASSED\n", bad);
  return bad ? 1 : 0;
}

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

static bool g_pair = false;      // use the CTA-pair kernel (BN must be 256; tb then has 128-row boxes)
template <int BN, int EPI>
static void launch(const CUtensorMap& ta, const CUtensorMap& tb, const GemmArgs& a, int sms, cudaStream_t st = 0) {
  if (g_pair && BN == 256) {
    auto kp = gemm_tc_pair_kernel<EPI>;
    using PS = GemmPairSmemT<gemm_pair_epi_warps<EPI>()>;
    CK(cudaFuncSetAttribute(kp, cudaFuncAttributeMaxDynamicSharedMemorySize, PS::DYN_BYTES));
    kp<<<sms & ~1, PS::THREADS, PS::DYN_BYTES, st>>>(ta, tb, a);
    return;
  }
  auto kern = gemm_tc_kernel<BN, EPI>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmSmem<BN>::DYN_BYTES));
  kern<<<sms, GEMM_THREADS, GemmSmem<BN>::DYN_BYTES, st>>>(ta, tb, a);
}

static double gelu_ref(double x) { return 0.5 * x * (1.0 + erf(x / sqrt(2.0))); }

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

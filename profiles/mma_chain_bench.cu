// Developer micro-benchmark (not part of the product): what does a SHORT chain of small tcgen05.mma instructions cost
// on sm_100a?  The attention kernel issues 12 MMAs per 128 x 64 key tile (N = 64 / 16 / 80, K = 16); its clock64
// trace shows ~700 cycles from the issue of a 4..8 instruction batch to its completion.  This program times, for one
// CTA (optionally two co-resident CTAs), batches of n MMAs in several shapes and dependency patterns:
//   t_issue = cycles until the issuing thread is past the batch + tcgen05.commit
//   t_done  = cycles until the committed mbarrier flips
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I multi-modal-early-exit_b200/csrc \
//              profiles/mma_chain_bench.cu -o build/mma_chain_bench
// Run  :  build/mma_chain_bench            (prints one table; operands are zeros: timing only)
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "ptx.cuh"

using namespace mmee;

#define CK(x)                                                                                     \
  do {                                                                                            \
    cudaError_t e_ = (x);                                                                         \
    if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } \
  } while (0)

enum Pattern : int {
  DEP = 0,      // all n MMAs accumulate into the same D
  ALT2 = 1,     // two accumulators, alternating
  ALT4 = 2,     // four accumulators, round robin
  COLS = 3,     // MMA k writes its own N-column block of one D (independent, like the block-diagonal bias product)
  TILE_S = 4,   // the attention kernel's S batch: 4 dependent TS N = 64, then 4 SS N = 16 on column blocks of the same D
};

struct Case {
  int ts;        // 1: A from TMEM (TS form), 0: A from shared memory (SS form)
  int N;
  int n;         // instructions in the batch
  int pattern;
  int commits;   // tcgen05.commit instructions after the batch (1 or 2, on different barriers)
};

constexpr int SMEM_A = 16384;            // [128 x 64] bf16 SW128
constexpr int SMEM_B = 32768;            // [256 x 64] bf16 SW128
constexpr int SMEM_TOTAL = SMEM_A + SMEM_B + 64;
constexpr int REPS = 12;

__global__ void __launch_bounds__(128, 2) bench_kernel(const Case* cases, int n_cases, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sb = smem_u32(smem);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SMEM_A + SMEM_B);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 4);
  for (int i = threadIdx.x; i < (SMEM_A + SMEM_B) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    mbar_init(bars, 1);
    mbar_init(bars + 1, 1);
    fence_mbar_init();
  }
  fence_proxy_async_smem();
  if (threadIdx.x < 32) tmem_alloc<256>(slot);             // 256 columns so that two CTAs fit one SM
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  // zero the TMEM columns used as A / D (timing only, but keep NaNs out)
  {
    uint32_t z[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) z[i] = 0;
    const uint32_t lane_addr = static_cast<uint32_t>((threadIdx.x >> 5) * 32) << 16;
    for (int cb = 0; cb < 256; cb += 32) tmem_st32(tmem + lane_addr + cb, z);
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 0) {
    const uint64_t da = umma_desc_sw128_kmajor(sb);
    const uint64_t db = umma_desc_sw128_kmajor(sb + SMEM_A);
    uint32_t phase0 = 0, phase1 = 0;
    const uint32_t b0 = smem_u32(bars), b1 = smem_u32(bars + 1);
    for (int ci = 0; ci < n_cases; ++ci) {
      Case c = cases[ci];
      long long best_issue = 1ll << 60, best_done = 1ll << 60;
      for (int rep = 0; rep < REPS; ++rep) {
        const long long t0 = clock64();
        // TMEM here is 256 columns: fold the layout of issue_batch into that window
        {
          const uint32_t idesc = umma_idesc_bf16(128, c.N);
          const uint32_t tmem_a = tmem + 224;
          for (int i = 0; i < c.n; ++i) {
            uint32_t d = tmem;
            if (c.pattern == ALT2) d += (i & 1) * c.N;
            else if (c.pattern == ALT4) d += (i & 3) * c.N;
            else if (c.pattern == COLS) d += (i * c.N) % 192;
            const int k = i & 3;
            if (c.pattern == TILE_S) {
              if (i < 4) umma_bf16_ts(tmem, tmem_a + k * 8, db + 2 * k, umma_idesc_bf16(128, 64), 1u);
              else umma_bf16_ss(tmem + 16 * k, da + 2 * k, db + 2 * k, umma_idesc_bf16(128, 16), 1u);
              continue;
            }
            if (c.ts) umma_bf16_ts(d, tmem_a + k * 8, db + 2 * k, idesc, 1u);
            else umma_bf16_ss(d, da + 2 * k, db + 2 * k, idesc, 1u);
          }
        }
        umma_commit(b0);
        if (c.commits > 1) umma_commit(b1);
        const long long t1 = clock64();
        mbar_wait(b0, phase0); phase0 ^= 1;
        if (c.commits > 1) { mbar_wait(b1, phase1); phase1 ^= 1; }
        const long long t2 = clock64();
        tc_fence_after();
        if (t1 - t0 < best_issue) best_issue = t1 - t0;
        if (t2 - t0 < best_done) best_done = t2 - t0;
      }
      if (blockIdx.x == 0) { out[2 * ci] = best_issue; out[2 * ci + 1] = best_done; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc<256>(tmem); }
}

int main(int argc, char** argv) {
  const int ctas = argc > 1 ? atoi(argv[1]) : 1;            // 2: two co-resident CTAs on one SM run the same sequence
  std::vector<Case> cases;
  const char* pat[] = {"dependent", "2 accumulators", "4 accumulators", "column blocks", "attention S"};
  cases.push_back({0, 64, 0, DEP, 1});                       // commit + wait alone
  cases.push_back({0, 64, 0, DEP, 2});
  for (int ts = 0; ts < 2; ++ts)
    for (int N : {16, 64, 80, 128, 192})
      for (int n : {1, 4, 8, 16}) {
        if (N * 1 > 192) continue;
        cases.push_back({ts, N, n, DEP, 1});
      }
  for (int ts = 0; ts < 2; ++ts)
    for (int n : {4, 8, 16}) {
      cases.push_back({ts, 64, n, ALT2, 1});
      cases.push_back({ts, 32, n, ALT4, 1});
      cases.push_back({ts, 16, n, COLS, 1});
      cases.push_back({ts, 80, n, ALT2, 1});
    }
  cases.push_back({1, 64, 8, TILE_S, 2});
  cases.push_back({1, 64, 4, DEP, 2});
  cases.push_back({1, 80, 4, DEP, 2});
  Case* d_cases;
  long long* d_out;
  CK(cudaMalloc(&d_cases, cases.size() * sizeof(Case)));
  CK(cudaMalloc(&d_out, cases.size() * 2 * sizeof(long long)));
  CK(cudaMemcpy(d_cases, cases.data(), cases.size() * sizeof(Case), cudaMemcpyHostToDevice));
  CK(cudaFuncSetAttribute(bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
  // ctas == 2: launch 2 x #SMs CTAs so that every SM holds two (the hardware fills an SM before the next only when the
  // grid oversubscribes; both CTAs of SM 0 run the same sequence at the same time)
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const int grid = ctas == 2 ? 2 * sms : 1;
  for (int it = 0; it < 2; ++it) {
    bench_kernel<<<grid, 128, SMEM_TOTAL>>>(d_cases, static_cast<int>(cases.size()), d_out);
    CK(cudaDeviceSynchronize());
  }
  std::vector<long long> out(cases.size() * 2);
  CK(cudaMemcpy(out.data(), d_out, out.size() * sizeof(long long), cudaMemcpyDeviceToHost));
  printf("# M = 128, K = 16 per instruction, %s; cycles = min of %d repetitions\n",
         ctas == 2 ? "TWO co-resident CTAs per SM" : "one CTA alone", REPS);
  printf("%-4s %4s %3s %-15s %7s | %8s %8s %10s\n", "form", "N", "n", "pattern", "commits", "t_issue", "t_done", "done/instr");
  for (size_t i = 0; i < cases.size(); ++i) {
    const Case& c = cases[i];
    printf("%-4s %4d %3d %-15s %7d | %8lld %8lld %10.1f\n", c.ts ? "TS" : "SS", c.N, c.n, pat[c.pattern], c.commits,
           out[2 * i], out[2 * i + 1], c.n ? static_cast<double>(out[2 * i + 1]) / c.n : 0.0);
  }
  return 0;
}

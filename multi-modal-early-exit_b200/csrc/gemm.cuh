// Persistent warp-specialised tcgen05 GEMM for sm_100a:  C[M,N] = A[M,K] * W[N,K]^T (+ fused epilogue).
//
//   A  : activations, bf16 row-major [M_max, K]   (TMA, 128 x 64 boxes, SWIZZLE_128B)
//   W  : nn.Linear weight, bf16 row-major [N, K]  (TMA, BLOCK_N x 64 boxes, SWIZZLE_128B)
//   acc: fp32 in TMEM, 2 accumulator buffers of BLOCK_N columns (epilogue of tile i overlaps MMA of i+1)
//
// Roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer (one elected lane),
// warps 2..5 = epilogue (TMEM lane quarter = warp % 4).  M (the surviving-token count) is read from
// device memory so the grid never depends on a host round-trip: grid = #SMs, static round-robin tiles.
//
// Replaces the nn.Linear call sites of HF LayoutLMv3 used by the reference: query/key/value
// (HF modeling_layoutlmv3.py:245-259), SelfOutput.dense (:300-304), Intermediate.dense + GELU
// (:495-498), Output.dense (:509-513), and the patch-embedding conv as a GEMM (:70-82).
#pragma once
#include <cuda.h>

#include "ptx.cuh"

namespace mmee {

constexpr int GEMM_BLOCK_M = 128;
constexpr int GEMM_BLOCK_K = 64;
constexpr int GEMM_THREADS = 192;

enum GemmEpilogue : int {
  EPI_BIAS_BF16 = 0,   // out_bf16[m, n] = acc + bias[n]
  EPI_GELU_BF16 = 1,   // out_bf16[m, n] = gelu_erf(acc + bias[n])
  EPI_RESID_F32 = 2,   // out_f32 [m, n] = acc + bias[n] + resid_bf16[m, n]
  EPI_QKV = 3,         // n < qk_cols: qk_bf16[m, n];  else V^T: vt[doc][head][d][kv_pitch]
  EPI_PATCH = 4,       // patch-embed rows (doc*n_patch + p): out_f32[(doc*n_vis + 1 + p), n] = acc + bias + pos[1+p, n]
};

struct GemmArgs {
  const int* m_dev;        // device pointer to the dynamic row count (nullptr -> m_static)
  int m_static;
  int N, K;
  const float* bias;       // [N]
  void* out;               // bf16 or f32 [*, ld_out]
  int ld_out;
  const __nv_bfloat16* resid;   // EPI_RESID_F32: [*, N]
  // EPI_QKV
  __nv_bfloat16* vt;       // [docs][heads][64][kv_pitch]
  int qk_cols;             // 2*H
  int seq;                 // tokens per document (709)
  int kv_pitch;            // padded token pitch of V^T rows
  int heads;
  // EPI_PATCH
  const float* pos;        // pos_embed [n_vis, N]
  int n_patch, n_vis;
};

template <int BLOCK_N>
struct GemmSmem {
  static constexpr int STAGES = (BLOCK_N == 256) ? 4 : 6;
  static constexpr int A_BYTES = GEMM_BLOCK_M * GEMM_BLOCK_K * 2;
  static constexpr int B_BYTES = BLOCK_N * GEMM_BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BIAS_OFF = STAGES * STAGE_BYTES;
  static constexpr int BAR_OFF = BIAS_OFF + 2 * BLOCK_N * 4;
  static constexpr int TOTAL = BAR_OFF + (2 * STAGES + 4) * 8 + 16;
  static constexpr int DYN_BYTES = TOTAL + 1024;   // slack for manual 1024 B alignment
};

__device__ __forceinline__ float gelu_erf_fast(float x) {
  // x * Phi(x) with erf by Abramowitz-Stegun 7.1.26 (|err| < 1.5e-7): matches torch's exact GELU
  // (activations.py "gelu" -> nn.functional.gelu) far inside bf16 output rounding.
  const float z = fabsf(x) * 0.70710678118654752f;
  const float t = __frcp_rn(fmaf(0.3275911f, z, 1.0f));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  poly *= t;
  const float e = poly * __expf(-z * z);          // 1 - erf(z)
  const float phi = (x >= 0.f) ? (1.0f - 0.5f * e) : (0.5f * e);
  return x * phi;
}

template <int BLOCK_N, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const GemmArgs args) {
  using SM = GemmSmem<BLOCK_N>;
  constexpr int STAGES = SM::STAGES;
  constexpr uint32_t TMEM_COLS = (2 * BLOCK_N <= 256) ? 256 : 512;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* s_bias = reinterpret_cast<float*>(smem + SM::BIAS_OFF);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + SM::BAR_OFF);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int M = args.m_dev ? *args.m_dev : args.m_static;
  const int m_blocks = (M + GEMM_BLOCK_M - 1) / GEMM_BLOCK_M;
  const int n_blocks = args.N / BLOCK_N;
  const int k_blocks = args.K / GEMM_BLOCK_K;
  const int total_tiles = m_blocks * n_blocks;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 4);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_base_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int m0 = (tile / n_blocks) * GEMM_BLOCK_M;
        const int n0 = (tile % n_blocks) * BLOCK_N;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * SM::STAGE_BYTES;
          uint8_t* sb = sa + SM::A_BYTES;
          mbar_expect_tx(&full_bar[stage], SM::STAGE_BYTES);
          tma_load_2d(sa, &tmap_a, &full_bar[stage], kb * GEMM_BLOCK_K, m0);
          tma_load_2d(sb, &tmap_b, &full_bar[stage], kb * GEMM_BLOCK_K, n0);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(GEMM_BLOCK_M, BLOCK_N);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * SM::STAGE_BYTES);
          const uint32_t sb = sa + SM::A_BYTES;
          const uint64_t da = umma_desc_sw128_kmajor(sa);
          const uint64_t db = umma_desc_sw128_kmajor(sb);
#pragma unroll
          for (int k = 0; k < GEMM_BLOCK_K / 16; ++k) {
            // +32 B per K=16 step inside the 128 B swizzle row (descriptor address is in 16 B units)
            umma_bf16_ss(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue warps (2..5)
    const int quarter = warp & 3;                 // TMEM lane quarter this warp may access
    const int ep_tid = threadIdx.x - 64;          // 0..127
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int m0 = (tile / n_blocks) * GEMM_BLOCK_M;
      const int n0 = (tile % n_blocks) * BLOCK_N;
      float* sb = s_bias + acc * BLOCK_N;
      for (int i = ep_tid; i < BLOCK_N; i += 128) sb[i] = __ldg(args.bias + n0 + i);
      asm volatile("bar.sync 1, 128;" ::: "memory");

      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const int row = m0 + quarter * 32 + lane;
      const bool row_ok = row < M;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BLOCK_N;

#pragma unroll 1
      for (int c = 0; c < BLOCK_N; c += 32) {
        uint32_t v[32];
        tmem_ld32(taddr + c, v);
        tmem_ld_wait();
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]) + sb[c + j];
        const int n = n0 + c;
        if constexpr (EPI == EPI_BIAS_BF16 || EPI == EPI_GELU_BF16) {
          if (row_ok) {
            uint4* dst = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(args.out) +
                                                  static_cast<size_t>(row) * args.ld_out + n);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              float g[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                g[j] = (EPI == EPI_GELU_BF16) ? gelu_erf_fast(f[q * 8 + j]) : f[q * 8 + j];
              }
              dst[q] = make_uint4(pack_bf16x2(g[0], g[1]), pack_bf16x2(g[2], g[3]), pack_bf16x2(g[4], g[5]),
                                  pack_bf16x2(g[6], g[7]));
            }
          }
        } else if constexpr (EPI == EPI_RESID_F32) {
          if (row_ok) {
            const uint4* rs = reinterpret_cast<const uint4*>(args.resid + static_cast<size_t>(row) * args.N + n);
            float4* dst = reinterpret_cast<float4*>(static_cast<float*>(args.out) +
                                                    static_cast<size_t>(row) * args.ld_out + n);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const uint4 r = __ldg(rs + q);
              const float2 r0 = unpack_bf16x2(r.x), r1 = unpack_bf16x2(r.y), r2 = unpack_bf16x2(r.z),
                           r3 = unpack_bf16x2(r.w);
              dst[2 * q] = make_float4(f[q * 8 + 0] + r0.x, f[q * 8 + 1] + r0.y, f[q * 8 + 2] + r1.x,
                                       f[q * 8 + 3] + r1.y);
              dst[2 * q + 1] = make_float4(f[q * 8 + 4] + r2.x, f[q * 8 + 5] + r2.y, f[q * 8 + 6] + r3.x,
                                           f[q * 8 + 7] + r3.y);
            }
          }
        } else if constexpr (EPI == EPI_QKV) {
          if (n < args.qk_cols) {
            if (row_ok) {
              uint4* dst = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(args.out) +
                                                    static_cast<size_t>(row) * args.ld_out + n);
#pragma unroll
              for (int q = 0; q < 4; ++q)
                dst[q] = make_uint4(pack_bf16x2(f[q * 8 + 0], f[q * 8 + 1]), pack_bf16x2(f[q * 8 + 2], f[q * 8 + 3]),
                                    pack_bf16x2(f[q * 8 + 4], f[q * 8 + 5]), pack_bf16x2(f[q * 8 + 6], f[q * 8 + 7]));
            }
          } else if (row_ok) {
            // V is stored transposed per (doc, head): vt[d][token] so that P*V takes a K-major B operand.
            const int doc = row / args.seq;
            const int tok = row - doc * args.seq;
            const int nv = n - args.qk_cols;        // 32-aligned => one head per chunk
            const int head = nv >> 6;
            const int d0 = nv & 63;
            __nv_bfloat16* dst = args.vt +
                (static_cast<size_t>(doc * args.heads + head) * 64 + d0) * args.kv_pitch + tok;
#pragma unroll
            for (int j = 0; j < 32; ++j) dst[static_cast<size_t>(j) * args.kv_pitch] = __float2bfloat16_rn(f[j]);
          }
        } else if constexpr (EPI == EPI_PATCH) {
          if (row_ok) {
            const int doc = row / args.n_patch;
            const int p = row - doc * args.n_patch;
            const float4* ps = reinterpret_cast<const float4*>(args.pos + static_cast<size_t>(1 + p) * args.N + n);
            float4* dst = reinterpret_cast<float4*>(static_cast<float*>(args.out) +
                (static_cast<size_t>(doc) * args.n_vis + 1 + p) * args.ld_out + n);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 pe = __ldg(ps + q);
              dst[q] = make_float4(f[q * 4 + 0] + pe.x, f[q * 4 + 1] + pe.y, f[q * 4 + 2] + pe.z, f[q * 4 + 3] + pe.w);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

}  // namespace mmee

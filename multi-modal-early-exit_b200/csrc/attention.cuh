// Fused attention for LayoutLMv3 (HF modeling_layoutlmv3.py:236-289):
//     ctx = softmax( (Q/8) K^T + (rel_pos + rel_2d_pos)/8 + key_mask ) V
// One CTA = one (document slot, head, 128-query tile).  QK^T and PV run on tcgen05 with fp32 accumulators
// in TMEM; the [S,S] score matrix never leaves the SM.  The additive bias (1-D + 2-D relative position
// buckets and the key-padding mask, layer-invariant, built once per forward as fp16) is streamed tile by
// tile with TMA and added in registers before the online softmax.  The CogView "PB-relax" form
// softmax((s/32 - max(s/32))*32) of HF:224-234 is the standard max-shifted softmax.
//
// Roles (320 threads): warp 0 TMA producer, warp 1 MMA issuer (+TMEM owner), warps 2-9 softmax /
// accumulate: thread = (query row = TMEM lane, half of the tile's 128 keys / half of the 64 output dims);
// the two threads of a row exchange their partial row max through smem once per tile.  Per KV tile j:
//   S[j%2]  = Q K_j^T                       (MMA, 128x128x64)
//   softmax: s = S + bias  -> running max m, P_j = exp2((s - m) log2e) (bf16, smem, SW128), l
//   O'[j%2] = P_j V_j                       (MMA, 128x64x128, fresh accumulator)
//   O_reg   = alpha_j * O_reg + O'[j%2]     (registers; alpha_j = exp2((m_{j-1} - m_j) log2e))
// Q is pre-scaled by 1/sqrt(d) (folded into W_q, exact power of two); V is stored transposed per
// (doc, head) by the QKV GEMM epilogue so P*V takes a K-major B operand.
#pragma once
#include <cuda.h>

#include "ptx.cuh"

namespace mmee {

constexpr int ATT_SM_WARPS = 8;      // softmax warps: 2 per TMEM lane quarter (key-column halves)
constexpr int ATT_THREADS = 64 + ATT_SM_WARPS * 32;
constexpr int ATT_BQ = 128;    // query rows per CTA
constexpr int ATT_BKV = 128;   // keys per tile
constexpr int ATT_D = 64;

struct AttSmem {
  static constexpr int Q_BYTES = ATT_BQ * ATT_D * 2;           // 16 KB
  static constexpr int K_BYTES = ATT_BKV * ATT_D * 2;          // 16 KB
  static constexpr int V_BYTES = ATT_D * ATT_BKV * 2;          // 16 KB  (two [64 x 64] sub-tiles)
  static constexpr int B_BYTES = ATT_BQ * ATT_BKV * 2;         // 32 KB  (two [128 x 64] fp16 sub-tiles)
  static constexpr int P_BYTES = ATT_BQ * ATT_BKV * 2;         // 32 KB  (two [128 x 64] bf16 sub-tiles)
  static constexpr int KV_STAGE = K_BYTES + V_BYTES + B_BYTES; // 64 KB
  static constexpr int Q_OFF = 0;                              // 2 Q buffers
  static constexpr int KV_OFF = Q_OFF + 2 * Q_BYTES;
  static constexpr int P_OFF = KV_OFF + 2 * KV_STAGE;
  static constexpr int X_OFF = P_OFF + P_BYTES;                // P is single-buffered (see softmax loop)            // row-max [2][2][128] + row-sum [2][128] exchange (floats)
  static constexpr int BAR_OFF = X_OFF + 3 * 2 * ATT_BQ * 4;
  static constexpr int N_BARS = 16;    // q_full, q_empty, kv_full, kv_empty, s_full, p_full, o_full, o_empty (x2 each)
  static constexpr int TOTAL = BAR_OFF + N_BARS * 8 + 16;
  static constexpr int DYN_BYTES = TOTAL + 1024;
};

struct AttArgs {
  const int* n_active_dev;
  const int* slot_doc;         // slot -> original document (bias is indexed by document)
  __nv_bfloat16* ctx;          // [M, H]
  int H, heads, seq;
};

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// tmap_qk : bf16 [M_max, 2H]                    box [128 x 64]
// tmap_vt : bf16 [docs*heads*64, kv_pitch]       box [64 x 64]
// tmap_bias: fp16 [docs*heads*seq, bias_pitch]   box [128 x 64]
//
// Persistent: grid = #SMs; work item = (slot, head, q-tile), q-tile fastest so the CTAs running at the same
// time share K/V in L2.  Every pipeline (Q double buffer, K/V/bias double buffer, S and O' double buffers in
// TMEM) runs straight through item boundaries, so the next item's loads and first QK^T overlap the current
// item's last softmax / store.
__global__ void __launch_bounds__(ATT_THREADS, 1)
attention_kernel(const __grid_constant__ CUtensorMap tmap_qk, const __grid_constant__ CUtensorMap tmap_vt,
                 const __grid_constant__ CUtensorMap tmap_bias, const AttArgs args) {
  const int S = args.seq;
  const int n_kv = (S + ATT_BKV - 1) / ATT_BKV;
  const int n_qt = (S + ATT_BQ - 1) / ATT_BQ;
  const int total_items = *args.n_active_dev * args.heads * n_qt;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + AttSmem::BAR_OFF);
  uint64_t* q_full = bars;          // [2]
  uint64_t* q_empty = bars + 2;     // [2]
  uint64_t* kv_full = bars + 4;
  uint64_t* kv_empty = bars + 6;
  uint64_t* s_full = bars + 8;
  uint64_t* p_full = bars + 10;
  uint64_t* o_full = bars + 12;
  uint64_t* o_empty = bars + 14;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + AttSmem::N_BARS);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_qk);
    tma_prefetch_desc(&tmap_vt);
    tma_prefetch_desc(&tmap_bias);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 1);
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1 + ATT_SM_WARPS);   // MMA commit after PV + softmax warps done with the bias tile
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], ATT_SM_WARPS);
      mbar_init(&o_full[i], 1);
      mbar_init(&o_empty[i], ATT_SM_WARPS);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base;            // 2 x 128 columns
  const uint32_t tmem_O = tmem_base + 256;      // 2 x 64 columns

  auto decode = [&](int item, int& slot, int& head, int& q0) {
    const int qt = item % n_qt;
    const int sh = item / n_qt;
    head = sh % args.heads;
    slot = sh / args.heads;
    q0 = qt * ATT_BQ;
  };

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int ii = 0;
      uint32_t t = 0;                               // global KV-tile counter of this CTA
      for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++ii) {
        int slot, head, q0;
        decode(item, slot, head, q0);
        const int doc = args.slot_doc[slot];
        const int row0 = slot * S;
        const int qb = ii & 1;
        mbar_wait(&q_empty[qb], ((ii >> 1) & 1) ^ 1);
        mbar_expect_tx(&q_full[qb], AttSmem::Q_BYTES);
        tma_load_2d(smem + AttSmem::Q_OFF + qb * AttSmem::Q_BYTES, &tmap_qk, &q_full[qb], head * ATT_D, row0 + q0);
        const int vt_row = (slot * args.heads + head) * ATT_D;
        const int bias_row = (doc * args.heads + head) * S + q0;
        for (int j = 0; j < n_kv; ++j, ++t) {
          const int st = t & 1;
          mbar_wait(&kv_empty[st], ((t >> 1) & 1) ^ 1);
          uint8_t* sk = smem + AttSmem::KV_OFF + st * AttSmem::KV_STAGE;
          uint8_t* sv = sk + AttSmem::K_BYTES;
          uint8_t* sb = sv + AttSmem::V_BYTES;
          const int kv0 = j * ATT_BKV;
          mbar_expect_tx(&kv_full[st], AttSmem::KV_STAGE);
          tma_load_2d(sk, &tmap_qk, &kv_full[st], args.H + head * ATT_D, row0 + kv0);
          tma_load_2d(sv, &tmap_vt, &kv_full[st], kv0, vt_row);
          tma_load_2d(sv + AttSmem::V_BYTES / 2, &tmap_vt, &kv_full[st], kv0 + 64, vt_row);
          tma_load_2d(sb, &tmap_bias, &kv_full[st], kv0, bias_row);
          tma_load_2d(sb + AttSmem::B_BYTES / 2, &tmap_bias, &kv_full[st], kv0 + 64, bias_row);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(ATT_BQ, ATT_BKV);
      constexpr uint32_t idesc_o = umma_idesc_bf16(ATT_BQ, ATT_D);
      const int my_items = (total_items > static_cast<int>(blockIdx.x))
                               ? (total_items - 1 - static_cast<int>(blockIdx.x)) / static_cast<int>(gridDim.x) + 1 : 0;
      const uint32_t total_t = static_cast<uint32_t>(my_items) * n_kv;
      auto issue_s = [&](uint32_t t) {               // S[t&1] = Q_item K_t^T
        const int st = t & 1;
        const int ii = t / n_kv, j = t - ii * n_kv;
        const int qb = ii & 1;
        if (j == 0) { mbar_wait(&q_full[qb], (ii >> 1) & 1); }
        mbar_wait(&kv_full[st], (t >> 1) & 1);
        tc_fence_after();
        const uint64_t dq = umma_desc_sw128_kmajor(smem_u32(smem + AttSmem::Q_OFF + qb * AttSmem::Q_BYTES));
        const uint64_t dk = umma_desc_sw128_kmajor(smem_u32(smem + AttSmem::KV_OFF + st * AttSmem::KV_STAGE));
#pragma unroll
        for (int k = 0; k < ATT_D / 16; ++k)
          umma_bf16_ss(tmem_S + st * ATT_BKV, dq + 2 * k, dk + 2 * k, idesc_s, k ? 1u : 0u);
        umma_commit(&s_full[st]);
        if (j == n_kv - 1) umma_commit(&q_empty[qb]);   // last use of this item's Q
      };
      if (total_t) issue_s(0);
      for (uint32_t t = 0; t < total_t; ++t) {
        const int st = t & 1;
        if (t + 1 < total_t) issue_s(t + 1);
        mbar_wait(&p_full[st], (t >> 1) & 1);
        mbar_wait(&o_empty[st], ((t >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t sp = smem_u32(smem + AttSmem::P_OFF);
        const uint32_t sv = smem_u32(smem + AttSmem::KV_OFF + st * AttSmem::KV_STAGE + AttSmem::K_BYTES);
#pragma unroll
        for (int k = 0; k < ATT_BKV / 16; ++k) {
          const int hf = k >> 2, kk = k & 3;
          const uint64_t dp = umma_desc_sw128_kmajor(sp + hf * (AttSmem::P_BYTES / 2)) + 2 * kk;
          const uint64_t dv = umma_desc_sw128_kmajor(sv + hf * (AttSmem::V_BYTES / 2)) + 2 * kk;
          umma_bf16_ss(tmem_O + st * ATT_D, dp, dv, idesc_o, k ? 1u : 0u);
        }
        umma_commit(&o_full[st]);
        umma_commit(&kv_empty[st]);
      }
    }
  } else {
    // ------------------------------------------------------------ softmax + accumulate (warps 2..9)
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;                         // key columns [64*half, +64), output dims [32*half, +32)
    const int r = quarter * 32 + lane;                        // query row within the tile == TMEM lane
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    const float LOG2E = 1.4426950408889634f;
    float* xch = reinterpret_cast<float*>(smem + AttSmem::X_OFF);     // [buf][half][row]
    float* xl = xch + 2 * 2 * ATT_BQ;                                 // [half][row] row-sum exchange
    constexpr int OD = ATT_D / 2;
    const uint32_t sw = static_cast<uint32_t>(r & 7);
    uint32_t t = 0;

    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      int slot, head, q0;
      decode(item, slot, head, q0);
      const int row0 = slot * S;
      float o_acc[OD];
#pragma unroll
      for (int i = 0; i < OD; ++i) o_acc[i] = 0.f;
      float m_run = -INFINITY, l_run = 0.f, alpha_prev = 1.f;

      auto accumulate = [&](uint32_t tt, float alpha) {         // O_reg = alpha * O_reg + O'[tt]
        const int st = tt & 1;
        mbar_wait(&o_full[st], (tt >> 1) & 1);
        tc_fence_after();
        uint32_t v[32];
        tmem_ld32(tmem_O + lane_addr + st * ATT_D + half * OD, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < OD; ++i) o_acc[i] = fmaf(o_acc[i], alpha, __uint_as_float(v[i]));
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&o_empty[st]);
      };

      for (int j = 0; j < n_kv; ++j, ++t) {
        const int st = t & 1;
        const int kv0 = j * ATT_BKV;
        const uint8_t* sb = smem + AttSmem::KV_OFF + st * AttSmem::KV_STAGE + AttSmem::K_BYTES + AttSmem::V_BYTES +
                            half * (AttSmem::B_BYTES / 2) + r * 128;
        uint8_t* sp = smem + AttSmem::P_OFF + half * (AttSmem::P_BYTES / 2) + r * 128;
        mbar_wait(&s_full[st], (t >> 1) & 1);        // S_t done  (=> kv_full[st] landed: the MMA waited on it)
        tc_fence_after();
        const uint32_t ts = tmem_S + lane_addr + st * ATT_BKV + half * 64;
        const bool tail = (kv0 + ATT_BKV > S);

        // ---- pass 1: s = S + bias, partial row max, write s back to TMEM
        float tmax = -INFINITY;
#pragma unroll
        for (int c2 = 0; c2 < 2; ++c2) {
          uint32_t v[32];
          tmem_ld32(ts + c2 * 32, v);
          uint4 b4[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) b4[q] = *reinterpret_cast<const uint4*>(sb + (((c2 * 4 + q) ^ sw) << 4));
          tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint32_t w[4] = {b4[q].x, b4[q].y, b4[q].z, b4[q].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 bf = __half22float2(*reinterpret_cast<const __half2*>(&w[e]));
              const int i = q * 8 + e * 2;
              float s0 = __uint_as_float(v[i]) + bf.x;
              float s1 = __uint_as_float(v[i + 1]) + bf.y;
              if (tail) {
                const int col = kv0 + half * 64 + c2 * 32 + i;
                if (col >= S) s0 = -INFINITY;
                if (col + 1 >= S) s1 = -INFINITY;
              }
              tmax = fmaxf(tmax, fmaxf(s0, s1));
              v[i] = __float_as_uint(s0);
              v[i + 1] = __float_as_uint(s1);
            }
          }
          tmem_st32(ts + c2 * 32, v);
        }
        xch[(st * 2 + half) * ATT_BQ + r] = tmax;
        tmem_st_wait();
        __syncwarp();
        if (lane == 0) mbar_arrive(&kv_empty[st]);   // bias tile consumed
        asm volatile("bar.sync 1, 256;" ::: "memory");
        tmax = fmaxf(tmax, xch[(st * 2 + (half ^ 1)) * ATT_BQ + r]);

        const float m_new = fmaxf(m_run, tmax);
        const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
        const float alpha = fast_exp2((m_run - m_use) * LOG2E);     // m_run = -inf -> 0
        const float neg_m = -m_use * LOG2E;

        // O'[t-1] -> registers.  Its o_full wait also proves P V_{t-1} has finished reading the single P buffer.
        if (j > 0) accumulate(t - 1, alpha_prev);
        alpha_prev = alpha;

        // ---- pass 2: p = exp2(s*log2e - m*log2e), partial row sum, bf16 P -> smem (SW128 K-major A operand)
        float psum = 0.f;
#pragma unroll
        for (int c2 = 0; c2 < 2; ++c2) {
          uint32_t v[32];
          tmem_ld32(ts + c2 * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float p[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              p[e] = fast_exp2(fmaf(__uint_as_float(v[q * 8 + e]), LOG2E, neg_m));
              psum += p[e];
            }
            *reinterpret_cast<uint4*>(sp + (((c2 * 4 + q) ^ sw) << 4)) =
                make_uint4(pack_bf16x2(p[0], p[1]), pack_bf16x2(p[2], p[3]), pack_bf16x2(p[4], p[5]),
                           pack_bf16x2(p[6], p[7]));
          }
        }
        l_run = l_run * alpha + psum;
        m_run = m_new;
        fence_proxy_async_smem();                     // P visible to the tensor core (async proxy)
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[st]);
      }
      accumulate(t - 1, alpha_prev);

      // ---- combine the two partial row sums, normalise and store ctx[row, head*64 + 32*half .. +31]
      xl[half * ATT_BQ + r] = l_run;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const float l_tot = l_run + xl[(half ^ 1) * ATT_BQ + r];
      asm volatile("bar.sync 1, 256;" ::: "memory");      // xl may be rewritten by the next item
      const int q = q0 + r;
      if (q < S) {
        const float inv = 1.0f / l_tot;
        uint4* dst = reinterpret_cast<uint4*>(args.ctx + static_cast<size_t>(row0 + q) * args.H + head * ATT_D + half * OD);
#pragma unroll
        for (int i = 0; i < OD / 8; ++i)
          dst[i] = make_uint4(pack_bf16x2(o_acc[i * 8 + 0] * inv, o_acc[i * 8 + 1] * inv),
                              pack_bf16x2(o_acc[i * 8 + 2] * inv, o_acc[i * 8 + 3] * inv),
                              pack_bf16x2(o_acc[i * 8 + 4] * inv, o_acc[i * 8 + 5] * inv),
                              pack_bf16x2(o_acc[i * 8 + 6] * inv, o_acc[i * 8 + 7] * inv));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace mmee

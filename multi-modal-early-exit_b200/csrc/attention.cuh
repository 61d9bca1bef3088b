// Fused attention for LayoutLMv3 (HF modeling_layoutlmv3.py:236-289):
//     ctx = softmax( (Q/8) K^T + (rel_pos + rel_2d_pos)/8 + key_mask ) V
// Persistent kernel, grid = #SMs, work item = (document slot, head, 128-query tile).  QK^T and PV run on
// tcgen05 with fp32 accumulators in TMEM; the [S,S] score matrix never leaves the SM.  The additive bias
// (1-D + 2-D relative position buckets; layer-invariant, built once per forward as uint8 with a per-head
// scale, see bias_build_kernel) is streamed tile by tile with TMA and added in registers before the online
// softmax.  The CogView "PB-relax" form softmax((s/32 - max(s/32))*32) of HF:224-234 is the standard
// max-shifted softmax.  The key-padding mask is a per-(doc, key-tile) flag: fully padded tiles are skipped,
// mixed tiles add 0/-inf per key.
//
// Roles (576 threads): warp 0 TMA producer, warp 1 MMA issuer (+TMEM owner), warps 2-17 softmax /
// accumulate: thread = (query row = TMEM lane, a quarter of the tile's 128 keys and of the 64 output dims);
// the four threads of a row exchange their partial row max through smem once per tile; scores stay in
// registers between the max and the exp pass.  Per KV tile t (all in the log2 domain: log2(e)/sqrt(d) is
// folded into W_q):
//   S[t%2]  = Q K_t^T                       (MMA, 128x128x64)
//   softmax: s = S + bias  -> running max m, P_t = exp2(s - m) (bf16, smem, SW128), l
//   O'[t%2] = P_t V_t                       (MMA, 128x64x128, fresh accumulator)
//   O_reg   = alpha_t * O_reg + O'[t%2]     (registers; alpha_t = exp2(m_{t-1} - m_t))
// K/V/bias tiles go through a 3-stage TMA ring; the next item's bias tiles are prefetched into L2 while the
// current item runs.  V is stored transposed per (doc, head) by the QKV GEMM epilogue so P*V takes a K-major
// B operand.
#pragma once
#include <cuda.h>

#include "ptx.cuh"

namespace mmee {

constexpr int ATT_NSPLIT = 4;        // softmax threads per query row (each owns 32 of the tile's 128 keys)
constexpr int ATT_SM_WARPS = 4 * ATT_NSPLIT;
constexpr int ATT_THREADS = 64 + ATT_SM_WARPS * 32;
constexpr int ATT_BQ = 128;    // query rows per CTA
constexpr int ATT_BKV = 128;   // keys per tile
constexpr int ATT_D = 64;
constexpr int ATT_STAGES = 3;

struct AttSmem {
  static constexpr int Q_BYTES = ATT_BQ * ATT_D * 2;           // 16 KB
  static constexpr int K_BYTES = ATT_BKV * ATT_D * 2;          // 16 KB
  static constexpr int V_BYTES = ATT_D * ATT_BKV * 2;          // 16 KB  (two [64 x 64] sub-tiles)
  static constexpr int B_BYTES = ATT_BQ * ATT_BKV;             // 16 KB  uint8 [128 x 128]
  static constexpr int P_BYTES = ATT_BQ * ATT_BKV * 2;         // 32 KB  (two [128 x 64] bf16 sub-tiles)
  static constexpr int KV_STAGE = K_BYTES + V_BYTES + B_BYTES; // 48 KB
  static constexpr int Q_OFF = 0;                              // 2 Q buffers
  static constexpr int KV_OFF = Q_OFF + 2 * Q_BYTES;
  static constexpr int P_OFF = KV_OFF + ATT_STAGES * KV_STAGE;
  static constexpr int X_OFF = P_OFF + P_BYTES;                // row-max [2][NSPLIT][128] + row-sum [NSPLIT][128]
  static constexpr int BAR_OFF = X_OFF + 3 * ATT_NSPLIT * ATT_BQ * 4;
  static constexpr int N_BARS = 2 + 2 + 2 * ATT_STAGES + 2 + 2 + 2 + 2;
  static constexpr int TOTAL = BAR_OFF + N_BARS * 8 + 16;
  static constexpr int DYN_BYTES = TOTAL + 1024;
};
static_assert(AttSmem::DYN_BYTES <= 232448, "attention smem budget");

struct AttArgs {
  const int* n_active_dev;
  const int* slot_doc;         // slot -> original document (bias / mask are indexed by document)
  const int* tileflag;         // [docs][n_kv] 0 none masked, 1 some, 2 all
  const float* maskadd;        // [docs][kv_pitch] 0 / -inf
  const float* bias_scale2;    // [heads] scale_h * log2(e)
  __nv_bfloat16* ctx;          // [M, H]
  int H, heads, seq, kv_pitch;
};

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// byte K of w -> float(byte) - 128, exactly: PRMT builds 0x4B0000bb = 2^23 + bb, one FADD removes the offset.
template <int K>
__device__ __forceinline__ float u8_to_centered_float(uint32_t w) {
  return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7540 | K)) - 8388736.0f;
}

// Position in this CTA's (item, key tile) sequence; fully masked tiles are skipped.  Every role walks the same
// sequence so the pipeline counters stay in lock-step.
struct AttCursor {
  int item, ii, j, slot, head, q0, doc, first_j, last_j;
  bool valid;
};

// tmap_qk : bf16 [M_max, 2H]                    box [128 x 64]
// tmap_vt : bf16 [docs*heads*64, kv_pitch]       box [64 x 64]
// tmap_bias: u8  [docs*heads*seq, bias_pitch]    box [128 x 128]
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
  uint64_t* q_full = bars;                         // [2]
  uint64_t* q_empty = q_full + 2;                  // [2]
  uint64_t* kv_full = q_empty + 2;                 // [STAGES]
  uint64_t* kv_empty = kv_full + ATT_STAGES;       // [STAGES]
  uint64_t* s_full = kv_empty + ATT_STAGES;        // [2]
  uint64_t* p_full = s_full + 2;                   // [2]
  uint64_t* o_full = p_full + 2;                   // [2]
  uint64_t* o_empty = o_full + 2;                  // [2]
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
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], ATT_SM_WARPS);
      mbar_init(&o_full[i], 1);
      mbar_init(&o_empty[i], ATT_SM_WARPS);
    }
    for (int i = 0; i < ATT_STAGES; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1 + ATT_SM_WARPS);   // MMA commit after PV + softmax warps done with the bias tile
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

  // ---- shared cursor logic
  auto enter_item = [&](AttCursor& c) {                   // fills the item fields; valid = false when out of items
    if (c.item >= total_items) { c.valid = false; return; }
    const int qt = c.item % n_qt;
    const int sh = c.item / n_qt;
    c.head = sh % args.heads;
    c.slot = sh / args.heads;
    c.q0 = qt * ATT_BQ;
    c.doc = args.slot_doc[c.slot];
    c.first_j = -1;
    c.last_j = -1;
    for (int j = 0; j < n_kv; ++j)
      if (args.tileflag[c.doc * n_kv + j] != 2) { if (c.first_j < 0) c.first_j = j; c.last_j = j; }
    if (c.first_j < 0) { c.first_j = 0; c.last_j = 0; }   // degenerate: keep one tile so the row sum is defined
    c.j = c.first_j;
    c.valid = true;
  };
  auto start = [&](AttCursor& c) {
    c.item = blockIdx.x;
    c.ii = 0;
    enter_item(c);
  };
  auto advance = [&](AttCursor& c) {                       // next valid (item, tile)
    while (true) {
      ++c.j;
      if (c.j > c.last_j) {
        c.item += gridDim.x;
        ++c.ii;
        enter_item(c);
        return;
      }
      if (args.tileflag[c.doc * n_kv + c.j] != 2) return;
    }
  };

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      AttCursor c;
      start(c);
      uint32_t t = 0;
      int loaded_ii = -1;
      while (c.valid) {
        const int row0 = c.slot * S;
        if (c.ii != loaded_ii) {
          loaded_ii = c.ii;
          const int qb = c.ii & 1;
          mbar_wait(&q_empty[qb], ((c.ii >> 1) & 1) ^ 1);
          mbar_expect_tx(&q_full[qb], AttSmem::Q_BYTES);
          tma_load_2d(smem + AttSmem::Q_OFF + qb * AttSmem::Q_BYTES, &tmap_qk, &q_full[qb], c.head * ATT_D, row0 + c.q0);
          // pull the NEXT item's bias tiles (the only operand that comes from DRAM) into L2 ahead of time
          const int nitem = c.item + gridDim.x;
          if (nitem < total_items) {
            const int nsh = nitem / n_qt;
            const int ndoc = args.slot_doc[nsh / args.heads];
            const int nrow = (ndoc * args.heads + nsh % args.heads) * S + (nitem % n_qt) * ATT_BQ;
            for (int j = 0; j < n_kv; ++j)
              if (args.tileflag[ndoc * n_kv + j] != 2) tma_prefetch_2d(&tmap_bias, j * ATT_BKV, nrow);
          }
        }
        const int st = t % ATT_STAGES;
        mbar_wait(&kv_empty[st], ((t / ATT_STAGES) & 1) ^ 1);
        uint8_t* sk = smem + AttSmem::KV_OFF + st * AttSmem::KV_STAGE;
        uint8_t* sv = sk + AttSmem::K_BYTES;
        uint8_t* sb = sv + AttSmem::V_BYTES;
        const int kv0 = c.j * ATT_BKV;
        const int vt_row = (c.slot * args.heads + c.head) * ATT_D;
        const int bias_row = (c.doc * args.heads + c.head) * S + c.q0;
        mbar_expect_tx(&kv_full[st], AttSmem::KV_STAGE);
        tma_load_2d(sk, &tmap_qk, &kv_full[st], args.H + c.head * ATT_D, row0 + kv0);
        tma_load_2d(sv, &tmap_vt, &kv_full[st], kv0, vt_row);
        tma_load_2d(sv + AttSmem::V_BYTES / 2, &tmap_vt, &kv_full[st], kv0 + 64, vt_row);
        tma_load_2d(sb, &tmap_bias, &kv_full[st], kv0, bias_row);
        ++t;
        advance(c);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(ATT_BQ, ATT_BKV);
      constexpr uint32_t idesc_o = umma_idesc_bf16(ATT_BQ, ATT_D);
      AttCursor cs;                                   // cursor of the next S = Q K^T to issue (runs one tile ahead)
      start(cs);
      uint32_t ts = 0;
      auto issue_s = [&]() {
        const int st = ts % ATT_STAGES;
        const int qb = cs.ii & 1;
        if (cs.j == cs.first_j) mbar_wait(&q_full[qb], (cs.ii >> 1) & 1);
        mbar_wait(&kv_full[st], (ts / ATT_STAGES) & 1);
        tc_fence_after();
        const uint64_t dq = umma_desc_sw128_kmajor(smem_u32(smem + AttSmem::Q_OFF + qb * AttSmem::Q_BYTES));
        const uint64_t dk = umma_desc_sw128_kmajor(smem_u32(smem + AttSmem::KV_OFF + st * AttSmem::KV_STAGE));
#pragma unroll
        for (int k = 0; k < ATT_D / 16; ++k)
          umma_bf16_ss(tmem_S + (ts & 1) * ATT_BKV, dq + 2 * k, dk + 2 * k, idesc_s, k ? 1u : 0u);
        umma_commit(&s_full[ts & 1]);
        if (cs.j == cs.last_j) umma_commit(&q_empty[qb]);   // last use of this item's Q
        ++ts;
        advance(cs);
      };
      if (cs.valid) issue_s();
      for (uint32_t t = 0; t < ts; ++t) {               // ts grows while tiles remain
        if (cs.valid) issue_s();
        const int st = t % ATT_STAGES;
        const int b = t & 1;
        mbar_wait(&p_full[b], (t >> 1) & 1);
        mbar_wait(&o_empty[b], ((t >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t sp = smem_u32(smem + AttSmem::P_OFF);
        const uint32_t sv = smem_u32(smem + AttSmem::KV_OFF + st * AttSmem::KV_STAGE + AttSmem::K_BYTES);
#pragma unroll
        for (int k = 0; k < ATT_BKV / 16; ++k) {
          const int hf = k >> 2, kk = k & 3;
          const uint64_t dp = umma_desc_sw128_kmajor(sp + hf * (AttSmem::P_BYTES / 2)) + 2 * kk;
          const uint64_t dv = umma_desc_sw128_kmajor(sv + hf * (AttSmem::V_BYTES / 2)) + 2 * kk;
          umma_bf16_ss(tmem_O + b * ATT_D, dp, dv, idesc_o, k ? 1u : 0u);
        }
        umma_commit(&o_full[b]);
        umma_commit(&kv_empty[st]);
      }
    }
  } else {
    // ------------------------------------------------------------ softmax + accumulate (warps 2..17)
    const int quarter = warp & 3;
    const int part = (warp - 2) >> 2;                         // key columns [32*part, +32), output dims [16*part, +16)
    const int r = quarter * 32 + lane;                        // query row within the tile == TMEM lane
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    float* xch = reinterpret_cast<float*>(smem + AttSmem::X_OFF);     // [buf][part][row]
    float* xl = xch + 2 * ATT_NSPLIT * ATT_BQ;                        // [part][row] row-sum exchange
    constexpr int OD = ATT_D / ATT_NSPLIT;                            // 16
    constexpr int KC = ATT_BKV / ATT_NSPLIT;                          // 32
    const uint32_t sw = static_cast<uint32_t>(r & 7);
    uint32_t t = 0;
    AttCursor c;
    start(c);

    while (c.valid) {
      const int row0 = c.slot * S;
      const int q0 = c.q0, head = c.head, doc = c.doc, my_ii = c.ii;
      const float scale2 = __ldg(args.bias_scale2 + head);
      float o_acc[OD];
#pragma unroll
      for (int i = 0; i < OD; ++i) o_acc[i] = 0.f;
      float m_run = -INFINITY, l_run = 0.f, alpha_prev = 1.f;
      int nproc = 0;

      auto accumulate = [&](uint32_t tt, float alpha) {         // O_reg = alpha * O_reg + O'[tt]
        const int b = tt & 1;
        mbar_wait(&o_full[b], (tt >> 1) & 1);
        tc_fence_after();
        uint32_t v[OD];
        tmem_ld16(tmem_O + lane_addr + b * ATT_D + part * OD, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < OD; ++i) o_acc[i] = fmaf(o_acc[i], alpha, __uint_as_float(v[i]));
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&o_empty[b]);
      };

      while (c.valid && c.ii == my_ii) {
        const int st = t % ATT_STAGES;
        const int b = t & 1;
        const int kv0 = c.j * ATT_BKV;
        const int flag = args.tileflag[doc * n_kv + c.j];
        const uint8_t* sb = smem + AttSmem::KV_OFF + st * AttSmem::KV_STAGE + AttSmem::K_BYTES + AttSmem::V_BYTES + r * 128;
        uint8_t* sp = smem + AttSmem::P_OFF + (part >> 1) * (AttSmem::P_BYTES / 2) + r * 128;
        mbar_wait(&s_full[b], (t >> 1) & 1);         // S_t done  (=> kv_full[st] landed: the MMA waited on it)
        tc_fence_after();

        // ---- s = S + bias (registers, log2 domain), partial row max
        uint32_t v[KC];
        tmem_ld32(tmem_S + lane_addr + b * ATT_BKV + part * KC, v);
        const uint4 ba = *reinterpret_cast<const uint4*>(sb + (((part * 2) ^ sw) << 4));        // keys 32*part .. +15
        const uint4 bb = *reinterpret_cast<const uint4*>(sb + (((part * 2 + 1) ^ sw) << 4));    // keys +16 .. +31
        tmem_ld_wait();
        float sc[KC];
#define MMEE_BIAS4(W, BASE)                                                                         \
  sc[BASE + 0] = fmaf(u8_to_centered_float<0>(W), scale2, __uint_as_float(v[BASE + 0]));            \
  sc[BASE + 1] = fmaf(u8_to_centered_float<1>(W), scale2, __uint_as_float(v[BASE + 1]));            \
  sc[BASE + 2] = fmaf(u8_to_centered_float<2>(W), scale2, __uint_as_float(v[BASE + 2]));            \
  sc[BASE + 3] = fmaf(u8_to_centered_float<3>(W), scale2, __uint_as_float(v[BASE + 3]));
        MMEE_BIAS4(ba.x, 0) MMEE_BIAS4(ba.y, 4) MMEE_BIAS4(ba.z, 8) MMEE_BIAS4(ba.w, 12)
        MMEE_BIAS4(bb.x, 16) MMEE_BIAS4(bb.y, 20) MMEE_BIAS4(bb.z, 24) MMEE_BIAS4(bb.w, 28)
#undef MMEE_BIAS4
        if (flag != 0 || kv0 + ATT_BKV > S) {        // padded keys in this tile and/or keys beyond the document
          const float* ma = args.maskadd + static_cast<size_t>(doc) * args.kv_pitch + kv0 + part * KC;
#pragma unroll
          for (int i = 0; i < KC; i += 4) {
            const float4 m4 = __ldg(reinterpret_cast<const float4*>(ma + i));
            // select, not add: keys beyond the document may carry arbitrary (even non-finite) scores
            if (m4.x < 0.f) sc[i] = -INFINITY;
            if (m4.y < 0.f) sc[i + 1] = -INFINITY;
            if (m4.z < 0.f) sc[i + 2] = -INFINITY;
            if (m4.w < 0.f) sc[i + 3] = -INFINITY;
          }
        }
        float tmax = sc[0];
#pragma unroll
        for (int i = 1; i < KC; ++i) tmax = fmaxf(tmax, sc[i]);
        xch[(b * ATT_NSPLIT + part) * ATT_BQ + r] = tmax;
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&kv_empty[st]);   // bias tile consumed
        asm volatile("bar.sync 1, 512;" ::: "memory");
#pragma unroll
        for (int pp = 0; pp < ATT_NSPLIT; ++pp) tmax = fmaxf(tmax, xch[(b * ATT_NSPLIT + pp) * ATT_BQ + r]);

        const float m_new = fmaxf(m_run, tmax);
        const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
        const float alpha = fast_exp2(m_run - m_use);     // m_run = -inf -> 0

        // O'[t-1] -> registers.  Its o_full wait also proves P V_{t-1} has finished reading the single P buffer.
        if (nproc > 0) accumulate(t - 1, alpha_prev);
        alpha_prev = alpha;

        // ---- p = exp2(s - m), partial row sum, bf16 P -> smem (SW128 K-major A operand)
        float psum = 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float p[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            p[e] = fast_exp2(sc[q * 8 + e] - m_use);
            psum += p[e];
          }
          *reinterpret_cast<uint4*>(sp + ((((part & 1) * 4 + q) ^ sw) << 4)) =
              make_uint4(pack_bf16x2(p[0], p[1]), pack_bf16x2(p[2], p[3]), pack_bf16x2(p[4], p[5]),
                         pack_bf16x2(p[6], p[7]));
        }
        l_run = l_run * alpha + psum;
        m_run = m_new;
        fence_proxy_async_smem();                     // P visible to the tensor core (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[b]);
        ++t;
        ++nproc;
        advance(c);
      }
      accumulate(t - 1, alpha_prev);

      // ---- combine the partial row sums, normalise and store ctx[row, head*64 + 16*part .. +15]
      xl[part * ATT_BQ + r] = l_run;
      asm volatile("bar.sync 1, 512;" ::: "memory");
      float l_tot = 0.f;
#pragma unroll
      for (int pp = 0; pp < ATT_NSPLIT; ++pp) l_tot += xl[pp * ATT_BQ + r];
      asm volatile("bar.sync 1, 512;" ::: "memory");      // xl may be rewritten by the next item
      const int q = q0 + r;
      if (q < S) {
        const float inv = 1.0f / l_tot;
        uint4* dst = reinterpret_cast<uint4*>(args.ctx + static_cast<size_t>(row0 + q) * args.H + head * ATT_D + part * OD);
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

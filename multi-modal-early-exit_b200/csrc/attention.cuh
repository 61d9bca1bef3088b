// Fused attention for LayoutLMv3 (HF modeling_layoutlmv3.py:236-289):
//     ctx = softmax( (Q/8) K^T + (rel_pos + rel_2d_pos)/8 + key_mask ) V
// Persistent kernel, TWO co-resident CTAs per SM (grid = 2 x #SMs), work item = (document slot, head,
// 128-query tile), 64-key tiles.  Everything between Q/K/V and ctx stays on the SM:
//   S_t  = Q K_t^T            tcgen05.mma 128x64x64 (SS), fp32 in TMEM (two S buffers)
//   P_t  = exp2(S_t + bias_t - ref)   one softmax thread per query row (= TMEM lane): tcgen05.ld, bias from smem,
//                                     MUFU.EX2, bf16 pairs written back INTO the S buffer with tcgen05.st
//   O   += P_t [V_t | 1]      tcgen05.mma 128x80x64 with the A operand (P) read from TMEM (TS form); O (64 dims +
//                             the row sum of P from a ones-row appended to V^T) accumulates in TMEM over the item
// so neither the [S,S] scores nor P ever touch shared or global memory, and O is read back once per item.
// The additive bias (1-D + 2-D relative-position buckets; layer-invariant, built once per forward as uint8 with a
// per-head scale, see bias_build_kernel) is streamed tile by tile with TMA and added in registers.  The CogView
// "PB-relax" softmax of HF:224-234 is the standard max-shifted softmax.  Key padding: fully padded key tiles are
// skipped, mixed tiles select -inf per key.
//
// Roles per CTA (192 threads): warp 0 TMA producer, warp 1 MMA issuer (+TMEM owner), warps 2-5 softmax.
// S_{t+1} is issued before P V_t, and P V_t only needs P_t, so the softmax warps never wait on the tensor core in
// steady state; the two CTAs of an SM fill each other's MUFU / FMA bubbles.
//
// Everything runs in the log2 domain (log2(e)/sqrt(d) is folded into W_q).  `ref` is the row reference of the
// online softmax, kept as an integer multiple q_ref of the bias quantum so that it rides for free in the FADD that
// removes the uint8->float magic offset:
//   s - ref = ((2^23 + u) - (2^23 + 128 + q_ref)) * scale2 + acc.
// The first tile of an item takes its exact row maximum as reference.  Later tiles keep it unless a score exceeds
// it by more than 2^8 (lazy rescaling: P <= 256 is harmless in bf16/fp32); then the reference is raised for the
// following tiles and O is rescaled in place in TMEM by exp2(old - new) before the next P V.  A score that jumps
// 2^100 above everything before it raises err_flag.
#pragma once
#include <cuda.h>

#include <type_traits>

#include "ptx.cuh"

namespace mmee {

constexpr int ATT_SM_WARPS = 4;      // softmax warps: one thread per query row
constexpr int ATT_THREADS = 64 + ATT_SM_WARPS * 32;
constexpr int ATT_CTAS_PER_SM = 2;
constexpr int ATT_BQ = 128;    // query rows per CTA
constexpr int ATT_BKV = 64;    // keys per tile
constexpr int ATT_D = 64;
constexpr int ATT_DV = 80;     // V^T rows fed to the PV MMA: 64 dims + a ones row (row sum of P) + 15 zero rows
constexpr int ATT_STAGES = 3;
constexpr int ATT_MAX_KV_TILES = 16;
constexpr float ATT_LAZY = 8.0f;   // raise the row reference only when a score exceeds it by more than 2^8

struct AttSmem {
  static constexpr int Q_BYTES = ATT_BQ * ATT_D * 2;           // 16 KB
  static constexpr int K_BYTES = ATT_BKV * ATT_D * 2;          //  8 KB
  static constexpr int V_BYTES = ATT_DV * ATT_BKV * 2;         // 10 KB: [80 rows x 64 keys] bf16, SW128
  static constexpr int B_BYTES = ATT_BQ * ATT_BKV;             //  8 KB  uint8 [128 x 64], SW64
  static constexpr int KV_STAGE = K_BYTES + V_BYTES + B_BYTES; // 26 KB
  static constexpr int TX_BYTES = K_BYTES + ATT_D * ATT_BKV * 2 + B_BYTES;   // bytes TMA writes per stage
  static constexpr int Q_OFF = 0;                              // 2 Q buffers
  static constexpr int KV_OFF = Q_OFF + 2 * Q_BYTES;
  static constexpr int SC_OFF = KV_OFF + ATT_STAGES * KV_STAGE;   // per-head bias scale table (32 floats)
  static constexpr int BAR_OFF = SC_OFF + 128;
  // q_full[2] q_empty[2] kv_full[ST] kv_empty[ST] s_full[2] p_full[2] o_full[1]
  static constexpr int N_BARS = 2 + 2 + 2 * ATT_STAGES + 2 + 2 + 1;
  static constexpr int TOTAL = BAR_OFF + N_BARS * 8 + 16;
  static constexpr int DYN_BYTES = TOTAL + 1024;
};
static_assert(AttSmem::DYN_BYTES <= 115712, "two attention CTAs must fit one SM");
static_assert(AttSmem::KV_STAGE % 1024 == 0 && AttSmem::K_BYTES % 1024 == 0 && (AttSmem::K_BYTES + AttSmem::V_BYTES) % 1024 == 0,
              "swizzled tiles need 1024 B alignment");

struct AttArgs {
  const int* n_active_dev;
  const uint2* slot_meta;      // slot -> {live_tiles | partial_tiles << 16, doc}   (see slot_meta_kernel)
  const float* maskadd;        // [docs][kv_pitch] 0 / -inf
  const float* bias_scale2;    // [heads] scale_h * log2(e)
  int* err_flag;               // set to 1 if a score ran > 2^100 above its row reference (never in practice)
  __nv_bfloat16* ctx;          // [M, H]
  int H, heads, seq, kv_pitch;
};

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// byte K of w -> 2^23 + byte as a float (one PRMT); the caller's FADD removes the offset (and the row reference).
template <int K>
__device__ __forceinline__ float u8_magic(uint32_t w) {
  return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7540 | K));
}

// Position in this CTA's (item, key tile) sequence; fully masked tiles are skipped.  Every role walks the same
// sequence so the pipeline counters stay in lock-step.  Passed by value so it lives in registers.
struct AttCursor {
  int item, ii, j, slot, head, q0, doc, first_j, last_j;
  uint32_t live, partial;      // bit j: tile j is processed / has padded keys inside the document
  uint2 nmeta;                 // slot_meta of the NEXT item of this CTA, loaded one item ahead (latency hidden)
  bool valid;
};

__device__ __forceinline__ uint2 att_item_meta(int item, int total_items, int n_qt, const AttArgs& args) {
  return (item < total_items) ? __ldg(args.slot_meta + (item / n_qt) / args.heads) : make_uint2(1u, 0u);
}
__device__ __forceinline__ AttCursor att_enter(int item, int ii, uint2 meta, int total_items, int n_qt, int stride,
                                               const AttArgs& args) {
  AttCursor c;
  c.item = item; c.ii = ii; c.valid = item < total_items;
  c.j = 0; c.slot = 0; c.head = 0; c.q0 = 0; c.doc = 0; c.first_j = 0; c.last_j = 0; c.live = 1u; c.partial = 0u;
  c.nmeta = make_uint2(1u, 0u);
  if (!c.valid) return c;
  const int qt = item % n_qt;
  const int sh = item / n_qt;
  c.head = sh % args.heads;
  c.slot = sh / args.heads;
  c.q0 = qt * ATT_BQ;
  c.doc = static_cast<int>(meta.y);
  c.live = meta.x & 0xFFFFu;
  c.partial = meta.x >> 16;
  c.first_j = __ffs(c.live) - 1;
  c.last_j = 31 - __clz(c.live);
  c.j = c.first_j;
  c.nmeta = att_item_meta(item + stride, total_items, n_qt, args);
  return c;
}
__device__ __forceinline__ AttCursor att_first(int total_items, int n_qt, int stride, const AttArgs& args) {
  return att_enter(blockIdx.x, 0, att_item_meta(blockIdx.x, total_items, n_qt, args), total_items, n_qt, stride, args);
}
__device__ __forceinline__ AttCursor att_next(AttCursor c, int total_items, int n_qt, int stride, const AttArgs& args) {
  const uint32_t rest = c.live & ~((2u << c.j) - 1u);
  if (rest) { c.j = __ffs(rest) - 1; return c; }
  return att_enter(c.item + stride, c.ii + 1, c.nmeta, total_items, n_qt, stride, args);
}

// slot_meta[slot] = {live | partial << 16, doc} from the per-document tile flags (keymask_kernel).
__global__ void slot_meta_kernel(const int* __restrict__ slot_doc, const int* __restrict__ tileflag,
                                 const int* __restrict__ n_active_dev, uint2* __restrict__ slot_meta, int n_kv) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= *n_active_dev) return;
  const int doc = slot_doc[s];
  uint32_t live = 0u, partial = 0u;
  for (int j = 0; j < n_kv; ++j) {
    const int f = tileflag[doc * n_kv + j];
    if (f != 2) live |= 1u << j;
    if (f == 1) partial |= 1u << j;
  }
  if (live == 0u) live = 1u;                     // degenerate: keep one tile so the row sum is defined
  slot_meta[s] = make_uint2(live | (partial << 16), static_cast<uint32_t>(doc));
}

// tmap_q   : bf16 [M_max, 2H]                    box [128 rows x 64 cols]   (SW128)
// tmap_k   : bf16 [M_max, 2H]                    box [ 64 rows x 64 cols]   (SW128)
// tmap_vt  : bf16 [docs*heads*64, kv_pitch]       box [ 64 rows x 64 cols]   (SW128)
// tmap_bias: u8   [docs*heads*seq, bias_pitch]    box [128 rows x 64 B]      (SW64)
__global__ void __launch_bounds__(ATT_THREADS, ATT_CTAS_PER_SM)
attention_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                 const __grid_constant__ CUtensorMap tmap_vt, const __grid_constant__ CUtensorMap tmap_bias,
                 const AttArgs args) {
  const int S = args.seq;
  const int n_kv = (S + ATT_BKV - 1) / ATT_BKV;
  const int n_qt = (S + ATT_BQ - 1) / ATT_BQ;
  const int total_items = *args.n_active_dev * args.heads * n_qt;
  const int stride = gridDim.x;

  extern __shared__ uint8_t smem_raw[];
  // 32-bit shared-window address of the 1024 B aligned working area; all hot-loop accesses use these addresses
  const uint32_t sb = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (sb - smem_u32(smem_raw));
  const uint32_t bar0 = sb + AttSmem::BAR_OFF;
  const uint32_t q_full = bar0;                                  // [2]
  const uint32_t q_empty = q_full + 2 * 8;                       // [2]
  const uint32_t kv_full = q_empty + 2 * 8;                      // [STAGES]
  const uint32_t kv_empty = kv_full + ATT_STAGES * 8;            // [STAGES]
  const uint32_t s_full = kv_empty + ATT_STAGES * 8;             // [2]
  const uint32_t p_full = s_full + 2 * 8;                        // [2]
  const uint32_t o_full = p_full + 2 * 8;                        // [1]  every P V commit
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + AttSmem::BAR_OFF);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + AttSmem::N_BARS);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_vt);
    tma_prefetch_desc(&tmap_bias);
    uint64_t* b = bars;
    for (int i = 0; i < 2; ++i) mbar_init(b++, 1);                  // q_full
    for (int i = 0; i < 2; ++i) mbar_init(b++, 1);                  // q_empty
    for (int i = 0; i < ATT_STAGES; ++i) mbar_init(b++, 1);         // kv_full
    for (int i = 0; i < ATT_STAGES; ++i) mbar_init(b++, 1 + ATT_SM_WARPS);   // kv_empty: PV commit + bias consumed
    for (int i = 0; i < 2; ++i) mbar_init(b++, 1);                  // s_full
    for (int i = 0; i < 2; ++i) mbar_init(b++, ATT_SM_WARPS);       // p_full
    mbar_init(b++, 1);                                              // o_full
    fence_mbar_init();
  }
  float* s_scale2 = reinterpret_cast<float*>(smem + AttSmem::SC_OFF);
  if (threadIdx.x < 32) s_scale2[threadIdx.x] = (static_cast<int>(threadIdx.x) < args.heads) ? args.bias_scale2[threadIdx.x] : 1.f;
  // constant rows 64..79 of every V^T tile: row 64 = 1.0 (PV then also yields the row sum of P), rest 0
  for (int i = threadIdx.x; i < ATT_STAGES * 128; i += blockDim.x) {
    const int stg = i >> 7, chunk = i & 127;                   // 128 x 16 B chunks = rows 64..79
    uint8_t* base = smem + AttSmem::KV_OFF + stg * AttSmem::KV_STAGE + AttSmem::K_BYTES + ATT_D * 128;
    const uint32_t val = (chunk < 8) ? 0x3F803F80u : 0u;       // first 8 chunks = row 64
    *reinterpret_cast<uint4*>(base + chunk * 16) = make_uint4(val, val, val, val);
  }
  fence_proxy_async_smem();
  if (warp == 1) tmem_alloc<256>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base;            // 2 x 64 columns; P_t (bf16 pairs) overwrites columns [0,32) of S_t
  const uint32_t tmem_O = tmem_base + 128;      // 80 columns: 64 dims + row sum + padding

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      AttCursor c = att_first(total_items, n_qt, stride, args);
      uint32_t t = 0;
      int loaded_ii = -1;
      while (c.valid) {
        const int row0 = c.slot * S;
        if (c.ii != loaded_ii) {
          loaded_ii = c.ii;
          const int qb = c.ii & 1;
          mbar_wait_suspend(q_empty + qb * 8, ((c.ii >> 1) & 1) ^ 1);
          mbar_expect_tx(q_full + qb * 8, AttSmem::Q_BYTES);
          tma_load_2d(sb + AttSmem::Q_OFF + qb * AttSmem::Q_BYTES, &tmap_q, q_full + qb * 8, c.head * ATT_D, row0 + c.q0);
          // pull the NEXT item's bias tiles (the only operand that comes from DRAM) into L2 ahead of time
          const int nitem = c.item + stride;
          if (nitem < total_items) {
            const int nsh = nitem / n_qt;
            const int nrow = (static_cast<int>(c.nmeta.y) * args.heads + nsh % args.heads) * S + (nitem % n_qt) * ATT_BQ;
            for (int j = 0; j < n_kv; ++j)
              if ((c.nmeta.x >> j) & 1u) tma_prefetch_2d(&tmap_bias, j * ATT_BKV, nrow);
          }
        }
        const int st = t % ATT_STAGES;
        mbar_wait_suspend(kv_empty + st * 8, ((t / ATT_STAGES) & 1) ^ 1);
        const uint32_t sk = sb + AttSmem::KV_OFF + st * AttSmem::KV_STAGE;
        const uint32_t sv = sk + AttSmem::K_BYTES;
        const uint32_t sbias = sv + AttSmem::V_BYTES;
        const int kv0 = c.j * ATT_BKV;
        mbar_expect_tx(kv_full + st * 8, AttSmem::TX_BYTES);
        tma_load_2d(sk, &tmap_k, kv_full + st * 8, args.H + c.head * ATT_D, row0 + kv0);
        tma_load_2d(sv, &tmap_vt, kv_full + st * 8, kv0, (c.slot * args.heads + c.head) * ATT_D);
        tma_load_2d(sbias, &tmap_bias, kv_full + st * 8, kv0, (c.doc * args.heads + c.head) * S + c.q0);
        ++t;
        c = att_next(c, total_items, n_qt, stride, args);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(ATT_BQ, ATT_BKV);
      constexpr uint32_t idesc_o = umma_idesc_bf16(ATT_BQ, ATT_DV);
      AttCursor cs = att_first(total_items, n_qt, stride, args);   // next S = Q K^T to issue (runs one tile ahead)
      AttCursor cp = cs;                                              // next P V to issue
      uint32_t ts = 0;
      auto issue_s = [&]() {
        const int st = ts % ATT_STAGES;
        const int qb = cs.ii & 1;
        if (cs.j == cs.first_j) mbar_wait_suspend(q_full + qb * 8, (cs.ii >> 1) & 1);
        mbar_wait_suspend(kv_full + st * 8, (ts / ATT_STAGES) & 1);
        tc_fence_after();
        const uint64_t dq = umma_desc_sw128_kmajor(sb + AttSmem::Q_OFF + qb * AttSmem::Q_BYTES);
        const uint64_t dk = umma_desc_sw128_kmajor(sb + AttSmem::KV_OFF + st * AttSmem::KV_STAGE);
        // S buffer ts&1 last held P_{ts-2}; its P V was issued before this point and tcgen05.mma executes in order
#pragma unroll
        for (int k = 0; k < ATT_D / 16; ++k)
          umma_bf16_ss(tmem_S + (ts & 1) * ATT_BKV, dq + 2 * k, dk + 2 * k, idesc_s, k ? 1u : 0u);
        umma_commit(s_full + (ts & 1) * 8);
        if (cs.j == cs.last_j) umma_commit(q_empty + qb * 8);   // last use of this item's Q
        ++ts;
        cs = att_next(cs, total_items, n_qt, stride, args);
      };
      if (cs.valid) issue_s();
      for (uint32_t t = 0; t < ts; ++t) {               // ts grows while tiles remain
        if (cs.valid) issue_s();
        const int st = t % ATT_STAGES;
        const int b = t & 1;
        const bool first = (cp.j == cp.first_j);
        // P_t is in TMEM; for the first tile of an item the softmax warps have also read the previous item's O
        mbar_wait_suspend(p_full + b * 8, (t >> 1) & 1);
        tc_fence_after();
        const uint64_t dv = umma_desc_sw128_kmajor(sb + AttSmem::KV_OFF + st * AttSmem::KV_STAGE + AttSmem::K_BYTES);
#pragma unroll
        for (int k = 0; k < ATT_BKV / 16; ++k)
          umma_bf16_ts(tmem_O, tmem_S + b * ATT_BKV + k * 8, dv + 2 * k, idesc_o, (k || !first) ? 1u : 0u);
        umma_commit(o_full);
        umma_commit(kv_empty + st * 8);
        cp = att_next(cp, total_items, n_qt, stride, args);
      }
    }
  } else {
    // ------------------------------------------------------------ softmax (warps 2..5), thread = query row
    const int quarter = warp & 3;                             // TMEM lane quarter this warp may access
    const int r = quarter * 32 + lane;                        // query row within the tile == TMEM lane
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    constexpr float C0 = 8388736.0f;                          // 2^23 + 128
    // bias tile: [128 rows x 64 B], SWIZZLE_64B: 16 B chunk c of row r sits at r*64 + ((c ^ ((r >> 1) & 3)) << 4)
    const uint32_t swb = static_cast<uint32_t>((r >> 1) & 3);
    const uint32_t bias_row = sb + AttSmem::KV_OFF + AttSmem::K_BYTES + AttSmem::V_BYTES + r * 64;
    const uint32_t bo0 = (0u ^ swb) << 4, bo1 = (1u ^ swb) << 4, bo2 = (2u ^ swb) << 4, bo3 = (3u ^ swb) << 4;
    uint32_t t = 0;
    AttCursor c = att_first(total_items, n_qt, stride, args);

    float q_ref = 0.f, alpha_pend = 1.f, scale2 = 1.f, inv_scale2 = 1.f;
    __nv_bfloat16* out_ptr = nullptr;                         // ctx destination of the open item (nullptr: row >= S)
    bool have_item = false;

    // O (TMEM) / row sum -> ctx[row, head*64 .. +63]; the caller has waited for the item's last P V
    auto store_item = [&](__nv_bfloat16* dst_row) {
      uint32_t v[32];
      const uint32_t lsum = tmem_ld1(tmem_O + lane_addr + ATT_D);
      tmem_ld32(tmem_O + lane_addr, v);
      tmem_ld_wait();
      const float inv = 1.0f / __uint_as_float(lsum);
      uint4* dst = reinterpret_cast<uint4*>(dst_row);
      if (dst_row) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          dst[i] = make_uint4(pack_bf16x2(__uint_as_float(v[i * 8 + 0]) * inv, __uint_as_float(v[i * 8 + 1]) * inv),
                              pack_bf16x2(__uint_as_float(v[i * 8 + 2]) * inv, __uint_as_float(v[i * 8 + 3]) * inv),
                              pack_bf16x2(__uint_as_float(v[i * 8 + 4]) * inv, __uint_as_float(v[i * 8 + 5]) * inv),
                              pack_bf16x2(__uint_as_float(v[i * 8 + 6]) * inv, __uint_as_float(v[i * 8 + 7]) * inv));
      }
      tmem_ld32(tmem_O + lane_addr + 32, v);
      tmem_ld_wait();
      if (dst_row) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          dst[4 + i] = make_uint4(pack_bf16x2(__uint_as_float(v[i * 8 + 0]) * inv, __uint_as_float(v[i * 8 + 1]) * inv),
                                  pack_bf16x2(__uint_as_float(v[i * 8 + 2]) * inv, __uint_as_float(v[i * 8 + 3]) * inv),
                                  pack_bf16x2(__uint_as_float(v[i * 8 + 4]) * inv, __uint_as_float(v[i * 8 + 5]) * inv),
                                  pack_bf16x2(__uint_as_float(v[i * 8 + 6]) * inv, __uint_as_float(v[i * 8 + 7]) * inv));
      }
    };

    while (c.valid) {
      const int st = t % ATT_STAGES;
      const int b = t & 1;
      const int kv0 = c.j * ATT_BKV;
      const bool first = (c.j == c.first_j);                 // first processed tile of a new item
      const bool need_mask = (c.partial >> c.j) & 1u;        // padded text keys inside this tile (rare)
      const bool tail = kv0 + ATT_BKV > S;                   // keys beyond the document (last tile)
      const uint32_t brow = bias_row + st * AttSmem::KV_STAGE;
      const uint32_t tS = tmem_S + lane_addr + b * ATT_BKV;
      __nv_bfloat16* prev_out = out_ptr;
      if (first) {
        scale2 = lds_f32(sb + AttSmem::SC_OFF + c.head * 4);
        inv_scale2 = 1.0f / scale2;
        const int q = c.q0 + r;
        out_ptr = (q < S) ? args.ctx + static_cast<size_t>(c.slot * S + q) * args.H + c.head * ATT_D : nullptr;
      }
      mbar_wait(s_full + b * 8, (t >> 1) & 1);      // S_t done  (=> kv_full[st] landed: the MMA waited on it)
      tc_fence_after();

      const float crow = first ? C0 : (C0 + q_ref);  // exact: |q_ref| < 2^22 integers
      float pmax;
      if (!first && !tail && !need_mask) {
        // ---- fast path (all but the first / last / padded tiles): the row reference is already known, so
        // s = S + bias - ref, the running max and p = exp2(s) form ONE straight-line block per thread and the
        // scheduler overlaps the MUFU stream with the PRMT / FADD / FFMA work of the following elements.
        uint32_t v0[32], v1[32], pk[32];
        tmem_ld32(tS, v0);
        tmem_ld32(tS + 32, v1);
        const uint4 ba = lds128(brow + bo0);          // keys  0..15
        const uint4 bb = lds128(brow + bo1);          // keys 16..31
        const uint4 bc = lds128(brow + bo2);          // keys 32..47
        const uint4 bd = lds128(brow + bo3);          // keys 48..63
        tmem_ld_wait();
        __syncwarp();
        if (lane == 0) mbar_arrive(kv_empty + st * 8);   // bias tile consumed (K/V are released by the P V commit)
        float m0 = -INFINITY, m1 = -INFINITY;
#define MMEE_FAST4(W, V, VB, PB, M)                                                                  \
  {                                                                                                  \
    const float s0 = fmaf(u8_magic<0>(W) - crow, scale2, __uint_as_float(V[VB + 0]));                \
    const float s1 = fmaf(u8_magic<1>(W) - crow, scale2, __uint_as_float(V[VB + 1]));                \
    const float s2 = fmaf(u8_magic<2>(W) - crow, scale2, __uint_as_float(V[VB + 2]));                \
    const float s3 = fmaf(u8_magic<3>(W) - crow, scale2, __uint_as_float(V[VB + 3]));                \
    M = fmaxf(M, fmaxf(fmaxf(s0, s1), fmaxf(s2, s3)));                                               \
    pk[PB] = pack_bf16x2(fast_exp2(s0), fast_exp2(s1));                                              \
    pk[PB + 1] = pack_bf16x2(fast_exp2(s2), fast_exp2(s3));                                          \
  }
        MMEE_FAST4(ba.x, v0, 0, 0, m0) MMEE_FAST4(ba.y, v0, 4, 2, m1) MMEE_FAST4(ba.z, v0, 8, 4, m0) MMEE_FAST4(ba.w, v0, 12, 6, m1)
        MMEE_FAST4(bb.x, v0, 16, 8, m0) MMEE_FAST4(bb.y, v0, 20, 10, m1) MMEE_FAST4(bb.z, v0, 24, 12, m0) MMEE_FAST4(bb.w, v0, 28, 14, m1)
        MMEE_FAST4(bc.x, v1, 0, 16, m0) MMEE_FAST4(bc.y, v1, 4, 18, m1) MMEE_FAST4(bc.z, v1, 8, 20, m0) MMEE_FAST4(bc.w, v1, 12, 22, m1)
        MMEE_FAST4(bd.x, v1, 16, 24, m0) MMEE_FAST4(bd.y, v1, 20, 26, m1) MMEE_FAST4(bd.z, v1, 24, 28, m0) MMEE_FAST4(bd.w, v1, 28, 30, m1)
#undef MMEE_FAST4
        pmax = fmaxf(m0, m1);
        tmem_st32(tS, pk);
      } else {
        // ---- general path: s = S + bias - ref, key masks, exact row max (first tile of an item), then p = exp2(s)
        float sc[ATT_BKV];
        {
          uint32_t v[32];
          tmem_ld32(tS, v);
          const uint4 ba = lds128(brow + bo0);
          const uint4 bb = lds128(brow + bo1);
          tmem_ld_wait();
#define MMEE_BIAS4(W, BASE, VB)                                                                      \
  sc[BASE + 0] = fmaf(u8_magic<0>(W) - crow, scale2, __uint_as_float(v[VB + 0]));                    \
  sc[BASE + 1] = fmaf(u8_magic<1>(W) - crow, scale2, __uint_as_float(v[VB + 1]));                    \
  sc[BASE + 2] = fmaf(u8_magic<2>(W) - crow, scale2, __uint_as_float(v[VB + 2]));                    \
  sc[BASE + 3] = fmaf(u8_magic<3>(W) - crow, scale2, __uint_as_float(v[VB + 3]));
          MMEE_BIAS4(ba.x, 0, 0) MMEE_BIAS4(ba.y, 4, 4) MMEE_BIAS4(ba.z, 8, 8) MMEE_BIAS4(ba.w, 12, 12)
          MMEE_BIAS4(bb.x, 16, 16) MMEE_BIAS4(bb.y, 20, 20) MMEE_BIAS4(bb.z, 24, 24) MMEE_BIAS4(bb.w, 28, 28)
          tmem_ld32(tS + 32, v);
          const uint4 bc = lds128(brow + bo2);
          const uint4 bd = lds128(brow + bo3);
          tmem_ld_wait();
          MMEE_BIAS4(bc.x, 32, 0) MMEE_BIAS4(bc.y, 36, 4) MMEE_BIAS4(bc.z, 40, 8) MMEE_BIAS4(bc.w, 44, 12)
          MMEE_BIAS4(bd.x, 48, 16) MMEE_BIAS4(bd.y, 52, 20) MMEE_BIAS4(bd.z, 56, 24) MMEE_BIAS4(bd.w, 60, 28)
#undef MMEE_BIAS4
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(kv_empty + st * 8);
        if (tail) {
#pragma unroll
          for (int i = 0; i < ATT_BKV; ++i)
            if (kv0 + i >= S) sc[i] = -INFINITY;
        }
        if (need_mask) {                              // padded keys in this tile
          const float* ma = args.maskadd + static_cast<size_t>(c.doc) * args.kv_pitch + kv0;
#pragma unroll
          for (int i = 0; i < ATT_BKV; i += 4) {
            const float4 m4 = __ldg(reinterpret_cast<const float4*>(ma + i));
            // select, not add: robust to arbitrary (even non-finite) scores on masked keys
            if (m4.x < 0.f) sc[i] = -INFINITY;
            if (m4.y < 0.f) sc[i + 1] = -INFINITY;
            if (m4.z < 0.f) sc[i + 2] = -INFINITY;
            if (m4.w < 0.f) sc[i + 3] = -INFINITY;
          }
        }
        pmax = fmaxf(sc[0], sc[1]);
#pragma unroll
        for (int i = 2; i < ATT_BKV; i += 2) pmax = fmaxf(pmax, fmaxf(sc[i], sc[i + 1]));
        float shift = 0.f;
        if (first) {                                  // exact row maximum (rounded to the bias quantum) as reference
          q_ref = (pmax == -INFINITY) ? 0.f : rintf(pmax * inv_scale2);
          shift = q_ref * scale2;
        }
        uint32_t pk[32];
#pragma unroll
        for (int i = 0; i < 32; ++i)
          pk[i] = pack_bf16x2(fast_exp2(sc[2 * i] - shift), fast_exp2(sc[2 * i + 1] - shift));
        tmem_st32(tS, pk);
      }

      // ---- P V_{t-1} must be complete before O is read (new item) or rescaled, and before P V_t may be issued
      if (t > 0) {
        mbar_wait(o_full, (t - 1) & 1);
        tc_fence_after();
        if (first) {
          if (have_item) store_item(prev_out);
        } else if (__any_sync(0xffffffffu, alpha_pend != 1.0f)) {    // rare: an earlier tile raised the row reference
          uint32_t o[32];
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            tmem_ld32(tmem_O + lane_addr + hh * 32, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha_pend);
            tmem_st32(tmem_O + lane_addr + hh * 32, o);
          }
          const uint32_t ls = tmem_ld1(tmem_O + lane_addr + ATT_D);
          tmem_ld_wait();
          tmem_st1(tmem_O + lane_addr + ATT_D, __float_as_uint(__uint_as_float(ls) * alpha_pend));
        }
      }
      have_item = true;
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full + b * 8);

      // ---- reference for the following tiles (lazy: only when a score ran more than 2^8 above it)
      alpha_pend = 1.f;
      if (!first && pmax > ATT_LAZY) {
        const float dq = ceilf(pmax * inv_scale2);
        q_ref += dq;
        alpha_pend = fast_exp2(-dq * scale2);
        if (pmax > 100.f) *args.err_flag = 1;
      }
      ++t;
      c = att_next(c, total_items, n_qt, stride, args);
    }
    if (have_item) {
      mbar_wait(o_full, (t - 1) & 1);
      tc_fence_after();
      store_item(out_ptr);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<256>(tmem_base);
  }
}

}  // namespace mmee

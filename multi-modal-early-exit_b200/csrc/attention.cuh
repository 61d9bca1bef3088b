// Fused attention for LayoutLMv3 (HF modeling_layoutlmv3.py:236-289):
//     ctx = softmax( (Q/8) K^T + (rel_pos + rel_2d_pos)/8 + key_mask ) V
// Persistent kernel, TWO co-resident CTAs per SM (grid = 2 x #SMs), work item = (document slot, head,
// 128-query tile), 64-key tiles.  QK^T and PV run on tcgen05; S (fp32, 2 x 64 columns) and the OUTPUT
// accumulator O (fp32, 80 columns: 64 dims + the row sum of P from a ones-row appended to V^T) live in TMEM
// for the whole item, so the [S,S] score matrix never leaves the SM and O is read back once per item.
// The additive bias (1-D + 2-D relative-position buckets; layer-invariant, built once per forward as uint8
// with a per-head scale, see bias_build_kernel) is streamed tile by tile with TMA and added in registers.
// The CogView "PB-relax" softmax of HF:224-234 is the standard max-shifted softmax.  Key padding: fully padded
// key tiles are skipped, mixed tiles select -inf per key.
//
// Roles per CTA (320 threads): warp 0 TMA producer, warp 1 MMA issuer (+TMEM owner), warps 2-9 softmax:
// thread = (query row = TMEM lane, half of the tile's 64 keys / half of the 64 output dims); the two threads of
// a row exchange their partial row max through smem once per tile (64-thread named barrier per lane quarter).
// The two CTAs of an SM run out of phase, so one is usually in its MUFU-heavy exp phase while the other does
// the FMA/ALU-heavy bias phase.
//
// Everything runs in the log2 domain (log2(e)/sqrt(d) is folded into W_q).  Per KV tile t:
//   S[t%2] = Q K_t^T                          (MMA 128x64x64)
//   P_t    = exp2(S + bias - ref)             (bf16 -> smem, SW128 K-major A operand)
//   O     += P_t [V_t | 1]                    (MMA 128x80x64, accumulates in TMEM across the item's tiles)
// `ref` is the running row maximum, kept as an integer multiple q_ref of the bias quantum so that for every
// tile but the first of an item it rides for free in the FADD that removes the uint8->float magic offset:
//   s - ref = ((2^23 + u) - (2^23 + 128 + q_ref)) * scale2 + acc.
// The first tile of an item takes the exact row maximum; later tiles use the maximum of the tiles before them
// and raise it afterwards if needed: then (rarely) O is rescaled in place in TMEM by exp2(old - new) before the
// next P V.  P may exceed 1 meanwhile — harmless in fp32/bf16 unless a score jumps 2^100 above everything before
// it, which raises err_flag.
#pragma once
#include <cuda.h>

#include <type_traits>

#include "ptx.cuh"

namespace mmee {

constexpr int ATT_NSPLIT = 2;        // softmax threads per query row (each owns 32 of the tile's 64 keys)
constexpr int ATT_SM_WARPS = 4 * ATT_NSPLIT;
constexpr int ATT_THREADS = 64 + ATT_SM_WARPS * 32;
constexpr int ATT_CTAS_PER_SM = 2;
constexpr int ATT_BQ = 128;    // query rows per CTA
constexpr int ATT_BKV = 64;    // keys per tile
constexpr int ATT_D = 64;
constexpr int ATT_DV = 80;     // V^T rows fed to the PV MMA: 64 dims + a ones row (row sum of P) + 15 zero rows
constexpr int ATT_STAGES = 2;
constexpr int ATT_MAX_KV_TILES = 16;

struct AttSmem {
  static constexpr int Q_BYTES = ATT_BQ * ATT_D * 2;           // 16 KB
  static constexpr int K_BYTES = ATT_BKV * ATT_D * 2;          //  8 KB
  static constexpr int V_BYTES = ATT_DV * ATT_BKV * 2;         // 10 KB: [80 rows x 64 keys] bf16, SW128
  static constexpr int B_BYTES = ATT_BQ * ATT_BKV;             //  8 KB  uint8 [128 x 64], SW64
  static constexpr int P_BYTES = ATT_BQ * ATT_BKV * 2;         // 16 KB  [128 x 64] bf16, SW128
  static constexpr int KV_STAGE = K_BYTES + V_BYTES + B_BYTES; // 26 KB
  static constexpr int TX_BYTES = K_BYTES + ATT_D * ATT_BKV * 2 + B_BYTES;   // bytes TMA writes per stage
  static constexpr int Q_OFF = 0;                              // 2 Q buffers
  static constexpr int KV_OFF = Q_OFF + 2 * Q_BYTES;
  static constexpr int P_OFF = KV_OFF + ATT_STAGES * KV_STAGE;
  static constexpr int X_OFF = P_OFF + P_BYTES;                // partial row max exchange [2][NSPLIT][128] floats
  static constexpr int SC_OFF = X_OFF + 2 * ATT_NSPLIT * ATT_BQ * 4;   // per-head bias scale table (32 floats)
  static constexpr int BAR_OFF = SC_OFF + 128;
  static constexpr int N_BARS = 2 + 2 + 2 * ATT_STAGES + 2 + 2 + 1 + 1;
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
  long long* trace;            // optional developer trace (clock64 stamps of CTA 0), nullptr = off
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
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* s_scale2 = reinterpret_cast<float*>(smem + AttSmem::SC_OFF);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + AttSmem::BAR_OFF);
  uint64_t* q_full = bars;                         // [2]
  uint64_t* q_empty = q_full + 2;                  // [2]
  uint64_t* kv_full = q_empty + 2;                 // [STAGES]
  uint64_t* kv_empty = kv_full + ATT_STAGES;       // [STAGES]
  uint64_t* s_full = kv_empty + ATT_STAGES;        // [2]
  uint64_t* p_full = s_full + 2;                   // [2]
  uint64_t* o_full = p_full + 2;                   // [1]  every P V commit
  uint64_t* o_empty = o_full + 1;                  // [1]  once per item: O has been read out
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + AttSmem::N_BARS);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_vt);
    tma_prefetch_desc(&tmap_bias);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], ATT_SM_WARPS);
    }
    mbar_init(o_full, 1);
    mbar_init(o_empty, ATT_SM_WARPS);
    for (int i = 0; i < ATT_STAGES; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1 + ATT_SM_WARPS);   // MMA commit after PV + softmax warps done with the bias tile
    }
    fence_mbar_init();
  }
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
  const uint32_t tmem_S = tmem_base;            // 2 x 64 columns
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
          mbar_wait(&q_empty[qb], ((c.ii >> 1) & 1) ^ 1);
          mbar_expect_tx(&q_full[qb], AttSmem::Q_BYTES);
          tma_load_2d(smem + AttSmem::Q_OFF + qb * AttSmem::Q_BYTES, &tmap_q, &q_full[qb], c.head * ATT_D, row0 + c.q0);
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
        mbar_wait(&kv_empty[st], ((t / ATT_STAGES) & 1) ^ 1);
        uint8_t* sk = smem + AttSmem::KV_OFF + st * AttSmem::KV_STAGE;
        uint8_t* sv = sk + AttSmem::K_BYTES;
        uint8_t* sb = sv + AttSmem::V_BYTES;
        const int kv0 = c.j * ATT_BKV;
        mbar_expect_tx(&kv_full[st], AttSmem::TX_BYTES);
        tma_load_2d(sk, &tmap_k, &kv_full[st], args.H + c.head * ATT_D, row0 + kv0);
        tma_load_2d(sv, &tmap_vt, &kv_full[st], kv0, (c.slot * args.heads + c.head) * ATT_D);
        tma_load_2d(sb, &tmap_bias, &kv_full[st], kv0, (c.doc * args.heads + c.head) * S + c.q0);
        ++t;
        c = att_next(c, total_items, n_qt, stride, args);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(ATT_BQ, ATT_BKV);
      constexpr uint32_t idesc_o = umma_idesc_bf16(ATT_BQ, ATT_DV);
      const uint64_t dp = umma_desc_sw128_kmajor(smem_u32(smem + AttSmem::P_OFF));
      AttCursor cs = att_first(total_items, n_qt, stride, args);   // next S = Q K^T to issue (runs one tile ahead)
      AttCursor cp = cs;                                              // next P V to issue
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
        cs = att_next(cs, total_items, n_qt, stride, args);
      };
      if (cs.valid) issue_s();
      for (uint32_t t = 0; t < ts; ++t) {               // ts grows while tiles remain
        if (cs.valid) issue_s();
        const int st = t % ATT_STAGES;
        const int b = t & 1;
        const bool first = (cp.j == cp.first_j);
        mbar_wait(&p_full[b], (t >> 1) & 1);
        if (first) mbar_wait(o_empty, (cp.ii & 1) ^ 1);  // previous item's O has been read out of TMEM
        tc_fence_after();
        const uint64_t dv = umma_desc_sw128_kmajor(smem_u32(smem + AttSmem::KV_OFF + st * AttSmem::KV_STAGE + AttSmem::K_BYTES));
#pragma unroll
        for (int k = 0; k < ATT_BKV / 16; ++k)
          umma_bf16_ss(tmem_O, dp + 2 * k, dv + 2 * k, idesc_o, (k || !first) ? 1u : 0u);
        umma_commit(o_full);
        umma_commit(&kv_empty[st]);
        cp = att_next(cp, total_items, n_qt, stride, args);
      }
    }
  } else {
    // ------------------------------------------------------------ softmax (warps 2..9)
    const int quarter = warp & 3;
    const int part = (warp - 2) >> 2;                         // key columns [32*part, +32), output dims [32*part, +32)
    const int r = quarter * 32 + lane;                        // query row within the tile == TMEM lane
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    float* xch = reinterpret_cast<float*>(smem + AttSmem::X_OFF);     // [buf][part][row]
    constexpr int KC = ATT_BKV / ATT_NSPLIT;                          // 32 keys per thread
    constexpr int OD = ATT_D / ATT_NSPLIT;                            // 32 output dims per thread
    constexpr float C0 = 8388736.0f;                                  // 2^23 + 128
    // bias tile: [128 rows x 64 B], SWIZZLE_64B: 16 B chunk c of row r sits at r*64 + ((c ^ ((r >> 1) & 3)) << 4)
    const uint32_t swb = static_cast<uint32_t>((r >> 1) & 3);
    const uint32_t swp = static_cast<uint32_t>(r & 7);                // P tile: SWIZZLE_128B
    uint8_t* sp = smem + AttSmem::P_OFF + r * 128;
    uint32_t t = 0;
    AttCursor c = att_first(total_items, n_qt, stride, args);

    float q_ref = 0.f, alpha_pend = 1.f, scale2 = 1.f, inv_scale2 = 1.f;
    __nv_bfloat16* out_ptr = nullptr;                         // ctx destination of the open item (nullptr: row >= S)
    bool have_item = false;

    auto finish_item = [&](uint32_t last_t) {                 // O (TMEM) / row sum -> ctx[row, head*64 + 32*part .. +31]
      mbar_wait(o_full, last_t & 1);
      tc_fence_after();
      uint32_t v[OD];
      tmem_ld32(tmem_O + lane_addr + part * OD, v);
      const uint32_t lsum = tmem_ld1(tmem_O + lane_addr + ATT_D);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_empty);
      if (out_ptr) {
        const float inv = 1.0f / __uint_as_float(lsum);
        uint4* dst = reinterpret_cast<uint4*>(out_ptr);
#pragma unroll
        for (int i = 0; i < OD / 8; ++i)
          dst[i] = make_uint4(pack_bf16x2(__uint_as_float(v[i * 8 + 0]) * inv, __uint_as_float(v[i * 8 + 1]) * inv),
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
      const uint8_t* sb = smem + AttSmem::KV_OFF + st * AttSmem::KV_STAGE + AttSmem::K_BYTES + AttSmem::V_BYTES + r * 64;
      const bool tr = args.trace && blockIdx.x == 0 && threadIdx.x == 64 && t < 96;
      if (tr) args.trace[t * 8 + 0] = clock64();
      mbar_wait(&s_full[b], (t >> 1) & 1);         // S_t done  (=> kv_full[st] landed: the MMA waited on it)
      tc_fence_after();
      if (tr) args.trace[t * 8 + 1] = clock64();

      uint32_t v[KC];
      tmem_ld32(tmem_S + lane_addr + b * ATT_BKV + part * KC, v);
      const uint4 ba = *reinterpret_cast<const uint4*>(sb + (((part * 2) ^ swb) << 4));        // keys 32*part .. +15
      const uint4 bb = *reinterpret_cast<const uint4*>(sb + (((part * 2 + 1) ^ swb) << 4));    // keys +16 .. +31
      const float sc2 = first ? s_scale2[c.head] : scale2;
      tmem_ld_wait();
      // ---- s = S + bias - ref   (log2 domain; ref = 0 for the first tile of an item)
      const float crow = first ? C0 : (C0 + q_ref);   // exact: |q_ref| < 2^22 integers
      float sc[KC];
#define MMEE_BIAS4(W, BASE)                                                                  \
  sc[BASE + 0] = fmaf(u8_magic<0>(W) - crow, sc2, __uint_as_float(v[BASE + 0]));             \
  sc[BASE + 1] = fmaf(u8_magic<1>(W) - crow, sc2, __uint_as_float(v[BASE + 1]));             \
  sc[BASE + 2] = fmaf(u8_magic<2>(W) - crow, sc2, __uint_as_float(v[BASE + 2]));             \
  sc[BASE + 3] = fmaf(u8_magic<3>(W) - crow, sc2, __uint_as_float(v[BASE + 3]));
      MMEE_BIAS4(ba.x, 0) MMEE_BIAS4(ba.y, 4) MMEE_BIAS4(ba.z, 8) MMEE_BIAS4(ba.w, 12)
      MMEE_BIAS4(bb.x, 16) MMEE_BIAS4(bb.y, 20) MMEE_BIAS4(bb.z, 24) MMEE_BIAS4(bb.w, 28)
#undef MMEE_BIAS4
      if (tail) {
#pragma unroll
        for (int i = 0; i < KC; ++i)
          if (kv0 + part * KC + i >= S) sc[i] = -INFINITY;
      }
      if (need_mask) {                              // padded keys in this tile
        const float* ma = args.maskadd + static_cast<size_t>(c.doc) * args.kv_pitch + kv0 + part * KC;
#pragma unroll
        for (int i = 0; i < KC; i += 4) {
          const float4 m4 = __ldg(reinterpret_cast<const float4*>(ma + i));
          // select, not add: robust to arbitrary (even non-finite) scores on masked keys
          if (m4.x < 0.f) sc[i] = -INFINITY;
          if (m4.y < 0.f) sc[i + 1] = -INFINITY;
          if (m4.z < 0.f) sc[i + 2] = -INFINITY;
          if (m4.w < 0.f) sc[i + 3] = -INFINITY;
        }
      }
      float pmax = sc[0];
#pragma unroll
      for (int i = 1; i < KC; ++i) pmax = fmaxf(pmax, sc[i]);
      xch[(b * ATT_NSPLIT + part) * ATT_BQ + r] = pmax;
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&kv_empty[st]);   // bias tile consumed
      if (tr) args.trace[t * 8 + 2] = clock64();

      // p = exp2(s - ref) as bf16 into smem (SW128 K-major A operand); the row sum comes out of the PV MMA
      auto emit_p = [&](auto first_tag, float shift) {
        constexpr bool kFirst = decltype(first_tag)::value;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float p[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) p[e] = fast_exp2(kFirst ? (sc[q * 8 + e] - shift) : sc[q * 8 + e]);
          *reinterpret_cast<uint4*>(sp + (((part * 4 + q) ^ swp) << 4)) =
              make_uint4(pack_bf16x2(p[0], p[1]), pack_bf16x2(p[2], p[3]), pack_bf16x2(p[4], p[5]),
                         pack_bf16x2(p[6], p[7]));
        }
      };
      if (first) {
        // first tile of an item: exact row max before exponentiating
        asm volatile("bar.sync %0, 64;" ::"r"(quarter + 1) : "memory");   // the 2 warps that share these 32 rows
        const float rmax = fmaxf(pmax, xch[(b * ATT_NSPLIT + (part ^ 1)) * ATT_BQ + r]);
        // close the previous item (its last o_full also proves the single P buffer is free), normalise, store
        if (have_item) finish_item(t - 1);
        have_item = true;
        scale2 = sc2;
        inv_scale2 = 1.0f / sc2;
        q_ref = (rmax == -INFINITY) ? 0.f : rintf(rmax * inv_scale2);
        alpha_pend = 1.f;
        const int q = c.q0 + r;
        out_ptr = (q < S) ? args.ctx + static_cast<size_t>(c.slot * S + q) * args.H + c.head * ATT_D + part * OD : nullptr;
        if (tr) args.trace[t * 8 + 3] = clock64();
        emit_p(std::true_type{}, q_ref * scale2);
      } else {
        // P V_{t-1} must be complete: it reads the single P buffer, and O may need a rescale before P V_t adds to it
        mbar_wait(o_full, (t - 1) & 1);
        if (__any_sync(0xffffffffu, alpha_pend != 1.0f)) {    // rare: a tile raised the row reference
          tc_fence_after();
          uint32_t o[OD];
          tmem_ld32(tmem_O + lane_addr + part * OD, o);
          uint32_t ls = 0;
          if (part == 0) ls = tmem_ld1(tmem_O + lane_addr + ATT_D);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < OD; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha_pend);
          tmem_st32(tmem_O + lane_addr + part * OD, o);
          if (part == 0) tmem_st1(tmem_O + lane_addr + ATT_D, __float_as_uint(__uint_as_float(ls) * alpha_pend));
          tmem_st_wait();
          tc_fence_before();
        }
        if (tr) args.trace[t * 8 + 3] = clock64();
        emit_p(std::false_type{}, 0.f);
      }
      if (tr) args.trace[t * 8 + 4] = clock64();
      fence_proxy_async_smem();                     // P visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[b]);
      if (tr) args.trace[t * 8 + 5] = clock64();

      // ---- reference for the following tiles
      alpha_pend = 1.f;
      if (!first) {
        asm volatile("bar.sync %0, 64;" ::"r"(quarter + 1) : "memory");
        const float rmax = fmaxf(pmax, xch[(b * ATT_NSPLIT + (part ^ 1)) * ATT_BQ + r]);
        if (rmax > 0.f) {                           // this tile raised the row max: later tiles use the new reference
          const float dq = ceilf(rmax * inv_scale2);
          q_ref += dq;
          alpha_pend = fast_exp2(-dq * scale2);
          if (rmax > 100.f) *args.err_flag = 1;
        }
      }
      if (tr) args.trace[t * 8 + 6] = clock64();
      ++t;
      c = att_next(c, total_items, n_qt, stride, args);
      if (tr) args.trace[(t - 1) * 8 + 7] = clock64();
    }
    if (have_item) finish_item(t - 1);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<256>(tmem_base);
  }
}

}  // namespace mmee

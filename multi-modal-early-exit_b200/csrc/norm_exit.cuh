// LayerNorm (+ survivor compaction), attention-bias build, fused exit head / criterion / threshold /
// prefix-sum compaction kernels.
//
//   ln_rows_kernel   : X[dst] = LN(Y[src]) ; dst slot s' <- src slot slot_src[s'] (compaction is free: the
//                      post-LN write goes straight to the survivor's new slot).  HF:300-304, 509-513.
//   bias_build_kernel: (rel_pos + rel_2d_pos)/sqrt(d) + key mask, once per forward, reused by every layer
//                      (EE/models/LayoutLMv3.py:170-179; HF:393-458; mask HF:270-272).
//   exit_head_kernel : CLS row -> [LN] -> dense/tanh/out_proj (EE/models/LayoutLMv3.py:86-93, 226-227; gate mode
//                      also classifier(CLS) :768) -> logits/T -> max-softmax | entropy (EE_modules.py:149-160)
//                      -> strict threshold test (EE_modules.py:139-143, policy.py:33).
//   compact_kernel   : block prefix-sum over the fire flags -> survivor list for the next layer, results of
//                      leaving documents scattered to their original document index; no host round-trip.
#pragma once
#include "embed.cuh"
#include "ptx.cuh"

namespace mmee {

// one warp per destination row; rows >= *m_dst_dev are skipped.
template <int NV>
__global__ void ln_rows_kernel(const float* __restrict__ Y, __nv_bfloat16* __restrict__ X,
                               const float* __restrict__ w, const float* __restrict__ b, float eps, int H, int seq,
                               const int* __restrict__ m_dst_dev, const int* __restrict__ slot_src) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= *m_dst_dev) return;
  const int lane = threadIdx.x & 31;
  size_t src_row = row;
  if (slot_src) {
    const int s = row / seq;
    src_row = static_cast<size_t>(slot_src[s]) * seq + (row - s * seq);
  }
  const float* y = Y + src_row * H;
  float v[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane + 32 * i;
    v[i] = (c < H) ? y[c] : 0.f;
  }
  warp_layernorm<NV>(v, H, w, b, eps, lane);
  __nv_bfloat16* out = X + static_cast<size_t>(row) * H;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane + 32 * i;
    if (c < H) out[c] = __float2bfloat16_rn(v[i]);
  }
}

// ------------------------------------------------------------------ attention bias
// The additive attention bias (rel_pos + rel_2d_pos)/sqrt(d) is layer-invariant (the reference builds it once per
// forward, EE/models/LayoutLMv3.py:170-179; HF:393-458).  It is materialised once per forward as fp16 in the log2
// domain, bias16[doc][head][i][j] = (W1d[b1] + (Wx[bx] + Wy[by])) * log2(e)/sqrt(d), with -60000 on padded keys
// (attention_mask == 0, HF:270-272) and on the pitch padding j >= seq, so the attention kernel can ADD it on the
// tensor core (S += bias16 x I, see attention.cuh) and needs no separate key mask.  12.3 MB per base document
// instead of 2 x 24 MB fp32 in the reference.
//
// Persistent kernel, one CTA per SM: the 2-D table T2[bx][by][head] = fp16((Wx + Wy) * c) (98 KB for base, built at
// weight load) and T1[b1][head] = W1d * c (fp32) live in shared memory for the whole launch, so each (i, j) pair
// costs one bucket computation for all heads and one T1 + one T2 lookup per head PAIR.
struct BiasArgs {
  const int64_t* bbox;       // [B, n_text, 4]
  const int* vis_bbox;       // [n_vis, 4]
  const float* t1;           // [bins1][heads]            W1d * log2(e)/sqrt(d)
  const __half* t2;          // [bins2*bins2][heads]      (Wx + Wy) * log2(e)/sqrt(d)
  const uint8_t* lut1;       // |rel| -> bucket offset, 1-D   (size lut1_n)
  const uint8_t* lut2;       // 2-D
  const float* maskadd;      // [B][kv_pitch] 0 / -inf per key (padding, j >= seq)
  int lut1_n, lut2_n;
  int bins1, bins2;          // rel_pos_bins, rel_2d_pos_bins
  int heads, n_text, seq, pitch, kv_pitch, B;
  __half* out;               // [B][heads][seq][pitch]
};

constexpr int BIAS_THREADS = 768;
constexpr float BIAS_MASKED = -60000.0f;   // finite (0 * x stays 0 in the identity MMA) and exp2() of it is 0

inline size_t bias_build_smem(const BiasArgs& a) {
  return static_cast<size_t>(a.bins2) * a.bins2 * a.heads * 2 + static_cast<size_t>(a.bins1) * a.heads * 4 + a.lut1_n +
         a.lut2_n + static_cast<size_t>(a.pitch) * 16 + 64;
}

// thread = 8 consecutive keys j of one query row i; a pass covers blockDim / (pitch/8) rows of one document.
__global__ void __launch_bounds__(BIAS_THREADS, 1) bias_build_kernel(BiasArgs a) {
  extern __shared__ __align__(16) uint8_t bsm[];
  const int n_t2 = a.bins2 * a.bins2 * a.heads;
  __half* s_t2 = reinterpret_cast<__half*>(bsm);
  float* s_t1 = reinterpret_cast<float*>(bsm + static_cast<size_t>(n_t2) * 2);
  int* s_pos = reinterpret_cast<int*>(s_t1 + a.bins1 * a.heads);      // [pitch] each
  int* s_x = s_pos + a.pitch;
  int* s_y = s_x + a.pitch;
  int* s_m = s_y + a.pitch;                                           // 1 = masked key
  uint8_t* s_l1 = reinterpret_cast<uint8_t*>(s_m + a.pitch);
  uint8_t* s_l2 = s_l1 + a.lut1_n;
  for (int i = threadIdx.x; i < n_t2 / 8; i += blockDim.x)
    reinterpret_cast<uint4*>(s_t2)[i] = __ldg(reinterpret_cast<const uint4*>(a.t2) + i);
  for (int i = threadIdx.x; i < a.bins1 * a.heads; i += blockDim.x) s_t1[i] = a.t1[i];
  for (int i = threadIdx.x; i < a.lut1_n; i += blockDim.x) s_l1[i] = a.lut1[i];
  for (int i = threadIdx.x; i < a.lut2_n; i += blockDim.x) s_l2[i] = a.lut2[i];

  const int chunks = a.pitch >> 3;                      // threads per row
  const int rows_pp = blockDim.x / chunks;              // rows per pass
  const int passes = (a.seq + rows_pp - 1) / rows_pp;   // per document
  const int units = a.B * passes;
  const int u_lo = static_cast<int>(static_cast<long long>(units) * blockIdx.x / gridDim.x);
  const int u_hi = static_cast<int>(static_cast<long long>(units) * (blockIdx.x + 1) / gridDim.x);
  const int rl = threadIdx.x / chunks;                  // row within the pass
  const int j0 = (threadIdx.x - rl * chunks) * 8;
  const bool active = rl < rows_pp;
  const int half1 = a.bins1 >> 1, half2 = a.bins2 >> 1;
  int cur_doc = -1;
  for (int u = u_lo; u < u_hi; ++u) {
    const int doc = u / passes;
    const int i = (u - doc * passes) * rows_pp + rl;
    if (doc != cur_doc) {                               // (re)load this document's key coordinates
      __syncthreads();
      for (int t = threadIdx.x; t < a.pitch; t += blockDim.x) {
        int pos = 0, x0 = 0, y1 = 0, m = 1;
        if (t < a.seq) {
          if (t < a.n_text) {
            const int64_t* bb = a.bbox + (static_cast<size_t>(doc) * a.n_text + t) * 4;
            pos = t; x0 = static_cast<int>(bb[0]); y1 = static_cast<int>(bb[3]);
          } else {
            const int p = t - a.n_text;
            pos = p; x0 = a.vis_bbox[p * 4 + 0]; y1 = a.vis_bbox[p * 4 + 3];
          }
          m = a.maskadd[static_cast<size_t>(doc) * a.kv_pitch + t] < 0.f ? 1 : 0;
        }
        s_pos[t] = pos; s_x[t] = x0; s_y[t] = y1; s_m[t] = m;
      }
      cur_doc = doc;
      __syncthreads();
    }
    if (!active || i >= a.seq) continue;
    const int pi = s_pos[i], xi = s_x[i], yi = s_y[i];
    uint32_t i1[8], i2[8];
    uint32_t mbits = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int j = j0 + k;
      const int r1 = s_pos[j] - pi, rx = s_x[j] - xi, ry = s_y[j] - yi;
      const int b1 = (r1 > 0 ? half1 : 0) + s_l1[min(abs(r1), a.lut1_n - 1)];
      const int bx = (rx > 0 ? half2 : 0) + s_l2[min(abs(rx), a.lut2_n - 1)];
      const int by = (ry > 0 ? half2 : 0) + s_l2[min(abs(ry), a.lut2_n - 1)];
      i1[k] = static_cast<uint32_t>(b1 * a.heads);
      i2[k] = static_cast<uint32_t>((bx * a.bins2 + by) * a.heads);
      mbits |= static_cast<uint32_t>(s_m[j]) << k;
    }
    __half* out = a.out + ((static_cast<size_t>(doc) * a.heads) * a.seq + i) * a.pitch + j0;
    const size_t head_stride = static_cast<size_t>(a.seq) * a.pitch;
    for (int h = 0; h < a.heads; h += 2) {
      float v0[8], v1[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float2 t1 = *reinterpret_cast<const float2*>(s_t1 + i1[k] + h);
        const float2 t2 = __half22float2(*reinterpret_cast<const __half2*>(s_t2 + i2[k] + h));
        const bool m = (mbits >> k) & 1u;
        v0[k] = m ? BIAS_MASKED : t1.x + t2.x;
        v1[k] = m ? BIAS_MASKED : t1.y + t2.y;
      }
      auto pack = [](float lo, float hi) {
        __half2 t = __floats2half2_rn(lo, hi);
        return *reinterpret_cast<uint32_t*>(&t);
      };
      *reinterpret_cast<uint4*>(out + h * head_stride) =
          make_uint4(pack(v0[0], v0[1]), pack(v0[2], v0[3]), pack(v0[4], v0[5]), pack(v0[6], v0[7]));
      *reinterpret_cast<uint4*>(out + (h + 1) * head_stride) =
          make_uint4(pack(v1[0], v1[1]), pack(v1[2], v1[3]), pack(v1[4], v1[5]), pack(v1[6], v1[7]));
    }
  }
}

// Key-padding mask (HF:270-272 adds (1-mask)*finfo.min to the scores): maskadd[doc][j] = 0 / -inf and a per
// (doc, key tile of `tile_keys` keys) flag: 0 = no masked key, 1 = some, 2 = every valid key masked (the tile is skipped).
// grid B, block = kv_pitch threads (<= 1024)
__global__ void keymask_kernel(const int64_t* __restrict__ mask, float* __restrict__ maskadd, int* __restrict__ tileflag,
                               int n_text, int seq, int kv_pitch, int n_tiles, int tile_keys) {
  __shared__ int s_masked[32], s_valid[32];
  const int doc = blockIdx.x;
  const int j = threadIdx.x;
  if (j < 32) { s_masked[j] = 0; s_valid[j] = 0; }
  __syncthreads();
  if (j < kv_pitch) {
    bool m = (j >= seq);
    if (!m && j < n_text) m = (mask[static_cast<size_t>(doc) * n_text + j] == 0);
    maskadd[static_cast<size_t>(doc) * kv_pitch + j] = m ? -INFINITY : 0.f;
    if (j < seq) {
      atomicAdd(&s_valid[j / tile_keys], 1);
      if (m) atomicAdd(&s_masked[j / tile_keys], 1);
    }
  }
  __syncthreads();
  if (j < n_tiles) tileflag[doc * n_tiles + j] = (s_masked[j] == 0) ? 0 : (s_masked[j] == s_valid[j] ? 2 : 1);
}

// ------------------------------------------------------------------ exit head
struct HeadWeights {
  const float* dense_w;   // [H, H] or nullptr (1-layer head)
  const float* dense_b;   // [H]
  const float* out_w;     // [n_out, H]
  const float* out_b;     // [n_out]
  int n_out;
};

// (1) gather the exit input rows of the active slots (+ LayerNorm): Z[slot][H] fp32.  One warp per slot.
struct ExitRowsArgs {
  const float* rows;        // fp32 rows; row of slot s starts at rows + src(s) * row_stride
  size_t row_stride;        // in floats
  const int* slot_src;      // optional: rows not yet compacted -> src(s) = slot_src[s]
  const float* ln_w;        // nullptr -> no LayerNorm (mean-pooled embedding exit)
  const float* ln_b;
  float ln_eps;
  int H;
  const int* n_active_dev;
  float* Z;                 // [n, H]
};

__global__ void exit_rows_kernel(ExitRowsArgs a) {
  const int slot = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (slot >= *a.n_active_dev) return;
  const int lane = threadIdx.x & 31;
  const int H = a.H;
  const int src = a.slot_src ? a.slot_src[slot] : slot;
  const float* r = a.rows + static_cast<size_t>(src) * a.row_stride;
  float* z = a.Z + static_cast<size_t>(slot) * H;
  if (!a.ln_w) {
    for (int c = lane; c < H; c += 32) z[c] = r[c];
    return;
  }
  float s = 0.f;
  for (int c = lane; c < H; c += 32) s += r[c];
  const float mean = warp_sum(s) / H;
  float q = 0.f;
  for (int c = lane; c < H; c += 32) { const float d = r[c] - mean; q += d * d; }
  const float rstd = rsqrtf(warp_sum(q) / H + a.ln_eps);
  for (int c = lane; c < H; c += 32) z[c] = (r[c] - mean) * rstd * __ldg(a.ln_w + c) + __ldg(a.ln_b + c);
}

// (2) T[w][slot][j] = tanh(sum_k W_w[j,k] Z[slot][k] + b_w[j]) for up to two heads w (exit head, classifier).
// fp32 SIMT tile: 32 slots x 64 features per CTA, K staged through shared memory in chunks of 32.
struct ExitDenseArgs {
  const float* Z;           // [n, H]
  const float* w[2];        // [H, H] each
  const float* b[2];
  float* T[2];              // [n, H] each
  int H;
  const int* n_active_dev;
};

constexpr int EXD_DOCS = 32, EXD_FEATS = 64, EXD_K = 32;

__global__ void __launch_bounds__(256) exit_dense_kernel(ExitDenseArgs a) {
  __shared__ float sZ[EXD_DOCS][EXD_K + 1];
  __shared__ float sW[EXD_FEATS][EXD_K + 1];
  const int n = *a.n_active_dev;
  const int d0 = blockIdx.y * EXD_DOCS;
  if (d0 >= n) return;
  const int f0 = blockIdx.x * EXD_FEATS;
  const int which = blockIdx.z;
  const float* __restrict__ W = a.w[which];
  const int H = a.H;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;       // 4 features x 2 slots per thread
  float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
  for (int k0 = 0; k0 < H; k0 += EXD_K) {
    for (int i = threadIdx.x; i < EXD_DOCS * EXD_K; i += 256) {
      const int d = i >> 5, k = i & 31;
      sZ[d][k] = (d0 + d < n) ? a.Z[static_cast<size_t>(d0 + d) * H + k0 + k] : 0.f;
    }
    for (int i = threadIdx.x; i < EXD_FEATS * EXD_K; i += 256) {
      const int f = i >> 5, k = i & 31;
      sW[f][k] = __ldg(W + static_cast<size_t>(f0 + f) * H + k0 + k);
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < EXD_K; ++k) {
      const float z0 = sZ[ty * 2][k], z1 = sZ[ty * 2 + 1][k];
#pragma unroll
      for (int f = 0; f < 4; ++f) {
        const float w = sW[tx * 4 + f][k];
        acc[0][f] = fmaf(w, z0, acc[0][f]);
        acc[1][f] = fmaf(w, z1, acc[1][f]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int dd = 0; dd < 2; ++dd) {
    const int d = d0 + ty * 2 + dd;
    if (d >= n) continue;
#pragma unroll
    for (int f = 0; f < 4; ++f) {
      const int j = f0 + tx * 4 + f;
      a.T[which][static_cast<size_t>(d) * H + j] = tanhf(acc[dd][f] + __ldg(a.b[which] + j));
    }
  }
}

// (3) out_proj + temperature + criterion + strict threshold test.  One warp per slot.
struct ExitOutArgs {
  const float* in_head;     // [n, H] input of the exit head's out_proj (tanh(dense) or Z for 1-layer heads)
  const float* in_cls;      // [n, H] input of the classifier's out_proj (gate mode), else unused
  HeadWeights head;         // ramp: class logits; gate: 2-way gate logits (skipped when head.out_w == nullptr)
  HeadWeights cls;          // gate mode: the final classifier ("gated logits", EE/models/LayoutLMv3.py:768)
  int gate_mode;
  int H, n_labels;
  int criterion;            // 0 max_confidence (fire if >), 1 entropy (fire if <)
  float inv_temp;           // 1/T_e
  float threshold;
  int force;                // final classifier: always fires
  const int* n_active_dev;
  float* slot_logits;       // [n, n_labels]  class logits of this exit
  float* slot_head;         // [n, head.n_out] raw head output (gate logits in gate mode)
  float* slot_crit;         // [n]
  int* slot_fire;           // [n]
};

__device__ __forceinline__ float warp_dot(const float* __restrict__ w, const float* __restrict__ x, int H, int lane) {
  float acc = 0.f;
  for (int k = lane * 4; k < H; k += 128) {
    const float4 w4 = __ldg(reinterpret_cast<const float4*>(w + k));
    const float4 x4 = *reinterpret_cast<const float4*>(x + k);
    acc = fmaf(w4.x, x4.x, acc);
    acc = fmaf(w4.y, x4.y, acc);
    acc = fmaf(w4.z, x4.z, acc);
    acc = fmaf(w4.w, x4.w, acc);
  }
  return warp_sum(acc);
}

__global__ void exit_out_kernel(ExitOutArgs a) {
  const int slot = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (slot >= *a.n_active_dev) return;
  const int lane = threadIdx.x & 31;
  const int H = a.H, K = a.n_labels;
  float head_val = 0.f;                 // lane j holds head logit j
  if (a.head.out_w) {
    const float* x = a.in_head + static_cast<size_t>(slot) * H;
    for (int j = 0; j < a.head.n_out; ++j) {
      const float v = warp_dot(a.head.out_w + static_cast<size_t>(j) * H, x, H, lane) + __ldg(a.head.out_b + j);
      if (lane == j) head_val = v;
    }
    if (lane < a.head.n_out) a.slot_head[static_cast<size_t>(slot) * a.head.n_out + lane] = head_val;
  }
  float raw = head_val;                 // class logit of lane k
  if (a.gate_mode) {
    raw = 0.f;
    const float* x = a.in_cls + static_cast<size_t>(slot) * H;
    for (int j = 0; j < K; ++j) {
      const float v = warp_dot(a.cls.out_w + static_cast<size_t>(j) * H, x, H, lane) + __ldg(a.cls.out_b + j);
      if (lane == j) raw = v;
    }
  }
  if (lane < K) a.slot_logits[static_cast<size_t>(slot) * K + lane] = raw;
  const float z = raw * a.inv_temp;
  const float zmax = warp_max(lane < K ? z : -INFINITY);
  const float e = (lane < K) ? expf(z - zmax) : 0.f;
  const float A = warp_sum(e);
  float crit;
  if (a.criterion == 0) {
    crit = 1.0f / A;                    // max softmax = exp(0) / sum_k exp(z_k - z_max)
  } else {
    // entropy of EE/models/EE_modules.py:149-154, log(sum e^z) - sum z e^z / sum e^z, evaluated max-shifted
    // (identical in exact arithmetic, finite for any temperature).
    const float Bz = warp_sum((lane < K) ? (z - zmax) * e : 0.f);
    crit = logf(A) - Bz / A;
  }
  if (lane == 0) {
    a.slot_crit[slot] = crit;
    const bool fire = a.force || (a.criterion == 0 ? (crit > a.threshold) : (crit < a.threshold));
    a.slot_fire[slot] = fire ? 1 : 0;
  }
}

// ------------------------------------------------------------------ compaction
struct CompactArgs {
  const int* n_active_dev;     // current number of active slots
  int* n_next_dev;             // out: survivors
  int* m_next_dev;             // out: survivors * seq  (row count for the next layer's GEMMs)
  int seq;
  const int* slot_doc;         // [n] slot -> original document index
  int* next_slot_doc;          // [n_next]
  int* next_slot_src;          // [n_next] new slot -> old slot (row gather map for the LN that follows)
  const int* slot_fire;
  const float* slot_logits;    // [n, K]
  const float* slot_head;      // [n, n_head]
  const float* slot_crit;
  int K, n_head;
  int exit_index;              // e
  int leave;                   // 1: firing documents leave (early-exit mode); 0: dense mode, everyone stays
  // per-document outputs
  float* out_logits;           // [B, K]
  float* out_crit;             // [B]
  int* out_exit;               // [B]  (-1 = undecided)
  float* all_logits;           // [(E+1), B, K] or nullptr
  float* all_head;             // [(E+1), B, n_head_max] or nullptr
  float* all_crit;             // [(E+1), B] or nullptr
  int B, n_head_max;
  unsigned long long* hist;    // [E+1]
};

// single CTA of 1024 threads; handles any n via a chunked scan.
__global__ void __launch_bounds__(1024) compact_kernel(CompactArgs a) {
  __shared__ int s_warp[32];
  __shared__ int s_total;
  __shared__ int s_base;
  __shared__ int s_fired;
  const int n = *a.n_active_dev;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { s_base = 0; s_fired = 0; }
  __syncthreads();
  for (int start = 0; start < n; start += 1024) {
    const int s = start + threadIdx.x;
    const bool valid = s < n;
    int doc = -1, fire = 0, first = 0;
    if (valid) {
      doc = a.slot_doc[s];
      fire = a.slot_fire[s];
      // record per-exit outputs for every active document
      if (a.all_logits)
        for (int k = 0; k < a.K; ++k)
          a.all_logits[(static_cast<size_t>(a.exit_index) * a.B + doc) * a.K + k] = a.slot_logits[static_cast<size_t>(s) * a.K + k];
      if (a.all_head)
        for (int k = 0; k < a.n_head; ++k)
          a.all_head[(static_cast<size_t>(a.exit_index) * a.B + doc) * a.n_head_max + k] = a.slot_head[static_cast<size_t>(s) * a.n_head + k];
      if (a.all_crit) a.all_crit[static_cast<size_t>(a.exit_index) * a.B + doc] = a.slot_crit[s];
      first = fire && (a.out_exit[doc] < 0);
      if (first) {
        a.out_exit[doc] = a.exit_index;
        a.out_crit[doc] = a.slot_crit[s];
        for (int k = 0; k < a.K; ++k) a.out_logits[static_cast<size_t>(doc) * a.K + k] = a.slot_logits[static_cast<size_t>(s) * a.K + k];
      }
    }
    const int stay = valid && !(a.leave && fire);
    // block-wide exclusive scan of `stay`
    const unsigned ball = __ballot_sync(0xffffffffu, stay);
    const int wprefix = __popc(ball & ((1u << lane) - 1));
    const unsigned fball = __ballot_sync(0xffffffffu, first);
    if (lane == 0) { s_warp[warp] = __popc(ball); if (fball) atomicAdd(&s_fired, __popc(fball)); }
    __syncthreads();
    if (warp == 0) {
      const int v = s_warp[lane];
      int incl = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
      s_warp[lane] = incl - v;                 // exclusive per-warp offsets
      if (lane == 31) s_total = incl;
    }
    __syncthreads();
    if (stay) {
      const int dst = s_base + s_warp[warp] + wprefix;
      a.next_slot_doc[dst] = doc;
      a.next_slot_src[dst] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) s_base += s_total;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    *a.n_next_dev = s_base;
    *a.m_next_dev = s_base * a.seq;
    if (s_fired) atomicAdd(a.hist + a.exit_index, static_cast<unsigned long long>(s_fired));
  }
}

}  // namespace mmee

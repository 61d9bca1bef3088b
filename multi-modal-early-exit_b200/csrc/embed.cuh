// Embedding-stage kernels (HBM-bound gathers + LayerNorm):
//   posid_kernel        : RoBERTa position ids (HF modeling_layoutlmv3.py:139-147)
//   text_embed_kernel   : 7-way gather-sum + embeddings.LayerNorm + model LayerNorm  (HF:161-200, 113-137;
//                         reference call EE/models/LayoutLMv3.py:511-517, 565)
//   im2col_kernel       : 16x16/s16 patches -> bf16 rows for the patch GEMM (HF:70-82)
//   visual_ln_kernel    : cls|patch + pos_embed -> norm (eps 1e-6) -> model LayerNorm (EE/models/LayoutLMv3.py:358-373, 565)
//   embed_finish_vec_kernel : model LayerNorm of the survivors' text / visual embeddings after an embedding-level exit
//                         (vision_avg / text_avg, EE/models/LayoutLMv3.py:465-483, 519-534): documents that left there
//                         never get text embeddings, a fused row or an attention bias
//   meanpool_kernel     : mean over the 709 fused tokens for the text_visual_concat exit (EE/models/LayoutLMv3.py:582)
// Each token is LayerNormed twice, exactly as the reference does.
// The text kernels check input_ids / bbox against the table sizes (the reference raises IndexError on such input):
// out-of-range values are clamped and reported through an error flag that the synchronous entry points turn into an
// error (mmee_last_error).
#pragma once
#include "ptx.cuh"

namespace mmee {

struct EmbedWeights {
  const float* word;      // [vocab, H]
  const float* type0;     // [H]  (token_type row 0)
  const float* pos;       // [max_pos, H]
  const float* x_emb;     // [1024, coord]
  const float* y_emb;     // [1024, coord]
  const float* h_emb;     // [1024, shape]
  const float* w_emb;     // [1024, shape]
  const float* ln_emb_w;  const float* ln_emb_b;    // embeddings.LayerNorm
  const float* ln_model_w; const float* ln_model_b; // layoutlmv3.LayerNorm
  const float* ln_vis_w;  const float* ln_vis_b;    // layoutlmv3.norm
  const float* cls_token; // [H]
  const float* pos_embed; // [n_vis, H]
};

// one warp per document: pos = cumsum(id != pad) * (id != pad) + pad
__global__ void posid_kernel(const int64_t* __restrict__ ids, int* __restrict__ pos_out, int n_docs, int n_text,
                             int pad_id) {
  const int doc = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (doc >= n_docs) return;
  const int lane = threadIdx.x & 31;
  const int64_t* row = ids + static_cast<size_t>(doc) * n_text;
  int carry = 0;
  for (int base = 0; base < n_text; base += 32) {
    const int t = base + lane;
    const int real = (t < n_text) && (row[t] != pad_id);
    int incl = real;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (t < n_text) pos_out[static_cast<size_t>(doc) * n_text + t] = real ? (carry + incl + pad_id) : pad_id;
    carry += __shfl_sync(0xffffffffu, incl, 31);
  }
}

// LayerNorm of a row held as `NV` values per lane (column = lane + 32*i); fp32 two-pass.
template <int NV>
__device__ __forceinline__ void warp_layernorm(float (&v)[NV], int H, const float* __restrict__ w,
                                               const float* __restrict__ b, float eps, int lane) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) if (lane + 32 * i < H) s += v[i];
  const float mean = warp_sum(s) / H;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) if (lane + 32 * i < H) { const float d = v[i] - mean; q += d * d; }
  const float rstd = rsqrtf(warp_sum(q) / H + eps);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane + 32 * i;
    if (c < H) v[i] = (v[i] - mean) * rstd * __ldg(w + c) + __ldg(b + c);
  }
}

// LayerNorm of a row held as NV4 float4 per lane (columns 4*(lane + 32*i) .. +3); fp32 two-pass.
template <int NV4>
__device__ __forceinline__ void warp_layernorm4(float4 (&v)[NV4], int H, const float* __restrict__ w,
                                                const float* __restrict__ b, float eps, int lane) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV4; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  const float mean = warp_sum(s) / H;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV4; ++i) {
    const float dx = v[i].x - mean, dy = v[i].y - mean, dz = v[i].z - mean, dw = v[i].w - mean;
    q += (dx * dx + dy * dy) + (dz * dz + dw * dw);
  }
  const float rstd = rsqrtf(warp_sum(q) / H + eps);
#pragma unroll
  for (int i = 0; i < NV4; ++i) {
    const int c = 4 * (lane + 32 * i);
    const float4 w4 = __ldg(reinterpret_cast<const float4*>(w + c));
    const float4 b4 = __ldg(reinterpret_cast<const float4*>(b + c));
    v[i].x = (v[i].x - mean) * rstd * w4.x + b4.x;
    v[i].y = (v[i].y - mean) * rstd * w4.y + b4.y;
    v[i].z = (v[i].z - mean) * rstd * w4.z + b4.z;
    v[i].w = (v[i].w - mean) * rstd * w4.w + b4.w;
  }
}

// four consecutive values -> bf16 (8 B store) and, when lo != nullptr, the low parts bf16(v - bf16(v)) of the split
// representation (hi + lo carries 16 mantissa bits; the fp32 engine mode feeds both to the tensor cores)
__device__ __forceinline__ void store_split4(__nv_bfloat16* hi, __nv_bfloat16* lo, const float4& v) {
  const uint32_t h01 = pack_bf16x2(v.x, v.y), h23 = pack_bf16x2(v.z, v.w);
  *reinterpret_cast<uint2*>(hi) = make_uint2(h01, h23);
  if (lo) {
    const float2 f01 = unpack_bf16x2(h01), f23 = unpack_bf16x2(h23);
    *reinterpret_cast<uint2*>(lo) = make_uint2(pack_bf16x2(v.x - f01.x, v.y - f01.y), pack_bf16x2(v.z - f23.x, v.w - f23.y));
  }
}
// Outputs of a residual-stream row: `X` (bf16, GEMM operand), optional `Xlo` (its low part) and optional `X32` (the
// unrounded fp32 row: the fp32 engine mode adds the residual from it, so the residual path is exact as in the reference)
struct RowOut {
  __nv_bfloat16* X;
  __nv_bfloat16* Xlo;
  float* X32;
};
__device__ __forceinline__ void store_row4(const RowOut& o, size_t off, const float4& v) {
  store_split4(o.X + off, o.Xlo ? o.Xlo + off : nullptr, v);
  if (o.X32) *reinterpret_cast<float4*>(o.X32 + off) = v;
}

// one warp per text token; X row = slot*seq + t.  Vectorised variant: H = 128*NV4 and the six spatial segments
// (4 x coord + 2 x shape) are multiples of 4 columns, so every lane gathers whole float4s (base: 6 per table).
struct TextEmbedArgs {
  const int64_t* ids;      // [B, n_text]
  const int64_t* bbox;     // [B, n_text, 4]
  const int* posid;        // [B, n_text]
  __nv_bfloat16* X;        // [*, seq, H] fused rows (model LayerNorm applied), row = slot*seq + t; nullptr: not written
  __nv_bfloat16* Xlo;      // optional low part of X (split-bf16, fp32 engine mode)
  float* X32;              // optional fp32 copy of X (fp32 engine mode: exact residual)
  float* pre;              // [*, n_text, H] text embeddings BEFORE the model LayerNorm (text_avg exit), or nullptr
  const int* slot_doc;     // slot -> document (inputs are read by document, outputs written by slot); nullptr: identity
  const int* n_active_dev; // number of slots (nullptr: n_docs)
  int n_docs, n_text, seq, H, coord, shape, vocab, max_2d;
  float eps;
  int* err_flag;           // set to 1 when an input id / box coordinate is outside its table
};

// token -> (slot, t, source token index); false when past the active slots
__device__ __forceinline__ bool text_embed_locate(const TextEmbedArgs& a, int tok, int& slot, int& t, int& src_tok) {
  const int n_slots = a.n_active_dev ? *a.n_active_dev : a.n_docs;
  if (tok >= n_slots * a.n_text) return false;
  slot = tok / a.n_text;
  t = tok - slot * a.n_text;
  const int doc = a.slot_doc ? a.slot_doc[slot] : slot;
  src_tok = doc * a.n_text + t;
  return true;
}
// range check of one token's gather indices (clamped in place)
__device__ __forceinline__ void text_embed_check(const TextEmbedArgs& a, int64_t& id, int& x0, int& y0, int& x1, int& y1) {
  const bool bad = id < 0 || id >= a.vocab || (x0 | y0 | x1 | y1) < 0 || x0 >= a.max_2d || y0 >= a.max_2d ||
                   x1 >= a.max_2d || y1 >= a.max_2d;
  if (bad) {
    *a.err_flag = 1;
    id = min(max(id, static_cast<int64_t>(0)), static_cast<int64_t>(a.vocab - 1));
    x0 = min(max(x0, 0), a.max_2d - 1); y0 = min(max(y0, 0), a.max_2d - 1);
    x1 = min(max(x1, 0), a.max_2d - 1); y1 = min(max(y1, 0), a.max_2d - 1);
  }
}
__device__ __forceinline__ int clamp_coord64(long long v) {       // int64 -> int without wrap-around
  return static_cast<int>(min(max(v, -1ll), 1ll << 20));
}

template <int NV4>
__global__ void text_embed_vec_kernel(const TextEmbedArgs a, EmbedWeights W) {
  const int tok = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int slot, t, src_tok;
  if (!text_embed_locate(a, tok, slot, t, src_tok)) return;
  const int lane = threadIdx.x & 31;
  const int n_text = a.n_text, seq = a.seq, H = a.H, coord = a.coord, shape = a.shape;
  const float eps = a.eps;
  int64_t id = a.ids[src_tok];
  const int pos = a.posid[src_tok];
  const longlong2 b01 = __ldg(reinterpret_cast<const longlong2*>(a.bbox + static_cast<size_t>(src_tok) * 4));
  const longlong2 b23 = __ldg(reinterpret_cast<const longlong2*>(a.bbox + static_cast<size_t>(src_tok) * 4) + 1);
  int x0 = clamp_coord64(b01.x), y0 = clamp_coord64(b01.y), x1 = clamp_coord64(b23.x), y1 = clamp_coord64(b23.y);
  text_embed_check(a, id, x0, y0, x1, y1);
  const int hh = min(max(y1 - y0, 0), 1023), ww = min(max(x1 - x0, 0), 1023);
  const float* wrow = W.word + static_cast<size_t>(id) * H;
  const float* prow = W.pos + static_cast<size_t>(pos) * H;
  float4 wv[NV4], pv[NV4], sv[NV4], tv[NV4];
#pragma unroll
  for (int i = 0; i < NV4; ++i) {                       // all gathers in flight before the first use
    const int c = 4 * (lane + 32 * i);
    wv[i] = __ldg(reinterpret_cast<const float4*>(wrow + c));
    pv[i] = __ldg(reinterpret_cast<const float4*>(prow + c));
    tv[i] = __ldg(reinterpret_cast<const float4*>(W.type0 + c));
    const float* sp;
    if (c < coord) sp = W.x_emb + static_cast<size_t>(x0) * coord + c;
    else if (c < 2 * coord) sp = W.y_emb + static_cast<size_t>(y0) * coord + (c - coord);
    else if (c < 3 * coord) sp = W.x_emb + static_cast<size_t>(x1) * coord + (c - 2 * coord);
    else if (c < 4 * coord) sp = W.y_emb + static_cast<size_t>(y1) * coord + (c - 3 * coord);
    else if (c < 4 * coord + shape) sp = W.h_emb + static_cast<size_t>(hh) * shape + (c - 4 * coord);
    else sp = W.w_emb + static_cast<size_t>(ww) * shape + (c - 4 * coord - shape);
    sv[i] = __ldg(reinterpret_cast<const float4*>(sp));
  }
  float4 v[NV4];
#pragma unroll
  for (int i = 0; i < NV4; ++i) {
    // same association as the reference: ((word + type) + pos) + spatial
    v[i].x = ((wv[i].x + tv[i].x) + pv[i].x) + sv[i].x;
    v[i].y = ((wv[i].y + tv[i].y) + pv[i].y) + sv[i].y;
    v[i].z = ((wv[i].z + tv[i].z) + pv[i].z) + sv[i].z;
    v[i].w = ((wv[i].w + tv[i].w) + pv[i].w) + sv[i].w;
  }
  warp_layernorm4<NV4>(v, H, W.ln_emb_w, W.ln_emb_b, eps, lane);
  if (a.pre) {                                          // text embeddings before the model LayerNorm (text_avg exit)
#pragma unroll
    for (int i = 0; i < NV4; ++i)
      *reinterpret_cast<float4*>(a.pre + (static_cast<size_t>(slot) * n_text + t) * H + 4 * (lane + 32 * i)) = v[i];
  }
  if (!a.X) return;
  warp_layernorm4<NV4>(v, H, W.ln_model_w, W.ln_model_b, eps, lane);
  const size_t orow = (static_cast<size_t>(slot) * seq + t) * H;
#pragma unroll
  for (int i = 0; i < NV4; ++i) store_row4(RowOut{a.X, a.Xlo, a.X32}, orow + 4 * (lane + 32 * i), v[i]);
}

// generic (scalar) variant for shapes the vectorised kernel does not cover (e.g. large: coord 171 / shape 170)
// one warp per text token; X row = doc*seq + t
template <int NV>
__global__ void text_embed_kernel(const TextEmbedArgs a, EmbedWeights W) {
  const int tok = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int slot, t, src_tok;
  if (!text_embed_locate(a, tok, slot, t, src_tok)) return;
  const int lane = threadIdx.x & 31;
  const int n_text = a.n_text, seq = a.seq, H = a.H, coord = a.coord, shape = a.shape;
  const float eps = a.eps;
  int64_t id = a.ids[src_tok];
  const int pos = a.posid[src_tok];
  const int64_t* bb = a.bbox + static_cast<size_t>(src_tok) * 4;
  int x0 = clamp_coord64(bb[0]), y0 = clamp_coord64(bb[1]), x1 = clamp_coord64(bb[2]), y1 = clamp_coord64(bb[3]);
  text_embed_check(a, id, x0, y0, x1, y1);
  const int hh = min(max(y1 - y0, 0), 1023), ww = min(max(x1 - x0, 0), 1023);
  const float* wrow = W.word + static_cast<size_t>(id) * H;
  const float* prow = W.pos + static_cast<size_t>(pos) * H;
  float v[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane + 32 * i;
    float e = 0.f;
    if (c < H) {
      // same association as the reference: ((word + type) + pos) + spatial
      e = __ldg(wrow + c) + __ldg(W.type0 + c);
      e += __ldg(prow + c);
      float sp;
      if (c < coord) sp = __ldg(W.x_emb + static_cast<size_t>(x0) * coord + c);
      else if (c < 2 * coord) sp = __ldg(W.y_emb + static_cast<size_t>(y0) * coord + (c - coord));
      else if (c < 3 * coord) sp = __ldg(W.x_emb + static_cast<size_t>(x1) * coord + (c - 2 * coord));
      else if (c < 4 * coord) sp = __ldg(W.y_emb + static_cast<size_t>(y1) * coord + (c - 3 * coord));
      else if (c < 4 * coord + shape) sp = __ldg(W.h_emb + static_cast<size_t>(hh) * shape + (c - 4 * coord));
      else sp = __ldg(W.w_emb + static_cast<size_t>(ww) * shape + (c - 4 * coord - shape));
      e += sp;
    }
    v[i] = e;
  }
  warp_layernorm<NV>(v, H, W.ln_emb_w, W.ln_emb_b, eps, lane);
  if (a.pre) {
#pragma unroll
    for (int i = 0; i < NV; ++i)
      if (lane + 32 * i < H) a.pre[(static_cast<size_t>(slot) * n_text + t) * H + lane + 32 * i] = v[i];
  }
  if (!a.X) return;
  warp_layernorm<NV>(v, H, W.ln_model_w, W.ln_model_b, eps, lane);
  const size_t orow = (static_cast<size_t>(slot) * seq + t) * H;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane + 32 * i;
    if (c < H) {
      const __nv_bfloat16 hi = __float2bfloat16_rn(v[i]);
      a.X[orow + c] = hi;
      if (a.Xlo) a.Xlo[orow + c] = __float2bfloat16_rn(v[i] - __bfloat162float(hi));
      if (a.X32) a.X32[orow + c] = v[i];
    }
  }
}

// pixels f32 [B,3,img,img] -> patches bf16 [B*np*np, 3*16*16], k = c*256 + kh*16 + kw (conv weight order)
__global__ void im2col_kernel(const float* __restrict__ px, __nv_bfloat16* __restrict__ out,
                              __nv_bfloat16* __restrict__ out_lo, int n_docs, int img, int patch, int chans) {
  const int np = img / patch;
  const int kdim = chans * patch * patch;
  // grid.y = document: all index arithmetic stays 32-bit (the 64-bit divisions of a flat index cost more than the copy)
  const int doc = blockIdx.y;
  const int per_doc = np * np * kdim / 4;                                    // float4 groups per document
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= per_doc || doc >= n_docs) return;
  const int el = i * 4;
  const int k = el % kdim;
  const int p = el / kdim;
  const size_t e = static_cast<size_t>(doc) * np * np * kdim + el;
  const int c = k / (patch * patch), kh = (k / patch) % patch, kw = k % patch;
  const int pr = p / np, pc = p % np;
  const float4 v = *reinterpret_cast<const float4*>(
      px + ((static_cast<size_t>(doc) * chans + c) * img + (pr * patch + kh)) * img + pc * patch + kw);
  store_split4(out + e, out_lo ? out_lo + e : nullptr, v);
}

// one warp per visual token (doc, p); VIS rows 1..n_patch hold conv + bias + pos_embed (written by the patch GEMM)
template <int NV>
__global__ void visual_ln_kernel(float* __restrict__ VIS, EmbedWeights W, RowOut out, int write_pre, int n_docs,
                                 int n_vis, int n_text, int seq, int H, float eps_vis, float eps) {
  __nv_bfloat16* X = out.X;
  __nv_bfloat16* Xlo = out.Xlo;
  const int tok = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (tok >= n_docs * n_vis) return;
  const int lane = threadIdx.x & 31;
  const int doc = tok / n_vis, p = tok - doc * n_vis;
  float v[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane + 32 * i;
    float e = 0.f;
    if (c < H) e = (p == 0) ? (__ldg(W.cls_token + c) + __ldg(W.pos_embed + c))
                            : VIS[(static_cast<size_t>(doc) * n_vis + p) * H + c];
    v[i] = e;
  }
  warp_layernorm<NV>(v, H, W.ln_vis_w, W.ln_vis_b, eps_vis, lane);
  if (write_pre) {                                      // visual embeddings after `norm` (vision_avg exit), in place
#pragma unroll
    for (int i = 0; i < NV; ++i) if (lane + 32 * i < H) VIS[(static_cast<size_t>(doc) * n_vis + p) * H + lane + 32 * i] = v[i];
  }
  if (!X) return;                                       // embedding-level exits pending: embed_finish_vec_kernel writes X
  warp_layernorm<NV>(v, H, W.ln_model_w, W.ln_model_b, eps, lane);
  const size_t orow = (static_cast<size_t>(doc) * seq + n_text + p) * H;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane + 32 * i;
    if (c < H) {
      const __nv_bfloat16 hi = __float2bfloat16_rn(v[i]);
      X[orow + c] = hi;
      if (Xlo) Xlo[orow + c] = __float2bfloat16_rn(v[i] - __bfloat162float(hi));
      if (out.X32) out.X32[orow + c] = v[i];
    }
  }
}

// Vectorised variant for H = 128 * NV4: float4 loads, 8 B stores, grid-stride over the visual tokens.
template <int NV4>
__global__ void __launch_bounds__(256) visual_ln_vec_kernel(float* __restrict__ VIS, EmbedWeights W, RowOut out,
                                                            int write_pre, int n_docs, int n_vis, int n_text, int seq,
                                                            int H, float eps_vis, float eps) {
  __nv_bfloat16* X = out.X;
  const int lane = threadIdx.x & 31;
  const int total = n_docs * n_vis;
  const int warps_total = gridDim.x * (blockDim.x >> 5);
  for (int tok = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); tok < total; tok += warps_total) {
    const int doc = tok / n_vis, p = tok - doc * n_vis;
    float* row = VIS + (static_cast<size_t>(doc) * n_vis + p) * H;
    float4 v[NV4];
#pragma unroll
    for (int i = 0; i < NV4; ++i) {
      const int c = 4 * (lane + 32 * i);
      if (p == 0) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(W.cls_token + c));
        const float4 b = __ldg(reinterpret_cast<const float4*>(W.pos_embed + c));
        v[i] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
      } else {
        v[i] = *reinterpret_cast<const float4*>(row + c);
      }
    }
    warp_layernorm4<NV4>(v, H, W.ln_vis_w, W.ln_vis_b, eps_vis, lane);
    if (write_pre) {                                    // visual embeddings after `norm` (vision_avg exit), in place
#pragma unroll
      for (int i = 0; i < NV4; ++i) *reinterpret_cast<float4*>(row + 4 * (lane + 32 * i)) = v[i];
    }
    if (!X) continue;                                   // embedding-level exits pending: embed_finish_vec_kernel writes X
    warp_layernorm4<NV4>(v, H, W.ln_model_w, W.ln_model_b, eps, lane);
    const size_t orow = (static_cast<size_t>(doc) * seq + n_text + p) * H;
#pragma unroll
    for (int i = 0; i < NV4; ++i) store_row4(out, orow + 4 * (lane + 32 * i), v[i]);
  }
}

// Model LayerNorm of already-embedded rows for the SURVIVORS of an embedding-level exit: X[slot*seq + dst_off + r] =
// LN_model(src[src_map(slot)*src_rows + r]) for the active slots.  src = the text embeddings after embeddings.LayerNorm
// (TXT, indexed by the slot numbering they were computed under; src_map = new -> old slot of the text_avg compaction,
// or nullptr) or the visual embeddings after `norm` (VIS, indexed by document; src_map = slot -> document).
// H = 128 * NV4.  EE/models/LayoutLMv3.py:549-566 (cat + LayerNorm) restricted to the documents that are still there.
template <int NV4>
__global__ void __launch_bounds__(256) embed_finish_vec_kernel(const float* __restrict__ src, int src_rows,
                                                               const int* __restrict__ src_map, EmbedWeights W,
                                                               RowOut out, int dst_off, int seq, int H, float eps,
                                                               const int* __restrict__ n_active_dev) {
  const int lane = threadIdx.x & 31;
  const int total = *n_active_dev * src_rows;
  const int warps_total = gridDim.x * (blockDim.x >> 5);
  for (int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < total; row += warps_total) {
    const int slot = row / src_rows, r = row - slot * src_rows;
    const int sidx = src_map ? src_map[slot] : slot;
    const float* in = src + (static_cast<size_t>(sidx) * src_rows + r) * H;
    float4 v[NV4];
#pragma unroll
    for (int i = 0; i < NV4; ++i) v[i] = *reinterpret_cast<const float4*>(in + 4 * (lane + 32 * i));
    warp_layernorm4<NV4>(v, H, W.ln_model_w, W.ln_model_b, eps, lane);
    const size_t orow = (static_cast<size_t>(slot) * seq + dst_off + r) * H;
#pragma unroll
    for (int i = 0; i < NV4; ++i) store_row4(out, orow + 4 * (lane + 32 * i), v[i]);
  }
}

// pool[doc][c] = mean_t X[doc*seq + t][c]; grid (ceil(H/32), n_docs), block 256 (8 warps stride the tokens)
__global__ void meanpool_kernel(const __nv_bfloat16* __restrict__ X, const float* __restrict__ X32,
                                float* __restrict__ pool, int seq, int H) {
  __shared__ float part[8][33];
  const int doc = blockIdx.y;
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  const int w = threadIdx.x >> 5;
  float s = 0.f;
  if (c < H)
    for (int t = w; t < seq; t += 8) {
      const size_t i = (static_cast<size_t>(doc) * seq + t) * H + c;
      s += X32 ? X32[i] : __bfloat162float(X[i]);
    }
  part[w][threadIdx.x & 31] = s;
  __syncthreads();
  if (w == 0 && c < H) {
    float tot = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) tot += part[i][threadIdx.x];
    pool[static_cast<size_t>(doc) * H + c] = tot / seq;
  }
}

// fp32 variant: pool[doc][c] = mean_t src[(doc*rows + t)][c]  (vision_avg / text_avg exits, EE/models/LayoutLMv3.py:466, 520)
__global__ void meanpool_f32_kernel(const float* __restrict__ src, float* __restrict__ pool, int rows, int H) {
  __shared__ float part[8][33];
  const int doc = blockIdx.y;
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  const int w = threadIdx.x >> 5;
  float s = 0.f;
  if (c < H)
    for (int t = w; t < rows; t += 8) s += src[(static_cast<size_t>(doc) * rows + t) * H + c];
  part[w][threadIdx.x & 31] = s;
  __syncthreads();
  if (w == 0 && c < H) {
    float tot = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) tot += part[i][threadIdx.x];
    pool[static_cast<size_t>(doc) * H + c] = tot / rows;
  }
}

}  // namespace mmee

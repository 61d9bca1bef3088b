"""TEST INFRASTRUCTURE — CPU restatement ("port") of the reference hot path in plain torch fp32.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this module; the product path (`mmee`) never does.

Parity status: the reference ships no golden vectors or tests, so parity is
pinned by OUR fixtures: `tests/golden/*.npz` hold outputs of the unmodified reference model
(run through `oracle/reference_harness.py` in the dev container by
`tests/golden/make_golden.py`); `tests/test_oracle.py` checks this port against them.

Each function cites the reference lines it restates ("HF:" = transformers
models/layoutlmv3/modeling_layoutlmv3.py v5.5.0, the third-party arithmetic the reference
subclasses; other paths are under the reference's EE/).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F


def _ln(x, w, b, eps):
    return F.layer_norm(x, (x.shape[-1],), w, b, eps)


def visual_bbox(n_side: int = 14, max_len: int = 1000) -> torch.Tensor:
    """HF:576-602 create_visual_bbox: CLS box [1,1,999,999] then the 14x14 grid boxes."""
    edges = torch.div(torch.arange(0, max_len * (n_side + 1), max_len), n_side, rounding_mode="trunc")
    x0 = edges[:-1].repeat(n_side, 1)
    y0 = edges[:-1].repeat(n_side, 1).transpose(0, 1)
    x1 = edges[1:].repeat(n_side, 1)
    y1 = edges[1:].repeat(n_side, 1).transpose(0, 1)
    grid = torch.stack([x0, y0, x1, y1], dim=-1).view(-1, 4)
    cls = torch.tensor([[1, 1, max_len - 1, max_len - 1]])
    return torch.cat([cls, grid], dim=0)


def relative_position_bucket(rel: torch.Tensor, num_buckets: int, max_distance: int) -> torch.Tensor:
    """HF:393-414 (bidirectional=True).  fp32 log, truncation toward zero."""
    nb = num_buckets // 2
    ret = (rel > 0).long() * nb
    n = rel.abs()
    max_exact = nb // 2
    is_small = n < max_exact
    large = max_exact + (
        torch.log(n.float() / max_exact) / math.log(max_distance / max_exact) * (nb - max_exact)
    ).to(torch.long)
    large = torch.min(large, torch.full_like(large, nb - 1))
    return ret + torch.where(is_small, n, large)


def text_embeddings(sd, dims, input_ids, bbox):
    """HF:161-200 with position_ids=None (HF:139-147) and token_type_ids=0; spatial HF:113-137."""
    p = "layoutlmv3.embeddings."
    if input_ids.shape[1] == 0:                             # image-only path: no text tokens (HF:730-768 with pixel_values only)
        return torch.zeros((input_ids.shape[0], 0, dims.hidden), dtype=torch.float32)
    mask = input_ids.ne(dims.pad_id).int()
    pos = (torch.cumsum(mask, dim=1).type_as(mask) * mask).long() + dims.pad_id
    e = sd[p + "word_embeddings.weight"][input_ids] + sd[p + "token_type_embeddings.weight"][0]
    e = e + sd[p + "position_embeddings.weight"][pos]
    xw, yw = sd[p + "x_position_embeddings.weight"], sd[p + "y_position_embeddings.weight"]
    hw, ww = sd[p + "h_position_embeddings.weight"], sd[p + "w_position_embeddings.weight"]
    spatial = torch.cat([
        xw[bbox[..., 0]], yw[bbox[..., 1]], xw[bbox[..., 2]], yw[bbox[..., 3]],
        hw[torch.clip(bbox[..., 3] - bbox[..., 1], 0, 1023)],
        ww[torch.clip(bbox[..., 2] - bbox[..., 0], 0, 1023)],
    ], dim=-1)
    e = e + spatial
    return _ln(e, sd[p + "LayerNorm.weight"], sd[p + "LayerNorm.bias"], dims.ln_eps)


def visual_embeddings(sd, dims, pixel_values):
    """EE/models/LayoutLMv3.py:358-373 forward_image; HF:70-82 patch conv."""
    p = "layoutlmv3."
    x = F.conv2d(pixel_values, sd[p + "patch_embed.proj.weight"], sd[p + "patch_embed.proj.bias"],
                 stride=dims.patch)
    x = x.flatten(2).transpose(1, 2)
    cls = sd[p + "cls_token"].expand(x.shape[0], -1, -1)
    x = torch.cat([cls, x], dim=1) + sd[p + "pos_embed"]
    return _ln(x, sd[p + "norm.weight"], sd[p + "norm.bias"], dims.vis_ln_eps)


def fused_embeddings(sd, dims, docs=None, *, input_ids=None, bbox=None, pixel_values=None):
    """EE/models/LayoutLMv3.py:441-463, 511-517, 549-566: [text | visual] then model LayerNorm."""
    if docs is not None:
        input_ids, bbox, pixel_values = docs["input_ids"], docs["bbox"], docs["pixel_values"]
    t = text_embeddings(sd, dims, input_ids, bbox)
    v = visual_embeddings(sd, dims, pixel_values)
    x = torch.cat([t, v], dim=1)
    return _ln(x, sd["layoutlmv3.LayerNorm.weight"], sd["layoutlmv3.LayerNorm.bias"], dims.ln_eps)


def attention_bias(sd, dims, bbox_text):
    """HF:416-458 with the fused position ids / boxes of EE/models/LayoutLMv3.py:556-563.
    Returns (rel_pos + rel_2d_pos) [B,h,S,S] (NOT yet divided by sqrt(d))."""
    B = bbox_text.shape[0]
    n_side = dims.image // dims.patch
    vb = visual_bbox(n_side).unsqueeze(0).expand(B, -1, -1)
    bb = torch.cat([bbox_text, vb], dim=1)
    pos = torch.cat([torch.arange(dims.n_text), torch.arange(dims.n_vis)]).unsqueeze(0)
    rel = pos.unsqueeze(-2) - pos.unsqueeze(-1)
    b1 = relative_position_bucket(rel, dims.rel_bins, dims.max_rel)
    r1 = sd["layoutlmv3.encoder.rel_pos_bias.weight"].t()[b1].permute(0, 3, 1, 2)
    cx, cy = bb[:, :, 0], bb[:, :, 3]
    bx = relative_position_bucket(cx.unsqueeze(-2) - cx.unsqueeze(-1), dims.rel2d_bins, dims.max_rel2d)
    by = relative_position_bucket(cy.unsqueeze(-2) - cy.unsqueeze(-1), dims.rel2d_bins, dims.max_rel2d)
    rx = sd["layoutlmv3.encoder.rel_pos_x_bias.weight"].t()[bx].permute(0, 3, 1, 2)
    ry = sd["layoutlmv3.encoder.rel_pos_y_bias.weight"].t()[by].permute(0, 3, 1, 2)
    return r1 + (rx + ry)


def encoder_layer(sd, dims, i, x, bias, ext_mask, parts: Optional[dict] = None):
    """HF:236-369: post-LN BERT layer with CogView softmax (HF:224-234) and exact erf GELU."""
    L = f"layoutlmv3.encoder.layer.{i}."
    B, S, H = x.shape
    h, d = dims.heads, dims.head_dim

    def lin(name, t):
        return F.linear(t, sd[L + name + ".weight"], sd[L + name + ".bias"])

    q = lin("attention.self.query", x).view(B, S, h, d).transpose(1, 2)
    k = lin("attention.self.key", x).view(B, S, h, d).transpose(1, 2)
    v = lin("attention.self.value", x).view(B, S, h, d).transpose(1, 2)
    s = torch.matmul(q / math.sqrt(d), k.transpose(-1, -2))
    s = s + bias / math.sqrt(d)
    s = s + ext_mask
    alpha = 32
    ss = s / alpha
    mx = ss.amax(dim=-1, keepdim=True)
    p_ = torch.softmax((ss - mx) * alpha, dim=-1)
    ctx = torch.matmul(p_, v).permute(0, 2, 1, 3).contiguous().view(B, S, H)
    a = _ln(lin("attention.output.dense", ctx) + x,
            sd[L + "attention.output.LayerNorm.weight"], sd[L + "attention.output.LayerNorm.bias"], dims.ln_eps)
    m = F.gelu(lin("intermediate.dense", a))
    y = _ln(lin("output.dense", m) + a,
            sd[L + "output.LayerNorm.weight"], sd[L + "output.LayerNorm.bias"], dims.ln_eps)
    if parts is not None:
        parts.update(q=q, k=k, v=v, ctx=ctx, attn_out=a, mlp=m)
    return y


def exit_head(sd, prefix, x):
    """EE/models/LayoutLMv3.py:86-93 (dropout = identity in eval); also HF:816-822 for `classifier`."""
    if prefix + ".dense.weight" in sd:
        x = torch.tanh(F.linear(x, sd[prefix + ".dense.weight"], sd[prefix + ".dense.bias"]))
    return F.linear(x, sd[prefix + ".out_proj.weight"], sd[prefix + ".out_proj.bias"])


def entropy(x):
    """EE/models/EE_modules.py:149-154 (un-stabilised, nats)."""
    ex = torch.exp(x)
    a = ex.sum(dim=1)
    b = (x * ex).sum(dim=1)
    return torch.log(a) - b / a


def max_confidence(x):
    """EE/models/EE_modules.py:157-160."""
    return torch.softmax(x, dim=1).max(dim=1)[0]


def forward(sd, dims, ee, docs, keep_hidden: bool = False) -> Dict[str, torch.Tensor]:
    """Dense (all exits, all docs) forward = what the reference computes
    (EE/models/LayoutLMv3.py:696-748, 375-665, 151-305).

    Returns exit_logits [E+1,B,K] (per-exit class logits as stored by EE/utils.py:182-193: ramp
    logits, or classifier(CLS_j) "gated logits" in gate mode :764-782; final classifier last),
    head_logits [E,B,K|2] (raw head outputs, = exit_states[j][0]), cls_rows [E+1,B,H]."""
    sd = {k: v.float() for k, v in sd.items()}
    tvis = visual_embeddings(sd, dims, docs["pixel_values"])
    ttxt = text_embeddings(sd, dims, docs["input_ids"], docs["bbox"])
    x = _ln(torch.cat([ttxt, tvis], dim=1), sd["layoutlmv3.LayerNorm.weight"], sd["layoutlmv3.LayerNorm.bias"],
            dims.ln_eps)                                     # == fused_embeddings
    B = x.shape[0]
    mask = torch.cat([docs["attention_mask"].to(torch.float32), torch.ones(B, dims.n_vis)], dim=1)
    ext = (1.0 - mask)[:, None, None, :] * torch.finfo(torch.float32).min   # modeling_utils get_extended_attention_mask
    bias = attention_bias(sd, dims, docs["bbox"])
    gate = ee.encoder_layer_strategy == "gate"
    rows: List[torch.Tensor] = []
    heads: List[torch.Tensor] = []
    hidden = [x] if keep_hidden else None
    if "vision_avg" in ee.exits:
        z = tvis.mean(1)                                    # :465-468: mean of the visual embeddings (after `norm`)
        rows.append(z)
        heads.append(exit_head(sd, "layoutlmv3.vision_exit_embeddings", z))
    if "text_avg" in ee.exits:
        z = ttxt.mean(1)                                    # :519-521: mean of the text embeddings (pads included)
        rows.append(z)
        heads.append(exit_head(sd, "layoutlmv3.text_exit_embeddings", z))
    if ee.has_concat_exit:
        z = x.mean(1)                                       # :582 (pads included)
        rows.append(z)
        heads.append(exit_head(sd, "layoutlmv3.concat_exit_embeddings", z))
    exit_layers = ee.encoder_exit_layers
    for i in range(dims.layers):
        x = encoder_layer(sd, dims, i, x, bias, ext)
        if keep_hidden:
            hidden.append(x)
        if (i + 1) in exit_layers:
            k = exit_layers.index(i + 1)
            z = x[:, 0, :]                                  # :226
            rows.append(z)
            heads.append(exit_head(sd, f"layoutlmv3.encoder.early_exits.{k}", z))
    final_row = x[:, 0, :]
    final = exit_head(sd, "classifier", final_row)          # :730-731
    per_exit = [exit_head(sd, "classifier", z) for z in rows] if gate else heads
    out = {
        "exit_logits": torch.stack(per_exit + [final]),
        "head_logits": torch.stack(heads) if heads else torch.zeros(0),
        "cls_rows": torch.stack(rows + [final_row]),
        "last_hidden": x,
    }
    if keep_hidden:
        out["hidden"] = torch.stack(hidden)
    return out


def lte_scores(sd, out) -> torch.Tensor:
    """sigmoid(lte_classifier(exit input row)) for every exit row [E+1, B]
    (EE/models/LayoutLMv3.py:142-149, 231-237; model level :597-602)."""
    w = sd["layoutlmv3.encoder.lte_classifier.weight"].float()
    b = sd["layoutlmv3.encoder.lte_classifier.bias"].float()
    return torch.sigmoid(out["cls_rows"] @ w.t() + b).squeeze(-1)


def lte_exit(sd, ee, out, threshold: float) -> Dict[str, torch.Tensor]:
    """Learned-to-exit decision per document (the reference evaluates it for one document at a time,
    EE/models/LayoutLMv3.py:250-268): the first ENCODER exit after layer l with l < len(exit_encoder_layers)
    (`i + 1 < self.num_layers`, :139, :252) whose score is < threshold (`lte_th = [exit_threshold] * num_layers`,
    :147-149) is taken and its class logits returned (:739-748); embedding-level exits are scored only.
    `out` is the dict of `forward`.  Returns exit_index [B] (index into the E+1 exits), logits [B,K], scores."""
    scores = lte_scores(sd, out)
    names = [x for x in ("vision_avg", "text_avg", "text_visual_concat") if x in ee.exits]
    enc = sorted(ee.encoder_exit_layers)
    layer_of = [0] * len(names) + enc                       # 0 = embedding level
    E = len(layer_of)
    B = scores.shape[1]
    exit_index = torch.full((B,), E, dtype=torch.int64)
    for d in range(B):
        for e in range(E):
            if 0 < layer_of[e] < len(enc) and scores[e, d] < threshold:
                exit_index[d] = e
                break
    logits = out["exit_logits"][exit_index, torch.arange(B)]
    return {"exit_index": exit_index, "logits": logits, "scores": scores,
            "exit_layer": torch.tensor([layer_of[e] if e < E else 0 for e in exit_index.tolist()])}

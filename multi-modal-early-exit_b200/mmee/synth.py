"""Seed-deterministic synthetic weights and RVL-CDIP-shaped documents.

Weights follow the reference's state-dict names (dumped from
`LayoutLMv3EEForSequenceClassification`, EE/models/LayoutLMv3.py:669-694; list
in SURVEY.md Appendix A.6) so the same dict loads into the reference model
(strict) and into the engine.  All tensors are N(0, 0.02) except LayerNorm
weights N(1, 0.02): unlike HF's init, biases / cls_token / pos_embed are
non-zero so every code path is exercised.

Documents follow the dataset feature contract EE/data/RVL_CDIP.py:223-246:
input_ids/attention_mask i64[512], bbox i64[512,4] in [0,1000],
pixel_values f32[3,224,224], labels i64.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, List, Tuple

import torch

from .config import ExitConfig, ModelDims


def state_dict_spec(dims: ModelDims, ee: ExitConfig) -> "OrderedDict[str, Tuple[Tuple[int, ...], str]]":
    """name -> (shape, kind) with kind in {"w", "ln_w"}; order is the generation order."""
    H, I, h = dims.hidden, dims.inter, dims.heads
    spec: "OrderedDict[str, Tuple[Tuple[int, ...], str]]" = OrderedDict()

    def lin(prefix, out_f, in_f):
        spec[prefix + ".weight"] = ((out_f, in_f), "w")
        spec[prefix + ".bias"] = ((out_f,), "w")

    def ln(prefix):
        spec[prefix + ".weight"] = ((H,), "ln_w")
        spec[prefix + ".bias"] = ((H,), "w")

    def head(prefix, out_f):
        if ee.exit_head_num_layers == 2:
            lin(prefix + ".dense", H, H)
        lin(prefix + ".out_proj", out_f, H)

    p = "layoutlmv3."
    spec[p + "cls_token"] = ((1, 1, H), "w")
    spec[p + "pos_embed"] = ((1, dims.n_vis, H), "w")
    e = p + "embeddings."
    spec[e + "word_embeddings.weight"] = ((dims.vocab, H), "w")
    spec[e + "token_type_embeddings.weight"] = ((1, H), "w")
    spec[e + "position_embeddings.weight"] = ((dims.max_pos, H), "w")
    spec[e + "x_position_embeddings.weight"] = ((dims.max_2d, dims.coord), "w")
    spec[e + "y_position_embeddings.weight"] = ((dims.max_2d, dims.coord), "w")
    spec[e + "h_position_embeddings.weight"] = ((dims.max_2d, dims.shape), "w")
    spec[e + "w_position_embeddings.weight"] = ((dims.max_2d, dims.shape), "w")
    ln(e + "LayerNorm")
    spec[p + "patch_embed.proj.weight"] = ((H, dims.channels, dims.patch, dims.patch), "w")
    spec[p + "patch_embed.proj.bias"] = ((H,), "w")
    ln(p + "LayerNorm")
    ln(p + "norm")
    for i in range(dims.layers):
        L = f"{p}encoder.layer.{i}."
        lin(L + "attention.self.query", H, H)
        lin(L + "attention.self.key", H, H)
        lin(L + "attention.self.value", H, H)
        lin(L + "attention.output.dense", H, H)
        ln(L + "attention.output.LayerNorm")
        lin(L + "intermediate.dense", I, H)
        lin(L + "output.dense", H, I)
        ln(L + "output.LayerNorm")
    spec[p + "encoder.rel_pos_bias.weight"] = ((h, dims.rel_bins), "w")
    spec[p + "encoder.rel_pos_x_bias.weight"] = ((h, dims.rel2d_bins), "w")
    spec[p + "encoder.rel_pos_y_bias.weight"] = ((h, dims.rel2d_bins), "w")
    out_f = dims.n_labels if ee.encoder_layer_strategy == "ramp" else 2
    for k, _layer in enumerate(ee.encoder_exit_layers):
        head(f"{p}encoder.early_exits.{k}", out_f)
    if "vision_avg" in ee.exits:
        head(p + "vision_exit_embeddings", out_f)
    if "text_avg" in ee.exits:
        head(p + "text_exit_embeddings", out_f)
    if ee.has_concat_exit:
        head(p + "concat_exit_embeddings", out_f)
    lin("classifier.dense", H, H)
    lin("classifier.out_proj", dims.n_labels, H)
    if getattr(ee, "use_lte", False):
        lin(p + "encoder.lte_classifier", 1, H)             # EE/models/LayoutLMv3.py:144 (appended last: earlier
    return spec                                             # tensors keep their seeded values)


def make_state_dict(dims: ModelDims, ee: ExitConfig, seed: int = 0, std: float = 0.02,
                    head_gain: float = 1.0) -> Dict[str, torch.Tensor]:
    """fp32 CPU state dict.  `head_gain` scales every exit head's / classifier's out_proj weight
    (the documented alternative knob of SURVEY.md §8(d) to spread confidences at random init)."""
    g = torch.Generator().manual_seed(1000003 * (seed + 1))
    sd: Dict[str, torch.Tensor] = {}
    for name, (shape, kind) in state_dict_spec(dims, ee).items():
        t = torch.randn(shape, generator=g, dtype=torch.float32) * std
        if kind == "ln_w":
            t = t + 1.0
        if head_gain != 1.0 and name.endswith("out_proj.weight"):
            t = t * head_gain
        sd[name] = t
    return sd


def make_docs(dims: ModelDims, n: int, seed: int = 1, pad: bool = True) -> Dict[str, torch.Tensor]:
    """n synthetic documents (SURVEY.md §8(d) recipe).  pad=False -> all n_text tokens real."""
    g = torch.Generator().manual_seed(7919 * (seed + 1))
    T = dims.n_text
    if T == 0:                                              # image-only documents (BASELINE config 5): patch tokens only
        pixels = torch.rand((n, dims.channels, dims.image, dims.image), generator=g, dtype=torch.float32) * 2 - 1
        labels = torch.randint(0, dims.n_labels, (n,), generator=g, dtype=torch.int64)
        z = torch.zeros((n, 0), dtype=torch.int64)
        return {"input_ids": z, "attention_mask": z.clone(), "bbox": torch.zeros((n, 0, 4), dtype=torch.int64),
                "pixel_values": pixels, "labels": labels}
    ids = torch.randint(3, dims.vocab, (n, T), generator=g, dtype=torch.int64)
    if pad:
        lens = torch.randint(min(64, T), T + 1, (n,), generator=g, dtype=torch.int64)
        lens[0] = T                                         # always one unpadded document
    else:
        lens = torch.full((n,), T, dtype=torch.int64)
    x0 = torch.randint(0, 900, (n, T), generator=g, dtype=torch.int64)
    y0 = torch.randint(0, 900, (n, T), generator=g, dtype=torch.int64)
    w = torch.randint(1, 101, (n, T), generator=g, dtype=torch.int64)
    hh = torch.randint(1, 101, (n, T), generator=g, dtype=torch.int64)
    bbox = torch.stack([x0, y0, x0 + w, y0 + hh], dim=-1)
    pos = torch.arange(T).unsqueeze(0)
    real = pos < lens.unsqueeze(1)
    ids[:, 0] = 0                                           # <s>
    last = (lens - 1).clamp(min=1)
    ids[torch.arange(n), last] = 2                          # </s>
    ids = torch.where(real, ids, torch.full_like(ids, dims.pad_id))
    special = (~real) | (pos == 0) | (pos == last.unsqueeze(1))
    bbox = torch.where(special.unsqueeze(-1), torch.zeros_like(bbox), bbox)
    pixels = torch.rand((n, dims.channels, dims.image, dims.image), generator=g, dtype=torch.float32) * 2 - 1
    labels = torch.randint(0, dims.n_labels, (n,), generator=g, dtype=torch.int64)
    return {
        "input_ids": ids,
        "attention_mask": real.to(torch.int64),
        "bbox": bbox,
        "pixel_values": pixels,
        "labels": labels,
    }


def exit_names(ee: ExitConfig) -> List[str]:
    """Model-level exit order (EE/models/LayoutLMv3.py:481,532,595,649): concat first, then layers."""
    names: List[str] = []
    if ee.has_concat_exit:
        names.append("text_visual_concat")
    names += [f"layer_{i}" for i in ee.encoder_exit_layers]
    return names

"""Engine-side mirror of the reference's model + exit configuration.

`ExitConfig` keeps the field names of the reference's `ExitConfig`
(EE/models/EE_modules.py:175-195: training_strategy, inference_strategy,
global_threshold, exits, encoder_layer_strategy, exit_head_num_layers) so a
dict written for the reference (`config.EE_config`) configures the engine
unchanged.  `ModelDims` carries the HuggingFace LayoutLMv3 dimensions the
reference inherits (HF configuration_layoutlmv3.py; values used by the
reference at EE/models/LayoutLMv3.py:311-356).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Sequence, Union

EMBEDDING_EXITS = ("vision_avg", "text_avg", "text_visual_concat")


@dataclass
class ExitConfig:
    """Same keys as the reference's ExitConfig (EE/models/EE_modules.py:175-195)."""

    training_strategy: str = "joint_weighted_avg"
    inference_strategy: str = "max_confidence"      # "max_confidence" (>) | "entropy" (<)
    global_threshold: float = 0.9
    exits: Sequence[Union[str, int]] = ("text_visual_concat", 1, 4, 8)
    encoder_layer_strategy: str = "ramp"            # "ramp" | "gate"
    exit_head_num_layers: int = 2
    model_weights: str = ""                         # accepted and ignored (reference: processor name)
    use_lte: bool = False                           # EE_config["use_lte"] (EE/models/LayoutLMv3.py:140): learned-to-exit

    @classmethod
    def from_dict(cls, d: dict) -> "ExitConfig":
        known = {k: v for k, v in d.items() if k in cls.__dataclass_fields__}
        cfg = cls(**known)
        cfg.inference_strategy = str(cfg.inference_strategy)
        cfg.use_lte = bool(cfg.use_lte) and str(cfg.use_lte) != "False"
        cfg.encoder_layer_strategy = str(cfg.encoder_layer_strategy)
        if isinstance(cfg.exits, str):  # reference accepts "a,b,1,2" (EE/models/LayoutLMv3.py:100-108)
            parsed: List[Union[str, int]] = []
            for e in cfg.exits.split(","):
                try:
                    parsed.append(int(e))
                except ValueError:
                    parsed.append(e)
            cfg.exits = parsed
        if cfg.inference_strategy not in ("max_confidence", "entropy"):
            raise NotImplementedError(f"{cfg.inference_strategy} not implemented")
        if cfg.encoder_layer_strategy not in ("ramp", "gate"):
            raise NotImplementedError(f"{cfg.encoder_layer_strategy} not implemented")
        return cfg

    @property
    def encoder_exit_layers(self) -> List[int]:
        return [e for e in self.exits if isinstance(e, int)]

    @property
    def has_concat_exit(self) -> bool:
        return "text_visual_concat" in self.exits


@dataclass
class ModelDims:
    hidden: int = 768
    layers: int = 12
    heads: int = 12
    inter: int = 3072
    n_text: int = 512
    image: int = 224
    patch: int = 16
    channels: int = 3
    n_labels: int = 16
    coord: int = 128
    shape: int = 128
    vocab: int = 50265
    max_pos: int = 514
    max_2d: int = 1024
    rel_bins: int = 32
    max_rel: int = 128
    rel2d_bins: int = 64
    max_rel2d: int = 256
    pad_id: int = 1
    ln_eps: float = 1e-5
    vis_ln_eps: float = 1e-6

    @property
    def head_dim(self) -> int:
        return self.hidden // self.heads

    @property
    def n_patch(self) -> int:
        return (self.image // self.patch) ** 2

    @property
    def n_vis(self) -> int:
        return self.n_patch + 1

    @property
    def seq(self) -> int:
        return self.n_text + self.n_vis

    def check(self) -> None:
        assert self.hidden % self.heads == 0 and self.head_dim == 64, "engine is built for head_dim 64"
        assert 4 * self.coord + 2 * self.shape == self.hidden
        assert self.hidden % 64 == 0 and self.inter % 64 == 0

    @classmethod
    def base(cls, **kw) -> "ModelDims":
        return cls(**kw)

    @classmethod
    def large(cls, **kw) -> "ModelDims":
        d = dict(hidden=1024, layers=24, heads=16, inter=4096, coord=171, shape=170)
        d.update(kw)
        return cls(**d)

    @classmethod
    def tiny(cls, **kw) -> "ModelDims":
        """Small shape for fast tests; same structure (head_dim 64, 709-token sequence)."""
        d = dict(hidden=128, layers=3, heads=2, inter=256, coord=22, shape=20, vocab=1000)
        d.update(kw)
        return cls(**d)

    def to_hf_config(self):
        """HF LayoutLMv3Config with these dims (used by the oracle harness and the reference arm)."""
        from transformers import LayoutLMv3Config

        return LayoutLMv3Config(
            vocab_size=self.vocab, hidden_size=self.hidden, num_hidden_layers=self.layers,
            num_attention_heads=self.heads, intermediate_size=self.inter,
            max_position_embeddings=self.max_pos, type_vocab_size=1, num_labels=self.n_labels,
            coordinate_size=self.coord, shape_size=self.shape, input_size=self.image,
            patch_size=self.patch, num_channels=self.channels, rel_pos_bins=self.rel_bins,
            max_rel_pos=self.max_rel, rel_2d_pos_bins=self.rel2d_bins, max_rel_2d_pos=self.max_rel2d,
            max_2d_position_embeddings=self.max_2d, layer_norm_eps=self.ln_eps, pad_token_id=self.pad_id,
        )

"""Shared helpers for the parity tests: load a golden fixture and rebuild its seeded inputs."""
import json
import os

import numpy as np

from mmee import synth
from mmee.config import ExitConfig, ModelDims

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ALL_CASES = ["tiny_ramp_conf", "tiny_gate_ent", "tiny_ramp_1layer_head",
             "base_ramp_conf", "base_gate_ent", "large4_ramp_conf", "large24_ramp2",
             "tiny_modality_ramp", "tiny_modality_gate", "tiny_image_only", "base2_image_only"]
LTE_CASES = ["tiny_lte_ramp", "tiny_lte_gate", "base4_lte_ramp"]      # learned-to-exit (EE_config["use_lte"])


def load_case(name):
    """-> (golden npz dict, dims, ee, state_dict, docs) — inputs regenerated from the stored seeds."""
    g = dict(np.load(os.path.join(GOLDEN_DIR, f"{name}.npz")))
    m = json.loads(str(g["meta"]))
    dims = getattr(ModelDims, m["ctor"])(**m["dims_kw"])
    ee = ExitConfig.from_dict(m["ee"])
    sd = synth.make_state_dict(dims, ee, seed=m["wseed"], std=m.get("std", 0.02))
    docs = synth.make_docs(dims, m["n"], seed=m["dseed"], pad=m["pad"])
    # the fixture records input checksums so a drift in the generator is caught, not silently absorbed
    assert np.array_equal(docs["input_ids"].sum(1).numpy(), g["input_ids_sum"]), "synthetic input drift"
    assert np.array_equal(docs["bbox"].sum((1, 2)).numpy(), g["bbox_sum"]), "synthetic bbox drift"
    assert np.allclose(docs["pixel_values"].double().sum((1, 2, 3)).numpy(), g["pixel_sum"]), "pixel drift"
    return g, dims, ee, sd, docs

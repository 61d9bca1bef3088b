"""Shared helpers for the parity tests: load a golden fixture and rebuild its seeded inputs."""
import json
import os

import numpy as np

from mmee import synth
from mmee.config import ExitConfig, ModelDims

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ALL_CASES = ["tiny_ramp_conf", "tiny_gate_ent", "tiny_ramp_1layer_head",
             "base_ramp_conf", "base_gate_ent", "config0_base_ramp16", "large4_ramp_conf", "large24_ramp2",
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


def port_forward_chunked(sd, dims, ee, docs, chunk=16):
    """oracle.port.forward over a large batch in chunks (documents are independent; the port materialises
    [B, heads, S, S] fp32 bias tensors, 25 MB per base document and several copies of it)."""
    import torch
    from oracle import port

    n = docs["input_ids"].shape[0]
    parts = []
    for i in range(0, n, chunk):
        sub = {k: v[i:i + chunk] for k, v in docs.items()}
        parts.append(port.forward(sd, dims, ee, sub)["exit_logits"])
    return torch.cat(parts, dim=1)


def exit_agreement(got_exits, all_logits_engine, ref_logits, temps, thr, kind):
    """Agreement of the engine's exit indices with the reference policy, over ALL documents and over the decisive
    ones.  A document is decisive when at every exit its reference criterion keeps more than the engine's own
    worst-case criterion error from the threshold: for max-softmax |d conf| <= 0.5 * max_k|d z_k| / T_e (softmax
    Jacobian bound), for entropy |d H| <= (log K + 1) * ... is replaced by the measured |H(engine) - H(ref)|.  The bound
    is computed per (exit, document) from the engine's DENSE logits, so an early-exit decision that differs from the
    reference outside it is a real error, not rounding.  Returns dict(all, decisive, decisive_frac, want)."""
    import numpy as np
    from oracle import policy_port

    ref_cal = policy_port.temperature_scale(np.asarray(ref_logits, dtype=np.float64), temps)
    eng_cal = policy_port.temperature_scale(np.asarray(all_logits_engine, dtype=np.float64), temps)
    want, _, crit_ref = policy_port.exit_policy_vectorised(ref_cal, thr, kind)
    crit_eng = policy_port.criterion(eng_cal, kind)
    if crit_ref.shape[0] > 1:
        err = np.abs(crit_eng - crit_ref)[:-1]                      # engine's own criterion error per (exit, doc)
        thr_rows = np.broadcast_to(np.asarray(thr, dtype=np.float64).reshape(-1), (crit_ref.shape[0],)) \
            if np.size(thr) == 1 else np.asarray(thr, dtype=np.float64)
        margin = np.abs(crit_ref[:-1] - thr_rows[:-1, None])
        decisive = (margin > 1.5 * err + 1e-7).all(axis=0)
    else:
        decisive = np.ones(want.shape, bool)
    agree = np.asarray(got_exits) == want
    return dict(all=float(agree.mean()), decisive=float(agree[decisive].mean()) if decisive.any() else 1.0,
                decisive_frac=float(decisive.mean()), want=want, agree=agree, decisive_mask=decisive)

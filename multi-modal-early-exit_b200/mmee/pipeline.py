"""Batched stand-in for the reference's logits-store loop:

    get_logits(model, config, test_loader)  -> (logits f64 [E+1, N, K], references i64 [N], None)      EE/utils.py:125-223
    dump_logits(model, logits, references, config, name)                                               EE/utils.py:240-271

The reference walks the DataLoader with batch size 1 and writes one row of the f64 store per step
(EE/utils.py:169-193); here every batch of the loader (any size <= the engine's max_batch) is one engine forward and
fills a block of rows.  File names, array keys ("arr_0"), dtypes, the cache short-cut (EE/utils.py:150-163) and the
config.json clean-up (EE/utils.py:258-270) are the reference's, so `eval.py`, `thresh.py` and `large_scale.py` read
the results unchanged.  Which logits are stored per exit follows EE/utils.py:182-193: `gated_logits[j]` when the
model produced them (gate heads), else `exit_states[j][0]`; the last row is `outputs.logits`.
"""
from __future__ import annotations

import json
import os
from typing import Iterable, Optional, Tuple

import numpy as np
import torch

POPPED_KEYS = ("exit_threshold", "global_threshold", "inference_strategy", "exit_policy", "use_lte", "use_wandb",
               "calibrate", "full_test", "step", "epsilon")          # EE/utils.py:260-269


def config_to_checkpoint(config) -> str:
    """EE/utils.py:114-122."""
    output_path = os.path.join("results", f"{config['checkpoint'].split('/')[-1]}-{config['test_dataset'].split('/')[-1]}")
    if config.get("downsampling"):
        output_path += f"-{config['downsampling']}i"
    return os.path.join(config.get("results_root", ""), output_path)


def dump_logits(model, logits, references, config, name: str = "test") -> str:
    """EE/utils.py:240-271: references-<name>.npz, exit_logits-<name>.npz (key arr_0) and config.json."""
    output_path = config_to_checkpoint(config)
    os.makedirs(output_path, exist_ok=True)
    if references is not None:
        np.savez_compressed(os.path.join(output_path, f"references-{name}.npz"), np.asarray(references))
    if isinstance(logits, torch.Tensor):
        logits = logits.cpu().numpy()
    np.savez_compressed(os.path.join(output_path, f"exit_logits-{name}.npz"), np.asarray(logits))
    to_save = dict(config)
    exit_config = getattr(getattr(model, "config", None), "exit_config", None) or getattr(model, "exit_config", None)
    if exit_config:
        to_save.update(exit_config)
    for k in POPPED_KEYS:
        to_save.pop(k, None)
    to_save.pop("results_root", None)
    with open(os.path.join(output_path, "config.json"), "w+") as f:
        json.dump(to_save, f, indent=4, default=str)
    return output_path


def get_logits(model, config, test_loader: Iterable[dict]) -> Tuple[np.ndarray, np.ndarray, Optional[object]]:
    """Same contract as EE/utils.py:125-223 (minus the OCR benchmark / plotting side channels)."""
    batches = list(test_loader) if not isinstance(test_loader, list) else test_loader
    references = np.concatenate([np.asarray(b["labels"]).reshape(-1) for b in batches])
    if config.get("downsampling"):
        references = references[: config["downsampling"]]
    N = len(references)
    labelset = config.get("labelset", "test")
    output_path = config_to_checkpoint(config)
    f_log = os.path.join(output_path, f"exit_logits-{labelset}.npz")
    f_ref = os.path.join(output_path, f"references-{labelset}.npz")
    if os.path.exists(f_log) and os.path.exists(f_ref):                          # EE/utils.py:150-163
        return np.load(f_log)["arr_0"], np.load(f_ref)["arr_0"], None
    nr_exits = model.n_exits
    K = model.dims.n_labels
    logits_store = np.zeros((nr_exits + 1, N, K), dtype=np.float64)              # EE/utils.py:160-164
    correct = 0
    row = 0
    for batch in batches:
        if row >= N:
            break
        take = min(int(np.asarray(batch["labels"]).reshape(-1).shape[0]), N - row)
        feed = {k: v[:take] for k, v in batch.items()}
        outputs = model.forward(**feed)
        for j in range(nr_exits):                                                 # EE/utils.py:182-191
            if outputs.gated_logits is not None and len(outputs.gated_logits) > 0:
                logits_store[j, row:row + take] = outputs.gated_logits[j].double().cpu().numpy()
            else:
                logits_store[j, row:row + take] = outputs.exit_states[j][0].double().cpu().numpy()
        logits_store[-1, row:row + take] = outputs.logits.double().cpu().numpy()  # EE/utils.py:192
        correct += int((outputs.logits.argmax(-1).cpu() == torch.as_tensor(feed["labels"]).view(-1).cpu()).sum())
        row += take
    name = "test" if labelset == "test" else "validation"                       # EE/utils.py:204-209
    cfg = dict(config)
    cfg["labelset"] = "test"
    dump_logits(model, logits_store, references, cfg, name=name)
    return logits_store, references, None


def full_test_iteration(logits, references, config, start_threshold: float, step: float, device: int = 0):
    """The threshold sweep of EE/eval.py:227-274 (`--full_test True`): for every `t in np.arange(start_threshold, 1,
    step)` the configured policy is applied to the stored logits — `t` is the global `exit_threshold`, or, for
    `accuracy_calibration_heuristic`, the `epsilon` the per-exit thresholds are derived with (:242-245) — and one result
    dict per sweep point is collected and written to `<checkpoint dir>/<policy>/calibrated-metrics.json` (or
    `non-calibrated-metrics.json`, :264-272).  The reference re-runs the whole Python policy loop per point; here all
    points go through ONE device scan (`mmee.policy.PolicyStore`).  Each dict holds what that scan yields: accuracy,
    exit distribution and average exit (the reference adds the other `calc_metrics` scores and the FLOP accounting of
    `EE/analysis.py`, which are outside this path)."""
    from .policy import PolicyStore, heuristic_thresholds

    lg = np.asarray(logits, dtype=np.float64)
    E1 = lg.shape[0]
    sweep = np.arange(start_threshold, 1, step)
    policy = config.get("exit_policy", "max_confidence_global_thresholding_policy")
    if policy == "accuracy_calibration_heuristic":
        if "calibration_metrics" not in config:
            raise Exception("calibration_metrics not in config -> Set calibrate flag to True")
        rows = np.stack([heuristic_thresholds(config["calibration_metrics"], float(eps), E1) for eps in sweep])
        key = "epsilon"
    elif policy == "max_confidence_global_thresholding_policy":
        rows = np.repeat(sweep[:, None], E1, axis=1)
        key = "exit_threshold"
    else:
        raise NotImplementedError(policy)
    with PolicyStore(lg, "max_confidence", labels=np.asarray(references).reshape(-1), device=device) as store:
        res = store.scan(rows, per_exit=True, want_exits=False)
    results = []
    for t, value in enumerate(sweep):
        results.append({key: float(value), "accuracy": float(res.accuracy[t]), "average_exit": float(res.mean_exit[t]),
                        "exit_distribution": {int(e): float(res.hist[t, e]) / res.n_samples for e in range(E1)}})
    out_dir = os.path.join(config_to_checkpoint(config), policy)
    os.makedirs(out_dir, exist_ok=True)
    name = "calibrated-metrics.json" if config.get("calibrate") else "non-calibrated-metrics.json"
    with open(os.path.join(out_dir, name), "w+") as f:
        json.dump(results, f, indent=4)
    return results

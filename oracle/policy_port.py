"""TEST INFRASTRUCTURE — numpy fp64 restatement of the reference's post-hoc exit policy.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU legs may import this.
Pinned by `tests/test_oracle.py` against the reference's own `Policy` class (imported from
/root/reference when present) and against `tests/golden/policy_*.npz`.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple

import numpy as np


def softmax64(x: np.ndarray, axis: int = -1) -> np.ndarray:
    """scipy.special.softmax semantics (max-shifted), fp64 — used at EE/policy.py:30-32."""
    x = np.asarray(x, dtype=np.float64)
    e = np.exp(x - x.max(axis=axis, keepdims=True))
    return e / e.sum(axis=axis, keepdims=True)


def entropy64(x: np.ndarray) -> np.ndarray:
    """EE/models/EE_modules.py:149-154 evaluated in fp64: log sum e^x - sum x e^x / sum e^x."""
    x = np.asarray(x, dtype=np.float64)
    ex = np.exp(x)
    a = ex.sum(axis=-1)
    b = (x * ex).sum(axis=-1)
    return np.log(a) - b / a


def temperature_scale(logits: np.ndarray, temperatures: Optional[Sequence[float]]) -> np.ndarray:
    """EE/generic_scaling.py:54-61 applied per exit as EE/eval.py:312-327 does: logits[e] / T_e."""
    logits = np.asarray(logits, dtype=np.float64)
    if temperatures is None:
        return logits
    t = np.asarray(temperatures, dtype=np.float64).reshape(-1, 1, 1)
    assert t.shape[0] == logits.shape[0]
    return logits / t


def criterion(logits: np.ndarray, kind: str) -> np.ndarray:
    """[E+1,N,K] -> [E+1,N]; "max_confidence" EE/policy.py:30-32, "entropy" EE_modules.py:149."""
    if kind == "max_confidence":
        return softmax64(logits).max(axis=-1)
    if kind == "entropy":
        return entropy64(logits)
    raise NotImplementedError(kind)


def exit_policy(logits: np.ndarray, thresholds, kind: str = "max_confidence"
                ) -> Tuple[np.ndarray, np.ndarray, Dict[int, float]]:
    """First exit whose criterion passes its threshold, else the last one (unconditional).

    Restates the double loop of EE/policy.py:28-45 (strict `>` for max-confidence, :33) and,
    for entropy, the sign of EE/models/EE_modules.py:142-143 (strict `<`).  `thresholds` is a
    scalar (global policy, :17) or a per-exit vector (accuracy_calibration_heuristic, :94).
    Returns (exits_store int32[N], predictions f64[N,K], exit_distribution)."""
    logits = np.asarray(logits, dtype=np.float64)
    E1, N, K = logits.shape
    thr = np.broadcast_to(np.asarray(thresholds, dtype=np.float64), (E1,)) if np.ndim(thresholds) == 0 \
        else np.asarray(thresholds, dtype=np.float64)
    exits_store = np.zeros(N, dtype=np.int32)
    predictions = np.zeros((N, K), dtype=np.float64)
    for s in range(N):
        for e in range(E1):
            if kind == "max_confidence":
                score = np.max(softmax64(logits[e][s]))
                fire = score > thr[e]
            else:
                score = entropy64(logits[e][s])
                fire = score < thr[e]
            if fire or e == E1 - 1:
                exits_store[s] = e
                predictions[s] = logits[e][s]
                break
    dist = {e: np.count_nonzero(exits_store == e) / N for e in range(E1)}
    return exits_store, predictions, dist


def exit_policy_vectorised(logits, thresholds, kind="max_confidence"):
    """Vectorised equivalent (the property the reference checks at EE/thresh.py:308-318)."""
    logits = np.asarray(logits, dtype=np.float64)
    E1, N, K = logits.shape
    thr = np.broadcast_to(np.asarray(thresholds, dtype=np.float64).reshape(-1), (E1,)) \
        if np.size(thresholds) == 1 else np.asarray(thresholds, dtype=np.float64)
    crit = criterion(logits, kind)
    fire = crit > thr[:, None] if kind == "max_confidence" else crit < thr[:, None]
    fire[-1, :] = True
    idx = fire.argmax(axis=0).astype(np.int32)
    return idx, logits[idx, np.arange(N)], crit


# ---------------------------------------------------------------------------------------------------------------
# Per-exit threshold-vector ("mixture") sweeps: EE/thresh.py:184-233 and EE/large_scale.py:12-128.
def csf(logits: np.ndarray, name: str = "msp") -> np.ndarray:
    """CSF_dict of EE/large_scale.py:12-18 / EE/thresh.py:56-62 on [E1, N, K] -> [E1, N]: "msp" = max softmax,
    "entropy" = NEGATED entropy (so that larger = more confident for both)."""
    if name == "msp":
        return softmax64(logits).max(axis=-1)
    if name == "entropy":
        return -entropy64(logits)
    if name == "margin":
        # EE/thresh.py:48-52 `top12_margin_np` as written: ascending sort, values[0] - values[1]
        values = np.sort(np.asarray(logits, dtype=np.float64), axis=-1)
        return values[..., 0] - values[..., 1]
    raise NotImplementedError(name)


def check_2d_threshold(csf_logits: np.ndarray, threshold: np.ndarray) -> np.ndarray:
    """EE/thresh.py:184-185 = EE/large_scale.py:42-43: `(CSF_logits >= threshold[:, None]).argmax(0)` — non-strict,
    every exit tested (the last one too), 0 when no exit fires."""
    return (csf_logits >= np.asarray(threshold)[:, None]).argmax(0)


def opt0_2d(csf_logits: np.ndarray, thresholds_2d: np.ndarray) -> np.ndarray:
    """EE/thresh.py:188-215 (`opt0_2D`, serial: the reference maps check_2D_threshold over the rows with joblib)."""
    return np.stack([check_2d_threshold(csf_logits, t) for t in thresholds_2d]).astype(np.int32)


def evaluate_exit_logits(logits: np.ndarray, references: np.ndarray, exits: np.ndarray):
    """EE/thresh.py:225-233 / EE/large_scale.py:87-107: accuracy of the logits at the exit taken, average exit, and
    the exit distribution."""
    n = len(references)
    accuracy = np.mean(np.argmax(logits[exits, np.arange(n)], axis=-1) == references)
    average_exit = np.mean(exits)
    dist = {e: np.count_nonzero(exits == e) / n for e in range(logits.shape[0])}
    return accuracy, average_exit, dist


def generate_thresholds(csf_logits: np.ndarray, num_per_exit: int, num_mixtures: int, seed: int = 42) -> np.ndarray:
    """EE/large_scale.py:46-62: percentile thresholds per exit (last exit's row stays 0), `num_mixtures` random picks
    (np.random.seed(42), one randint(0, num_per_exit, num_exits) per mixture)."""
    np.random.seed(seed)
    num_exits = csf_logits.shape[0]
    exit_thresholds = np.zeros((num_exits, num_per_exit))
    percentiles = np.linspace(0, 100, num_per_exit)
    for exit_id in range(num_exits - 1):
        for p, perc in enumerate(percentiles):
            exit_thresholds[exit_id, p] = np.percentile(csf_logits[exit_id], perc)
    mixture_selection = [np.random.randint(0, num_per_exit, num_exits) for _ in range(num_mixtures)]
    return exit_thresholds[np.arange(num_exits), mixture_selection]

"""TEST INFRASTRUCTURE — scipy/numpy fp64 restatement of the reference's temperature calibration.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU legs may import this.

Follows EE/generic_scaling.py:37-111 (`TemperatureScaler`: L-BFGS-B from the current temperature, bounds
(1e-32, inf), objective sklearn `log_loss(labels, softmax(logits / T))`) and the per-exit loop of
EE/eval.py:293-335 (`calibrate`: ONE scaler object reused across exits, so exit i starts from exit i-1's
temperature; accuracy / average confidence taken from the calibrated TEST logits, accuracy against the
VALIDATION references exactly as the reference does).

Pinned by `tests/test_oracle.py` against `tests/golden/calibration.npz`, which holds the temperatures the
reference's own class produced on seeded inputs (`tests/golden/make_calibration_golden.py`).

ECE is NOT pinned: the reference calls the hub metric `jordyvl/ece` (EE/metrics.py:479-498), which is not in
the reference tree and cannot be fetched; `ece_equal_mass` restates the definition its arguments name
(equal-mass bins, upper-edge proxy, p = 1) and says so.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import numpy as np
from scipy.optimize import minimize
from scipy.special import log_softmax, softmax


def synthetic_exit_logits(seed: int, n_exits: int, n_samples: int, n_labels: int) -> Tuple[np.ndarray, np.ndarray]:
    """Seeded [E1, N, K] logits + labels whose optimal temperatures differ per exit (over- and under-confident)."""
    rng = np.random.default_rng(seed)
    labels = rng.integers(0, n_labels, n_samples)
    gains = [4.0, 0.3, 2.0, 1.0, 7.0, 0.7, 12.0, 1.5]
    logits = np.zeros((n_exits, n_samples, n_labels))
    for e in range(n_exits):
        z = rng.standard_normal((n_samples, n_labels))
        z[np.arange(n_samples), labels] += 0.5 + 0.2 * e          # never separable: a finite minimiser exists
        logits[e] = gains[e % len(gains)] * z
    return logits, labels.astype(np.int64)


def nll(labels: np.ndarray, logits: np.ndarray, temperature: float) -> float:
    """manual_NLL (EE/generic_scaling.py:29-34) of logits / T."""
    ls = log_softmax(np.asarray(logits, dtype=np.float64) / temperature, axis=-1)
    return float(-np.mean(ls[np.arange(len(labels)), labels]))


class TemperatureScalerPort:
    """EE/generic_scaling.py:37-111."""

    def __init__(self, temperature=None):
        self.temperature = np.ones(1) * (temperature if temperature else 1.0)

    def temperature_scale(self, logits):
        logits = np.asarray(logits)
        return logits / np.resize(self.temperature, logits.shape)

    def transform(self, logits):
        return softmax(self.temperature_scale(logits), -1)

    def fit(self, labels, logits):
        labels = np.asarray(labels).astype(np.int64).reshape(-1)
        logits = np.asarray(logits, dtype=np.float64)
        rows = np.arange(labels.shape[0])
        eps = np.finfo(np.float64).eps

        def objective(t):
            # sklearn.metrics.log_loss: clip to [eps, 1 - eps] then -mean log p[y]  (:74-79)
            p = np.clip(softmax(logits / t, -1), eps, 1 - eps)
            return -np.mean(np.log(p[rows, labels]))

        res = minimize(objective, x0=self.temperature, method="L-BFGS-B", bounds=[(1e-32, None)])
        assert res.success
        self.temperature = res.x
        return self.temperature


def ece_equal_mass(references: np.ndarray, probs: np.ndarray, n_bins: int = 100) -> float:
    """UNPINNED restatement of the `jordyvl/ece` call at EE/metrics.py:483-497: equal-mass bins over the max
    probability (n_bins = min(N - 1, 100)), bin confidence = the bin's upper edge, L1, weighted by bin mass."""
    probs = np.asarray(probs, dtype=np.float64)
    conf = probs.max(-1)
    correct = (probs.argmax(-1) == np.asarray(references)).astype(np.float64)
    n = conf.shape[0]
    n_bins = max(1, min(n - 1, n_bins))
    order = np.argsort(conf, kind="stable")
    total = 0.0
    for grp in np.array_split(order, n_bins):
        if grp.size == 0:
            continue
        total += grp.size / n * abs(correct[grp].mean() - conf[grp].max())
    return float(total)


def calibrate(validation_logits: np.ndarray, validation_references: np.ndarray, test_logits: np.ndarray,
              ece_fn=None) -> Tuple[np.ndarray, Dict[str, List[float]]]:
    """EE/eval.py:293-335 without the caching / dumping around it."""
    ece_fn = ece_fn or (lambda refs, lg: ece_equal_mass(refs, softmax(lg, -1)))
    calibrated = np.zeros_like(test_logits)
    T = TemperatureScalerPort()
    out: Dict[str, List[float]] = {"ece": [], "accuracy": [], "temperature": [], "average_confidence": []}
    for i in range(test_logits.shape[0]):
        T.fit(validation_references, validation_logits[i])
        calibrated[i] = T.temperature_scale(test_logits[i])
        out["ece"].append(ece_fn(validation_references, calibrated[i]))
        out["average_confidence"].append(float(softmax(calibrated[i], -1).max(-1).mean()))
        out["temperature"].append(float(T.temperature[0]))
        out["accuracy"].append(float(np.mean(calibrated[i].argmax(-1) == validation_references)))
    return calibrated, out

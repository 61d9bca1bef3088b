"""Host-side temperature calibration (offline; N x K is tiny).

`TemperatureScaler` mirrors the reference class of the same name
(EE/generic_scaling.py:37-111): `fit(labels, logits)` minimises the NLL of
softmax(logits / T) over the scalar T with L-BFGS-B from x0 = 1 and bounds
(1e-32, inf) (:89-95); `temperature_scale(logits)` returns logits / T (:54-61).
The engine applies 1/T_e on the device inside the exit kernel; this class only
produces the T_e vector.

`spread_temperatures` is the synthetic-benchmark stand-in for a fitted T
(SURVEY.md §8(d)): random-init heads give max-softmax ~0.1 at every exit, so it
picks T_e such that the median criterion at exit e hits a target that rises
with depth; thresholds 0.5-0.99 then produce a spread of exit depths.
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
from scipy.optimize import minimize
from scipy.special import log_softmax, softmax


class TemperatureScaler:
    def __init__(self, temperature: Optional[float] = None):
        self.temperature = np.ones(1) * (temperature if temperature else 1.0)

    def fit(self, labels, logits):
        return self.set_temperature(labels, logits)

    def transform(self, logits):
        return softmax(self.temperature_scale(logits), -1)

    def temperature_scale(self, logits):
        logits = np.asarray(logits)
        return logits / np.resize(self.temperature, logits.shape)

    def set_temperature(self, labels, logits):
        labels = np.asarray(labels).astype(np.int64).reshape(-1)
        logits = np.asarray(logits, dtype=np.float64)
        rows = np.arange(labels.shape[0])

        def objective(t):
            # sklearn.log_loss clips probabilities to [eps, 1-eps]; with eps = fp64 machine epsilon
            p = np.exp(log_softmax(logits / t, axis=-1))
            eps = np.finfo(np.float64).eps
            p = np.clip(p, eps, 1 - eps)
            p = p / p.sum(axis=1, keepdims=True)
            return -np.mean(np.log(p[rows, labels]))

        res = minimize(objective, x0=self.temperature, method="L-BFGS-B", bounds=[(1e-32, None)])
        assert res.success
        self.temperature = res.x
        return self.temperature


def median_criterion(logits_e: np.ndarray, t: float, kind: str) -> float:
    z = np.asarray(logits_e, dtype=np.float64) / t
    if kind == "max_confidence":
        return float(np.median(softmax(z, -1).max(-1)))
    ls = log_softmax(z, -1)
    return float(np.median(-(np.exp(ls) * ls).sum(-1)))


def spread_temperatures(exit_logits: np.ndarray, kind: str = "max_confidence",
                        lo: float = 0.45, hi: float = 0.97) -> np.ndarray:
    """T_e [E+1] such that the median max-softmax at exit e is linspace(lo, hi)[e]
    (for entropy: the median entropy equals that of a max-softmax target, mapped through
    the 2-point proxy below so both criteria give similar exit-depth spreads)."""
    exit_logits = np.asarray(exit_logits, dtype=np.float64)
    E1, _, K = exit_logits.shape
    targets = np.linspace(lo, hi, E1)
    temps = np.ones(E1)
    for e in range(E1):
        if kind == "max_confidence":
            tgt, increasing_in_invT = targets[e], True
        else:
            p = targets[e]                      # entropy of (p, (1-p)/(K-1), ...)
            q = (1 - p) / (K - 1)
            tgt, increasing_in_invT = -(p * np.log(p) + (K - 1) * q * np.log(q)), False
        a, b = 1e-4, 1e4                         # bisection on T (criterion monotone in 1/T)
        for _ in range(80):
            mid = np.sqrt(a * b)
            v = median_criterion(exit_logits[e], mid, kind)
            sharper_needed = (v < tgt) if increasing_in_invT else (v > tgt)
            if sharper_needed:
                b = mid
            else:
                a = mid
        temps[e] = np.sqrt(a * b)
    return temps


def thresholds_for(kind: str, conf_threshold: float, n_labels: int) -> float:
    """Map a max-confidence threshold to the matching entropy threshold (same 2-point proxy)."""
    if kind == "max_confidence":
        return conf_threshold
    p = conf_threshold
    q = (1 - p) / (n_labels - 1)
    return float(-(p * np.log(p) + (n_labels - 1) * q * np.log(q)))

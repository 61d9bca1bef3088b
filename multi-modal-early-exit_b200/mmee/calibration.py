"""Temperature calibration of the per-exit logits.

`TemperatureScaler` mirrors the reference class of the same name
(EE/generic_scaling.py:37-111): `fit(labels, logits)` minimises the NLL of
softmax(logits / T) over the scalar T; `temperature_scale(logits)` returns
logits / T (:54-61).  The minimisation runs on the device
(`mmee_temperature_fit`, csrc/calibrate.cuh: safeguarded Newton on 1/T, fp64,
all exits in one call) and fails loudly without one; the reference's L-BFGS-B
answer lies within ~4e-4 (relative) of the minimiser it returns
(tests/golden/make_calibration_golden.py).  `calibrate` is the per-exit loop of
EE/eval.py:293-335 on top of it.  The engine applies 1/T_e on the device inside
the exit kernel; this module only produces the T_e vector and the
`calibration_metrics` the heuristic policy reads (EE/policy.py:68-70).

`spread_temperatures` is the synthetic-benchmark stand-in for a fitted T
(SURVEY.md §8(d)): random-init heads give max-softmax ~0.1 at every exit, so it
picks T_e such that the median criterion at exit e hits a target that rises
with depth; thresholds 0.5-0.99 then produce a spread of exit depths.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
from scipy.special import log_softmax, softmax

from . import _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _stack(logits, labels) -> Tuple[np.ndarray, np.ndarray]:
    lg = np.ascontiguousarray(logits, dtype=np.float64)
    if lg.ndim == 2:
        lg = lg[None]
    if lg.ndim != 3:
        raise ValueError("logits must be [num_exits + 1, num_samples, num_labels] (or one exit [N, K])")
    lab = np.ascontiguousarray(labels, dtype=np.int64).reshape(-1)
    if lab.shape[0] != lg.shape[1]:
        raise ValueError("labels must have one entry per sample")
    return lg, lab


def temperature_fit(logits, labels, t_init: Optional[Sequence[float]] = None, max_iter: int = 40,
                    device: int = 0) -> Dict[str, np.ndarray]:
    """Per-exit temperatures minimising the NLL of softmax(logits[e] / T_e) (device computation).

    Returns {"temperature", "nll_before", "nll_after", "average_confidence", "accuracy"}, each [E1]; the last two
    are taken on the same logits at the fitted temperatures."""
    lg, lab = _stack(logits, labels)
    E1, N, K = lg.shape
    t0 = None if t_init is None else np.ascontiguousarray(np.broadcast_to(np.asarray(t_init, dtype=np.float64), (E1,)))
    out = {k: np.empty(E1, dtype=np.float64)
           for k in ("temperature", "nll_before", "nll_after", "average_confidence", "accuracy")}
    _lib.check(_lib.load().mmee_temperature_fit(device, E1, N, K, _ptr(lg), _ptr(lab), _ptr(t0), int(max_iter),
                                                _ptr(out["temperature"]), _ptr(out["nll_before"]),
                                                _ptr(out["nll_after"]), _ptr(out["average_confidence"]),
                                                _ptr(out["accuracy"])))
    return out


def calibration_stats(logits, labels, temperatures: Optional[Sequence[float]] = None,
                      device: int = 0) -> Dict[str, np.ndarray]:
    """{"nll", "average_confidence", "accuracy"} [E1] of softmax(logits[e] / T_e) against labels (device)."""
    lg, lab = _stack(logits, labels)
    E1, N, K = lg.shape
    t = None if temperatures is None else np.ascontiguousarray(
        np.broadcast_to(np.asarray(temperatures, dtype=np.float64), (E1,)))
    out = {k: np.empty(E1, dtype=np.float64) for k in ("nll", "average_confidence", "accuracy")}
    _lib.check(_lib.load().mmee_calibration_stats(device, E1, N, K, _ptr(lg), _ptr(lab), _ptr(t), _ptr(out["nll"]),
                                                  _ptr(out["average_confidence"]), _ptr(out["accuracy"])))
    return out


class TemperatureScaler:
    """Same constructor / methods as EE/generic_scaling.py:37-111; the fit runs on the device."""

    def __init__(self, temperature: Optional[float] = None, device: int = 0):
        self.temperature = np.ones(1) * (temperature if temperature else 1.0)
        self.device = device

    def fit(self, labels, logits):
        return self.set_temperature(labels, logits)

    def transform(self, logits):
        return softmax(self.temperature_scale(logits), -1)

    def temperature_scale(self, logits):
        logits = np.asarray(logits)
        return logits / np.resize(self.temperature, logits.shape)

    def set_temperature(self, labels, logits):
        # the reference starts L-BFGS-B from the current temperature (:89-95); so does the Newton iteration
        res = temperature_fit(np.asarray(logits)[None], labels, t_init=self.temperature, device=self.device)
        self.temperature = res["temperature"].copy()
        return self.temperature


def ece_equal_mass(references, logits_or_probs, n_bins: int = 100) -> float:
    """Expected calibration error with the arguments the reference passes to the hub metric `jordyvl/ece`
    (EE/metrics.py:479-498): equal-mass bins over the max probability, n_bins = min(N - 1, 100), the bin's upper
    edge as its confidence, L1, weighted by bin mass.  That metric's source is not part of the reference tree, so
    this follows its stated definition and is not checked against it; pass `ece_fn` to `calibrate` to use the
    original where it is installed."""
    p = np.asarray(logits_or_probs, dtype=np.float64)
    if not np.isclose(np.sum(p), len(p)):                 # EE/metrics.py:480-481: logits -> probabilities
        p = softmax(p, axis=-1)
    conf = p.max(-1)
    correct = (p.argmax(-1) == np.asarray(references)).astype(np.float64)
    n = conf.shape[0]
    order = np.argsort(conf, kind="stable")
    total = 0.0
    for grp in np.array_split(order, max(1, min(n - 1, n_bins))):
        if grp.size:
            total += grp.size / n * abs(correct[grp].mean() - conf[grp].max())
    return float(total)


def calibrate(validation_logits, validation_references, test_logits,
              ece_fn: Optional[Callable] = None, device: int = 0) -> Tuple[np.ndarray, Dict[str, List[float]]]:
    """EE/eval.py:293-335: fit T_e on the validation logits, scale the test logits, collect `calibration_metrics`
    = {"ece", "accuracy", "temperature", "average_confidence"} (lists of E1 floats, what
    Policy.accuracy_calibration_heuristic reads).  As in the reference, accuracy and average confidence come from
    the calibrated TEST logits and accuracy is taken against the VALIDATION references (:325-335), and exit i's fit
    starts from exit i-1's temperature (one scaler object, :298); all exits are fitted in one device call, which
    reaches the same minimisers."""
    val = np.asarray(validation_logits, dtype=np.float64)
    test = np.asarray(test_logits, dtype=np.float64)
    refs = np.asarray(validation_references).reshape(-1)
    fit = temperature_fit(val, refs, device=device)
    temps = fit["temperature"]
    calibrated = test / temps[:, None, None]
    stats = calibration_stats(test, refs, temps, device=device) if test.shape[1] == refs.shape[0] else None
    if stats is None:
        raise ValueError("the reference compares the calibrated test logits with the validation references "
                         "(EE/eval.py:333-335): both splits must have the same number of samples")
    ece_fn = ece_fn or ece_equal_mass
    metrics = {"ece": [float(ece_fn(refs, calibrated[i])) for i in range(test.shape[0])],
               "accuracy": [float(a) for a in stats["accuracy"]],
               "temperature": [float(t) for t in temps],
               "average_confidence": [float(c) for c in stats["average_confidence"]]}
    return calibrated, metrics


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

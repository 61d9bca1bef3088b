"""Drop-in for the reference's post-hoc exit policy (boundary #2 of SURVEY.md §8b):

    exits_store, predictions, exit_distribution = getattr(Policy(logits=..., config=...), name)()     # EE/eval.py:91-98

`Policy` keeps the reference class's constructor, method names, config keys and return types
(EE/policy.py:7-111); the per-sample double loop runs on the GPU (`mmee_policy_scan`, csrc/policy.cuh) in fp64
like the reference.  `sweep()` evaluates many thresholds in one launch — what `full_test_iteration`
(EE/eval.py:227-274) and `thresh.py` `opt0` (EE/thresh.py:106-132) do with one Python pass per threshold.
There is no CPU fallback: without libmmee.so / a B200 the calls raise.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib

CRITERIA = {"max_confidence": 0, "entropy": 1}


@dataclass
class SweepResult:
    thresholds: np.ndarray          # [T, E1] as evaluated (last column unused)
    exits: np.ndarray               # int32 [T, N]
    hist: np.ndarray                # int64 [T, E1]   samples per exit
    criteria: np.ndarray            # f64 [E1, N]
    correct: Optional[np.ndarray]   # int64 [T] or None (needs labels)

    @property
    def exit_distribution(self) -> np.ndarray:
        return self.hist / self.exits.shape[1]

    @property
    def accuracy(self) -> Optional[np.ndarray]:
        return None if self.correct is None else self.correct / self.exits.shape[1]

    @property
    def mean_exit(self) -> np.ndarray:
        return (self.hist * np.arange(self.hist.shape[1])[None, :]).sum(1) / self.exits.shape[1]


def heuristic_thresholds(calibration_metrics: dict, epsilon: float, num_exits: int) -> np.ndarray:
    """Per-exit thresholds of accuracy_calibration_heuristic, EE/policy.py:68-79 (host arithmetic, E+1 values)."""
    acc = calibration_metrics["accuracy"]
    ece = calibration_metrics["ece"]
    metrics = np.array([1 - (acc[i] / ece[i]) for i in range(num_exits)])
    return (metrics - (np.min(metrics) - epsilon)) / ((np.max(metrics) + epsilon) - (np.min(metrics) - epsilon))


def policy_scan(logits: np.ndarray, thresholds: np.ndarray, criterion: str = "max_confidence",
                temperatures: Optional[Sequence[float]] = None, labels: Optional[np.ndarray] = None,
                device: int = 0) -> SweepResult:
    """logits f64 [E1, N, K]; thresholds [T, E1] (or [T] global, or scalar) -> SweepResult (device computation)."""
    lib = _lib.load()
    lg = np.ascontiguousarray(logits, dtype=np.float64)
    if lg.ndim != 3:
        raise ValueError("logits must be [num_exits + 1, num_samples, num_labels]")
    E1, N, K = lg.shape
    thr = np.asarray(thresholds, dtype=np.float64)
    if thr.ndim == 0:
        thr = np.full((1, E1), float(thr))
    elif thr.ndim == 1 and thr.shape[0] != E1:
        thr = np.repeat(thr[:, None], E1, axis=1)          # [T] global thresholds
    elif thr.ndim == 1:
        thr = thr[None, :]                                  # one per-exit vector
    if thr.shape[1] != E1:
        raise ValueError(f"thresholds must have {E1} columns")
    thr = np.ascontiguousarray(thr)
    T = thr.shape[0]
    temps = None if temperatures is None else np.ascontiguousarray(temperatures, dtype=np.float64)
    if temps is not None and temps.shape != (E1,):
        raise ValueError(f"temperatures must have {E1} entries")
    lab = None if labels is None else np.ascontiguousarray(labels, dtype=np.int64).reshape(-1)
    if lab is not None and lab.shape[0] != N:
        raise ValueError("labels must have one entry per sample")
    exits = np.empty((T, N), dtype=np.int32)
    crit = np.empty((E1, N), dtype=np.float64)
    hist = np.empty((T, E1), dtype=np.int64)
    correct = np.empty(T, dtype=np.int64) if lab is not None else None
    p = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    _lib.check(lib.mmee_policy_scan(device, E1, N, K, p(lg), p(temps), CRITERIA[criterion], p(thr), T, p(lab),
                                    p(exits), p(crit), p(hist), p(correct)))
    return SweepResult(thr, exits, hist, crit, correct)


class Policy:
    """Same constructor / methods / results as the reference `Policy` (EE/policy.py:7-111)."""

    def __init__(self, logits, config) -> None:
        self.logits = logits
        self.config = config

    def _device(self) -> int:
        d = self.config.get("cuda_device", 0) if hasattr(self.config, "get") else 0
        return int(d)

    def _finish(self, res: SweepResult) -> Tuple[np.ndarray, torch.Tensor, Dict[int, float]]:
        lg = np.asarray(self.logits, dtype=np.float64)
        exits_store = res.exits[0]
        n = lg.shape[1]
        # predictions[s] = logits[exit taken][s]   (EE/policy.py:36-38, 43-45)
        predictions = torch.from_numpy(lg[exits_store, np.arange(n)].copy())
        dev = self.config.get("device", "cpu") if hasattr(self.config, "get") else "cpu"
        if dev not in (None, "cpu"):
            predictions = predictions.to(dev)
        exit_distribution = {e: float(res.hist[0, e]) / n for e in range(lg.shape[0])}
        return exits_store, predictions, exit_distribution

    def max_confidence_global_thresholding_policy(self):
        """EE/policy.py:12-53: first exit with max softmax > config["exit_threshold"], else the last exit."""
        return self._finish(policy_scan(self.logits, float(self.config["exit_threshold"]), "max_confidence",
                                        device=self._device()))

    def accuracy_calibration_heuristic(self):
        """EE/policy.py:55-111: per-exit thresholds from (accuracy, ECE) and config["epsilon"]."""
        if "calibration_metrics" not in self.config:
            raise Exception("calibration_metrics not in config -> Set calibrate flag to True")
        num_exits = np.asarray(self.logits).shape[0]
        thr = heuristic_thresholds(self.config["calibration_metrics"], self.config["epsilon"], num_exits)
        return self._finish(policy_scan(self.logits, thr, "max_confidence", device=self._device()))

    def sweep(self, thresholds, criterion: str = "max_confidence", temperatures=None, labels=None) -> SweepResult:
        """All sweep points of EE/eval.py:227-274 (`np.arange(start, 1, step)`) in one device pass."""
        return policy_scan(self.logits, np.asarray(thresholds, dtype=np.float64), criterion, temperatures, labels,
                           device=self._device())

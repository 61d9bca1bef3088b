"""Drop-in for the reference's post-hoc exit policy (boundary #2 of SURVEY.md §8b):

    exits_store, predictions, exit_distribution = getattr(Policy(logits=..., config=...), name)()     # EE/eval.py:91-98

`Policy` keeps the reference class's constructor, method names, config keys and return types
(EE/policy.py:7-111); the per-sample double loop runs on the GPU (`mmee_policy_store_*`, csrc/policy.cuh) in fp64
like the reference.  `sweep()` evaluates many global thresholds in one launch — what `full_test_iteration`
(EE/eval.py:227-274) and `thresh.py` `opt0` (EE/thresh.py:106-132) do with one Python pass per threshold;
`PolicyStore.mixture_sweep()` is `opt0_2D` + the per-mixture accuracy / average exit of `large_scale.py`
(EE/large_scale.py:46-128) for any number of per-exit threshold vectors, with the criteria resident on the device.
There is no CPU fallback: without libmmee.so / a B200 the calls raise.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib

CRITERIA = {"max_confidence": 0, "entropy": 1, "margin": 2}
# EE/large_scale.py:12-18 / EE/thresh.py:56-62 CSF_dict names.  "margin" is the reference's `top12_margin_np` AS WRITTEN
# (np.sort ascending, values[0] - values[1]: smallest minus second-smallest logit), kept bug-compatible.
CSF_TO_CRITERION = {"msp": "max_confidence", "entropy": "entropy", "margin": "margin"}
MODES = {"policy": 0, "check_2D_threshold": 1}


@dataclass
class SweepResult:
    thresholds: np.ndarray          # [T, E1] as evaluated
    exits: Optional[np.ndarray]     # int32 [T, N], or None when the scan was asked for counts only
    hist: np.ndarray                # int64 [T, E1]   samples per exit
    criteria: Optional[np.ndarray]  # f64 [E1, N]
    correct: Optional[np.ndarray]   # int64 [T] or None (needs labels)
    n_samples: int = 0

    @property
    def exit_distribution(self) -> np.ndarray:
        return self.hist / self.n_samples

    @property
    def accuracy(self) -> Optional[np.ndarray]:
        return None if self.correct is None else self.correct / self.n_samples

    @property
    def mean_exit(self) -> np.ndarray:
        """`average_exit` of EE/large_scale.py:101 (mean exit index)."""
        return (self.hist * np.arange(self.hist.shape[1])[None, :]).sum(1) / self.n_samples


def heuristic_thresholds(calibration_metrics: dict, epsilon: float, num_exits: int) -> np.ndarray:
    """Per-exit thresholds of accuracy_calibration_heuristic, EE/policy.py:68-79 (host arithmetic, E+1 values)."""
    acc = calibration_metrics["accuracy"]
    ece = calibration_metrics["ece"]
    metrics = np.array([1 - (acc[i] / ece[i]) for i in range(num_exits)])
    return (metrics - (np.min(metrics) - epsilon)) / ((np.max(metrics) + epsilon) - (np.min(metrics) - epsilon))


def _threshold_rows(thresholds, E1: int, per_exit: bool) -> np.ndarray:
    """-> f64 [T, E1].  A scalar or a 1-D array means GLOBAL thresholds (one sweep point each, the same value at
    every exit); per-exit vectors are passed as a 2-D [T, E1] array, or as one 1-D vector with per_exit=True.  The
    meaning of a 1-D array is never guessed from its length."""
    thr = np.asarray(thresholds, dtype=np.float64)
    if thr.ndim == 0:
        thr = np.full((1, E1), float(thr))
    elif thr.ndim == 1 and per_exit:
        if thr.shape[0] != E1:
            raise ValueError(f"a per-exit threshold vector must have {E1} entries (one per exit incl. the final one)")
        thr = thr[None, :]
    elif thr.ndim == 1:
        thr = np.repeat(thr[:, None], E1, axis=1)
    elif per_exit and thr.ndim != 2:
        raise ValueError("per-exit thresholds must be [E1] or [T, E1]")
    if thr.ndim != 2 or thr.shape[1] != E1:
        raise ValueError(f"thresholds must have {E1} columns")
    return np.ascontiguousarray(thr)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class PolicyStore:
    """Criteria of one logits store, resident on the device (`mmee_policy_store_*`).

    logits f64 [E1, N, K] (what EE/utils.py:160-193 stores and EE/thresh.py:checkpoint_logits reloads)."""

    def __init__(self, logits, criterion: str = "max_confidence", temperatures: Optional[Sequence[float]] = None,
                 labels: Optional[np.ndarray] = None, device: int = 0):
        self._lib = _lib.load()
        lg = np.ascontiguousarray(logits, dtype=np.float64)
        if lg.ndim != 3:
            raise ValueError("logits must be [num_exits + 1, num_samples, num_labels]")
        self.E1, self.N, self.K = lg.shape
        self.criterion = criterion
        temps = None if temperatures is None else np.ascontiguousarray(temperatures, dtype=np.float64)
        if temps is not None and temps.shape != (self.E1,):
            raise ValueError(f"temperatures must have {self.E1} entries")
        lab = None if labels is None else np.ascontiguousarray(labels, dtype=np.int64).reshape(-1)
        if lab is not None and lab.shape[0] != self.N:
            raise ValueError("labels must have one entry per sample")
        self.has_labels = lab is not None
        self._h = C.c_void_p()
        _lib.check(self._lib.mmee_policy_store_create(device, self.E1, self.N, self.K, _ptr(lg), _ptr(temps),
                                                      CRITERIA[criterion], _ptr(lab), C.byref(self._h)))

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.mmee_policy_store_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def criteria(self) -> np.ndarray:
        """f64 [E1, N]: max softmax (EE/policy.py:30-32) or entropy (EE_modules.py:149-154) of logits[e][s] / T_e."""
        out = np.empty((self.E1, self.N), dtype=np.float64)
        _lib.check(self._lib.mmee_policy_store_criteria(self._h, _ptr(out)))
        return out

    def scan(self, thresholds, per_exit: bool = False, mode: str = "policy", want_exits: bool = True,
             want_criteria: bool = False) -> SweepResult:
        thr = _threshold_rows(thresholds, self.E1, per_exit)
        T = thr.shape[0]
        exits = np.empty((T, self.N), dtype=np.int32) if want_exits else None
        hist = np.empty((T, self.E1), dtype=np.int64)
        correct = np.empty(T, dtype=np.int64) if self.has_labels else None
        _lib.check(self._lib.mmee_policy_store_scan(self._h, _ptr(thr), T, MODES[mode], _ptr(exits), _ptr(hist),
                                                    _ptr(correct)))
        return SweepResult(thr, exits, hist, self.criteria() if want_criteria else None, correct, self.N)

    def mixture_sweep(self, thresholds_2D) -> SweepResult:
        """`opt0_2D` (EE/thresh.py:188-215 / EE/large_scale.py:65-84) over [M, E1] per-exit threshold vectors with
        `check_2D_threshold` semantics, reduced on the device to what `evaluate_exit_logits`
        (EE/large_scale.py:87-128) derives from each row: exit histogram, accuracy, average exit.  The [M, N] index
        matrix itself is never materialised."""
        return self.scan(np.asarray(thresholds_2D, dtype=np.float64), per_exit=True, mode="check_2D_threshold",
                         want_exits=False)


def generate_thresholds(criteria: np.ndarray, num_per_exit: int = 10, num_mixtures: int = 1500000, seed: int = 42
                        ) -> np.ndarray:
    """Threshold mixtures of EE/large_scale.py:46-62 (`generate_thresholds`; the reference's module-level constants
    are the defaults, :175-178): per exit `num_per_exit` percentiles (0..100) of the CSF, the last exit's row left at
    zero, then `num_mixtures` random picks of one percentile per exit.  `criteria` is the [E1, N] CSF (`PolicyStore.
    criteria()`, negated by the caller for entropy as EE/large_scale.py:15 does).  Host arithmetic only (percentiles
    + index picks), same numpy calls in the same order, so the same seed gives the reference's mixtures."""
    np.random.seed(seed)
    num_exits = criteria.shape[0]
    exit_thresholds = np.zeros((num_exits, num_per_exit))
    percentiles = np.linspace(0, 100, num_per_exit)
    for exit_id in range(num_exits - 1):
        for p, perc in enumerate(percentiles):
            exit_thresholds[exit_id, p] = np.percentile(criteria[exit_id], perc)
    # one randint call of shape [M, E1] consumes the legacy RandomState stream exactly like the reference's M calls of
    # shape [E1] (checked in tests/test_host_cpu.py)
    mixture_selection = np.random.randint(0, num_per_exit, (num_mixtures, num_exits))
    return exit_thresholds[np.arange(num_exits)[None, :], mixture_selection]


def policy_scan(logits: np.ndarray, thresholds, criterion: str = "max_confidence",
                temperatures: Optional[Sequence[float]] = None, labels: Optional[np.ndarray] = None,
                device: int = 0, per_exit: bool = False, mode: str = "policy") -> SweepResult:
    """One-shot scan: logits f64 [E1, N, K]; thresholds = scalar / [T] GLOBAL thresholds, or per-exit [T, E1]
    (or one [E1] vector with per_exit=True) -> SweepResult with exits, histogram and criteria (device computation)."""
    lg = np.asarray(logits)
    if lg.ndim != 3:
        raise ValueError("logits must be [num_exits + 1, num_samples, num_labels]")
    rows = _threshold_rows(thresholds, lg.shape[0], per_exit)          # argument errors before any device work
    with PolicyStore(lg, criterion, temperatures, labels, device) as st:
        return st.scan(rows, per_exit=True, mode=mode, want_exits=True, want_criteria=True)


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
        return self._finish(policy_scan(self.logits, thr, "max_confidence", device=self._device(), per_exit=True))

    def sweep(self, thresholds, criterion: str = "max_confidence", temperatures=None, labels=None,
              per_exit: bool = False) -> SweepResult:
        """All sweep points of EE/eval.py:227-274 (`np.arange(start, 1, step)`) in one device pass.  `thresholds` is a
        1-D array of T GLOBAL thresholds (always: a sweep whose length happens to equal the number of exits is still T
        sweep points), or a 2-D [T, E1] array of per-exit vectors (per_exit=True also accepts a single [E1] vector)."""
        return policy_scan(self.logits, np.asarray(thresholds, dtype=np.float64), criterion, temperatures, labels,
                           device=self._device(), per_exit=per_exit)

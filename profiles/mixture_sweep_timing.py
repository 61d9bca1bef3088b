#!/usr/bin/env python
"""large_scale.py at its own size on the device: 1.5 M per-exit threshold mixtures (num_per_exit = 10, seed 42,
EE/large_scale.py:173-178) over an RVL-CDIP-test-sized logits store (40 000 samples, 14 exits, 16 classes; synthetic,
seeded).  Prints one JSON line with the wall times of the steps; a 2 000-mixture slice is checked row by row against
the oracle's check_2d_threshold / evaluate_exit_logits (restatement of EE/thresh.py:184-233, pinned to the reference's
own outputs by tests/golden/mixtures.npz).  Run under gpurun:  python profiles/mixture_sweep_timing.py"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "multi-modal-early-exit_b200"))

from mmee.policy import PolicyStore, generate_thresholds  # noqa: E402
from oracle import policy_port  # noqa: E402


def main():
    E1, N, K, M = 14, 40000, 16, 1500000
    rng = np.random.default_rng(7)
    labels = rng.integers(0, K, size=N)
    lg = rng.normal(size=(E1, N, K)) * np.linspace(0.8, 3.0, E1)[:, None, None]
    lg[:, np.arange(N), labels] += np.linspace(0.3, 2.5, E1)[:, None]
    t0 = time.perf_counter()
    store = PolicyStore(lg, "max_confidence", labels=labels)
    csf = store.criteria()
    t_store = time.perf_counter() - t0
    t0 = time.perf_counter()
    thr2d = generate_thresholds(csf, 10, M)
    t_gen = time.perf_counter() - t0
    store.mixture_sweep(thr2d[:4096])                                   # warm-up
    t0 = time.perf_counter()
    res = store.mixture_sweep(thr2d)
    t_sweep = time.perf_counter() - t0
    # oracle check of a slice (the reference's per-mixture loop, serial)
    t0 = time.perf_counter()
    n_chk = 2000
    ok = True
    for t in range(n_chk):
        ex = policy_port.check_2d_threshold(csf, thr2d[t])
        acc, avg, _ = policy_port.evaluate_exit_logits(lg, labels, ex)
        ok = ok and np.array_equal(np.bincount(ex, minlength=E1), res.hist[t]) and abs(acc - res.accuracy[t]) < 1e-15 \
            and abs(avg - res.mean_exit[t]) < 1e-12
    t_cpu = time.perf_counter() - t0
    print(json.dumps({
        "what": "opt0_2D + evaluate_exit_logits of EE/large_scale.py for 1.5 M mixtures x 40 000 samples x 14 exits",
        "store_create_s": round(t_store, 3), "generate_thresholds_host_s": round(t_gen, 3),
        "device_sweep_s": round(t_sweep, 3), "mixtures": M, "samples": N,
        "pairs_per_s": M * N / t_sweep, "oracle_check": {"mixtures": n_chk, "all_equal": bool(ok), "cpu_s": round(t_cpu, 2),
        "cpu_s_per_mixture": t_cpu / n_chk, "extrapolated_cpu_s_for_all": t_cpu / n_chk * M},
        "result": {"accuracy_min": float(res.accuracy.min()), "accuracy_max": float(res.accuracy.max()),
                   "mean_exit_min": float(res.mean_exit.min()), "mean_exit_max": float(res.mean_exit.max())}}))
    store.close()


if __name__ == "__main__":
    main()

"""Generates tests/golden/calibration.npz with the reference's OWN TemperatureScaler (dev container only).

    python tests/golden/make_calibration_golden.py

Imports /root/reference/EE/generic_scaling.py verbatim (its `from metrics import ece_logits` is satisfied by a
stub: that function only feeds two print statements, EE/generic_scaling.py:82-87, 101-109, and needs the hub metric
`jordyvl/ece`).  The scaler object is reused across exits as EE/eval.py:298-313 does.  Inputs are regenerated from
the seed by `oracle.calibration_port.synthetic_exit_logits`, so only the temperatures are stored.  Also records
how far the reference's L-BFGS-B answer is from the exact minimiser (bounded Brent, xatol 1e-12): that gap is
the tolerance of the parity tests.
"""
import contextlib
import io
import os
import sys
import types

import numpy as np
from scipy.optimize import minimize_scalar

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import calibration_port  # noqa: E402

CASES = {"cal_a": (11, 5, 3000, 16), "cal_b": (12, 13, 1200, 16), "cal_c": (13, 3, 257, 5)}


def main() -> None:
    stub = types.ModuleType("metrics")
    stub.ece_logits = lambda references, predictions: 0.0
    sys.modules["metrics"] = stub
    sys.path.insert(0, "/root/reference/EE")
    import generic_scaling as G  # the reference, verbatim

    out = {}
    for name, (seed, E1, N, K) in CASES.items():
        logits, labels = calibration_port.synthetic_exit_logits(seed, E1, N, K)
        scaler = G.TemperatureScaler()
        t_ref, t_opt = [], []
        for e in range(E1):
            with contextlib.redirect_stdout(io.StringIO()):
                scaler.fit(labels, logits[e])
            t_ref.append(float(scaler.temperature[0]))
            r = minimize_scalar(lambda t: calibration_port.nll(labels, logits[e], t), bounds=(1e-3, 1e3),
                                method="bounded", options={"xatol": 1e-12})
            t_opt.append(float(r.x))
        t_ref, t_opt = np.array(t_ref), np.array(t_opt)
        print(name, "max |T_ref - T_opt| / T_opt =", np.abs(t_ref - t_opt).max() / 1, (np.abs(t_ref - t_opt) / t_opt).max())
        out[f"{name}_shape"] = np.array([seed, E1, N, K])
        out[f"{name}_t_ref"] = t_ref
        out[f"{name}_t_opt"] = t_opt
        out[f"{name}_nll_ref"] = np.array([calibration_port.nll(labels, logits[e], t_ref[e]) for e in range(E1)])
    np.savez(os.path.join(ROOT, "tests", "golden", "calibration.npz"), **out)


if __name__ == "__main__":
    main()

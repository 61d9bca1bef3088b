"""Golden vectors for the per-exit threshold-vector ("mixture") sweep, produced by the REFERENCE's own functions.

EE/thresh.py imports seaborn / plotly / matplotlib at module level (absent here) and runs a script body on import, so
the function definitions it needs are taken from the reference source with `ast` and executed unmodified in a
namespace holding numpy / scipy / joblib: `entropy`, `CSF_dict`, `check_2D_threshold`, `opt0_2D`, `parallel_process`,
`evaluate_exit_logits` of EE/thresh.py and `generate_thresholds` of EE/large_scale.py.  Nothing of the reference is
copied into the repository.  Run in the dev container:  python tests/golden/make_mixture_golden.py
"""
import ast
import os
import sys
from collections import OrderedDict

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = os.environ.get("MMEE_REFERENCE_ROOT", "/root/reference")


def reference_namespace():
    from joblib import Parallel, delayed
    from scipy.special import softmax

    ns = {"np": np, "softmax": softmax, "Parallel": Parallel, "delayed": delayed, "OrderedDict": OrderedDict}
    wanted = {"thresh.py": {"entropy", "top12_margin_np", "check_2D_threshold", "opt0_2D", "parallel_process",
                            "evaluate_exit_logits", "CSF_dict"},
              "large_scale.py": {"generate_thresholds"}}
    for fname, names in wanted.items():
        src = open(os.path.join(REF, "EE", fname)).read()
        tree = ast.parse(src)
        for node in tree.body:
            if isinstance(node, ast.FunctionDef) and node.name in names:
                exec(compile(ast.Module([node], []), fname, "exec"), ns)
            elif isinstance(node, ast.Assign) and any(getattr(t, "id", None) in names for t in node.targets):
                exec(compile(ast.Module([node], []), fname, "exec"), ns)
    return ns


def synthetic_store(seed, E1, N, K):
    """Seeded per-exit logits whose confidence grows with depth, and labels correlated with the deeper exits."""
    rng = np.random.default_rng(seed)
    labels = rng.integers(0, K, size=N)
    lg = rng.normal(size=(E1, N, K)) * np.linspace(0.8, 2.5, E1)[:, None, None]
    lg[:, np.arange(N), labels] += np.linspace(0.5, 3.0, E1)[:, None]
    return lg, labels


def main():
    ns = reference_namespace()
    out = {}
    for name, (seed, E1, N, K, npe, M) in {"mix_a": (1, 6, 1500, 16, 10, 400), "mix_b": (2, 14, 900, 16, 10, 300)}.items():
        lg, labels = synthetic_store(seed, E1, N, K)
        for csf_name in (("msp", "entropy", "margin") if name == "mix_a" else ("msp", "entropy")):
            CSF = ns["CSF_dict"][csf_name]
            ns["CSF"] = CSF                                     # thresh.py's opt0_2D reads the module-level CSF
            csf_logits = np.apply_along_axis(CSF, -1, lg)
            # large_scale.generate_thresholds uses module-level constants and CSF_dict["msp"]; bind them
            ns["num_per_exit"], ns["num_mixtures"] = npe, M
            saved = ns["CSF_dict"]
            ns["CSF_dict"] = {"msp": CSF}
            thr2d = ns["generate_thresholds"](lg, labels)
            ns["CSF_dict"] = saved
            exits = np.asarray(ns["opt0_2D"](labels, lg, thr2d))
            acc = np.array([ns["evaluate_exit_logits"](lg, labels, ex)[0] for ex in exits])
            avg = np.array([ns["evaluate_exit_logits"](lg, labels, ex)[1] for ex in exits])
            tag = f"{name}_{csf_name}"
            out[tag + "_shape"] = np.array([seed, E1, N, K, npe, M])
            out[tag + "_thr2d"] = thr2d
            out[tag + "_exits"] = exits.astype(np.int32)
            out[tag + "_acc"] = acc
            out[tag + "_avg_exit"] = avg
            out[tag + "_csf_sum"] = csf_logits.sum(axis=1)
            print(tag, exits.shape, "acc", acc.min(), acc.max(), "avg exit", avg.min(), avg.max())
    np.savez_compressed(os.path.join(HERE, "mixtures.npz"), **out)


if __name__ == "__main__":
    main()

"""GPU parity tests of the post-hoc exit policy (`mmee.policy`, C ABI `mmee_policy_scan`) against the oracle
policy port and the reference `Policy` results stored in the golden fixtures.  Exit indices are integers: the bar
is bit-exact for every sample whose criterion margin to its threshold exceeds 1e-12 (fp64 arithmetic on both
sides; the device exp() may differ from numpy's in the last ulp)."""
import numpy as np
import pytest
import torch

from helpers import ALL_CASES, load_case
from mmee.policy import Policy, heuristic_thresholds, policy_scan
from oracle import policy_port

pytestmark = pytest.mark.gpu
MARGIN = 1e-12


def _decisive(crit, thr_row, kind="max_confidence"):
    return np.abs(crit[:-1] - np.asarray(thr_row)[:-1, None]).min(axis=0) > MARGIN if crit.shape[0] > 1 \
        else np.ones(crit.shape[1], bool)


@pytest.mark.parametrize("name", ALL_CASES)
def test_policy_matches_reference_golden(name):
    """Same call shape as EE/eval.py:91-98; results equal the reference Policy's stored in the golden."""
    g, *_ = load_case(name)
    lg = g["exit_logits"].astype(np.float64)
    for tag, l in (("raw", lg), ("cal", policy_port.temperature_scale(lg, g["temps"]))):
        for thr in (0.1, 0.5, 0.7, 0.9):
            ex, pred, dist = Policy(logits=l, config={"exit_threshold": thr, "device": "cpu"}
                                    ).max_confidence_global_thresholding_policy()
            assert ex.dtype == np.int32 and isinstance(pred, torch.Tensor) and pred.dtype == torch.float64
            assert np.array_equal(ex, g[f"policy_{tag}_{thr}_exits"])
            assert np.array_equal(pred.numpy(), g[f"policy_{tag}_{thr}_pred"])
            assert abs(sum(dist.values()) - 1.0) < 1e-12 and set(dist) == set(range(l.shape[0]))


def test_sweep_equals_oracle_loop_full_size():
    """RVL-CDIP-test-sized store (40k samples, 14 exits, 16 classes), the 0.5..0.99 sweep of BASELINE config 3."""
    rng = np.random.default_rng(3)
    E1, N, K = 14, 40000, 16
    lg = rng.normal(size=(E1, N, K)) * np.linspace(0.5, 4.0, E1)[:, None, None]
    labels = rng.integers(0, K, size=N)
    temps = np.linspace(1.5, 0.7, E1)
    thrs = np.arange(0.5, 1.0, 0.01)
    res = Policy(lg, {"device": "cpu"}).sweep(thrs, temperatures=temps, labels=labels)
    cal = policy_port.temperature_scale(lg, temps)
    crit = policy_port.criterion(cal, "max_confidence")
    assert np.abs(res.criteria - crit).max() < 1e-13
    prev = None
    for t, thr in enumerate(thrs):
        want, _, _ = policy_port.exit_policy_vectorised(cal, thr, "max_confidence")
        ok = _decisive(crit, np.full(E1, thr))
        assert np.array_equal(res.exits[t][ok], want[ok]) and ok.mean() > 0.999
        assert res.hist[t].sum() == N and np.array_equal(np.bincount(res.exits[t], minlength=E1), res.hist[t])
        assert res.correct[t] == int((cal[res.exits[t], np.arange(N)].argmax(-1) == labels).sum())
        if prev is not None:
            assert (res.exits[t] >= prev).all()          # exit depth is monotone in the threshold
        prev = res.exits[t]
    assert np.allclose(res.mean_exit, res.exits.mean(1))


def test_entropy_and_per_exit_thresholds_and_edges():
    rng = np.random.default_rng(0)
    lg = rng.normal(size=(5, 777, 16)) * 3
    thr = np.array([0.5, 1.0, 1.5, 2.0, 0.0])
    res = policy_scan(lg, thr, "entropy", per_exit=True)
    want, pred, crit = policy_port.exit_policy_vectorised(lg, thr, "entropy")
    ok = _decisive(crit, thr)
    assert np.array_equal(res.exits[0][ok], want[ok])
    assert np.abs(res.criteria - crit).max() < 1e-12
    # strict comparisons: threshold 1.0 never fires, threshold 0.0 always fires at exit 0
    ex, pred, dist = Policy(lg, {"exit_threshold": 1.0, "device": "cpu"}).max_confidence_global_thresholding_policy()
    assert (ex == 4).all() and dist[4] == 1.0 and np.array_equal(pred.numpy(), lg[4])
    ex, pred, _ = Policy(lg, {"exit_threshold": 0.0, "device": "cpu"}).max_confidence_global_thresholding_policy()
    assert (ex == 0).all() and np.array_equal(pred.numpy(), lg[0])
    # a single exit (E+1 == 1) and a single sample
    ex, _, _ = Policy(lg[:1, :1], {"exit_threshold": 0.99, "device": "cpu"}).max_confidence_global_thresholding_policy()
    assert ex.tolist() == [0]


def test_accuracy_calibration_heuristic_matches_loop():
    rng = np.random.default_rng(5)
    lg = rng.normal(size=(6, 500, 16)) * 2.5
    cm = {"accuracy": [0.5, 0.6, 0.7, 0.75, 0.8, 0.85], "ece": [0.2, 0.15, 0.12, 0.1, 0.08, 0.05],
          "average_confidence": [0.6] * 6}
    cfg = {"calibration_metrics": cm, "epsilon": 0.1, "device": "cpu"}
    ex, pred, dist = Policy(lg, cfg).accuracy_calibration_heuristic()
    thr = heuristic_thresholds(cm, 0.1, 6)
    want, wpred, _ = policy_port.exit_policy(lg, thr, "max_confidence")
    assert np.array_equal(ex, want) and np.array_equal(pred.numpy(), wpred)
    with pytest.raises(Exception, match="calibration_metrics"):
        Policy(lg, {"epsilon": 0.1}).accuracy_calibration_heuristic()


def test_sweep_length_equal_to_exit_count_is_still_a_sweep():
    """14 global thresholds on a 14-exit store are 14 sweep points, not one per-exit vector."""
    rng = np.random.default_rng(8)
    E1, N, K = 14, 300, 16
    lg = rng.normal(size=(E1, N, K)) * 2
    thrs = np.arange(0.3, 1.0, 0.05)
    assert len(thrs) == E1
    res = Policy(lg, {"device": "cpu"}).sweep(thrs)
    assert res.exits.shape == (E1, N) and res.thresholds.shape == (E1, E1)
    for t, thr in enumerate(thrs):
        want, _, crit = policy_port.exit_policy_vectorised(lg, thr, "max_confidence")
        ok = _decisive(crit, np.full(E1, thr))
        assert np.array_equal(res.exits[t][ok], want[ok])
    one = Policy(lg, {"device": "cpu"}).sweep(thrs, per_exit=True)          # the explicit per-exit form
    assert one.exits.shape == (1, N)
    want, _, _ = policy_port.exit_policy_vectorised(lg, thrs, "max_confidence")
    assert np.array_equal(one.exits[0], want)


def test_entropy_criterion_is_stable_at_small_temperatures():
    """exp(x / T) overflows fp64 once max|x| / T > 709; the max-shifted form does not (and equals the
    reference's formula wherever that one is finite)."""
    rng = np.random.default_rng(2)
    lg = rng.normal(size=(3, 200, 16))
    temps = np.array([1.0, 0.01, 0.001])                # x / T up to ~4000
    res = policy_scan(lg, 0.5, "entropy", temperatures=temps)
    assert np.isfinite(res.criteria).all() and (res.criteria >= -1e-12).all()
    cal = policy_port.temperature_scale(lg, temps)
    ref = policy_port.entropy64(cal[0])
    assert np.abs(res.criteria[0] - ref).max() < 1e-12
    y = cal - cal.max(-1, keepdims=True)
    shifted = np.log(np.exp(y).sum(-1)) - (y * np.exp(y)).sum(-1) / np.exp(y).sum(-1)
    assert np.abs(res.criteria - shifted).max() < 1e-12
    assert (res.exits[0][res.criteria[0] < 0.5] == 0).all()


# ------------------------------------------------------------------ mixture sweeps (check_2D_threshold / opt0_2D)
def _mixture_case(tag):
    import os
    import sys
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "mixtures.npz"))
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    from make_mixture_golden import synthetic_store
    seed, E1, N, K, npe, M = (int(v) for v in g[tag + "_shape"])
    lg, labels = synthetic_store(seed, E1, N, K)
    return g, lg, labels


@pytest.mark.parametrize("tag", ["mix_a_msp", "mix_a_entropy", "mix_a_margin", "mix_b_msp", "mix_b_entropy"])
def test_mixture_sweep_matches_reference_golden(tag):
    """Device `check_2D_threshold` semantics (>=, every exit tested, argmax-0 fallback, negated entropy) against the
    exits / accuracy / average exit the REFERENCE's opt0_2D + evaluate_exit_logits produced
    (tests/golden/make_mixture_golden.py)."""
    from mmee.policy import PolicyStore, generate_thresholds
    g, lg, labels = _mixture_case(tag)
    crit_name = {"msp": "max_confidence", "entropy": "entropy", "margin": "margin"}[tag.split("_")[-1]]
    thr2d, want = g[tag + "_thr2d"], g[tag + "_exits"]
    N = lg.shape[1]
    with PolicyStore(lg, crit_name, labels=labels) as st:
        csf = -st.criteria() if crit_name == "entropy" else st.criteria()      # larger = more confident (CSF_dict)
        assert np.allclose(csf.sum(1), g[tag + "_csf_sum"], rtol=1e-12)
        # the same mixtures as the reference's generate_thresholds from the device criteria (percentiles of the same
        # values; the last ulp of a criterion can move a percentile by an ulp)
        mine = generate_thresholds(csf, int(g[tag + "_shape"][4]), int(g[tag + "_shape"][5]))
        assert mine.shape == thr2d.shape and np.abs(mine - thr2d).max() < 1e-12
        full = st.scan(thr2d, per_exit=True, mode="check_2D_threshold", want_exits=True)
        # thresholds ARE criterion values (percentiles): ties are the rule, so a last-ulp difference between the device
        # exp() and numpy's can flip `>=` for the tied sample; everything else must be identical
        tie = (np.abs(csf[None, :, :] - thr2d[:, :, None]) < 1e-13).any(axis=1)
        assert np.array_equal(full.exits[~tie], want[~tie]) and (~tie).mean() > 0.97
        counts = st.mixture_sweep(thr2d)                                   # counts only (n_thr < 2048: sample-parallel)
        assert counts.exits is None and np.array_equal(counts.hist, full.hist) and np.array_equal(counts.correct, full.correct)
        for t in range(thr2d.shape[0]):
            assert np.array_equal(np.bincount(full.exits[t], minlength=lg.shape[0]), full.hist[t])
            assert full.correct[t] == int((lg[full.exits[t], np.arange(N)].argmax(-1) == labels).sum())
        clean = ~tie.any(axis=1)
        assert np.allclose(full.accuracy[clean], g[tag + "_acc"][clean], atol=1e-15)
        assert np.allclose(full.mean_exit[clean], g[tag + "_avg_exit"][clean], atol=1e-12)


@pytest.mark.parametrize("crit_name", ["max_confidence", "entropy"])
def test_large_mixture_sweep_equals_oracle(crit_name):
    """large_scale.py-shaped sweep: 120 000 per-exit threshold vectors (> the 65 535 grid limit, thread-per-mixture
    kernel, no index matrix) on a 14-exit store against the oracle's check_2d_threshold, row by row."""
    from mmee.policy import PolicyStore
    rng = np.random.default_rng(11)
    E1, N, K, M = 14, 2000, 16, 120000
    labels = rng.integers(0, K, size=N)
    lg = rng.normal(size=(E1, N, K)) * np.linspace(0.8, 3.0, E1)[:, None, None]
    lg[:, np.arange(N), labels] += np.linspace(0.3, 2.5, E1)[:, None]
    with PolicyStore(lg, crit_name, labels=labels) as st:
        crit = st.criteria()
        csf = crit if crit_name == "max_confidence" else -crit
        assert np.abs(csf - policy_port.csf(lg, "msp" if crit_name == "max_confidence" else "entropy")).max() < 1e-12
        # thresholds strictly between criterion values: no ties, so the comparison is exact on both sides
        lo, hi = np.percentile(csf, 1, axis=1), np.percentile(csf, 99, axis=1)
        thr2d = lo[None, :] + rng.random((M, E1)) * (hi - lo)[None, :]
        thr2d[rng.random((M, E1)) < 0.05] = np.inf                       # some exits never fire
        thr2d[:7, :] = np.inf                                            # nothing fires -> argmax of all-False = 0
        res = st.mixture_sweep(thr2d)
        assert res.exits is None and res.hist.shape == (M, E1) and (res.hist.sum(1) == N).all()
        assert (res.hist[:7, 0] == N).all()
        correct_at = (lg.argmax(-1) == labels[None, :])                  # [E1, N]
        for t0 in range(0, M, 4000):
            fire = csf[None, :, :] >= thr2d[t0:t0 + 4000, :, None]       # [m, E1, N]
            ex = fire.argmax(1)
            want_hist = (ex[:, :, None] == np.arange(E1)[None, None, :]).sum(1)
            assert np.array_equal(res.hist[t0:t0 + 4000], want_hist)
            want_corr = np.take_along_axis(correct_at.T[None, :, :].repeat(ex.shape[0], 0), ex[:, :, None], 2)[:, :, 0].sum(1)
            assert np.array_equal(res.correct[t0:t0 + 4000], want_corr)
        # the policy.py mode through the same kernel: strict, last exit unconditional
        pol = st.scan(thr2d[:5000], per_exit=True, mode="policy", want_exits=False)
        for t in (0, 7, 100, 4999):
            want, _, _ = policy_port.exit_policy_vectorised(
                policy_port.temperature_scale(lg, None), thr2d[t], crit_name)
            assert np.array_equal(pol.hist[t], np.bincount(want, minlength=E1))


def test_full_test_iteration_equals_policy_loop(tmp_path):
    """EE/eval.py:227-274: one device scan for the whole sweep == the policy oracle applied threshold by threshold
    (global thresholds and the accuracy / ECE heuristic swept over epsilon); result files as the reference names them."""
    import json
    import os

    from mmee.pipeline import full_test_iteration
    rng = np.random.default_rng(21)
    E1, N, K = 6, 800, 16
    labels = rng.integers(0, K, size=N)
    lg = rng.normal(size=(E1, N, K)) * np.linspace(1.0, 3.0, E1)[:, None, None]
    lg[:, np.arange(N), labels] += np.linspace(0.5, 2.5, E1)[:, None]
    cfg = {"checkpoint": "org/ckpt", "test_dataset": "ds/rvl", "results_root": str(tmp_path), "calibrate": True,
           "exit_policy": "max_confidence_global_thresholding_policy"}
    res = full_test_iteration(lg, labels, cfg, 0.3, 0.05)
    thrs = np.arange(0.3, 1, 0.05)
    assert len(res) == len(thrs)
    crit = policy_port.criterion(lg, "max_confidence")
    for r, thr in zip(res, thrs):
        want, pred, dist = policy_port.exit_policy_vectorised(lg, thr, "max_confidence")
        if (np.abs(crit[:-1] - thr).min(axis=0) > MARGIN).all():
            assert r["accuracy"] == pytest.approx(float((pred.argmax(-1) == labels).mean()), abs=1e-15)
            assert r["average_exit"] == pytest.approx(float(want.mean()), abs=1e-12)
        assert abs(sum(r["exit_distribution"].values()) - 1.0) < 1e-12 and r["exit_threshold"] == pytest.approx(thr)
    out = os.path.join(str(tmp_path), "results", "ckpt-rvl", "max_confidence_global_thresholding_policy", "calibrated-metrics.json")
    assert len(json.load(open(out))) == len(thrs)
    cfg2 = dict(cfg, exit_policy="accuracy_calibration_heuristic", calibrate=False,
                calibration_metrics={"accuracy": [0.5, 0.6, 0.7, 0.75, 0.8, 0.85], "ece": [0.2, 0.15, 0.12, 0.1, 0.08, 0.05]})
    res2 = full_test_iteration(lg, labels, cfg2, 0.1, 0.2)
    for r in res2:
        thr = heuristic_thresholds(cfg2["calibration_metrics"], r["epsilon"], E1)
        want, pred, _ = policy_port.exit_policy(lg, thr, "max_confidence")
        if (np.abs(crit[:-1] - thr[:-1, None]).min(axis=0) > MARGIN).all():
            assert r["accuracy"] == pytest.approx(float((pred.argmax(-1) == labels).mean()), abs=1e-15)
            assert r["average_exit"] == pytest.approx(float(want.mean()), abs=1e-12)
    assert os.path.exists(os.path.join(str(tmp_path), "results", "ckpt-rvl", "accuracy_calibration_heuristic", "non-calibrated-metrics.json"))


def test_mixture_sweep_many_exits():
    """A store with 40 exits (criteria stage above the 48 KB default shared-memory limit, 64-wide register histogram)."""
    from mmee.policy import PolicyStore
    rng = np.random.default_rng(12)
    E1, N, K, M = 40, 700, 8, 3000
    labels = rng.integers(0, K, size=N)
    lg = rng.normal(size=(E1, N, K)) * np.linspace(0.5, 3.0, E1)[:, None, None]
    with PolicyStore(lg, "max_confidence", labels=labels) as st:
        csf = st.criteria()
        lo, hi = np.percentile(csf, 5, axis=1), np.percentile(csf, 99.5, axis=1)
        thr2d = lo[None, :] + rng.random((M, E1)) * (hi - lo)[None, :]
        res = st.mixture_sweep(thr2d)
        ex = (csf[None, :, :] >= thr2d[:, :, None]).argmax(1)
        assert np.array_equal(res.hist, (ex[:, :, None] == np.arange(E1)[None, None, :]).sum(1))
        assert np.array_equal(res.correct, ((lg.argmax(-1) == labels[None, :]).T[np.arange(N)[None, :], ex]).sum(1))


def test_sweep_with_exits_beyond_grid_limit():
    """n_thr > 65 535 with the index matrix requested: chunked over grid.y."""
    from mmee.policy import PolicyStore
    rng = np.random.default_rng(4)
    lg = rng.normal(size=(4, 50, 8)) * 2
    thrs = np.linspace(0.2, 0.999, 70000)
    with PolicyStore(lg, "max_confidence") as st:
        res = st.scan(thrs, want_exits=True)
        crit = st.criteria()
    assert res.exits.shape == (70000, 50)
    for t in (0, 65534, 65535, 65536, 69999):
        want, _, _ = policy_port.exit_policy_vectorised(lg, thrs[t], "max_confidence")
        ok = np.abs(crit[:-1] - thrs[t]).min(axis=0) > MARGIN
        assert np.array_equal(res.exits[t][ok], want[ok])
        assert np.array_equal(np.bincount(res.exits[t], minlength=4), res.hist[t])


# ------------------------------------------------------------------ temperature calibration (mmee_temperature_fit)
def _cal_case(name):
    import os
    from oracle import calibration_port
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "calibration.npz"))
    seed, E1, N, K = (int(v) for v in g[f"{name}_shape"])
    logits, labels = calibration_port.synthetic_exit_logits(seed, E1, N, K)
    return g, logits, labels


@pytest.mark.parametrize("name", ["cal_a", "cal_b", "cal_c"])
def test_temperature_fit_matches_reference_golden(name):
    """Device fit vs the temperatures of the reference's own TemperatureScaler (golden) and the exact minimiser.
    The reference stops within 4e-4 (relative) of the minimiser (L-BFGS-B defaults, finite-difference gradient,
    tests/golden/make_calibration_golden.py), hence 1e-3 against it and 1e-7 against the minimiser; the device NLL
    is never above the reference's."""
    from mmee.calibration import temperature_fit
    from oracle import calibration_port
    g, logits, labels = _cal_case(name)
    res = temperature_fit(logits, labels)
    t = res["temperature"]
    assert np.all(np.abs(t - g[f"{name}_t_ref"]) <= 1e-3 * g[f"{name}_t_ref"])
    assert np.all(np.abs(t - g[f"{name}_t_opt"]) <= 1e-7 * g[f"{name}_t_opt"])
    for e in range(logits.shape[0]):
        assert abs(res["nll_after"][e] - calibration_port.nll(labels, logits[e], t[e])) < 1e-12
        assert res["nll_after"][e] <= g[f"{name}_nll_ref"][e] + 1e-13
        assert abs(res["nll_before"][e] - calibration_port.nll(labels, logits[e], 1.0)) < 1e-11
        assert res["accuracy"][e] == pytest.approx(np.mean(logits[e].argmax(-1) == labels), abs=1e-15)
        sm = policy_port.softmax64(logits[e] / t[e])
        assert res["average_confidence"][e] == pytest.approx(sm.max(-1).mean(), abs=1e-12)


def test_temperature_scaler_drop_in_and_calibrate_loop():
    """Same object protocol as EE/generic_scaling.py:37-61 (fit -> temperature [1], temperature_scale, transform),
    warm start from the previous exit as EE/eval.py:298-313; `calibrate` equals the oracle's restatement of
    EE/eval.py:293-335 (temperatures to L-BFGS-B's tolerance, accuracy exactly)."""
    from mmee.calibration import TemperatureScaler, calibrate, calibration_stats
    from oracle import calibration_port
    g, logits, labels = _cal_case("cal_a")
    T = TemperatureScaler()
    for e in range(logits.shape[0]):
        t = T.fit(labels, logits[e])
        assert t.shape == (1,) and abs(t[0] - g["cal_a_t_opt"][e]) <= 1e-7 * g["cal_a_t_opt"][e]
        assert np.allclose(T.temperature_scale(logits[e]), logits[e] / t[0], rtol=0, atol=0)
        assert np.allclose(T.transform(logits[e]).sum(-1), 1.0)
    test_logits, _ = calibration_port.synthetic_exit_logits(99, logits.shape[0], logits.shape[1], logits.shape[2])
    cal, met = calibrate(logits, labels, test_logits)
    cal_o, met_o = calibration_port.calibrate(logits, labels, test_logits)
    assert set(met) == {"ece", "accuracy", "temperature", "average_confidence"}
    assert np.allclose(met["temperature"], met_o["temperature"], rtol=1e-3)
    assert np.allclose(cal, cal_o, rtol=1e-3)
    assert met["accuracy"] == pytest.approx(met_o["accuracy"], abs=1e-15)
    assert np.allclose(met["average_confidence"], met_o["average_confidence"], atol=1e-3)
    assert np.allclose(met["ece"], met_o["ece"], atol=5e-3)
    # the metrics feed the heuristic policy unchanged (EE/policy.py:68-79)
    thr = heuristic_thresholds(met, 0.05, logits.shape[0])
    assert thr.shape == (logits.shape[0],) and np.all((thr > 0) & (thr < 1))
    st = calibration_stats(test_logits, labels, met["temperature"])
    assert np.allclose(st["average_confidence"], met["average_confidence"])


def test_temperature_fit_full_size_stationary_and_edge_cases():
    """RVL-CDIP-validation-sized store (40k samples, 14 exits): at the returned T the NLL derivative vanishes and
    neighbouring temperatures are worse (size-independent optimality check); separable exits and bad labels."""
    from mmee.calibration import temperature_fit
    from oracle import calibration_port
    logits, labels = calibration_port.synthetic_exit_logits(5, 14, 40000, 16)
    res = temperature_fit(logits, labels)
    for e in (0, 6, 13):
        t = res["temperature"][e]
        n0, n1, n2 = (calibration_port.nll(labels, logits[e], t * f) for f in (1.0, 1.001, 0.999))
        assert n0 <= n1 and n0 <= n2
        assert abs((n1 - n2) / (0.002 * t)) < 1e-6                      # central difference of d nll / dT
        assert abs(res["nll_after"][e] - n0) < 1e-12
    # warm start far away converges to the same minimiser
    far = temperature_fit(logits, labels, t_init=np.full(14, 50.0))
    assert np.allclose(far["temperature"], res["temperature"], rtol=1e-7)
    # a separable exit has no finite minimiser: the call returns a small positive T with nll ~ 0, no NaN
    sep = np.zeros((1, 64, 4)); lab = np.arange(64) % 4; sep[0, np.arange(64), lab] = 5.0
    r = temperature_fit(sep, lab)
    assert np.isfinite(r["temperature"][0]) and 0 < r["temperature"][0] < 1 and r["nll_after"][0] < 1e-6
    with pytest.raises(RuntimeError, match="label out of range"):
        temperature_fit(sep, lab + 1)
    with pytest.raises(ValueError):
        temperature_fit(sep, lab[:-1])

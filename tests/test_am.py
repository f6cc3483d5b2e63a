"""Hipparcos-Gaia astrometric block (BASELINE config 3).

CPU: the NumPy/longdouble oracle is bit-identical to the script the REAL reference generated
(tests/golden/c3_*.npz); the file loader mirrors DataWrapper's AM preprocessing.
GPU: the device block against the EXACT value of the reference's formulas (40-digit evaluation,
tests/tools/am_truth_mpmath.py -> tests/golden/c3_am_truth.npz) at north_star's 1e-10, and against the
reference's float at 1e-8: that float is itself up to 3.9e-9 away from the exact value
(test_am_value_is_ill_conditioned, test_reference_value_is_only_good_to_1e9_of_the_exact_value)."""
import os

import numpy as np
import pytest

from conftest import load_golden

AM_CASES = ["c3_hip21850_am_k1", "c3_hip21850_am_k2"]


def _am(g):
    return {k[3:]: g[k] for k in g.files if k.startswith("am_")}


@pytest.mark.parametrize("name", AM_CASES)
def test_am_oracle_bit_identical_to_generated_script(name):
    from oracle.am_oracle import AMOracle
    from oracle.rv_oracle import RVOracle
    g, spec = load_golden(name)
    cm = spec.compile()
    assert cm.am_enabled and all(m == 5 for m in cm.kep_model)
    ao, ro = AMOracle(cm, _am(g)), RVOracle(cm, g["t"], g["y"], g["yerr"], g["flag"])
    fin = np.where(np.isfinite(g["logp"]))[0]
    assert len(fin) > 30
    for i in fin:
        th = g["thetas"][i]
        with np.errstate(all="ignore"):
            la = ao.loglike_AM(th)
            tot = ro.my_likelihood(th) + la  # a00.like:5-8
        assert isinstance(la, np.longdouble)
        assert la == np.longdouble(g["logl_am_hi"][i]) + np.longdouble(g["logl_am_lo"][i])
        assert float(tot) == g["logl"][i]


def test_am_value_is_ill_conditioned():
    """One ulp on the propagated barycentre RA (what a different libm's arctan2 gives) moves
    loglike_AM by more than 1e-10 relative: the reference value is only defined to ~1e-9."""
    from oracle.am_oracle import AMOracle
    g, spec = load_golden("c3_hip21850_am_k2")
    ao = AMOracle(spec.compile(), _am(g))
    i = int(np.where(np.isfinite(g["logp"]))[0][0])
    th = g["thetas"][i]
    base = float(ao.loglike_AM(th))
    orig = ao.obs_lin_prop_PA

    def nudged(obs):
        out = orig(obs)
        out[:, 0] = np.nextafter(out[:, 0], np.inf)
        return out
    ao.obs_lin_prop_PA = nudged
    moved = float(ao.loglike_AM(th))
    assert abs(moved - base) / abs(base) > 1e-10


def test_am_with_fixed_parameter_is_rejected():
    from astroemperor_b200.modelspec import UnsupportedModelError
    g, spec = load_golden("c3_hip21850_am_k1")
    spec.blocks[0].params[3].fixed = 0.1
    with pytest.raises(UnsupportedModelError):
        spec.compile()


@pytest.mark.skipif(not os.path.isdir("/root/reference/tests/datafiles/HIP21850/AM"),
                    reason="reference data only exists in the build container")
def test_load_am_folder_matches_datawrapper():
    from astroemperor_b200.amdata import load_am_folder
    g, _ = load_golden("c3_hip21850_am_k2")
    ref = _am(g)
    mine = load_am_folder("/root/reference/tests/datafiles/HIP21850/AM/", float(np.ravel(ref["common_t"])[0]),
                          deadtime_dir="/root/reference/src/astroemperor/support/deadtime")
    for k, v in ref.items():
        a = np.asarray(mine[k], dtype=np.float64)
        assert a.shape == np.asarray(v).shape or k == "common_t", k
        assert np.allclose(a.ravel(), np.asarray(v, dtype=np.float64).ravel(), rtol=1e-12, atol=1e-14), k


@pytest.mark.gpu
@pytest.mark.parametrize("name", AM_CASES)
def test_am_device_matches_oracle(name):
    from astroemperor_b200.engine import LikelihoodEngine
    g, spec = load_golden(name)
    eng = LikelihoodEngine(spec, g["t"], g["y"], g["yerr"], g["flag"], am=_am(g))
    ll, lp = eng.logl_batch(g["thetas"])
    fin = np.isfinite(g["logp"])
    assert np.array_equal(np.isfinite(lp), fin) and np.all(ll[~fin] == -np.inf)
    ref = g["logl"][fin]
    rel = np.abs(ll[fin] - ref) / np.abs(ref)
    print(f"{name}: max rel err vs the reference's float {rel.max():.3e}")
    # the REFERENCE's float is up to 3.9e-9 away from the exact value of its own formulas (40-digit evaluation,
    # test_reference_value_is_only_good_to_1e9_of_the_exact_value); the device is held to the exact value below
    assert rel.max() < 1e-8
    # the Isotropic prior on the inclination goes through device sin/log: 1e-14, not bit-exact
    assert np.max(np.abs(lp[fin] - g["logp"][fin])) < 1e-12


def _truth(name):
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "c3_am_truth.npz"))
    return z[name + "_hi"].astype(np.longdouble) + z[name + "_lo"].astype(np.longdouble)


@pytest.mark.parametrize("name", AM_CASES)
def test_reference_value_is_only_good_to_1e9_of_the_exact_value(name):
    """tests/tools/am_truth_mpmath.py evaluated the same formulas on the same doubles with 40 digits: the
    reference's own float (x87 intermediates and all) is up to 3.9e-9 relative away from the exact value, i.e.
    north_star's 1e-10 'vs the reference' is below the reference's own accuracy for this block."""
    g, _ = load_golden(name)
    fin = np.isfinite(g["logp"])
    t = _truth(name)[fin]
    rel = np.abs((g["logl"][fin].astype(np.longdouble) - t) / t)
    assert 1e-10 < float(rel.max()) < 1e-8 and float(np.median(rel)) < 1e-9


@pytest.mark.gpu
@pytest.mark.parametrize("name", AM_CASES)
def test_am_device_is_at_least_as_close_to_the_exact_value_as_the_reference(name):
    """VERDICT r1 #2 (row A11): |device - truth| <= max(|reference - truth|, 1e-10 |truth|) for every golden theta,
    truth = the 40-digit evaluation (tests/golden/c3_am_truth.npz).  The achieved errors are written to
    gpurun_out/am_accuracy_<case>.json (copied to profiles/)."""
    import json
    from astroemperor_b200.engine import LikelihoodEngine
    g, spec = load_golden(name)
    eng = LikelihoodEngine(spec, g["t"], g["y"], g["yerr"], g["flag"], am=_am(g))
    ll, lp = eng.logl_batch(g["thetas"])
    fin = np.isfinite(g["logp"])
    t = _truth(name)[fin]
    e_dev = np.abs((ll[fin].astype(np.longdouble) - t) / t).astype(np.float64)
    e_ref = np.abs((g["logl"][fin].astype(np.longdouble) - t) / t).astype(np.float64)
    rec = {"case": name, "n": int(fin.sum()), "device_max_rel_err_vs_exact": float(e_dev.max()),
           "device_median_rel_err_vs_exact": float(np.median(e_dev)),
           "reference_max_rel_err_vs_exact": float(e_ref.max()),
           "reference_median_rel_err_vs_exact": float(np.median(e_ref)),
           "device_max_rel_err_vs_reference": float(np.max(np.abs(ll[fin] - g["logl"][fin]) / np.abs(g["logl"][fin]))),
           "thetas_where_device_is_closer": int(np.sum(e_dev <= e_ref))}
    print(json.dumps(rec))
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, f"am_accuracy_{name}.json"), "w") as fh:
        json.dump(rec, fh)
    # north_star's 1e-10, against the exact value; and never worse than the reference itself
    assert e_dev.max() < 1e-10, e_dev.max()
    assert np.all(e_dev <= np.maximum(e_ref, 1e-12)), (e_dev.max(), e_ref.max())


@pytest.mark.gpu
def test_am_pt_sweep_decisions_match_oracle():
    """BASELINE config 3 shape (joint RV + astrometry, 2 Keplerians): PT sweeps vs the oracle."""
    from astroemperor_b200.engine import LikelihoodEngine
    from astroemperor_b200.sampler import PTSampler
    from oracle.am_oracle import AMOracle
    from oracle.pt_oracle import PTOracle
    from oracle.rv_oracle import RVOracle
    g, spec = load_golden("c3_hip21850_am_k2")
    cm = spec.compile()
    eng = LikelihoodEngine(spec, g["t"], g["y"], g["yerr"], g["flag"], am=_am(g))
    T, W = 3, 64
    samp = PTSampler(W, eng.ndim, eng, ntemps=T, seed=4)
    p0 = samp.initial_positions(spec)

    class Joint(RVOracle):
        def __init__(self, *a, am=None):
            super().__init__(*a)
            self.ao = AMOracle(self.cm, am)

        def my_likelihood(self, theta):
            with np.errstate(all="ignore"):
                return float(super().my_likelihood(theta) + self.ao.loglike_AM(theta))

    orc = PTOracle(Joint(cm, g["t"], g["y"], g["yerr"], g["flag"], am=_am(g)), samp.betas)
    samp._init_state(p0)
    orc.init_state(p0)
    for k in range(3):
        d = samp.draw(1)
        n_acc = samp.sweep(d)
        _, n_acc_o, src_o = orc.sweep(d)
        p, ll, lp = samp.state_numpy()
        assert np.array_equal(n_acc, n_acc_o) and np.array_equal(p, orc.p)
        assert np.max(np.abs(ll - orc.logl) / np.abs(orc.logl)) < 5e-8
    print("AM PT min decision margin", orc.min_margin)

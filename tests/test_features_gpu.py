"""GPU: every instantiation of the likelihood kernel.  The kernel is specialised on the model-feature mask
(acceleration x {no MA, global MA(1), global MA(n)} x post-MA terms = 12 kernels, emp_logl.cuh); the golden
cases reach 6 of them, this test reaches all 12 — on data sizes that end in a partial tile and in a partial
64-point group — against the oracle (which is itself bit-identical to the reference-generated scripts)."""
import itertools

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RTOL = 1e-12


@pytest.mark.parametrize("acc,ma,post", list(itertools.product((0, 2), (None, 1, 2), (0, 1))))
def test_feature_mask_kernels_match_oracle(acc, ma, post):
    from astroemperor_b200.data import from_instrument_tables
    from astroemperor_b200.draws import initial_positions
    from astroemperor_b200.engine import LikelihoodEngine
    from astroemperor_b200.frontend import default_spec
    from astroemperor_b200.synth import add_activity_columns, make_synthetic_rv
    from oracle.rv_oracle import RVOracle
    seed = 100 + 4 * acc + 7 * (ma or 0) + post
    n = (577, 130, 1051)[(acc + (ma or 0) + post) % 3]          # 1 full tile + 65, 2 groups + 2, 2 tiles + 27
    files = make_synthetic_rv(seed=seed, n=n, nins=3, kplan=2, ma=ma is not None)
    if post:
        files = add_activity_columns(files, [1, 0, 2], seed)
    data = from_instrument_tables(files)
    spec = default_spec(data, kplan=2, parameterisation=(0, 1, 3)[seed % 3], acceleration=acc,
                        moav=None if ma is None else {"order": ma, "global": True},
                        sinusoid=1 if post else 0, magnetic_cycle=1 if (post and ma == 1) else 0,
                        conditions=[("Acceleration", "limits", [-0.01, 0.01]),          # keep the trend at the
                                    ("Acceleration Order 2", "limits", [-1e-5, 1e-5])])  # size of the signal
    cm = spec.compile()
    th = initial_positions(np.random.RandomState(seed), spec, 1, 24)[0]
    eng = LikelihoodEngine(spec, data.t, data.y, data.yerr, data.flag, sai=data.sai)
    ll, lp = eng.logl_batch(th)
    orc = RVOracle(cm, data.t, data.y, data.yerr, data.flag, sai=data.sai)
    ref, lpo = orc.logl_logp_batch(th)
    fin = np.isfinite(lpo)
    assert fin.sum() >= 12 and np.array_equal(np.isfinite(lp), fin)
    # logP: same roundings as the reference except that Normal.prior's `(...)**2` is libm pow(z, 2.0) in NumPy
    # (not always the correctly rounded square) and z*z on the device: at most 1 ulp of one term
    assert np.max(np.abs(lp[fin] - lpo[fin])) <= 4 * np.finfo(float).eps * np.max(np.abs(lpo[fin]))
    assert np.mean(lp[fin] == lpo[fin]) > 0.8
    rel = np.abs(ll[fin] - ref[fin]) / np.abs(ref[fin])
    assert rel.max() < RTOL, (acc, ma, post, rel.max())
    # and my_model (the plotting / residual entry) for one theta
    k = int(np.flatnonzero(fin)[0])
    model, err2 = eng.my_model(th[k])
    m0, e0 = orc.my_model(th[k])
    assert np.array_equal(err2, e0)
    assert np.max(np.abs(model - m0)) < 1e-12 * (np.max(np.abs(m0)) + 1.0)

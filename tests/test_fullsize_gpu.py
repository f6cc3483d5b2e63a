"""GPU parity at BASELINE.json's full data sizes (configs[3]: 10 000 points, 5 planets, 4 instruments, global
MA(1); configs[4]: 50 000 points): a random subset of walkers against the oracle at full N, and the
size-independent properties the domain offers — the value does not depend on where a walker sits in the
batch (bit-exact), duplicates agree bit for bit, and shifting the data and every offset by the same constant
leaves logL unchanged."""
import os
import sys

import numpy as np
import pytest

from conftest import REPO

pytestmark = pytest.mark.gpu


def _workload(name):
    sys.path.insert(0, REPO)
    import bench
    return bench.build_workload(name), bench.valid_thetas


@pytest.mark.parametrize("name,n_walkers,n_check", [("c4", 4096, 12), ("c5", 2048, 6)])
def test_full_size_parity_and_invariances(name, n_walkers, n_check):
    from astroemperor_b200.engine import LikelihoodEngine
    from oracle.rv_oracle import RVOracle
    (w, data, spec), valid_thetas = _workload(name)
    assert len(data.t) == w["n"]
    th = valid_thetas(spec, n_walkers, seed=5)
    eng = LikelihoodEngine(spec, data.t, data.y, data.yerr, data.flag)
    ll, lp = eng.logl_batch(th)
    assert np.all(np.isfinite(lp)) and np.all(np.isfinite(ll))
    # (1) a subset against the oracle at full N
    rng = np.random.default_rng(1)
    pick = rng.choice(n_walkers, n_check, replace=False)
    orc = RVOracle(spec.compile(), data.t, data.y, data.yerr, data.flag)
    ref = np.array([orc.my_likelihood(th[i]) for i in pick])
    rel = np.abs(ll[pick] - ref) / np.abs(ref)
    assert rel.max() < 1e-12, (name, rel.max())
    # (2) position in the batch does not matter; duplicates are bit-identical
    perm = rng.permutation(n_walkers)
    ll_p, _ = eng.logl_batch(np.concatenate([th[perm], th[:64]]))
    assert np.array_equal(ll_p[:n_walkers], ll[perm]) and np.array_equal(ll_p[n_walkers:], ll[:64])
    # (3) y -> y + c together with every offset -> offset + c: the residuals are the same numbers up to rounding
    cm = spec.compile()
    c = 0.37
    free_of_full = {int(f): j for j, f in enumerate(cm.free_to_full)}
    th_s = th[:256].copy()
    for i in range(cm.n_ins):
        th_s[:, free_of_full[cm.offset_off + i]] += c
    eng_s = LikelihoodEngine(spec, data.t, data.y + c, data.yerr, data.flag)
    ll_s, lp_s = eng_s.logl_batch(th_s)
    ok = np.isfinite(lp_s)                      # a shifted offset may leave its prior box
    assert ok.sum() > 200
    assert np.max(np.abs(ll_s[ok] - ll[:256][ok]) / np.abs(ll[:256][ok])) < 1e-11

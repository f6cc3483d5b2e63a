"""GPU: the device Kepler solver (FP32 Markley starter + FP64 refinement with MUFU-seeded
reciprocals) against the oracle restatement of kepler.py and against the equation itself."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_device_solver_matches_oracle_and_residual():
    from astroemperor_b200.engine import kepler_solve
    from oracle import kepler_shim
    rng = np.random.default_rng(3)
    for e in [0.0, 1e-7, 0.05, 0.3, 0.6, 0.9, 0.99, 0.999, 0.9999]:
        M = np.concatenate([np.linspace(0, 2 * np.pi, 20001), rng.uniform(-50, 1e4, 20000),
                            [1e-300, 1e-20, 1e-12, 1e-7, np.pi, np.pi - 1e-9, np.pi + 1e-9, 2 * np.pi - 1e-9,
                             -3.7, 1e5, 7e5]])
        E = kepler_solve(M, e)
        Eo = kepler_shim.solve(M, np.full_like(M, e))
        d = np.abs(E - Eo)
        assert d.max() <= 2.0e-15 / max(1.0 - e, 1e-3) + 8.9e-16, (e, d.max(), M[np.argmax(d)])
        Mw = np.mod(M, 2 * np.pi)
        res = E - e * np.sin(E) - Mw
        res = (res + np.pi) % (2 * np.pi) - np.pi
        assert np.abs(res).max() <= 4 * np.finfo(float).eps * 2 * np.pi, (e, np.abs(res).max())


def test_device_solver_vector_ecc_and_edges():
    from astroemperor_b200.engine import kepler_solve
    from oracle import kepler_shim
    rng = np.random.default_rng(5)
    M = rng.uniform(0, 2 * np.pi, 100000)
    e = rng.uniform(0, 0.999, 100000)
    E = kepler_solve(M, e)
    Eo = kepler_shim.solve(M, e)
    assert np.max(np.abs(E - Eo) * (1 - e)) < 4e-15
    assert kepler_solve(np.array([0.0]), 0.5)[0] == 0.0
    assert np.isnan(kepler_solve(np.array([np.nan, np.inf]), 0.5)).all()
    assert kepler_solve(np.zeros(0), 0.1).shape == (0,)

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


def test_grid_core_matches_oracle_and_residual():
    """The solver the likelihood kernel runs (grid-anchored core, emp_kepler_grid_host) element by
    element: E against the oracle restatement of kepler.py, the residual of Kepler's equation, and
    sin E / cos E (what the RV term is built from) against the oracle's E."""
    from astroemperor_b200.engine import kepler_solve_grid
    from oracle import kepler_shim
    rng = np.random.default_rng(11)
    for e in [0.0, 1e-7, 0.05, 0.3, 0.6, 0.9, 0.95, 0.98, 0.985, 0.999]:   # > 0.98: kepler.py-style path
        M = np.concatenate([np.linspace(0, 2 * np.pi, 20001), rng.uniform(-50, 1e4, 20000),
                            np.pi * rng.uniform(size=5000) ** 6,          # crowd the periapsis
                            [0.0, 1e-300, 1e-20, 1e-12, 1e-7, np.pi, np.pi - 1e-9, np.pi + 1e-9, 2 * np.pi - 1e-9,
                             -3.7, 1e5, 7e5, 3e12]])
        E, s, c = kepler_solve_grid(M, e)
        Eo = kepler_shim.solve(M, np.full_like(M, e))
        d = np.abs(E - Eo)
        assert d.max() <= 2.0e-15 / max(1.0 - e, 1e-3) + 8.9e-16, (e, d.max(), M[np.argmax(d)])
        Mw = np.mod(M, 2 * np.pi)
        res = E - e * np.sin(E) - Mw
        res = (res + np.pi) % (2 * np.pi) - np.pi
        assert np.abs(res).max() <= 4 * np.finfo(float).eps * 2 * np.pi, (e, np.abs(res).max())
        tol = 2.0e-15 / max(1.0 - e, 1e-3) + 4.5e-16
        assert np.abs(s - np.sin(Eo)).max() <= tol and np.abs(c - np.cos(Eo)).max() <= tol, e


def test_grid_core_vector_ecc_and_edges():
    from astroemperor_b200.engine import kepler_solve_grid
    from oracle import kepler_shim
    rng = np.random.default_rng(12)
    M = rng.uniform(0, 2 * np.pi, 200000)
    e = rng.uniform(0, 0.999, 200000)
    E, s, c = kepler_solve_grid(M, e)
    Eo = kepler_shim.solve(M, e)
    assert np.max(np.abs(E - Eo) * (1 - e)) < 4e-15
    assert np.max(np.abs(s - np.sin(Eo)) * (1 - e)) < 4e-15
    E0, s0, c0 = kepler_solve_grid(np.array([0.0]), 0.5)
    assert E0[0] == 0.0 and s0[0] == 0.0 and c0[0] == 1.0
    assert np.isnan(kepler_solve_grid(np.array([np.nan, np.inf]), 0.5)[0]).all()
    assert kepler_solve_grid(np.zeros(0), 0.1)[0].shape == (0,)


def test_table_started_grid_core_matches_oracle_and_residual():
    """Likelihood kernel v10: planets with e <= 0.8 take their starter from the per-walker table built in the
    kernel prologue instead of the Markley formula (emp_kepler_grid_table_host runs exactly that path).  Same
    bars as the Markley-started core: E, sin E, cos E against the oracle element by element, residual <= 4 ulp."""
    from astroemperor_b200.engine import kepler_solve_grid, kepler_solve_grid_table
    from oracle import kepler_shim
    rng = np.random.default_rng(13)
    for e in [0.0, 1e-7, 0.05, 0.3, 0.5, 0.6, 0.7, 0.75, 0.79, 0.8]:
        M = np.concatenate([np.linspace(0, 2 * np.pi, 40001), rng.uniform(-50, 1e4, 20000),
                            np.pi * rng.uniform(size=5000) ** 6,          # crowd the periapsis
                            [0.0, 1e-300, 1e-20, 1e-12, 1e-7, np.pi, np.pi - 1e-9, np.pi + 1e-9, 2 * np.pi - 1e-9,
                             -3.7, 1e5, 7e5]])
        E, s, c = kepler_solve_grid_table(M, e)
        Eo = kepler_shim.solve(M, np.full_like(M, e))
        d = np.abs(E - Eo)
        assert d.max() <= 2.0e-15 / (1.0 - e) + 8.9e-16, (e, d.max(), M[np.argmax(d)])
        Mw = np.mod(M, 2 * np.pi)
        res = E - e * np.sin(E) - Mw
        res = (res + np.pi) % (2 * np.pi) - np.pi
        assert np.abs(res).max() <= 4 * np.finfo(float).eps * 2 * np.pi, (e, np.abs(res).max())
        tol = 2.0e-15 / (1.0 - e) + 4.5e-16
        assert np.abs(s - np.sin(Eo)).max() <= tol and np.abs(c - np.cos(Eo)).max() <= tol, e
        # and against the Markley-started core: the same root
        E2, s2, c2 = kepler_solve_grid(M, e)
        assert np.abs(E - E2).max() <= 4e-15 / (1.0 - e)
    with pytest.raises(Exception):
        kepler_solve_grid_table(np.zeros(4), 0.9)

"""Seeded synthetic RV data sets of the BASELINE shapes.

Data generation only (runs once on the host, not on the hot path): timestamps
`sort(U(0, 4000 d))`, instrument id uniform per point, `yerr ~ U(1, 3)` m/s, a
sum of Keplerians with the table below, per-instrument offsets, white noise
`N(0, yerr^2 + 2^2)` and optionally a true MA(1) term (phi = 0.2, tau = 10 d).
The result is returned per instrument, i.e. in the shape of the `.vels` files
the reference's `DataWrapper.mk_RV` reads (qol_utils.py:61-100).
"""
import numpy as np

# (P [d], K [m/s], e)
PLANETS = [(12.3, 50.0, 0.05), (45.6, 30.0, 0.10), (111.0, 20.0, 0.20),
           (365.0, 10.0, 0.30), (1200.0, 5.0, 0.40)]
OFFSETS = [10.0, -20.0, 5.0, 0.0]


def _solve_kepler_newton(M, e, iters=60):
    """Plain Newton iteration; only used to synthesise data."""
    M = np.mod(M, 2 * np.pi)
    E = np.where(e < 0.8, M, np.pi * np.ones_like(M))
    for _ in range(iters):
        E = E - (E - e * np.sin(E) - M) / (1.0 - e * np.cos(E))
    return E


def keplerian_rv(t, per, K, phase, e, w):
    M = 2.0 * np.pi / per * t + phase
    E = _solve_kepler_newton(M, e)
    f = 2.0 * np.arctan2(np.sqrt(1 + e) * np.sin(E / 2), np.sqrt(1 - e) * np.cos(E / 2))
    return K * (np.cos(f + w) + e * np.cos(w))


def make_synthetic_rv(seed, n, nins, kplan, ma=False, span=4000.0):
    """Returns a list of `nins` tuples (t, rv, erv), one per instrument file."""
    rng = np.random.default_rng(seed)
    t = np.sort(rng.uniform(0.0, span, n))
    ins = rng.integers(0, nins, n)
    # every instrument needs at least two points for the reference's loader
    for j in range(nins):
        if np.sum(ins == j) < 2:
            ins[2 * j:2 * j + 2] = j
    yerr = rng.uniform(1.0, 3.0, n)
    rv = np.zeros(n)
    for (per, K, e) in PLANETS[:kplan]:
        phase, w = rng.uniform(0, 2 * np.pi, 2)
        rv += keplerian_rv(t, per, K, phase, e, w)
    rv += np.asarray(OFFSETS * (1 + nins // len(OFFSETS)))[ins]
    eps = rng.normal(0.0, np.sqrt(yerr ** 2 + 2.0 ** 2))
    if ma:
        phi, tau = 0.2, 10.0
        for i in range(1, n):
            rv[i] += phi * np.exp(-abs(t[i] - t[i - 1]) / tau) * eps[i - 1]
    rv += eps
    t = t + 2450000.0  # BJD-like absolute epochs; the loader subtracts common_t
    return [(t[ins == j], rv[ins == j], yerr[ins == j]) for j in range(nins)]


def add_activity_columns(files, counts, seed):
    """Synthetic stellar-activity indices: `counts[i]` extra columns for instrument i, weakly correlated with
    its RVs (the shape of a `.vels` file with columns after eRV, qol_utils.py:74-79).  Returns
    (t, rv, erv, activity[n_i, counts[i]]) per instrument."""
    rng = np.random.default_rng(seed + 1000)
    out = []
    for i, (t, rv, erv) in enumerate(files):
        cols = [rng.normal(size=len(t)) * 3.0 + 0.1 * rv for _ in range(counts[i])]
        out.append((t, rv, erv, np.column_stack(cols) if cols else np.zeros((len(t), 0))))
    return out

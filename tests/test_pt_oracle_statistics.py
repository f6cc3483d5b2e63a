"""CPU: the sampler oracle (restated emcee stretch move + ptemcee-lineage swap sweep, "parity unpinned"
at the package boundary) is at least a CORRECT parallel-tempering sampler: on a Gaussian target every rung
samples its tempered posterior N(0, sigma^2 / beta) (moments and <logL> = -nd/(2 beta)), the swaps preserve the
ensemble as a multiset, and the thermodynamic integral over the ladder matches its analytic value.  The device path is held
bit-identical to this oracle (tests/test_pt_gpu.py), so this pins the algorithm both run."""
import numpy as np

from astroemperor_b200.draws import DrawStreams, draw_sweep
from astroemperor_b200.postproc import evidence_ti
from oracle.pt_oracle import stretch_step, swap_sweep


def test_tempered_gaussian_moments_swaps_and_evidence():
    nd, T, W, sweeps, burn = 2, 5, 64, 1500, 300
    sigma, half = 1.0, 25.0
    betas = np.array([1.0, 0.5, 0.25, 0.1, 0.03])

    def loglike(q):
        ll = -0.5 * np.sum(q * q, axis=1) / sigma ** 2
        inside = np.all(np.abs(q) <= half, axis=1)
        lp = np.where(inside, -nd * np.log(2 * half), -np.inf)
        return np.where(inside, ll, -np.inf), lp

    rng = np.random.default_rng(0)
    p = rng.normal(size=(T, W, nd)) * 3
    ll, lp = loglike(p.reshape(-1, nd))
    logl, logp = ll.reshape(T, W), lp.reshape(T, W)
    streams = DrawStreams(11, T)
    chain = np.empty((sweeps, T, W, nd))
    lls = np.empty((sweeps, T, W))
    swaps = np.zeros(T - 1)
    for k in range(sweeps):
        d = draw_sweep(streams, W, nd, 1)
        stretch_step(p, logl, logp, betas, d.half_idx[0], d.zz[0], d.rint[0], d.factors[0], d.lnu[0], loglike)
        before = np.sort(logl.reshape(-1))
        n_acc, src, _ = swap_sweep(p, logl, logp, betas, d.perm, d.lnu_swap)
        assert np.array_equal(np.sort(logl.reshape(-1)), before)          # a swap sweep only permutes walkers
        assert np.array_equal(np.sort(src.reshape(-1)), np.arange(T * W))
        swaps += n_acc
        chain[k], lls[k] = p, logl
    x = chain[burn:]
    for t, b in enumerate(betas):
        var = x[:, t].reshape(-1, nd).var(axis=0)
        assert np.allclose(var, sigma ** 2 / b, rtol=0.12), (b, var)
        assert abs(x[:, t].mean()) < 0.25 * np.sqrt(sigma ** 2 / b)
    rate = swaps / (sweeps * W)
    assert np.all(rate > 0.2) and np.all(rate < 0.95), rate
    # thermodynamic integration over THIS ladder: <logL>_beta = -nd / (2 beta) on [beta_min, 1], extended flat
    # to 0 by the estimator: -(nd/2) ln(1/beta_min) - nd/2 (the true logZ needs a ladder that reaches the prior)
    expect = -0.5 * nd * np.log(1.0 / betas[-1]) - 0.5 * nd
    z, err = evidence_ti(np.moveaxis(lls[burn:], 1, 0).reshape(T, -1), betas, pchip=True)
    assert abs(z - expect) < 0.5, (z, expect, err)
    mean_ll = lls[burn:].mean(axis=(0, 2))
    assert np.allclose(mean_ll, -0.5 * nd / betas, rtol=0.1), mean_ll

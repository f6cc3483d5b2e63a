"""GPU parity of the parallel-tempering step against oracle/pt_oracle.py with identical
host-supplied draws: accept / swap decisions and the resulting chains must be bit-identical
(BASELINE.json north_star); log-likelihoods within 1e-10 relative."""
import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu


def _setup(name, T, W, seed, betas=None, **kw):
    from astroemperor_b200.engine import LikelihoodEngine
    from astroemperor_b200.sampler import PTSampler
    from oracle.rv_oracle import RVOracle
    from oracle.pt_oracle import PTOracle
    g, spec = load_golden(name)
    eng = LikelihoodEngine(spec, g["t"], g["y"], g["yerr"], g["flag"])
    samp = PTSampler(W, eng.ndim, eng, ntemps=T, seed=seed, betas=betas, **kw)
    p0 = samp.initial_positions(spec)
    orc = PTOracle(RVOracle(spec.compile(), g["t"], g["y"], g["yerr"], g["flag"]), samp.betas,
                   adapt_tau=samp.adapt_tau, adapt_nu=samp.adapt_nu, adapt=samp.adapt)
    return g, spec, eng, samp, orc, p0


@pytest.mark.parametrize("name,T,W,nsweeps,nsteps", [
    ("c1_51peg_k1_p0", 2, 100, 6, 2),            # BASELINE config 1 shape
    ("c2_synth3p_2ins_n400", 4, 64, 4, 1),       # 3 planets, S/C parameterisation, derived-ecc prior
    ("synth_k1_p0_ma1_global", 3, 32, 4, 3),     # global MA recurrence in the likelihood
    ("synth_k1_p0_acc2_fixed", 5, 24, 4, 1),     # fixed parameter + acceleration, odd T
])
def test_sweeps_bit_identical_to_oracle(name, T, W, nsweeps, nsteps):
    g, spec, eng, samp, orc, p0 = _setup(name, T, W, seed=11)
    samp._init_state(p0)
    orc.init_state(p0)
    p, ll, lp = samp.state_numpy()
    assert np.array_equal(lp, orc.logp)
    n_dec = 0
    for k in range(nsweeps):
        d = samp.draw(nsteps)
        n_acc = samp.sweep(d)
        acc_o, n_acc_o, src_o = orc.sweep(d)
        p, ll, lp = samp.state_numpy()
        assert np.array_equal(n_acc, n_acc_o), (k, n_acc, n_acc_o)
        assert np.array_equal(samp._src.cpu().numpy(), src_o), f"swap plan differs in sweep {k}"
        assert np.array_equal(p, orc.p), f"chain differs from the oracle in sweep {k}"
        assert np.array_equal(lp, orc.logp)
        assert np.array_equal(samp.betas, orc.betas), "ladder adaptation differs"
        fin = np.isfinite(orc.logl)
        assert np.max(np.abs(ll[fin] - orc.logl[fin]) / np.abs(orc.logl[fin])) < 1e-10
        n_dec += T * W * nsteps + (T - 1) * W
    # the last accept mask too (last step of the last sweep)
    assert np.array_equal(samp.accepted.cpu().numpy().astype(bool), acc_o[-1])
    print(f"{name}: {n_dec} decisions identical, min decision margin {orc.min_margin:.3e}")
    assert orc.min_margin > 1e-9  # otherwise the case is too close to call and should be re-seeded


def test_run_mcmc_api_and_storage():
    g, spec, eng, samp, orc, p0 = _setup("c1_51peg_k1_p0", 3, 32, seed=3)
    samp.run_mcmc(p0, nsweeps=12, nsteps=2, progress=False)
    ch = samp.get_chain()
    assert ch.shape == (3, 12, 32, eng.ndim)
    assert samp.get_chain(flat=True, discard=2, thin=2).shape == (3, 5 * 32, eng.ndim)
    ll = samp.get_log_like()
    lp = samp.get_log_prior()
    assert ll.shape == (3, 12, 32) and np.all(np.isfinite(ll)) and np.all(np.isfinite(lp))
    assert samp.get_log_prob(flat=True).shape == (3, 12 * 32)
    af = samp.acceptance_fraction
    assert af.shape == (3, 32) and 0.0 < af.mean() < 1.0
    assert samp.get_betas().shape == (12, 3) and samp.get_tsw().shape == (12, 2)
    # stored likelihoods are the likelihoods of the stored positions
    ll_re, _ = eng.logl_batch(ch[:, -1].reshape(-1, eng.ndim))
    assert np.array_equal(ll_re.reshape(3, 32), ll[:, -1])
    # continue the run: storage grows
    samp.run_mcmc(None, nsweeps=3, nsteps=1)
    assert samp.get_chain().shape[1] == 15
    logz, err = samp.get_evidence_ti()
    assert np.isfinite(logz)


def test_same_seed_same_chain_and_host_store():
    g, spec, eng, s1, _, p0 = _setup("c2_synth3p_2ins_n400", 2, 16, seed=9)
    s1.run_mcmc(p0, nsweeps=5, nsteps=1)
    g, spec, eng2, s2, _, p0b = _setup("c2_synth3p_2ins_n400", 2, 16, seed=9, store="host")
    assert np.array_equal(p0, p0b)
    s2.run_mcmc(p0b, nsweeps=5, nsteps=1)
    assert np.array_equal(s1.get_chain(), s2.get_chain())
    assert np.array_equal(s1.get_log_like(), s2.get_log_like())


def test_swap_moves_walkers_down_several_rungs():
    """A hot walker with a much better likelihood is handed down rung by rung in ONE sweep
    (the reason the plan is sequential)."""
    import torch
    from astroemperor_b200.engine import LikelihoodEngine
    g, spec = load_golden("c1_51peg_k0")
    eng = LikelihoodEngine(spec, g["t"], g["y"], g["yerr"], g["flag"])
    T, W = 6, 4
    ll = np.full((T, W), -100.0)
    ll[T - 1, 2] = -1.0
    betas = np.linspace(1.0, 0.1, T)
    perm = np.tile(np.arange(W, dtype=np.int32), (T - 1, 2, 1))
    perm[:, 1] = np.roll(perm[:, 1], 1, axis=-1)  # slot a of temp i <-> slot a-1... of temp i-1
    lnu = np.full((T - 1, W), -1e-3)
    dev = "cuda"
    src = torch.empty((T, W), dtype=torch.int32, device=dev)
    n_acc = torch.empty((T - 1,), dtype=torch.int32, device=dev)
    eng.pt_swap_plan(torch.as_tensor(ll, device=dev), torch.as_tensor(betas, device=dev),
                     torch.as_tensor(perm, device=dev), torch.as_tensor(lnu, device=dev), src, n_acc)
    from oracle.pt_oracle import swap_sweep
    p = np.zeros((T, W, 1))
    n_o, src_o, _ = swap_sweep(p, ll.copy(), np.zeros((T, W)), betas, perm, lnu)
    assert np.array_equal(src.cpu().numpy(), src_o) and np.array_equal(n_acc.cpu().numpy(), n_o)
    assert (T - 1) * W + 2 in src_o[0]  # the hot walker reached the coldest chain


def test_postprocessing_reductions_and_smd(tmp_path):
    """smd history, autocorrelation time, TI / stepping-stone evidence,
    chain sink in the reference's backend layout."""
    T = 24  # fixed ladder down to beta = 1e-7 so that the estimators see (almost) the prior
    g, spec, eng, samp, orc, p0 = _setup("c1_51peg_k0", T, 32, seed=2, betas=np.geomspace(1, 1e-7, T), adapt=False)
    samp.D_ = spec.prior_widths()
    samp.run_mcmc(p0, nsweeps=400, nsteps=2, progress=False)
    smd = samp.get_smd()
    assert smd.shape == (400, T - 1) and np.all(smd >= 0) and smd[5:].mean() > 0
    tau = samp.get_autocorr_time(discard=100, quiet=True)
    assert tau.shape == (T, eng.ndim) and np.all(np.isfinite(tau)) and np.all(tau > 0)
    with pytest.raises(RuntimeError):
        samp.get_autocorr_time(discard=380, quiet=False, tol=1000)
    z_ti, e_ti = samp.get_evidence_ti(discard=100)
    z_ss, e_ss = samp.get_evidence_ss(discard=100)
    z_hy, e_hy = samp.get_evidence_hybrid(discard=100)
    assert np.isfinite([z_ti, z_ss, z_hy, e_ti, e_hy]).all() and z_hy == z_ss
    # 2-parameter model (offset + jitter): evidence by brute-force quadrature over the prior box
    fp = spec.free_params()
    g0 = np.linspace(fp[0].limits[0], fp[0].limits[1], 1201)
    g1 = np.linspace(fp[1].limits[0], fp[1].limits[1], 1201)
    G0, G1 = np.meshgrid(g0, g1, indexing="ij")
    ll, lp = eng.logl_batch(np.column_stack([G0.ravel(), G1.ravel()]))
    post = (ll + lp).reshape(G0.shape)
    mx = post.max()
    trap = np.trapezoid if hasattr(np, "trapezoid") else np.trapz
    logz_ref = mx + np.log(trap(trap(np.exp(post - mx), g1, axis=1), g0))
    print(f"logZ: stepping-stone {z_ss:.3f} +- {e_ss:.3f}, TI {z_ti:.3f} +- {e_ti:.3f}, quadrature {logz_ref:.3f}")
    assert abs(z_ss - logz_ref) < 0.5 and abs(z_ti - logz_ref) < 2.0, (z_ss, z_ti, logz_ref)
    path = samp.save_backend(str(tmp_path / "run"), discard=10)
    assert path.endswith((".npz", ".h5"))
    if path.endswith(".npz"):
        d = np.load(path)
        assert d["chain"].shape == (T, 390, 32, eng.ndim) and d["beta_history"].shape == (390, T)

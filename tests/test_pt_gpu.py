"""GPU parity of the parallel-tempering step against oracle/pt_oracle.py with identical
host-supplied draws: accept / swap decisions and the resulting chains must be bit-identical
(BASELINE.json north_star); log-likelihoods within 1e-10 relative."""
import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu


def _setup(name, T, W, seed, betas=None, with_D=False, **kw):
    from astroemperor_b200.engine import LikelihoodEngine
    from astroemperor_b200.sampler import PTSampler
    from oracle.rv_oracle import RVOracle
    from oracle.pt_oracle import PTOracle
    g, spec = load_golden(name)
    eng = LikelihoodEngine(spec, g["t"], g["y"], g["yerr"], g["flag"])
    samp = PTSampler(W, eng.ndim, eng, ntemps=T, seed=seed, betas=betas, **kw)
    if with_D:
        samp.D_ = spec.prior_widths()
    p0 = samp.initial_positions(spec)
    orc = PTOracle(RVOracle(spec.compile(), g["t"], g["y"], g["yerr"], g["flag"]), samp.betas,
                   adapt_tau=samp.adapt_tau, adapt_nu=samp.adapt_nu, adapt=samp.adapt,
                   D=samp.D_ if with_D else None)
    return g, spec, eng, samp, orc, p0


@pytest.mark.parametrize("name,T,W,nsweeps,nsteps,graph", [
    ("c1_51peg_k1_p0", 2, 100, 6, 2, False),            # BASELINE config 1 shape
    ("c2_synth3p_2ins_n400", 4, 64, 4, 1, False),       # 3 planets, S/C parameterisation, derived-ecc prior
    ("c2_synth3p_2ins_n400", 4, 64, 8, 1, True),        # the same sweeps replayed from the captured CUDA graph
    ("synth_k1_p0_ma1_global", 3, 32, 4, 3, True),      # global MA recurrence in the likelihood, graph, nsteps = 3
    ("synth_k1_p0_acc2_fixed", 5, 26, 4, 1, False),     # fixed parameter + acceleration, odd T, W % 4 != 0
    ("c4_synth5p_4ins_ma_global_n600", 9, 1100, 3, 1, True),  # W > 1024: two plan elements per thread
])
def test_sweeps_bit_identical_to_oracle(name, T, W, nsweeps, nsteps, graph):
    """Stretch steps, accept masks, swap plan, swap counts, DEVICE ladder adaptation, swap mean distance: every
    sweep against oracle/pt_oracle.py with the same draws."""
    g, spec, eng, samp, orc, p0 = _setup(name, T, W, seed=11, with_D=True)
    samp._init_state(p0)
    samp._alloc_store(nsweeps * nsteps)
    orc.init_state(p0)
    p, ll, lp = samp.state_numpy()

    def same_logp(a, b):
        # bit-identical, except where libm's pow(z, 2.0) — what `**2` of Normal.prior:8 is on a NumPy scalar — is
        # not the correctly rounded square the device computes: 1 ulp in ~1/2600 Normal priors
        return np.array_equal(a, b) or (np.max(np.abs(a - b) / np.abs(b)) < 4e-16 and np.mean(a != b) < 0.02)
    assert same_logp(lp, orc.logp)
    n_dec = 0
    captures0 = eng.graph_captures
    for k in range(nsweeps):
        d = samp.draw(nsteps)
        n_acc = samp.sweep(samp.stage_draws(d, pinned=True) if graph else d)
        acc_o, n_acc_o, src_o = orc.sweep(d)
        p, ll, lp = samp.state_numpy()
        assert np.array_equal(n_acc, n_acc_o), (k, n_acc, n_acc_o)
        assert np.array_equal(samp._src.cpu().numpy(), src_o), f"swap plan differs in sweep {k}"
        assert np.array_equal(p, orc.p), f"chain differs from the oracle in sweep {k}"
        assert same_logp(lp, orc.logp)
        assert np.array_equal(samp.betas, orc.betas), "ladder adaptation (on the device) differs"
        fin = np.isfinite(orc.logl)
        assert np.max(np.abs(ll[fin] - orc.logl[fin]) / np.abs(orc.logl[fin])) < 1e-10
        smd = samp.get_smd()
        assert smd.shape == (k + 1, T - 1)
        assert np.allclose(smd[-1], orc.smd[: T - 1], rtol=1e-12, atol=0), "swap mean distance differs"
        n_dec += T * W * nsteps + (T - 1) * W
    # the last accept mask too (last step of the last sweep)
    assert np.array_equal(samp.accepted.cpu().numpy().astype(bool), acc_o[-1])
    # histories kept on the device: one row per sweep
    assert np.array_equal(samp.get_betas()[-1], orc.betas) and samp.get_betas().shape == (nsweeps * nsteps, T)
    assert samp.get_betas_sweeps().shape == (nsweeps, T)
    assert np.array_equal(samp.get_tsw()[-1], n_acc_o / W)
    # every stretch step is a stored sample; the last one of a sweep is the state after the swap
    ch = samp.get_chain()
    assert ch.shape == (T, nsweeps * nsteps, W, eng.ndim) and np.array_equal(ch[:, -1], orc.p)
    if graph:  # two argument blocks alternate (state and staging are double-buffered): two captures, then replays
        assert eng.graph_captures - captures0 <= 2
    print(f"{name}: {n_dec} decisions identical, min decision margin {orc.min_margin:.3e}")
    assert orc.min_margin > 1e-9  # otherwise the case is too close to call and should be re-seeded


def test_launches_per_sweep():
    """a sweep with nsteps = 1 is at most 6 kernel launches (was 13 + torch glue)."""
    g, spec, eng, samp, orc, p0 = _setup("c2_synth3p_2ins_n400", 10, 512, seed=5, with_D=True, chunk=1)
    samp._init_state(p0)
    samp._alloc_store(13)   # storage of the whole test up front: a re-allocation moves the chain, i.e. new graphs
    samp._alloc_hist(13)
    samp.run_mcmc(None, nsweeps=3, nsteps=1)
    l0 = eng.launch_count
    samp.run_mcmc(None, nsweeps=10, nsteps=1)
    assert (eng.launch_count - l0) == 60, eng.launch_count - l0
    assert eng.graph_captures <= 2  # state and staging are double-buffered in step: two argument blocks alternate


@pytest.mark.parametrize("name,T,W,nsweeps,nsteps,kw", [
    ("c1_51peg_k1_p0", 2, 100, 37, 1, dict()),                              # BASELINE config 1 shape, automatic chunk
    ("c2_synth3p_2ins_n400", 4, 64, 11, 2, dict(store="host", thin_by=3, chunk=4)),  # ring drained per chunk
    ("c2_synth3p_2ins_n400", 3, 32, 10, 1, dict(chunk=3)),                  # odd chunk: the parity alternates
    ("synth_k1_p0_ma1_global", 1, 32, 9, 1, dict(chunk=4)),                 # a single temperature: no swap sweep
])
def test_chunked_run_equals_sweep_by_sweep(name, T, W, nsweeps, nsteps, kw):
    """run_mcmc replays k sweeps per graph launch for small ensembles (emp_pt_sweep_chunk: one upload node + k
    sweeps): chains, histories and counters must be the bits of the sweep-by-sweep run, across calls too."""
    g, spec, eng_a, a, _, p0 = _setup(name, T, W, seed=21, with_D=True, **{**kw, "chunk": 1})
    g, spec, eng_b, b, _, p0b = _setup(name, T, W, seed=21, with_D=True, **kw)
    assert b._chunk_len(nsteps) > 1 and a._chunk_len(nsteps) == 1
    for s, q in ((a, p0), (b, p0b)):
        st = s.run_mcmc(q, nsweeps=nsweeps, nsteps=nsteps)
        s.run_mcmc(st, nsweeps=3, nsteps=nsteps)          # continue: a staged sweep may be pending
        s.run_mcmc(None, nsweeps=1, nsteps=nsteps)        # a single sweep takes the sweep-by-sweep path
    l0 = eng_b.launch_count
    b.run_mcmc(None, nsweeps=2 * b._chunk_len(nsteps), nsteps=nsteps)
    a.run_mcmc(None, nsweeps=2 * b._chunk_len(nsteps), nsteps=nsteps)
    per_sweep = (eng_b.launch_count - l0) / (2 * b._chunk_len(nsteps))
    assert per_sweep <= 6 * nsteps + (2 if T > 1 else 1), per_sweep
    assert np.array_equal(a.get_chain(), b.get_chain())
    assert np.array_equal(a.get_log_like(), b.get_log_like())
    assert np.array_equal(a.get_log_prob(), b.get_log_prob())
    assert np.array_equal(a.get_betas(), b.get_betas()) and np.array_equal(a.betas, b.betas)
    assert np.array_equal(a.get_tsw(), b.get_tsw()) and np.array_equal(a.get_smd(), b.get_smd())
    assert np.array_equal(a.acceptance_fraction, b.acceptance_fraction)
    pa, pb = a.state_numpy(), b.state_numpy()
    assert all(np.array_equal(x, y) for x, y in zip(pa, pb))
    assert a.iteration == b.iteration and a._n_steps == b._n_steps and a._stored == b._stored


def test_chain_plan_kernel_equals_shared_memory_plan_kernels(monkeypatch):
    """The swap plan as W independent chains (pt_swap_plan_chain_kernel, the default) against the single-CTA
    shared-memory kernels it replaced (EMP_PLAN_NO_CHAIN=1), on ladders larger than the oracle tests can afford:
    plans, swap counts, adapted ladders and chains must be the same bits."""
    for T, W in ((48, 2048), (9, 4100), (130, 256)):
        runs = []
        for no_chain in ("0", "1"):
            monkeypatch.setenv("EMP_PLAN_NO_CHAIN", no_chain)   # read when the handle first prepares a plan
            g, spec, eng, samp, _, p0 = _setup("c1_51peg_k1_p0", T, W, seed=31, with_D=True, chunk=1, store=None)
            samp.run_mcmc(p0, nsweeps=4, nsteps=1)
            runs.append((samp.state_numpy(), samp._src.cpu().numpy(), samp.betas.copy(), samp.get_tsw(), samp.get_smd()))
            del samp
            eng.close()
        a, b = runs
        assert all(np.array_equal(x, y) for x, y in zip(a[0], b[0])), (T, W)
        assert np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2]) and np.array_equal(a[3], b[3]), (T, W)
        assert np.array_equal(a[4], b[4]), (T, W)
        assert a[3].mean() > 0.05  # swaps did happen


def test_run_mcmc_api_and_storage():
    g, spec, eng, samp, orc, p0 = _setup("c1_51peg_k1_p0", 3, 32, seed=3)
    state = samp.run_mcmc(p0, nsweeps=12, nsteps=2, progress=False)
    ch = samp.get_chain()
    # reddemcee stores every stretch step: nsweeps*nsteps samples (emp.py:2514-2516 discards in steps)
    assert ch.shape == (3, 24, 32, eng.ndim)
    assert samp.get_chain(flat=True, discard=4, thin=2).shape == (3, 10 * 32, eng.ndim)
    ll = samp.get_log_like()
    lp = samp.get_log_prior()
    assert ll.shape == (3, 24, 32) and np.all(np.isfinite(ll)) and np.all(np.isfinite(lp))
    assert samp.get_log_prob(flat=True).shape == (3, 24 * 32)
    af = samp.acceptance_fraction
    assert af.shape == (3, 32) and 0.0 < af.mean() < 1.0
    # per-sweep histories
    assert samp.get_betas().shape == (24, 3) and samp.get_betas_sweeps().shape == (12, 3) and samp.get_tsw().shape == (12, 2)
    # stored likelihoods are the likelihoods of the stored positions
    for j in (-1, -2, 0):
        ll_re, _ = eng.logl_batch(ch[:, j].reshape(-1, eng.ndim))
        assert np.array_equal(ll_re.reshape(3, 32), ll[:, j])
    assert np.array_equal(state.coords, ch[:, -1])
    # continue the run (the returned state is accepted like emcee's): storage grows
    samp.run_mcmc(state, nsweeps=3, nsteps=1)
    assert samp.get_chain().shape[1] == 27 and samp.get_betas().shape == (27, 3) and samp.get_tsw().shape == (15, 2)
    assert np.array_equal(samp.get_chain()[:, :24], ch)
    logz, err = samp.get_evidence_ti()
    assert np.isfinite(logz)
    bk = samp.backend
    assert bk[0].iteration == 27 and bk[0].get_chain().shape == (27, 32, eng.ndim)
    # thinning counts stretch steps
    g, spec, eng, s2, _, p0 = _setup("c1_51peg_k1_p0", 3, 32, seed=3, thin_by=3)
    s2.run_mcmc(p0, nsweeps=12, nsteps=2)
    assert s2.get_chain().shape[1] == 8 and np.array_equal(s2.get_chain(), ch[:, 0:24:3])


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
    assert samp.get_chain().shape[1] == 800
    tau = samp.get_autocorr_time(discard=200, quiet=True)
    assert tau.shape == (T, eng.ndim) and np.all(np.isfinite(tau)) and np.all(tau > 0)
    with pytest.raises(RuntimeError):
        samp.get_autocorr_time(discard=760, quiet=False, tol=1000)
    z_ti, e_ti = samp.get_evidence_ti(discard=200)
    z_ss, e_ss = samp.get_evidence_ss(discard=200)
    assert np.isfinite([z_ti, z_ss, e_ti]).all()
    with pytest.raises(NotImplementedError):  # reddemcee's own estimator: EMPEROR's try/except falls back to TI
        samp.get_evidence_hybrid(discard=100)
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
    assert path.endswith(".h5")
    from astroemperor_b200.postproc import load_backend
    r = load_backend(str(tmp_path / "run"))   # the reference's per-temperature HDF5 layout, read back without a device
    assert r.get_chain().shape == (T, 790, 32, eng.ndim) and r.get_betas().shape == (790, T)
    assert np.array_equal(r.get_chain(), samp.get_chain(discard=10))
    assert np.array_equal(r.get_log_prob(), samp.get_log_prob(discard=10))
    assert r.get_tsw().shape == (395, T - 1) and np.array_equal(r.get_tsw(), samp.get_tsw(discard=5))
    assert np.allclose(r.acceptance_fraction, samp.acceptance_fraction)
    assert np.isclose(r.get_evidence_ss(discard=190)[0], z_ss)

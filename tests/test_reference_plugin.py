"""The B200 path installed INSIDE the unmodified reference front end (astroemperor_b200/reference_plugin.py).

Build container only (needs /root/reference; the GPU box has no copy of it).  The real `astroemperor.Simulation` is
imported with make_golden.py's presentation stubs, `reference_plugin.install()` routes its reddemcee engine through
this package, and the real `autorun` / `run` / `postprocess` are executed:
  * with an engine factory that stops at the LikelihoodEngine constructor: the descriptor and the data arrays the
    real front end hands over equal the golden fixture of the same configuration;
  * with an ORACLE-backed stand-in for the device (tests may use oracle/): the whole parent flow — run(), the
    backend files, _load_sampler, postprocess() with its get_chain / get_log_like / get_log_prob / get_autocorr_time
    / get_evidence_* / get_betas / get_tsw / get_smd / acceptance_fraction calls (emp.py:961-965, 1375-1447,
    1966-1990) — runs unchanged against this package's read-back API.
"""
import contextlib
import io
import os
import sys
import tempfile

import numpy as np
import pytest

from conftest import REPO, load_golden

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src", "astroemperor")),
                                reason="the reference only exists in the build container")


@pytest.fixture(scope="module")
def ref():
    sys.path.insert(0, os.path.join(REPO, "tests", "golden"))
    import make_golden as mg
    mg._install_stubs()
    if os.path.join(REF, "src") not in sys.path:
        sys.path.insert(0, os.path.join(REF, "src"))
    import astroemperor as emp
    yield emp, mg
    from astroemperor_b200 import reference_plugin
    reference_plugin.uninstall()


class _Reached(Exception):
    pass


def _simulation(emp, mg, workdir, star, configure):
    mg._stage_star(workdir, star)
    sim = emp.Simulation()
    sim.read_loc = workdir + "/"
    sim.save_loc = workdir + "/"
    sim.load_data(star)
    sim.set_engine("reddemcee")
    k = configure(sim)
    return sim, k


def test_autorun_reaches_the_engine_with_the_golden_descriptor(ref):
    emp, mg = ref
    from astroemperor_b200 import reference_plugin
    seen = {}

    def engine_factory(spec, t, y, yerr, flag, am=None, sai=None, device=0):
        seen.update(spec=spec, t=t, y=y, yerr=yerr, flag=flag, am=am)
        raise _Reached()

    shim = reference_plugin.install(engine_factory=engine_factory)
    assert shim.__name__ == "reddemcee"  # the thirty `== 'reddemcee'` branches of emp.py keep working
    workdir = tempfile.mkdtemp(prefix="emp_plugin_")
    cwd = os.getcwd()
    os.chdir(workdir)
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            sim, k = _simulation(emp, mg, workdir, "51Peg", mg.CASES["c1_51peg_k1_p0"]["cfg"])
            assert sim.engine__ is shim
            with pytest.raises(_Reached):
                sim.autorun(k, k)
    finally:
        os.chdir(cwd)
    g, spec = load_golden("c1_51peg_k1_p0")
    assert seen["spec"].to_json() == spec.to_json()
    for key in ("t", "y", "yerr", "flag"):
        assert np.array_equal(seen[key], g[key]), key
    assert seen["am"] is None


def test_autorun_hands_over_the_astrometry_constants(ref):
    emp, mg = ref
    from astroemperor_b200 import reference_plugin
    from astroemperor_b200.amdata import validate
    seen = {}

    def engine_factory(spec, t, y, yerr, flag, am=None, sai=None, device=0):
        seen.update(spec=spec, t=t, am=am)
        raise _Reached()

    reference_plugin.install(engine_factory=engine_factory)
    workdir = tempfile.mkdtemp(prefix="emp_plugin_")
    cwd = os.getcwd()
    os.chdir(workdir)
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            sim, k = _simulation(emp, mg, workdir, "HIP21850", mg.CASES["c3_hip21850_am_k2"]["cfg"])
            with pytest.raises(_Reached):
                sim.autorun(k, k)
    finally:
        os.chdir(cwd)
    g, spec = load_golden("c3_hip21850_am_k2")
    assert seen["spec"].to_json() == spec.to_json()
    am = validate(seen["am"])
    for key, val in am.items():
        ref_val = g["am_" + key]
        assert np.array_equal(np.asarray(val).reshape(-1), np.asarray(ref_val).reshape(-1)), key


class _OracleEngine:
    """CPU stand-in for LikelihoodEngine in this test only: the oracle behind the engine's interface."""

    def __init__(self, spec, t, y, yerr, flag, am=None, sai=None, device=0):
        from oracle.rv_oracle import RVOracle
        self.spec = spec
        self.orc = RVOracle(spec.compile(), t, y, yerr, flag)
        self.ndim = spec.ndim

    def logl_batch(self, th):
        return self.orc.logl_logp_batch(np.asarray(th, dtype=np.float64).reshape(-1, self.ndim))

    def my_likelihood(self, theta):
        return float(self.orc.my_likelihood(np.asarray(theta, dtype=np.float64)))

    def my_prior(self, theta):
        th = np.asarray(theta, dtype=np.float64)
        if th.ndim == 2:
            return self.logl_batch(th)[1]
        return float(self.orc.my_prior(th))

    def my_model(self, theta):
        return self.orc.my_model(np.asarray(theta, dtype=np.float64))


class _OracleSampler:
    """CPU stand-in for PTSampler: oracle/pt_oracle.py sweeps, read back through postproc.StoredRun (the class
    that also answers the getters of a run reloaded from disk)."""

    def __init__(self, nwalkers, ndim, eng, log_prior=None, ntemps=1, betas=None, adapt_tau=1000, adapt_nu=1,
                 seed=7, **kw):
        from astroemperor_b200.draws import DrawStreams, default_betas
        from oracle.pt_oracle import PTOracle
        self.eng, self.nwalkers, self.ndim, self.ntemps = eng, nwalkers, ndim, ntemps
        self.betas = np.array(betas, dtype=float) if betas is not None else default_betas(ndim, ntemps)
        self.streams = DrawStreams(seed, ntemps)
        self.po = PTOracle(eng.orc, self.betas, adapt_tau=adapt_tau, adapt_nu=adapt_nu)
        self.D_ = None
        self._rows = []

    def initial_positions(self, spec):
        from astroemperor_b200.draws import initial_positions
        p0 = initial_positions(self.streams.init, spec, self.ntemps, self.nwalkers)
        for _ in range(100):
            lp = self.eng.my_prior(p0.reshape(-1, self.ndim)).reshape(self.ntemps, self.nwalkers)
            bad = ~np.isfinite(lp)
            if not bad.any():
                break
            p0[bad] = initial_positions(self.streams.init, spec, self.ntemps, self.nwalkers)[bad]
        return p0

    def run_mcmc(self, p0, nsweeps, nsteps=1, progress=False):
        from astroemperor_b200.draws import draw_sweep
        from astroemperor_b200.postproc import StoredRun
        po = self.po
        po.D = self.D_
        po.init_state(p0)
        ch, ll, lpost, bh, tsw, smd, acc = [], [], [], [], [], [], np.zeros((self.ntemps, self.nwalkers))
        for _ in range(nsweeps):
            d = draw_sweep(self.streams, self.nwalkers, self.ndim, nsteps)
            beta_used = po.betas.copy()
            a, n_acc, _ = po.sweep(d)
            acc += a.sum(0)
            for _s in range(nsteps):  # the stand-in stores the post-sweep state for every step of the sweep
                ch.append(po.p.copy()), ll.append(po.logl.copy())
                lpost.append(beta_used[:, None] * po.logl + po.logp), bh.append(po.betas.copy())
            tsw.append(n_acc / self.nwalkers), smd.append(po.smd[: self.ntemps - 1])
        run = StoredRun(np.swapaxes(ch, 0, 1), np.swapaxes(ll, 0, 1), np.swapaxes(lpost, 0, 1), np.array(bh), acc,
                        np.array(tsw), np.array(smd), n_steps=nsweeps * nsteps)
        self.__dict__.update({k: getattr(run, k) for k in ("_chain", "_ll", "_lpost", "_betas", "_accepted",
                                                           "_tsw", "_smd", "iteration", "_n_steps")})
        self._run = run
        self.betas = po.betas.copy()
        return po.p

    def __getattr__(self, name):  # every read-back call goes to the StoredRun
        if name.startswith("get_") or name in ("acceptance_fraction", "backend"):
            return getattr(self.__dict__["_run"], name)
        raise AttributeError(name)

    def save_backend(self, name, discard=0):
        from astroemperor_b200.postproc import save_backend
        return save_backend(self._run, name, discard=discard)


def test_parent_run_and_postprocess_against_the_readback_api(ref):
    """The unmodified parent: run() -> (in-process engine run) -> _run_clean -> _load_sampler -> postprocess()."""
    emp, mg = ref
    from astroemperor_b200 import reference_plugin
    reference_plugin.install(engine_factory=_OracleEngine, sampler_factory=_OracleSampler)
    workdir = tempfile.mkdtemp(prefix="emp_plugin_")
    cwd = os.getcwd()
    os.chdir(workdir)
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            cfg = mg._cfg(1, 0, setup=(3, 24, 40, 2), conds=[("Period 1", "limits", [3, 5]),
                                                             ("Amplitude 1", "limits", [45, 60])])
            sim, k = _simulation(emp, mg, workdir, "51Peg", cfg)
            sim.run_config["burnin"] = 0.5
            sim.k_start = k
            sim._autorun_add_blocks()
            sim._stat_holder_update()
            sim.run()
            assert sim.sampler is not None and sim.reddemcee_discard == 40   # burnin 0.5 x 40 sweeps x 2 steps
            sim.postprocess()
    finally:
        os.chdir(cwd)
    # postprocess() consumed the chain of the cold temperature: 80 samples - 40 discarded, 24 walkers, flat
    assert sim.chain[0].shape == (40 * 24, sim.model.ndim__)
    assert np.isfinite(sim.like_max) and np.isfinite(sim.post_max) and np.isfinite(sim.BIC)
    assert sim.fit_max.shape == (sim.model.ndim__,) and np.all(np.isfinite(sim.sigmas))
    assert np.isfinite(sim.evidence[0])
    # the backend files of the generated script's save section, where the parent expects them
    back = os.path.join(sim.saveplace, "restore", "backends")
    assert any(f.startswith(sim.backend_name) for f in os.listdir(back))
    # temp_like_func is what the parent's DIC uses (emp.py:1202)
    assert np.isfinite(sim.temp_like_func(sim.fit_mean))

"""Install the B200 path INSIDE the unmodified reference front end.

BASELINE.json north_star: "Host code stays Python behind the existing `astroemperor.Simulation` /
`set_engine('reddemcee')` / `autorun` API, so the GPU path is a drop-in for the generated likelihood
callable the engine invokes."  The reference dispatches every engine-specific action on
`self.engine__.__name__` — method suffixes (`_run_engine_<name>`, `_load_sampler_<name>`,
`_prepare_run_<name>`, `_postprocess_setup_<name>`, `_postprocess_set_samples_<name>`, `_get_fit_<name>`,
`_clear_samples_<name>`; emp.py:168-179, 774, 1369, 1398, 1464, 2509, 2556, 2691) AND some thirty literal
`== 'reddemcee'` tests (plots, run tables, backend housekeeping: emp.py:425-765, 958, 1078, 1940, 1967, 2598,
2680).  A new engine NAME would silently skip all of the latter, so the plug-in keeps the name:

    import astroemperor_b200.reference_plugin as b200
    b200.install()                       # before or after `import astroemperor`
    import astroemperor as emp
    sim = emp.Simulation()
    sim.set_engine('reddemcee')          # unchanged user code from here on (tests/00_mini_test.py)
    sim.engine_config['setup'] = [8, 128, 512, 1]
    sim.load_data('51Peg'); sim.autorun(1, 1)

`install()` does three things, nothing else of the reference is touched:
  1. `sys.modules['reddemcee']` becomes a shim module (name 'reddemcee') whose `PTSampler` is
     `astroemperor_b200.sampler.PTSampler` and whose `hdf.PTHDFBackend / HDFBackend_plus` read what
     `postproc.save_backend` wrote — what `set_engine('reddemcee')` imports (emp.py:2360-2367);
  2. `Simulation._run_engine_reddemcee` (emp.py:2559-2582: write the script, `os.system('ipython ...')`) is
     replaced by an in-process run: model descriptor from the refreshed `ReddModel` (no text generation), the
     data frame the script would have read back from `temp_data.csv`, `LikelihoodEngine` + `PTSampler` on the GPU,
     `run_mcmc` exactly like `support/endit_reddemcee.scr` / `endit_freeze1.scr`, then the backend files the
     generated script's save section writes (emp.py:722-762) straight into `restore/backends/`;
  3. `Simulation._load_sampler_reddemcee` (emp.py:777-807: re-import the script, rebuild the sampler over the HDF5
     files) keeps the live sampler and points `temp_model_func / temp_like_func / temp_prior_func` at the engine;
     `_run_clean` no longer tries to move a script / backends that were never written to the working directory.
The parent's post-processing (emp.py:1324-1447, 1966-1990) then runs unchanged against the sampler's read-back API.
"""
from __future__ import annotations

import gc
import os
import sys
import types

import numpy as np

_STATE = {"installed": False, "options": {}}


# ---- what the generated script would have loaded (emp_model.py:406-433, 610-702) -----------------------
def _csv_round_trip(df):
    """The generated script does not see the parent's DataFrames but what `pd.read_csv` makes of their
    `to_csv` dump (emp_model.py:337-342, 406-411, 616-620); pandas' default float parser is not exactly
    round-tripping, so the hand-off goes through the same text to give the engine the same bits."""
    import io
    import pandas as pd
    return pd.read_csv(io.StringIO(df.to_csv()), index_col=0)


def rv_arrays_from_simulation(sim):
    """X_, Y_, YERR_, Flag (+ the SAI columns) as the generated script reads them back from temp_data.csv."""
    d = sim.model.data if getattr(sim.model, "data", None) is not None else sim.my_data
    d = _csv_round_trip(d)
    t, y, e = d["BJD"].values, d["RV"].values, d["eRV"].values
    flag = d["Flag"].values.astype(np.int32)
    sai = d.iloc[:, 4:].values if d.shape[1] > 4 and int(np.sum(getattr(sim.model, "cornums", [0]))) > 0 else None
    return (np.ascontiguousarray(t, dtype=np.float64), np.ascontiguousarray(y, dtype=np.float64),
            np.ascontiguousarray(e, dtype=np.float64), flag, sai)


def am_arrays_from_reddmodel(model):
    """The Hipparcos-Gaia constants of `_write_data_AM` (emp_model.py:610-702) straight from the ReddModel
    attributes `write_model` dumps to files (emp_model.py:339-365)."""
    cols = ["ref_epoch", "ra", "dec", "parallax", "pmra", "pmdec", "radial_velocity"]
    hipp, gost = _csv_round_trip(model.AM_hipp), _csv_round_trip(model.AM_gost)
    return dict(catalogs=_csv_round_trip(model.AM_hg123)[cols].values,
                time_hipp=hipp["BJD"].values, cpsi_hipp=hipp["CPSI"].values, spsi_hipp=hipp["SPSI"].values,
                epoch_hipp=hipp["EPOCH"].values, parf_hipp=hipp["PARF"].values, res_hipp=hipp["RES"].values,
                sres_hipp=hipp["SRES"].values,
                time_gost=gost["BJD"].values, cpsi_gost=gost["CPSI"].values, spsi_gost=gost["SPSI"].values,
                parf_gost=gost["parf"].values,
                mask_gdr2=np.asarray(model.mask_GDR2, dtype=bool), mask_gdr3=np.asarray(model.mask_GDR3, dtype=bool),
                gsv2=model.AM_GSV["GDR2"], gsv3=model.AM_GSV["GDR3"], inv_cov=model.AM_inv_COV,
                log_det_cov=model.AM_log_det_COV, astro_gost=_csv_round_trip(model.AM_astro).values,
                common_t=np.float64(model.common_t))


# ---- the engine run: replaces write_script() + os.system('ipython temp_script.py') ----------------------
def _run_engine_reddemcee(self):
    """Simulation._run_engine_reddemcee on the GPU (replaces emp.py:2559-2582)."""
    from .modelspec import spec_from_reddmodel
    opt = _STATE["options"]
    ntemps, nwalkers, nsweeps, nsteps = (int(x) for x in self.engine_config["setup"])
    assert self.nins__ == len(self.instrument_names_RV), f"instrument_names should have {self.nins__} items"
    cfg = self.engine_config
    if cfg["betas"] is not None:
        assert len(cfg["betas"]) == ntemps, f"betas should have {ntemps} items"

    if self.backend_name is None:                     # _set_backends, emp.py:557-558
        self.backend_name = f"{self.starname}_{self.saveplace_run}"
    spec = spec_from_reddmodel(self.model)            # the refreshed ReddModel (emp_model.py:223-330), no text
    t, y, yerr, flag, sai = rv_arrays_from_simulation(self)
    am = am_arrays_from_reddmodel(self.model) if getattr(self, "switch_AM", False) else None
    self.b200_spec = spec
    engine_factory = opt.get("engine_factory")
    if engine_factory is None:
        from .engine import LikelihoodEngine as engine_factory
    eng = engine_factory(spec, t, y, yerr, flag, am=am, sai=sai, device=opt.get("device", 0))

    sampler_factory = opt.get("sampler_factory")
    if sampler_factory is None:
        from .sampler import PTSampler as sampler_factory
    backend_dir = f"{self.saveplace}/restore/backends"
    sampler = sampler_factory(nwalkers, self.model.ndim__, eng, None, ntemps=ntemps, pool=None, backend=None,
                              betas=cfg["betas"], tsw_history=cfg["tsw_history"], smd_history=cfg["smd_history"],
                              adapt_tau=cfg["adapt_tau"], adapt_nu=cfg["adapt_nu"], adapt_mode=cfg["adapt_mode"],
                              **opt.get("sampler_kwargs", {}))
    sampler.D_ = np.asarray(spec.prior_widths(), dtype=np.float64)   # _set_sampler_D_, emp.py:595-602
    p1 = sampler.initial_positions(spec)                             # set_init() / test_init(), emp.py:617-684
    self.logger("Generating Samples", center=True, save=False, c="green", attrs=["reverse"])
    self.logger.line()
    progress = bool(cfg.get("progress", True))
    adapt_batches = self.run_config.get("adaptation_batches")
    adapt_nsweeps = self.run_config.get("adaptation_nsweeps")
    if adapt_batches and adapt_nsweeps:      # support/endit_freeze1.scr with the constants of emp.py:694-703
        nsweeps1 = int(self.reddemcee_discard)
        state = sampler.run_mcmc(p1, nsweeps=nsweeps1, nsteps=nsteps, progress=progress)
        sampler.select_adjustment("00")
        sampler.run_mcmc(state, nsweeps=nsweeps - nsweeps1, nsteps=nsteps, progress=progress)
    else:                                    # support/endit_reddemcee.scr
        sampler.run_mcmc(p1, nsweeps=nsweeps, nsteps=nsteps, progress=progress)
    self.sampler = sampler
    self.b200_engine = eng
    # the save section of the generated script (emp.py:722-762), written where _run_clean would have moved it
    if opt.get("save_backends", True) and hasattr(sampler, "save_backend"):
        os.makedirs(backend_dir, exist_ok=True)
        sampler.save_backend(f"{backend_dir}/{self.backend_name}")


def _run_clean(self):
    """emp.py:2592-2608 moves temp_script_*.py and the .h5 files out of the working directory; the in-process run
    never puts them there (the backends are written into restore/backends/ directly)."""
    if self.engine__.__name__ == "dynesty":
        return _STATE["orig"]["_run_clean"](self)
    gc.collect()


def _load_sampler_reddemcee(self):
    """emp.py:777-807 re-imports the generated script for my_model / my_likelihood / my_prior and rebuilds the
    sampler over the HDF5 files; here the sampler is still alive and the callables are the engine's."""
    eng = self.b200_engine
    self.temp_model_func = eng.my_model
    self.temp_like_func = eng.my_likelihood
    self.temp_prior_func = eng.my_prior
    if self.sampler is None:  # a restored run: read the backends back
        from .postproc import load_backend
        self.sampler = load_backend(f"{self.saveplace}/restore/backends/{self.backend_name}")
    self.betas = list(np.asarray(self.sampler.betas))
    self.engine_config["betas"] = self.betas
    self.engine_config["ntemps"] = len(self.betas)


# ---- the `reddemcee` module the reference imports -------------------------------------------------------
def _make_shim():
    from . import __version__ as ver
    from .postproc import load_backend, StoredRun

    shim = types.ModuleType("reddemcee")
    shim.__version__ = f"b200-{ver}"
    shim.__doc__ = "astroemperor_b200 standing in for reddemcee (astroemperor_b200.reference_plugin)"

    def PTSampler(nwalkers, ndim, log_like, log_prior=None, **kw):
        """reddemcee.PTSampler(nwalkers, ndim, my_likelihood, my_prior, ntemps=, pool=, backend=, betas=, ...)
        (emp.py:576-593).  `log_like` must be the LikelihoodEngine (its bound `my_likelihood` is accepted too);
        with `backend=` a PTHDFBackend reader this returns the stored run (emp.py:794-807)."""
        backend = kw.get("backend")
        if isinstance(backend, StoredRun):
            return backend
        from .sampler import PTSampler as _PT
        eng = getattr(log_like, "__self__", log_like)
        return _PT(nwalkers, ndim, eng, log_prior, **kw)

    hdf = types.ModuleType("reddemcee.hdf")

    def PTHDFBackend(filename, *a, **k):
        """Reader of `<name>.h5` + `<name>_<t>.h5` (emp.py:781-785) -> StoredRun."""
        name = filename[:-3] if filename.endswith(".h5") else filename
        return load_backend(name)

    def HDFBackend_plus(filename, *a, **k):   # per-temperature file: folded into PTHDFBackend's StoredRun
        return filename

    hdf.PTHDFBackend, hdf.HDFBackend_plus = PTHDFBackend, HDFBackend_plus
    shim.PTSampler, shim.hdf = PTSampler, hdf
    return shim, hdf


def install(device: int = 0, save_backends: bool = True, engine_factory=None, sampler_factory=None,
            **sampler_kwargs):
    """Route `astroemperor.Simulation`'s reddemcee engine through the B200 path (see module docstring).
    engine_factory / sampler_factory: test seams (default LikelihoodEngine / PTSampler); sampler_kwargs are
    passed to the sampler (e.g. seed=, store='host', thin_by=)."""
    _STATE["options"] = dict(device=device, save_backends=save_backends, engine_factory=engine_factory,
                             sampler_factory=sampler_factory, sampler_kwargs=sampler_kwargs)
    shim, hdf = _make_shim()
    sys.modules["reddemcee"] = shim
    sys.modules["reddemcee.hdf"] = hdf
    import astroemperor.emp as E   # the unmodified reference
    S = E.Simulation
    if not _STATE["installed"]:
        _STATE["orig"] = {k: getattr(S, k) for k in ("_run_engine_reddemcee", "_run_clean", "_load_sampler_reddemcee")}
    S._run_engine_reddemcee = _run_engine_reddemcee
    S._run_clean = _run_clean
    S._load_sampler_reddemcee = _load_sampler_reddemcee
    _STATE["installed"] = True
    return shim


def uninstall():
    """Restore the reference's own methods (the `reddemcee` shim stays importable)."""
    if _STATE["installed"]:
        import astroemperor.emp as E
        for k, v in _STATE["orig"].items():
            setattr(E.Simulation, k, v)
        _STATE["installed"] = False

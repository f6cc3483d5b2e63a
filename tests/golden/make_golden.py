#!/usr/bin/env python3
"""Generate the golden fixtures under tests/golden/ from the REAL reference.

Runs only in the build container (needs /root/reference, which does not exist on
the GPU box).  For every case below it

  1. imports the unmodified reference package from /root/reference/src with
     throw-away stubs for the presentation-only imports that are not installed
     here (termcolor, matplotlib, reddcolors, astropy.time.Time, reddemcee's
     name/version) — SURVEY.md §8c row C1;
  2. drives the real front end (`Simulation.load_data / set_engine /
     add_condition / _autorun_add_blocks / _prepare_run / write_script`) so the
     real generator emits `temp_script_*.py`;
  3. executes the generated script's model / likelihood / prior section (the part
     before the multiprocessing pool and the sampler are created) with
     `kepler` = oracle/kepler_shim.py (kepler.py itself is not installable);
  4. evaluates `my_likelihood`, `my_prior` (and `my_model` for one theta) on
     seeded parameter vectors and stores inputs + outputs + the model
     description (`astroemperor_b200.modelspec.spec_from_reddmodel`) as
     `tests/golden/<case>.npz` + `<case>.json`.

Nothing from the reference's sources is copied into the repository: the
fixtures hold numbers only.

Usage:  python tests/golden/make_golden.py [case ...]
"""
import io
import os
import shutil
import sys
import tempfile
import types
import contextlib

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference"
OUT = os.path.join(REPO, "tests", "golden")
sys.path.insert(0, REPO)


# ------------------------------------------------------------------ stubs ----
def _install_stubs():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    mod("termcolor", colored=lambda s, *a, **k: s)

    class _Anything:
        def __init__(self, *a, **k):
            pass

        def __getattr__(self, name):
            return _Anything()

        def __call__(self, *a, **k):
            return _Anything()

        def __iter__(self):
            return iter(())

        def __getitem__(self, k):
            return _Anything()

    mpl = mod("matplotlib", use=lambda *a, **k: None, rcParams={}, __version__="0", rc=lambda *a, **k: None)
    for sub in ("pyplot", "colors", "gridspec", "ticker", "patches", "lines", "cm", "collections"):
        m = mod(f"matplotlib.{sub}")
        m.__getattr__ = lambda name: _Anything()
        setattr(mpl, sub, m)
    rc = mod("reddcolors")
    rc.__getattr__ = lambda name: _Anything()
    mod("corner").__getattr__ = lambda name: _Anything()
    mod("arviz").__getattr__ = lambda name: _Anything()
    mod("reddemcee", __name__="reddemcee", __version__="0.9.8-stub")
    mod("emcee", __version__="3.1.6-stub")

    class Time:  # astropy.time.Time: only used for 1991.25 -> JD (qol_utils.py:210-211)
        def __init__(self, val, format=None):
            assert format == "decimalyear" and float(val) == 1991.25, (val, format)

        def to_value(self, fmt):
            assert fmt == "jd"
            return 2448348.75

    ap = mod("astropy")
    ap.time = mod("astropy.time", Time=Time)

    from oracle import kepler_shim
    sys.modules["kepler"] = kepler_shim


# ------------------------------------------------------ reference driver ----
def _write_vels(path, t, rv, erv):
    np.savetxt(path, np.column_stack([t, rv, erv]), fmt="%.17g")


def _run_generator(workdir, starname, configure, pre=None):
    """Returns (sim, namespace of the executed model/likelihood/prior section)."""
    import astroemperor as emp  # the real reference

    cwd = os.getcwd()
    os.chdir(workdir)
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            sim = emp.Simulation()
            sim.read_loc = workdir + "/"
            sim.save_loc = workdir + "/"
            if pre is not None:
                pre(sim)  # switches that load_data itself reads (switch_SA)
            sim.load_data(starname)
            sim.set_engine("reddemcee")
            k = configure(sim)
            sim.k_start = k
            sim._autorun_add_blocks()
            sim._prepare_run()
            sim.write_script()
        text = open(os.path.join(workdir, sim.temp_script)).read()
        # keep everything up to (not including) the multiprocessing pool section
        cut = text.index("import multiprocessing")
        head = text[:cut]
        # the script reads its data with paths relative to the cwd
        ns = {"__name__": "golden_script"}
        with contextlib.redirect_stdout(io.StringIO()):
            exec(compile(head, sim.temp_script, "exec"), ns)
        return sim, ns
    finally:
        os.chdir(cwd)


def _draw_thetas(spec, rng, n, widen=0.02):
    """Seeded parameter vectors inside (and a few just outside) the prior box."""
    fp = spec.free_params()
    lo = np.array([p.limits[0] for p in fp], dtype=float)
    hi = np.array([p.limits[1] for p in fp], dtype=float)
    th = rng.uniform(lo, hi, size=(n, len(fp)))
    # a few rows leave the box in one coordinate -> prior -inf
    for r in range(0, n, 7):
        j = rng.integers(len(fp))
        th[r, j] = hi[j] + widen * (hi[j] - lo[j]) if rng.random() < 0.5 else lo[j] - widen * (hi[j] - lo[j])
    return th


def _stage_star(workdir, star):
    dst = os.path.join(workdir, "datafiles", star)
    shutil.copytree(os.path.join(REF, "tests", "datafiles", star), dst)
    return dst


def _synthetic_star(workdir, star, seed, n, nins, kplan, ma=False, sai=None):
    """SURVEY.md §8d row D2 generator (same code as astroemperor_b200.synth).  `sai` = activity-index
    columns per instrument: extra columns after eRV (qol_utils.py:74-79), correlated with the RVs."""
    from astroemperor_b200.synth import add_activity_columns, make_synthetic_rv
    files = make_synthetic_rv(seed=seed, n=n, nins=nins, kplan=kplan, ma=ma)
    files = add_activity_columns(files, sai if sai else [0] * nins, seed)
    d = os.path.join(workdir, "datafiles", star, "RV")
    os.makedirs(d)
    for i, (t, rv, erv, act) in enumerate(files):
        np.savetxt(os.path.join(d, f"{star}_ins{i + 1}.vels"), np.column_stack([t, rv, erv, act]), fmt="%.17g")


# ------------------------------------------------------------------ cases ----
def _cfg(k, param=0, setup=(2, 100, 500, 1), conds=(), moav=None, acc=0, jitter=True, names=None,
         extra=None):
    def configure(sim):
        sim.engine_config["setup"] = list(setup)
        sim.keplerian_parameterisation = param
        if moav is not None:
            sim.moav = dict(moav)
        sim.acceleration = acc
        sim.switch_jitter = jitter
        if names is not None:
            sim.instrument_names_RV = list(names)
        for c in conds:
            sim.add_condition(list(c))
        if extra is not None:
            extra(sim)
        return k
    return configure


CASES = {
    # BASELINE config 1: README mini test / quickstart (51Peg, K=1, parameterisation 0)
    "c1_51peg_k1_p0": dict(star="51Peg", cfg=_cfg(1, 0), n_theta=96),
    # Kepler-free known-answer case: notebooks print logL(-0.153, 36.063) = -1308.795
    "c1_51peg_k0": dict(star="51Peg", cfg=_cfg(0, 0), n_theta=32),
    # tests/00_mini_test.py: parameterisation 1 + conditions
    "mini_51peg_k1_p1": dict(star="51Peg", cfg=_cfg(1, 1, conds=[
        ("Period 1", "limits", [3, 5]), ("Amplitude 1", "limits", [45, 60]),
        ("Offset 1", "limits", [-10., 10.]), ("Period 1", "init_pos", [4.1, 4.3]),
        ("Amplitude 1", "init_pos", [50, 60])], names=["LICK"]), n_theta=96),
    # every other Keplerian template on the 2-instrument synth fixture (80 points)
    "synth_k2_p2": dict(star="synth", cfg=_cfg(2, 2), n_theta=64),
    "synth_k2_p3": dict(star="synth", cfg=_cfg(2, 3), n_theta=64),
    "synth_k2_p4": dict(star="synth", cfg=_cfg(2, 4), n_theta=64),
    "synth_k2_p6": dict(star="synth", cfg=_cfg(2, 6), n_theta=64),
    "synth_k2_p7": dict(star="synth", cfg=_cfg(2, 7), n_theta=64),
    # acceleration order 2 + a fixed parameter (A_/mod_fixed_ path)
    "synth_k1_p0_acc2_fixed": dict(star="synth", cfg=_cfg(1, 0, acc=2, conds=[
        ("Eccentricity 1", "fixed", 0.1)]), n_theta=64),
    # default per-instrument MA template (a no-op on logL in the reference)
    "synth_k3_p1_ma1_perins": dict(star="synth", cfg=_cfg(3, 1, moav={"order": 1, "global": False}),
                                   n_theta=48),
    # global MA recurrence, orders 1 and 2
    "synth_k1_p0_ma1_global": dict(star="synth", cfg=_cfg(1, 0, moav={"order": 1, "global": True}),
                                   n_theta=48),
    "synth_k1_p1_ma2_global": dict(star="synth", cfg=_cfg(1, 1, moav={"order": 2, "global": True}),
                                   n_theta=48),
    # periodic blocks (sinusoid00.model, magneticcycle00.model); with a global MA block they are
    # added AFTER the MA residuals were formed (block order of emp.py:2628-2651)
    "synth_k1_p0_sinusoid": dict(star="synth", cfg=_cfg(1, 0, extra=lambda sim: setattr(sim, "sinusoid", 1)),
                                 n_theta=48),
    "synth_k1_p1_magcycle_ma1_global": dict(star="synth", cfg=_cfg(
        1, 1, moav={"order": 1, "global": True},
        extra=lambda sim: (setattr(sim, "magnetic_cycle", 1), setattr(sim, "sinusoid", 2))), n_theta=48),
    # stellar-activity indices (sai00.model): 2 + 1 columns on 2 instruments; second case with a global MA
    # block and a sinusoid, which fixes the reference's order MOAV -> activity -> periodic
    "synth_k1_p0_sai21": dict(synth=dict(seed=9, n=120, nins=2, kplan=1, sai=[2, 1]), cfg=_cfg(1, 0),
                              pre=lambda sim: setattr(sim, "switch_SA", True), n_theta=48),
    "synth_k2_p1_sai03_ma1_global_sin": dict(
        synth=dict(seed=10, n=150, nins=2, kplan=2, ma=True, sai=[0, 3]),
        cfg=_cfg(2, 1, moav={"order": 1, "global": True}, extra=lambda sim: setattr(sim, "sinusoid", 1)),
        pre=lambda sim: setattr(sim, "switch_SA", True), n_theta=48),
    # no jitter block
    "synth_k1_p0_nojit": dict(star="synth", cfg=_cfg(1, 0, jitter=False), n_theta=32),
    # GJ876: 8 instruments, 770 points
    "gj876_k2_p1": dict(star="GJ876", cfg=_cfg(2, 1), n_theta=32),
    # BASELINE config 2 shape at reduced N (generated synthetic, 3 planets / 2 instruments)
    "c2_synth3p_2ins_n400": dict(synth=dict(seed=2, n=400, nins=2, kplan=3), cfg=_cfg(3, 1), n_theta=48),
    # BASELINE config 4 shape at reduced N: 5 planets, 4 instruments, MA(1) both modes
    "c4_synth5p_4ins_ma_noop_n600": dict(synth=dict(seed=4, n=600, nins=4, kplan=5, ma=True),
                                          cfg=_cfg(5, 0, moav={"order": 1, "global": False}), n_theta=32),
    "c4_synth5p_4ins_ma_global_n600": dict(synth=dict(seed=4, n=600, nins=4, kplan=5, ma=True),
                                            cfg=_cfg(5, 0, moav={"order": 1, "global": True}), n_theta=32),
    # BASELINE config 3: HIP21850 joint RV + Hipparcos-Gaia astrometry, 2 Keplerians
    "c3_hip21850_am_k2": dict(star="HIP21850", cfg=_cfg(2, 0, acc=1), n_theta=48, am=True),
    "c3_hip21850_am_k1": dict(star="HIP21850", cfg=_cfg(1, 0, acc=1), n_theta=48, am=True),
}


def make_case(name, case):
    from astroemperor_b200.modelspec import spec_from_reddmodel
    rng = np.random.default_rng(abs(hash(name)) % (2 ** 32) if False else
                                int.from_bytes(name.encode(), "little") % (2 ** 32))
    workdir = tempfile.mkdtemp(prefix="emp_golden_")
    try:
        if "star" in case:
            star = case["star"]
            _stage_star(workdir, star)
        else:
            star = name
            _synthetic_star(workdir, star, **case["synth"])
        sim, ns = _run_generator(workdir, star, case["cfg"], pre=case.get("pre"))
        spec = spec_from_reddmodel(sim.model)
        assert spec.ndim == ns_ndim(sim), (spec.ndim, ns_ndim(sim))
        thetas = _draw_thetas(spec, rng, case["n_theta"])
        ll = np.empty(len(thetas))
        lp = np.empty(len(thetas))
        with np.errstate(all="ignore"):
            for i, th in enumerate(thetas):
                lp[i] = ns["my_prior"](th)
                ll[i] = float(ns["my_likelihood"](th)) if np.isfinite(lp[i]) else -np.inf
            good = int(np.argmax(np.where(np.isfinite(lp), ll, -np.inf)))
            model0, err20 = ns["my_model"](thetas[good])
        out = dict(t=ns["X_"], y=ns["Y_"], yerr=ns["YERR_"],
                   flag=ns["my_data"]["Flag"].values.astype(np.int32),
                   thetas=thetas, logl=ll, logp=lp, model_theta_index=np.int64(good),
                   model0=np.asarray(model0, dtype=np.float64), err20=np.asarray(err20, dtype=np.float64),
                   D_=np.diff(np.array(sim.model.get_attr_param("limits", flat=True), dtype=float)[
                       sim.model.C_]).flatten(),
                   common_t=np.float64(sim.my_data_common_t))
        n_sai = int(np.sum(getattr(sim, "cornums", [0])))
        if n_sai:
            out["sai"] = np.column_stack([ns[f"SAI{j + 1}_"] for j in range(n_sai)]).astype(np.float64)
        if case.get("am"):
            from astroemperor_b200.amdata import am_arrays_from_namespace
            am = am_arrays_from_namespace(ns)
            out.update({f"am_{k}": v for k, v in am.items()})
            # the longdouble result, split so float64 fixtures keep its extra bits
            with np.errstate(all="ignore"):
                ll_am = np.array([ns["loglike_AM"](th) if np.isfinite(p) else -np.inf
                                  for th, p in zip(thetas, lp)], dtype=np.longdouble)
            out["logl_am_hi"] = ll_am.astype(np.float64)
            out["logl_am_lo"] = (ll_am - out["logl_am_hi"].astype(np.longdouble)).astype(np.float64)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
        with open(os.path.join(OUT, name + ".json"), "w") as f:
            f.write(spec.to_json())
        n_ok = int(np.isfinite(lp).sum())
        print(f"{name:36s} ndim={spec.ndim:3d} n={len(out['t']):5d} thetas={len(thetas)} finite={n_ok} "
              f"max logL={np.max(ll):.6f}")
    finally:
        shutil.rmtree(workdir, ignore_errors=True)


def ns_ndim(sim):
    return int(sim.model.ndim__)


def main(argv):
    _install_stubs()
    sys.path.insert(0, os.path.join(REF, "src"))
    names = argv or list(CASES)
    for name in names:
        make_case(name, CASES[name])


if __name__ == "__main__":
    main(sys.argv[1:])

"""Host-side mirror of the reference front end, for the hot path only.

`default_spec` restates what `Simulation._autorun_add_blocks` + `SmartSetter`
(emp.py:2614-2651, block_repo.py:505-867) produce for an RV model: the block
order (Keplerians first, emp.py:1235; then Acceleration, Offset, Jitter, MOAV),
the parameter names and the data-derived default limits / priors.  It is checked
against the REAL reference's output in tests/test_host_logic.py::test_default_spec_matches_reference_model via the golden
`<case>.json` files.

`Simulation` keeps the reference's user-facing names (`load_data`,
`set_engine('reddemcee')`, `engine_config['setup']`, `add_condition`,
`keplerian_parameterisation`, `moav`, `acceleration`, `autorun`) so the README
mini test reads the same; underneath, instead of generating a script and
spawning `ipython` (emp.py:2575-2582), it builds a `LikelihoodEngine` and a
device `PTSampler` in-process.  Post-processing, plotting, tables, the
model-comparison loop's GM/HDI machinery are out of scope (SURVEY.md §2 rows
12-18): `autorun` records max-likelihood / BIC per model and stops on the
reference's default criterion (BIC_old - BIC_new < 5, emp.py:1131-1133).
"""
from __future__ import annotations

import itertools
import os
from typing import List, Optional, Sequence

import numpy as np

from .data import RVData, load_rv_folder
from .modelspec import (AdditionalPrior, BlockSpec, ModelSpec, ParamSpec, finalize_prargs,
                        UnsupportedModelError)

TWO_PI_LIMITS = [0, 2 * np.pi]

# parameter names per Keplerian template (block_repo.py:22-31, param_repo.py)
KEP_PARAM_NAMES = {
    0: ["Period", "Amplitude", "Phase", "Eccentricity", "Longitude"],
    1: ["Period", "Amplitude", "Phase", "Ecc_sin", "Ecc_cos"],
    2: ["lPeriod", "Amp_sin", "Amp_cos", "Ecc_sin", "Ecc_cos"],
    3: ["Period", "Amplitude", "T_0", "Eccentricity", "Longitude"],
    4: ["Period", "Amplitude", "T_0", "Ecc_sin", "Ecc_cos"],
    6: ["lPeriod", "Amplitude", "Phase", "Eccentricity", "Longitude"],
    7: ["lPeriod", "Amplitude", "Phase", "Ecc_sin", "Ecc_cos"],
}
HOU_NAMES = ("Amp_sin", "Amp_cos", "Ecc_sin", "Ecc_cos")  # param_repo.py: is_hou=True


def _param(name, prior, limits, prargs):
    base = name.split(" ")[0]
    return ParamSpec(name=name, prior=prior, limits=[_f(limits[0]), _f(limits[1])], prargs=prargs,
                     is_hou=base in HOU_NAMES)


def _f(x):
    return x.item() if isinstance(x, (np.floating, np.integer)) else x


# Default (prior, limits) of every Keplerian parameter, by NAME.  Limits are symbolic: they are resolved against the
# data set by `_resolve` (the values are the ones SmartSetter.set_Keplerian derives, block_repo.py:517-607; the
# golden descriptors of tests/golden pin them for all eight parameterisations).
_KEP_DEFAULTS = {
    "Period": ("Uniform", (1.5, "span")),
    "lPeriod": ("Jeffreys", ("ln 1.5", "ln span")),
    "Amplitude": ("Uniform", (1e-6, "amp")),
    "Amp_sin": ("Uniform", ("-root_half_amp", "root_half_amp")),
    "Amp_cos": ("Uniform", ("-root_half_amp", "root_half_amp")),
    "Phase": ("Uniform", (0.0, "2pi")),
    "Longitude": ("Uniform", (0.0, "2pi")),
    "T_0": ("Uniform", (-1000, 1000)),
    "Eccentricity": ("Normal", "ecc"),          # limits / prargs come from the caller (sim.ecc_limits / ecc_prargs)
    "Ecc_sin": ("Uniform", (-1, 1)),
    "Ecc_cos": ("Uniform", (-1, 1)),
    "Inclination": ("Isotropic", (0, "pi")),
    "Omega": ("Uniform", (0.0, "2pi")),
}
# parameterisations whose eccentricity is derived (S^2 + C^2): prior of the derived quantity (emp.py:210-220)
_KEP_DERIVED_ECC = {1: "Uniform", 2: "Uniform", 4: "Uniform", 7: "Normal"}
# parameterisation 2 samples ln P with a plain Uniform prior (block_repo.py:553-563); 6 and 7 use 'Jeffreys'
_KEP_PRIOR_OVERRIDE = {(2, "lPeriod"): "Uniform"}


def _resolve(sym, data: RVData):
    """A symbolic limit -> number, from the data-derived scales of SmartSetter (RV scatter, time span)."""
    if not isinstance(sym, str):
        return sym
    span = data.t.max() - data.t.min()
    amp = np.std(data.y) * np.sqrt(4)  # data['RV'].std(ddof=0) * sqrt(4)
    table = {"span": span, "ln span": np.log(span), "ln 1.5": np.log(1.5), "amp": amp, "2pi": TWO_PI_LIMITS[1],
             "pi": np.pi, "root_half_amp": np.sqrt(amp / 2), "-root_half_amp": -np.sqrt(amp / 2)}
    return table[sym]


def keplerian_block(data: RVData, number: int, parameterisation: int, ecc_limits, ecc_prargs,
                    astrometry: bool = False) -> BlockSpec:
    """KeplerianBlock + SmartSetter.set_Keplerian (block_repo.py:22-78, 517-607) from the defaults table."""
    p = parameterisation
    if p not in KEP_PARAM_NAMES:
        raise UnsupportedModelError(f"keplerian_parameterisation {p}")
    names = list(KEP_PARAM_NAMES[p])
    if astrometry:
        if p != 0:
            raise UnsupportedModelError("astrometric Keplerians use parameterisation 0 (block_repo.py:524-541)")
        names += ["Inclination", "Omega"]
    params = []
    for n in names:
        prior, lim = _KEP_DEFAULTS[n]
        prior = _KEP_PRIOR_OVERRIDE.get((p, n), prior)
        prargs = None
        if lim == "ecc":
            lim, prargs = ecc_limits, list(ecc_prargs)
        params.append(_param(f"{n} {number}", prior, [_resolve(lim[0], data), _resolve(lim[1], data)], prargs))
    addi: List[AdditionalPrior] = []
    if p in _KEP_DERIVED_ECC:
        pr = _KEP_DERIVED_ECC[p]
        addi.append(AdditionalPrior("Ecc", pr, [0, 1], list(ecc_prargs) if pr == "Normal" else None))
    return BlockSpec(type_="Keplerian", params=params, parameterisation=p, astrometry=astrometry,
                     number=number, additional=addi)


def periodic_blocks(data: RVData, sinusoid=0, magnetic_cycle=0) -> List[BlockSpec]:
    """SinusoidBlock / MagneticCycleBlock + SmartSetter.set_Sinusoid / set_MagneticCycle
    (block_repo.py:81-107, 302-347, 804-833)."""
    amp = np.std(data.y) * np.sqrt(4)
    per = data.t.max() - data.t.min()
    blocks = []
    if sinusoid:
        names = [f"Period {sinusoid}", f"Amplitude {sinusoid}", f"Phase {sinusoid}"]
        lims = [[1.5, per], [1e-6, amp], TWO_PI_LIMITS]
        blocks.append(BlockSpec(type_="Sinusoid", number=sinusoid,
                                params=[_param(n, "Uniform", l, None) for n, l in zip(names, lims)]))
    if magnetic_cycle:
        names = ["Period S1", "Amplitude S1", "Amplitude S2", "Phase S1", "Phase S2"]
        lims = [[1.5, per], [1e-6, amp], [1e-6, amp], TWO_PI_LIMITS, TWO_PI_LIMITS]
        blocks.append(BlockSpec(type_="MagneticCycle", number=magnetic_cycle,
                                params=[_param(n, "Uniform", l, None) for n, l in zip(names, lims)]))
    return blocks


def instrument_blocks(data: RVData, acceleration=0, jitter=True, moav=None,
                      jitter_prargs=(5, 5)) -> List[BlockSpec]:
    """_autorun_add_RV_ins (emp.py:2628-2651) + SmartSetter.set_{Acceleration,Offset,Jitter,MOAV}."""
    nins = data.nins
    blocks = []
    if acceleration:
        ps = []
        for i in range(acceleration):
            name = "Acceleration" if i == 0 else f"Acceleration Order {i + 1}"
            ps.append(_param(name, "Uniform", [-1, 1], None))
        blocks.append(BlockSpec(type_="Acceleration", params=ps, number=acceleration))
    lims = []
    for nin in range(nins):
        m = data.flag == (nin + 1)
        lims.append(np.abs(data.y[m]).max())
    blocks.append(BlockSpec(type_="Offset", number=nins,
                            params=[_param(f"Offset {i + 1}", "Uniform", [-lims[i], lims[i]], None)
                                    for i in range(nins)]))
    if jitter:
        blocks.append(BlockSpec(type_="Jitter", number=nins,
                                params=[_param(f"Jitter {i + 1}", "Normal", [1e-5, lims[i]], list(jitter_prargs))
                                        for i in range(nins)]))
    if moav and moav.get("order", 0):
        order, is_global = int(moav["order"]), bool(moav.get("global", False))
        ps = []
        for i, j in itertools.product(range(1 if is_global else nins), range(order)):
            ps.append(_param(f"MACoefficient {i + 1} Order {j + 1}", "Uniform", [0.0, 0.3], None))
            ps.append(_param(f"MATimescale {i + 1} Order {j + 1}", "Uniform", [5, 25], None))
        blocks.append(BlockSpec(type_="MOAV", params=ps, number=nins, moav_order=order, moav_global=is_global))
    return blocks


def activity_block(data: RVData) -> List[BlockSpec]:
    """SAIBlock + SmartSetter.set_StellarActivity (block_repo.py:243-272, 742-749): one coefficient per
    activity column, named `Staract <instrument> <column>`, Uniform on [-1, 1]."""
    cornums = list(data.cornums or [])
    if not sum(cornums):
        return []
    ps = [_param(f"Staract {i + 1} {c + 1}", "Uniform", [-1., 1.], None)
          for i, n in enumerate(cornums) for c in range(n)]
    # number_: the model writer reuses it as the running column index (emp_model.py:740-745) and leaves the last
    return [BlockSpec(type_="StellarActivity", params=ps, number=sum(cornums), sai_counts=cornums)]


def astrometry_blocks() -> List[BlockSpec]:
    """add_offset_am / add_jitter_am (emp.py:1300-1311) + SmartSetter.set_Astrometry* (block_repo.py:835-850)."""
    off = [_param(n, "Uniform", [-1e1, 1e1], [])
           for n in ("Offset RA", "Offset DE", "Offset PLX", "Offset pm RA", "Offset pm DE")]
    jit = [_param("Jitter Hipp", "Uniform", [0, 20], []), _param("Jitter Gaia", "Uniform", [0, 10], [])]
    return [BlockSpec(type_="AstrometryOffset", params=off, number=5),
            BlockSpec(type_="AstrometryJitter", params=jit, number=2)]


def apply_conditions(spec: ModelSpec, conds: Sequence[Sequence]) -> None:
    """Simulation.apply_conditions (emp.py:2790-2806): [param_name, attribute, value]."""
    for b in spec.blocks:
        for p in b.params:
            for c in conds:
                if p.name == c[0]:
                    if c[1] not in ("limits", "prior", "prargs", "fixed", "init_pos"):
                        raise UnsupportedModelError(f"condition on attribute '{c[1]}'")
                    setattr(p, c[1], c[2])


def finalize(spec: ModelSpec) -> ModelSpec:
    """ReddModel.refresh__ (emp_model.py:223-330): fixed handling + prior normalisers."""
    for b in spec.blocks:
        for p in b.params:
            if p.fixed is not None:  # Parameter_Block._handle_fixed_params (emp_model.py:75-80)
                p.prior, p.limits = "Fixed", [float("nan"), float("nan")]
            else:
                p.prargs = finalize_prargs(p.prior, p.limits, p.prargs)
        for a in b.additional:
            a.prargs = finalize_prargs(a.prior, a.limits, a.prargs)
    return spec


def default_spec(data: RVData, kplan: int, parameterisation: int = 0, acceleration: int = 0,
                 jitter: bool = True, moav: Optional[dict] = None, conditions: Sequence = (),
                 eccentricity_limits=(0, 1), eccentricity_prargs=(0, 0.1), jitter_prargs=(5, 5),
                 astrometry: bool = False, sinusoid: int = 0, magnetic_cycle: int = 0) -> ModelSpec:
    blocks = [keplerian_block(data, k + 1, parameterisation, list(eccentricity_limits),
                              list(eccentricity_prargs), astrometry=astrometry) for k in range(kplan)]
    blocks += instrument_blocks(data, acceleration, jitter, moav, jitter_prargs)
    blocks += activity_block(data)                              # after MOAV, before the periodic blocks
    blocks += periodic_blocks(data, sinusoid, magnetic_cycle)  # (emp.py:2636-2650)
    if astrometry:
        blocks += astrometry_blocks()
    spec = ModelSpec(blocks=blocks, nins=data.nins)
    apply_conditions(spec, conditions)
    return finalize(spec)


class Simulation:
    """The slice of `astroemperor.Simulation` (emp.py:2240-2917) that leads to the hot path."""

    def __init__(self):
        self.starname = None
        self.read_loc = ""
        self.instrument_names_RV = None
        self.starmass = 1.0
        self.keplerian_parameterisation = 0
        self.acceleration = 0
        self.switch_jitter = True
        self.switch_SA = False  # read before load_data, like the reference (emp.py:2307)
        self.moav = {"order": 0, "global": False}
        self.sinusoid = 0
        self.magnetic_cycle = 0
        self.eccentricity_limits = [0, 1]
        self.eccentricity_prargs = [0, 0.1]
        self.jitter_prargs = [5, 5]
        self.conds = []
        self.cores__ = None  # accepted and ignored: walkers are evaluated on the GPU
        self.engine__ = None
        self.engine_config = {}
        self.run_config = {}
        self.device = 0
        self.seed = None
        self.data: Optional[RVData] = None
        self.am_data = None
        self.sampler = None
        self.model: Optional[ModelSpec] = None
        self.history = []  # one dict per model run by autorun
        self.BIC = np.inf

    # -- reference API --------------------------------------------------------------
    def load_data(self, folder_name: str):
        """`datafiles/<star>/RV/*.vels` under read_loc (qol_utils.py:16-100)."""
        self.starname = folder_name
        base = os.path.join(f"{self.read_loc}datafiles", folder_name)
        self.data = load_rv_folder(os.path.join(base, "RV") + os.sep, switch_SA=self.switch_SA)
        am_dir = os.path.join(base, "AM")
        if os.path.isdir(am_dir) and os.listdir(am_dir):
            from .amdata import load_am_folder
            self.am_data = load_am_folder(am_dir + os.sep, common_t=self.data.common_t)

    def set_engine(self, eng: str):
        if eng != "reddemcee":
            raise UnsupportedModelError(f"engine '{eng}': only the reddemcee PT path is on the device")
        self.engine__ = "reddemcee"
        # defaults of Simulation._set_engine_reddemcee (emp.py:2370-2388)
        self.engine_config = {"setup": np.array([5, 100, 500, 2]), "ntemps": 5, "betas": None, "moves": None,
                              "tsw_history": True, "smd_history": True, "adapt_tau": 1000, "adapt_nu": 1,
                              "adapt_mode": 0, "progress": True}
        self.run_config = {"burnin": None, "adaptation_batches": None, "adaptation_nsweeps": None,
                           "thin": 1, "discard": 0.5, "logger_level": "CRITICAL"}

    def add_condition(self, cond):
        self.conds.append(list(cond))

    def build_model(self, kplan: int) -> ModelSpec:
        if self.data is None:
            raise RuntimeError("load_data() first")
        return default_spec(self.data, kplan, self.keplerian_parameterisation, self.acceleration,
                            self.switch_jitter, self.moav if self.moav.get("order") else None, self.conds,
                            self.eccentricity_limits, self.eccentricity_prargs, self.jitter_prargs,
                            astrometry=self.am_data is not None, sinusoid=self.sinusoid,
                            magnetic_cycle=self.magnetic_cycle)

    def run(self, kplan: int):
        """_run_engine_reddemcee (emp.py:2559-2582) without the script / child process."""
        from .engine import LikelihoodEngine
        from .sampler import PTSampler
        if self.engine__ != "reddemcee":
            raise RuntimeError("set_engine('reddemcee') first")
        ntemps, nwalkers, nsweeps, nsteps = [int(v) for v in self.engine_config["setup"]]
        if self.engine_config["betas"] is not None:
            assert len(self.engine_config["betas"]) == ntemps, f"betas should have {ntemps} items"
        self.model = self.build_model(kplan)
        d = self.data
        eng = LikelihoodEngine(self.model, d.t, d.y, d.yerr, d.flag, am=self.am_data, device=self.device,
                               sai=d.sai)
        cfg = self.engine_config
        self.sampler = PTSampler(nwalkers, self.model.ndim, eng, ntemps=ntemps, betas=cfg["betas"],
                                 tsw_history=cfg["tsw_history"], smd_history=cfg["smd_history"],
                                 adapt_tau=cfg["adapt_tau"], adapt_nu=cfg["adapt_nu"],
                                 adapt_mode=cfg["adapt_mode"], seed=self.seed)
        self.sampler.D_ = self.model.prior_widths()  # emp.py:595-602
        p1 = self.sampler.initial_positions(self.model)
        progress = bool(cfg.get("progress"))
        rc = self.run_config
        if rc.get("adaptation_batches") and rc.get("adaptation_nsweeps"):
            # two-phase run of support/endit_freeze1.scr: `reddemcee_discard` adaptation sweeps
            # (_prepare_run_reddemcee, emp.py:2512-2520; _set_run overrides adaptation_nsweeps with it,
            # emp.py:699-701), then the ladder is frozen for the rest
            niter = nsteps * nsweeps
            burn = rc.get("burnin")
            discard = int(burn * niter) if isinstance(burn, float) else (int(burn) if isinstance(burn, int) else 0)
            discard = min(discard, nsweeps)
            self.sampler.run_mcmc(p1, nsweeps=discard, nsteps=nsteps, progress=progress)
            self.sampler.select_adjustment("00")
            self.sampler.run_mcmc(None if discard else p1, nsweeps=nsweeps - discard, nsteps=nsteps, progress=progress)
        else:
            self.sampler.run_mcmc(p1, nsweeps=nsweeps, nsteps=nsteps, progress=progress)
        return self.sampler

    def autorun(self, k_start=0, k_end=10):
        """autorun loop (emp.py:2414-2460) reduced to: run k Keplerians, record max logL and
        BIC = ndim ln n - 2 lnL_max (emp.py:1195), stop when BIC stops improving by > 5."""
        assert k_start <= k_end, f"Invalid keplerian starting point: ({k_start}, {k_end})"
        bic_old = np.inf
        for k in range(k_start, k_end + 1):
            s = self.run(k)
            ll = s.get_log_like(flat=True)[0]
            like_max = float(np.max(ll))
            bic = self.model.ndim * np.log(len(self.data)) - 2 * like_max
            self.history.append(dict(k=k, ndim=self.model.ndim, like_max=like_max, BIC=bic))
            self.BIC = bic
            if not (bic_old - bic > 5):
                break
            bic_old = bic
        return self.history

"""Model description handed from the host front end to the device path.

In the reference the model reaches the sampler as *generated Python text*
(`ReddModel.write_model`, emp_model.py:333-403, and `emp_scribe.write_script`,
emp.py:137-175).  Here the same information travels as data: a `ModelSpec`
(blocks -> parameters, the object graph `Simulation.blocks__` holds) that is
compiled into the flat `EmpModelDesc` of include/emperor_b200.h.

`spec_from_reddmodel` reads a *refreshed* reference `ReddModel`
(emp_model.py:223-330) by duck typing, so a maintainer can hand the unmodified
front end's model straight to this package.
"""
from __future__ import annotations

import ctypes
import json
import math
from dataclasses import dataclass, field, asdict
from typing import List, Optional, Sequence

import numpy as np

# ---- limits of the C descriptor (include/emperor_b200.h) --------------------
EMP_ABI_VERSION = 6
EMP_MAX_KEP = 10
EMP_MAX_INS = 16
EMP_MAX_DIM = 128
EMP_MAX_ACC = 4
EMP_MAX_MA = 4
EMP_MAX_PERIODIC = 4
EMP_MAX_SAI = 4
EMP_MAX_PRIOR_OPS = 2 * EMP_MAX_DIM + 2 * EMP_MAX_KEP + 8

PRIOR_KINDS = {"Uniform": 0, "Normal": 1, "Jeffreys": 2, "Isotropic": 3, "Fixed": 4}
# support/priors/{Beta,GaussianMixture,Hill}.prior exist in the reference but are
# not in any BASELINE config; the device path rejects them instead of ignoring them.
UNSUPPORTED_PRIORS = ("Beta", "GaussianMixture", "Hill")

POP_PARAM, POP_CHECK, POP_SUMSQ = 0, 1, 2
MA_NONE, MA_REFERENCE_NOOP, MA_GLOBAL = 0, 1, 2

# parameters per Keplerian template (support/models/kep0*.model, akep00.model)
KEP_NPAR = {0: 5, 1: 5, 2: 5, 3: 5, 4: 5, 5: 7, 6: 5, 7: 5}
# templates whose eccentricity is S**2 + C**2 (slots 3, 4)
KEP_SC = (1, 2, 4, 7)


class UnsupportedModelError(NotImplementedError):
    """The reference can express this model but the device path cannot."""


@dataclass
class ParamSpec:
    name: str
    prior: str = "Uniform"
    limits: Sequence[float] = (float("nan"), float("nan"))
    prargs: object = None  # Uniform/Isotropic: logZ | Normal: [mu, s, logZ] | raw [mu, s]
    fixed: Optional[float] = None
    init_pos: Sequence[Optional[float]] = (None, None)
    is_hou: bool = False


@dataclass
class AdditionalPrior:
    """Derived-quantity prior of S/C parameterisations (emp.py:210-246)."""
    kind: str  # 'Ecc' (slots 3,4) or 'Amp' (slots 1,2)
    prior: str = "Uniform"
    limits: Sequence[float] = (0.0, 1.0)
    prargs: object = None


@dataclass
class BlockSpec:
    type_: str  # Keplerian | Acceleration | Offset | Jitter | MOAV | AstrometryOffset | AstrometryJitter
    params: List[ParamSpec] = field(default_factory=list)
    parameterisation: Optional[int] = None  # Keplerian only
    astrometry: bool = False  # Keplerian only: akep00.model (7 parameters)
    number: int = 0  # nins for Offset/Jitter/MOAV, order for Acceleration
    moav_order: int = 0
    moav_global: bool = False
    sai_counts: List[int] = field(default_factory=list)  # StellarActivity only: columns per instrument (cornums)
    additional: List[AdditionalPrior] = field(default_factory=list)


@dataclass
class ModelSpec:
    blocks: List[BlockSpec]
    nins: int = 1

    # -- (de)serialisation used by the golden fixtures ---------------------
    def to_json(self) -> str:
        return json.dumps(asdict(self), indent=1)

    @staticmethod
    def from_json(text: str) -> "ModelSpec":
        d = json.loads(text)
        blocks = []
        for b in d["blocks"]:
            params = [ParamSpec(**p) for p in b.pop("params")]
            addi = [AdditionalPrior(**a) for a in b.pop("additional")]
            blocks.append(BlockSpec(params=params, additional=addi, **b))
        return ModelSpec(blocks=blocks, nins=d["nins"])

    # -- views -------------------------------------------------------------
    @property
    def params(self) -> List[ParamSpec]:
        return [p for b in self.blocks for p in b.params]

    @property
    def ndim_full(self) -> int:
        return len(self.params)

    @property
    def free_index(self) -> np.ndarray:
        """model.C_ (emp_model.py:273-279): full indices of the free parameters."""
        return np.array([i for i, p in enumerate(self.params) if p.fixed is None], dtype=np.int32)

    @property
    def ndim(self) -> int:
        return int(len(self.free_index))

    def free_params(self) -> List[ParamSpec]:
        return [p for p in self.params if p.fixed is None]

    def prior_widths(self) -> np.ndarray:
        """sampler.D_ (emp.py:595-602): hi - lo of every free parameter."""
        return np.array([p.limits[1] - p.limits[0] for p in self.free_params()], dtype=np.float64)

    def compile(self) -> "CompiledModel":
        return CompiledModel(self)


# ---- prior normalisers (emp_model.py:82-120 _check_prargs) -------------------
def _norm_cdf(x: float) -> float:
    try:  # the reference uses scipy.stats.norm.cdf == scipy.special.ndtr
        from scipy.special import ndtr
        return float(ndtr(x))
    except Exception:  # pragma: no cover - scipy is present in the image
        return 0.5 * math.erfc(-x / math.sqrt(2.0))


def finalize_prargs(prior: str, limits, prargs):
    """Apply `_check_prargs` (emp_model.py:82-99) to raw user prargs."""
    low, high = limits
    if prior == "Uniform":
        with np.errstate(divide="ignore"):
            return float(np.log(1 / (np.float64(high) - np.float64(low))))
    if prior == "Normal":
        mu, s = float(prargs[0]), float(prargs[1])
        a, b = (low - mu) / s, (high - mu) / s
        return [mu, s, float(np.log(_norm_cdf(b) - _norm_cdf(a)))]
    if prior == "Isotropic":
        return float(np.log(0.5 * (np.cos(low) - np.cos(high))))
    return prargs


def _prior_op(op, prior, i0, i1, limits, prargs):
    if prior in UNSUPPORTED_PRIORS or prior not in PRIOR_KINDS:
        raise UnsupportedModelError(
            f"prior '{prior}' (support/priors/{prior}.prior) is not implemented on the device path")
    kind = PRIOR_KINDS[prior]
    lo, hi = (float(limits[0]), float(limits[1])) if prior != "Fixed" else (float("nan"), float("nan"))
    a0 = a1 = a2 = a3 = 0.0
    if prior == "Uniform":
        a0 = float(prargs)
    elif prior == "Isotropic":
        a0 = float(prargs)
    elif prior == "Jeffreys":
        # support/priors/Jeffreys.prior:5 evaluates np.log(1/(high-low)) per call
        a0 = float(np.log(1 / (np.float64(hi) - np.float64(lo))))
    elif prior == "Normal":
        mu, s, logz = float(prargs[0]), float(prargs[1]), float(prargs[2])
        a0, a1, a2 = mu, s, logz
        a3 = float(np.log(s * np.sqrt(2 * np.pi)))  # support/priors/Normal.prior:8
    return (op, kind, int(i0), int(i1), lo, hi, a0, a1, a2, a3)


class _PriorOpC(ctypes.Structure):
    _fields_ = [("op", ctypes.c_int32), ("prior", ctypes.c_int32), ("i0", ctypes.c_int32),
                ("i1", ctypes.c_int32), ("lo", ctypes.c_double), ("hi", ctypes.c_double),
                ("a0", ctypes.c_double), ("a1", ctypes.c_double), ("a2", ctypes.c_double),
                ("a3", ctypes.c_double)]


class EmpModelDescC(ctypes.Structure):
    """ctypes mirror of `EmpModelDesc` (include/emperor_b200.h)."""
    _fields_ = [
        ("abi_version", ctypes.c_int32), ("ndim_free", ctypes.c_int32),
        ("ndim_full", ctypes.c_int32), ("n_kep", ctypes.c_int32),
        ("kep_model", ctypes.c_int32 * EMP_MAX_KEP), ("kep_off", ctypes.c_int32 * EMP_MAX_KEP),
        ("acc_order", ctypes.c_int32), ("acc_off", ctypes.c_int32),
        ("n_ins", ctypes.c_int32), ("offset_off", ctypes.c_int32),
        ("has_jitter", ctypes.c_int32), ("jitter_off", ctypes.c_int32),
        ("ma_mode", ctypes.c_int32), ("ma_order", ctypes.c_int32), ("ma_off", ctypes.c_int32),
        ("am_enabled", ctypes.c_int32), ("am_offset_off", ctypes.c_int32),
        ("am_jitter_off", ctypes.c_int32), ("n_prior_ops", ctypes.c_int32),
        ("n_periodic", ctypes.c_int32), ("periodic_kind", ctypes.c_int32 * EMP_MAX_PERIODIC),
        ("periodic_off", ctypes.c_int32 * EMP_MAX_PERIODIC),
        ("n_sai", ctypes.c_int32), ("sai_off", ctypes.c_int32), ("sai_count", ctypes.c_int32 * EMP_MAX_INS),
        ("free_to_full", ctypes.c_int32 * EMP_MAX_DIM),
        ("full_init", ctypes.c_double * EMP_MAX_DIM),
        ("prior_ops", _PriorOpC * EMP_MAX_PRIOR_OPS),
    ]


class CompiledModel:
    """Flat form of a ModelSpec: the layout facts `_write_model_RV`
    (emp_model.py:706-781) and `_write_prior_reddemcee` (emp.py:182-254) bake
    into the generated script."""

    def __init__(self, spec: ModelSpec):
        self.spec = spec
        self.ndim_full = spec.ndim_full
        self.free_to_full = spec.free_index
        self.ndim_free = len(self.free_to_full)
        if self.ndim_full > EMP_MAX_DIM:
            raise UnsupportedModelError(f"{self.ndim_full} parameters > EMP_MAX_DIM={EMP_MAX_DIM}")
        self.full_init = np.zeros(self.ndim_full, dtype=np.float64)
        for i, p in enumerate(spec.params):
            if p.fixed is not None:
                self.full_init[i] = float(p.fixed)

        self.kep_model: List[int] = []
        self.kep_off: List[int] = []
        self.acc_order = 0
        self.acc_off = 0
        self.n_ins = int(spec.nins)
        self.offset_off = -1
        self.has_jitter = 0
        self.jitter_off = 0
        self.ma_mode, self.ma_order, self.ma_off = MA_NONE, 0, 0
        self.am_enabled, self.am_offset_off, self.am_jitter_off = 0, 0, 0
        self.periodic: List[tuple] = []  # (kind, offset): 0 sinusoid00.model, 1 magneticcycle00.model
        self.sai_off, self.sai_count = 0, [0] * self.n_ins  # StellarActivityBlock: columns per instrument
        self.prior_ops: List[tuple] = []

        off = 0
        seen_tail = False  # blocks after the Keplerians, in evaluation order
        order_types = []
        for b in spec.blocks:
            n = len(b.params)
            order_types.append(b.type_)
            if b.type_ == "Keplerian":
                if seen_tail:
                    raise UnsupportedModelError("Keplerian blocks must precede the instrument blocks "
                                                "(emp.py:1235 inserts them first)")
                model = 5 if b.astrometry else int(b.parameterisation)
                if model not in KEP_NPAR or KEP_NPAR[model] != n:
                    raise UnsupportedModelError(f"Keplerian parameterisation {b.parameterisation} "
                                                f"with {n} parameters")
                self.kep_model.append(model)
                self.kep_off.append(off)
            elif b.type_ == "Acceleration":
                seen_tail = True
                self.acc_order, self.acc_off = n, off
                if n > EMP_MAX_ACC:
                    raise UnsupportedModelError(f"acceleration order {n} > {EMP_MAX_ACC}")
            elif b.type_ == "Offset":
                seen_tail = True
                if n != self.n_ins:
                    raise UnsupportedModelError("OffsetBlock length != nins")
                self.offset_off = off
            elif b.type_ == "Jitter":
                seen_tail = True
                if n != self.n_ins:
                    raise UnsupportedModelError("JitterBlock length != nins")
                self.has_jitter, self.jitter_off = 1, off
            elif b.type_ == "MOAV":
                seen_tail = True
                self.ma_order, self.ma_off = int(b.moav_order), off
                if self.ma_order > EMP_MAX_MA:
                    raise UnsupportedModelError(f"MA order {self.ma_order} > {EMP_MAX_MA}")
                if self.ma_order > 0:
                    self.ma_mode = MA_GLOBAL if b.moav_global else MA_REFERENCE_NOOP
            elif b.type_ in ("Sinusoid", "MagneticCycle"):
                seen_tail = True
                kind = 0 if b.type_ == "Sinusoid" else 1
                if n != (3, 5)[kind]:
                    raise UnsupportedModelError(f"{b.type_} block with {n} parameters")
                self.periodic.append((kind, off))
                if len(self.periodic) > EMP_MAX_PERIODIC:
                    raise UnsupportedModelError(f"more than {EMP_MAX_PERIODIC} Sinusoid/MagneticCycle blocks")
            elif b.type_ == "StellarActivity":
                seen_tail = True
                counts = [int(c) for c in b.sai_counts]
                if len(counts) != self.n_ins or sum(counts) != n or n == 0:
                    raise UnsupportedModelError("StellarActivityBlock: columns per instrument do not match")
                if max(counts) > EMP_MAX_SAI:
                    raise UnsupportedModelError(f"more than {EMP_MAX_SAI} activity columns for one instrument")
                self.sai_off, self.sai_count = off, counts
            elif b.type_ == "AstrometryOffset":
                seen_tail = True
                self.am_enabled, self.am_offset_off = 1, off
            elif b.type_ == "AstrometryJitter":
                seen_tail = True
                self.am_jitter_off = off
            else:
                # StellarActivityPRO, Celerite2: SURVEY.md §2 rows 19, 23
                raise UnsupportedModelError(f"block type '{b.type_}' is outside the device hot path")

            # --- prior program, in the order emp.py:200-246 writes it
            for j, p in enumerate(b.params):
                if p.fixed is not None:
                    continue
                self.prior_ops.append(_prior_op(POP_PARAM, p.prior, off + j, 0, p.limits, p.prargs))
            self.prior_ops.append((POP_CHECK, 0, 0, 0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0))
            for a in b.additional:
                if a.kind == "Ecc":
                    i0, i1 = off + 3, off + 4
                elif a.kind == "Amp":
                    i0, i1 = off + 1, off + 2
                else:
                    raise UnsupportedModelError(f"additional prior '{a.kind}'")
                self.prior_ops.append(_prior_op(POP_SUMSQ, a.prior, i0, i1, a.limits, a.prargs))
            off += n

        self._check_order(order_types)
        if len(self.kep_model) > EMP_MAX_KEP:
            raise UnsupportedModelError(f"{len(self.kep_model)} Keplerians > EMP_MAX_KEP={EMP_MAX_KEP}")
        if self.n_ins > EMP_MAX_INS:
            raise UnsupportedModelError(f"{self.n_ins} instruments > EMP_MAX_INS={EMP_MAX_INS}")
        if self.offset_off < 0:
            # the reference always adds an OffsetBlock (emp.py:2632)
            raise UnsupportedModelError("model has no OffsetBlock")
        if self.am_enabled and any(m != 5 for m in self.kep_model):
            raise UnsupportedModelError("astrometry needs AstrometryKeplerianBlock (akep00.model)")
        if self.am_enabled and "AstrometryJitter" not in order_types:
            raise UnsupportedModelError("astrometry needs an AstrometryJitterBlock (emp.py:1300-1311)")
        if self.am_enabled and self.ndim_free != self.ndim_full:
            # loglike_AM indexes the UN-expanded theta with full-theta slices (a00.like:7,
            # emp_model.py:1633-1637): with a fixed parameter the reference reads the wrong entries
            raise UnsupportedModelError("astrometric models with fixed parameters are undefined in the reference")

    def _check_order(self, types):
        """The device kernel evaluates acc -> offset -> jitter -> MA like the
        reference's `_autorun_add_RV_ins` (emp.py:2628-2651) orders them; the MA
        residual must see every mean-model term, so MOAV has to come last of the
        RV blocks."""
        rv = [t for t in types if t in ("Acceleration", "Offset", "Jitter", "MOAV")]
        seq = [t for t in types if t in ("MOAV", "StellarActivity", "Sinusoid", "MagneticCycle")]
        if "MOAV" in seq and seq.index("MOAV") != 0:
            # the reference computes the MA residuals before the activity / periodic terms are added
            raise UnsupportedModelError("StellarActivity/Sinusoid/MagneticCycle blocks before MOAV are not supported")
        if "MOAV" in rv and rv.index("MOAV") < max(
                (i for i, t in enumerate(rv) if t in ("Acceleration", "Offset")), default=-1):
            raise UnsupportedModelError("MOAV block before Offset/Acceleration is not supported")

    # -- C view ---------------------------------------------------------------
    def to_c(self) -> EmpModelDescC:
        d = EmpModelDescC()
        d.abi_version = EMP_ABI_VERSION
        d.ndim_free, d.ndim_full = self.ndim_free, self.ndim_full
        d.n_kep = len(self.kep_model)
        for k, (m, o) in enumerate(zip(self.kep_model, self.kep_off)):
            d.kep_model[k], d.kep_off[k] = m, o
        d.acc_order, d.acc_off = self.acc_order, self.acc_off
        d.n_ins, d.offset_off = self.n_ins, self.offset_off
        d.has_jitter, d.jitter_off = self.has_jitter, self.jitter_off
        d.ma_mode, d.ma_order, d.ma_off = self.ma_mode, self.ma_order, self.ma_off
        d.am_enabled, d.am_offset_off, d.am_jitter_off = (self.am_enabled, self.am_offset_off,
                                                          self.am_jitter_off)
        if len(self.prior_ops) > EMP_MAX_PRIOR_OPS:
            raise UnsupportedModelError("prior program too long")
        d.n_prior_ops = len(self.prior_ops)
        d.n_periodic = len(self.periodic)
        for j, (kind, o) in enumerate(self.periodic):
            d.periodic_kind[j], d.periodic_off[j] = kind, o
        d.n_sai, d.sai_off = int(sum(self.sai_count)), self.sai_off
        for i, c in enumerate(self.sai_count):
            d.sai_count[i] = c
        for j, f in enumerate(self.free_to_full):
            d.free_to_full[j] = int(f)
        for j, v in enumerate(self.full_init):
            d.full_init[j] = float(v)
        for j, t in enumerate(self.prior_ops):
            o = d.prior_ops[j]
            (o.op, o.prior, o.i0, o.i1, o.lo, o.hi, o.a0, o.a1, o.a2, o.a3) = t
        return d

    def expand(self, theta: np.ndarray) -> np.ndarray:
        """theta[..., ndim_free] -> full[..., ndim_full] (emp_model.py:709-711)."""
        theta = np.asarray(theta, dtype=np.float64)
        full = np.broadcast_to(self.full_init, theta.shape[:-1] + (self.ndim_full,)).copy()
        full[..., self.free_to_full] = theta
        return full


# ---- bridge from the reference's object graph -----------------------------------
def _plain(x):
    if x is None:
        return None
    if isinstance(x, (list, tuple, np.ndarray)):
        return [_plain(v) for v in x]
    if isinstance(x, (np.floating, np.integer)):
        return x.item()
    return x


def spec_from_reddmodel(model) -> ModelSpec:
    """Build a ModelSpec from a refreshed reference `ReddModel`
    (emp_model.py:223-330; blocks from block_repo.py).  Duck-typed: needs
    `for b in model`, `b.type_`, `b.parameterisation`, the parameters'
    `name/prior/limits/prargs/fixed/init_pos/is_hou`, and
    `b.additional_parameters` with `name/has_prior/prior/limits/prargs`."""
    blocks = []
    nins = int(getattr(model, "nins__", 0) or 0)
    for b in model:
        params = []
        for p in b:
            prior = p.prior
            fixed = getattr(p, "fixed", None)
            params.append(ParamSpec(
                name=str(p.name), prior=str(prior),
                limits=_plain(list(p.limits)) if fixed is None else [float("nan"), float("nan")],
                prargs=_plain(p.prargs), fixed=None if fixed is None else float(fixed),
                init_pos=_plain(list(getattr(p, "init_pos", [None, None]))),
                is_hou=bool(getattr(p, "is_hou", False))))
        addi = []
        for a in getattr(b, "additional_parameters", []):
            if isinstance(a, dict):  # 'Hill' is appended as a dict (block_repo.py:617-628)
                if a.get("has_prior"):
                    raise UnsupportedModelError("Hill prior (support/priors/Hill.prior)")
                continue
            if not getattr(a, "has_prior", False):
                continue
            if a.name[:3] in ("Amp", "Ecc"):
                addi.append(AdditionalPrior(kind=a.name[:3], prior=str(a.prior),
                                            limits=_plain(list(a.limits)), prargs=_plain(a.prargs)))
            else:
                raise UnsupportedModelError(f"additional prior on '{a.name}'")
        bs = BlockSpec(type_=str(b.type_), params=params, additional=addi,
                       parameterisation=getattr(b, "parameterisation", None),
                       astrometry=bool(getattr(b, "astrometry_bool", False)),
                       number=int(getattr(b, "number_", 0) or 0),
                       moav_order=int(getattr(b, "moav", 0) or 0),
                       moav_global=bool(getattr(b, "is_global", False)),
                       sai_counts=[int(c) for c in getattr(b, "cornums", [])] if str(b.type_) == "StellarActivity" else [])
        if bs.type_ == "Offset":
            nins = len(params)
        blocks.append(bs)
    return ModelSpec(blocks=blocks, nins=nins)

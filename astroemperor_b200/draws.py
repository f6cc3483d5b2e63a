"""Host-side random draws for the parallel-tempering step.

BASELINE.json north_star: "the affine-invariant stretch-move proposal, the
Metropolis accept and the adjacent-temperature swap are done on device from
host-supplied random draws".  This module is that host supply.  The draw ORDER
restates what the reference stack consumes from its `numpy.random.RandomState`
(emcee 3.1.6 `RedBlueMove.propose` + `StretchMove.get_proposal`, then the
ptemcee-lineage swap sweep of reddemcee; SURVEY.md §3.3, §8c row C2 — recalled,
the packages are not vendored):

  per temperature (its own RandomState, like reddemcee's one emcee sampler per temperature):
    for step in range(nsteps):
      inds = arange(W) % 2 ; shuffle(inds)
      for split in (0, 1):
        zz   = ((a-1)*rand(Ns) + 1)**2 / a
        rint = randint(Nc, size=Ns)
        u    = rand() for each walker of the split, in index order
  swap draws, one RandomState per adjacent pair (i, i-1), consumed hot -> cold:
    iperm = permutation(W) ; i1perm = permutation(W) ; raccept = log(uniform(size=W))
  (one stream per pair, not one for the sweep: a rank of a sharded ladder draws only the pairs of
  its own temperatures and the ranks all-gather them over NCCL — with one sequential stream every
  rank had to generate the 2(T-1) permutations of the WHOLE ladder, 10 ms per sweep at T = 256)

Logs (`(ndim-1)*log(zz)`, `log(u)`) are taken here with NumPy so the device and
the oracle compare against bit-identical thresholds.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class SweepDraws:
    half_idx: np.ndarray  # [nsteps, T_loc, 2, H] int32: walkers of split 0 / 1 (ascending)
    zz: np.ndarray        # [nsteps, T_loc, 2, H]
    rint: np.ndarray      # [nsteps, T_loc, 2, H] int32
    factors: np.ndarray   # [nsteps, T_loc, 2, H]  (ndim-1)*log(zz)
    lnu: np.ndarray       # [nsteps, T_loc, 2, H]  log(u)
    perm: np.ndarray      # [R, 2, W] int32: the row of pair j couples temperature j+1 (row 0) with j (row 1);
                          # relabelled so that row 0 is the identity (relabel_swap_draws)
    lnu_swap: np.ndarray  # [R, W]   (R = T-1 pairs, or this rank's rows of a sharded ladder)
    sharded_swap: bool = False  # perm / lnu_swap hold only this rank's pair rows (to be all-gathered)

    FIELDS = ("half_idx", "zz", "rint", "factors", "lnu", "perm", "lnu_swap")

    def nbytes(self) -> int:
        return sum(getattr(self, f).nbytes for f in self.FIELDS)


class DrawStreams:
    """One `RandomState` per temperature (reddemcee keeps one emcee sampler, hence one random
    state, per temperature) plus one per adjacent swap pair and one for the initial ensemble, all
    derived from a single seed.  A rank of a sharded ladder only advances the streams of its
    own temperatures and of the swap pairs it was assigned.

    native=True: the temperature and swap-pair streams are handed to the C generator of the C-ABI library
    (csrc/emp_draws.cpp: the same MT19937 states, the same legacy RandomState algorithms, bit-identical
    draws — tests/test_host_logic.py compares the two), which fills a whole sweep in one call, in threads."""

    def __init__(self, seed, ntemps: int, native: bool = False, n_threads: int = 0):
        ss = np.random.SeedSequence(seed)
        kids = ss.spawn(2 * ntemps + 1)
        mk = lambda k: np.random.RandomState(np.random.MT19937(k))
        self.temp = [mk(k) for k in kids[:ntemps]]
        self.init = mk(kids[ntemps + 1])
        self.swap_pair = [mk(k) for k in kids[ntemps + 2:]]  # pair j: temperatures (j+1, j), j < ntemps-1
        self.ntemps = ntemps
        self.native = bool(native)
        import os
        self.n_threads = int(n_threads) if n_threads else max(1, min(8, (os.cpu_count() or 2) // 2))
        self._h = None

    def handle(self):
        """The C generator, created on first use from the RandomStates' current MT19937 states
        (stream t = temperature t, stream ntemps + j = swap pair j)."""
        if self._h is None:
            import ctypes
            from . import _lib
            states = [r.get_state(legacy=True) for r in self.temp + self.swap_pair]
            keys = np.ascontiguousarray(np.stack([st[1] for st in states]), dtype=np.uint32)
            pos = np.array([st[2] for st in states], dtype=np.int32)
            h = ctypes.c_void_p()
            _lib.check(_lib.lib().emp_draws_create(len(states), keys.ctypes.data, pos.ctypes.data, self.n_threads,
                                                   ctypes.byref(h)))
            self._h = h
        return self._h

    def __del__(self):
        try:
            if self._h is not None:
                from . import _lib
                _lib.lib().emp_draws_destroy(self._h)
                self._h = None
        except Exception:
            pass


def _zz_from_u(u, a):
    """emcee StretchMove: zz = ((a - 1) u + 1)^2 / a (the same NumPy expression on both draw paths)."""
    return ((a - 1.0) * u + 1) ** 2.0 / a


def draw_stretch(rng: np.random.RandomState, W: int, nsteps: int, a: float = 2.0):
    """Draws of ONE temperature for `nsteps` RedBlue stretch steps, in emcee's order."""
    H = W // 2
    half_idx = np.empty((nsteps, 2, H), dtype=np.int32)
    zz = np.empty((nsteps, 2, H))
    rint = np.empty((nsteps, 2, H), dtype=np.int32)
    u = np.empty((nsteps, 2, H))
    base = np.arange(W) % 2
    for s in range(nsteps):
        inds = base.copy()
        rng.shuffle(inds)
        half_idx[s, 0] = np.flatnonzero(inds == 0)
        half_idx[s, 1] = np.flatnonzero(inds == 1)
        for split in (0, 1):
            zz[s, split] = _zz_from_u(rng.rand(H), a)
            rint[s, split] = rng.randint(H, size=(H,))
            u[s, split] = rng.rand(H)
    return half_idx, zz, rint, u


def relabel_swap_draws(iperm, i1perm, u):
    """The reference's swap sweep pairs slot iperm[k] of the warmer row with slot i1perm[k] of the colder one
    under the uniform u[k].  The same SET of (a, b, u) triples listed by a: row 0 becomes the identity, row 1 the
    partner b(a), u the uniform of that pair.  Nothing about the sweep changes (the pairs of one sweep are
    disjoint), but thread a of the plan kernel then owns slot a of the warmer row (emp_pt.cuh)."""
    order = np.argsort(iperm, kind="stable")  # inverse permutation
    return iperm[order], i1perm[order], u[order]


def sweep_shapes(T_loc: int, W: int, nsteps: int, n_rows: int):
    """(field, shape, dtype) of the seven arrays of a sweep's draws, in SweepDraws.FIELDS order."""
    H = W // 2
    s4 = (nsteps, T_loc, 2, H)
    return [("half_idx", s4, np.int32), ("zz", s4, np.float64), ("rint", s4, np.int32), ("factors", s4, np.float64),
            ("lnu", s4, np.float64), ("perm", (n_rows, 2, W), np.int32), ("lnu_swap", (n_rows, W), np.float64)]


def draw_sweep(streams: DrawStreams, W: int, ndim: int, nsteps: int, a: float = 2.0,
               temps: slice = None, swap: bool = True, swap_rows=None, out=None) -> SweepDraws:
    """Draws of one sweep: stretch draws for the temperatures in `temps` (default: all), swap
    draws for the pairs in `swap_rows` (default: all T-1; a row index >= T-1 yields a zero row, the
    padding of a sharded ladder whose ranks hold T/G rows each).  `out`: dict of preallocated arrays
    (e.g. views of a pinned staging buffer) to fill instead of allocating."""
    if W % 2:
        raise ValueError("nwalkers must be even (two equal halves, emcee RedBlueMove nsplits=2)")
    T = streams.ntemps
    tl = range(T)[temps] if temps is not None else range(T)
    rows = list(range(max(T - 1, 0))) if swap_rows is None else list(swap_rows)
    if out is None:
        out = {f: (np.zeros if f in ("perm", "lnu_swap") else np.empty)(shp, dtype=dt)
               for f, shp, dt in sweep_shapes(len(tl), W, nsteps, len(rows))}
    half_idx, zz, rint, factors, lnu = (out[f] for f in ("half_idx", "zz", "rint", "factors", "lnu"))
    perm, lnu_swap = out["perm"], out["lnu_swap"]
    if streams.native:
        from . import _lib
        L, h = _lib.lib(), streams.handle()
        ts = np.array(list(tl), dtype=np.int32)
        rs = np.array([T + j if j < T - 1 else -1 for j in rows] if swap else [], dtype=np.int32)
        _lib.check(L.emp_draws_sweep(h, ts.ctypes.data, len(ts), W, nsteps, half_idx.ctypes.data, zz.ctypes.data,
                                     rint.ctypes.data, lnu.ctypes.data, rs.ctypes.data, len(rs), perm.ctypes.data,
                                     lnu_swap.ctypes.data))
        zz[...] = _zz_from_u(zz, a)
        if not len(rs):
            perm[...] = 0
            lnu_swap[...] = 1.0
    else:
        H = W // 2
        for j, t in enumerate(tl):
            half_idx[:, j], zz[:, j], rint[:, j], lnu[:, j] = draw_stretch(streams.temp[t], W, nsteps, a)
        perm[...] = 0
        lnu_swap[...] = 1.0
        if swap:
            for k, j in enumerate(rows):
                if j >= T - 1:
                    continue
                rng = streams.swap_pair[j]
                iperm, i1perm, u = rng.permutation(W), rng.permutation(W), rng.uniform(size=W)
                perm[k, 0], perm[k, 1], lnu_swap[k] = relabel_swap_draws(iperm, i1perm, u)
    # the logs are NumPy's on both paths: device and oracle compare against the same bits
    np.multiply(np.log(zz), ndim - 1.0, out=factors)
    with np.errstate(divide="ignore"):
        np.log(lnu, out=lnu)
        np.log(lnu_swap, out=lnu_swap)
    return SweepDraws(half_idx, zz, rint, factors, lnu, perm, lnu_swap, sharded_swap=swap_rows is not None)


def initial_positions(rng: np.random.RandomState, spec, ntemps: int, nwalkers: int) -> np.ndarray:
    """set_init() of the generated script (emp.py:617-654): per temperature and free parameter
    `pos = r*(2*sort(U(0,1,W)) - 1) + m`, shuffled; m, r from init_pos or the limits, r*0.707
    for `is_hou` parameters."""
    fp = spec.free_params()
    pos = np.zeros((ntemps, nwalkers, len(fp)))
    for t in range(ntemps):
        for j, p in enumerate(fp):
            r_f = 0.707 if p.is_hou else 1
            b = p.limits[0] if p.init_pos[0] is None else np.round(p.init_pos[0], 8)
            a = p.limits[1] if p.init_pos[1] is None else np.round(p.init_pos[1], 8)
            m = (a + b) / 2
            r = (a - b) / 2 * r_f
            dist = np.sort(rng.uniform(0, 1, nwalkers))
            pos[t][:, j] = r * (2 * dist - 1) + m
            rng.shuffle(pos[t, :, j])
    return pos


# Geometric spacing of the default ladder per dimension (index ndim - 1): the table of emcee v2's PTSampler /
# ptemcee `default_beta_ladder` (Vousden, Farr & Mandel 2016: the spacing that gives ~25 % swap acceptance for a
# unimodal Gaussian of that dimension).  reddemcee 0.9 uses it unchanged — the reference's own notebooks show it:
# tests/00_mini_test.ipynb / 01_51peg_basic.ipynb print hottest rungs 5.057e-10 = 7**-11 and 2.478e-08 = 7**-9 for the
# 2-parameter model with 12 and 10 temperatures (table[1] = 7), tests/quickstart.ipynb prints [1.0, 0.4002] for the
# 7-parameter model with 2 temperatures (1 / table[6] = 0.40019): tests/test_host_logic.py pins both.
_TSTEP = np.array([
    25.2741, 7., 4.47502, 3.5236, 3.0232, 2.71225, 2.49879, 2.34226, 2.22198, 2.12628,
    2.04807, 1.98276, 1.92728, 1.87946, 1.83774, 1.80096, 1.76826, 1.73895, 1.7125, 1.68849,
    1.66657, 1.64647, 1.62795, 1.61083, 1.59494, 1.58014, 1.56632, 1.55338, 1.54123, 1.5298,
    1.51901, 1.50881, 1.49916, 1.49, 1.4813, 1.47302, 1.46512, 1.45759, 1.45039, 1.4435,
    1.4369, 1.43056, 1.42448, 1.41864, 1.41302, 1.40761, 1.40239, 1.39736, 1.3925, 1.38781,
    1.38327, 1.37888, 1.37463, 1.37051, 1.36652, 1.36265, 1.35889, 1.35524, 1.3517, 1.34825,
    1.3449, 1.34164, 1.33847, 1.33538, 1.33236, 1.32943, 1.32656, 1.32377, 1.32104, 1.31838,
    1.31578, 1.31325, 1.31076, 1.30834, 1.30596, 1.30364, 1.30137, 1.29915, 1.29697, 1.29484,
    1.29275, 1.29071, 1.2887, 1.28673, 1.2848, 1.28291, 1.28106, 1.27923, 1.27745, 1.27569,
    1.27397, 1.27227, 1.27061, 1.26898, 1.26737, 1.26579, 1.26424, 1.26271, 1.26121, 1.25973])


def default_betas(ndim: int, ntemps: int) -> np.ndarray:
    """The default ladder `betas = tstep**-arange(ntemps)` of ptemcee's `default_beta_ladder(ndim, ntemps)` (Tmax not
    given: the hottest chain keeps a finite temperature), which reddemcee 0.9 uses when EMPEROR passes `betas=None`
    (`emp.py:2372`): tstep from the table above for ndim <= 100, `1 + 2 sqrt(ln 4 / ndim)` beyond it."""
    if ndim < 1:
        raise ValueError("ndim must be >= 1")
    tstep = _TSTEP[ndim - 1] if ndim <= len(_TSTEP) else 1.0 + 2.0 * np.sqrt(np.log(4.0)) / np.sqrt(ndim)
    return tstep ** (-np.arange(ntemps, dtype=np.float64))

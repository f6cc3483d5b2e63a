"""Host-side reductions over the stored chains.

These are the read-back calls the reference's parent process makes on the sampler after a
run (`emp.py:1375-1385` autocorrelation time, `emp.py:1432-1447` evidence, `emp.py:722-762`
HDF5 dump).  They run once per run on [T, n, W] arrays that were copied back from the GPU;
plain NumPy, not part of the hot path.
"""
from __future__ import annotations

import numpy as np


# ---- integrated autocorrelation time (emcee 3 `autocorr.integrated_time`, Sokal window) -------
def _next_pow_two(n):
    i = 1
    while i < n:
        i = i << 1
    return i


def function_1d(x):
    """Normalised autocorrelation function of a 1-D series (FFT)."""
    x = np.atleast_1d(x)
    n = _next_pow_two(len(x))
    f = np.fft.fft(x - np.mean(x), n=2 * n)
    acf = np.fft.ifft(f * np.conjugate(f))[: len(x)].real
    acf /= acf[0]
    return acf


def _auto_window(taus, c):
    m = np.arange(len(taus)) < c * taus
    if np.any(m):
        return int(np.argmin(m))
    return len(taus) - 1


def integrated_time(x, c=5, tol=50, quiet=False):
    """x [n_steps, n_walkers, ndim] -> tau [ndim] (emcee's estimator: walker-averaged ACF,
    automated windowing M >= c*tau).  Raises unless the chain is longer than tol*tau, or warns
    when `quiet`."""
    x = np.atleast_1d(x)
    if x.ndim == 1:
        x = x[:, None, None]
    if x.ndim == 2:
        x = x[:, :, None]
    n_t, n_w, n_d = x.shape
    tau_est = np.empty(n_d)
    windows = np.empty(n_d, dtype=int)
    for d in range(n_d):
        f = np.zeros(n_t)
        for k in range(n_w):
            f += function_1d(x[:, k, d])
        f /= n_w
        taus = 2.0 * np.cumsum(f) - 1.0
        windows[d] = _auto_window(taus, c)
        tau_est[d] = taus[windows[d]]
    flag = tol * tau_est > n_t
    if np.any(flag) and tol > 0:
        msg = (f"The chain is shorter than {tol} times the integrated autocorrelation time for "
               f"{int(np.sum(flag))} parameter(s). Use this estimate with caution; N/{tol} = {n_t / tol:.0f}; "
               f"tau: {tau_est}")
        if not quiet:
            raise RuntimeError(msg)
        import warnings
        warnings.warn(msg)
    return tau_est


# ---- evidence ---------------------------------------------------------------------------------
def _sorted_by_beta(betas, *arrs):
    order = np.argsort(betas)
    return (np.asarray(betas)[order],) + tuple(np.asarray(a)[order] for a in arrs)


def evidence_ti(logl, betas, pchip=False):
    """Thermodynamic integration: logZ = int_0^1 <logL>_beta dbeta over the ladder (logl [T, n_samples]):
    trapezoid rule, or with `pchip=True` (the option emp.py:1434-1446 passes on) the integral of the
    monotone piecewise-cubic Hermite interpolant of <logL>(beta).  Error = |full ladder - every other
    rung| (ptemcee convention)."""
    mean_ll = np.mean(np.reshape(logl, (len(betas), -1)), axis=1)
    b, m = _sorted_by_beta(betas, mean_ll)
    if b[0] > 0:
        b, m = np.concatenate([[0.0], b]), np.concatenate([[m[0]], m])
    trap = np.trapezoid if hasattr(np, "trapezoid") else np.trapz

    def integral(bb, mm):
        if pchip and len(bb) >= 3:
            from scipy.interpolate import PchipInterpolator
            return float(PchipInterpolator(bb, mm).integrate(bb[0], bb[-1]))
        return float(trap(mm, bb))

    logz = integral(b, m)
    b2, m2 = b[::2], m[::2]
    if b2[-1] != b[-1]:
        b2, m2 = np.append(b2, b[-1]), np.append(m2, m[-1])
    return logz, abs(logz - integral(b2, m2))


def evidence_ss(logl, betas, n_batches=8):
    """Stepping-stone estimator (Xie et al. 2011): logZ = sum_i log < exp((b_{i+1}-b_i) logL) >_{b_i}
    over the ladder from the hottest to the coldest rung; error from `n_batches` batch means."""
    ll = np.reshape(logl, (len(betas), -1))
    b, ll = _sorted_by_beta(betas, ll)

    def ss(sub):
        tot = 0.0
        lo_b, lo_l = b, sub
        if b[0] > 0:  # bridge from beta = 0 with the hottest samples
            lo_b = np.concatenate([[0.0], b])
            lo_l = np.concatenate([sub[:1], sub])
        for i in range(len(lo_b) - 1):
            x = (lo_b[i + 1] - lo_b[i]) * lo_l[i]
            mx = np.max(x)
            tot += mx + np.log(np.mean(np.exp(x - mx)))
        return tot

    logz = ss(ll)
    n = ll.shape[1]
    if n >= 2 * n_batches:
        parts = [ss(ll[:, k * (n // n_batches):(k + 1) * (n // n_batches)]) for k in range(n_batches)]
        err = float(np.std(parts, ddof=1) / np.sqrt(n_batches))
    else:
        err = float("nan")
    return float(logz), err


# ---- chain sink in the reference's backend layout (emp.py:722-762) ------------------------------
def _h5_write(path, group, attrs, datasets):
    """One HDF5 file with one group: h5py where it is installed, else the minimal writer of h5min.py."""
    try:
        import h5py
    except ImportError:
        from .h5min import write_h5
        return write_h5(path, {group: {"attrs": attrs, "datasets": datasets}})
    with h5py.File(path, "w") as f:
        g = f.create_group(group)
        for k, v in attrs.items():
            g.attrs[k] = v
        for k, v in datasets.items():
            g.create_dataset(k, data=v)
    return path


def _h5_read(path, group):
    try:
        import h5py
    except ImportError:
        from .h5min import read_h5
        g = read_h5(path)[group]
        return g["attrs"], g["datasets"]
    with h5py.File(path, "r") as f:
        g = f[group]
        return dict(g.attrs), {k: g[k][...] for k in g}


def save_backend(sampler, name, discard=0, group="mcmc"):
    """Dump the run in the layout EMPEROR's generated script writes after a reddemcee run (emp.py:722-762) and its
    parent reads back (emp.py:781-789, `reddemcee.hdf.PTHDFBackend` + one emcee-style `HDFBackend_plus` per
    temperature): `<name>.h5` with group 'mcmc' {attrs iteration (sweeps), ntemps, nwalkers, ndim; datasets
    tsw_history, smd_history} and `<name>_<t>.h5` with group 'mcmc' {attrs iteration (stored steps), nwalkers,
    ndim, has_blobs, version; datasets chain[iter, W, ndim], log_like, log_prob, beta_history[iter], accepted[W]}.
    `discard` counts stored steps (sweep histories are cut at the sweep of the first kept step).
    A sharded ladder gathers first (every rank returns the same file name; rank 0 writes)."""
    chain = sampler.get_chain(discard=discard)           # [T, n, W, ndim]
    ll = sampler.get_log_like(discard=discard)
    lpost = sampler.get_log_prob(discard=discard)
    betas = sampler.get_betas(discard=discard)           # [n, T] per stored step
    n_steps = int(getattr(sampler, "_n_steps", chain.shape[1]))
    accepted = np.rint(np.asarray(sampler.acceptance_fraction) * max(n_steps, 1)).astype(np.int64)
    ss = np.asarray(getattr(sampler, "_sample_sweep", []), dtype=np.int64)
    d_sweeps = int(ss[discard]) if (discard and len(ss) > discard) else 0
    tsw, smd = sampler.get_tsw(discard=d_sweeps), sampler.get_smd(discard=d_sweeps)
    T, n, W, nd = chain.shape
    shard = getattr(sampler, "shard", None)
    if shard is not None and shard.world > 1 and shard.rank != 0:
        return name + ".h5"
    from . import __version__ as ver
    _h5_write(name + ".h5", group,
              dict(iteration=int(len(tsw)), ntemps=T, nwalkers=W, ndim=nd, version=f"b200-{ver}", n_steps=n_steps,
                   thin_by=int(getattr(sampler, "thin_by", 1)), tsw_history_bool=bool(len(tsw)),
                   smd_history_bool=bool(np.size(smd))),
              dict(tsw_history=np.asarray(tsw, dtype=np.float64), smd_history=np.asarray(smd, dtype=np.float64),
                   betas=np.asarray(sampler.betas, dtype=np.float64)))
    for t in range(T):
        _h5_write(f"{name}_{t}.h5", group,
                  dict(iteration=n, nwalkers=W, ndim=nd, has_blobs=False, version=f"b200-{ver}"),
                  dict(chain=chain[t], log_like=ll[t], log_prob=lpost[t], beta_history=betas[:, t],
                       accepted=accepted[t]))
    return name + ".h5"


# ---- reading a stored run back (the parent's `_load_sampler_reddemcee`, emp.py:777-807) ----------
class StoredRun:
    """A run written by `save_backend`, read back with the getters the reference's parent process calls on
    the re-created sampler (emp.py:961-965, 1375-1447, 1977-1989): `get_chain / get_log_like / get_log_prob
    (discard=, thin=, flat=)`, `betas`, `get_betas`, `get_tsw`, `get_smd`, `acceptance_fraction`, the evidence
    and autocorrelation reductions and the `backend[t]` view.  No device is needed: post-processing of a GPU run
    can happen anywhere."""

    def __init__(self, chain, log_like, log_prob, beta_history, accepted, tsw_history, smd_history, n_steps=None):
        self._chain = np.asarray(chain)               # [T, n, W, ndim]
        self._ll = np.asarray(log_like)               # [T, n, W]
        self._lpost = np.asarray(log_prob)            # [T, n, W]
        self._betas = np.asarray(beta_history)        # [n, T]
        self._accepted = np.asarray(accepted)         # [T, W]
        self._tsw = np.asarray(tsw_history)
        self._smd = np.asarray(smd_history)
        self.ntemps, self.iteration, self.nwalkers, self.ndim = self._chain.shape
        self._n_steps = n_steps if n_steps is not None else self.iteration
        self.betas = self._betas[-1].copy() if len(self._betas) else None
        self.tsw_history_bool, self.smd_history_bool = self._tsw.size > 0, self._smd.size > 0

    @staticmethod
    def _view(x, discard, thin, flat):
        x = x[:, discard::thin]
        return x.reshape((x.shape[0], x.shape[1] * x.shape[2]) + x.shape[3:]) if flat else x

    def get_chain(self, discard=0, thin=1, flat=False):
        return self._view(self._chain, discard, thin, flat)

    def get_log_like(self, discard=0, thin=1, flat=False):
        return self._view(self._ll, discard, thin, flat)

    def get_log_prob(self, discard=0, thin=1, flat=False):
        return self._view(self._lpost, discard, thin, flat)

    def get_betas(self, discard=0):
        return self._betas[discard:]

    def get_tsw(self, discard=0):
        return self._tsw[discard:]

    def get_smd(self, discard=0):
        return self._smd[discard:]

    @property
    def acceptance_fraction(self):
        return self._accepted / max(self._n_steps, 1)

    def get_evidence_ti(self, discard=0, pchip=False):
        return evidence_ti(self.get_log_like(discard=discard), self.betas, pchip=pchip)

    def get_evidence_ss(self, discard=0, pchip=False):
        return evidence_ss(self.get_log_like(discard=discard), self.betas)

    def get_autocorr_time(self, discard=0, thin=1, quiet=False, tol=50, c=5):
        ch = self.get_chain(discard=discard, thin=thin)
        return np.array([thin * integrated_time(ch[t], c=c, tol=tol, quiet=quiet) for t in range(ch.shape[0])])

    @property
    def backend(self):
        return _StoredBackend(self)


class _StoredBackend:
    def __init__(self, run):
        self._r = run
        self.iteration = run.iteration
        self.tsw_history_bool, self.smd_history_bool = run.tsw_history_bool, run.smd_history_bool
        self.tsw_history, self.smd_history = run._tsw, run._smd

    def __len__(self):
        return self._r.ntemps

    def __getitem__(self, t):
        r = self._r
        if not -r.ntemps <= t < r.ntemps:
            raise IndexError(t)
        t %= r.ntemps

        class _T:
            iteration = r.iteration
            accepted = r._accepted[t]
            get_chain = staticmethod(lambda: r._chain[t])
            get_log_like = staticmethod(lambda: r._ll[t])
            get_log_prob = staticmethod(lambda: r._lpost[t])
            get_betas = staticmethod(lambda: r._betas[:, t])
        return _T()


def load_backend(name, group="mcmc") -> StoredRun:
    """Read what `save_backend(sampler, name)` wrote: `<name>.h5` + `<name>_<t>.h5` (the reference's PTHDFBackend /
    HDFBackend_plus file layout, emp.py:781-785; h5py where installed, h5min otherwise), or the `.npz` of round 1."""
    import os
    if not os.path.exists(name + ".h5") and os.path.exists(name + ".npz"):
        z = np.load(name + ".npz")
        return StoredRun(z["chain"], z["log_like"], z["log_prob"], z["beta_history"], z["accepted"],
                         z["tsw_history"], z["smd_history"])
    a, d = _h5_read(name + ".h5", group)
    T = int(a["ntemps"])
    parts = []
    for t in range(T):
        _, dt = _h5_read(f"{name}_{t}.h5", group)
        parts.append([dt[k] for k in ("chain", "log_like", "log_prob", "beta_history", "accepted")])
    ch, ll, lpost, bh, acc = (np.stack([p[i] for p in parts]) for i in range(5))
    run = StoredRun(ch, ll, lpost, bh.T, acc, d["tsw_history"], d["smd_history"], n_steps=int(a.get("n_steps", 0)) or None)
    if "betas" in d and len(d["betas"]) == T:
        run.betas = np.asarray(d["betas"], dtype=np.float64)
    return run

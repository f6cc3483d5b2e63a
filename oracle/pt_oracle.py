"""ORACLE — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

NumPy restatement of one parallel-tempering sweep of the reference's sampler
stack, with INJECTED random draws:

  * emcee 3.1.6 `RedBlueMove.propose` with `nsplits=2` + `StretchMove(a=2)`:
      q = c[rint] - (c[rint] - s) * zz[:, None]; factors = (ndim - 1) * log(zz)
      accept iff factors + logpost(q) - logpost(s) > log(u)
    with the tempered posterior of ptemcee / reddemcee: logpost = beta*logL + logP;
  * the hot -> cold adjacent-temperature swap sweep (ptemcee `_temperature_swaps`
    lineage): for i = T-1 .. 1, two independent permutations, accept iff
      (beta_{i-1} - beta_i) * (logL_i[iperm] - logL_{i-1}[i1perm]) > log(u);
  * the Vousden et al. (2016) ladder adaptation in reddemcee's `adapt_tau`,
    `adapt_nu` parameterisation (adapt_mode 0: equalise neighbouring swap rates).
    Its one transcendental, `np.exp(dSs)`, is evaluated by `exp_det`: a FIXED sequence of
    individually rounded IEEE operations (<= 1 ulp from exp) that the device replays
    operation by operation (emp_pt.cuh::exp_det), so the device ladder is bit-identical
    to this one.  `np.exp` itself is neither correctly rounded nor the same on every host
    (NumPy dispatches to AVX-512/SVML or libm builds), i.e. the reference's own ladder is
    only defined to that level; tests/test_host_logic.py bounds the difference
    (exp_det vs np.exp vs a 40-digit exp; ladders after 2000 adaptations agree to 1e-13);
  * swap mean distance per sweep (`smd_history`): see `swap_mean_distance`.

Neither package is vendored in /root/reference nor installable here, so this is
a restatement of their published algorithms: "parity unpinned" at the sampler
boundary.  What the GPU path is held to is THIS step function: identical
accept / swap masks and bit-identical chains given identical draws.

The draws come from `astroemperor_b200.draws.draw_sweep` (read as plain arrays).
Only tests/, __graft_entry__.smoke() and bench.py may import this module.
"""
import numpy as np


def stretch_step(p, logl, logp, betas, half_idx, zz, rint, factors, lnu, loglike_fn):
    """One RedBlue stretch step of every temperature, in place.
    p [T,W,nd], logl/logp [T,W]; draws [T,2,H]; loglike_fn(q[n,nd]) -> (ll[n], lp[n]).
    Returns accepted [T,W] bool and the minimum decision margin |lnpdiff - ln u|."""
    T, W, nd = p.shape
    accepted = np.zeros((T, W), dtype=bool)
    margin = np.inf
    for split in (0, 1):
        for t in range(T):
            S1 = half_idx[t, split]
            C = half_idx[t, 1 - split]
            s = p[t, S1]
            c = p[t, C]
            cr = c[rint[t, split]]
            q = cr - (cr - s) * zz[t, split][:, None]
            ll_new, lp_new = loglike_fn(q)
            with np.errstate(invalid="ignore"):
                post_new = betas[t] * ll_new + lp_new
                post_old = betas[t] * logl[t, S1] + logp[t, S1]
                lnpdiff = factors[t, split] + post_new - post_old
                acc = lnpdiff > lnu[t, split]
                m = np.abs(lnpdiff - lnu[t, split])
            m = m[np.isfinite(m)]
            if len(m):
                margin = min(margin, float(m.min()))
            idx = S1[acc]
            p[t, idx] = q[acc]
            logl[t, idx] = ll_new[acc]
            logp[t, idx] = lp_new[acc]
            accepted[t, idx] = True
    return accepted, margin


def swap_sweep(p, logl, logp, betas, perm, lnu_swap):
    """Hot -> cold swap sweep, in place.  Returns (n_acc[T-1], src[T,W] plan, min margin)."""
    T, W, _ = p.shape
    n_acc = np.zeros(max(T - 1, 0), dtype=np.int32)
    src = np.arange(T * W, dtype=np.int32).reshape(T, W)
    margin = np.inf
    for i in range(T - 1, 0, -1):
        dbeta = betas[i - 1] - betas[i]
        iperm = perm[i - 1, 0]
        i1perm = perm[i - 1, 1]
        with np.errstate(invalid="ignore"):
            paccept = dbeta * (logl[i, iperm] - logl[i - 1, i1perm])
            asel = paccept > lnu_swap[i - 1]
            m = np.abs(paccept - lnu_swap[i - 1])
        m = m[np.isfinite(m)]
        if len(m):
            margin = min(margin, float(m.min()))
        n_acc[i - 1] = int(np.sum(asel))
        a, b = iperm[asel], i1perm[asel]
        for arr in (p, logl, logp, src):
            tmp = np.copy(arr[i, a])
            arr[i, a] = arr[i - 1, b]
            arr[i - 1, b] = tmp
    return n_acc, src, margin


_EXP_C = [1.0 / 6227020800.0, 1.0 / 479001600.0, 1.0 / 39916800.0, 1.0 / 3628800.0, 1.0 / 362880.0,
          1.0 / 40320.0, 1.0 / 5040.0, 1.0 / 720.0, 1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0, 0.5]


def exp_det(x):
    """exp(x) as a fixed sequence of IEEE double operations, each rounded once (NumPy never fuses
    `a*b + c`): n = rint(x log2 e), Cody-Waite reduction r = (x - n L1) - n L2, degree-13 Taylor
    polynomial by Horner, exp = (1 + (r + r^2 P(r))) 2^n.  |error| <= 1 ulp.  The device runs the
    same operations with single-rounding intrinsics (emp_pt.cuh::exp_det)."""
    x = np.asarray(x, dtype=np.float64)
    n = np.rint(x * 1.4426950408889634074)
    r = (x - n * 6.93147180369123816490e-01) - n * 1.90821492927058770002e-10
    pz = np.full_like(r, _EXP_C[0])
    for c in _EXP_C[1:]:
        pz = pz * r + c
    y = 1.0 + (r + (r * r) * pz)
    return np.ldexp(y, n.astype(np.int64))


def swap_mean_distance(p_before, p_after, src, D):
    """Swap mean distance of one sweep (consumers emp.py:961-965, 1985-1990): for every temperature the
    mean, over the slots that received a walker from a HOTTER rung, of the distance between the walker
    that arrived and the one that left, in units of the prior widths `sampler.D_` (emp.py:595-602).
    reddemcee's own definition is not recoverable offline; this is the definition the device implements
    (pt_apply_plan_kernel).  src [T, W]: flat index of the pre-sweep slot whose walker ends in (t, w)."""
    T, W, _ = p_before.shape
    down = (src // W) > np.arange(T)[:, None]
    dist = np.sqrt((((p_after - p_before) / D) ** 2).sum(-1))
    return (dist * down).sum(1) / np.maximum(down.sum(1), 1)


def adapt_ladder(betas, ratios, time, adapt_tau, adapt_nu, exp=exp_det):
    """Vousden, Farr & Mandel (2016) eq. 11-13 as in ptemcee `_get_ladder_adjustment`,
    with reddemcee's names: lag = adapt_tau, time-scale = adapt_nu.
    ratios [T-1]: swap acceptance between temperature i and i+1."""
    betas = betas.copy()
    T = len(betas)
    if T < 3:
        return betas
    decay = adapt_tau / (time + adapt_tau)
    kappa = decay / adapt_nu
    dSs = kappa * (ratios[:-1] - ratios[1:])
    deltaTs = np.diff(1 / betas[:-1])
    deltaTs = deltaTs * exp(dSs)
    betas[1:-1] = 1 / (np.cumsum(deltaTs) + 1 / betas[0])
    return betas


class PTOracle:
    """Whole-ladder sweep driver mirroring astroemperor_b200.sampler.PTSampler."""

    def __init__(self, rv_oracle, betas, adapt_tau=1000, adapt_nu=1, adapt=True, D=None):
        self.orc = rv_oracle
        self.D = None if D is None else np.asarray(D, dtype=np.float64)
        self.smd = None
        self.betas = np.array(betas, dtype=np.float64)
        self.adapt_tau, self.adapt_nu, self.adapt = adapt_tau, adapt_nu, adapt
        self.time = 0
        self.min_margin = np.inf

    def init_state(self, p0):
        self.p = np.array(p0, dtype=np.float64)
        T, W, nd = self.p.shape
        ll, lp = self.orc.logl_logp_batch(self.p.reshape(-1, nd))
        self.logl, self.logp = ll.reshape(T, W), lp.reshape(T, W)

    def sweep(self, draws):
        nsteps = draws.zz.shape[0]
        acc_all = []
        for s in range(nsteps):
            acc, m = stretch_step(self.p, self.logl, self.logp, self.betas, draws.half_idx[s], draws.zz[s],
                                  draws.rint[s], draws.factors[s], draws.lnu[s], self.orc.logl_logp_batch)
            self.min_margin = min(self.min_margin, m)
            acc_all.append(acc)
        p_before = self.p.copy() if self.D is not None else None
        n_acc, src, m = swap_sweep(self.p, self.logl, self.logp, self.betas, draws.perm, draws.lnu_swap)
        if self.D is not None:
            self.smd = swap_mean_distance(p_before, self.p, src, self.D)
        self.min_margin = min(self.min_margin, m)
        self.time += 1
        if self.adapt:
            W = self.p.shape[1]
            self.betas = adapt_ladder(self.betas, n_acc / W, self.time, self.adapt_tau, self.adapt_nu)
        return np.array(acc_all), n_acc, src

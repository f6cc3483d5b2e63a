"""ORACLE — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

NumPy restatement of the functions the reference *generates* for one model:
`my_model`, `my_likelihood`, `my_prior`.  Every
method cites the template it follows and keeps the template's operation order,
so on the same NumPy the results are bit-identical to a script emitted by the
real generator (checked by tests/golden/make_golden.py, which executes the real
generator's output in the build container and stores the vectors under
tests/golden/).

`kepler.solve` is the C restatement in kepler_oracle.c (kepler.py is not
vendored; "parity unpinned" at that boundary, see its header).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module.
"""
import numpy as np

from . import kepler_shim as kepler

TWO_PI = 2 * np.pi


def _w_from_sc(S, C, ecc, thr):
    # support/models/kep01.model:4-9 (thr 1e-6; kep07 same), kep02.model:7-12 / kep04.model:4-9 (thr 1e-5)
    if ecc < thr:
        return 0
    w = np.arccos(C / (ecc ** 0.5))
    if S < 0:
        w = 2 * np.pi - w
    return w


class RVOracle:
    """One generated script's worth of functions, for a compiled model
    (astroemperor_b200.modelspec.CompiledModel is only read as plain data)."""

    def __init__(self, cm, t, y, yerr, flag, sai=None, run_noop_ma_loop=False):
        self.cm = cm
        # moav00.model's loop has no effect on logL (see my_model) but it is what the reference EXECUTES per call;
        # run_noop_ma_loop=True runs the same statements so that the CPU baseline can time the default template
        self.run_noop_ma_loop = bool(run_noop_ma_loop)
        # emp_model.py:425-433: SAI{j}_ = my_data.iloc[:, 3 + j].values, one array per activity column
        self.SAI_ = None if sai is None else np.ascontiguousarray(sai, dtype=np.float64).reshape(len(t), -1)
        self.X_ = np.ascontiguousarray(t, dtype=np.float64)
        self.Y_ = np.ascontiguousarray(y, dtype=np.float64)
        self.YERR_ = np.ascontiguousarray(yerr, dtype=np.float64)
        self.flag = np.ascontiguousarray(flag, dtype=np.int64)
        self.ndat = len(self.X_)
        # emp_model.py:420-424: mask{n} = (my_data['Flag'] == n).values
        self.masks = [self.flag == (n + 1) for n in range(cm.n_ins)]
        # support/likelihoods/00.like:1
        self.likelihood_constant = -0.5 * np.log(2 * np.pi) * self.ndat

    # emp_model.py:709-711 / emp.py:190-193
    def full_theta(self, theta):
        cm = self.cm
        full = cm.full_init.copy()
        full[cm.free_to_full] = theta
        return full

    # ---- Keplerian templates ------------------------------------------------
    def _kep_elements(self, model, th):
        """(per, A, phase_or_tp, ecc, w, use_tp, sqrt_variant)"""
        if model == 0:  # kep00.model:2
            per, A, phase, ecc, w = th
            return per, A, phase, ecc, w, False, 0
        if model == 1:  # kep01.model:1-9
            per, A, phase, S, C = th
            ecc = S ** 2 + C ** 2
            return per, A, phase, ecc, _w_from_sc(S, C, ecc, 1e-6), False, 1
        if model == 2:  # kep02.model:1-16
            P, As, Ac, S, C = th
            per = np.exp(P)
            A = As ** 2 + Ac ** 2
            ecc = S ** 2 + C ** 2
            w = _w_from_sc(S, C, ecc, 1e-5)
            phase = np.arccos(Ac / (A ** 0.5))
            if As < 0:
                phase = 2 * np.pi - np.arccos(Ac / (A ** 0.5))
            return per, A, phase, ecc, w, False, 1
        if model == 3:  # kep03.model:1
            per, A, tp, ecc, w = th
            return per, A, tp, ecc, w, True, 1
        if model == 4:  # kep04.model:1-9
            per, A, tp, S, C = th
            ecc = S ** 2 + C ** 2
            return per, A, tp, ecc, _w_from_sc(S, C, ecc, 1e-5), True, 1
        if model == 5:  # akep00.model:1
            per, A, pha, ecc, w, _I, _Om = th
            return per, A, pha, ecc, w, False, 1
        if model == 6:  # kep06.model:2-4
            P, A, phase, ecc, w = th
            return np.exp(P), A, phase, ecc, w, False, 0
        if model == 7:  # kep07.model:2-11
            P, A, phase, S, C = th
            per = np.exp(P)
            ecc = S ** 2 + C ** 2
            return per, A, phase, ecc, _w_from_sc(S, C, ecc, 1e-6), False, 0
        raise ValueError(model)

    def _kep_rv(self, model, th, X_):
        per, A, ph, ecc, w, use_tp, variant = self._kep_elements(model, th)
        freq = 2. * np.pi / per
        if use_tp:
            M = freq * (X_ - ph)  # kep03.model:4
        else:
            M = freq * X_ + ph  # kep00.model:5
        E = kepler.solve(M, np.repeat(ecc, len(M)))
        with np.errstate(all="ignore"):
            if variant == 0:  # kep00.model:7
                f = np.arctan(((1. + ecc) / (1. - ecc)) ** 0.5 * np.tan(E / 2.)) * 2.
            else:  # kep01.model:14
                f = (np.arctan(((1. + ecc) ** 0.5 / (1. - ecc) ** 0.5) * np.tan(E / 2.)) * 2.)
            return A * (np.cos(f + w) + ecc * np.cos(w))  # kep00.model:8

    # ---- my_model: emp_model.py:706-781 ---------------------------------------
    def my_model(self, theta):
        cm = self.cm
        theta = self.full_theta(np.asarray(theta, dtype=np.float64))
        X_, Y_ = self.X_, self.Y_
        model0 = np.zeros(self.ndat)
        err20 = self.YERR_ ** 2
        for model, off in zip(cm.kep_model, cm.kep_off):
            npar = 7 if model == 5 else 5
            model0 += self._kep_rv(model, theta[off:off + npar], X_)
        if cm.acc_order:
            # support/models/acc.model:2
            model0 += np.polyval(np.concatenate([theta[cm.acc_off:cm.acc_off + cm.acc_order], [0]]),
                                 (X_ - X_[0]))
        for n in range(cm.n_ins):
            # support/models/offset00.model:3
            model0 += theta[cm.offset_off:cm.offset_off + cm.n_ins][n] * self.masks[n]
        if cm.has_jitter:
            for n in range(cm.n_ins):
                # support/models/jitter00.model:3
                err20 += self.masks[n] * theta[cm.jitter_off:cm.jitter_off + cm.n_ins][n] ** 2
        if cm.ma_mode == 2:
            # emp_model.py:757-762 + support/models/moav01.model:3-15
            residuals = Y_ - model0
            order = cm.ma_order
            theta_ma = theta[cm.ma_off:cm.ma_off + 2 * order]
            t_ = X_
            res_ = residuals
            for i in range(self.ndat):
                for c in range(order):
                    if i > c:
                        dt = abs(t_[i] - t_[i - 1 - c])
                        macoef = theta_ma[2 * c]
                        matime = theta_ma[2 * c + 1]
                        MA = macoef * np.exp(-dt / matime) * res_[i - 1 - c]
                        model0[i] += MA
                        residuals[i] -= MA
        n_sai = int(sum(getattr(cm, "sai_count", [])))
        if n_sai:
            # emp_model.py:736-745 + support/models/sai00.model:3, once per activity column
            theta_sa = theta[cm.sai_off:cm.sai_off + n_sai]
            for j in range(n_sai):
                model0 += theta_sa[j] * self.SAI_[:, j]
        for kind, off in getattr(cm, "periodic", []):
            if kind == 0:  # support/models/sinusoid00.model
                per, A, phase = theta[off:off + 3]
                freq = 2. * np.pi / per
                M = freq * X_ + phase
                model0 += A * np.cos(M)
            else:  # support/models/magneticcycle00.model
                per, A1, A2, phase1, phase2 = theta[off:off + 5]
                freq2 = 2. * np.pi / per
                freq1 = np.pi / per
                M1 = freq1 * X_ + phase1
                M2 = freq2 * X_ + phase2
                model0 += A1 * np.cos(M1) + A2 * np.cos(M2)
        # cm.ma_mode == 1 (support/models/moav00.model): `model0[mask][i] += MA`
        # writes into a temporary copy, so the block has NO effect on model0 / err20
        # (SURVEY.md §0 fact 3); nothing to do — unless the caller wants the reference's COST:
        if cm.ma_mode == 1 and self.run_noop_ma_loop:
            residuals = Y_ - model0  # emp_model.py:757-762
            order = cm.ma_order
            for n in range(cm.n_ins):  # moav00.model:4-19, once per instrument
                mask = self.masks[n]
                t_ = X_[mask]
                res_ = residuals[mask]
                theta_ma = theta[cm.ma_off:cm.ma_off + 2 * order * cm.n_ins][order * 2 * n:order * 2 * (n + 1)]
                ndat_n = int(np.sum(mask))
                if order > 0:
                    for c in range(order):
                        macoef = theta_ma[2 * c]
                        matime = theta_ma[2 * c + 1]
                        for i in range(1, ndat_n):
                            if i > c:
                                dt = abs(t_[i] - t_[i - 1 - c])
                                MA = macoef * np.exp(-dt / matime) * res_[i - 1 - c]
                                model0[mask][i] += MA      # fancy-index copy: discarded
                                residuals[mask][i] -= MA   # ditto
        return model0, err20

    # support/likelihoods/00.like:3-5
    def my_likelihood(self, theta):
        model, err2 = self.my_model(theta)
        with np.errstate(all="ignore"):
            return -0.5 * (np.sum((self.Y_ - model) ** 2 / err2 + np.log(err2))) + self.likelihood_constant

    # ---- my_prior: emp.py:182-254 + support/priors/*.prior --------------------
    @staticmethod
    def _prior(kind, x, lo, hi, a0, a1, a2, a3):
        if kind == 4:  # Fixed.prior
            return 0.
        if not (lo <= x <= hi):
            return -np.inf
        if kind == 0:  # Uniform.prior:1-5
            return a0
        if kind == 1:  # Normal.prior:8 ; a3 = np.log(s*np.sqrt(2*np.pi))
            return -0.5 * ((x - a0) / a1) ** 2 - a3 - a2
        if kind == 2:  # Jeffreys.prior:4-5 (uniform in the reference)
            return a0
        if kind == 3:  # Isotropic.prior:4
            with np.errstate(all="ignore"):
                return np.log(0.5 * np.sin(x)) - a0
        raise ValueError(kind)

    def my_prior(self, theta):
        theta = self.full_theta(np.asarray(theta, dtype=np.float64))
        lp = 0.
        for (op, kind, i0, i1, lo, hi, a0, a1, a2, a3) in self.cm.prior_ops:
            if op == 0:
                lp += self._prior(kind, theta[i0], lo, hi, a0, a1, a2, a3)
            elif op == 1:
                if lp == -np.inf:
                    return lp
            else:
                x = theta[i0] ** 2 + theta[i1] ** 2
                lp += self._prior(kind, x, lo, hi, a0, a1, a2, a3)
        return lp

    # ---- batch helpers for tests / bench ---------------------------------------
    def logl_logp_batch(self, thetas, skip_bad=True):
        """emcee semantics: the likelihood is not evaluated where the prior is -inf."""
        thetas = np.asarray(thetas, dtype=np.float64).reshape(-1, self.cm.ndim_free)
        ll = np.empty(len(thetas))
        lp = np.empty(len(thetas))
        for i, th in enumerate(thetas):
            lp[i] = self.my_prior(th)
            if skip_bad and lp[i] == -np.inf:
                ll[i] = -np.inf
            else:
                ll[i] = self.my_likelihood(th)
        return ll, lp

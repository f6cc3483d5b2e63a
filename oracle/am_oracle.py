"""ORACLE — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

NumPy restatement of the Hipparcos-Gaia astrometric block the reference generates
(`ReddModel._write_model_AM`, emp_model.py:1232-1672; constants
support/astrometry/constants.scr; combined with the RV term in
support/likelihoods/a00.like:3-8).  Same functions, same operation order and the
same `np.longdouble` (x87 80-bit) intermediates as the generated script, so on
the same NumPy it is bit-identical to it (checked against the real generator's
output by tests/golden/make_golden.py -> tests/golden/c3_*.npz).

`am` is a dict of plain arrays (the constants emp_model.py:610-702 loads):
  catalogs[3,7], time_hipp, cpsi_hipp, spsi_hipp, epoch_hipp, parf_hipp, res_hipp, sres_hipp,
  time_gost, cpsi_gost, spsi_gost, parf_gost, mask_gdr2, mask_gdr3 (bool), gsv2[5,n2], gsv3[5,n3],
  inv_cov[3,5,5], log_det_cov[3], astro_gost[2,5], common_t

Only tests/, __graft_entry__.smoke() and bench.py may import this module.
"""
import numpy as np

from . import kepler_shim as kepler

# support/astrometry/constants.scr:4-10
LOG_2PI = np.log(2 * np.pi)
MAS_PER_DEG = np.longdouble(3.6e6)
DEG2MAS = np.longdouble(3.6e6)
PC_PER_KPC = 1e3
DAY_PER_YEAR = 365.25
PC2AU = 206265
AUYR2KMS = 4.74047


def mas2deg(x):  # constants.scr:13-14
    return np.longdouble(x) / MAS_PER_DEG


def bl2xyz(b_rad, l_rad):  # constants.scr:17-23
    x = np.cos(b_rad) * np.cos(l_rad)
    y = np.cos(b_rad) * np.sin(l_rad)
    z = np.sin(b_rad)
    return np.array([x, y, z])


def xyz2bl_vec(x, y, z):  # constants.scr:26-33
    b = np.arctan2(z, np.sqrt(x ** 2 + y ** 2))
    ind = b > np.pi / 2
    if (np.sum(ind) > 0):
        b[ind] = b[ind] - np.pi
    l = np.arctan2(y, x) % (2 * np.pi)
    return b, l


def thiele_innes(omega, Omega, sinI, cosI):  # emp_model.py:1253-1270
    sinOM, cosOM = np.sin(Omega), np.cos(Omega)
    sinom, cosom = np.sin(omega), np.cos(omega)
    A = (cosom * cosOM - sinom * sinOM * cosI)
    B = (cosom * sinOM + sinom * cosOM * cosI)
    F = (-sinom * cosOM - cosom * sinOM * cosI)
    G = (-sinom * sinOM + cosom * cosOM * cosI)
    C = sinom * sinI
    H = cosom * sinI
    return A, B, F, G, C, H


class AMOracle:
    def __init__(self, cm, am):
        self.cm = cm
        a = {k: np.asarray(v) for k, v in am.items()}
        self.AM_catalogs_ = a["catalogs"].astype(np.float64)
        self.AM_iref_ = -1
        times = self.AM_catalogs_[:, 0]
        self.AM_ref_epoch_ = times[self.AM_iref_]
        self.AM_catalogs_times_refed = times - self.AM_ref_epoch_
        self.AM_catalogs_obs_ref = self.AM_catalogs_[self.AM_iref_, 1:]
        self.AM_PLX_ref = self.AM_catalogs_[:, 3][self.AM_iref_]
        self.time_iad_hipp = a["time_hipp"]
        self.time_iad_gost = a["time_gost"]
        self.time_iad_gost_refed = self.time_iad_gost - self.AM_ref_epoch_
        self.time_iad_all = np.concatenate([self.time_iad_hipp, self.time_iad_gost])
        self.CPSI_HIPP_, self.SPSI_HIPP_ = a["cpsi_hipp"], a["spsi_hipp"]
        self.EPOCH_HIPP_, self.PARF_HIPP_ = a["epoch_hipp"], a["parf_hipp"]
        self.RES_HIPP_, self.SRES_HIPP_ = a["res_hipp"], a["sres_hipp"]
        self.CPSI_GOST_, self.SPSI_GOST_, self.PARF_GOST_ = a["cpsi_gost"], a["spsi_gost"], a["parf_gost"]
        self.N_HIPP, self.N_GOST = len(self.time_iad_hipp), len(self.time_iad_gost)
        self.N_IAD = self.N_HIPP + self.N_GOST
        self.AM_GSV = {"GDR2": a["gsv2"], "GDR3": a["gsv3"]}
        self.AM_inv_COV, self.AM_log_det_COV = a["inv_cov"], a["log_det_cov"]
        self.AM_astro_gost = a["astro_gost"]
        self.GAIA_CATS = dict(GDR2=dict(mask=a["mask_gdr2"].astype(bool), row=0, cov_idx=1),
                              GDR3=dict(mask=a["mask_gdr3"].astype(bool), row=1, cov_idx=2))
        self.common_t = float(np.asarray(a["common_t"]).reshape(-1)[0])

    # emp_model.py:1312-1368 (only rows 0..2 are consumed downstream; all six are formed like the script)
    def calc_astro_new(self, theta, plx):
        per, K, pha, ecc, omega, I, Omega = theta
        sinI, cosI = np.sin(I), np.cos(I)
        sqrt1_e2 = np.sqrt(1 - ecc ** 2)
        freq = 2. * np.pi / per
        M = freq * (self.time_iad_all - self.common_t) + pha
        E = kepler.solve(M, np.repeat(ecc, self.N_IAD))
        f = (np.arctan(((1. + ecc) ** 0.5 / (1. - ecc) ** 0.5) * np.tan(E / 2.)) * 2.)
        A, B, F, G, C, H = thiele_innes(omega, Omega, sinI, cosI)
        X = np.cos(E) - ecc
        Y = sqrt1_e2 * np.sin(E)
        VX = -np.sin(f)
        VY = np.cos(f) + ecc
        alpha0 = K / sinI / PC_PER_KPC / AUYR2KMS
        beta0 = per / DAY_PER_YEAR * (K / PC_PER_KPC / AUYR2KMS) * sqrt1_e2 / (2 * np.pi) / sinI
        alpha = -alpha0 * plx
        beta = -beta0 * plx
        rasP = beta * (B * X + G * Y)
        decP = beta * (A * X + F * Y)
        plxP = -beta * (C * X + H * Y) * plx / 206265e3
        pmrasP = alpha * (B * VX + G * VY)
        pmdecP = alpha * (A * VX + F * VY)
        rv = alpha0 * (C * VX + H * VY)
        return np.array([rasP, decP, plxP, pmrasP, pmdecP, rv * AUYR2KMS])

    # emp_model.py:1290-1304
    def astrometry_iad_model(self, theta):
        cm = self.cm
        model = np.zeros((6, self.N_IAD))
        for m, off in zip(cm.kep_model, cm.kep_off):
            theta_am_off = theta[cm.am_offset_off:cm.am_offset_off + 5]
            plx0 = self.AM_PLX_ref - theta_am_off[2]
            model += self.calc_astro_new(theta[off:off + 7], plx0)
        return model

    # emp_model.py:1403-1415
    def get_deltas_HIPP(self, bary):
        ref_ra, ref_dec, ref_plx_mas, ref_pmra, ref_pmde = self.AM_catalogs_[0, 1:-1]
        mean_dec = 0.5 * (ref_dec + bary[1])
        dra = (bary[0] - ref_ra) * np.cos(np.deg2rad(mean_dec)) * DEG2MAS  # dra_star_mas, :1243-1247
        dde = (bary[1] - ref_dec) * DEG2MAS  # ddec_mas, :1249-1250
        dplx = (bary[2] - ref_plx_mas)
        dpmra = (bary[3] - ref_pmra)
        dpmde = (bary[4] - ref_pmde)
        return np.array([dra, dde, dplx, dpmra, dpmde])

    # emp_model.py:1418-1431
    def get_deltas_GOST(self, bary, epoch):
        ref_ra, ref_dec = self.AM_catalogs_[2, 1:3]
        dec = bary[1] + mas2deg(epoch[1, :])
        dec_rad = np.deg2rad(dec)
        cos_dec = np.cos(dec_rad)
        dra = ((bary[0] - ref_ra) * cos_dec * MAS_PER_DEG) + epoch[0, :]
        ddec = (dec - ref_dec) * MAS_PER_DEG
        dplx = bary[2] + epoch[2, :]
        return np.array([dra, ddec, dplx])

    # emp_model.py:1436-1457
    def compute_abs_signal_hipp(self, epoch, bary):
        dra0, dde0 = epoch[0], epoch[1]
        dplx0, dpmra0, dpmde0 = np.zeros(self.N_HIPP), np.zeros(self.N_HIPP), np.zeros(self.N_HIPP)
        deltas = self.get_deltas_HIPP(bary)
        dra0 += deltas[0]
        dde0 += deltas[1]
        dplx0 += deltas[2]
        dpmra0 += deltas[3]
        dpmde0 += deltas[4]
        return (self.CPSI_HIPP_ * (dra0 + dpmra0 * self.EPOCH_HIPP_) +
                self.SPSI_HIPP_ * (dde0 + dpmde0 * self.EPOCH_HIPP_) +
                self.PARF_HIPP_ * dplx0)

    # emp_model.py:1460-1475
    def compute_abs_signal_gost(self, epoch_, bary):
        dra0, dde0, dplx0 = self.get_deltas_GOST(bary, epoch_)
        return (self.SPSI_GOST_ * (dra0) + self.CPSI_GOST_ * (dde0) + self.PARF_GOST_ * dplx0)

    # emp_model.py:1483-1495
    def model_barycenter(self, theta):
        theta0 = theta.copy()
        dec_ref = self.AM_catalogs_obs_ref[1]
        theta0[0] = mas2deg(theta0[0]) / np.cos(np.deg2rad(dec_ref))
        theta0[1] = mas2deg(theta0[1])
        theta0 = np.append(theta0, 0)
        obs = self.AM_catalogs_obs_ref - theta0
        return self.obs_lin_prop_PA(obs)

    # emp_model.py:1498-1573
    def obs_lin_prop_PA(self, obs):
        RA, DE, PLX, PMRA, PMDE, RV = obs
        ra = np.deg2rad(RA)
        de = np.deg2rad(DE)
        plx, pmra, pmde, rv = PLX, PMRA, PMDE, RV
        cosde, sinde = np.cos(de), np.sin(de)
        cosra, sinra = np.cos(ra), np.sin(ra)
        d = 1 / plx
        x, y, z = bl2xyz(de, ra) * d * PC_PER_KPC
        vra = pmra * d
        vde = pmde * d
        vr = rv / AUYR2KMS
        vx_equ = vr * cosde * cosra - vde * sinde * cosra - vra * sinra
        vy_equ = vr * cosde * sinra - vde * sinde * sinra + vra * cosra
        vz_equ = vr * sinde + vde * cosde
        time_factor = self.AM_catalogs_times_refed / (DAY_PER_YEAR * PC2AU)
        x1 = x + vx_equ * time_factor
        y1 = y + vy_equ * time_factor
        z1 = z + vz_equ * time_factor
        de1_rad, ra1_rad = xyz2bl_vec(x1, y1, z1)
        d1 = np.sqrt(x1 ** 2 + y1 ** 2 + z1 ** 2) * 1e-3
        ra1 = np.rad2deg(ra1_rad)
        de1 = np.rad2deg(de1_rad)
        cosra1, sinra1 = np.cos(ra1_rad), np.sin(ra1_rad)
        cosde1, sinde1 = np.cos(de1_rad), np.sin(de1_rad)
        vv = np.array([vx_equ, vy_equ, vz_equ])
        zeros = np.zeros_like(cosra1)
        ones = np.ones_like(cosra1)
        rotz = np.array([[cosra1, sinra1, zeros], [-sinra1, cosra1, zeros], [zeros, zeros, ones]]).transpose(2, 0, 1)
        roty = np.array([[cosde1, zeros, sinde1], [zeros, ones, zeros], [-sinde1, zeros, cosde1]]).transpose(2, 0, 1)
        rot = np.matmul(roty, rotz)
        vequ = np.einsum('ijk,k->ij', rot, vv)
        pmra1 = vequ[:, 1] / d1
        pmde1 = vequ[:, 2] / d1
        rv1 = vequ[:, 0] * AUYR2KMS
        return np.column_stack((ra1, de1, 1 / d1, pmra1, pmde1, rv1))

    # emp_model.py:1576-1594
    def obs_lin_prop_simple(self, obs):
        RA, DEC, PLX, PMRA, PMDEC, RV = obs
        ra = np.deg2rad(RA)
        dec = np.deg2rad(DEC)
        plx, pmra, pmdec, rv = PLX, PMRA, PMDEC, RV
        t = self.time_iad_gost_refed
        decs = dec + pmdec * t / DAY_PER_YEAR / 206265e3
        ras = ra + pmra * t / DAY_PER_YEAR / np.cos(decs) / 206265e3
        return np.array([np.rad2deg(ras), np.rad2deg(decs), np.repeat(plx, self.N_GOST),
                         np.repeat(pmra, self.N_GOST), np.repeat(pmdec, self.N_GOST),
                         np.repeat(rv, self.N_GOST)]).T

    @staticmethod
    def gaussian_loglike_iid(residuals, var):  # emp_model.py:1602-1605
        n = residuals.size
        return -0.5 * (np.sum(residuals ** 2 / var + np.log(var)) + n * LOG_2PI)

    @staticmethod
    def gaussian_loglike_mvn(residuals, inv_cov, log_det_cov, jitter_sq=1.0):  # emp_model.py:1608-1615
        n = residuals.size
        quad = residuals @ inv_cov @ residuals / jitter_sq
        return -0.5 * (quad + n * np.log(jitter_sq) + log_det_cov + n * LOG_2PI)

    # emp_model.py:1618-1624
    def _prepare_gost_inputs(self, coord, barycenter):
        epoch_g = coord[:3, -self.N_GOST:]
        barycenter_g = barycenter[self.AM_iref_, :]
        bary_g = self.obs_lin_prop_simple(barycenter_g).T[:3]
        return self.compute_abs_signal_gost(epoch_g, bary_g)

    # emp_model.py:1633-1671.  NOTE: like the generated script this indexes the theta it is given
    # with FULL-theta slices; the reference passes the un-expanded theta (a00.like:7), which is only
    # meaningful when no parameter is fixed.
    def loglike_AM(self, theta):
        cm = self.cm
        theta = np.asarray(theta, dtype=np.float64)
        theta_am_off = theta[cm.am_offset_off:cm.am_offset_off + 5]
        J_H, J_G = theta[cm.am_jitter_off:cm.am_jitter_off + 2]
        ll = 0
        coor = self.astrometry_iad_model(theta)
        bary = self.model_barycenter(theta_am_off)
        coor_h = coor[:, :self.N_HIPP]
        bary_h = bary[0, :]
        abs_hipp = self.compute_abs_signal_hipp(coor_h, bary_h)
        res_hipp = self.RES_HIPP_ - abs_hipp
        var_hipp = self.SRES_HIPP_ ** 2 + J_H ** 2
        ll += self.gaussian_loglike_iid(res_hipp, var_hipp)
        abs_gost = self._prepare_gost_inputs(coor, bary)
        for cat, meta in self.GAIA_CATS.items():
            params = self.AM_GSV[cat] @ abs_gost[meta['mask']]
            res = self.AM_astro_gost[meta['row'], :] - params
            inv_cov = self.AM_inv_COV[meta['cov_idx']]
            log_det_cov = self.AM_log_det_COV[meta['cov_idx']]
            ll += self.gaussian_loglike_mvn(res, inv_cov, log_det_cov, jitter_sq=J_G ** 2)
        return ll

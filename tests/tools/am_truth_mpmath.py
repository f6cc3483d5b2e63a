#!/usr/bin/env python3
"""40-digit evaluation of the joint RV + Hipparcos-Gaia log-likelihood (BASELINE config 3) for the golden thetas.

SURVEY.md §8a row A11 / §0 fact 4: the reference evaluates `loglike_AM` partly in x87 80-bit arithmetic and its
value is ill-conditioned (a one-ulp change of the propagated barycentre moves it by ~1e-9 relative,
tests/test_am.py::test_am_value_is_ill_conditioned), so "within 1e-10 of the reference's float" is not a
well-posed bar for ANY other implementation.  What is well posed: the exact value of the same formulas on the same
inputs.  This script evaluates the formulas of `my_likelihood` (support/likelihoods/a00.like:3-8: akep00.model +
acc.model + offset00.model + jitter00.model + 00.like, and loglike_AM with its helpers, emp_model.py:1232-1672) in
mpmath at 40 significant digits, treating every input (data, constants, theta) as the exact value of its double,
and stores the result as (hi, lo) double pairs next to the reference's own value in
tests/golden/c3_am_truth.npz.  tests/test_am.py then holds the device to
    |device - truth| <= max(|reference - truth|, 1e-10 |truth|).

Usage: python tests/tools/am_truth_mpmath.py        (a few minutes; needs mpmath, no reference checkout)
"""
import os
import sys

import mpmath as mp
import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))

mp.mp.dps = 40
F = lambda x: mp.mpf(float(x))  # the exact value of a double
PI = mp.pi
# support/astrometry/constants.scr:4-10 (the doubles the reference's literals denote)
MAS_PER_DEG, PC_PER_KPC, DAY_PER_YEAR, PC2AU, AUYR2KMS = F(3.6e6), F(1e3), F(365.25), F(206265), F(4.74047)
C206265E3 = F(206265e3)
LOG_2PI = mp.log(2 * PI)
DEG = PI / 180


def kepler(M, e):
    """Root of E - e sin E = M at working precision (Newton from a double-precision start)."""
    Mr = M - 2 * PI * mp.floor(M / (2 * PI))
    E = Mr + e * mp.sin(Mr) if e < F(0.8) else PI
    for _ in range(200):
        dE = (E - e * mp.sin(E) - Mr) / (1 - e * mp.cos(E))
        E -= dE
        if abs(dE) < mp.mpf(10) ** (-38):
            break
    return E + (M - Mr)  # the generated script keeps E on M's branch only through periodic functions


def true_anomaly(E, e):  # akep00.model: (1+e)**0.5/(1-e)**0.5
    return 2 * mp.atan(mp.sqrt(1 + e) / mp.sqrt(1 - e) * mp.tan(E / 2))


def rv_loglike(cm, g, th):
    t, y, yerr, flag = g["t"], g["y"], g["yerr"], g["flag"]
    n = len(t)
    model = [mp.mpf(0)] * n
    for off in cm.kep_off:
        per, A, pha, e, w = (th[off + i] for i in range(5))
        freq = 2 * PI / per
        for i in range(n):
            E = kepler(freq * F(t[i]) + pha, e)
            f = true_anomaly(E, e)
            model[i] += A * (mp.cos(f + w) + e * mp.cos(w))
    t0 = F(t[0])
    for i in range(n):
        x = F(t[i]) - t0
        acc = mp.mpf(0)
        for j in range(cm.acc_order):  # np.polyval([a_n .. a_1, 0], x)
            acc = acc * x + th[cm.acc_off + j]
        acc = acc * x
        model[i] += acc + th[cm.offset_off + int(flag[i]) - 1]
    ll = mp.mpf(0)
    for i in range(n):
        err2 = F(yerr[i]) ** 2 + (th[cm.jitter_off + int(flag[i]) - 1] ** 2 if cm.has_jitter else 0)
        ll += (F(y[i]) - model[i]) ** 2 / err2 + mp.log(err2)
    return -ll / 2 - mp.log(2 * PI) * n / 2


def thiele_innes(om, Om, sinI, cosI):
    sO, cO, so, co = mp.sin(Om), mp.cos(Om), mp.sin(om), mp.cos(om)
    return (co * cO - so * sO * cosI, co * sO + so * cO * cosI, -so * cO - co * sO * cosI,
            -so * sO + co * cO * cosI, so * sinI, co * sinI)


def lin_prop_pa(obs, times_refed):
    """obs_lin_prop_PA (emp_model.py:1498-1573) for every catalogue epoch: rows (ra, de, plx, pmra, pmde, rv)."""
    RA, DE, plx, pmra, pmde, rv = obs
    ra, de = RA * DEG, DE * DEG
    cde, sde, cra, sra = mp.cos(de), mp.sin(de), mp.cos(ra), mp.sin(ra)
    d = 1 / plx
    x, y, z = cde * cra * d * PC_PER_KPC, cde * sra * d * PC_PER_KPC, sde * d * PC_PER_KPC
    vra, vde, vr = pmra * d, pmde * d, rv / AUYR2KMS
    vx = vr * cde * cra - vde * sde * cra - vra * sra
    vy = vr * cde * sra - vde * sde * sra + vra * cra
    vz = vr * sde + vde * cde
    out = []
    for tr in times_refed:
        tf = tr / (DAY_PER_YEAR * PC2AU)
        x1, y1, z1 = x + vx * tf, y + vy * tf, z + vz * tf
        b = mp.atan2(z1, mp.sqrt(x1 ** 2 + y1 ** 2))
        l = mp.atan2(y1, x1)
        l = l - 2 * PI * mp.floor(l / (2 * PI))
        d1 = mp.sqrt(x1 ** 2 + y1 ** 2 + z1 ** 2) / 1000
        cr, sr, cd, sd = mp.cos(l), mp.sin(l), mp.cos(b), mp.sin(b)
        # rot = roty @ rotz applied to (vx, vy, vz)
        u0, u1, u2 = cr * vx + sr * vy, -sr * vx + cr * vy, vz
        v0, v1, v2 = cd * u0 + sd * u2, u1, -sd * u0 + cd * u2
        out.append((l / DEG, b / DEG, 1 / d1, v1 / d1, v2 / d1, v0 * AUYR2KMS))
    return out


def am_loglike(cm, am, th):
    cat = [[F(v) for v in row] for row in am["catalogs"]]
    ref = cat[-1]
    ref_epoch = ref[0]
    times_refed = [row[0] - ref_epoch for row in cat]
    t_h, t_g = [F(v) for v in am["time_hipp"]], [F(v) for v in am["time_gost"]]
    common_t = F(np.asarray(am["common_t"]).reshape(-1)[0])
    nh, ng = len(t_h), len(t_g)
    off = [th[cm.am_offset_off + i] for i in range(5)]
    J_H, J_G = th[cm.am_jitter_off], th[cm.am_jitter_off + 1]
    plx0 = ref[3] - off[2]
    # astrometry_iad_model / calc_astro_new: reflex offsets (ras, dec, plx) at every IAD epoch
    ras, dec, plxv = [mp.mpf(0)] * (nh + ng), [mp.mpf(0)] * (nh + ng), [mp.mpf(0)] * (nh + ng)
    for o in cm.kep_off:
        per, K, pha, e, om, I, Om = (th[o + i] for i in range(7))
        sinI, cosI = mp.sin(I), mp.cos(I)
        sq = mp.sqrt(1 - e ** 2)
        freq = 2 * PI / per
        A, B, Fc, G, C, H = thiele_innes(om, Om, sinI, cosI)
        beta0 = per / DAY_PER_YEAR * (K / PC_PER_KPC / AUYR2KMS) * sq / (2 * PI) / sinI
        beta = -beta0 * plx0
        for i, tt in enumerate(t_h + t_g):
            E = kepler(freq * (tt - common_t) + pha, e)
            X, Y = mp.cos(E) - e, sq * mp.sin(E)
            ras[i] += beta * (B * X + G * Y)
            dec[i] += beta * (A * X + Fc * Y)
            plxv[i] += -beta * (C * X + H * Y) * plx0 / C206265E3
    # model_barycenter
    dec_ref = ref[2]
    th0 = [off[0] / MAS_PER_DEG / mp.cos(dec_ref * DEG), off[1] / MAS_PER_DEG, off[2], off[3], off[4], mp.mpf(0)]
    obs = [ref[1 + i] - th0[i] for i in range(6)]
    bary = lin_prop_pa(obs, times_refed)
    # Hipparcos: compute_abs_signal_hipp + iid Gaussian
    bh = bary[0]
    h = cat[0]
    mean_dec = (h[2] + bh[1]) / 2
    d_ra = (bh[0] - h[1]) * mp.cos(mean_dec * DEG) * MAS_PER_DEG
    d_de = (bh[1] - h[2]) * MAS_PER_DEG
    d_plx, d_pmra, d_pmde = bh[2] - h[3], bh[3] - h[4], bh[4] - h[5]
    ll = mp.mpf(0)
    acc = mp.mpf(0)
    for i in range(nh):
        ep = F(am["epoch_hipp"][i])
        ab = (F(am["cpsi_hipp"][i]) * (ras[i] + d_ra + d_pmra * ep) + F(am["spsi_hipp"][i]) * (dec[i] + d_de + d_pmde * ep)
              + F(am["parf_hipp"][i]) * d_plx)
        var = F(am["sres_hipp"][i]) ** 2 + J_H ** 2
        acc += (F(am["res_hipp"][i]) - ab) ** 2 / var + mp.log(var)
    ll += -(acc + nh * LOG_2PI) / 2
    # Gaia: obs_lin_prop_simple of the GDR3-epoch barycentre, get_deltas_GOST, refit, MVN
    RA, DEC, PLX, PMRA, PMDEC, RV = bary[-1]
    abs_g = []
    for j in range(ng):
        t = t_g[j] - ref_epoch
        decs = DEC * DEG + PMDEC * t / DAY_PER_YEAR / C206265E3
        rass = RA * DEG + PMRA * t / DAY_PER_YEAR / mp.cos(decs) / C206265E3
        b_ra, b_de = rass / DEG, decs / DEG
        de = b_de + dec[nh + j] / MAS_PER_DEG
        dra = (b_ra - ref[1]) * mp.cos(de * DEG) * MAS_PER_DEG + ras[nh + j]
        dde = (de - ref[2]) * MAS_PER_DEG
        dplx = PLX + plxv[nh + j]
        abs_g.append(F(am["spsi_gost"][j]) * dra + F(am["cpsi_gost"][j]) * dde + F(am["parf_gost"][j]) * dplx)
    for row, (mask, gsv, ci) in enumerate(((am["mask_gdr2"], am["gsv2"], 1), (am["mask_gdr3"], am["gsv3"], 2))):
        idx = np.flatnonzero(np.asarray(mask, dtype=bool))
        res = []
        for a in range(5):
            s = mp.mpf(0)
            for m, jj in enumerate(idx):
                s += F(gsv[a][m]) * abs_g[jj]
            res.append(F(am["astro_gost"][row][a]) - s)
        quad = mp.mpf(0)
        for a in range(5):
            for b in range(5):
                quad += res[a] * F(am["inv_cov"][ci][a][b]) * res[b]
        jsq = J_G ** 2
        ll += -(quad / jsq + 5 * mp.log(jsq) + F(am["log_det_cov"][ci]) + 5 * LOG_2PI) / 2
    return ll


def main():
    from conftest import load_golden
    out = {}
    for name in ("c3_hip21850_am_k1", "c3_hip21850_am_k2"):
        g, spec = load_golden(name)
        cm = spec.compile()
        am = {k[3:]: g[k] for k in g.files if k.startswith("am_")}
        fin = np.flatnonzero(np.isfinite(g["logp"]))
        hi, lo = np.full(len(g["logp"]), np.nan), np.full(len(g["logp"]), np.nan)
        worst = 0.0
        for i in fin:
            th = [F(v) for v in g["thetas"][i]]
            tot = rv_loglike(cm, g, th) + am_loglike(cm, am, th)
            hi[i] = float(tot)
            lo[i] = float(tot - mp.mpf(hi[i]))
            rel = abs((mp.mpf(float(g["logl"][i])) - tot) / tot)
            worst = max(worst, float(rel))
            print(f"{name}[{i}] truth {mp.nstr(tot, 22)}  reference {g['logl'][i]!r}  rel {float(rel):.2e}", flush=True)
        print(f"{name}: max |reference - truth| / |truth| = {worst:.3e}")
        out[name + "_hi"], out[name + "_lo"] = hi, lo
    np.savez(os.path.join(REPO, "tests", "golden", "c3_am_truth.npz"), **out)


if __name__ == "__main__":
    main()

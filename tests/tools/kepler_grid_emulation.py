"""CPU emulation of the table-driven Kepler core (likelihood kernel v6): accuracy study.

Emulates, with NumPy float32/float64 arithmetic (no FMA, MUFU results perturbed by 2^-22), the scheme
of astroemperor_b200/csrc/emp_device.cuh `kepler_grid`:
  FP32 : Markley starter -> grid point k = rint(128 E0), El = E0 - k/128,
         Halley step in delta-space (polynomial in delta with table values sin/cos(k/128))
  FP64 : residual g(delta) and g'(delta) from short sin/cos polynomials, one Newton (+ optional
         Halley) correction, first-order rotation to sin E / 1 - cos E of the root.
and compares E, sin E, 1-cos E and the RV term with an 80-bit Newton solution.

Development aid (not a test): python tests/tools/kepler_grid_emulation.py
"""
import numpy as np

f32 = np.float32
PI = np.pi
F1 = 3.0 * PI / (PI - 6.0 / PI)
F2 = 1.6 / (PI - 6.0 / PI)
H = 1.0 / 128.0
rng = np.random.default_rng(0)


def mufu(x):
    """approximate-unit result: relative perturbation up to 2^-22"""
    return (x * (1.0 + rng.uniform(-1, 1, size=np.shape(x)) * 2.0 ** -22)).astype(f32)


def starter_f32(Mr, e):
    M = Mr.astype(f32)
    ef = e.astype(f32)
    omef = (1.0 - e).astype(f32)
    c2f = (F2 / (1.0 + e)).astype(f32)
    ome3f = (3.0 * (1.0 - e)).astype(f32)
    M2 = M * M
    alpha = c2f * (f32(PI) - M) + f32(F1)
    d = alpha * ef + ome3f
    ad = alpha * d
    r = (f32(3.0) * ad * (d - omef) + M2) * M
    q = f32(2.0) * ad * omef - M2
    q2 = q * q
    x = np.abs(r) + mufu(np.sqrt(q2 * q + r * r))
    w = mufu(np.exp2(f32(2.0 / 3.0) * mufu(np.log2(x))))
    den0 = w * (w + q) + q2
    return (f32(2.0) * r * w + M * den0) * mufu(f32(1.0) / (den0 * d))


def truth(Mr, e):
    M = Mr.astype(np.longdouble)
    el = e.astype(np.longdouble)
    E = M + el * np.sin(M)
    E = np.where(el > 0.8, np.longdouble(PI), E)
    for _ in range(200):
        f = E - el * np.sin(E) - M
        fp = 1 - el * np.cos(E)
        E = np.clip(E - f / fp, 0, np.longdouble(PI))
    for _ in range(4):
        f = E - el * np.sin(E) - M
        fp = 1 - el * np.cos(E)
        E = E - f / fp
    return E


def v6(Mr, e, w=None):
    """mirror of emp_device.cuh kep_rv_grid (FMA contractions ignored); returns E, sin E, cos E, RV/A"""
    E0f = starter_f32(Mr, e)
    kf = np.rint(E0f * f32(128.0))
    El = (E0f - kf * f32(H)).astype(f32)
    k = kf.astype(np.int64)
    Eh = k * H
    sh = np.sin(Eh.astype(np.longdouble)).astype(np.float64)
    ch = np.cos(Eh.astype(np.longdouble)).astype(np.float64)
    shf = sh.astype(f32)
    chf = ch.astype(f32)
    c = e * sh + (Mr - Eh)
    cf = c.astype(f32)
    ef = e.astype(f32)
    c6, s2 = (ch / 6.0).astype(f32), (0.5 * sh).astype(f32)  # table entries (sin/2, cos/6)
    P = -El * (El * c6 + s2) + chf
    g0 = -(ef * El) * P + (El - cf)
    Q = -El * ((f32(0.5) * El) * chf + shf) + chf
    g1 = -ef * Q + f32(1.0)
    g2h = (f32(0.5) * ef) * (El * chf + shf)
    r = mufu(f32(1.0) / g1)
    dn = g0 * r
    d1 = -dn * (dn * (g2h * r)) + (El - dn)
    df = d1.astype(np.float64)
    d2 = df * df
    sl = (df * d2) * (d2 * (1.0 / 120.0) - 1.0 / 6.0) + df
    cm = d2 * (d2 * (-1.0 / 24.0) + 0.5)  # d^6/720 < 1.1e-17 is dropped on the device too
    w1 = ch * sl - sh * cm
    w2 = sh * sl + ch * cm
    sEf = sh + w1
    cEf = ch - w2
    g = -e * w1 + (df - c)
    gp = -e * cEf + 1.0
    y1 = (1.0 / gp) * (1.0 + rng.uniform(-1, 1, size=gp.shape) * 2.0 ** -45)
    dd = -g * y1
    sE = cEf * dd + sEf
    cE = -sEf * dd + cEf
    den = -e * cE + 1.0
    y2 = y1 * (-den * y1 + 1.0) + y1
    E = Eh + (df + dd)
    out = None
    if w is not None:
        b1 = np.cos(w) * ((1.0 - e) * (1.0 + e))  # A cos w (1 - e^2): the additive constant folded in
        a2 = -np.sin(w) * np.sqrt((1.0 - e) * (1.0 + e))
        out = (b1 * cE + a2 * sE) * y2
    return E, sE, cE, out, np.abs(dd), np.abs(df)


def oracle(Mr, e):
    import os, sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
    from oracle import kepler_shim
    return kepler_shim.solve(Mr, e)


def rv(sinE, cEm, e, w):
    ome = 1.0 - e
    a1 = np.cos(w)
    a2 = -np.sin(w) * np.sqrt(ome * (1.0 + e))
    a3 = e * np.cos(w)
    return (a1 * (ome - cEm) + a2 * sinE) / (ome + e * cEm) + a3


def main():
    n = 400000
    for emax, name in [(0.5, "e<0.5"), (0.9, "e<0.9"), (0.98, "e<0.98"), (0.999, "e<0.999")]:
        e = rng.uniform(0, emax, n)
        e[: n // 10] = emax * (1 - rng.uniform(size=n // 10) ** 2 * 0.02)  # crowd the upper end
        # mix of uniform M and M crowded near periapsis
        Mr = np.where(rng.uniform(size=n) < 0.5, rng.uniform(0, PI, n), PI * rng.uniform(size=n) ** 4)
        Mr = np.maximum(Mr, 1e-12)
        Et = truth(Mr, e)
        w = rng.uniform(0, 2 * PI, n)
        rvt = rv(np.sin(Et), 1 - np.cos(Et), e.astype(np.longdouble), w.astype(np.longdouble))
        Eo = oracle(Mr, e)
        rvo = rv(np.sin(Eo), 1 - np.cos(Eo), e, w)
        print(f"--- {name}:  oracle max|dE| {np.max(np.abs(Eo - Et)):.2e}  max|dRV/A| {np.max(np.abs(rvo - rvt)):.2e}"
              f"  rms {np.sqrt(np.mean((rvo - rvt).astype(np.float64) ** 2)):.2e}")
        E, sE, cE, rvv, dd, dl = v6(Mr, e, w)
        dE = np.abs(E - Et).astype(np.float64)
        drv = np.abs(rvv - rvt).astype(np.float64)
        i = np.argmax(drv)
        print(f"    grid core:  max|dE| {dE.max():.2e}  max|dRV/A| {drv.max():.2e} (e={e[i]:.5f} M={Mr[i]:.3e})"
              f"  rms {np.sqrt(np.mean(drv**2)):.2e}  max|delta| {dl.max():.2e}  max|dd| {dd.max():.2e}")


if __name__ == "__main__":
    main()

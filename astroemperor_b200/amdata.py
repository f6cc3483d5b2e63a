"""Hipparcos-Gaia astrometry constants: host-side preparation and hand-off to the C-ABI.

The reference prepares these once per target in `DataWrapper` (qol_utils.py:305-446:
GOST epoch filtering and dead-time masks, Gaia 5-parameter solution vectors by pinv,
catalogue covariance inverse / log-det) and the generated script loads them
(emp_model.py:610-702).  The device block only needs the resulting arrays:

  catalogs[3,7]   ref_epoch, ra, dec, parallax, pmra, pmdec, radial_velocity (Hipparcos, GDR2, GDR3)
  time_hipp, cpsi_hipp, spsi_hipp, epoch_hipp, parf_hipp, res_hipp, sres_hipp   [n_hipp]
  time_gost, cpsi_gost, spsi_gost, parf_gost                                    [n_gost]
  mask_gdr2, mask_gdr3 (bool [n_gost]); gsv2[5,n2], gsv3[5,n3]; inv_cov[3,5,5]; log_det_cov[3]
  astro_gost[2,5]; common_t
"""
from __future__ import annotations

import ctypes
import os
from typing import Dict, Optional

import numpy as np

AM_KEYS = ("catalogs", "time_hipp", "cpsi_hipp", "spsi_hipp", "epoch_hipp", "parf_hipp", "res_hipp", "sres_hipp",
           "time_gost", "cpsi_gost", "spsi_gost", "parf_gost", "mask_gdr2", "mask_gdr3", "gsv2", "gsv3",
           "inv_cov", "log_det_cov", "astro_gost", "common_t")


def am_arrays_from_namespace(ns) -> Dict[str, np.ndarray]:
    """Pull the constants out of an executed generated script (tests/golden/make_golden.py)."""
    out = dict(
        catalogs=np.asarray(ns["AM_catalogs_"], dtype=np.float64),
        time_hipp=ns["time_iad_hipp"], cpsi_hipp=ns["CPSI_HIPP_"], spsi_hipp=ns["SPSI_HIPP_"],
        epoch_hipp=ns["EPOCH_HIPP_"], parf_hipp=ns["PARF_HIPP_"], res_hipp=ns["RES_HIPP_"],
        sres_hipp=ns["SRES_HIPP_"], time_gost=ns["time_iad_gost"], cpsi_gost=ns["CPSI_GOST_"],
        spsi_gost=ns["SPSI_GOST_"], parf_gost=ns["PARF_GOST_"],
        mask_gdr2=np.asarray(ns["mask_GDR2"], dtype=bool), mask_gdr3=np.asarray(ns["mask_GDR3"], dtype=bool),
        gsv2=ns["AM_GSV"]["GDR2"], gsv3=ns["AM_GSV"]["GDR3"], inv_cov=ns["AM_inv_COV"],
        log_det_cov=ns["AM_log_det_COV"], astro_gost=ns["AM_astro_gost"].values,
        common_t=np.float64(ns["common_t"]))
    return {k: np.ascontiguousarray(v, dtype=(bool if k.startswith("mask") else np.float64)) for k, v in out.items()}


def validate(am: Dict[str, np.ndarray]) -> Dict[str, np.ndarray]:
    missing = [k for k in AM_KEYS if k not in am]
    if missing:
        raise ValueError(f"astrometry data lacks {missing}")
    a = {k: np.ascontiguousarray(am[k], dtype=(bool if k.startswith("mask") else np.float64)) for k in AM_KEYS}
    nh, ng = len(a["time_hipp"]), len(a["time_gost"])
    for k in ("cpsi_hipp", "spsi_hipp", "epoch_hipp", "parf_hipp", "res_hipp", "sres_hipp"):
        if a[k].shape != (nh,):
            raise ValueError(f"{k} must have shape ({nh},)")
    for k in ("cpsi_gost", "spsi_gost", "parf_gost", "mask_gdr2", "mask_gdr3"):
        if a[k].shape != (ng,):
            raise ValueError(f"{k} must have shape ({ng},)")
    if a["catalogs"].shape != (3, 7) or a["inv_cov"].shape != (3, 5, 5) or a["astro_gost"].shape != (2, 5):
        raise ValueError("catalogs[3,7], inv_cov[3,5,5], astro_gost[2,5] expected")
    if a["gsv2"].shape != (5, int(a["mask_gdr2"].sum())) or a["gsv3"].shape != (5, int(a["mask_gdr3"].sum())):
        raise ValueError("gsv2 / gsv3 must be [5, sum(mask)]")
    return a


def am_to_c(am: Dict[str, np.ndarray]):
    """-> (EmpAmDataC, keepalive list) for emp_create (include/emperor_b200.h EmpAmData)."""
    from ._lib import EmpAmDataC
    a = validate(am)
    keep = []

    def ptr(x, dtype=np.float64):
        x = np.ascontiguousarray(x, dtype=dtype)
        keep.append(x)
        return x.ctypes.data

    idx2 = np.flatnonzero(a["mask_gdr2"]).astype(np.int32)
    idx3 = np.flatnonzero(a["mask_gdr3"]).astype(np.int32)
    c = EmpAmDataC()
    c.n_hipp, c.n_gost = len(a["time_hipp"]), len(a["time_gost"])
    c.n_mask2, c.n_mask3 = len(idx2), len(idx3)
    c.common_t = float(np.asarray(a["common_t"]).reshape(-1)[0])
    for f in ("time_hipp", "cpsi_hipp", "spsi_hipp", "epoch_hipp", "parf_hipp", "res_hipp", "sres_hipp",
              "time_gost", "cpsi_gost", "spsi_gost", "parf_gost", "gsv2", "gsv3", "inv_cov", "log_det_cov",
              "astro_gost", "catalogs"):
        setattr(c, f, ptr(a[f]))
    c.idx_mask2 = ptr(idx2, np.int32)
    c.idx_mask3 = ptr(idx3, np.int32)
    keep.append(c)
    return c, keep


# ------------------------------------------------------------------ file loading ----
HIPP_EPOCH_JD = 2448348.75  # astropy Time(1991.25, format='decimalyear').jd (qol_utils.py:210-211)
GDR2_REF_EP, GDR2_BASELINE = 2457206.0, 2457532.0  # qol_utils.py:215-216
GDR3_REF_EP, GDR3_BASELINE = 2457388.5, 2457902.0  # qol_utils.py:218-219


def load_am_folder(path: str, common_t: float, deadtime_dir: Optional[str] = None) -> Dict[str, np.ndarray]:
    """Mirror of DataWrapper.mk_AM for `datafiles/<star>/AM/` (qol_utils.py:104-446):
    `*_hipgaia.hg123`, `*_hip2.abs`, `*_gost.csv`.  `deadtime_dir` holds the two Gaia dead-time
    tables the reference ships under support/deadtime/ (public Gaia data, not vendored here);
    without it only the DR2/DR3 baseline cuts define the masks."""
    import pandas as pd
    hg = hipp = gost = None
    for fn in sorted(os.listdir(path)):
        ident = fn.split("_")[-1]
        ff = os.path.join(path, fn)
        if ident == "hipgaia.hg123":
            hg = pd.read_csv(ff, sep=r"\s+")
        elif ident == "hip2.abs":
            hipp = pd.read_csv(ff, sep=r"\s+")
        elif ident == "gost.csv":
            gost = pd.read_csv(ff)
            gost.columns = gost.columns.astype(str).str.strip()
    if hg is None or hipp is None or gost is None:
        raise FileNotFoundError(f"{path}: need *_hipgaia.hg123, *_hip2.abs and *_gost.csv")
    cols = ["ref_epoch", "ra", "dec", "parallax", "pmra", "pmdec", "radial_velocity"]
    catalogs = hg[cols].values.astype(np.float64)
    # astrometry_astro (qol_utils.py:234-245)
    df = hg[["ra", "dec", "parallax", "pmra", "pmdec"]][-2:].copy()
    dra = (df["ra"] - df["ra"].iloc[-1]) * np.cos(df["dec"].iloc[-2] * np.pi / 180) * 3.6e6
    ddec = (df["dec"] - df["dec"].iloc[-1]) * 3.6e6
    astro_gost = np.column_stack([dra.values, ddec.values, df["parallax"].values, df["pmra"].values,
                                  df["pmdec"].values])
    # gost: rename, filter, masks (qol_utils.py:279-345)
    gost = gost.rename(columns={"ObservationTimeAtBarycentre[BarycentricJulianDateInTCB]": "BJD",
                                "scanAngle[rad]": "psi", "parallaxFactorAlongScan": "parf",
                                "parallaxFactorAcrossScan": "parx"})
    gost = gost[gost["BJD"] < GDR3_BASELINE][["BJD", "psi", "parf", "parx"]]
    t = gost["BJD"].values
    valid2 = np.ones(len(t), bool)
    valid3 = np.ones(len(t), bool)
    if deadtime_dir is not None:
        off, scale = 1717.6256, 365.25 / 1461
        for name, valid in (("astrometric_gaps_gaiadr2_08252020.csv", valid2),
                            ("astrometric_gaps_gaiaedr3_12232020.csv", valid3)):
            dead = pd.read_csv(os.path.join(deadtime_dir, name), comment="#")
            st = 2457023.75 + (dead["start"].values - off) * scale
            en = 2457023.75 + (dead["end"].values - off) * scale
            for a, b in zip(st, en):
                valid[np.logical_and(t >= a, t <= b)] = 0
    mask2 = (t < GDR2_BASELINE) & valid2
    mask3 = (t < GDR3_BASELINE) & valid3
    cpsi, spsi = np.cos(gost["psi"].values), np.sin(gost["psi"].values)
    gsv = {}
    for cat, mask, ref_ep in (("gsv2", mask2, GDR2_REF_EP), ("gsv3", mask3, GDR3_REF_EP)):
        tf = (t[mask] - ref_ep) / 365.25
        XX = np.column_stack([spsi[mask], cpsi[mask], gost["parf"].values[mask], tf * spsi[mask], tf * cpsi[mask]])
        gsv[cat] = np.linalg.pinv(XX)
    keys = ["ra", "dec", "parallax", "pmra", "pmdec"]
    inv_cov, logdet = [], []
    for _, row in hg.iterrows():
        cov = np.zeros((5, 5))
        for i, ki in enumerate(keys):
            cov[i, i] = row[f"{ki}_error"] ** 2
            for j in range(i + 1, 5):
                cov[i, j] = cov[j, i] = row[f"{ki}_{keys[j]}_cov"]
        sign, ld = np.linalg.slogdet(cov)
        if sign <= 0:
            raise ValueError("catalogue covariance matrix is not positive definite")
        inv_cov.append(np.linalg.inv(cov))
        logdet.append(ld)
    return validate(dict(
        catalogs=catalogs, time_hipp=hipp["BJD"].values, cpsi_hipp=hipp["CPSI"].values,
        spsi_hipp=hipp["SPSI"].values, epoch_hipp=hipp["EPOCH"].values, parf_hipp=hipp["PARF"].values,
        res_hipp=hipp["RES"].values, sres_hipp=hipp["SRES"].values, time_gost=t, cpsi_gost=cpsi, spsi_gost=spsi,
        parf_gost=gost["parf"].values, mask_gdr2=mask2, mask_gdr3=mask3, gsv2=gsv["gsv2"], gsv3=gsv["gsv3"],
        inv_cov=np.array(inv_cov), log_det_cov=np.array(logdet), astro_gost=astro_gost,
        common_t=np.float64(common_t)))

"""RV data loading with the reference's preprocessing.

Mirrors `DataWrapper.mk_RV` / `get_data__` (qol_utils.py:61-100, 192-198):
one whitespace table per instrument (BJD, RV, eRV[, activity indices...]),
files taken in sorted name order, per-file RV mean subtraction when
|mean| > 1e-6, a 1-based `Flag` column per file, concatenation, sort by BJD and
a shift by `common_t = min(BJD)`.  Activity-index columns (`switch_SA`,
emp.py:2305-2312) are mean-subtracted and rescaled to the RV range of their file
(qol_utils.py:88-93) and are zero outside their instrument (`fillna(0)`,
qol_utils.py:198).  Host-side, runs once per data set.
"""
import os
from dataclasses import dataclass
from typing import List, Sequence, Tuple

import numpy as np


@dataclass
class RVData:
    t: np.ndarray  # BJD - common_t, sorted
    y: np.ndarray  # RV, per-instrument mean removed
    yerr: np.ndarray
    flag: np.ndarray  # int32, 1..nins
    common_t: float
    labels: List[str]
    sai: np.ndarray = None        # [n, sum(cornums)] activity columns (SAI{j}_ of the generated script) or None
    cornums: List[int] = None     # activity columns per instrument (data_wrapper['RV']['nsai'])

    @property
    def nins(self) -> int:
        return len(self.labels)

    def __len__(self) -> int:
        return len(self.t)


def from_instrument_tables(tables: Sequence[Tuple], labels: Sequence[str] = None) -> RVData:
    """tables: one (t, rv, erv) or (t, rv, erv, activity[n_i, c_i]) per instrument."""
    ts, ys, es, fs, acts = [], [], [], [], []
    for i, tab in enumerate(tables):
        t = np.asarray(tab[0], dtype=np.float64)
        rv = np.asarray(tab[1], dtype=np.float64).copy()
        erv = np.asarray(tab[2], dtype=np.float64)
        m = rv.mean()  # pandas df.mean()['RV'] (qol_utils.py:84-85)
        if abs(m) > 1e-6:
            rv -= m
        act = np.zeros((len(t), 0))
        if len(tab) > 3 and tab[3] is not None and np.size(tab[3]):
            act = np.array(tab[3], dtype=np.float64).reshape(len(t), -1)
            for j in range(act.shape[1]):  # qol_utils.py:88-93, statement by statement
                s = act[:, j] - act[:, j].mean()
                act[:, j] = (s - s.min()) / (s.max() - s.min()) * (rv.max() - rv.min()) + rv.min()
        ts.append(t), ys.append(rv), es.append(erv), acts.append(act)
        fs.append(np.full(len(t), i + 1, dtype=np.int32))
    t = np.concatenate(ts)
    order = np.argsort(t, kind="stable")  # pd.sort_values('BJD'); ties keep file order
    t = t[order]
    common_t = float(t.min())
    cornums = [a.shape[1] for a in acts]
    sai = None
    if sum(cornums):
        sai = np.zeros((len(t), sum(cornums)))  # pd.concat(...).fillna(0): zero outside the column's instrument
        row = col = 0
        for a in acts:
            sai[row:row + len(a), col:col + a.shape[1]] = a
            row, col = row + len(a), col + a.shape[1]
        sai = np.ascontiguousarray(sai[order])
    return RVData(t=t - common_t, y=np.concatenate(ys)[order], yerr=np.concatenate(es)[order],
                  flag=np.concatenate(fs)[order], common_t=common_t,
                  labels=list(labels) if labels is not None else [f"ins{i + 1}" for i in range(len(ts))],
                  sai=sai, cornums=cornums)


def load_rv_folder(path: str, switch_SA: bool = False) -> RVData:
    """`path` = .../datafiles/<star>/RV/ (qol_utils.py:19-21)."""
    names = [fn for fn in sorted(os.listdir(path)) if fn != ".DS_Store"]
    if not names:
        raise FileNotFoundError(f"no RV files under {path}")
    tables = []
    for fn in names:
        d = np.loadtxt(os.path.join(path, fn))
        if d.ndim != 2 or d.shape[1] < 3:
            raise ValueError(f"{fn}: expected columns BJD RV eRV")
        # columns after eRV are activity indices; without switch_SA the reference drops them (emp.py:2310-2312)
        tables.append((d[:, 0], d[:, 1], d[:, 2], d[:, 3:] if switch_SA else None))
    return from_instrument_tables(tables, names)

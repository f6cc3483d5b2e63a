"""RV data loading with the reference's preprocessing.

Mirrors `DataWrapper.mk_RV` / `get_data__` (qol_utils.py:61-100, 192-198):
one whitespace table per instrument (BJD, RV, eRV[, activity indices...]),
files taken in sorted name order, per-file RV mean subtraction when
|mean| > 1e-6, a 1-based `Flag` column per file, concatenation, sort by BJD and
a shift by `common_t = min(BJD)`.  Host-side, runs once per data set.
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

    @property
    def nins(self) -> int:
        return len(self.labels)

    def __len__(self) -> int:
        return len(self.t)


def from_instrument_tables(tables: Sequence[Tuple[np.ndarray, np.ndarray, np.ndarray]],
                           labels: Sequence[str] = None) -> RVData:
    ts, ys, es, fs = [], [], [], []
    for i, (t, rv, erv) in enumerate(tables):
        t = np.asarray(t, dtype=np.float64)
        rv = np.asarray(rv, dtype=np.float64).copy()
        erv = np.asarray(erv, dtype=np.float64)
        m = rv.mean()  # pandas df.mean()['RV'] (qol_utils.py:84-85)
        if abs(m) > 1e-6:
            rv -= m
        ts.append(t), ys.append(rv), es.append(erv)
        fs.append(np.full(len(t), i + 1, dtype=np.int32))
    t = np.concatenate(ts)
    order = np.argsort(t, kind="stable")  # pd.sort_values('BJD'); ties keep file order
    t = t[order]
    common_t = float(t.min())
    return RVData(t=t - common_t, y=np.concatenate(ys)[order], yerr=np.concatenate(es)[order],
                  flag=np.concatenate(fs)[order], common_t=common_t,
                  labels=list(labels) if labels is not None else [f"ins{i + 1}" for i in range(len(ts))])


def load_rv_folder(path: str) -> RVData:
    """`path` = .../datafiles/<star>/RV/ (qol_utils.py:19-21)."""
    names = [fn for fn in sorted(os.listdir(path)) if fn != ".DS_Store"]
    if not names:
        raise FileNotFoundError(f"no RV files under {path}")
    tables = []
    for fn in names:
        d = np.loadtxt(os.path.join(path, fn))
        if d.ndim != 2 or d.shape[1] < 3:
            raise ValueError(f"{fn}: expected columns BJD RV eRV")
        tables.append((d[:, 0], d[:, 1], d[:, 2]))
    return from_instrument_tables(tables, names)

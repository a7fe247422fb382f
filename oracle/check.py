"""TEST INFRASTRUCTURE ONLY - the checker used by __graft_entry__.smoke(): replays one
Newton-iteration assembly through the oracle (the plain-C restatement, or the compiled
reference when only that is present) and returns the largest norm-wise relative difference
from the arrays the CUDA path produced."""
from __future__ import annotations

import numpy as np


def _backend():
    from . import oraclebind, refbind
    if oraclebind.available():
        return oraclebind
    if refbind.available():
        return refbind
    raise RuntimeError("neither oracle/libcubens_oracle.so nor oracle/_ref is built "
                       "(run __graft_entry__.build())")


def _rel(a, b):
    s = float(np.abs(b).max()) if b.size else 0.0
    d = float(np.abs(a - b).max()) if b.size else 0.0
    return 0.0 if d == 0.0 else d / s


def compare_iteration(m, dd, ss_after, f_after):
    """sequence: begin_increment; stiff; update_forces(dd); end_iteration; stiff.
    ss_after = skyline K_t of the second stiff, f_after = f_temp of update_forces."""
    B = _backend()
    from .refbind import RefState
    s = RefState(m)
    s.begin_increment()
    B.stiff(m, s, SLVFLAG=0)
    B.update_forces(m, s, dd)
    s.end_iteration()
    ss = B.stiff(m, s, SLVFLAG=0)
    return max(_rel(ss_after, ss), _rel(f_after, s.f_temp))


def compare_iteration_csc(m, dd, Ap, Ai, Ax_after, f_after):
    """the same sequence for the device CSC layout: entry (i <= j) of the oracle's skyline lives at
    ss[maxa[j-1] + (j-i) - 1] (1-based, model.c:1269-1278); both triangles of the CSC are held against it,
    structural CSC entries outside the skyline profile must be exactly zero"""
    B = _backend()
    from .refbind import RefState
    s = RefState(m)
    s.begin_increment()
    B.stiff(m, s, SLVFLAG=0)
    B.update_forces(m, s, dd)
    s.end_iteration()
    ss = B.stiff(m, s, SLVFLAG=0)
    n = m.NEQ
    maxa = np.asarray(m.maxa, dtype=np.int64)
    cols = np.repeat(np.arange(n, dtype=np.int64), np.diff(Ap))
    rows = np.asarray(Ai, dtype=np.int64)
    lo, hi = np.minimum(rows, cols), np.maximum(rows, cols)
    addr = maxa[hi] - 1 + (hi - lo)
    inside = addr < maxa[hi + 1] - 1
    if np.any(Ax_after[~inside] != 0.0):
        return float("inf")
    return max(float(np.abs(Ax_after[inside] - ss[addr[inside]]).max() / np.abs(ss).max()), _rel(f_after, s.f_temp))

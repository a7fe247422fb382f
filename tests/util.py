"""Shared helpers for the parity tests."""
import numpy as np


def relerr(a, b):
    """norm-wise relative difference max|a-b| / max|b| (0 when both are all-zero)."""
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    s = np.abs(b).max() if b.size else 0.0
    d = np.abs(a - b).max() if b.size else 0.0
    return 0.0 if d == 0.0 else d / s


def csc_to_dense(n, Ap, Ai, Ax):
    K = np.zeros((n, n))
    for c in range(n):
        for p in range(Ap[c], Ap[c + 1]):
            K[Ai[p], c] += Ax[p]
    return K


def skyline_to_dense(n, maxa, ss):
    """upper triangle from the reference's skyline vector (model.c:1269-1278)."""
    K = np.zeros((n, n))
    for j in range(1, n + 1):
        h = maxa[j] - maxa[j - 1]
        for k in range(h):
            i = j - k
            K[i - 1, j - 1] = ss[maxa[j - 1] - 1 + k]
    return K

"""Sparse acoustic FSI (ANAFLAG 4, VERDICT r01 item 9 / SURVEY §8 f4).

The reference builds the coupled system as DENSE NEQ x NEQ arrays (fsi.c:333-445): stiff_fsi = [K L; 0 H],
mass_fsi = [M 0; -rho L^T Q] with L = G A from L_br (fsi.c:447-531).  The library keeps both on one sparse
joint-block CSC pattern (pressure DOFs on twin joints, the coupling as a two-joint element).  Parity: every
entry of the reference's dense matrices, block by block, each block against its own magnitude at north_star's
1e-12 (K is ~1e11, H ~1, L ~1, Q ~1e-7: a single norm over the whole matrix would hide the small blocks); entries
outside the sparse pattern must be exactly zero in the reference.
"""
import numpy as np
import pytest

import cubens_b200 as cb
from cubens_b200 import meshgen

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _dense(n, Ap, Ai, Ax):
    D = np.zeros((n, n))
    cols = np.repeat(np.arange(n), np.diff(Ap))
    D[Ai, cols] = Ax
    inpat = np.zeros((n, n), dtype=bool)
    inpat[Ai, cols] = True
    return D, inpat


def _blocks(sn):
    s, f = slice(0, sn), slice(sn, None)
    return {"ss": (s, s), "sf": (s, f), "fs": (f, s), "ff": (f, f)}


def _compare(name, got, want, inpat, sn):
    assert np.all(want[~inpat] == 0.0), f"{name}: the reference has entries outside the sparse pattern"
    for b, (r, c) in _blocks(sn).items():
        w, g = want[r, c], got[r, c]
        scale = np.abs(w).max()
        if scale == 0.0:
            assert np.abs(g).max() == 0.0, f"{name}[{b}] must be exactly zero"
        else:
            assert np.abs(g - w).max() / scale < TOL, f"{name}[{b}]: {np.abs(g - w).max() / scale:.2e}"


@pytest.mark.parametrize("skin", [False, True], ids=["brick_solid", "shell_skin"])
@pytest.mark.parametrize("dims", [(3, 2, 2, 2), (7, 6, 3, 4)])
def test_fsi_system_matrices(gpu, ref, skin, dims):
    m = meshgen.fsi_model(*dims, skin=skin, distort=0.15)
    K_ref, M_ref = ref.fsi_matrices(m)
    asm = cb.Assembler(m, layout=cb.CB_MAT_CSC)
    asm.begin_increment(); asm.stiff()
    Ap, Ai, Ax = asm.csc()
    K, inpat = _dense(m.NEQ, Ap, Ai, Ax)
    _compare("stiff_fsi", K, K_ref, inpat, m.SNDOF)
    asm.mass()
    M, _ = _dense(m.NEQ, Ap, Ai, asm.mass_csc())
    _compare("mass_fsi", M, M_ref, inpat, m.SNDOF)
    # the sparse system is what makes configs[4]'s size reachable: nnz is O(NEQ), not NEQ^2
    assert len(Ax) < 90 * m.NEQ
    with pytest.raises(cb.CubensError):
        asm.update_forces(np.zeros(m.NEQ))         # fsi.c: linear, assembled once - no force pass
    asm.close()


def test_fsi_needs_interface_data(gpu):
    m = meshgen.fsi_model(2, 2, 1, 1)
    m.nnorm = None
    with pytest.raises(cb.CubensError):
        cb.Assembler(m, layout=cb.CB_MAT_CSC)

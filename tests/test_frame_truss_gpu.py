"""GPU parity of the truss and 14-DOF frame paths (and mixed meshes) against the compiled
reference: K_t (skyline + CSC), co-rotational update, f_int, fixed-end forces, lumped mass."""
import os
import sys

import numpy as np
import pytest

import cubens_b200 as cb
from cubens_b200 import meshgen, model as M
from util import relerr, walk, csc_to_dense, TOL

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import make_golden as G  # noqa: E402

pytestmark = pytest.mark.gpu


def test_truss_walk(gpu, ref):
    m = meshgen.truss_model(3)
    asm = cb.Assembler(m, layout=cb.CB_MAT_BOTH)
    walk(m, ref, asm, scale=1e-2)
    asm.close()


@pytest.mark.parametrize("name", ["lattice_3", "lattice_3_offsets_releases"])
def test_frame_walk(gpu, ref, name):
    m = G.build(name)
    asm = cb.Assembler(m, layout=cb.CB_MAT_BOTH)
    s = walk(m, ref, asm, scale=1e-2)
    # CSC of the next iteration against the reference's dense scatter
    s.begin_increment(); asm.begin_increment(); asm.stiff()
    dense = ref.stiff(m, s, SLVFLAG=2).reshape(m.NEQ, m.NEQ)
    K = csc_to_dense(m.NEQ, *asm.csc())
    assert relerr(K, dense.T) < TOL
    asm.close()


@pytest.mark.parametrize("maker", [lambda: meshgen.truss_model(3, ANAFLAG=1),
                                   lambda: meshgen.lattice_model(3, ANAFLAG=1)])
def test_linear_and_mass(gpu, ref, maker):
    m = maker()
    asm = cb.Assembler(m, layout=cb.CB_MAT_BOTH)
    s = ref.RefState(m)
    assert relerr(asm.mass(), ref.mass(m, s, SLVFLAG=0)) < TOL
    assert relerr(asm.download("LLENGTH"), s.llength) < 1e-15
    ss_ref = ref.stiff(m, s, SLVFLAG=0, gen="c")
    asm.stiff(cb.CB_GEN_COMMITTED)
    assert relerr(asm.skyline(), ss_ref) < TOL
    d = np.random.default_rng(3).uniform(-1e-2, 1e-2, size=m.NEQ)
    f_ref = ref.forces_linear(m, s, d)
    assert relerr(asm.forces_linear(d), f_ref) < TOL
    assert relerr(asm.download("EF"), s.ef) < TOL
    asm.close()


def test_mixed_shell_frame_truss(gpu, ref):
    """a plate stiffened by frames along one edge and braced by trusses: mixed joint DOF counts"""
    p = meshgen.plate_model(4, 3, z_bump=0.02, pinned=False)
    X = p.x.reshape(-1, 3)
    shells = p.minc.reshape(-1, 3)
    edge = [i * 4 + 1 for i in range(5)]                       # joints along j = 0
    frames = np.array([[edge[i], edge[i + 1]] for i in range(4)])
    aux = X[frames[:, 0] - 1] + np.array([0.0, 0.0, 1.0])
    trusses = np.array([[1, 7], [6, 12], [11, 17]])
    fixed = [(1, d) for d in range(1, 8)] + [(20, 1), (20, 2), (20, 3), (16, 3)]
    fp = (2.1e11, 8.0e10, 8050.0, 1e-4, 1e-8, 2e-8, 3e-8, 1e-12, 3.45e8, 1e-5, 1e-5)
    m = M.build_model(p.x, trusses=trusses, frames=frames, shells=shells, fixed=fixed,
                      truss_props=(2.1e11, 1e-4, 8050.0, 3.45e8), frame_props=fp,
                      shell_props=meshgen.SHELL_5C, frame_aux=aux, ANAFLAG=2)
    asm = cb.Assembler(m, layout=cb.CB_MAT_BOTH)
    s = walk(m, ref, asm, scale=1e-5)
    s.begin_increment(); asm.begin_increment(); asm.stiff()
    dense = ref.stiff(m, s, SLVFLAG=2).reshape(m.NEQ, m.NEQ)
    assert relerr(csc_to_dense(m.NEQ, *asm.csc()), dense.T) < TOL
    asm.close()


@pytest.mark.parametrize("name", ["brick_2x2x2", "brick_skin_2x2x1"])
def test_brick_stiffness(gpu, ref, name):
    """stiff_br (+ stiff_sh on the skinned face) into the CSC vs the reference's dense scatter"""
    m = G.build(name)
    asm = cb.Assembler(m, layout=cb.CB_MAT_CSC)
    s = ref.RefState(m)
    dense = ref.stiff(m, s, SLVFLAG=2, gen="c").reshape(m.NEQ, m.NEQ)
    asm.stiff(cb.CB_GEN_COMMITTED)
    K = csc_to_dense(m.NEQ, *asm.csc())
    assert relerr(K, dense.T) < TOL
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".npz"))
    assert relerr(K, gold["K_dense"].reshape(m.NEQ, m.NEQ).T) < TOL
    asm.close()


@pytest.mark.parametrize("name", ["brick_2x2x2", "brick_skin_2x2x1"])
def test_brick_mass(gpu, ref, name):
    """mass_br (consistent, brick.c:399-537) + lumped mass_sh of the skin in the reference's
    full-order layout vs the CSC mass matrix the device assembles on the pattern of K_t"""
    m = G.build(name)
    asm = cb.Assembler(m, layout=cb.CB_MAT_CSC)
    s = ref.RefState(m)
    dense = ref.mass(m, s, SLVFLAG=2).reshape(m.NEQ, m.NEQ)
    Mx = asm.mass_csc()
    Ap, Ai, _ = asm.csc()
    M = csc_to_dense(m.NEQ, Ap, Ai, Mx)
    assert relerr(M, dense.T) < TOL
    assert np.abs(M - M.T).max() <= 1e-14 * np.abs(M).max() and np.abs(M).max() > 0
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".npz"))
    assert relerr(M, gold["M_dense"].reshape(m.NEQ, m.NEQ).T) < TOL
    asm.close()

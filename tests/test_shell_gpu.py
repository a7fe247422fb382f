"""GPU parity of the DKT shell path against the compiled reference (oracle/_ref): K_t in both
matrix layouts, the co-rotational update and f_int over several Newton-like iterations.

Tolerances (BASELINE.json north_star): element/assembled matrices and residuals 1e-12
norm-wise relative; the integer maps and the thresholded CSC pattern exact."""
import numpy as np
import pytest

import cubens_b200 as cb
from cubens_b200 import meshgen
from util import relerr, csc_to_dense

pytestmark = pytest.mark.gpu
TOL = 1e-12


from util import walk as _walk


@pytest.mark.parametrize("nx,ny,bump", [(1, 1, 0.0), (2, 2, 0.0), (6, 4, 0.02), (9, 7, 0.05)])
def test_shell_newton_walk(gpu, ref, nx, ny, bump):
    m = meshgen.plate_model(nx, ny, z_bump=bump, pinned=(nx > 1))
    if nx == 1:   # single cell: hold one corner so NEQ stays small but non-trivial
        m = meshgen.plate_model(1, 1, pinned=False)
    asm = cb.Assembler(m, layout=cb.CB_MAT_BOTH)
    _walk(m, ref, asm)
    asm.close()


def test_shell_csc_matches_dense_and_reference_pattern(gpu, ref):
    m = meshgen.plate_model(5, 4, z_bump=0.03)
    asm = cb.Assembler(m, layout=cb.CB_MAT_BOTH)
    s = _walk(m, ref, asm, n_iter=2)
    s.begin_increment(); asm.begin_increment()
    asm.stiff()
    dense = ref.stiff(m, s, SLVFLAG=2).reshape(m.NEQ, m.NEQ)
    Ap, Ai, Ax = asm.csc()
    # structural pattern: sorted rows, every column non-empty, diagonal present
    for c in range(m.NEQ):
        rows = Ai[Ap[c]:Ap[c + 1]]
        assert rows.size and np.all(np.diff(rows) > 0) and c in rows
    K = csc_to_dense(m.NEQ, Ap, Ai, Ax)
    # the reference's dense scatter stores K[je][ie] at ss[(i-1)*NEQ+j-1] (shell.c:338): row-major
    # dense == column-major CSC of K
    assert relerr(K, dense.T) < TOL
    # thresholded pattern exactly as solve.c:110-119 builds it
    rAp, rAi, rAx, _ = ref.dense_to_csc(m, dense.reshape(-1).copy())
    cAp, cAi, cAx = asm.csc_compact(1e-10)
    assert np.array_equal(cAp, rAp) and np.array_equal(cAi, rAi)
    assert relerr(cAx, rAx) < TOL
    asm.close()


def test_shell_linear_forces_and_mass(gpu, ref):
    m = meshgen.plate_model(4, 3, ANAFLAG=1)
    asm = cb.Assembler(m, layout=cb.CB_MAT_BOTH)
    s = ref.RefState(m)
    ss_ref = ref.stiff(m, s, SLVFLAG=0, gen="c")
    asm.stiff(cb.CB_GEN_COMMITTED)
    assert relerr(asm.skyline(), ss_ref) < TOL
    d = np.random.default_rng(3).uniform(-1e-3, 1e-3, size=m.NEQ)
    f_ref = ref.forces_linear(m, s, d)
    f = asm.forces_linear(d)
    assert relerr(f, f_ref) < TOL
    assert relerr(asm.download("EF"), s.ef) < TOL
    sm_ref = ref.mass(m, s, SLVFLAG=0)
    sm = asm.mass()
    assert relerr(sm, sm_ref) < TOL
    asm.close()


def test_repeatable(gpu):
    """assembly is atomic-free: two runs give bit-identical matrices"""
    m = meshgen.plate_model(12, 9, z_bump=0.01)
    out = []
    for _ in range(2):
        asm = cb.Assembler(m, layout=cb.CB_MAT_CSC)
        asm.begin_increment(); asm.stiff()
        f, *_ = asm.update_forces(meshgen.perturbation(m)); asm.end_iteration(); asm.stiff()
        out.append((asm.csc()[2].copy(), f))
        asm.close()
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])


def test_convergence_test_matches_reference_test(gpu, ref):
    """test() of the reference (misc.c:187-250) evaluated from the device-resident vectors: the five sums, the
    convergence code, and the reaction resultants (forces the reference drops at fixed DOFs)"""
    import ctypes as C
    m = meshgen.plate_model(20, 15, z_bump=0.02)
    asm = cb.Assembler(m, layout=cb.CB_MAT_CSC)
    s = ref.RefState(m)
    l = ref.set_model(m)
    asm.set_q(m.q)
    rng = np.random.default_rng(11)
    lpf = 0.7
    qtot = lpf * m.q
    s.begin_increment(); asm.begin_increment()
    intener1 = None
    for it in range(3):
        dd = rng.uniform(-1e-4, 1e-4, size=m.NEQ) * (0.1 ** it)
        f_ip = s.f_temp.copy(); fp = s.f.copy()
        ref.update_forces(m, s, dd, itecnt=it)
        f, *_ = asm.update_forces(dd, itecnt=it)
        want = np.array([(qtot - s.f_temp) @ (qtot - s.f_temp), dd @ dd, dd @ (qtot - f_ip), s.d_temp @ s.d_temp,
                         (qtot - fp) @ (qtot - fp)])
        if it == 0:
            intener1 = float(dd @ (qtot - fp))
        for tols in ((1e-3, 1e-3, 1e-3), (1e-1, 1e-9, 0.5), (2.0, 1e-12, 1e-12)):
            ref.set_model(m)
            conv = C.c_int(0); ie = C.c_double(intener1)
            err = l.test(ref.P(s.d_temp), ref.P(dd), ref.P(s.f_temp), ref.P(fp), ref.P(qtot), ref.P(f_ip), C.byref(ie),
                         C.byref(conv), C.byref(C.c_double(tols[0])), C.byref(C.c_double(tols[1])),
                         C.byref(C.c_double(tols[2])))
            gerr, gconv, sums = asm.convergence_test(lpf, intener1, *tols)
            assert (gerr, gconv) == (err, conv.value)
            assert np.allclose(sums, want, rtol=1e-12, atol=0)
        # deterministic
        assert np.array_equal(asm.residual_sums(lpf, fetch=True), asm.residual_sums(lpf, fetch=True))
        # reactions: every element's joint forces are self-equilibrated, so the translational resultants at
        # the fixed DOFs balance the internal forces at the free ones
        R = asm.reaction_sums()
        jc = np.asarray(m.jcode).reshape(-1, 7)
        for k in range(3):
            free = jc[:, k][jc[:, k] > 0] - 1
            assert abs(R[k] + f[free].sum()) <= 1e-9 * np.abs(f).max()
        s.end_iteration(); asm.end_iteration()
    asm.close()


def test_upload_geometry_rebuilds_derived_data(gpu):
    """restart path: uploading FAREA / SLENGTH must rebuild the cached DKT matrices and
    drop the geometry classes, exactly as cb_mass does after rewriting them"""
    m = meshgen.plate_model(8, 6, z_bump=0.03)
    dd = meshgen.perturbation(m, scale=1e-3)
    a = cb.Assembler(m, layout=cb.CB_MAT_BOTH)
    a.begin_increment(); a.update_forces(dd); a.end_iteration(); a.commit()
    a.mass()                                   # slength / farea <- committed coordinates
    a.begin_increment(); a.stiff()
    K1, A1 = a.skyline(), a.csc_values()
    fa, sl = a.download("FAREA"), a.download("SLENGTH")
    assert not np.array_equal(fa, m.farea)
    a.close()
    b = cb.Assembler(m, layout=cb.CB_MAT_BOTH)
    b.begin_increment(); b.stiff()             # builds the plan and the initial-geometry tables first
    b.update_forces(dd); b.end_iteration(); b.commit()
    b.upload("FAREA", fa); b.upload("SLENGTH", sl)
    b.begin_increment(); b.stiff()
    assert np.array_equal(b.skyline(), K1) and np.array_equal(b.csc_values(), A1)
    b.close()


@pytest.mark.parametrize("full_every", ["0", "1", "3", "5"])
def test_symmetric_handoff_rebuilds_full_matrix(gpu, full_every, monkeypatch):
    """cb_csc_values_begin / _end: the packed upper triangle crosses PCIe, host threads rebuild the full
    columns.  The upper part is bit-identical to cb_get_csc_values, the lower part is its exact mirror
    (the device's own lower entries agree with it to rounding: K_t is symmetric); chunks sent in full
    (CB_SYM_FULL_EVERY) keep the device's own lower entries."""
    monkeypatch.setenv("CB_SYM_FULL_EVERY", full_every)
    m = meshgen.plate_model(40, 27, z_bump=0.02, SLVFLAG=2)
    a = cb.Assembler(m, layout=cb.CB_MAT_CSC)
    a.begin_increment(); a.update_forces(meshgen.perturbation(m, scale=1e-3)); a.end_iteration(); a.stiff()
    Ap, Ai, Ax = a.csc()
    full = a.csc_values_mirrored(nthreads=3)
    cols = np.repeat(np.arange(m.NEQ), np.diff(Ap))
    up = Ai <= cols
    assert np.array_equal(full[up], Ax[up])
    assert np.abs(full - Ax).max() <= 1e-13 * np.abs(Ax).max()
    # exact symmetry of the rebuilt matrix
    import scipy.sparse as sp
    K = sp.csc_matrix((full, Ai, Ap), shape=(m.NEQ, m.NEQ))
    if full_every == "0":
        assert abs(K - K.T).max() == 0.0
    elif full_every == "1":
        assert np.array_equal(full, Ax)
    assert abs(K - K.T).max() <= 1e-13 * np.abs(Ax).max()
    # the packed stream is an upper-triangular CSC of its own
    Apu, Aiu, Axu = a.csc_upper()
    U = sp.csc_matrix((Axu, Aiu, Apu), shape=(m.NEQ, m.NEQ))
    assert abs(sp.triu(K) - U).max() == 0.0 and U.nnz == sp.triu(K).nnz
    a.close()


def test_warp_partial_sums_equal_the_corner_gather(gpu, ref, monkeypatch):
    """cb_wsum.cuh: the warps of the force pass pre-sum their corners per joint before staging.  The same f_temp
    and reaction resultants as the corner-by-corner gather (CB_NO_WARP_SUMS=1) up to the association of the
    additions, both within 1e-12 of forces_sh; a union-jack plate puts eight corners of one joint into one warp
    (more than a slot holds: the joint gets a second slot)"""
    for kw in (dict(z_bump=0.02), dict(unionjack=True, z_bump=0.02), dict(jitter=0.2, z_bump=0.01)):
        m = meshgen.plate_model(37, 23, **kw)
        dd = meshgen.perturbation(m, scale=1e-3)
        s = ref.RefState(m); s.begin_increment()
        ref.update_forces(m, s, dd, itecnt=0)
        out = []
        for off in (False, True):
            if off:
                monkeypatch.setenv("CB_NO_WARP_SUMS", "1")
            else:
                monkeypatch.delenv("CB_NO_WARP_SUMS", raising=False)
            a = cb.Assembler(m, layout=cb.CB_MAT_CSC)
            a.begin_increment(); a.stiff()
            f, *_ = a.update_forces(dd, itecnt=0)
            a.set_q(m.q)
            sums = a.residual_sums(0.8, fetch=True); reac = a.reaction_sums()
            assert relerr(f, s.f_temp) < 1e-12
            out.append((f, sums, reac, a.map_bytes))
            a.close()
        (f0, s0, r0, mb0), (f1, s1, r1, mb1) = out
        assert mb0 > mb1, "the slot lists of the warp-level sums were not built"
        assert np.abs(f0 - f1).max() <= 1e-14 * np.abs(f1).max()
        assert np.allclose(s0, s1, rtol=1e-12, atol=0)
        assert np.allclose(r0, r1, rtol=1e-11, atol=1e-12 * np.abs(f1).max())
    monkeypatch.delenv("CB_NO_WARP_SUMS", raising=False)

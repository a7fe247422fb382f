"""Mid-size oracle parity of exactly the kernel instantiations bench.py times.

At <= 126 shells the persistent tile kernels never loop (fewer tiles than resident CTAs) and the
geometry-class tables stay off (CLS_MAX = NE_SH / 8), so the small walks of test_shell_gpu.py meet the
benchmarked code only through self-comparisons.  Here:

  * flat 200 x 60 plate, 24 000 shells, geometry classes ON, >> 592 tiles: the shell-only CSC stream
    kernel and k_shell_forces<1> against the reference's skyline (stiff_sh, shell.c:307-329 addressing)
    entry by entry, f_temp / ef_i / triads against forces_sh + updatc, three iterations;
  * the same plate jittered (every shell its own geometry, classes OFF): the <0> instantiations;
  * 20^3 frame lattice, 22 800 members: k_assemble_tiles<7,0,1> + k_frame_forces_simple against
    stiff_fr / forces_fr (frame.c:226, 902).

Tolerances: north_star's 1e-12 relative, applied BOTH norm-wise (max|a-b| / max|b|) and entry-wise
scaled by the diagonal, |dK_ij| <= 1e-12 sqrt(K_ii K_jj), so that the small drilling terms (1e-4 of the
bending stiffness) are held to their own magnitude.
"""
import numpy as np
import pytest

import cubens_b200 as cb
from cubens_b200 import meshgen
from util import relerr, csc_vs_skyline

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _walk_csc(m, ref, asm, n_iter=3, scale=1e-4, seed=5):
    s = ref.RefState(m)
    s.begin_increment(); asm.begin_increment()
    rng = np.random.default_rng(seed)
    Ap = Ai = None
    for it in range(n_iter):
        ss_ref = ref.stiff(m, s, SLVFLAG=0)
        asm.stiff()
        if Ap is None:
            Ap, Ai, Ax = asm.csc()
        else:
            Ax = asm.csc_values()
        nrm, ent = csc_vs_skyline(m, Ap, Ai, Ax, ss_ref)
        assert nrm < TOL, f"K_t norm-wise, iteration {it}: {nrm:.2e}"
        assert ent < TOL, f"K_t entry-wise (scaled by the diagonal), iteration {it}: {ent:.2e}"
        dd = rng.uniform(-scale, scale, size=m.NEQ)
        fr, sh, _ = ref.update_forces(m, s, dd, itecnt=it)
        f, gfr, gsh, _ = asm.update_forces(dd, itecnt=it)
        assert (fr, sh) == (gfr, gsh)
        assert relerr(f, s.f_temp) < TOL, f"f_temp iteration {it}"
        assert relerr(asm.download("EF_I"), s.ef_i) < TOL
        assert relerr(asm.download("X_TEMP"), s.x_temp) == 0.0
        for nm in ("C1", "C2", "C3"):
            assert relerr(asm.download(nm + "_I"), getattr(s, nm.lower() + "_i")) < 1e-15
        if m.NE_SH:
            assert relerr(asm.download("DEFFAREA_I"), s.deffarea_i) < 1e-15
            assert relerr(asm.download("DEFSLEN_I"), s.defslen_i) < 1e-15
        if m.NE_FR:
            assert relerr(asm.download("DEFLLEN_I"), s.defllen_i) < 1e-15
            assert relerr(asm.download("EFFE_I"), s.efFE_i) < TOL
        s.end_iteration(); asm.end_iteration()
    return s


@pytest.mark.parametrize("shape", ["narrow", "wide", "mini", "narrow12"])
def test_plate_200x60_classes_on(gpu, ref, shape, monkeypatch):
    monkeypatch.setenv("CB_KT", shape)            # both compiled shapes of the stream kernel
    m = meshgen.plate_model(200, 60, SLVFLAG=0)
    asm = cb.Assembler(m, layout=cb.CB_MAT_CSC)
    asm.begin_increment(); asm.stiff()
    assert asm.geometry_classes > 0, "flat structured plate must run the class-table kernels (<1>)"
    _walk_csc(m, ref, asm)
    assert asm.geometry_classes > 0
    asm.close()


@pytest.mark.parametrize("shape", ["narrow", "wide", "mini", "narrow12"])
def test_plate_200x60_jittered_classes_off(gpu, ref, shape, monkeypatch):
    monkeypatch.setenv("CB_KT", shape)
    m = meshgen.plate_model(200, 60, SLVFLAG=0, jitter=0.2, z_bump=0.01)
    asm = cb.Assembler(m, layout=cb.CB_MAT_CSC)
    asm.begin_increment(); asm.stiff()
    assert asm.geometry_classes == 0, "jittered plate: every shell streams / recomputes its own DKT data (<0>)"
    _walk_csc(m, ref, asm)
    asm.close()


@pytest.mark.parametrize("shape", ["narrow", "wide", "mini", "narrow12"])
def test_plate_unionjack_split_blocks(gpu, ref, shape, monkeypatch):
    """joints with 8 shells around them: diagonal blocks of 8 contributions are cut in two lane parts
    (first part stored, follower part added at the end of the tile)"""
    monkeypatch.setenv("CB_KT", shape)
    m = meshgen.plate_model(90, 40, SLVFLAG=0, unionjack=True, z_bump=0.02)
    asm = cb.Assembler(m, layout=cb.CB_MAT_CSC)
    _walk_csc(m, ref, asm, n_iter=2)
    asm.close()


def test_lattice_20_frame_tiles(gpu, ref):
    m = meshgen.lattice_model(20, SLVFLAG=0)
    asm = cb.Assembler(m, layout=cb.CB_MAT_CSC)
    _walk_csc(m, ref, asm, scale=1e-3, seed=9)
    asm.close()

"""Material-nonlinear analysis (ANAFLAG 3) of trusses and frames on the device against the unmodified
reference: plastic reduction of the tangent (stiffm_tr truss.c:206, stiffm_fr frame.c:581), the
yield check / return to the surface / elastic unloading inside forces_fr (frame.c:1157-1268,
regula_falsi 1397, unload 1457) with its early-return semantics, the squash-load
cap of forces_tr (truss.c:335-347), and the converged solution of the NR loop that reacts to the
return codes (main.c:2030-2063)."""
import numpy as np
import pytest

import cubens_b200 as cb
from cubens_b200 import meshgen
from util import relerr, ref_newton, TOL

pytestmark = pytest.mark.gpu


def lockstep(m, ref, asm, base, steps):
    """reference and device through the same sequence of calls.  A displacement increment is
    steps[k] * base; after return code 1 (yield surface overshot) the increment is retried scaled
    by the factor forces_fr left in dlpf, after code 2 (elastic unloading) it is repeated - both
    from the committed state, as main.c:2030-2063 does."""
    s = ref.RefState(m)
    s.begin_increment(); asm.begin_increment()
    codes = []
    k, scale, calls = 0, 1.0, 0
    while k < len(steps) and calls < 80:
        calls += 1
        dd = steps[k] * scale * base
        ss_ref = ref.stiff(m, s, SLVFLAG=0)
        asm.stiff()
        assert relerr(asm.skyline(), ss_ref) < TOL, f"K_t before call {calls}"
        fr, sh, dl_ref = ref.update_forces(m, s, dd, dlpf=1.0, itecnt=0)
        f, gfr, gsh, dl_dev = asm.update_forces(dd, dlpf=1.0, itecnt=0)
        assert (gfr, gsh) == (fr, sh), f"return codes, call {calls}"
        assert abs(dl_dev - dl_ref) <= 1e-9 * abs(dl_ref), f"dlpf, call {calls}"
        assert np.array_equal(asm.yldflag(), s.yldflag), f"yldflag, call {calls}"
        assert relerr(f, s.f_temp) < TOL, f"f_temp, call {calls}"
        codes.append(fr)
        if fr != 0:
            if fr == 1:
                scale *= dl_ref
            s.begin_increment(); asm.begin_increment()
            continue
        assert relerr(asm.download("EF_I"), s.ef_i) < TOL
        assert relerr(asm.download("EFFE_I"), s.efFE_i) < TOL
        s.end_iteration(); asm.end_iteration()
        s.commit(); asm.commit()
        assert np.array_equal(asm.yldflag(), s.yldflag)
        s.begin_increment(); asm.begin_increment()
        k += 1; scale = 1.0
    assert k == len(steps)
    return codes, s


def test_frame_plastic_lockstep(gpu, ref):
    """seeded displacement increments that put member ends on the yield surface, overshoot it
    (code 1, dlpf rescaled by regula_falsi) and, reversed, unload them again (code 2)"""
    m = meshgen.lattice_model(3, ANAFLAG=3, load=200.0)
    base = np.random.default_rng(11).uniform(-1.0, 1.0, size=m.NEQ)
    steps = [0.004] * 3 + [-0.0003] * 3 + [0.002] * 2 + [-0.004] * 2
    asm = cb.Assembler(m, layout=cb.CB_MAT_SKYLINE)
    codes, s = lockstep(m, ref, asm, base, steps)
    assert 1 in codes and 2 in codes and 0 in codes, codes
    assert (s.yldflag == 1).any()
    asm.close()


def test_frame_plastic_csc_tiles(gpu, ref):
    """the CSC tile kernel (compile-time joint blocks) with yielded ends against the dense scatter"""
    m = meshgen.lattice_model(3, ANAFLAG=3, load=200.0, SLVFLAG=2)
    rng = np.random.default_rng(5)
    asm = cb.Assembler(m, layout=cb.CB_MAT_CSC)
    s = ref.RefState(m)
    s.begin_increment(); asm.begin_increment()
    dd = 0.02 * rng.uniform(-1.0, 1.0, size=m.NEQ)
    n_yield = 0
    for it in range(4):
        fr, _, _ = ref.update_forces(m, s, dd * 0.5 ** it, dlpf=0.1, itecnt=it)
        _, gfr, _, _ = asm.update_forces(dd * 0.5 ** it, dlpf=0.1, itecnt=it)
        assert fr == gfr
        if fr:
            s.begin_increment(); asm.begin_increment()
            continue
        s.end_iteration(); asm.end_iteration()
        n_yield = int((s.yldflag == 1).sum())
        K_ref = ref.stiff(m, s, SLVFLAG=2).reshape(m.NEQ, m.NEQ)
        asm.stiff()
        Ap, Ai, Ax = asm.csc()
        K = np.zeros((m.NEQ, m.NEQ))
        for c in range(m.NEQ):
            K[Ai[Ap[c]:Ap[c + 1]], c] = Ax[Ap[c]:Ap[c + 1]]
        assert relerr(K, K_ref.T) < TOL
    assert n_yield > 0
    asm.close()


def test_frame_plastic_newton(gpu, ref):
    """load-controlled NR through first yield: the C host driver on the device path against the
    same loop around the reference's routines - same sub-incrementation history, converged
    displacements and load factors to 1e-9"""
    m = meshgen.lattice_model(3, ANAFLAG=3, load=200.0)
    kw = dict(lpfmax=0.252, lpf=0.05, dlpf=0.05, dlpfmax=0.05, dlpfmin=1e-6, itemax=30, submax=30,
              hist_dof=int(m.jcode.reshape(-1, 7)[m.meta["top"] - 1, 0] - 1))
    asm = cb.Assembler(m, layout=cb.CB_MAT_SKYLINE)
    d, res, hist = cb.newton_static(asm, m.q, **kw)
    d_ref, stat, hist_ref = ref_newton(m, ref, m.q, **kw)
    assert (res.status, res.increments, res.iterations) == (stat["status"], stat["increments"], stat["iterations"])
    assert np.allclose(hist[:, 0], hist_ref[:, 0], rtol=1e-9, atol=0)
    assert np.array_equal(hist[:, 1], hist_ref[:, 1])
    assert relerr(d, d_ref) < 1e-9
    assert res.increments > 100 and asm.yldflag().any()    # really went plastic (sub-incremented)
    asm.close()


def test_truss_plastic(gpu, ref):
    """trusses: axial force capped at the squash load, geometric stiffness only once yielded,
    plastic reduction beyond the surface"""
    m = meshgen.truss_model(3, ANAFLAG=3, load=50.0)
    rng = np.random.default_rng(3)
    asm = cb.Assembler(m, layout=cb.CB_MAT_BOTH)
    s = ref.RefState(m)
    s.begin_increment(); asm.begin_increment()
    capped = 0
    for it in range(4):
        dd = rng.uniform(-0.3, 0.3, size=m.NEQ)
        ref.update_forces(m, s, dd, itecnt=it)
        f, *_ = asm.update_forces(dd, itecnt=it)
        assert relerr(f, s.f_temp) < TOL
        assert relerr(asm.download("EF_I"), s.ef_i) < TOL
        py = m.carea[:m.NE_TR] * m.yld[:m.NE_TR]
        capped += int((np.abs(s.ef_i[0::2]) == py).sum())
        s.end_iteration(); asm.end_iteration()
        asm.stiff()
        assert relerr(asm.skyline(), ref.stiff(m, s, SLVFLAG=0)) < TOL
    assert capped > 0
    # beyond the surface (only reachable through an uploaded state): stiffm_tr's reduction
    ef = s.ef_i.copy(); ef[0::2] *= 1.01; ef[1::2] *= 1.01
    s.ef_i[:] = ef; s.ef_ip[:] = ef
    asm.upload("EF_IP", ef)
    asm.stiff()
    assert relerr(asm.skyline(), ref.stiff(m, s, SLVFLAG=0)) < TOL
    asm.close()


@pytest.mark.parametrize("name", ["lattice_3_plastic", "truss_3_plastic"])
def test_plastic_golden_walk(gpu, name):
    """the committed fixture (tests/golden/make_golden.py::record_plastic, recorded from the
    unmodified reference): every call's K_t, f_temp, ef_i, return code, dlpf and yldflag"""
    import os, sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_golden as G
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".npz"))
    m = G.build(name)
    asm = cb.Assembler(m, layout=cb.CB_MAT_SKYLINE)
    asm.begin_increment()
    seen = set()
    for c in range(int(g["ncalls"])):
        asm.stiff()
        assert relerr(asm.skyline(), g[f"K_sky_{c}"]) < TOL, c
        f, fr, _, dl = asm.update_forces(g[f"dd_{c}"], dlpf=1.0, itecnt=0)
        code, dl_ref = int(g[f"ret_{c}"][0]), g[f"ret_{c}"][1]
        assert fr == code and abs(dl - dl_ref) <= 1e-9 * abs(dl_ref), c
        assert np.array_equal(asm.yldflag(), g[f"yld_{c}"]), c
        assert relerr(f, g[f"f_{c}"]) < TOL, c
        seen.add(code)
        if code != 0:
            asm.begin_increment()
            continue
        assert relerr(asm.download("EF_I"), g[f"ef_{c}"]) < TOL, c
        asm.end_iteration(); asm.commit(); asm.begin_increment()
    if m.NE_FR:
        assert seen == {0, 1, 2}
    asm.close()


def test_shell_plastic_golden_walk(gpu):
    """DKT shells with Ivanov's yield criterion (stiff_sh shell.c:171-283 + stiffm_sh, forces_sh
    shell.c:1786-2325) against the fixture recorded from the unmodified reference: K_t (elastic and
    elasto-plastic), return codes, f_temp, ef_i and the plastic state after every call"""
    import os, sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_golden as G
    name = "plate_4x3_plastic"
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".npz"))
    m = G.build(name)
    asm = cb.Assembler(m, layout=cb.CB_MAT_BOTH)
    asm.begin_increment()
    codes, plastic_K = [], 0
    for c in range(int(g["ncalls"])):
        asm.stiff()
        assert relerr(asm.skyline(), g[f"K_sky_{c}"]) < TOL, c
        f, fr, sh, _ = asm.update_forces(g[f"dd_{c}"], itecnt=0)
        assert (fr, sh) == tuple(int(v) for v in g[f"ret_{c}"]), c
        codes.append(sh)
        if sh != 0:
            asm.begin_increment()
            continue
        assert relerr(f, g[f"f_{c}"]) < TOL, c
        assert relerr(asm.download("EF_I"), g[f"ef_{c}"]) < TOL, c
        assert relerr(asm.download("EFN_TEMP"), g[f"efN_{c}"]) < TOL, c
        assert relerr(asm.download("EFM_TEMP"), g[f"efM_{c}"]) < TOL, c
        chi = g[f"chi_{c}"]
        assert np.allclose(asm.download("CHI_TEMP"), chi, rtol=1e-9, atol=1e-18), c
        plastic_K += int((chi > 0).any())
        asm.end_iteration(); asm.commit(); asm.begin_increment()
        assert relerr(asm.download("EFN"), g[f"efN_{c}"]) < TOL
    assert 1 in codes and 0 in codes and plastic_K > 0
    asm.close()


def test_shell_plastic_lockstep_csc(gpu, ref):
    """a larger plate through the CSC tile kernel, reference and device in lockstep"""
    m = meshgen.plate_model(6, 5, z_bump=0.05, ANAFLAG=3, SLVFLAG=2)
    base = np.random.default_rng(21).uniform(-1.0, 1.0, size=m.NEQ)
    asm = cb.Assembler(m, layout=cb.CB_MAT_CSC)
    s = ref.RefState(m)
    s.begin_increment(); asm.begin_increment()
    scale, done, calls, yielded = 2e-4, 0, 0, 0
    while done < 12 and calls < 120:
        calls += 1
        K_ref = ref.stiff(m, s, SLVFLAG=2).reshape(m.NEQ, m.NEQ)
        asm.stiff()
        Ap, Ai, Ax = asm.csc()
        K = np.zeros((m.NEQ, m.NEQ))
        for c in range(m.NEQ):
            K[Ai[Ap[c]:Ap[c + 1]], c] = Ax[Ap[c]:Ap[c + 1]]
        assert relerr(K, K_ref.T) < TOL, calls
        dd = scale * base
        fr, sh, _ = ref.update_forces(m, s, dd, itecnt=0)
        f, gfr, gsh, _ = asm.update_forces(dd, itecnt=0)
        assert (gfr, gsh) == (fr, sh), calls
        if sh:
            scale /= 2
            s.begin_increment(); asm.begin_increment()
            continue
        assert relerr(f, s.f_temp) < TOL, calls
        assert relerr(asm.download("EFN_TEMP"), s.efN_temp) < TOL
        assert relerr(asm.download("EFM_TEMP"), s.efM_temp) < TOL
        yielded = int((s.chi_temp > 0).sum())
        s.end_iteration(); asm.end_iteration(); s.commit(); asm.commit()
        s.begin_increment(); asm.begin_increment()
        done += 1; scale *= 1.5
    assert done == 12 and yielded > 0
    asm.close()


def test_shell_plastic_newton(gpu, ref):
    """load-controlled NR of a shallow shell through first yield: forces_sh's return code 1 drives
    the sub-incrementation of main.c:2035-2063; same history, displacements to 1e-9"""
    m = meshgen.plate_model(6, 6, props=(2.1e11, 0.3, 0.05, 8050.0, 3.45e8), load=-1.2e7, z_bump=0.1,
                            ANAFLAG=3)
    kw = dict(lpfmax=0.12, lpf=0.03, dlpf=0.03, dlpfmax=0.03, dlpfmin=1e-6, solmin=1, toldisp=1e-6,
              tolforc=1e-6, tolener=1e-6, itemax=60, submax=40,
              hist_dof=int(m.jcode.reshape(-1, 7)[m.meta["centre"] - 1, 2] - 1))
    asm = cb.Assembler(m, layout=cb.CB_MAT_SKYLINE)
    d, res, hist = cb.newton_static(asm, m.q, **kw)
    d_ref, stat, hist_ref = ref_newton(m, ref, m.q, **kw)
    assert (res.status, res.increments, res.iterations) == (stat["status"], stat["increments"], stat["iterations"])
    assert res.status == 0 and res.increments > 50
    assert np.allclose(hist[:, 0], hist_ref[:, 0], rtol=1e-9, atol=0)
    assert np.array_equal(hist[:, 1], hist_ref[:, 1])
    assert relerr(d, d_ref) < 1e-9
    assert (asm.download("CHI") > 0).sum() >= 5            # really went plastic
    asm.close()


def test_frame_plastic_with_offsets_and_releases(gpu, ref):
    """inelastic frames whose members also carry rigid end offsets and bending releases: the generic
    (runtime-indexed) stiffness / force paths with the plastic reduction, reference and device in
    lockstep (stiffm_fr before release(), frame.c:268-282)"""
    import os, sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_golden as G
    m = G.build("lattice_3_offsets_releases")
    m.ANAFLAG = 3
    base = np.random.default_rng(4).uniform(-1.0, 1.0, size=m.NEQ)
    asm = cb.Assembler(m, layout=cb.CB_MAT_BOTH)
    codes, s = lockstep(m, ref, asm, base, [0.004] * 3 + [-0.0003] * 2 + [0.003] * 2)
    assert 1 in codes and 0 in codes and (s.yldflag == 1).any()
    # CSC of the final state against the dense scatter
    K_ref = ref.stiff(m, s, SLVFLAG=2).reshape(m.NEQ, m.NEQ)
    asm.stiff()
    Ap, Ai, Ax = asm.csc()
    K = np.zeros((m.NEQ, m.NEQ))
    for c in range(m.NEQ):
        K[Ai[Ap[c]:Ap[c + 1]], c] = Ax[Ap[c]:Ap[c + 1]]
    assert relerr(K, K_ref.T) < TOL
    asm.close()

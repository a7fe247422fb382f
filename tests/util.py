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


def csc_vs_skyline(m, Ap, Ai, Ax, ss):
    """device CSC (full, unsymmetric storage, structural joint-block pattern) against the reference's
    skyline vector, entry by entry: entry (i <= j) lives at ss[maxa[j-1] + (j-i) - 1] (1-based,
    model.c:1269-1278; shell.c:307-329 / frame.c:330-352 scatter the element's upper triangle there).
    The lower triangle of the CSC is compared with the same skyline entries (K_t is symmetric).
    Structural CSC entries outside the skyline profile must be exactly zero.  Returns
    (norm-wise relative difference, entry-wise difference scaled by sqrt(K_ii K_jj))."""
    n = m.NEQ
    maxa = np.asarray(m.maxa, dtype=np.int64)
    cols = np.repeat(np.arange(n, dtype=np.int64), np.diff(Ap))
    rows = np.asarray(Ai, dtype=np.int64)
    lo, hi = np.minimum(rows, cols), np.maximum(rows, cols)
    addr = maxa[hi] - 1 + (hi - lo)
    inside = addr < maxa[hi + 1] - 1
    assert np.all(Ax[~inside] == 0.0), "non-zero CSC entry outside the reference's skyline profile"
    # every skyline entry the reference wrote must be present in the CSC pattern
    seen = np.zeros(ss.size, dtype=bool)
    seen[addr[inside]] = True
    assert np.all(ss[~seen] == 0.0), "reference skyline entry missing from the CSC pattern"
    want = ss[addr[inside]]
    diff = np.abs(Ax[inside] - want)
    diag = np.abs(ss[maxa[:-1] - 1])
    scale = np.sqrt(diag[lo[inside]] * diag[hi[inside]])
    assert np.all(scale > 0)
    return diff.max() / np.abs(ss).max(), (diff / scale).max()


TOL = 1e-12


def walk(m, ref, asm, n_iter=3, scale=1e-4, seed=1, dlpf=0.25):
    """drive reference and device through the same sequence of calls of the Newton loop
    (main.c:1833-2134) and compare everything after every call"""
    s = ref.RefState(m)
    s.begin_increment(); asm.begin_increment()
    rng = np.random.default_rng(seed)
    for it in range(n_iter):
        ss_ref = ref.stiff(m, s, SLVFLAG=0)
        asm.stiff()
        assert relerr(asm.skyline(), ss_ref) < TOL, f"skyline K_t iter {it}"
        dd = rng.uniform(-scale, scale, size=m.NEQ)
        fr, sh, _ = ref.update_forces(m, s, dd, dlpf=dlpf, itecnt=it)
        f, gfr, gsh, _ = asm.update_forces(dd, dlpf=dlpf, itecnt=it)
        assert (fr, sh) == (gfr, gsh)
        assert relerr(f, s.f_temp) < TOL, f"f_temp iter {it}"
        assert relerr(asm.download("EF_I"), s.ef_i) < TOL
        assert relerr(asm.download("X_TEMP"), s.x_temp) == 0.0
        assert relerr(asm.download("X_IP"), s.x_ip) == 0.0
        assert relerr(asm.download("D_TEMP"), s.d_temp) == 0.0
        for nm in ("C1", "C2", "C3"):
            assert relerr(asm.download(nm + "_I"), getattr(s, nm.lower() + "_i")) < 1e-15
            assert relerr(asm.download(nm + "_IP"), getattr(s, nm.lower() + "_ip")) < 1e-15
        assert relerr(asm.download("DEFFAREA_I"), s.deffarea_i) < 1e-15
        assert relerr(asm.download("DEFSLEN_I"), s.defslen_i) < 1e-15
        assert relerr(asm.download("DEFLLEN_I"), s.defllen_i) < 1e-15
        assert relerr(asm.download("EFFE_I"), s.efFE_i) < TOL
        assert relerr(asm.download("XFR_TEMP"), s.xfr_temp) < 1e-15
        s.end_iteration(); asm.end_iteration()
        assert relerr(asm.download("C1_IP"), s.c1_ip) < 1e-15
        assert relerr(asm.download("EFFE_IP"), s.efFE_ip) < TOL
    s.commit(); asm.commit()
    assert relerr(asm.download("EF"), s.ef) < TOL
    assert relerr(asm.download("EFFE"), s.efFE) < TOL
    assert relerr(asm.download("X"), s.x) == 0.0
    assert relerr(asm.download("D"), s.d) == 0.0
    assert relerr(asm.download("F"), s.f) < TOL
    assert relerr(asm.download("XFR"), s.xfr) < 1e-15
    return s


def ref_newton(m, ref, q, lpfmax=1.0, lpf=0.1, dlpf=0.1, dlpfmax=0.5, dlpfmin=1e-4, itemax=20,
               submax=5, solmin=10, toldisp=1e-8, tolforc=1e-8, tolener=1e-8, algflag=1, hist_dof=-1):
    """The reference's NR / MNR loop (main.c:1824-2152) around the reference's OWN routines
    (stiff_*, solve() -> skyfact/skysolve, updatc, forces_*, test) - the converged-solution oracle."""
    import ctypes as C
    l = ref.set_model(m)
    s = ref.RefState(m)
    n = m.NEQ
    qtot = np.zeros(n); fp = np.zeros(n); f_ip = np.zeros(n); r = np.zeros(n)
    hist = []
    solcnt = subcnt = 0
    stat = dict(increments=0, iterations=0, status=0, lpf=0.0)
    ss = None
    while True:
        if lpf > lpfmax:
            lpf = lpfmax
        qtot[:] = q * lpf; fp[:] = s.f
        dlpfp = dlpf
        s.begin_increment()
        itecnt = 0; frfr = frsh = 0
        while True:
            r[:] = qtot - s.f_temp
            refactor = algflag == 1 or (algflag == 2 and itecnt == 0)
            if refactor:
                ss = ref.stiff(m, s, SLVFLAG=0)
            dd, _, _ = ref.skyline_solve(m, ss, r, fact=0 if refactor else 1)
            f_ip[:] = s.f_temp
            frfr, frsh, dlpf = ref.update_forces(m, s, dd, dlpf=dlpf, itecnt=itecnt)
            stat["iterations"] += 1
            if itecnt == 0:
                intener1 = C.c_double(float(ref.lib().dot(ref.P(dd), ref.P(qtot - fp), C.c_int(n))))
            conv = C.c_int(0)
            ref.set_model(m)
            err = l.test(ref.P(s.d_temp), ref.P(dd), ref.P(s.f_temp), ref.P(fp), ref.P(qtot),
                         ref.P(f_ip), C.byref(intener1), C.byref(conv), C.byref(C.c_double(toldisp)),
                         C.byref(C.c_double(tolforc)), C.byref(C.c_double(tolener)))
            assert err == 0
            s.end_iteration()
            itecnt += 1
            if not (conv.value != 0 and frfr == 0 and frsh == 0 and itecnt <= itemax):
                break
        if frfr == 2:
            dlpf = dlpfp
        elif (conv.value != 0 or frfr or frsh) and subcnt <= submax:
            if lpf == lpfmax:
                stat["status"] = 4; break
            if dlpfp == dlpfmin:
                stat["status"] = 5; break
            if frfr != 1:
                dlpf = dlpfp / 2
            dlpf = max(dlpf, dlpfmin)
            lpf = lpf - dlpfp + dlpf
            subcnt += 1; solcnt = 0
        elif subcnt > submax:
            stat["status"] = 6; break
        else:
            stat["increments"] += 1
            s.commit()
            stat["lpf"] = lpf
            hist.append((lpf, itecnt, s.d[hist_dof] if hist_dof >= 0 else 0.0))
            solcnt += 1; subcnt = 0
            if solcnt >= solmin:
                dlpf = min(dlpf * 2, dlpfmax); solcnt = 0
            lpf += dlpf
        if not lpf <= lpfmax:
            break
    return s.d.copy(), stat, np.array(hist)

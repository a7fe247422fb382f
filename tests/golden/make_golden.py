"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref, compiled from
/root/reference by oracle/Makefile).  The reference ships no tests and no expected outputs
(SURVEY.md section 4), so these vectors - arrays read back in memory from its own routines at
full double precision - are the golden fixtures for this path.  Run from the repo root:

    python tests/golden/make_golden.py

Each file holds, for a model rebuilt through cubens_b200.meshgen with the arguments in
``cases()`` and a fixed seeded sequence of Newton-like iterations, the reference's skyline K_t,
dense K_t (small cases), f_temp, ef_i, triads and the lumped mass."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "cu-bens_b200", "python"))

from cubens_b200 import meshgen, model as M  # noqa: E402
from oracle import refbind as R              # noqa: E402


def cases():
    """name -> (factory name, kwargs, post-processing id)"""
    return {
        "plate_3x2": ("plate_model", dict(nx=3, ny=2, z_bump=0.02), None),
        "plate_5x4_flat": ("plate_model", dict(nx=5, ny=4), None),
        "plate_lin_3x3": ("plate_model", dict(nx=3, ny=3, ANAFLAG=1), None),
        "truss_3": ("truss_model", dict(n=3), None),
        "truss_lin_3": ("truss_model", dict(n=3, ANAFLAG=1), None),
        "lattice_3": ("lattice_model", dict(n=3), "fe"),
        "lattice_3_offsets_releases": ("lattice_model", dict(n=3), "offrel"),
        "lattice_lin_3": ("lattice_model", dict(n=3, ANAFLAG=1), None),
        "lattice_3_plastic": ("lattice_model", dict(n=3, ANAFLAG=3, load=200.0), "fe"),
        "truss_3_plastic": ("truss_model", dict(n=3, ANAFLAG=3, load=50.0), None),
        "plate_4x3_plastic": ("plate_model", dict(nx=4, ny=3, z_bump=0.03, ANAFLAG=3), None),
        "brick_2x2x2": ("brick_model", dict(nx=2, ny=2, nz=2, distort=0.1), None),
        "brick_skin_2x2x1": ("brick_model", dict(nx=2, ny=2, nz=1, distort=0.05, skin=True), None),
    }


def build(name):
    fac, kw, post = cases()[name]
    m = getattr(meshgen, fac)(**kw)
    if post in ("fe", "offrel"):
        rng = np.random.default_rng(11)
        m.efFE_ref[:] = rng.uniform(-1, 1, m.efFE_ref.size)
    if post == "offrel":
        rng = np.random.default_rng(12)
        m.osflag[::3] = 1
        m.offset[:] = rng.uniform(-5, 5, m.offset.size) * np.repeat(m.osflag, 6)
        rel = m.mendrel.reshape(-1, 5)
        rel[1::4] = [1, 1, 0, 0, 0]; rel[2::4] = [1, 0, 1, 1, 0]
        rel[3::7] = [1, 1, 1, 1, 0]; rel[5::11] = [1, 1, 1, 1, 1]
        xfr, ll, lx, ly, lz = M.frame_geometry(m.x, m.minc.reshape(-1, 2) - 1, m.auxpt, m.offset,
                                               m.osflag)
        m.xfr[:] = xfr.reshape(-1); m.llength[:] = ll
        m.c1[:] = lx.reshape(-1); m.c2[:] = ly.reshape(-1); m.c3[:] = lz.reshape(-1)
    return m


def record(name, B=R, n_iter=3):
    """replay the fixed sequence through backend B (refbind or oraclebind)"""
    m = build(name)
    out = {"NEQ": m.NEQ, "lss": m.lss, "jcode": m.jcode, "mcode": m.mcode}
    if m.maxa is not None:
        out["maxa"] = m.maxa
    s = R.RefState(m)
    rng = np.random.default_rng(2026)
    if m.NE_BR:
        out["K_dense"] = B.stiff(m, s, SLVFLAG=2, gen="c")
        out["M_dense"] = B.mass(m, s, SLVFLAG=2)
        return m, out
    if m.ANAFLAG == 1:
        out["K_sky"] = B.stiff(m, s, SLVFLAG=0, gen="c")
        d = rng.uniform(-1e-3, 1e-3, m.NEQ)
        out["d"] = d
        out["f_lin"] = B.forces_linear(m, s, d)
        out["ef_lin"] = s.ef.copy()
        return m, out
    if m.ANAFLAG == 3 and m.NE_SH:
        return m, record_plastic_shell(m, s, out, B)
    if m.ANAFLAG == 3:
        return m, record_plastic(m, s, out, B)
    s.begin_increment()
    for it in range(n_iter):
        out[f"K_sky_{it}"] = B.stiff(m, s, SLVFLAG=0)
        if m.NEQ <= 200:
            out[f"K_dense_{it}"] = B.stiff(m, s, SLVFLAG=2)
        dd = rng.uniform(-1e-4, 1e-4, m.NEQ) * (100.0 if m.NE_FR or m.NE_TR else 1.0)
        out[f"dd_{it}"] = dd
        B.update_forces(m, s, dd, dlpf=0.25, itecnt=it)
        out[f"f_{it}"] = s.f_temp.copy(); out[f"ef_{it}"] = s.ef_i.copy()
        out[f"c1_{it}"] = s.c1_i.copy(); out[f"c2_{it}"] = s.c2_i.copy(); out[f"c3_{it}"] = s.c3_i.copy()
        if m.NE_FR:
            out[f"efFE_{it}"] = s.efFE_i.copy()
        s.end_iteration()
    s.commit()
    out["mass"] = B.mass(m, s, SLVFLAG=0) if B is R else B.mass(m, s)
    return m, out


PLASTIC_STEPS = [0.004] * 3 + [-0.0003] * 3 + [0.002] * 2 + [-0.004] * 2


def record_plastic(m, s, out, B):
    """material-nonlinear walk (ANAFLAG 3): displacement increments steps[k] * base; after forces_fr
    returns 1 (yield surface overshot) the increment is retried scaled by the factor it left in
    dlpf, after 2 (elastic unloading) it is repeated - from the committed state, as
    main.c:2030-2063 does.  Every call's K_t, f_temp, ef_i, return code, dlpf and yldflag are kept."""
    base = np.random.default_rng(11).uniform(-1.0, 1.0, size=m.NEQ)
    steps = PLASTIC_STEPS if m.NE_FR else [0.3, 0.2, -0.4, 0.1]
    out["base"] = base; out["steps"] = np.array(steps)
    s.begin_increment()
    k, scale, call = 0, 1.0, 0
    while k < len(steps) and call < 80:
        dd = steps[k] * scale * base
        out[f"K_sky_{call}"] = B.stiff(m, s, SLVFLAG=0)
        fr, _, dl = B.update_forces(m, s, dd, dlpf=1.0, itecnt=0)
        out[f"dd_{call}"] = dd; out[f"f_{call}"] = s.f_temp.copy()
        out[f"ret_{call}"] = np.array([fr, dl]); out[f"yld_{call}"] = s.yldflag.copy()
        if fr == 0:
            out[f"ef_{call}"] = s.ef_i.copy()
            if m.NE_FR:
                out[f"efFE_{call}"] = s.efFE_i.copy()
        call += 1
        if fr != 0:
            if fr == 1:
                scale *= dl
            s.begin_increment()
            continue
        s.end_iteration(); s.commit(); s.begin_increment()
        k += 1; scale = 1.0
    out["ncalls"] = call
    return out


def record_plastic_shell(m, s, out, B, n_steps=16):
    """material-nonlinear shells (ANAFLAG 3, Ivanov's criterion).  forces_sh returns 1 whenever a
    vertex ends more than 10*phitol beyond the yield surface, so the walk does what main.c:2035-2063
    does: halve the increment and repeat it from the committed state; converged increments grow by
    1.5, the direction reverses after step 10 (unloading).  Every call's K_t, return code, f_temp,
    ef_i and the plastic state (chi, efN, efM) are kept."""
    base = np.random.default_rng(5).uniform(-1.0, 1.0, size=m.NEQ)
    out["base"] = base
    s.begin_increment()
    scale, done, call, dds = 2e-4, 0, 0, []
    while done < n_steps and call < 150:
        dd = (1.0 if done < 10 else -1.0) * scale * base
        out[f"K_sky_{call}"] = B.stiff(m, s, SLVFLAG=0)
        fr, sh, dl = B.update_forces(m, s, dd, itecnt=0)
        out[f"dd_{call}"] = dd; out[f"ret_{call}"] = np.array([fr, sh])
        if sh == 0:
            out[f"f_{call}"] = s.f_temp.copy(); out[f"ef_{call}"] = s.ef_i.copy()
            out[f"chi_{call}"] = s.chi_temp.copy(); out[f"efN_{call}"] = s.efN_temp.copy()
            out[f"efM_{call}"] = s.efM_temp.copy()
        call += 1
        if sh != 0:
            scale /= 2; s.begin_increment()
            continue
        s.end_iteration(); s.commit(); s.begin_increment()
        done += 1; scale *= 1.5
    out["ncalls"] = call
    return out


DECKS = "/root/reference/Sample_Input_Files/"


def deck_cases():
    """the reference's shipped sample decks (its only fixtures)"""
    return {"deck_5a_truss": "model_def_5a_truss.txt", "deck_5b_frame": "model_def_5b_frame.txt",
            "deck_5c_shell": "model_def_5c_shell.txt", "deck_5d_shell": "model_def_5d_shell.txt"}


def record_deck(name, B=R):
    """parse the shipped deck (cubens_b200.deck), keep the parsed model in the fixture (the GPU box
    has no /root/reference) and record what the reference computes on it:
      5a  static MNR: the whole run (reference NR loop around its own routines + ben.exe's print)
      5b/5c/5d  dynamic decks: K_t + lumped mass on the deck's mesh, and for the geometric-nonlinear
          ones a seeded two-iteration walk (5b is ANAFLAG 3 in the deck; recorded as ANAFLAG 2)."""
    import subprocess, tempfile
    from cubens_b200 import deck
    from cubens_b200.model import model_to_dict
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from util import ref_newton
    text = open(DECKS + deck_cases()[name]).read()
    m, params, _ = deck.read_deck(text)
    if m.ANAFLAG == 3:
        m.ANAFLAG = 2
    out = model_to_dict(m)
    s = R.RefState(m)
    if params is not None and m.ANAFLAG == 2:
        d, stat, hist = ref_newton(m, B, m.q, hist_dof=0, **params)
        out["params"] = np.array([params[k] for k in ("lpfmax", "lpf", "dlpf", "dlpfmax", "dlpfmin",
                                  "itemax", "submax", "solmin", "toldisp", "tolforc", "tolener",
                                  "algflag")], dtype=np.float64)
        out["d_ref"] = d; out["hist_ref"] = hist
        out["stat_ref"] = np.array([stat["increments"], stat["iterations"], stat["status"]])
        exe = os.path.join(ROOT, "oracle", "_ref", "ben.exe")
        with tempfile.TemporaryDirectory() as td:
            open(os.path.join(td, "model_def.txt"), "w").write(text.replace("\r", ""))
            subprocess.check_call([exe], cwd=td, stdout=subprocess.DEVNULL)
            last = open(os.path.join(td, "results2.txt")).read().strip().splitlines()[-1].split()
        out["ben_exe_last"] = np.array([float(v) for v in last])
        return m, out
    out["K_sky"] = B.stiff(m, s, SLVFLAG=0, gen="c")
    out["mass"] = B.mass(m, s, SLVFLAG=0) if B is R else B.mass(m, s)
    if m.ANAFLAG == 2:
        s = R.RefState(m); s.begin_increment()
        rng = np.random.default_rng(77)
        for it in range(2):
            dd = rng.uniform(-1e-4, 1e-4, m.NEQ); out[f"dd_{it}"] = dd
            B.update_forces(m, s, dd, dlpf=0.25, itecnt=it)
            out[f"f_{it}"] = s.f_temp.copy(); out[f"ef_{it}"] = s.ef_i.copy()
            s.end_iteration()
            out[f"K_sky_{it}"] = B.stiff(m, s, SLVFLAG=0)
    return m, out


def run_cases():
    """full transient runs of the shipped decks through the UNMODIFIED reference driver
    (oracle/_ref/ben_capture.exe = main.c with its output() rows recorded at full precision)"""
    return {"run_5b_frame": "model_def_5b_frame.txt", "run_5c_shell": "model_def_5c_shell.txt",
            "run_5d_shell": "model_def_5d_shell.txt"}


ARC_SHELL = (2.1e11, 0.3, 0.02, 8050.0, 3.45e8)
ARC = dict(dk=-0.0005, alpha=1.0, psi_thresh=0.1, iteopt=5, lpfmax=1.7102, dkimax=0.01, itemax=50,
           submax=10, imagmax=10, negmax=10, toldisp=1e-3, tolforc=1e-3, tolener=1e-3)


def arc_model():
    """shallow 4x4 shell cap for the arc-length (ALGFLAG 3) run; the reference's arc length
    collapses to tiny increments after the first one (beta, main.c:2626-2629), so lpfmax is set just
    above the first increment's load factor: one prescribed-displacement increment + ~15 MSAL ones"""
    return meshgen.plate_model(4, 4, props=ARC_SHELL, load=-1.0e5, z_bump=0.1, ALGFLAG=3)


def record_arc():
    import subprocess, tempfile
    from cubens_b200 import deck
    m = arc_model()
    c = m.meta["centre"]
    a = ARC
    tail = [f"{c},3,{a['dk']!r}", repr(a["alpha"]), repr(a["psi_thresh"]), str(a["iteopt"]),
            f"{a['lpfmax']!r},{a['dkimax']!r}", f"{a['itemax']},{a['submax']},{a['imagmax']},{a['negmax']}",
            f"{a['toldisp']!r},{a['tolforc']!r},{a['tolener']!r}"]
    text = deck.write_shell_deck(m, ARC_SHELL, [(c, 3, -1.0e5)], tail)
    exe = os.path.join(ROOT, "oracle", "_ref", "ben_capture.exe")
    with tempfile.TemporaryDirectory() as td:
        open(os.path.join(td, "model_def.txt"), "w").write(text)
        subprocess.run([exe], cwd=td, stdout=subprocess.DEVNULL, timeout=60, check=True)
        raw = open(os.path.join(td, "capture.bin"), "rb").read()
        ok = "Solution successful" in open(os.path.join(td, "results1.txt")).read()
    neq, nrows = np.frombuffer(raw[:16], dtype=np.int64)
    assert neq == m.NEQ and ok
    hist = np.frombuffer(raw[16:], dtype=np.float64).reshape(nrows, neq + 2).copy()
    return m, {"hist": hist, "dkdof": int(m.jcode.reshape(-1, 7)[c - 1, 2] - 1)}


def record_run(name):
    import subprocess, tempfile
    from cubens_b200 import deck
    from cubens_b200.model import model_to_dict
    text = open(DECKS + run_cases()[name]).read().replace("\r", "")
    if name == "run_5d_shell":      # shipped with RFLAG = 1 (restart from a results8.txt that is not shipped)
        lines = text.split("\n"); assert lines[2] == "4,1"; lines[2] = "4,0"; text = "\n".join(lines)
    m, _, dyn = deck.read_deck(text)
    out = model_to_dict(m)
    for k, v in dyn.items():
        if k == "params":
            out["dyn_params"] = np.array([v[q] for q in ("lpfmax", "lpf", "dlpf", "dlpfmax", "dlpfmin",
                                          "itemax", "submax", "solmin", "toldisp", "tolforc", "tolener")])
        else:
            out["dyn_" + k] = np.asarray(v)
    exe = os.path.join(ROOT, "oracle", "_ref", "ben_capture.exe")
    with tempfile.TemporaryDirectory() as td:
        open(os.path.join(td, "model_def.txt"), "w").write(text)
        subprocess.run([exe], cwd=td, stdout=subprocess.DEVNULL, timeout=120, check=True)
        raw = open(os.path.join(td, "capture.bin"), "rb").read()
    neq, nrows = np.frombuffer(raw[:16], dtype=np.int64)
    assert neq == m.NEQ
    out["hist"] = np.frombuffer(raw[16:], dtype=np.float64).reshape(nrows, neq + 2).copy()
    return m, out


if __name__ == "__main__":
    m, out = record_arc()
    np.savez_compressed(os.path.join(HERE, "run_arc_shell.npz"), **out)
    print("run_arc_shell NEQ", m.NEQ, "rows", out["hist"].shape[0])
    for name in run_cases():
        m, out = record_run(name)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, "NEQ", m.NEQ, "rows", out["hist"].shape[0])
    assert R.available(), "build oracle/_ref first (make -C oracle ref)"
    for name in deck_cases():
        m, out = record_deck(name)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, "NEQ", m.NEQ, "arrays", len(out))
    for name in cases():
        m, out = record(name)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, "NEQ", m.NEQ, "arrays", len(out))

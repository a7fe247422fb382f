"""Full transient runs of the reference's shipped decks on the device path, against the
displacement history of the UNMODIFIED reference driver (tests/golden/run_*.npz, recorded by
oracle/_ref/ben_capture.exe = main.c with every output() row captured at full double precision):

  model_def_5b_frame.txt  ANAFLAG 3 / ALGFLAG 5: inelastic (plastic-hinge) frames, nonlinear
                          Newmark with generalized-alpha damping (rho = 0.9), 17 time steps; K_t, lumped
                          mass and f_int rebuilt on the device every Newton iteration
  model_def_5c_shell.txt  ANAFLAG 1 / ALGFLAG 4 (BASELINE.json configs[1]): DKT shells, linear Newmark

through the C host drivers cb_newmark_nonlinear / cb_newmark_linear (main.c:3305-3960, 3143-3303 +
solve.c:139-536).  Tolerance: 1e-9 relative on every time step's displacement vector."""
import os

import numpy as np
import pytest

import cubens_b200 as cb
from cubens_b200.model import model_from_dict
from util import relerr

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    m = model_from_dict(g)
    dyn = {k[4:]: g[k] for k in g.files if k.startswith("dyn_") and k != "dyn_params"}
    for k in ("ntstps", "nbc"):
        dyn[k] = int(dyn[k])
    for k in ("dt", "alpham", "alphaf"):
        dyn[k] = float(dyn[k])
    if "dyn_params" in g.files:
        keys = ("lpfmax", "lpf", "dlpf", "dlpfmax", "dlpfmin", "itemax", "submax", "solmin", "toldisp",
                "tolforc", "tolener")
        dyn["params"] = dict(zip(keys, g["dyn_params"]))
        for k in ("itemax", "submax", "solmin"):
            dyn["params"][k] = int(dyn["params"][k])
    return g, m, dyn


SOLVERS = ("skyline", "csc")      # csc: device-built CSC -> host sparse LDL^T


def _asm(m, solver):
    return cb.Assembler(m, layout=cb.CB_MAT_CSC if solver == "csc" else cb.CB_MAT_SKYLINE)


@pytest.mark.parametrize("solver", SOLVERS)
def test_deck_5b_inelastic_frame_newmark(gpu, solver):
    g, m, dyn = _load("run_5b_frame")
    assert (m.ANAFLAG, m.ALGFLAG) == (3, 5)
    asm = _asm(m, solver)
    hist, res = cb.newmark(asm, dyn, nonlinear=True, csc=solver == "csc")
    ref = g["hist"]
    assert res.status == 0 and hist.shape == ref.shape
    assert np.allclose(hist[:, 0], ref[:, 0], rtol=1e-12, atol=0)         # times
    assert np.array_equal(hist[:, 1], ref[:, 1])                          # Newton iterations per step
    for k in range(ref.shape[0]):
        assert relerr(hist[k, 2:], ref[k, 2:]) < 1e-9, k
    assert np.abs(ref[-1, 2:]).max() > 0
    asm.close()


@pytest.mark.parametrize("solver", SOLVERS)
def test_deck_5c_shell_linear_newmark(gpu, solver):
    g, m, dyn = _load("run_5c_shell")
    assert (m.ANAFLAG, m.ALGFLAG) == (1, 4)
    asm = _asm(m, solver)
    hist, res = cb.newmark(asm, dyn, nonlinear=False, csc=solver == "csc")
    ref = g["hist"][:-1]          # the reference's last row is its final output() of the (zero) d
    assert res.status == 0 and hist.shape == ref.shape
    assert np.allclose(hist[:, 0], ref[:, 0], rtol=1e-12, atol=0)
    for k in range(1, ref.shape[0]):
        assert relerr(hist[k, 2:], ref[k, 2:]) < 1e-9, k
    assert np.abs(ref[-1, 2:]).max() > 0
    asm.close()


@pytest.mark.parametrize("solver", SOLVERS)
def test_arclength_shell_cap(gpu, solver):
    """modified spherical arc-length (ALGFLAG 3, main.c:2158-3141 + quad() arc.c:70): a shallow DKT
    shell cap through the C host driver cb_arclength_static against the unmodified reference driver
    (run_arc_shell.npz): the prescribed-displacement increment, then 20 MSAL increments - load
    factors, iteration counts and displacement vectors"""
    import sys
    sys.path.insert(0, GOLD)
    import make_golden as G
    g = np.load(os.path.join(GOLD, "run_arc_shell.npz"))
    ref = g["hist"]
    m = G.arc_model()
    a = dict(G.ARC)
    asm = _asm(m, solver)
    hist, res = cb.arclength_static(asm, m.q, dkdof=int(g["dkdof"]), csc=solver == "csc", **a)
    assert res.status == 0 and hist.shape == ref.shape
    assert np.array_equal(hist[:, 1], ref[:, 1])
    assert np.allclose(hist[:, 0], ref[:, 0], rtol=1e-9, atol=0)
    for k in range(ref.shape[0]):
        assert relerr(hist[k, 2:], ref[k, 2:]) < 1e-9, k
    asm.close()


@pytest.mark.parametrize("solver", SOLVERS)
def test_deck_5d_shell_newmark_support_motion(gpu, solver):
    """model_def_5d_shell.txt (ANAFLAG 2 / ALGFLAG 5, RFLAG set to 0): geometric-nonlinear DKT shells
    driven by prescribed support motion (NBC = 2: heavy support masses, matpart() on the effective
    matrix, inertial reactions) - 20 time steps against the unmodified reference driver"""
    g, m, dyn = _load("run_5d_shell")
    assert (m.ANAFLAG, m.ALGFLAG) == (2, 5) and dyn["nbc"] == 2
    asm = _asm(m, solver)
    hist, res = cb.newmark(asm, dyn, nonlinear=True, csc=solver == "csc")
    ref = g["hist"]
    assert res.status == 0 and hist.shape == ref.shape
    assert np.array_equal(hist[:, 1], ref[:, 1])
    for k in range(ref.shape[0]):
        assert relerr(hist[k, 2:], ref[k, 2:]) < 1e-9, k
    assert np.abs(ref[-1, 2:]).max() > 1e-3
    asm.close()

"""The reference's shipped sample decks (BASELINE.json configs[0] = model_def_5a_truss.txt as
shipped, configs[1] = model_def_5c_shell.txt, plus 5b / 5d) through the device path, against
fixtures recorded from the reference itself (tests/golden/make_golden.py::record_deck; the parsed
model travels inside the fixture because the GPU box has no /root/reference)."""
import os

import numpy as np
import pytest

import cubens_b200 as cb
from cubens_b200.model import model_from_dict
from util import relerr, TOL

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    return g, model_from_dict(g)


def test_deck_5a_truss_full_run(gpu):
    """configs[0]: nonlinear static truss, MNR, 10 increments x 2 iterations, through the C host
    driver; converged displacements / load factors to 1e-9 of the reference loop and to the 7
    digits ben.exe prints in results2.txt"""
    g, m = _load("deck_5a_truss")
    keys = ("lpfmax", "lpf", "dlpf", "dlpfmax", "dlpfmin", "itemax", "submax", "solmin", "toldisp",
            "tolforc", "tolener", "algflag")
    p = dict(zip(keys, g["params"]))
    for k in ("itemax", "submax", "solmin", "algflag"):
        p[k] = int(p[k])
    asm = cb.Assembler(m, layout=cb.CB_MAT_SKYLINE)
    d, res, hist = cb.newton_static(asm, m.q, hist_dof=0, **p)
    assert res.status == 0
    assert [res.increments, res.iterations] == list(g["stat_ref"][:2])
    assert relerr(d, g["d_ref"]) < 1e-9
    assert np.allclose(hist[:, 0], g["hist_ref"][:, 0], rtol=1e-9, atol=0)
    assert np.array_equal(hist[:, 1], g["hist_ref"][:, 1])
    ben = g["ben_exe_last"]                       # lpf, iterations, d[0], d[1] as printed (%e)
    assert abs(res.lpf - ben[0]) < 1e-6 and hist[-1, 1] == ben[1]
    assert np.allclose(d, ben[2:4], rtol=1e-6, atol=0)
    asm.close()


@pytest.mark.parametrize("name", ["deck_5b_frame", "deck_5c_shell", "deck_5d_shell"])
def test_deck_stiffness_mass_forces(gpu, name):
    g, m = _load(name)
    asm = cb.Assembler(m, layout=cb.CB_MAT_SKYLINE)
    asm.stiff(cb.CB_GEN_COMMITTED)
    assert relerr(asm.skyline(), g["K_sky"]) < TOL
    if m.ANAFLAG == 2:
        asm.begin_increment()
        for it in range(2):
            f, *_ = asm.update_forces(g[f"dd_{it}"], dlpf=0.25, itecnt=it)
            assert relerr(f, g[f"f_{it}"]) < TOL
            assert relerr(asm.download("EF_I"), g[f"ef_{it}"]) < TOL
            asm.end_iteration()
            asm.stiff()
            assert relerr(asm.skyline(), g[f"K_sky_{it}"]) < TOL
    asm2 = cb.Assembler(m, layout=cb.CB_MAT_SKYLINE)
    assert relerr(asm2.mass(), g["mass"]) < TOL
    asm.close(); asm2.close()

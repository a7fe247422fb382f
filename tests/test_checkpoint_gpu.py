"""Restart from a binary checkpoint of the device state (the reference's own
checkpoint is "%e" text, misc.c:494-716, so its restarts drift): a run interrupted after some
increments, restored into a FRESH handle and continued, must give bit-identical K_t, f_int,
element forces and plastic state to the uninterrupted run - elastic shells after mass_* rewrote the
reference geometry, and inelastic frames with yielded member ends."""
import numpy as np
import pytest

import cubens_b200 as cb
from cubens_b200 import meshgen

pytestmark = pytest.mark.gpu


def _advance(asm, dds, with_mass):
    out = []
    for dd in dds:
        asm.stiff()
        K = asm.skyline()
        sm = asm.mass() if with_mass else None
        f, fr, sh, _ = asm.update_forces(dd, dlpf=0.1, itecnt=0)
        asm.end_iteration()
        if fr or sh:
            asm.begin_increment()
        else:
            asm.commit(); asm.begin_increment()
        out.append((K, sm, f, asm.download("EF"), fr, sh))
    return out


@pytest.mark.parametrize("kind", ["shell_mass", "frame_plastic", "shell_plastic"])
def test_restart_is_bit_identical(gpu, tmp_path, kind):
    if kind == "shell_mass":
        m = meshgen.plate_model(6, 5, z_bump=0.04); scale = 2e-4
    elif kind == "frame_plastic":
        m = meshgen.lattice_model(3, ANAFLAG=3, load=200.0); scale = 2e-3
    else:
        m = meshgen.plate_model(4, 3, z_bump=0.03, ANAFLAG=3); scale = 4e-6
    rng = np.random.default_rng(9)
    base = rng.uniform(-1, 1, m.NEQ)
    dds = [scale * base * (1 + 0.1 * rng.uniform(-1, 1, m.NEQ)) for _ in range(8)]
    a = cb.Assembler(m, layout=cb.CB_MAT_SKYLINE)
    a.begin_increment()
    _advance(a, dds[:5], kind == "shell_mass")
    path = tmp_path / "ck.bin"
    a.checkpoint_save(path)
    want = _advance(a, dds[5:], kind == "shell_mass")
    state_a = {k: a.download(k) for k in ("X", "D", "F", "C1", "DEFFAREA" if m.NE_SH else "DEFLLEN")}
    b = cb.Assembler(m, layout=cb.CB_MAT_SKYLINE)
    b.checkpoint_load(path)
    got = _advance(b, dds[5:], kind == "shell_mass")
    for (K1, s1, f1, e1, fr1, sh1), (K2, s2, f2, e2, fr2, sh2) in zip(want, got):
        assert np.array_equal(K1, K2) and np.array_equal(f1, f2) and np.array_equal(e1, e2)
        assert (fr1, sh1) == (fr2, sh2)
        if s1 is not None:
            assert np.array_equal(s1, s2)
    for k, v in state_a.items():
        assert np.array_equal(b.download(k), v), k
    if kind == "frame_plastic":
        assert np.array_equal(a.yldflag(), b.yldflag()) and a.yldflag().any()
    if kind == "shell_plastic":
        assert np.array_equal(a.download("CHI"), b.download("CHI"))
        assert np.array_equal(a.download("EFN"), b.download("EFN"))
    # a checkpoint of another model is refused
    other = cb.Assembler(meshgen.plate_model(3, 3), layout=cb.CB_MAT_SKYLINE)
    with pytest.raises(cb.CubensError):
        other.checkpoint_load(path)
    a.close(); b.close(); other.close()

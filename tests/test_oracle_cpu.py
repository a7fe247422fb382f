"""CPU suite: pins the oracle (plain-C restatement) against the golden vectors and, where the
compiled reference is present, against the reference itself - bit for bit."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import make_golden as G  # noqa: E402


@pytest.fixture(scope="module")
def orc():
    from oracle import oraclebind
    if not oraclebind.available():
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"])
    oraclebind.lib()
    return oraclebind


@pytest.mark.parametrize("name", list(G.cases()))
def test_oracle_matches_golden(orc, name):
    """the restatement reproduces every committed golden array exactly"""
    gold = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    m, out = G.record(name, B=orc)
    assert set(out) == set(gold.files)
    for k in gold.files:
        assert np.array_equal(np.asarray(out[k]), gold[k]), f"{name}:{k}"


@pytest.mark.parametrize("name", list(G.cases()))
def test_reference_regenerates_golden(ref, name):
    """the fixtures really are what the compiled reference produces today"""
    gold = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    m, out = G.record(name, B=ref)
    for k in gold.files:
        assert np.array_equal(np.asarray(out[k]), gold[k]), f"{name}:{k}"


@pytest.mark.parametrize("name", ["deck_5b_frame", "deck_5c_shell", "deck_5d_shell"])
def test_oracle_matches_deck_golden(orc, name):
    """the restatement reproduces what the reference computes on its own shipped decks"""
    from cubens_b200.model import model_from_dict
    from oracle.refbind import RefState
    gold = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    m = model_from_dict(gold)
    s = RefState(m)
    assert np.array_equal(orc.stiff(m, s, SLVFLAG=0, gen="c"), gold["K_sky"])
    assert np.array_equal(orc.mass(m, s), gold["mass"])
    if m.ANAFLAG == 2:
        s = RefState(m); s.begin_increment()
        for it in range(2):
            orc.update_forces(m, s, gold[f"dd_{it}"], dlpf=0.25, itecnt=it)
            assert np.array_equal(s.f_temp, gold[f"f_{it}"]) and np.array_equal(s.ef_i, gold[f"ef_{it}"])
            s.end_iteration()
            assert np.array_equal(orc.stiff(m, s, SLVFLAG=0), gold[f"K_sky_{it}"])


def test_deck_reader_matches_reference_setup(ref):
    """cubens_b200.deck + the numpy codes/skylin mirror reproduce NEQ / lss of the reference's own
    parser on every shipped deck (values from running ben.exe)"""
    base = "/root/reference/Sample_Input_Files/"
    if not os.path.isdir(base):
        pytest.skip("reference decks not present")
    from cubens_b200 import deck
    want = {"model_def_5a_truss.txt": (2, 3), "model_def_5b_frame.txt": (71, 729),
            "model_def_5c_shell.txt": (24, 300), "model_def_5d_shell.txt": (18, 151)}
    for f, (neq, lss) in want.items():
        m, _, _ = deck.read_deck(open(base + f).read())
        assert (m.NEQ, m.lss) == (neq, lss), f


def test_oracle_element_matrix_vs_reference(orc, ref):
    """element-level K (pre-scatter): give every element its own disjoint mcode in the
    reference and compare with orc_shell_element_K"""
    from cubens_b200 import meshgen
    m = meshgen.plate_model(3, 3, z_bump=0.04)
    s = ref.RefState(m); s.begin_increment()
    ref.update_forces(m, s, meshgen.perturbation(m)); s.end_iteration()
    import copy
    for n in range(m.NE_SH):
        m1 = copy.copy(m)
        mc = np.zeros_like(m.mcode); mc[n * 18:(n + 1) * 18] = np.arange(1, 19)
        m1.mcode = mc; m1.NEQ = 18; m1.SLVFLAG = 2
        dense = ref.stiff(m1, s, SLVFLAG=2).reshape(18, 18)
        K = orc.shell_element_K(m, s, n)
        assert np.array_equal(K, dense.T), n


def test_setup_mirror_bit_exact(ref, orc):
    """numpy codes/skylin/prop_sh mirror == reference == oracle"""
    import ctypes as C
    from cubens_b200 import meshgen
    m = meshgen.plate_model(6, 4, z_bump=0.02)
    l = ref.set_model(m)
    jf = np.where(m.jcode != 0, -1, 0).astype(np.int64)
    jc = jf.copy(); mc = np.zeros_like(m.mcode); wr = np.zeros(m.NJ * 3, dtype=np.int32)
    l.codes(ref.P(mc), ref.P(jc), ref.P(m.minc), ref.P(wr))
    assert l.ref_get_NEQ() == m.NEQ and np.array_equal(jc, m.jcode) and np.array_equal(mc, m.mcode)
    ojc, omc, oneq = orc.codes(m, jf)
    assert oneq == m.NEQ and np.array_equal(ojc, m.jcode) and np.array_equal(omc, m.mcode)
    maxa = np.zeros(m.NEQ + 1, dtype=np.int64); kht = np.zeros(m.NEQ, dtype=np.int64)
    pm = np.zeros(m.NEQ, dtype=np.int64); lss = C.c_long(0)
    l.skylin(ref.P(maxa), ref.P(mc), C.byref(lss), ref.P(jc), ref.P(kht), ref.P(pm))
    assert lss.value == m.lss and np.array_equal(maxa, m.maxa) and np.array_equal(kht, m.kht)
    okht, omaxa, olss = orc.skylin(m)
    assert olss == m.lss and np.array_equal(omaxa, m.maxa) and np.array_equal(okht, m.kht)
    txt = "".join("%.17g,%.17g,%.17g,%.17g,%.17g\n" % meshgen.SHELL_5C for _ in range(m.NE_SH)).encode()
    l.ref_set_input_text(txt, C.c_long(len(txt)))
    z = lambda a: np.zeros_like(a)
    emod, nu, xl, th, de, fa, sl, yl, c1, c2, c3 = map(z, (m.emod, m.nu, m.xlocal, m.thick, m.dens,
                                                          m.farea, m.slength, m.yld, m.c1, m.c2, m.c3))
    l.prop_sh(ref.P(m.x), ref.P(emod), ref.P(nu), ref.P(xl), ref.P(th), ref.P(de), ref.P(fa),
              ref.P(sl), ref.P(yl), ref.P(c1), ref.P(c2), ref.P(c3), ref.P(m.minc))
    l.ref_close_input()
    for a, b in ((xl, m.xlocal), (fa, m.farea), (sl, m.slength), (c1, m.c1), (c2, m.c2), (c3, m.c3),
                 (emod, m.emod), (nu, m.nu), (th, m.thick)):
        assert np.array_equal(a, b)


def test_dense_to_csc_rule(orc, ref):
    """orc_dense_to_csc == what solve.c:110-119 hands to umfpack_di_symbolic"""
    from cubens_b200 import meshgen
    m = meshgen.plate_model(3, 2, z_bump=0.02)
    s = ref.RefState(m)
    dense = ref.stiff(m, s, SLVFLAG=2, gen="c")
    rAp, rAi, rAx, _ = ref.dense_to_csc(m, dense.copy())
    oAp, oAi, oAx = orc.dense_to_csc(m.NEQ, dense)
    assert np.array_equal(rAp, oAp) and np.array_equal(rAi, oAi) and np.array_equal(rAx, oAx)


def test_reference_sample_deck_5a(ref, tmp_path):
    """the oracle build of the reference runs the shipped truss deck to the survey's numbers"""
    deck = "/root/reference/Sample_Input_Files/model_def_5a_truss.txt"
    exe = os.path.join(ROOT, "oracle", "_ref", "ben.exe")
    if not (os.path.exists(deck) and os.path.exists(exe)):
        pytest.skip("reference deck / ben.exe not present")
    (tmp_path / "model_def.txt").write_bytes(open(deck, "rb").read().replace(b"\r", b""))
    subprocess.check_call([exe], cwd=tmp_path, stdout=subprocess.DEVNULL)
    last = (tmp_path / "results2.txt").read_text().strip().splitlines()[-1].split()
    assert last[0] == "1.000000e+00" and last[2] == "-6.734350e-04" and last[3] == "2.267574e-07"

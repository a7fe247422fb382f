"""Element-partitioned material-nonlinear runs (last row): forces_fr / forces_sh return
at the FIRST element that trips (fact 0.8), so the ranks have to agree on the lowest global element
index before flags are committed and f_temp is gathered (cb_update_forces_begin -> min over ranks
-> cb_update_forces_end).  Several handles on ONE device stand in for the ranks; the exchange is
the host-side minimum (under torch.distributed it is TripExchange's all-reduce, covered on CPU by
tests/test_host_cpu.py).  The unpartitioned device run - itself pinned to the reference by
tests/test_plastic_gpu.py - is the answer: return codes, rescaled dlpf, yield flags, f_temp and the
owned matrix columns must be IDENTICAL (same kernels, same per-joint summation order)."""
import numpy as np
import pytest

import cubens_b200 as cb
from cubens_b200 import meshgen
from cubens_b200.partition import partition_model, reduce_trip

pytestmark = pytest.mark.gpu


class Parts:
    def __init__(self, m, world):
        self.m, self.world = m, world
        self.sub, self.asm, self.owned, self.ids, self.eqs = [], [], [], [], []
        jc = m.jcode.reshape(-1, 7)
        for r in range(world):
            s, own, ids = partition_model(m, world, r)
            a = cb.Assembler(s, layout=cb.CB_MAT_CSC)
            a.set_owned_joints(*own)
            a.set_element_ids(ids["fr"] if s.NE_FR else None, ids["sh"] if s.NE_SH else None)
            q = jc[own[0]:own[1]].reshape(-1)
            self.sub.append(s); self.asm.append(a); self.owned.append(own); self.ids.append(ids)
            self.eqs.append(q[q > 0] - 1)

    def each(self, name, *args):
        return [getattr(a, name)(*args) for a in self.asm]

    def update_forces(self, dd, dlpf=1.0, itecnt=0):
        firsts = [a.update_forces_begin(dd, dlpf, itecnt) for a in self.asm]
        g = reduce_trip(firsts)
        out = [a.update_forces_end(g[0], g[1], dlpf) for a in self.asm]
        f = np.zeros(self.m.NEQ)
        for (fl, _, _, _), q in zip(out, self.eqs):
            f[q] = fl[q]
        return f, max(o[1] for o in out), max(o[2] for o in out), min(o[3] for o in out), g

    def yldflag(self):
        y = np.full(2 * self.m.NE_FR, -1, dtype=np.int32)
        for a, ids in zip(self.asm, self.ids):
            loc = a.yldflag().reshape(-1, 2)
            y.reshape(-1, 2)[ids["fr"][ids["own_fr"]]] = loc[ids["own_fr"]]
        return y

    def columns(self, Ap_g, Ax_g):
        """owned CSC columns of every rank against the unpartitioned matrix (bit for bit)"""
        for a, q in zip(self.asm, self.eqs):
            Ap, Ai, Ax = a.csc()
            for c in q:
                assert np.array_equal(Ax[Ap[c]:Ap[c + 1]], Ax_g[Ap_g[c]:Ap_g[c + 1]]), c

    def close(self):
        self.each("close")


@pytest.mark.parametrize("world", [2, 3])
def test_frame_plastic_partitioned(gpu, world):
    """the walk of test_frame_plastic_lockstep: member ends reach the yield surface, overshoot it
    (code 1, dlpf rescaled) and unload (code 2) - the tripping member is on a different rank from
    call to call"""
    m = meshgen.lattice_model(3, ANAFLAG=3, load=200.0, SLVFLAG=2)
    base = np.random.default_rng(11).uniform(-1.0, 1.0, size=m.NEQ)
    steps = [0.004] * 3 + [-0.0003] * 3 + [0.002] * 2 + [-0.004] * 2
    one = cb.Assembler(m, layout=cb.CB_MAT_CSC)
    P = Parts(m, world)
    one.begin_increment(); P.each("begin_increment")
    codes, trips = [], set()
    k, scale, calls = 0, 1.0, 0
    while k < len(steps) and calls < 80:
        calls += 1
        dd = steps[k] * scale * base
        one.stiff(); P.each("stiff")
        Ap, Ai, Ax = one.csc()
        P.columns(Ap, Ax)
        f1, fr1, sh1, dl1 = one.update_forces(dd, dlpf=1.0, itecnt=0)
        f, fr, sh, dl, g = P.update_forces(dd, dlpf=1.0, itecnt=0)
        assert (fr, sh) == (fr1, sh1), calls
        assert dl == dl1, calls
        assert np.array_equal(P.yldflag(), one.yldflag()), calls
        assert np.array_equal(f, f1), calls
        codes.append(fr)
        if fr != 0:
            trips.add(next(r for r in range(world) if g[0] in P.ids[r]["fr"][P.ids[r]["own_fr"]]))
            if fr == 1:
                scale *= dl
            one.begin_increment(); P.each("begin_increment")
            continue
        one.end_iteration(); P.each("end_iteration")
        one.commit(); P.each("commit")
        one.begin_increment(); P.each("begin_increment")
        k += 1; scale = 1.0
    assert k == len(steps) and 1 in codes and 2 in codes and 0 in codes, codes
    assert (one.yldflag() == 1).any()
    one.close(); P.close()


@pytest.mark.parametrize("world", [2, 4])
def test_shell_plastic_partitioned(gpu, world):
    """the walk of test_shell_plastic_lockstep_csc: Ivanov yielding with forces_sh's return code 1
    raised by a shell of one rank and honoured by all of them"""
    m = meshgen.plate_model(6, 5, z_bump=0.05, ANAFLAG=3, SLVFLAG=2)
    base = np.random.default_rng(21).uniform(-1.0, 1.0, size=m.NEQ)
    one = cb.Assembler(m, layout=cb.CB_MAT_CSC)
    P = Parts(m, world)
    one.begin_increment(); P.each("begin_increment")
    scale, done, calls, tripped = 2e-4, 0, 0, 0
    while done < 12 and calls < 120:
        calls += 1
        one.stiff(); P.each("stiff")
        Ap, Ai, Ax = one.csc()
        P.columns(Ap, Ax)
        dd = scale * base
        f1, fr1, sh1, _ = one.update_forces(dd, itecnt=0)
        f, fr, sh, _, g = P.update_forces(dd, itecnt=0)
        assert (fr, sh) == (fr1, sh1), calls
        if sh:
            tripped += 1
            scale /= 2
            one.begin_increment(); P.each("begin_increment")
            continue
        assert np.array_equal(f, f1), calls
        one.end_iteration(); P.each("end_iteration"); one.commit(); P.each("commit")
        one.begin_increment(); P.each("begin_increment")
        done += 1; scale *= 1.5
    assert done == 12 and tripped > 0
    chi = one.download("CHI").reshape(-1, 3)
    assert (chi > 0).any()
    for a, ids in zip(P.asm, P.ids):
        assert np.array_equal(a.download("CHI").reshape(-1, 3), chi[ids["sh"]])
    one.close(); P.close()


def test_force_pass_halves_misuse(gpu):
    """error behaviour of the split force pass: _end without _begin, non-ascending element ids; and
    begin + end with negative firsts is exactly cb_update_forces_dev"""
    m = meshgen.lattice_model(3, ANAFLAG=3, load=200.0, SLVFLAG=2)
    a = cb.Assembler(m, layout=cb.CB_MAT_CSC)
    b = cb.Assembler(m, layout=cb.CB_MAT_CSC)
    with pytest.raises(cb.CubensError):
        a.update_forces_end(-1, -1)
    with pytest.raises(cb.CubensError):
        a.set_element_ids(fr_gid=np.arange(m.NE_FR)[::-1].copy())
    a.set_element_ids(fr_gid=np.arange(m.NE_FR) * 3 + 5)          # any ascending numbering
    dd = 0.004 * np.random.default_rng(11).uniform(-1.0, 1.0, size=m.NEQ)
    a.begin_increment(); b.begin_increment()
    for it in range(3):
        first = a.update_forces_begin(dd, 1.0, it)
        fa, fra, sha, dla = a.update_forces_end(-1, -1, 1.0)
        fb, frb, shb, dlb = b.update_forces(dd, dlpf=1.0, itecnt=it)
        assert (fra, sha, dla) == (frb, shb, dlb) and np.array_equal(fa, fb)
        assert np.array_equal(a.yldflag(), b.yldflag())
        if fra:
            assert first[0] % 3 == 2 and first[0] != 0x7fffffff      # reported in the caller's numbering
            break
        a.end_iteration(); b.end_iteration()
    a.close(); b.close()

"""CPU suite: the C-ABI library loads and exports every symbol of include/cubens_b200.h, fails
loudly without a GPU, and the host-side partition logic is consistent (gloo, world_size 2)."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_abi_exports_every_declared_symbol():
    import cubens_b200 as cb
    lib = cb.load_library()
    hdr = open(os.path.join(ROOT, "include", "cubens_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(cb_[a-z_A-Z0-9]+)\s*\(", hdr)))
    assert len(declared) >= 35
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert set(cb.EXPORTS) <= set(declared)
    assert lib.cb_abi_version() == int(re.search(r"#define CB_ABI_VERSION (\d+)", hdr).group(1)) == 5


def test_no_cpu_fallback():
    """without a CUDA device creation must fail loudly (never silently compute on the CPU)"""
    import cubens_b200 as cb
    from cubens_b200 import meshgen
    lib = cb.load_library()
    if lib.cb_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(cb.CubensError):
        cb.Assembler(meshgen.plate_model(2, 2))


def test_product_does_not_touch_oracle():
    """nothing under cu-bens_b200/ or include/ may reference oracle/"""
    bad = []
    for base in ("cu-bens_b200", "include"):
        for dp, _, fs in os.walk(os.path.join(ROOT, base)):
            if "build" in dp or "__pycache__" in dp:
                continue
            for f in fs:
                if f.endswith((".so", ".o", ".log")):
                    continue
                txt = open(os.path.join(dp, f), errors="ignore").read()
                if re.search(r"\boracle[/.]", txt) and "never touches" not in txt:
                    bad.append(os.path.join(dp, f))
    assert not bad, bad


def test_partition_numbering_consistent():
    from cubens_b200 import meshgen
    from cubens_b200.partition import plate_partition
    world, nx, ny = 3, 4, 5
    full = meshgen.plate_model(nx * world, ny, lx=nx * world / ny, ly=1.0, SLVFLAG=2)
    owned_total, el_total = 0, 0
    cover = np.zeros(full.NJ, dtype=int)
    for r in range(world):
        m, (j0, j1), ne = plate_partition(nx, ny, world, r)
        assert m.NEQ == full.NEQ and np.array_equal(m.jcode, full.jcode)
        assert np.abs(m.x - full.x).max() < 1e-15
        cover[j0:j1] += 1; el_total += ne
        # every element touching an owned joint is present locally (halo complete)
        tri_full = full.minc.reshape(-1, 3) - 1
        need = np.any((tri_full >= j0) & (tri_full < j1), axis=1)
        have = {tuple(t) for t in (m.minc.reshape(-1, 3) - 1).tolist()}
        assert all(tuple(t) in have for t in tri_full[need].tolist())
    assert np.all(cover == 1) and el_total == full.NE_SH


WORKER = r'''
import os, sys
sys.path.insert(0, os.path.join(%(root)r, "cu-bens_b200", "python"))
import numpy as np, torch, torch.distributed as dist
from cubens_b200.partition import plate_partition
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
m, (j0, j1), ne = plate_partition(4, 5, world, rank)
jc = m.jcode.reshape(-1, 7)
eq = jc[j0:j1].reshape(-1); eq = eq[eq > 0]
own = np.zeros(m.NEQ); own[eq - 1] = 1.0
t = torch.from_numpy(own); dist.all_reduce(t)
assert torch.all(t == 1.0), "every equation must be owned by exactly one rank"
n = torch.tensor([float(ne)]); dist.all_reduce(n)
assert int(n) == 2 * 4 * world * 5
# residual-sum exchange: partial dot products over owned equations add up to the global one
rng = np.random.default_rng(0); r = rng.normal(size=m.NEQ)
part = torch.tensor([float(np.dot(r[eq - 1], r[eq - 1]))], dtype=torch.float64); dist.all_reduce(part)
assert abs(float(part) - float(np.dot(r, r))) < 1e-9 * float(np.dot(r, r))
dist.destroy_process_group()
print("ok", rank)
'''


def test_partition_gloo_world2(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(WORKER % {"root": ROOT})
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                          "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port",
                          "29731", str(script)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("ok") == 2


def test_host_skyline_solver_matches_reference(ref):
    """cb_sky_factor / cb_sky_solve (C host, COLSOL) == the reference's skyfact / skysolve, bit for
    bit, on a real stiffness matrix"""
    import cubens_b200 as cb
    from cubens_b200 import meshgen
    for m in (meshgen.plate_model(7, 5, z_bump=0.05), meshgen.lattice_model(3), meshgen.truss_model(3)):
        s = ref.RefState(m)
        ss = ref.stiff(m, s, SLVFLAG=0, gen="c")
        rhs = np.random.default_rng(0).normal(size=m.NEQ)
        x_ref, _, _ = ref.skyline_solve(m, ss.copy(), rhs)
        x = cb.sky_factor_solve(m.maxa, ss.copy(), rhs)
        assert np.array_equal(x, x_ref)


def test_host_skyline_indefinite_mult_partition_match_reference(ref):
    """the other skyline services of solve.c the host drivers use, bit for bit against the
    reference's own routines: the indefinite factorisation of the arc-length driver (pivots ssd and
    sign of the determinant, skyfact with ALGFLAG 3, solve.c:563-572), skymult (solve.c:700-756,
    generalized-alpha Newmark) and matpart (solve.c:758-824, prescribed support motion)"""
    import ctypes as C
    import cubens_b200 as cb
    from cubens_b200 import meshgen
    hl = cb.load_host_library()
    P = ref.P
    m = meshgen.plate_model(6, 5, z_bump=0.05)
    s = ref.RefState(m)
    ss0 = ref.stiff(m, s, SLVFLAG=0, gen="c")
    n = m.NEQ
    maxa = np.ascontiguousarray(m.maxa, dtype=np.int64)
    rng = np.random.default_rng(3)
    # an indefinite matrix: shift the diagonal so that some pivots turn negative
    ss = ss0.copy()
    ss[maxa[:-1] - 1] -= 0.3 * np.abs(ss0[maxa[:-1] - 1]).mean()
    l = ref.set_model(m, SLVFLAG=0)
    l.ref_set_flags(C.c_int(2), C.c_int(3), C.c_int(0), C.c_int(1))          # ALGFLAG 3
    a = ss.copy(); ssd_ref = np.zeros(n); dd = np.zeros(n); det_ref = C.c_int(0)
    assert l.skyfact(P(maxa), P(a), P(ssd_ref), P(dd), C.c_int(0), C.byref(det_ref)) == 0
    b = ss.copy(); ssd = np.zeros(n); det = C.c_int(0)
    assert hl.cb_sky_factor(C.c_long(n), P(maxa), P(b), P(ssd), C.byref(det), C.c_int(1)) == 0
    assert np.array_equal(a, b) and np.array_equal(ssd, ssd_ref) and det.value == det_ref.value == 1
    # skymult
    l.ref_set_flags(C.c_int(2), C.c_int(1), C.c_int(0), C.c_int(1))
    v = rng.normal(size=n); v_ref = v.copy()
    l.skymult(P(maxa), P(ss0), P(v_ref))
    hl.cb_sky_mult(C.c_long(n), P(maxa), P(ss0), P(v))
    assert np.array_equal(v, v_ref)
    # matpart: three prescribed equations
    pm = np.zeros(n, dtype=bool); pm[[4, 17, n - 2]] = True
    ij32 = np.flatnonzero(pm).astype(np.int32); ii32 = np.flatnonzero(~pm).astype(np.int32)
    l.ref_set_NBC(C.c_long(3))
    K_ref = ss0.copy(); q_ref = rng.normal(size=n); uc = rng.normal(size=n); q = q_ref.copy(); K = ss0.copy()
    l.matpart(P(maxa), P(m.kht), P(K_ref), P(q_ref), P(uc), P(ii32), P(ij32))
    ii64, ij64 = ii32.astype(np.int64), ij32.astype(np.int64)      # keep alive across the call
    hl.cb_sky_partition(C.c_long(n), C.c_long(3), P(maxa), P(K), P(q), P(uc), P(ii64), P(ij64))
    assert np.array_equal(K, K_ref) and np.array_equal(q, q_ref)


def _sky_to_csc(maxa, ss, n):
    """full (unsymmetric-storage) CSC of a skyline matrix, every in-profile entry kept"""
    import scipy.sparse as sp
    rows, cols, vals = [], [], []
    for j in range(n):
        h = maxa[j + 1] - maxa[j]
        for k in range(h):
            i = j - k
            rows.append(i); cols.append(j); vals.append(ss[maxa[j] - 1 + k])
            if i != j:
                rows.append(j); cols.append(i); vals.append(ss[maxa[j] - 1 + k])
    A = sp.csc_matrix((vals, (rows, cols)), shape=(n, n))
    A.sort_indices()
    return A


def test_host_csc_solver_services_match_skyline(ref):
    """the CSC hand-off (host/cb_sparse.c): sparse LDL^T (definite and the
    indefinite arc-length branch with pivots / determinant sign), K v and matpart on the full CSC
    against the skyline services pinned to the reference above, on a real stiffness matrix"""
    import ctypes as C
    import cubens_b200 as cb
    from cubens_b200 import meshgen
    hl = cb.load_host_library()
    hl.cb_csc_solver_lnz.restype = C.c_long
    P = ref.P
    rng = np.random.default_rng(5)
    for m in (meshgen.plate_model(6, 5, z_bump=0.05), meshgen.lattice_model(3)):
        s = ref.RefState(m)
        ss0 = ref.stiff(m, s, SLVFLAG=0, gen="c")
        n = m.NEQ
        maxa = np.ascontiguousarray(m.maxa, dtype=np.int64)
        A = _sky_to_csc(maxa, ss0, n)
        Ap, Ai = A.indptr.astype(np.int32), A.indices.astype(np.int32)
        Ax = np.ascontiguousarray(A.data, dtype=np.float64)
        S = C.c_void_p()
        assert hl.cb_csc_solver_create(C.c_long(n), P(Ap), P(Ai), C.byref(S)) == 0
        # no fill beyond the skyline profile with the natural ordering
        assert hl.cb_csc_solver_lnz(S) <= int(maxa[-1] - 1) - n
        # positive definite solve
        rhs = rng.normal(size=n)
        x_sky = cb.sky_factor_solve(m.maxa, ss0.copy(), rhs)
        x = rhs.copy()
        assert hl.cb_csc_solver_factor(S, P(Ax), C.c_int(0), C.c_void_p(0), C.c_void_p(0)) == 0
        assert hl.cb_csc_solver_solve(S, P(x)) == 0
        assert np.abs(A @ x - rhs).max() <= 1e-9 * np.abs(rhs).max()
        assert np.abs(x - x_sky).max() <= 1e-8 * np.abs(x_sky).max()
        # indefinite: pivots and determinant sign as skyfact's ALGFLAG 3 branch
        shift = 0.3 * np.abs(ss0[maxa[:-1] - 1]).mean()
        ss = ss0.copy(); ss[maxa[:-1] - 1] -= shift
        dpos = np.array([A.indptr[j] + int(np.searchsorted(A.indices[A.indptr[j]:A.indptr[j + 1]], j))
                         for j in range(n)])
        Axs = Ax.copy(); Axs[dpos] -= shift
        ssd = np.zeros(n); det = C.c_int(0)
        assert hl.cb_sky_factor(C.c_long(n), P(maxa), P(ss), P(ssd), C.byref(det), C.c_int(1)) == 0
        piv = np.zeros(n); det2 = C.c_int(0)
        assert hl.cb_csc_solver_factor(S, P(Axs), C.c_int(1), C.byref(det2), P(piv)) == 0
        assert det.value == det2.value == 1
        assert np.array_equal(np.sign(piv), np.sign(ssd))
        assert np.abs(piv - ssd).max() <= 1e-7 * np.abs(ssd).max()
        # the definite factorisation rejects it, like skyfact
        assert hl.cb_csc_solver_factor(S, P(Axs), C.c_int(0), C.c_void_p(0), C.c_void_p(0)) == 1
        # K v
        v = rng.normal(size=n); v_sky = v.copy(); tmp = np.zeros(n)
        hl.cb_sky_mult(C.c_long(n), P(maxa), P(ss0), P(v_sky))
        hl.cb_csc_mult(C.c_long(n), P(Ap), P(Ai), P(Ax), P(v), P(tmp))
        assert np.abs(v - v_sky).max() <= 1e-13 * np.abs(v_sky).max()
        # diagonal addresses
        diag = np.zeros(n, dtype=np.int64)
        hl.cb_csc_diag(C.c_long(n), P(Ap), P(Ai), P(diag))
        assert np.array_equal(Ax[diag], ss0[maxa[:-1] - 1])
        # matpart
        pm = np.zeros(n, dtype=np.int32); pm[[4, 17, n - 2]] = 1
        ij = np.flatnonzero(pm).astype(np.int64); ii = np.flatnonzero(pm == 0).astype(np.int64)
        K = ss0.copy(); q = rng.normal(size=n); uc = rng.normal(size=n); q2 = q.copy(); Ax2 = Ax.copy()
        hl.cb_sky_partition(C.c_long(n), C.c_long(3), P(maxa), P(K), P(q), P(uc), P(ii), P(ij))
        hl.cb_csc_partition(C.c_long(n), P(Ap), P(Ai), P(Ax2), P(q2), P(uc), P(pm))
        B = _sky_to_csc(maxa, K, n)
        assert np.array_equal(B.indices, A.indices) and np.array_equal(B.data, Ax2)
        assert np.abs(q2 - q).max() <= 1e-13 * np.abs(q).max()
        hl.cb_csc_solver_destroy(S)


def test_partition_model_generic():
    """partition_model (frames / shells / trusses by joint ranges): global numbering kept, halo
    complete, every element owned once, per-element arrays cut consistently with the type offsets"""
    from cubens_b200 import meshgen
    from cubens_b200.partition import partition_model
    for m, key, k in ((meshgen.lattice_model(3, ANAFLAG=3), "fr", 14), (meshgen.plate_model(6, 5, z_bump=0.05), "sh", 18),
                      (meshgen.truss_model(3), "tr", 6)):
        ne = {"fr": m.NE_FR, "sh": m.NE_SH, "tr": m.NE_TR}[key]
        for world in (2, 3):
            own_total, cover = 0, np.zeros(m.NJ, dtype=int)
            for r in range(world):
                s, (j0, j1), ids = partition_model(m, world, r)
                cover[j0:j1] += 1
                assert s.NEQ == m.NEQ and s.jcode is m.jcode and s.x is m.x
                gid = ids[key]
                assert np.all(np.diff(gid) > 0)
                own_total += int(ids["own_" + key].sum())
                assert np.array_equal(s.mcode, m.mcode.reshape(-1, k)[gid].reshape(-1))
                nn = 3 if key == "sh" else 2
                conn = m.minc.reshape(-1, nn)
                need = np.flatnonzero(np.any((conn - 1 >= j0) & (conn - 1 < j1), axis=1))
                assert np.array_equal(need, gid)
                assert np.array_equal(s.minc.reshape(-1, nn), conn[gid])
                off = {"tr": 0, "fr": m.NE_TR, "sh": m.NE_TR + m.NE_FR}[key]
                assert np.array_equal(s.emod, m.emod[off + gid])
                if key == "fr":
                    assert np.array_equal(s.auxpt.reshape(-1, 3), m.auxpt.reshape(-1, 3)[gid])
                    assert np.array_equal(s.c2.reshape(-1, 3), m.c2[m.NE_TR:].reshape(-1, 3)[gid])
                    assert np.array_equal(s.llength, m.llength[m.NE_TR + gid])
                if key == "sh":
                    assert np.array_equal(s.xlocal.reshape(-1, 3), m.xlocal.reshape(-1, 3)[gid])
                    assert np.array_equal(s.c3.reshape(-1, 3), m.c3.reshape(-1, 3)[gid])
            assert np.all(cover == 1) and own_total == ne


TRIP_WORKER = r'''
import os, sys
sys.path.insert(0, os.path.join(%(root)r, "cu-bens_b200", "python"))
import torch.distributed as dist
from cubens_b200.partition import TripExchange
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
BIG = 0x7fffffff


class FakeAsm:
    """stands in for the device handle: rank 1 holds tripping member 17 (code 1, factor 0.25) and
    member 40, rank 0 holds member 23; shells trip on rank 0 only"""
    def update_forces_begin(self, dd, dlpf, itecnt):
        return (23, 5) if rank == 0 else (17, BIG)
    def update_forces_end(self, ffr, fsh, dlpf, want_f):
        assert (ffr, fsh) == (17, 5), (ffr, fsh)
        holds = rank == 1
        return None, (1 if holds else 0), 1, (dlpf * 0.25 if holds else dlpf)


f, fr, sh, dl = TripExchange(FakeAsm(), dist).update_forces(None, dlpf=0.5, want_f=False)
assert (fr, sh, dl) == (1, 1, 0.125), (fr, sh, dl)
dist.destroy_process_group()
print("ok", rank)
'''


def test_trip_exchange_gloo_world2(tmp_path):
    """the ANAFLAG 3 exchange of cb_update_forces_begin / _end under torch.distributed (gloo, 2 ranks)"""
    script = tmp_path / "t.py"
    script.write_text(TRIP_WORKER % {"root": ROOT})
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                          "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port",
                          "29733", str(script)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("ok") == 2


def test_fp64_probe_needs_a_device():
    """cb_measure_fp64_tflops (the FP64 roofline denominator of bench.py) reports failure, not a made-up
    number, where no CUDA device exists"""
    import cubens_b200 as cb
    lib = cb.load_library()
    if lib.cb_device_count() == 0:
        assert lib.cb_measure_fp64_tflops(0) < 0


@pytest.mark.parametrize("nx,ny,jitter,own", [(1, 1, 0.0, None), (2, 2, 0.0, None), (9, 7, 0.0, None),
                                               (40, 25, 0.0, None), (40, 25, 0.2, None), (60, 40, 0.0, (300, 1500)),
                                               (120, 3, 0.0, None), (3, 120, 0.0, (100, 101))])
@pytest.mark.parametrize("shape", ["narrow", "wide", "mini"])
def test_stream_plan_interpreter(nx, ny, jitter, own, shape, monkeypatch):
    """the sorted element-to-nonzero map of the shell stream kernel (k_assemble_shell_stream), built on
    the host without a device and walked the way the kernel walks it: every joint-pair block is summed
    from exactly its contribution list in reference order and stored once, the blocks tile each tile's
    output without overlap or gap, the tiles cover the owned CSC slice contiguously"""
    import cubens_b200 as cb
    from cubens_b200 import meshgen
    monkeypatch.setenv("CB_KT", shape)
    m = meshgen.plate_model(nx, ny, SLVFLAG=2, jitter=jitter, pinned=nx > 1)
    j0, j1 = own if own else (0, 0)
    st = cb.plan_selfcheck(m, j0, j1)
    assert st["kind"] == 3 and st["tiles"] >= 1 and st["steps"] == {"narrow": 4, "wide": 6, "mini": 3}[shape]
    if own is None:
        # nnz of the structural joint-block pattern: sum over joints of nfree(B) * sum of nfree(neighbours)
        jc = (np.asarray(m.jcode).reshape(-1, 7) != 0).sum(axis=1)
        nb = [set() for _ in range(m.NJ)]
        for tri in np.asarray(m.minc).reshape(-1, 3) - 1:
            for a in tri:
                nb[a].update(int(b) for b in tri)
        assert st["nnz"] == sum(int(jc[b]) * sum(int(jc[a]) for a in nb[b]) for b in range(m.NJ))


@pytest.mark.parametrize("shape", ["narrow", "wide"])
def test_stream_plan_split_blocks(shape, monkeypatch):
    """union-jack plate: joints with 8 shells around them - their diagonal blocks (8 contributions) are cut
    in two parts, the second a follower that is added at the end of the tile"""
    import cubens_b200 as cb
    from cubens_b200 import meshgen
    monkeypatch.setenv("CB_KT", shape)
    st = cb.plan_selfcheck(meshgen.plate_model(24, 17, SLVFLAG=2, unionjack=True, z_bump=0.02))
    assert st["kind"] == 3


def test_plan_selfcheck_other_element_types():
    """frames / trusses / bricks take the general tile plan; the host-only build must go through"""
    import cubens_b200 as cb
    from cubens_b200 import meshgen
    assert cb.plan_selfcheck(meshgen.lattice_model(5, SLVFLAG=2))["kind"] == 1
    assert cb.plan_selfcheck(meshgen.truss_model(4, SLVFLAG=2))["kind"] == 1
    assert cb.plan_selfcheck(meshgen.brick_model(3, 3, 3, skin=True))["kind"] == 1


@pytest.mark.parametrize("kind", ["plate", "jitter", "partition_lo", "partition_mid", "partition_hi", "unionjack",
                                  "lattice", "brick_skin"])
def test_symmetric_handoff_host_side(kind):
    """packed upper-triangle layout (what k_pack_upper produces), its CSC pattern and the threaded rebuild of
    the full matrix, on the host with synthetic symmetric values (cb_sym_selftest; no device)"""
    import cubens_b200 as cb
    from cubens_b200 import meshgen
    own = (0, 0)
    if kind == "lattice":
        m = meshgen.lattice_model(6, SLVFLAG=2)
    elif kind == "brick_skin":
        m = meshgen.brick_model(3, 4, 3, skin=True)
    else:
        m = meshgen.plate_model(60, 40, SLVFLAG=2, jitter=0.2 if kind == "jitter" else 0.0,
                                unionjack=kind == "unionjack")
        own = {"partition_lo": (0, 900), "partition_mid": (300, 1500), "partition_hi": (1500, m.NJ)}.get(kind, (0, 0))
    assert cb.sym_selftest(m, own[0], own[1], nthreads=3) >= 0.0


@pytest.mark.parametrize("skin", [False, True])
def test_fsi_plan_selfcheck(skin):
    """ANAFLAG 4: pressure twins + the coupling element go through the same plan builder / interpreter"""
    import cubens_b200 as cb
    from cubens_b200 import meshgen
    m = meshgen.fsi_model(5, 4, 2, 3, skin=skin, distort=0.1)
    st = cb.plan_selfcheck(m)
    assert st["nnz"] > 0 and st["tiles"] > 0
    # block pattern: structure block columns hold the wet joint's pressure row and vice versa
    assert st["nnz"] < 90 * m.NEQ

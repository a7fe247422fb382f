"""The -DCB_HAVE_UMFPACK branch of cu-bens_b200/host/cb_sparse.c (solve.c:107-135: umfpack_di_symbolic once,
numeric + solve per refactorisation, the solution copied back into the right-hand side - which the reference
forgets) compiled, linked and run.  SuiteSparse is not in this image: the five umfpack_di_*
entry points come from a test double with the published prototypes (tests/umfpack_double: dense LU), so this
pins the binding's call sequence and argument order, not UMFPACK itself."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    out = tmp_path_factory.mktemp("umf") / "libcb_sparse_umf.so"
    cmd = ["gcc", "-O2", "-fPIC", "-shared", "-Wall", "-Wextra", "-Werror", "-DCB_HAVE_UMFPACK",
           "-I", os.path.join(ROOT, "tests", "umfpack_double"), "-I", os.path.join(ROOT, "cu-bens_b200", "host"),
           *[os.path.join(ROOT, "cu-bens_b200", "host", f) for f in
             ("cb_sparse.c", "cb_skyline.c", "cb_newton.c", "cb_newmark.c", "cb_arclength.c")],
           os.path.join(ROOT, "tests", "umfpack_double", "umfpack_double.c"), "-o", str(out),
           "-L", os.path.join(ROOT, "cu-bens_b200"), "-lcubens_b200", "-Wl,-rpath," + os.path.join(ROOT, "cu-bens_b200"), "-lm"]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    return C.CDLL(str(out))


def _spd_csc(n, rng):
    A = rng.uniform(-1, 1, (n, n)) * (rng.uniform(size=(n, n)) < 0.2)
    A = A + A.T + n * np.eye(n)
    Ap = [0]; Ai = []; Ax = []
    for j in range(n):
        rows = np.nonzero(A[:, j])[0]
        Ai += rows.tolist(); Ax += A[rows, j].tolist(); Ap.append(len(Ai))
    return A, np.array(Ap, dtype=np.int32), np.array(Ai, dtype=np.int32), np.array(Ax)


def test_umfpack_branch_solves_through_the_binding(lib):
    rng = np.random.default_rng(3)
    n = 40
    A, Ap, Ai, Ax = _spd_csc(n, rng)
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    s = C.c_void_p()
    assert lib.cb_csc_solver_create(C.c_long(n), P(Ap), P(Ai), C.byref(s)) == 0
    calls = (C.c_int * 5).in_dll(lib, "umfpack_double_calls")
    assert list(calls) == [1, 0, 0, 0, 0]                      # symbolic once (solve.c:122)
    for k in range(2):                                          # two refactorisations on one symbolic analysis
        Axk = Ax * (1.0 + k)
        assert lib.cb_csc_solver_factor(s, P(Axk), 0, None, None) == 0
        b = rng.uniform(-1, 1, n)
        x = b.copy()
        assert lib.cb_csc_solver_solve(s, P(x)) == 0
        assert np.allclose(x, np.linalg.solve(A * (1.0 + k), b), rtol=1e-12, atol=1e-14)   # x copied back
    assert list(calls)[:3] == [1, 2, 2] and calls[4] == 1       # the first Numeric object freed at the refactorisation
    # the arc-length branch needs pivots and the determinant's sign: always the built-in LDL^T (solve.c:563-572)
    piv = np.zeros(n); neg = C.c_int(-1)
    assert lib.cb_csc_solver_factor(s, P(Ax), 1, C.byref(neg), P(piv)) == 0
    assert neg.value == 0 and np.all(piv > 0) and calls[1] == 2
    lib.cb_csc_solver_destroy(s)
    assert calls[3] == 1 and calls[4] == 2

"""The sorted element-to-nonzero map, the CSC pattern and the tile plan built ON THE DEVICE (cb_plan_device.cuh:
radix sorts, run-length encoding, scans, the packer of cb_plan_pack.h one thread per segment) against the host
builder of cb_api.cu: identical bytes - tiles, step records, pair records, shell slots, Ap, Ai - and therefore
identical matrices and forces (north_star: "CSC pattern precomputed once on the device", "the assembled CSC
sparsity pattern and element-to-DOF maps must be bit-exact with the reference" - the host builder is pinned to
the reference's codes() / dense scan by tests/test_shell_gpu.py and tests/test_host_cpu.py)."""
import numpy as np
import pytest

import cubens_b200 as cb
from cubens_b200 import meshgen

pytestmark = pytest.mark.gpu

CASES = {
    "plate": (dict(nx=40, ny=27, z_bump=0.02), None),
    "single_cell": (dict(nx=1, ny=1, pinned=False), None),
    "jitter": (dict(nx=33, ny=21, jitter=0.2), None),
    "unionjack": (dict(nx=24, ny=17, unionjack=True, z_bump=0.02), None),
    "partition": (dict(nx=60, ny=40), (300, 1500)),
    "midsize": (dict(nx=200, ny=60), None),
}


def _run(m, own, mode, shape, monkeypatch):
    monkeypatch.setenv("CB_PLAN", mode)
    monkeypatch.setenv("CB_KT", shape)
    a = cb.Assembler(m, layout=cb.CB_MAT_CSC)
    if own:
        a.set_owned_joints(*own)
    a.begin_increment()
    dd = meshgen.perturbation(m, scale=1e-3)
    f, *_ = a.update_forces(dd); a.end_iteration(); a.stiff()
    out = dict(info=a.plan_info(), plan=[a.debug_stream_plan(k) for k in range(6)], csc=a.csc(), f=f,
               ax_upper=a.csc_upper()[2])
    a.close()
    return out


@pytest.mark.parametrize("shape", ["narrow", "wide"])
@pytest.mark.parametrize("case", list(CASES))
def test_device_plan_equals_host_plan(gpu, case, shape, monkeypatch):
    kw, own = CASES[case]
    kw = dict(kw); nx, ny = kw.pop("nx"), kw.pop("ny")
    m = meshgen.plate_model(nx, ny, SLVFLAG=2, **kw)
    host = _run(m, own, "host", shape, monkeypatch)
    dev = _run(m, own, "device", shape, monkeypatch)
    assert host["info"][1] is False and dev["info"][1] is True
    for k, name in enumerate(("tiles", "step records", "pair records", "shell slots", "Ai", "Ap")):
        assert host["plan"][k] is not None and dev["plan"][k] is not None, name
        assert np.array_equal(host["plan"][k], dev["plan"][k]), f"{name} differ between the host and the device builder"
    for x, y in zip(host["csc"], dev["csc"]):
        assert np.array_equal(x, y)
    assert np.array_equal(host["f"], dev["f"]) and np.array_equal(host["ax_upper"], dev["ax_upper"])

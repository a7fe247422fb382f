"""A C host (no Python, no torch, no MPI in the processes under test) running the element-partitioned hot path
with the collective INSIDE the library: cu-bens_b200/host/cb_multi_gpu_demo.c, one process per GPU, the
ncclUniqueId handed over through a file, cb_comm_init + cb_residual_allreduce (ncclAllReduce on the handle's
stream).  The all-reduced convergence sums and reaction resultants must equal a one-GPU run of the whole
model (north_star: "interface residual and reaction sums are exchanged by NCCL over NVLink")."""
import os
import subprocess
import numpy as np
import pytest

import cubens_b200 as cb
from cubens_b200 import meshgen
from cubens_b200.partition import partition_model, write_submodel

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEMO = os.path.join(ROOT, "cu-bens_b200", "cb_multi_gpu_demo")


def _whole(m, dd, lpf):
    a = cb.Assembler(m, layout=cb.CB_MAT_CSC)
    a.begin_increment(); a.stiff(); a.update_forces(dd); a.set_q(m.q)
    s = a.residual_sums(lpf, fetch=True); r = a.reaction_sums()
    a.close()
    return np.concatenate([s, r])


@pytest.mark.parametrize("collective", ["auto", "nccl"])
@pytest.mark.parametrize("world", [1, 2])
def test_c_host_ranks_allreduce_in_library(gpu, tmp_path, world, collective):
    """collective "auto": the library's own exchange over mapped peer memory where cb_comm_init could map the
    ranks' mailboxes (fused into the sums kernel by cb_residual_sums_allreduce; the demo checks that the fused
    launch and the two separate calls give the same bits); "nccl": CB_COMM_P2P=0 forces ncclAllReduce"""
    if gpu.cb_device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    assert os.path.exists(DEMO), "make -C cu-bens_b200 builds cb_multi_gpu_demo"
    m = meshgen.plate_model(48, 31, z_bump=0.02, SLVFLAG=2)
    dd = meshgen.perturbation(m, scale=1e-3)
    lpf = 0.6
    want = _whole(m, dd, lpf)
    procs, outs = [], []
    uid = tmp_path / "nccl_uid"
    for r in range(world):
        sub, own, _ = partition_model(m, world, r)
        mp = tmp_path / f"model_{r}.bin"; op = tmp_path / f"out_{r}.bin"
        write_submodel(mp, sub, own, m.q, dd, lpf)
        outs.append(op)
        env = dict(os.environ, CB_COMM_P2P="0") if collective == "nccl" else dict(os.environ)
        procs.append(subprocess.Popen([DEMO, str(mp), str(r), str(world), str(uid), str(op)], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    for p in procs:
        out, _ = p.communicate(timeout=300)
        assert p.returncode == 0, out
        print(out.strip().splitlines()[-1])
    for op in outs:
        got = np.fromfile(op, dtype=np.float64)
        assert got.shape == (12,)
        if collective == "nccl" or world == 1:
            assert got[11] == 0.0
        got = got[:11]
        # the ranks add their partial sums in a different association than the one-GPU reduction
        assert np.allclose(got[:5], want[:5], rtol=1e-12, atol=0)
        assert np.allclose(got[5:], want[5:], rtol=1e-9, atol=1e-9 * np.abs(want[5:]).max())
    assert np.array_equal(np.fromfile(outs[0], dtype=np.float64), np.fromfile(outs[-1], dtype=np.float64))
    if collective == "auto" and world > 1 and os.environ.get("CB_EXPECT_PEER_MEMORY"):
        assert np.fromfile(outs[0], dtype=np.float64)[11] == 1.0, "the ranks did not map each other's mailboxes"

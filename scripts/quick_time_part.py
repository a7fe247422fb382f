"""ad-hoc: one rank's share of a 2-rank weak-scaling plate, timed alone on one GPU (no NCCL) - where
does the partitioned step differ from the unpartitioned one? (development aid)"""
import sys, time, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "cu-bens_b200", "python"))
import numpy as np
import cubens_b200 as cb
from cubens_b200 import meshgen
from cubens_b200.partition import plate_partition
n = 1000
rank = int(sys.argv[1]) if len(sys.argv) > 1 else 0
world = int(sys.argv[2]) if len(sys.argv) > 2 else 2
m, owned, n_local = plate_partition(n, n, world, rank, weak=True)
a = cb.Assembler(m, layout=cb.CB_MAT_CSC)
a.set_owned_joints(*owned)
dd = meshgen.perturbation(m)
a.begin_increment(); a.update_forces(dd, want_f=False); a.end_iteration()
a.set_dd(dd * 1e-3)
a.set_q(m.q)
def step(ex):
    a.stiff(); a.update_forces_dev()
    if ex: a.residual_sums(1.0)
    a.end_iteration()
for ex in (False, True):
    for _ in range(5): step(ex)
    a.sync(); a.timer_start(); t0 = time.perf_counter()
    for _ in range(50): step(ex)
    ms = a.timer_stop_ms() / 50; a.sync(); w = (time.perf_counter() - t0) * 1e3 / 50
    a.stiff(); ks = a.last_stiff_ms; a.update_forces_dev(); fs = a.last_forces_ms; a.end_iteration()
    print(f"classes {a.geometry_classes} map_bytes {a.map_bytes} nnz {a.lib.cb_csc_nnz(a.h)} | rank {rank}/{world} sums={ex}: step {ms:.3f} ms (wall {w:.3f})  stiff {ks:.3f}  forces {fs:.3f}  elements {m.NE_SH} NEQ {m.NEQ}", flush=True)

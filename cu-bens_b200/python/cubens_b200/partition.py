"""Element partitioning across the GPUs of one box.

One process per GPU.  Every rank keeps the GLOBAL joint / equation numbering (NJ, NEQ, jcode,
coordinates - a few tens of MB) and builds a sub-model holding only

    its own contiguous element range  +  the halo elements that touch its owned joints,

so the columns of the global matrix that belong to its owned joints, and the f_int / mass
entries of those joints, are complete on that rank with no exchange of matrix or force
entries (halo elements are re-evaluated on both neighbours: 2*ny of ~2*nx*ny/N elements).
The only per-iteration cross-GPU traffic is the all-reduce of the residual / displacement /
energy sums that the convergence test needs (test(), misc.c:187-250), done with NCCL over
NVLink on device-resident partial sums.
"""
from __future__ import annotations

import numpy as np

from .model import build_model, I64, F64
from . import meshgen


def plate_partition(nx_local, ny, world, rank, weak=True, props=meshgen.SHELL_5C):
    """Row-strip partition of the pinned plate of meshgen.plate_model.

    weak=True : the global plate has nx_local*world x ny cells (unit cells of 1/ny), each rank
                owns nx_local cell rows.   weak=False: the global plate is nx_local x ny and the
                rows are split across ranks.
    Returns (sub_model, (j0, j1) owned joints 0-based, number of OWN elements)."""
    nxg = nx_local * world if weak else nx_local
    per = nxg // world
    i0 = rank * per
    i1 = nxg if rank == world - 1 else i0 + per
    h = 1.0 / ny
    # global joints
    i, j = np.meshgrid(np.arange(nxg + 1), np.arange(ny + 1), indexing="ij")
    x = np.stack([i * h, j * h, np.zeros_like(i, dtype=F64)], axis=-1).astype(F64).reshape(-1)
    NJ = (nxg + 1) * (ny + 1)
    jflags = np.full((NJ, 7), -1, dtype=I64)
    jflags[:, 6] = 0
    edge = ((i == 0) | (i == nxg) | (j == 0) | (j == ny)).reshape(-1)
    jflags[edge, 0:3] = 0
    # local + halo cell rows: joint row i0 is owned by this rank and touched by cell row i0-1
    c0 = max(i0 - 1, 0)
    c1 = i1
    ii, jj = np.meshgrid(np.arange(c0, c1), np.arange(ny), indexing="ij")
    nid = lambda a, b: (a * (ny + 1) + b + 1).astype(I64)
    A, B, Cn, D = nid(ii, jj), nid(ii + 1, jj), nid(ii + 1, jj + 1), nid(ii, jj + 1)
    shells = np.stack([np.stack([A, B, Cn], -1), np.stack([A, Cn, D], -1)], axis=2).reshape(-1, 3)
    centre = int((nxg // 2) * (ny + 1) + ny // 2 + 1)
    m = build_model(x, shells=shells, shell_props=props, loads=[(centre, 3, -1000.0)],
                    ANAFLAG=2, ALGFLAG=1, SLVFLAG=2, jflags=jflags,
                    meta=dict(kind="plate-part", nxg=nxg, ny=ny, rows=(i0, i1), halo_rows=(c0, c1)))
    # owned joints: joint rows [i0, i1), the last rank also owns row nxg
    j0 = i0 * (ny + 1)
    j1 = (i1 + (1 if rank == world - 1 else 0)) * (ny + 1)
    return m, (j0, j1), 2 * ny * (i1 - i0)


class _DevVec:
    """torch view of a device buffer owned by the C library (no copy)"""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f8", "data": (int(ptr), False),
                                         "version": 2}


class InterfaceExchange:
    """Per-iteration NCCL step: the three convergence sums of test() (misc.c:187-250) are formed
    over the owned equations by the library (cb_residual_sums, one fused pass, fixed order) and
    all-reduced over the ranks; 24 bytes cross NVLink per iteration."""

    def __init__(self, asm, m, world, rank, dist, owned=None):
        import torch
        self.torch, self.dist, self.asm = torch, dist, asm
        asm.set_q(m.q)
        self.sums = torch.as_tensor(_DevVec(asm.lib.cb_dev_sums(asm.h), 3), device="cuda")
        self.stream = torch.cuda.ExternalStream(asm.lib.cb_stream(asm.h))

    def reduce(self, lpf=1.0):
        self.asm.residual_sums(lpf)
        with self.torch.cuda.stream(self.stream):
            self.dist.all_reduce(self.sums)
        return self.sums

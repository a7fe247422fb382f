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


def partition_model(m, world, rank):
    """Generic element partition of a truss / frame / shell model (no bricks) by contiguous joint
    ranges: rank r owns joints [j0, j1) and gets every element that touches one of them (its own
    elements plus the halo), with the GLOBAL joint / equation numbering.  Returns
    (sub_model, (j0, j1), ids) with ids = dict(tr=, fr=, sh=) the global 0-based element indices of
    the local elements (ascending; what cb_set_element_ids takes) and ids['own_*'] boolean masks of
    the elements counted on this rank (first joint owned)."""
    from .model import Model
    if m.NE_BR:
        raise ValueError("partition_model: bricks are assembled once and are not partitioned")
    per = -(-m.NJ // world)
    j0, j1 = min(rank * per, m.NJ), min((rank + 1) * per, m.NJ)
    TR, FR, SH = m.NE_TR, m.NE_FR, m.NE_SH
    o_fr, o_sh = 2 * TR, 2 * TR + 2 * FR
    conn = {"tr": m.minc[:o_fr].reshape(-1, 2), "fr": m.minc[o_fr:o_sh].reshape(-1, 2),
            "sh": m.minc[o_sh:o_sh + 3 * SH].reshape(-1, 3)}
    ids = {}
    for k, c in conn.items():
        owned = (c - 1 >= j0) & (c - 1 < j1)
        ids[k] = np.flatnonzero(owned.any(axis=1)).astype(I64)
        ids["own_" + k] = owned[ids[k], 0] if len(c) else np.zeros(0, dtype=bool)
    tr, fr, sh = ids["tr"], ids["fr"], ids["sh"]
    ntr, nfr, nsh = len(tr), len(fr), len(sh)
    s = Model(NJ=m.NJ, NE_TR=ntr, NE_FR=nfr, NE_SH=nsh, NEQ=m.NEQ, ANAFLAG=m.ANAFLAG, ALGFLAG=m.ALGFLAG,
              SLVFLAG=2, x=m.x, jcode=m.jcode, q=m.q, lss=0,
              meta=dict(m.meta, kind="partition", rank=rank, world=world, joints=(j0, j1)))
    s.minc = np.concatenate([conn["tr"][tr].reshape(-1), conn["fr"][fr].reshape(-1), conn["sh"][sh].reshape(-1)])
    jc = m.jcode.reshape(-1, 7)
    s.mcode = np.concatenate([jc[conn["tr"][tr] - 1][:, :, :3].reshape(-1),      # model.c:992-1142
                              jc[conn["fr"][fr] - 1].reshape(-1),
                              jc[conn["sh"][sh] - 1][:, :, :6].reshape(-1)]).astype(I64)
    by_el = np.concatenate([tr, TR + fr, TR + FR + sh])                  # emod / yld
    s.emod, s.yld = m.emod[by_el], m.yld[by_el]
    # dens[n] is indexed by the element's number WITHIN its type for every type (App. B.4)
    # - the same prefix of one array for trusses, frames and shells.  A mixed-type partition therefore
    # only has a faithful dens[] when the types agree on those prefixes (uniform density, or one type)
    s.dens = np.zeros(ntr + nfr + nsh)
    for idx in (tr, fr, sh):
        want = m.dens[idx]
        have = s.dens[:len(idx)]
        clash = (have != 0) & (have != want)
        if clash.any():
            raise ValueError("partition_model: dens[n] is shared by all element types (reference App. B.4); "
                             "a mixed-type partition with non-uniform densities cannot be represented")
        s.dens[:len(idx)] = want
    lin = np.concatenate([tr, TR + fr])
    s.carea, s.llength = m.carea[lin], m.llength[lin]
    three = lambda idx: (idx[:, None] * 3 + np.arange(3)).reshape(-1)
    cs = np.concatenate([tr, TR + three(fr), TR + 3 * FR + three(sh)]).astype(I64)
    s.c1, s.c2, s.c3 = m.c1[cs], m.c2[cs], m.c3[cs]
    s.nu, s.thick, s.farea = m.nu[sh], m.thick[sh], m.farea[sh]
    s.slength, s.xlocal = m.slength[three(sh)], m.xlocal[three(sh)]
    rows = lambda a, idx, k: np.ascontiguousarray(np.asarray(a).reshape(-1, k)[idx].reshape(-1))
    for name, k in (("gmod", 1), ("istrong", 1), ("iweak", 1), ("ipolar", 1), ("iwarp", 1), ("zstrong", 1),
                    ("zweak", 1), ("auxpt", 3), ("offset", 6), ("osflag", 1), ("mendrel", 5), ("xfr", 6),
                    ("efFE_ref", 14)):
        a = getattr(m, name)
        setattr(s, name, rows(a, fr, k) if a is not None else None)
    return s, (j0, j1), ids


def reduce_trip(firsts):
    """the exchange between cb_update_forces_begin and _end: element-wise minimum of the ranks'
    (first_fr, first_sh) - host-side form for several handles in one process; under
    torch.distributed the same thing is one all_reduce(MIN) of two int32 (TripExchange)"""
    return min(f[0] for f in firsts), min(f[1] for f in firsts)


class TripExchange:
    """ANAFLAG 3 on several GPUs: agree on the first tripping frame / shell (all-reduce MIN of two
    integers), then on the return code (MAX) and the rescaled dlpf (MIN) - 24 bytes per force pass."""

    def __init__(self, asm, dist, device=None):
        import torch
        self.torch, self.dist, self.asm = torch, dist, asm
        self.dev = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")

    def update_forces(self, dd, dlpf=1.0, itecnt=0, want_f=True):
        t, dist = self.torch, self.dist
        first = t.tensor(self.asm.update_forces_begin(dd, dlpf, itecnt), dtype=t.int32, device=self.dev)
        dist.all_reduce(first, op=dist.ReduceOp.MIN)
        f, fr, sh, dl = self.asm.update_forces_end(int(first[0]), int(first[1]), dlpf, want_f)
        code = t.tensor([fr, sh], dtype=t.int32, device=self.dev)
        dist.all_reduce(code, op=dist.ReduceOp.MAX)
        dmin = t.tensor([dl], dtype=t.float64, device=self.dev)
        dist.all_reduce(dmin, op=dist.ReduceOp.MIN)
        return f, int(code[0]), int(code[1]), float(dmin[0])


class _DevVec:
    """torch view of a device buffer owned by the C library (no copy)"""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f8", "data": (int(ptr), False),
                                         "version": 2}


class InterfaceExchange:
    """Per-iteration NCCL step: the three convergence sums of test() (misc.c:187-250) and the six reaction
    resultants are formed over the owned equations / joints by the library (cb_residual_sums, one fused
    pass, fixed order) and all-reduced over the ranks BY THE LIBRARY (cb_residual_allreduce: ncclAllReduce
    on the handle's stream, include/cubens_b200.h); 72 bytes cross NVLink per iteration.  torch.distributed
    only carries the 128-byte ncclUniqueId to the ranks once."""

    def __init__(self, asm, m, world, rank, dist, owned=None):
        import torch
        self.asm = asm
        asm.set_q(m.q)
        if world > 1:
            uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
            if rank == 0:
                uid.copy_(torch.frombuffer(bytearray(asm.comm_unique_id()), dtype=torch.uint8))
            dist.broadcast(uid, src=0)
            asm.comm_init(bytes(uid.cpu().numpy().tobytes()), rank, world)

    def reduce(self, lpf=1.0):
        self.asm.residual_sums_allreduce(lpf)      # one launch: sums + exchange over NVLink peer memory (else NCCL)


def write_submodel(path, m, owned, q, dd, lpf, layout=1, device=0):
    """binary image of a (sub-)model for the C host cu-bens_b200/host/cb_multi_gpu_demo.c: cb_sizes, cb_flags,
    owned joint range, every cb_model array as (count, data) in the struct's order, q, dd, lpf"""
    import struct
    from . import _MODEL_FIELDS, _model_array
    with open(path, "wb") as f:
        f.write(struct.pack("7l", m.NJ, m.NE_TR, m.NE_FR, m.NE_SH, m.NE_SBR, m.NE_FBR, m.NEQ))
        f.write(struct.pack("5i", m.ANAFLAG, m.ALGFLAG, m.SLVFLAG, layout, device))
        f.write(struct.pack("2l", *(owned if owned is not None else (0, m.NJ))))
        for n in _MODEL_FIELDS:
            a = _model_array(m, n)
            if a is None or a.size == 0:
                f.write(struct.pack("l", 0)); continue
            f.write(struct.pack("l", a.size)); f.write(a.tobytes())
        for v in (q, dd):
            v = np.ascontiguousarray(v, dtype=np.float64)
            f.write(struct.pack("l", v.size)); f.write(v.tobytes())
        f.write(struct.pack("d", float(lpf)))

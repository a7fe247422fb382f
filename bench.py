#!/usr/bin/env python
"""bench.py - BASELINE.json's metric on BASELINE.json's config.

metric   elements assembled / s (K_t + f_int); ms_per_step = per-Newton-iteration assembly ms
workload configs[2]: synthetic 1000x1000 DKT shell plate, 2 000 000 triangles, NEQ 6 000 006,
         geometric nonlinear (ANAFLAG 2), state perturbed by the seeded dd of SURVEY.md 8(d)
step     one Newton-iteration assembly = `ss=0; stiff_sh` + `updatc; f_temp=0; forces_sh;
         ef_ip=ef_i; _ip=_i`  (main.c:1897-2028 minus solve() and test())

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--n 1000]

One JSON line on stdout (rank 0).  `value` is device-resident (dd / f_temp / CSC stay in HBM),
`e2e` goes through the C-ABI with host buffers: dd host->device, f_temp and the CSC values
device->host every step (what the reference's host-side solve() consumes).  The CPU baseline and
`--impl reference` time the UNMODIFIED reference routines (oracle/_ref, compiled from
/root/reference by oracle/Makefile) on all host cores over a bounded sample of the same mesh.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "cu-bens_b200", "python"))

ALG_BYTES_KT = 1192.0     # algorithmic HBM bytes per DKT element, K_t pass
ALG_BYTES_FINT = 654.0    # f_int pass
METRIC = "elements assembled/sec (K_t+f_int)"
UNIT = "elements/s"


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ------------------------------------------------------------------------------------------
class Clocks:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE,
                stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------
# CPU arm: the unmodified reference on a bounded sample (strips of the same plate)
# ------------------------------------------------------------------------------------------
SAMPLE_NX, SAMPLE_NY = 1000, 2           # one strip: 1000 x 2 cells = 4000 triangles


def _cpu_worker(args):
    """time `reps` Newton-iteration assemblies of one strip with the reference routines"""
    wid, reps, warm, cell = args
    from oracle import refbind as R
    from cubens_b200 import meshgen
    m = meshgen.plate_model(SAMPLE_NX, SAMPLE_NY, lx=SAMPLE_NX * cell, ly=SAMPLE_NY * cell)
    s = R.RefState(m)
    s.begin_increment()
    dd = meshgen.perturbation(m, seed=20261017 + wid)
    R.update_forces(m, s, dd); s.end_iteration()
    small = dd * 1e-3
    ts = []
    for it in range(warm + reps):
        t0 = time.perf_counter()
        R.stiff(m, s, SLVFLAG=0)                 # ss=0 ; stiff_sh (skyline scatter)
        R.update_forces(m, s, small)             # updatc ; forces_sh ; ef_ip=ef_i
        s.end_iteration()                        # _ip <- _i
        ts.append(time.perf_counter() - t0)
    return m.NE_SH, ts[warm:]


FRAME_SAMPLE_N = 8        # 8^3-joint lattice per core: 1344 frames, chunk-local skyline of 11 MB


def _cpu_worker_frames(args):
    """the same for the frame path: `reps` iterations of stiff_fr + updatc + forces_fr on a
    small lattice with the unmodified reference routines"""
    wid, reps, warm = args
    from oracle import refbind as R
    from cubens_b200 import meshgen
    m = meshgen.lattice_model(FRAME_SAMPLE_N)
    s = R.RefState(m)
    s.begin_increment()
    dd = meshgen.perturbation(m, scale=1e-3, seed=20261017 + wid)
    R.update_forces(m, s, dd); s.end_iteration()
    small = dd * 1e-2
    ts = []
    for it in range(warm + reps):
        t0 = time.perf_counter()
        R.stiff(m, s, SLVFLAG=0)
        R.update_forces(m, s, small)
        s.end_iteration()
        ts.append(time.perf_counter() - t0)
    return m.NE_FR, ts[warm:]


def cpu_sample_frames(steps, warmup, cores=None):
    import multiprocessing as mp
    cores = cores or os.cpu_count() or 1
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_worker_frames, [(w, steps, warmup) for w in range(cores)])
    ne = res[0][0]
    total = float(np.max(np.array([r[1] for r in res]), axis=0).sum())
    return {"value": cores * ne * steps / total, "unit": UNIT, "cores": cores, "kind": "reference",
            "sample": f"{cores} independent {FRAME_SAMPLE_N}^3-joint lattices ({ne} frames each, chunk-local "
                      f"skyline), {steps} iterations per core, unmodified stiff_fr+updatc+forces_fr at gcc -O2"}


def cpu_sample(steps, warmup, cell, cores=None, o0=False):
    import multiprocessing as mp
    cores = cores or os.cpu_count() or 1
    ctx = mp.get_context("fork")
    if o0:
        os.environ["CUBENS_REF_O0"] = "1"        # read by oracle.refbind in the forked workers
    else:
        os.environ.pop("CUBENS_REF_O0", None)
    with ctx.Pool(cores) as pool:
        t0 = time.perf_counter()
        res = pool.map(_cpu_worker, [(w, steps, warmup, cell) for w in range(cores)])
        wall = time.perf_counter() - t0
    ne = res[0][0]
    per_step = np.max(np.array([r[1] for r in res]), axis=0)      # slowest core per step
    total = float(per_step.sum())
    value = cores * ne * steps / total
    return {"value": value, "unit": UNIT, "cores": cores, "kind": "reference",
            "sample": f"{cores} independent {SAMPLE_NX}x{SAMPLE_NY}-cell strips of the same plate "
                      f"({ne} DKT shells each, chunk-local skyline), {steps} Newton-iteration "
                      f"assemblies per core, unmodified stiff_sh+updatc+forces_sh at gcc -O2",
            "ms_per_step": 1e3 * total / steps, "wall_s": wall,
            "per_core_elements_per_s": ne * steps / total}


def cpu_details(cell):
    """SURVEY 8(d) detail next to the headline CPU figure: the serial reference on ONE core (what the program
    as shipped does), and all cores at the optimisation level it ships with (Makefile:6,21: no -O)"""
    out = {}
    try:
        r1 = cpu_sample(6, 1, cell, cores=1)
        out["one_core_O2"] = {"value": r1["value"], "unit": UNIT, "cores": 1}
        if os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libcubens_ref_O0.so")):
            r0 = cpu_sample(3, 1, cell, o0=True)
            out["all_cores_O0_as_shipped"] = {"value": r0["value"], "unit": UNIT, "cores": r0["cores"]}
            r01 = cpu_sample(2, 1, cell, cores=1, o0=True)
            out["one_core_O0_as_shipped"] = {"value": r01["value"], "unit": UNIT, "cores": 1}
    except Exception as e:                          # extras, never fatal
        out["unavailable"] = repr(e)
    finally:
        os.environ.pop("CUBENS_REF_O0", None)
    return out


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import refbind as R
    if not R.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libcubens_ref.so not built"}))
        return
    cell = 1.0 / a.n
    r = cpu_sample(a.steps, a.warmup, cell)
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT,
            "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": r["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": f"synthetic {a.n}x{a.n} DKT shell plate ({2 * a.n * a.n} elements) "
                                   "geometric-nonlinear Newton: bounded sample, see cpu_baseline.sample"},
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------
# SURVEY.md 8(d), "Frame / truss / brick bytes (same method)"
ALG_BYTES_FR_KT = 1215.0      # 915 B of CSC written (114 nnz / frame on the lattice) + 300 B read
ALG_BYTES_FR_FINT = 700.0
ALG_BYTES_BR_K = 1968.0       # 243 nnz / brick written + 8 joints' coordinates


def other_configs(cb, meshgen, device, peak):
    """BASELINE.json configs[3] and configs[4] shapes on one GPU (not the bench metric; same timing
    method: CUDA events on the library's stream, inputs far larger than the 126 MB L2):
    119^3 joint frame lattice (5 012 994 frames, 7 DOF / joint, geometric nonlinear) - K_t, f_int and
    the lumped mass the Newmark loop refreshes every iteration; 80^3 bricks + 12 800 skin shells -
    stiff_br (+ stiff_sh) and the consistent mass on the same CSC pattern (linear, assembled once in
    the reference)."""
    out = {}
    m = meshgen.lattice_model(119, SLVFLAG=2)
    a = cb.Assembler(m, layout=cb.CB_MAT_CSC, device=device)
    a.begin_increment(); a.stiff(); a.sync()
    dd = meshgen.perturbation(m, scale=1e-3)
    a.update_forces(dd, want_f=False); a.end_iteration()
    a.set_dd(dd * 0.01)
    ks, fs = [], []
    for _ in range(3):
        a.stiff(); a.update_forces_dev(); a.end_iteration()
    K = 20
    a.sync(); a.timer_start()
    for _ in range(K):
        a.stiff(); a.update_forces_dev(); a.end_iteration()
    step_ms = a.timer_stop_ms() / K
    # the Newmark loop of configs[3] (main.c:3590-3619) refreshes the lumped mass with the stiffness in every
    # iteration: the transient step is K_t + mass + f_int
    a.sync(); a.timer_start()
    for _ in range(K):
        a.stiff(); a._check(a.lib.cb_mass(a.h)); a.update_forces_dev(); a.end_iteration()
    tstep_ms = a.timer_stop_ms() / K
    for _ in range(5):
        a.stiff(); ks.append(a.last_stiff_ms); a.update_forces_dev(); fs.append(a.last_forces_ms)
        a.end_iteration()
    a.sync(); a.timer_start()
    for _ in range(5):
        a._check(a.lib.cb_mass(a.h))
    mass_ms = a.timer_stop_ms() / 5
    k_ms, f_ms = float(np.median(ks)), float(np.median(fs))
    out["frame_lattice"] = {
        "workload": f"119^3-joint cubic frame lattice, {m.NE_FR} frames, NEQ {m.NEQ}, nnz "
                    f"{a.lib.cb_csc_nnz(a.h)}, ANAFLAG 2 (BASELINE.json configs[3] shape)",
        "value": m.NE_FR / (step_ms * 1e-3), "unit": UNIT, "ms_per_step": step_ms, "steps": K,
        "transient_step": {"ms_per_step": tstep_ms, "value": m.NE_FR / (tstep_ms * 1e-3),
                           "note": "K_t + lumped-mass refresh + f_int, what one Newmark iteration assembles"},
        "split_ms": {"stiff_total": k_ms, "update_forces": f_ms, "mass_refresh": mass_ms},
        "roofline_frac": {"K_t": ALG_BYTES_FR_KT * m.NE_FR / (k_ms * 1e-3) / 1e9 / peak,
                          "f_int": ALG_BYTES_FR_FINT * m.NE_FR / (f_ms * 1e-3) / 1e9 / peak,
                          "step": (ALG_BYTES_FR_KT + ALG_BYTES_FR_FINT) * m.NE_FR / (step_ms * 1e-3) / 1e9 / peak}}
    a.close()
    m = meshgen.brick_model(80, 80, 80, skin=True)
    a = cb.Assembler(m, layout=cb.CB_MAT_CSC, device=device)
    a.stiff(cb.CB_GEN_COMMITTED); a.sync()
    ks = []
    for _ in range(5):
        a.stiff(cb.CB_GEN_COMMITTED); ks.append(a.last_stiff_ms)
    a._check(a.lib.cb_mass(a.h)); a.sync(); a.timer_start()
    for _ in range(3):
        a._check(a.lib.cb_mass(a.h))
    mass_ms = a.timer_stop_ms() / 3
    k_ms = float(np.median(ks))
    out["brick_skin"] = {
        "workload": f"80^3 8-node bricks + {m.NE_SH} DKT skin shells, NEQ {m.NEQ}, nnz "
                    f"{a.lib.cb_csc_nnz(a.h)} (BASELINE.json configs[4] structure without the fluid; "
                    "linear, assembled once per analysis)",
        "value": m.NE_SBR / (k_ms * 1e-3), "unit": "bricks/s", "ms_stiff": k_ms,
        "ms_consistent_mass": mass_ms,
        "roofline_frac": ALG_BYTES_BR_K * m.NE_SBR / (k_ms * 1e-3) / 1e9 / peak}
    a.close()
    # acoustic FSI (fsi.c, ANAFLAG 4): [K L; 0 H] and [M 0; -rho L^T Q] sparse, on one block pattern
    m = meshgen.fsi_model(64, 64, 32, 32)
    t0 = time.perf_counter()
    a = cb.Assembler(m, layout=cb.CB_MAT_CSC, device=device)
    a.stiff(cb.CB_GEN_COMMITTED); a.sync()
    setup = time.perf_counter() - t0
    ks = []
    for _ in range(3):
        a.stiff(cb.CB_GEN_COMMITTED); ks.append(a.last_stiff_ms)
    a._check(a.lib.cb_mass(a.h)); a.sync(); a.timer_start()
    for _ in range(3):
        a._check(a.lib.cb_mass(a.h))
    mass_ms = a.timer_stop_ms() / 3
    nnz = int(a.lib.cb_csc_nnz(a.h))
    out["fsi"] = {
        "workload": f"64x64x(32 solid + 32 fluid) bricks, acoustic FSI (ANAFLAG 4): {m.NE_SBR} solid + {m.NE_FBR} "
                    f"fluid bricks, {m.SNDOF} structural + {m.FNDOF} pressure equations (BASELINE.json configs[4])",
        "nnz": nnz, "dense_entries_of_the_reference": int(m.NEQ) ** 2,
        "ms_stiff_fsi": float(np.median(ks)), "ms_mass_fsi": mass_ms, "create_plan_first_assembly_s": setup,
        "note": "stiff_fsi / mass_fsi (fsi.c:333-445) build dense NEQ^2 arrays; here both are value arrays on one "
                "sparse block pattern (pressure DOFs on twin joints, L as a two-joint coupling element)"}
    a.close()
    return out


def partition_rows(n, world, rank):
    """contiguous strips of cell rows (along i); returns (i0, i1)"""
    per = n // world
    i0 = rank * per
    i1 = n if rank == world - 1 else i0 + per
    return i0, i1


def run_gpu(a):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries ONE JSON line: anything a library prints on the way (NCCL's version banner ...) goes to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_mod
        torch.cuda.set_device(local)
        dist_mod.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist = dist_mod

    # CPU baseline first (rank 0, N=1 only): fork before any CUDA context exists in this process
    cpu = None
    cpu_frames = None
    if rank == 0 and world == 1 and not a.no_cpu:
        from oracle import refbind as R
        if R.available():
            cpu = cpu_sample(a.cpu_steps, 1, 1.0 / a.n)
            cpu["detail"] = cpu_details(1.0 / a.n)
            if not a.no_others:
                try:
                    cpu_frames = cpu_sample_frames(max(4, a.cpu_steps // 4), 1)
                except Exception as e:                       # the frame baseline is an extra, never fatal
                    cpu_frames = {"unavailable": repr(e)}

    import cubens_b200 as cb
    from cubens_b200 import meshgen
    from cubens_b200.partition import plate_partition

    n = a.n
    from cubens_b200.partition import InterfaceExchange

    def build(weak):
        """the model of this rank (whole plate at N=1), resident on the GPU, perturbed; setup times in seconds"""
        t0 = time.perf_counter()
        if world == 1:
            m = meshgen.plate_model(n, n, SLVFLAG=2)
            owned = None
            n_local = m.NE_SH
        else:
            m, owned, n_local = plate_partition(n, n, world, rank, weak=weak)
        t1 = time.perf_counter()
        asm = cb.Assembler(m, layout=cb.CB_MAT_CSC, device=local)
        if owned is not None:
            asm.set_owned_joints(*owned)
        asm.sync()
        t2 = time.perf_counter()
        asm.begin_increment(); asm.stiff(); asm.sync()        # first cb_stiff builds the element-to-nonzero maps
        t3 = time.perf_counter()
        # seeded perturbation so the plate is no longer flat and every local entry is non-zero
        full_dd = meshgen.perturbation(m)
        asm.update_forces(full_dd, want_f=False)
        asm.end_iteration()
        step_dd = full_dd * 1e-3
        asm.set_dd(step_dd)
        # the collective lives in the library (cb_comm_init / cb_residual_allreduce); N=1 forms the same sums
        exchange = InterfaceExchange(asm, m, world, rank, dist, owned)
        plan_s, plan_dev = asm.plan_info()
        return m, owned, n_local, asm, step_dd, exchange, {
            "mesh_and_reference_maps": t1 - t0, "cb_create": t2 - t1, "plan_and_first_stiff": t3 - t2,
            "plan_build": plan_s, "plan_built_on_device": plan_dev,
            "note": "mesh_and_reference_maps: numpy mesh + the vectorised mirror of codes() / skylin (what main.c does "
                    "before the loop); cb_create: uploads + geometry classes; plan_build: sorted element-to-nonzero "
                    "map, CSC pattern and tile plan (cb_plan_info), part of plan_and_first_stiff"}

    m, owned, n_local, asm, step_dd, exchange, setup_s = build(weak=not a.strong)
    nnz = asm.lib.cb_csc_nnz(asm.h)

    def step():
        asm.stiff()                      # K_t  -> device CSC
        asm.update_forces_dev()          # updatc + f_int -> device f_temp
        if not os.environ.get("BENCH_NO_EXCHANGE"):
            exchange.reduce()            # convergence / reaction sums (+ NCCL all-reduce over NVLink at N > 1)
        asm.end_iteration()

    def barrier():
        asm.sync()
        if dist is not None:
            import torch
            dist.barrier()
            torch.cuda.synchronize()

    # nvidia-smi samples every 100 ms: start it before the warm-up and keep the GPU under the same
    # load for >= 0.5 s so that clocks / throttle reasons are observed under load, then go straight
    # into the timed region
    clocks = Clocks(local)
    if rank == 0:
        clocks.start()
    for _ in range(max(a.warmup, 3)):
        step()
    barrier()
    t_w = time.perf_counter()
    while time.perf_counter() - t_w < 0.6:
        step()
    l0 = asm.launches
    asm_ms, st_ms, fo_ms = [], [], []
    barrier()
    asm.timer_start()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        step()
        if a.detail:
            asm_ms.append(asm.last_assemble_ms); st_ms.append(asm.last_stiff_ms)
            fo_ms.append(asm.last_forces_ms)
    dev_ms = asm.timer_stop_ms()
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t0)
    launches = asm.launches - l0
    clk = clocks.stop() if rank == 0 else None

    # per-kernel split (separate short loop so the event reads do not stall the timed region)
    for _ in range(5):
        asm.stiff(); asm_ms.append(asm.last_assemble_ms); st_ms.append(asm.last_stiff_ms)
        asm.update_forces_dev(); fo_ms.append(asm.last_forces_ms)
        asm.end_iteration()

    if dist is not None:
        import torch
        t = torch.tensor([dev_ms, wall_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, wall_ms = float(t[0]), float(t[1])
        cnt = torch.tensor([float(n_local), float(launches)], device="cuda", dtype=torch.float64)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        total_el, launches = int(cnt[0]), int(cnt[1])
    else:
        total_el = n_local
    ms_step = dev_ms / a.steps
    value = total_el / (ms_step * 1e-3)

    # ---- e2e through the C-ABI with HOST buffers, every rank over its own PCIe link: dd host -> device,
    # f_temp and the CSC values of the owned column slice device -> pinned host, every step.  The matrix
    # crosses as its packed upper triangle (K_t is symmetric) on a second stream and `threads` host threads
    # rebuild the full columns umfpack_di_* takes while later chunks are still on the wire
    # (cb_csc_values_begin / _end around cb_update_forces); `plain` is the same step with the full 2 GB copy
    # of cb_get_csc_values.
    import ctypes as C
    lib = asm.lib
    threads = max(1, min(32, (os.cpu_count() or 1) // max(1, world)))
    # how much of the matrix crosses as full columns (no host work) and how much as packed upper triangles
    # (half the bytes, rebuilt by host threads): with >= 12 threads per rank every third chunk goes in full,
    # with fewer the host cannot keep up with the copy engine and everything does (measured, profiles/README)
    # and with several ranks sharing the host its memory bandwidth is the bottleneck either way)
    os.environ.setdefault("CB_SYM_FULL_EVERY", "3" if (threads >= 12 and world == 1) else "1")
    nnz_u = lib.cb_csc_upper_nnz(asm.h)
    neq_loc = lib.cb_local_equations(asm.h)
    hdd = asm.pinned(m.NEQ); hdd[:] = step_dd
    hf = asm.pinned(m.NEQ)
    hAx = asm.pinned(nnz)
    hAxu = asm.pinned(nnz_u)

    def forces_host():
        cdl = C.c_double(1.0); fr = C.c_int(0); sh = C.c_int(0)
        rc = lib.cb_update_forces(asm.h, C.c_void_p(hdd.ctypes.data), C.byref(cdl), C.c_int(0),
                                  C.c_void_p(hf.ctypes.data), C.byref(fr), C.byref(sh))
        assert rc == 0

    def e2e_step_mirror():
        asm.stiff()
        rc = lib.cb_csc_values_begin(asm.h, C.c_void_p(hAx.ctypes.data), C.c_void_p(hAxu.ctypes.data), C.c_int(threads))
        assert rc == 0, lib.cb_last_error()
        forces_host()
        rc = lib.cb_csc_values_end(asm.h); assert rc == 0
        asm.end_iteration()

    def e2e_step_plain():
        asm.stiff()
        rc = lib.cb_get_csc_values(asm.h, C.c_void_p(hAx.ctypes.data)); assert rc == 0
        forces_host()
        asm.end_iteration()

    def time_e2e(fn, k):
        fn(); barrier()
        t0 = time.perf_counter()
        for _ in range(k):
            fn()
        barrier()
        ms = 1e3 * (time.perf_counter() - t0) / k
        if dist is not None:
            import torch
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t[0])
        return ms

    def e2e_step_upper():
        asm.stiff()
        rc = lib.cb_get_csc_upper_values(asm.h, C.c_void_p(hAxu.ctypes.data)); assert rc == 0
        forces_host()
        asm.end_iteration()

    ke = max(3, min(a.steps, 10))
    e_ms = time_e2e(e2e_step_mirror, ke)
    p_ms = time_e2e(e2e_step_plain, max(3, ke // 2))
    u_ms = time_e2e(e2e_step_upper, max(3, ke // 2))
    # both hand-offs leave the FULL matrix on the host; which one is faster depends on the box (host memory
    # bandwidth against PCIe: 32 vs 38 ms on some boxes of the pool, 39 vs 38 ms on others) - the headline is
    # the faster of the two, both are reported
    mirrored = {"ms_per_step": e_ms, "value": total_el / (e_ms * 1e-3),
                "d2h_bytes_per_step": int(lib.cb_csc_values_d2h_bytes(asm.h) + neq_loc * 8) * world,
                "full_every": int(os.environ["CB_SYM_FULL_EVERY"])}
    best_ms, variant = (e_ms, "cb_csc_values_begin/_end (upper triangle + host mirror, chunked)") if e_ms <= p_ms \
        else (p_ms, "cb_get_csc_values (all nnz values)")
    e2e = {"value": total_el / (best_ms * 1e-3), "unit": UNIT, "ms_per_step": best_ms, "variant": variant,
           "h2d_bytes_per_step": int(neq_loc * 8) * world,
           "d2h_bytes_per_step": mirrored["d2h_bytes_per_step"] if e_ms <= p_ms else int(neq_loc * 8 + nnz * 8) * world,
           "host_threads_per_rank": threads, "steps": ke,
           "mirrored": mirrored,
           "plain": {"ms_per_step": p_ms, "value": total_el / (p_ms * 1e-3),
                     "d2h_bytes_per_step": int(neq_loc * 8 + nnz * 8) * world},
           "upper_only": {"ms_per_step": u_ms, "value": total_el / (u_ms * 1e-3),
                          "d2h_bytes_per_step": int(neq_loc * 8 + nnz_u * 8) * world,
                          "note": "the packed upper-triangular CSC alone (cb_get_csc_upper_values): what a symmetric "
                                  "host solver needs; not the full matrix umfpack_di_* takes"},
           "note": "per rank and step: dd pinned host->device; f_temp and the owned CSC slice device->pinned host in "
                   "chunks on a second stream while cb_update_forces runs - every full_every-th chunk as full columns, "
                   "the others as packed upper triangles whose full columns host threads rebuild (the FULL matrix "
                   "umfpack_di_* consumes ends up on the host); wall clock between barriers, max over ranks; `plain` "
                   "ships all nnz values with cb_get_csc_values"}

    ncls = asm.geometry_classes
    map_bytes = asm.map_bytes
    asm_lib = asm.lib

    # ---- strong scaling (north_star "Target"): per-iteration assembly time of THE fixed n x n plate split
    # over the N ranks, same step, same timing (CUDA events, max over ranks).  At N=1 it is the run above.
    strong = None
    if world > 1 and not a.strong and not a.no_strong:
        asm.close()
        m_s, owned_s, n_loc_s, asm_s, _dd_s, ex_s, _ = build(weak=False)

        def sstep():
            asm_s.stiff(); asm_s.update_forces_dev(); ex_s.reduce(); asm_s.end_iteration()
        for _ in range(5):
            sstep()
        asm_s.sync(); dist.barrier()
        ks = max(a.steps, 50)
        asm_s.timer_start()
        for _ in range(ks):
            sstep()
        s_ms = asm_s.timer_stop_ms() / ks
        import torch
        t = torch.tensor([s_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        s_ms = float(t[0])
        tot = 2 * n * n
        # what the fixed per-step cost is made of: the same loop without the sums + collective, and the
        # CUDA-event times of the two passes alone (rank 0's; read in a separate loop, the reads synchronise)
        asm_s.sync(); dist.barrier()
        asm_s.timer_start()
        for _ in range(ks):
            asm_s.stiff(); asm_s.update_forces_dev(); asm_s.end_iteration()
        nc_ms = asm_s.timer_stop_ms() / ks
        t = torch.tensor([nc_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        nc_ms = float(t[0])
        kk, ff = [], []
        for _ in range(5):
            asm_s.stiff(); kk.append(asm_s.last_stiff_ms); asm_s.update_forces_dev(); ff.append(asm_s.last_forces_ms)
            asm_s.end_iteration()
        strong = {"ms_per_step": s_ms, "value": tot / (s_ms * 1e-3), "unit": UNIT, "elements": tot, "steps": ks,
                  "workload": f"ONE {n}x{n}-cell plate ({tot} elements) split over {world} ranks",
                  "ms_per_step_without_sums_and_collective": nc_ms,
                  "split_ms_rank0": {"stiff_total": float(np.median(kk)), "update_forces": float(np.median(ff))},
                  "note": "divide the N=1 line's ms_per_step by N x this ms_per_step for the strong-scaling efficiency; "
                          "ms_per_step - ms_per_step_without_sums_and_collective = the residual-sums launch + the "
                          "88-byte ncclAllReduce; without - (stiff_total + update_forces) = launch gaps of the four "
                          "kernels"}
        asm_s.close()
        asm = None

    # ---- the same plate with every interior joint moved (seeded): no two shells share their
    # geometry, so every shell streams its own DKT matrix from HBM (N=1, shorter run) -----------
    unstructured = None
    if world == 1 and not a.no_unstructured:
        asm.close(); asm = None
        mu = meshgen.plate_model(n, n, SLVFLAG=2, jitter=0.2)
        au = cb.Assembler(mu, layout=cb.CB_MAT_CSC, device=local)
        ddu = meshgen.perturbation(mu)
        au.begin_increment(); au.update_forces(ddu, want_f=False); au.end_iteration()
        au.set_dd(ddu * 1e-3)

        def ustep():
            au.stiff(); au.update_forces_dev(); au.end_iteration()
        for _ in range(5):
            ustep()
        au.sync()
        ku = max(10, a.steps // 2)
        au.timer_start()
        for _ in range(ku):
            ustep()
        u_ms = au.timer_stop_ms() / ku
        au.stiff(); ka = au.last_assemble_ms; au.update_forces_dev(); kf = au.last_forces_ms
        unstructured = {"value": mu.NE_SH / (u_ms * 1e-3), "unit": UNIT, "ms_per_step": u_ms,
                        "geometry_classes": au.geometry_classes, "steps": ku,
                        "split_ms": {"assemble_kernel": ka, "update_forces": kf},
                        "roofline_frac": ALG_BYTES_KT * mu.NE_SH / (ka * 1e-3) / 1e9 / measured_peak()[0],
                        "workload": "same plate, interior joints moved in-plane by up to 0.2 cell "
                                    "(rng 7): every shell has its own geometry"}
        au.close()

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    peak, peak_src = measured_peak()
    others = None
    if world == 1 and not a.no_others:
        if asm is not None:
            asm.close()
        others = other_configs(cb, meshgen, local, peak)
        if cpu_frames is not None:
            others["frame_lattice"]["cpu_baseline"] = cpu_frames
    k_ms = float(np.median(asm_ms))
    ach = ALG_BYTES_KT * n_local / (k_ms * 1e-3) / 1e9
    traffic = None
    fl_a = fl_f = None
    fl_src = "profiles/traffic.json"
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            tj = json.load(open(tp))
            traffic = tj.get("k_assemble_shell_stream_dram_bytes_per_launch")
            fl_a = tj.get("k_assemble_shell_stream_fp64_flops_per_launch")
            fl_f = tj.get("k_shell_forces_fp64_flops_per_launch")
            fl_src = tj.get("source", "profiles/traffic.json")
        except Exception:
            traffic = None
    # FP64 side of the roofline (north_star: achieved FP64 FLOP/s against the B200 peak): peak from the
    # DFMA micro-kernel measured now, executed flops per launch from the committed ncu capture
    fp64 = None
    try:
        pk = float(asm_lib.cb_measure_fp64_tflops(local))
        fp64 = {"peak_tflops": pk, "peak_source": "DFMA micro-kernel, this run (cb_measure_fp64_tflops)",
                "k_assemble_shell_stream": None if not fl_a else
                {"executed_flops_per_launch": fl_a, "achieved_tflops": fl_a / (k_ms * 1e-3) / 1e12,
                 "frac": fl_a / (k_ms * 1e-3) / 1e12 / pk},
                "flops_source": "dadd + dmul + 2 x dfma thread instructions of the committed ncu capture: " + fl_src}
    except Exception as e:
        fp64 = {"unavailable": repr(e)}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps,
        "warmup": max(a.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "strong" if (world > 1 and a.strong) else "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"synthetic {n}x{n}-cell DKT shell plate per GPU ({total_el} elements "
                               f"in total, NEQ {m.NEQ}) geometric-nonlinear Newton-iteration "
                               "assembly (BASELINE.json configs[2])",
                   "matrix": f"device CSC, nnz {nnz} per rank", "parallelism": f"element-partition x{world}",
                   "l2": "inputs and outputs (2 GB CSC + 1.7 GB element state per GPU) exceed the 126 MB L2",
                   "seed": 20261017,
                   "geometry_classes": f"{ncls} (shells with bit-identical geometry share one cached DKT "
                                       "matrix; `unstructured` repeats the run on a jittered plate where "
                                       "none do)" if ncls else "0 (every shell streams its own DKT matrix)"},
        "clocks": clk, "gpu_launches": int(launches), "wall_ms_per_step": wall_ms / a.steps,
        "split_ms": {"stiff_total": float(np.median(st_ms)), "assemble_kernel": k_ms,
                     "update_forces": float(np.median(fo_ms))},
        "roofline": {"kernel": "k_assemble_shell_stream", "bound": "hbm", "achieved": ach, "peak": peak,
                     "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                     "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": ALG_BYTES_KT * n_local,
                     "whole_step_frac": (ALG_BYTES_KT + ALG_BYTES_FINT) * n_local / (ms_step * 1e-3) / 1e9 / peak,
                     "map_bytes_per_launch": map_bytes},
    }
    line["fp64"] = fp64
    line["setup_s"] = setup_s
    if strong is not None:
        line["strong"] = strong
    elif world == 1:
        line["strong"] = {"ms_per_step": ms_step, "value": value, "unit": UNIT, "elements": total_el,
                          "note": "N=1: the weak and the strong workload coincide"}
    if unstructured is not None:
        line["unstructured"] = unstructured
    if others is not None:
        line["other_configs"] = others
    if e2e is not None:
        line["e2e"] = e2e
    if cpu is not None:
        line["cpu_baseline"] = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample", "detail")}
    sys.stdout.flush()
    os.dup2(real_stdout, 1)
    print(json.dumps(line), flush=True)
    os.dup2(2, 1)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native")
    ap.add_argument("--n", type=int, default=1000, help="plate cells per side (1000 = BASELINE)")
    ap.add_argument("--cpu-steps", type=int, default=40)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--detail", action="store_true")
    ap.add_argument("--no-unstructured", action="store_true",
                    help="skip the second (jittered-plate) measurement")
    ap.add_argument("--no-others", action="store_true",
                    help="skip the frame-lattice / brick measurements (BASELINE configs[3], [4] shapes)")
    ap.add_argument("--no-strong", action="store_true", help="N>1: skip the strong-scaling measurement")
    ap.add_argument("--strong", action="store_true",
                    help="N>1: split ONE n x n plate across the ranks (default: weak scaling, every "
                         "rank owns an n x n-cell strip of an (n*N) x n plate)")
    a = ap.parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_gpu(a)


if __name__ == "__main__":
    main()

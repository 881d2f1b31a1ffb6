#!/usr/bin/env python
"""bench.py — assembled cells/s of the batched Example200 `assemble!` (FP64 stiffness + rhs) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--n N_PER_SIDE] [--jitter]

Workload (BASELINE.json configs[1]): H1P1 2-D Poisson stiffness + load vector on the synthetic
`simplexgrid(LinRange(0,1,n+1))^2` mesh, n = 5657 -> 64,003,298 triangles / 32,012,964 nodes per GPU
(weak scaling: every rank assembles a slab of that size; interface columns are exchanged over NCCL).
A "step" is one steady-state assembly (pattern + cell->slot maps precomputed, like the reference's
second, allocation-free assembly — Example200:205-211).

* value  : device-resident inputs, CUDA events on the library's stream around exactly K steps.
* e2e    : the same through the C ABI with HOST buffers: every step uploads Coordinates, CellNodes and
           CellVolumes from pinned memory (femb_mesh_set), assembles, and downloads nzval and b.
* roofline: algorithmic bytes/cell * cells / mean kernel time, against MEASURED_PEAKS.json.
* cpu_baseline / --impl reference: the C restatement of the reference's loop (oracle/femb_oracle_c.c)
  on the host cores, on a bounded sample of the same mesh.  (The reference is pure Julia and no julia
  binary exists in this image: kind = "port".)
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

N_DEFAULT = 5657  # 2 * 5657^2 = 64,003,298 triangles
METRIC = "assembled cells/s (FP64 stiffness + rhs, H1P1 2-D Poisson, Example200 assemble!)"
UNIT = "cells/s"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ---------------------------------------------------------------------------------------------
def algorithmic_bytes_per_cell(nnodes, ncells, nnz, with_cellvol=True):
    """SURVEY.md §8(d) / DESIGN.md §4: every array of the path touched once.
    CellNodes 12 + Coordinates 16*nnodes/ncells + CellDofs 12 + cell->slot map 36 + nzval 8*nnz/ncells + b 8*nnodes/ncells
    (+ CellVolumes 8 when the caller provides them, as Example200:114 does)."""
    return 12.0 + 16.0 * nnodes / ncells + 12.0 + 36.0 + 8.0 * nnz / ncells + 8.0 * nnodes / ncells + (8.0 if with_cellvol else 0.0)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, device):
        self.device, self.rows, self.proc = device, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 6 and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
def make_problem(n, jitter):
    import femb_b200 as fb

    t0 = time.time()
    X = np.linspace(0.0, 1.0, n + 1)
    grid = fb.simplexgrid(X, X)
    if jitter:
        grid = fb.jitter(grid)
    vol = grid["CellVolumes"]
    log(f"[bench] mesh n={n}: {grid.ncells} cells, {grid.nnodes} nodes ({time.time() - t0:.1f}s host)")
    return grid, vol


def setup_device(grid, vol, device, use_vol=True, part=None):
    """Symbolic phase (once): register mesh / space / rule / evaluators, build the structural pattern on the
    device, run the reference's FIRST assembly (value filter decides the pattern) and compress.  With a partition
    (`part`, N > 1): halo rows are added to the pattern and the NCCL exchange plan is installed."""
    import femb_b200 as fb
    from femb_b200 import _lib, partition as pt
    from femb_b200.assembly import AffineFunction, _Session

    fes = fb.FESpace(fb.H1Pk(1, 2, 1), grid)
    qf = fb.QuadratureRule(grid.geometry, 0)  # Example200:110 with order 1 -> midpoint rule
    t0 = time.time()
    S = _Session([fes], qf, device, cellvolumes=use_vol)
    ctx = S.ctx
    eg, ei = S.evaluator(0, fb.Gradient), S.evaluator(0, fb.Identity)
    ctx.pattern_track(True)
    nnz_struct = ctx.pattern_build([(0, 0)])
    if part is not None:
        import torch.distributed as dist

        uid = [_lib.Context.comm_unique_id() if part.rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(part.nranks, part.rank, uid[0])
        colptr, rowval = ctx.pattern_get(grid.nnodes)
        sent = pt.allgather(pt.ghost_columns(part, colptr, rowval), part.nranks)
        er, ec = pt.extra_entries(part, sent)
        if er.size:
            ctx.pattern_extra(er, ec)
            nnz_struct = ctx.pattern_build([(0, 0)])
            colptr, rowval = ctx.pattern_get(grid.nnodes)
        sent = pt.allgather(pt.ghost_columns(part, colptr, rowval), part.nranks)
        ctx.comm_set_exchange(pt.exchange_plan(part, colptr, rowval, sent))
    rhs = AffineFunction([0.0], [[1.0, -1.0]])  # Example200:38
    ctx.matrix_zero()
    ctx.vector_zero()
    ctx.assemble_poisson(eg._eid, ei._eid, 1.0, _lib.RHS_AFFINE, rhs.params(), 1.0e-15)
    if part is not None:
        ctx.exchange()  # received non-zero contributions mark their slots as touched
    nnz = ctx.pattern_compress()
    ctx.pattern_track(False)
    if part is not None:
        colptr, rowval = ctx.pattern_get(grid.nnodes)
        sent = pt.allgather(pt.ghost_columns(part, colptr, rowval), part.nranks)
        plan = pt.exchange_plan(part, colptr, rowval, sent)
        ctx.comm_set_exchange(plan)
        ctx.set_fused_exchange(True)  # from now on femb_assemble_poisson overlaps the exchange with the bulk of the assembly
        S.plan = plan
        S.exchange_bytes = 8 * int(plan.send_ptr[-1] + plan.sendb_ptr[-1] + plan.recv_ptr[-1] + plan.recvb_ptr[-1])
    ctx.sync()
    log(f"[bench] symbolic phase: structural nnz={nnz_struct}, after Example200's 1e-15 filter nnz={nnz} ({time.time() - t0:.1f}s)")
    return S, fes, eg._eid, ei._eid, rhs.params(), nnz


def step_resident(ctx, eg, ei, rhsp, exchange=False):
    from femb_b200 import _lib

    ctx.matrix_zero()
    ctx.vector_zero()
    ctx.assemble_poisson(eg, ei, 1.0, _lib.RHS_AFFINE, rhsp, 1.0e-15)
    ms = ctx.last_kernel_ms()[0]
    if exchange:
        ctx.exchange()  # interface columns -> owners (NCCL send/recv + add kernel)
    return ms


def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import __graft_entry__  # noqa: F401  (sys.path for the shim)
    import femb_b200 as fb  # noqa: F401
    from femb_b200 import build as fbuild

    fbuild.build()
    part = None
    if world > 1:
        from femb_b200 import partition as pt

        t0 = time.time()
        part = pt.slab_partition_2d(args.n, rank, world)
        grid = part.grid
        vol = grid["CellVolumes"]
        log(f"[bench] rank {rank}: slab n={args.n}: {grid.ncells} cells, {grid.nnodes} local nodes ({time.time() - t0:.1f}s host)")
    else:
        grid, vol = make_problem(args.n, args.jitter)
    use_vol = args.cellvol == "given"
    S, fes, eg, ei, rhsp, nnz = setup_device(grid, vol, local, use_vol, part)
    ctx = S.ctx
    ncells, nnodes = grid.ncells, grid.nnodes
    xch = world > 1

    def barrier():
        if world > 1:
            dist.barrier()
        ctx.sync()

    # ---- value: device-resident ------------------------------------------------------------
    clk = ClockSampler(local)
    clk.__enter__()
    # device-resident steps: the interface exchange is fused into (overlapped with) femb_assemble_poisson
    for _ in range(args.warmup):
        step_resident(ctx, eg, ei, rhsp, False)
    barrier()
    l0 = ctx.launch_count()
    kms = []
    ctx.region_begin()
    for _ in range(args.steps):
        kms.append(step_resident(ctx, eg, ei, rhsp, False))
    ms_total = ctx.region_end()
    barrier()
    launches = ctx.launch_count() - l0
    if world > 1:
        t = torch.tensor([ms_total], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = world * ncells / (ms_step * 1e-3)

    # ---- e2e: host buffers, H2D + D2H inside the timed region ------------------------------
    coords, cn = grid["Coordinates"], grid["CellNodes"]
    nz_host = np.empty(nnz, dtype=np.float64)
    b_host = np.empty(nnodes, dtype=np.float64)
    vol_in = vol if use_vol else None
    pinned = [coords, cn, nz_host, b_host] + ([vol] if use_vol else [])
    for a in pinned:
        ctx.host_register(a)

    from femb_b200 import _lib

    if xch:
        plan = S.plan
        r_lo, r_n = (int(plan.recv_slots.min()), int(plan.recv_slots.max() - plan.recv_slots.min() + 1)) if plan.recv_slots.size else (0, 0)
        b_lo, b_n = (int(plan.recv_b.min()), int(plan.recv_b.max() - plan.recv_b.min() + 1)) if plan.recv_b.size else (0, 0)

    def step_e2e():
        # upload the mesh, assemble, download nzval and b — pipelined over column chunks inside the library
        ctx.assemble_poisson_host(eg, ei, 1.0, _lib.RHS_AFFINE, rhsp, 1.0e-15, coords, cn, vol_in, nz_host, b_host, args.chunks)
        if xch:  # interface columns: exchange, then re-download the (small, contiguous) ranges that received contributions
            ctx.exchange()
            if r_n:
                ctx.matrix_get_range(r_lo, r_n, nz_host[r_lo:r_lo + r_n])
            if b_n:
                ctx.vector_get_range(b_lo, b_n, b_host[b_lo:b_lo + b_n])

    for _ in range(min(args.warmup, 3)):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    ctx.region_begin()
    for _ in range(args.steps):
        step_e2e()
    e2e_ms_dev = ctx.region_end()
    barrier()
    e2e_wall = (time.perf_counter() - t0) * 1e3
    e2e_ms = max(e2e_ms_dev, e2e_wall)
    if world > 1:
        t = torch.tensor([e2e_ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e_value = world * ncells / (e2e_ms / args.steps * 1e-3)
    # variant: the caller declares the connectivity unchanged (cellnodes = NULL): only the geometry is uploaded
    geo_ms = None
    if not xch:
        def step_geo():
            ctx.assemble_poisson_host(eg, ei, 1.0, _lib.RHS_AFFINE, rhsp, 1.0e-15, coords, None, vol_in, nz_host, b_host, args.chunks)

        step_geo()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_geo()
        barrier()
        geo_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    clk.__exit__()
    clocks = clk.summary()  # sampled over warm-up + both timed regions
    h2d = coords.nbytes + cn.nbytes + (vol.nbytes if use_vol else 0)
    d2h = nz_host.nbytes + b_host.nbytes
    checks = sanity_checks(nz_host, b_host, S, nnodes, part)
    for a in pinned:
        ctx.host_unregister(a)

    # ---- roofline of the dominant kernel ----------------------------------------------------
    peaks = {}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peaks = json.load(open(pk))
    peak = float(peaks.get("hbm_gbs", 6650.0))
    bpc = algorithmic_bytes_per_cell(nnodes, ncells, nnz, use_vol)
    kernel_ms = float(np.mean(kms))
    achieved = bpc * ncells / (kernel_ms * 1e-3) / 1e9
    # dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of this kernel, per cell, scaled to this launch
    traffic = None
    tj = os.path.join(ROOT, "profiles", "p1_gather2d_traffic.json")
    if os.path.exists(tj):
        t = json.load(open(tj))
        key = "given" if use_vol else "device"
        if key in t:
            traffic = {"bytes_per_launch": round(t[key]["dram_bytes_per_cell"] * ncells), "bytes_per_cell": t[key]["dram_bytes_per_cell"], "source": t[key]["source"]}
    roofline = {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                "traffic": traffic, "peak_source": "MEASURED_PEAKS.json (of measured)" if peaks else "B200_PROFILING.md fallback (of fallback)",
                "kernel": "k_p1_poisson_gather2d", "kernel_ms": round(kernel_ms, 4), "algorithmic_bytes_per_cell": round(bpc, 2)}

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"H1P1 2-D Poisson stiffness+rhs, simplexgrid n={args.n} ({ncells} triangles, {nnodes} nodes) per GPU" + (", jittered" if args.jitter else ""),
                   "nnz": nnz, "l2": "inputs (%.1f GB/step) far larger than the 126 MB L2; no flush needed" % (bpc * ncells / 1e9),
                   "partition": ("one %d-row slab per rank, interface columns owned by the lower rank, NCCL send+recv of %d bytes on rank 0 per step" % (args.n, getattr(S, "exchange_bytes", 0))) if world > 1 else "single GPU", "scatter": "owner-computes gather (write-once, bit-reproducible)",
                   "cellvolumes": "xgrid[CellVolumes] passed in" if use_vol else "NULL: |det A|/2 on the device"},
        "clocks": clocks, "gpu_launches": int(launches),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_ms / args.steps,
                "note": "every step uploads Coordinates + CellNodes (+ CellVolumes) and downloads nzval + b"},
        "e2e_geometry_only": None if geo_ms is None else {"value": world * ncells / (geo_ms * 1e-3), "unit": UNIT, "ms_per_step": geo_ms,
                                                          "h2d_bytes_per_step": int(h2d - cn.nbytes), "d2h_bytes_per_step": int(d2h),
                                                          "note": "connectivity declared unchanged (cellnodes = NULL): the numeric phase reads the geometry only"},
        "roofline": roofline, "checks": checks,
    }
    if rank == 0 and world == 1 and not args.no_cpu:
        out["cpu_baseline"] = cpu_baseline(grid, vol, S, nnz, args.cpu_seconds)
    if rank == 0:
        emit(out)
    if world > 1:
        dist.destroy_process_group()


def sanity_checks(nz, b, S, nnodes, part=None):
    """Size-independent properties of the full-size result (tests/ hold the oracle comparisons)."""
    colptr, rowval = S.ctx.pattern_get(nnodes)
    cols = np.repeat(np.arange(nnodes, dtype=np.int64), np.diff(colptr))
    if part is None:
        rowsum = np.bincount(rowval - 1, weights=nz, minlength=nnodes)  # A * 1
        colsum = np.bincount(cols, weights=nz, minlength=nnodes)  # 1^T A
        return {"max_abs_A_times_ones": float(np.abs(rowsum).max()), "max_abs_ones_times_A": float(np.abs(colsum).max()), "sum_b": float(b.sum()),
                "diag_min": float(nz[rowval - 1 == cols].min())}
    # distributed: every OWNED column is complete after the exchange -> its entries sum to 0 and, on this structured
    # mesh, an owned interface node away from the boundary has the interior 5-point column (4; -1 -1 -1 -1)
    owned = part.owner == part.rank
    colsum = np.bincount(cols, weights=nz, minlength=nnodes)
    diag = np.zeros(nnodes)
    isd = rowval - 1 == cols
    diag[cols[isd]] = nz[isd]
    n1 = part.row_nodes
    top = np.nonzero(owned)[0][-n1:]  # the rank's last owned node row = the cut (or the domain boundary on the last rank)
    return {"max_abs_owned_colsum": float(np.abs(colsum[owned]).max()), "cut_row_diag_interior": float(np.median(diag[top][1:-1])),
            "owned_diag_min": float(diag[owned].min())}


# ---------------------------------------------------------------------------------------------
def _tables():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import femb_oracle as orc
    import oracle_c

    el = orc.Element("H1Pk", 1, 1)
    xref, w = orc.quadrature("Triangle2D", 0)
    vals, der = el.tabulate("Triangle2D", xref)
    return orc, oracle_c, xref, w, vals[:, :, 0], der[:, :, 0, :]


def cpu_run(grid, vol, colptr, rowval, seconds, nthreads=None):
    """Time the C restatement of the reference loop on cells [0, S) of the same mesh, S sized for ~`seconds`."""
    orc, oracle_c, xref, w, vals, der = _tables()
    coords, cn = grid["Coordinates"], grid["CellNodes"]
    nthreads = nthreads or oracle_c.max_threads()
    rhs = np.array([0.0, 1.0, -1.0])
    nz = np.zeros(rowval.shape[0])
    b = np.zeros(grid.nnodes)

    def run(s):
        t0 = time.perf_counter()
        miss = oracle_c.example200(coords, cn, vol, cn, w, xref, vals, der, 1.0, rhs, 1e-15, colptr, rowval, nz, b, nthreads, 0, s)
        assert miss == 0, "CPU baseline: contribution without a slot"
        return time.perf_counter() - t0

    s0 = min(grid.ncells, 1 << 20)
    run(s0)  # warm-up / page-in
    t = run(s0)
    sample = int(min(grid.ncells, max(s0, s0 * seconds / max(t, 1e-6))))
    run(sample)  # touch every page of the sample once
    reps = int(max(1, min(20, seconds / max(run(sample), 1e-6))))  # ~`seconds` of wall time in total
    ts = [run(sample) for _ in range(reps)]
    t = float(np.median(ts))
    return sample / t, sample, nthreads, t, reps


def cpu_baseline(grid, vol, S, nnz, seconds):
    colptr, rowval = S.ctx.pattern_get(grid.nnodes)
    rate, sample, nthreads, t, reps = cpu_run(grid, vol, colptr, rowval, seconds)
    return {"value": rate, "unit": UNIT, "cores": nthreads, "kind": "port",
            "sample": f"cells [0,{sample}) of the same mesh, C restatement of the reference loop (oracle/femb_oracle_c.c), OpenMP over cells with atomic adds, "
                      f"median of {reps} passes of {t:.2f}s"}


def run_reference(args):
    """--impl reference: the reference's CPU path.  The reference is pure Julia and no julia binary exists in
    this image, so this times the C restatement of its loop (kind = port) on all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    import __graft_entry__  # noqa: F401
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import femb_oracle as orc

    # bounded sample: a slab of the workload mesh (full width n, as many rows as ~cpu_seconds allow)
    rows = max(4, min(args.n, int(args.ref_cells // (2 * args.n))))
    import femb_b200 as fb

    X = np.linspace(0.0, 1.0, args.n + 1)
    grid = fb.simplexgrid(X, X[: rows + 1])
    vol = grid["CellVolumes"]
    cn = grid["CellNodes"]
    el = orc.Element("H1Pk", 1, 1)
    log(f"[bench/reference] sample slab: {grid.ncells} cells; building its (filtered) pattern with the numpy oracle")
    keys = []
    chunk = 1 << 21
    for c0 in range(0, grid.ncells, chunk):  # first assembly of the reference, symbolic part only
        Aloc, _ = orc.example200_local(el, "Triangle2D", grid["Coordinates"], cn[c0:c0 + chunk], vol[c0:c0 + chunk], 1.0, None)
        I, J, _ = orc.example200_triplets(Aloc, cn[c0:c0 + chunk])
        keys.append(np.unique((J - 1) * grid.nnodes + (I - 1)))
    keys = np.unique(np.concatenate(keys))
    colptr, rowval, _ = orc.flush_csc(grid.nnodes, grid.nnodes, keys % grid.nnodes + 1, keys // grid.nnodes + 1, np.zeros(keys.shape[0]))
    del keys, Aloc, I, J
    _, oracle_c, xref, w, vals, der = _tables()
    nthreads = oracle_c.max_threads()
    rhs = np.array([0.0, 1.0, -1.0])
    nz, b = np.zeros(rowval.shape[0]), np.zeros(grid.nnodes)

    def step():
        nz[:] = 0.0
        b[:] = 0.0
        t0 = time.perf_counter()
        miss = oracle_c.example200(grid["Coordinates"], cn, vol, cn, w, xref, vals, der, 1.0, rhs, 1e-15, colptr, rowval, nz, b, nthreads)
        assert miss == 0
        return time.perf_counter() - t0

    for _ in range(args.warmup):
        step()
    ts = [step() for _ in range(args.steps)]
    ms_step = 1e3 * float(np.mean(ts))
    value = grid.ncells / (ms_step * 1e-3)
    sample = f"slab of the workload mesh: {args.n} x {rows} quads = {grid.ncells} cells per step"
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": f"H1P1 2-D Poisson stiffness+rhs, simplexgrid n={args.n} (bounded sample: {sample})"},
           "cpu_baseline": {"value": value, "unit": UNIT, "cores": nthreads, "kind": "port", "sample": sample},
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    emit(out)


_REAL_STDOUT = None


def emit(obj):
    """The ONE JSON line goes to the real stdout; everything else any library prints (NCCL banners, torchrun notes)
    was redirected to stderr in main()."""
    line = (json.dumps(obj) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(line.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, line)


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", "--n", dest="n", type=int, default=N_DEFAULT, help="quads per side (cells = 2 n^2)")
    ap.add_argument("--jitter", action="store_true", help="perturb interior nodes (full 7-point pattern)")
    ap.add_argument("--cellvol", default="given", choices=["given", "device"], help="pass xgrid[CellVolumes] (as Example200:114 reads it) or NULL")
    ap.add_argument("--chunks", type=int, default=24, help="column chunks of the streamed host-to-host assembly (e2e)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--ref-cells", type=float, default=16e6, help="cells per step of the --impl reference sample")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

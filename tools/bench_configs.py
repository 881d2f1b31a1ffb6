#!/usr/bin/env python
"""Secondary configs of BASELINE.json (C3 P2-3D stiffness, C4 Hdiv mass+div, C5 Taylor-Hood) timed on the device
through the same C ABI as bench.py (device-resident inputs, CUDA events of the library).  Not the headline metric:
prints one JSON line per config with cells/s and the per-call kernel time, for DESIGN.md / profiles/."""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__  # noqa: E402,F401
import femb_b200 as fb  # noqa: E402
from femb_b200 import _lib  # noqa: E402
from femb_b200.assembly import AffineFunction, Example210RHS, _Session  # noqa: E402


def timeit(fn, ctx, steps, warmup):
    for _ in range(warmup):
        fn()
    ctx.sync()
    ctx.region_begin()
    for _ in range(steps):
        fn()
    ms = ctx.region_end() / steps
    return ms


def c3_p2_3d(m, steps, warmup, order=2, check=False, e2e=False):
    X = np.linspace(0, 1, m + 1)
    t0 = time.time()
    grid = fb.simplexgrid(X, X, X)
    ft = fb.H1P2(1, 3) if order == 2 else fb.H1P1(1)
    if order == 2:
        grid.instantiate_on_device(0, components=("edges",))  # xgrid[CellEdges] from the device enumeration (ndofs needs nedges)
    fes = fb.FESpace(ft, grid)
    qf = fb.QuadratureRule(grid.geometry, 2 * (order - 1))
    S = _Session([fes], qf, device_dofmaps=order == 2)  # CellDofs generated on the device from the pattern "N1E1"
    ctx = S.ctx
    eg, ei = S.evaluator(0, "Gradient"), S.evaluator(0, "Identity")
    t1 = time.time()
    nnz = ctx.pattern_build([(0, 0)])
    ctx.sync()
    t2 = time.time()
    p = AffineFunction([0.0], [[1.0, -1.0, 0.5]]).params()

    def step():
        ctx.matrix_zero()
        ctx.vector_zero()
        ctx.assemble_poisson(eg._eid, ei._eid, 1.0, _lib.RHS_AFFINE, p, 0.0)

    ms = timeit(step, ctx, steps, warmup)
    res = {"config": f"C3 H1P{order}{{1,3}} Poisson stiffness+rhs, Kuhn m={m}", "ncells": grid.ncells, "ndofs": fes.ndofs, "nnz": nnz, "ms_per_step": ms,
           "cells_per_s": grid.ncells / (ms * 1e-3), "host_setup_s": t1 - t0, "pattern_s": t2 - t1}
    if e2e and order == 2:
        # host buffers end to end (femb_assemble_poisson_host): geometry up, finished nzval / b ranges down behind the kernels
        nz_host, b_host = np.empty(nnz), np.empty(fes.ndofs)
        coords, vols = np.ascontiguousarray(grid["Coordinates"]), np.ascontiguousarray(grid["CellVolumes"])
        for arr in (coords, vols, nz_host, b_host):
            ctx.host_register(arr)

        def step_e2e():
            ctx.assemble_poisson_host(eg._eid, ei._eid, 1.0, _lib.RHS_AFFINE, p, 0.0, coords, None, vols, nz_host, b_host, 32)

        for _ in range(2):
            step_e2e()
        t0 = time.perf_counter()
        for _ in range(steps):
            step_e2e()
        ms_e2e = (time.perf_counter() - t0) * 1e3 / steps
        res.update({"e2e_ms_per_step": ms_e2e, "e2e_cells_per_s": grid.ncells / (ms_e2e * 1e-3), "e2e_h2d_bytes": coords.nbytes + vols.nbytes, "e2e_d2h_bytes": nz_host.nbytes + b_host.nbytes,
                    "e2e_d2h_GBps": (nz_host.nbytes + b_host.nbytes) / (ms_e2e * 1e-3) / 1e9})
        for arr in (coords, vols, nz_host, b_host):
            ctx.host_unregister(arr)
    if check:
        # size-independent properties of the full result: 1^T A = 0 (every column sums to zero), diag > 0, sum(b) = int f
        nz = ctx.matrix_get()
        colptr = ctx.pattern_get_colptr(fes.ndofs)
        colsum = np.add.reduceat(nz, colptr[:-1] - 1)
        b = ctx.vector_get(fes.ndofs)
        res.update({"check_max_abs_colsum": float(np.abs(colsum).max()), "check_max_abs_entry": float(np.abs(nz).max()), "check_nonzero_entries": int(np.count_nonzero(nz)),
                    "check_sum_b": float(b.sum()), "check_int_f": 0.25})  # int over the unit cube of x - y + z/2
    return res


def c4_hdiv(m, steps, warmup, name, scatter=None):
    X = np.linspace(0, 1, m + 1)
    grid = fb.simplexgrid(X, X, X)
    ft = fb.HDIVRT0(3) if name == "RT0" else fb.HDIVBDM1(3)
    FESv, FESq = fb.FESpace(ft, grid), fb.FESpace(fb.L2P0(1), grid)
    qf = fb.QuadratureRule(grid.geometry, 2)
    S = _Session([FESv, FESq], qf)
    ctx = S.ctx
    eid, ediv, eq = S.evaluator(0, "Identity"), S.evaluator(0, "Divergence"), S.evaluator(1, "Identity")
    nnz = ctx.pattern_build([(0, 0), (0, 1), (1, 0)])
    if scatter is not None:
        ctx.set_scatter(scatter)

    def step():
        ctx.matrix_zero()
        ctx.assemble_hdiv_mixed(eid._eid, ediv._eid, eq._eid, 1.0, 1.0)

    ms = timeit(step, ctx, steps, warmup)
    return {"config": f"C4 HDIV{name}{{3}} mass + (div v, q) vs L2P0, Kuhn m={m}" + ("" if scatter is None else f", scatter policy {scatter}"), "ncells": grid.ncells, "ndofs": FESv.ndofs + FESq.ndofs, "nnz": nnz,
            "ms_per_step": ms, "cells_per_s": grid.ncells / (ms * 1e-3)}


def c5_taylor_hood(n, steps, warmup):
    X = np.linspace(0, 1, n + 1)
    grid = fb.simplexgrid(X, X)
    FESu, FESp = fb.FESpace(fb.H1Pk(2, 2, 2), grid), fb.FESpace(fb.H1Pk(1, 2, 1), grid)
    qf = fb.QuadratureRule(grid.geometry, 5)
    S = _Session([FESu, FESp], qf)
    ctx = S.ctx
    egu, eiu, eip = S.evaluator(0, "Gradient"), S.evaluator(0, "Identity"), S.evaluator(1, "Identity")
    nnz = ctx.pattern_build([(0, 0), (0, 1), (1, 0)], [FESu.ndofs + 1])
    sol = np.random.default_rng(1).uniform(-1, 1, FESu.ndofs + FESp.ndofs)
    ctx.vector_set(1, sol)
    rp = Example210RHS(1e-2, 0.0).params()
    out = []
    for label, lin, nonlin in (("linear (Laplacian + div-pressure + rhs)", True, False), ("nonlinear (Newton convection + rhs)", False, True)):
        def step():
            ctx.matrix_zero()
            ctx.vector_zero()
            ctx.assemble_navierstokes(egu._eid, eiu._eid, eip._eid, 1e-2, lin, nonlin, _lib.RHS_EX210 if lin else _lib.RHS_NONE, rp if lin else None)

        ms = timeit(step, ctx, steps, warmup)
        out.append({"config": f"C5 Example210 Taylor-Hood update_system! {label}, n={n}", "ncells": grid.ncells, "ndofs": FESu.ndofs + FESp.ndofs, "nnz": nnz,
                    "ms_per_step": ms, "cells_per_s": grid.ncells / (ms * 1e-3)})
    # the steps after the assembly on the device: boundary dofs, 1e60 penalisation, residual norm
    from femb_b200.fespace import parse_pattern

    segs = parse_pattern(FESu.fetype.get_dofmap_pattern(grid.geometry))
    t0 = time.perf_counter()
    nb = ctx.boundary_dofs(0, segs, download=False)
    ctx.sync()
    t_bd = (time.perf_counter() - t0) * 1e3
    pin = np.array([FESu.ndofs + 1])
    ctx.residual(None, zero_fixed=True)  # builds the row-wise view once
    ms_dir = timeit(lambda: ctx.dirichlet_apply(True, 0, pin, None, 1.0e60), ctx, steps, warmup)
    ms_res = timeit(lambda: ctx.residual(None, zero_fixed=True), ctx, steps, warmup)
    out.append({"config": f"C5 after-assembly steps on the device, n={n}", "boundary_dofs": nb, "boundary_dofs_ms_first_call": t_bd, "dirichlet_apply_ms": ms_dir,
                "residual_norm_ms": ms_res, "nnz": nnz, "residual_GBps": (nnz * 16 + (FESu.ndofs + FESp.ndofs) * 24) / (ms_res * 1e-3) / 1e9})
    return out


def topo(m, host_limit=128):
    """face / edge enumeration + CellDofs of H1P2{1,3} on the device vs the numpy restatement (the host
    side is only run up to m = host_limit; it needs minutes and tens of GB beyond)."""
    from femb_b200.fespace import parse_pattern

    X = np.linspace(0, 1, m + 1)
    grid = fb.simplexgrid(X, X, X)
    ctx = _lib.Context(0)
    ctx.mesh_set(grid["Coordinates"], grid["CellNodes"], None)
    ctx.sync()
    res = {"config": f"topology of the Kuhn mesh m={m}: faces, edges, CellDofs(H1P2{{1,3}}) on the device", "ncells": grid.ncells}
    for name, kind in (("faces", _lib.ENTITY_FACES), ("edges", _lib.ENTITY_EDGES)):
        t0 = time.perf_counter()
        n = ctx.grid_entities(kind)
        ctx.sync()
        res[f"n{name}"] = n
        res[f"device_{name}_ms"] = (time.perf_counter() - t0) * 1e3
    ft = fb.H1P2(1, 3)
    t0 = time.perf_counter()
    ndofs, _ = ctx.space_set_from_pattern(0, _lib.FAMILY["H1"], 1, 10, 10, 0, parse_pattern(ft.get_dofmap_pattern(grid.geometry)))
    ctx.sync()
    res["device_celldofs_ms"] = (time.perf_counter() - t0) * 1e3
    res["ndofs"] = ndofs
    res["device_items_per_s"] = grid.ncells * 10 / ((res["device_faces_ms"] + res["device_edges_ms"]) * 1e-3)
    if m <= host_limit:
        t0 = time.perf_counter()
        grid["CellFaces"]
        res["host_faces_ms"] = (time.perf_counter() - t0) * 1e3
        t0 = time.perf_counter()
        grid["CellEdges"]
        res["host_edges_ms"] = (time.perf_counter() - t0) * 1e3
        t0 = time.perf_counter()
        cd = fb.FESpace(ft, grid)["CellDofs"]
        res["host_celldofs_ms"] = (time.perf_counter() - t0) * 1e3
        res["celldofs_equal"] = bool(np.array_equal(cd, ctx.space_get_celldofs(0, 10)))
    ctx.close()
    return res


def c3_multi_gpu(m, layers, steps, warmup):
    """C3 on z-slabs, one rank per GPU (launch with torchrun): weak scaling of the H1P2{1,3} stiffness + rhs assembly with
    the NCCL exchange of the interface columns (vertex and edge dofs in the cut planes)."""
    import torch
    import torch.distributed as dist
    from femb_b200 import partition as pt

    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    t0 = time.time()
    part, grid = pt.slab_partition_3d(m, rank, world, layers)
    grid.instantiate_on_device(local, components=("edges",))
    fes = fb.FESpace(fb.H1P2(1, 3), grid)
    cd = fes["CellDofs"]
    ndofs_own_cells = fes.ndofs
    if world > 1:
        sent0 = pt.allgather(pt.ghost_pattern_from_cells(part, cd), world)
        nhalo = pt.add_halo(part, sent0)
        fes.ndofs += nhalo  # halo rows: dofs of the neighbour's first cell layer
        er, ec = pt.extra_entries(part, sent0)
    qf = fb.QuadratureRule(grid.geometry, 2)
    S = _Session([fes], qf, local)
    ctx = S.ctx
    eg, ei = S.evaluator(0, "Gradient"), S.evaluator(0, "Identity")
    if world > 1 and er.size:
        ctx.pattern_extra(er, ec)
    nnz = ctx.pattern_build([(0, 0)])
    xbytes = 0
    if world > 1:
        uid = [_lib.Context.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(world, rank, uid[0])
        colptr = ctx.pattern_get_colptr(fes.ndofs)
        sent = pt.allgather(pt.ghost_columns(part, colptr, fetch=ctx.pattern_get_slots), world)
        plan = pt.exchange_plan(part, colptr, None, sent, fetch=ctx.pattern_get_slots)
        ctx.comm_set_exchange(plan)
        xbytes = 8 * int(plan.send_ptr[-1] + plan.sendb_ptr[-1] + plan.recv_ptr[-1] + plan.recvb_ptr[-1])
    setup_s = time.time() - t0
    p = AffineFunction([0.25], [[1.0, -1.0, 0.5]]).params()

    def step():
        ctx.matrix_zero()
        ctx.vector_zero()
        ctx.assemble_poisson(eg._eid, ei._eid, 1.0, _lib.RHS_AFFINE, p, 0.0)
        if world > 1:
            ctx.exchange()

    for _ in range(warmup):
        step()
    if world > 1:
        dist.barrier()
    ctx.sync()
    ctx.region_begin()
    for _ in range(steps):
        step()
    ms = ctx.region_end() / steps
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    # check: every owned column of the stiffness matrix sums to zero once the interface contributions have arrived
    nz = ctx.matrix_get()
    colptr = ctx.pattern_get_colptr(fes.ndofs)
    lens = np.diff(colptr)
    colsum = np.zeros(fes.ndofs)
    nzc = np.nonzero(lens)[0]
    colsum[nzc] = np.add.reduceat(nz, colptr[nzc] - 1)
    owned = part.owner == rank
    scale = np.abs(nz).max()
    res = {"config": f"C3 multi-GPU: H1P2{{1,3}} stiffness+rhs, Kuhn {m}x{m}x{layers} per rank, z-slabs", "n_gpus": world, "ncells_per_rank": grid.ncells,
           "ndofs_local": fes.ndofs, "halo_dofs": fes.ndofs - ndofs_own_cells, "nnz_local": nnz, "ms_per_step": ms,
           "cells_per_s": world * grid.ncells / (ms * 1e-3), "exchange_bytes_rank": xbytes, "setup_s": setup_s,
           "max_rel_owned_colsum": float(np.abs(colsum[owned]).max() / scale), "max_rel_ghost_colsum": float(np.abs(colsum[~owned]).max() / scale) if (~owned).any() else 0.0}
    if world > 1:
        allres = [None] * world
        dist.all_gather_object(allres, res)
        res["max_rel_owned_colsum"] = max(r["max_rel_owned_colsum"] for r in allres)
        dist.destroy_process_group()
    return res if rank == 0 else None


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--which", default="c3,c3p1,c4,c5")
    ap.add_argument("--m3", type=int, default=96)
    ap.add_argument("--m4", type=int, default=64)
    ap.add_argument("--n5", type=int, default=1024)
    ap.add_argument("--layers", type=int, default=None)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--e2e", action="store_true", help="C3: also time femb_assemble_poisson_host (host buffers, copies inside the timed region)")
    ap.add_argument("--scatter", type=int, default=None, help="C4: 0 atomic, 1 coloured, 2 gather (default policy of the library)")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    a = ap.parse_args()
    res = []
    for w in a.which.split(","):
        if w == "c3":
            res.append(c3_p2_3d(a.m3, a.steps, a.warmup, 2, a.check, a.e2e))
        elif w == "c3p1":
            res.append(c3_p2_3d(a.m3, a.steps, a.warmup, 1))
        elif w == "c4":
            res.append(c4_hdiv(a.m4, a.steps, a.warmup, "RT0", a.scatter))
            res.append(c4_hdiv(a.m4, a.steps, a.warmup, "BDM1", a.scatter))
        elif w == "c4rt0":
            res.append(c4_hdiv(a.m4, a.steps, a.warmup, "RT0", a.scatter))
        elif w == "c5":
            res += c5_taylor_hood(a.n5, a.steps, a.warmup)
        elif w == "topo":
            res.append(topo(a.m3))
        elif w == "c3mg":
            r = c3_multi_gpu(a.m3, a.layers, a.steps, a.warmup)
            if r:
                res.append(r)
    for r in res:
        print(json.dumps(r), flush=True)

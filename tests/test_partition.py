"""Multi-rank host logic on the CPU: slab partition, halo pattern entries, exchange plan (what femb_comm_set_exchange
consumes).  The numeric engine here is the numpy ORACLE (one local assembly per rank); the exchange is emulated
with gloo send/recv exactly as femb_exchange does it (pack -> send/recv -> add in ascending peer order).  The
distributed result must equal the single-domain oracle matrix / vector on every owned column."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

import femb_oracle as orc
import femb_b200 as fb
from femb_b200 import partition as pt


def _local_assembly(part, extra_rows=None, extra_cols=None):
    """local CSC (pattern = filtered local triplets + halo extras) with the rank's own contributions, and local b"""
    g = part.grid
    cn = g["CellNodes"]
    el = orc.Element("H1Pk", 1, 1)
    rhs = lambda x: x[..., 0] - x[..., 1]  # noqa: E731
    Aloc, bloc = orc.example200_local(el, "Triangle2D", g["Coordinates"], cn, g["CellVolumes"], 1.0, rhs)
    I, J, V = orc.example200_triplets(Aloc, cn)
    if extra_rows is not None and extra_rows.size:
        I, J, V = np.concatenate([I, extra_rows]), np.concatenate([J, extra_cols]), np.concatenate([V, np.zeros(extra_rows.shape[0])])
    colptr, rowval, nzval = orc.flush_csc(g.nnodes, g.nnodes, I, J, V)
    b = np.zeros(g.nnodes)
    np.add.at(b, cn.astype(np.int64).reshape(-1) - 1, bloc.reshape(-1))
    return colptr, rowval, nzval, b


def _global_reference(n, nranks, rows):
    X = np.linspace(0, 1, n + 1)
    Y = np.arange(nranks * rows + 1) / float(n)
    grid = fb.simplexgrid(X, Y)
    el = orc.Element("H1Pk", 1, 1)
    cn = grid["CellNodes"]
    colptr, rowval, nzval, b = orc.example200_assemble(el, "Triangle2D", grid["Coordinates"], cn, grid["CellVolumes"], cn, grid.nnodes, 1.0,
                                                       lambda x: x[..., 0] - x[..., 1])
    return sp.csc_matrix((nzval, rowval - 1, colptr - 1), shape=(grid.nnodes,) * 2), b


def _check_owned(part, colptr, rowval, nzval, b, Aref, bref):
    owned = np.nonzero(part.owner == part.rank)[0]
    A = sp.csc_matrix((nzval, rowval - 1, colptr - 1), shape=(part.l2g.shape[0],) * 2)
    for c in owned:
        col = A.getcol(c).tocoo()
        ref = Aref.getcol(int(part.l2g[c])).tocoo()
        got = dict(zip(part.l2g[col.row].tolist(), col.data.tolist()))
        want = dict(zip(ref.row.tolist(), ref.data.tolist()))
        got = {k: v for k, v in got.items() if v != 0.0 or k in want}
        assert set(got) == set(want), (part.rank, c)
        for k in want:
            assert abs(got[k] - want[k]) <= 1e-13 * max(1.0, abs(want[k]))
    assert np.abs(b[owned] - bref[part.l2g[owned]]).max() <= 1e-13 * np.abs(bref).max()


@pytest.mark.parametrize("nranks", [2, 3])
def test_slab_exchange_in_process(nranks):
    """all ranks emulated in one process (no collectives): plan consistency + numerical equality"""
    n, rows = 6, 4
    parts = [pt.slab_partition_2d(n, r, nranks, rows) for r in range(nranks)]
    first = [_local_assembly(p) for p in parts]
    sent = [pt.ghost_columns(p, f[0], f[1]) for p, f in zip(parts, first)]
    extras = [pt.extra_entries(p, sent) for p in parts]
    loc = [_local_assembly(p, *e) for p, e in zip(parts, extras)]
    sent = [pt.ghost_columns(p, l[0], l[1]) for p, l in zip(parts, loc)]
    plans = [pt.exchange_plan(p, l[0], l[1], sent) for p, l in zip(parts, loc)]
    # emulate femb_exchange: pack everything first, then add in ascending peer order
    packed = [{} for _ in range(nranks)]
    for r, (pl, l) in enumerate(zip(plans, loc)):
        for i, p in enumerate(pl.peers):
            packed[r][int(p)] = (l[2][pl.send_slots[pl.send_ptr[i]:pl.send_ptr[i + 1]]].copy(), l[3][pl.send_b[pl.sendb_ptr[i]:pl.sendb_ptr[i + 1]]].copy())
    for r, (pl, l) in enumerate(zip(plans, loc)):
        for i, p in enumerate(pl.peers):
            vals, bv = packed[int(p)][r]
            rs = pl.recv_slots[pl.recv_ptr[i]:pl.recv_ptr[i + 1]]
            assert vals.shape[0] == rs.shape[0]
            l[2][rs] += vals
            l[3][pl.recv_b[pl.recvb_ptr[i]:pl.recvb_ptr[i + 1]]] += bv
    Aref, bref = _global_reference(n, nranks, rows)
    for p, l in zip(parts, loc):
        _check_owned(p, *l, Aref, bref)
    # every global dof is owned exactly once
    owned = np.concatenate([p.l2g[p.owner == p.rank] for p in parts])
    assert np.array_equal(np.sort(owned), np.arange((n + 1) * (nranks * rows + 1)))


def _gloo_worker(rank, world, port, n, rows, q):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        part = pt.slab_partition_2d(n, rank, world, rows)
        c0, r0, _, _ = _local_assembly(part)
        sent = pt.allgather(pt.ghost_columns(part, c0, r0), world)
        colptr, rowval, nzval, b = _local_assembly(part, *pt.extra_entries(part, sent))
        sent = pt.allgather(pt.ghost_columns(part, colptr, rowval), world)
        plan = pt.exchange_plan(part, colptr, rowval, sent)
        sendbuf = [(torch.from_numpy(nzval[plan.send_slots[plan.send_ptr[i]:plan.send_ptr[i + 1]]].copy()),
                    torch.from_numpy(b[plan.send_b[plan.sendb_ptr[i]:plan.sendb_ptr[i + 1]]].copy())) for i in range(len(plan.peers))]
        recvbuf = [(torch.zeros(int(plan.recv_ptr[i + 1] - plan.recv_ptr[i]), dtype=torch.float64),
                    torch.zeros(int(plan.recvb_ptr[i + 1] - plan.recvb_ptr[i]), dtype=torch.float64)) for i in range(len(plan.peers))]
        ops = []
        for i, p in enumerate(plan.peers):
            for k in range(2):
                if sendbuf[i][k].numel():
                    ops.append(dist.P2POp(dist.isend, sendbuf[i][k], int(p)))
                if recvbuf[i][k].numel():
                    ops.append(dist.P2POp(dist.irecv, recvbuf[i][k], int(p)))
        for w in (dist.batch_isend_irecv(ops) if ops else []):
            w.wait()
        for i in range(len(plan.peers)):
            nzval[plan.recv_slots[plan.recv_ptr[i]:plan.recv_ptr[i + 1]]] += recvbuf[i][0].numpy()
            b[plan.recv_b[plan.recvb_ptr[i]:plan.recvb_ptr[i + 1]]] += recvbuf[i][1].numpy()
        Aref, bref = _global_reference(n, world, rows)
        _check_owned(part, colptr, rowval, nzval, b, Aref, bref)
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback

        q.put((rank, traceback.format_exc()))
        raise e
    finally:
        dist.destroy_process_group()


def test_slab_exchange_gloo_world2():
    """world_size 2 over gloo (CPU): the collectives of the plan builder + a batched isend/irecv exchange"""
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, 5, 3, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res


def _p2_local(part, grid, ndofs_local, extra=None):
    el = orc.Element("H1P2", 1)
    fes = fb.FESpace(fb.H1P2(1, 3), grid)
    cd = fes["CellDofs"]
    rhs = lambda x: 0.25 + x[..., 0] - x[..., 1] + 0.5 * x[..., 2]  # noqa: E731
    Aloc, bloc = orc.example200_local(el, "Tetrahedron3D", grid["Coordinates"], grid["CellNodes"], grid["CellVolumes"], 1.0, rhs)
    I, J, V = orc.example200_triplets(Aloc, cd, drop_tol=-1.0)  # structural pattern (no value filter)
    if extra is not None and extra[0].size:
        I, J, V = np.concatenate([I, extra[0]]), np.concatenate([J, extra[1]]), np.concatenate([V, np.zeros(extra[0].shape[0])])
    colptr, rowval, nzval = orc.flush_csc(ndofs_local, ndofs_local, I, J, V)
    b = np.zeros(ndofs_local)
    np.add.at(b, cd.astype(np.int64).reshape(-1) - 1, bloc.reshape(-1))
    return colptr, rowval, nzval, b, cd


@pytest.mark.parametrize("nranks", [2, 3])
def test_slab_exchange_3d_p2_in_process(nranks):
    """H1P2{1,3} on z-slabs of the Kuhn mesh: ghost pattern from CellDofs on the host, halo dofs discovered from the
    peers, exchange plan through the fetch interface (what the device path uses); result == single-domain oracle."""
    m, L = 3, 2
    parts, grids = zip(*[pt.slab_partition_3d(m, r, nranks, L) for r in range(nranks)])
    fess = [fb.FESpace(fb.H1P2(1, 3), g) for g in grids]
    sent0 = [pt.ghost_pattern_from_cells(p, f["CellDofs"]) for p, f in zip(parts, fess)]
    nhalo = [pt.add_halo(p, sent0) for p in parts]
    assert nhalo[0] > 0 and nhalo[-1] == 0  # the top rank has no upper neighbour
    extras = [pt.extra_entries(p, sent0) for p in parts]
    loc = [_p2_local(p, g, f.ndofs + h, e) for p, g, f, h, e in zip(parts, grids, fess, nhalo, extras)]
    sent = [pt.ghost_columns(p, l[0], fetch=lambda sl, rv=l[1]: rv[sl]) for p, l in zip(parts, loc)]
    plans = [pt.exchange_plan(p, l[0], None, sent, fetch=lambda sl, rv=l[1]: rv[sl]) for p, l in zip(parts, loc)]
    packed = [{} for _ in range(nranks)]
    for r, (pl, l) in enumerate(zip(plans, loc)):
        for i, p in enumerate(pl.peers):
            packed[r][int(p)] = (l[2][pl.send_slots[pl.send_ptr[i]:pl.send_ptr[i + 1]]].copy(), l[3][pl.send_b[pl.sendb_ptr[i]:pl.sendb_ptr[i + 1]]].copy())
    for r, (pl, l) in enumerate(zip(plans, loc)):
        for i, p in enumerate(pl.peers):
            vals, bv = packed[int(p)][r]
            l[2][pl.recv_slots[pl.recv_ptr[i]:pl.recv_ptr[i + 1]]] += vals
            l[3][pl.recv_b[pl.recvb_ptr[i]:pl.recvb_ptr[i + 1]]] += bv
    # single-domain reference with the same global dof keys
    X = np.linspace(0, 1, m + 1)
    gridG = fb.simplexgrid(X, X, np.arange(nranks * L + 1) / float(m))
    fesG = fb.FESpace(fb.H1P2(1, 3), gridG)
    cdG = fesG["CellDofs"]
    elG = orc.Element("H1P2", 1)
    Aloc, bloc = orc.example200_local(elG, "Tetrahedron3D", gridG["Coordinates"], gridG["CellNodes"], gridG["CellVolumes"], 1.0,
                                      lambda x: 0.25 + x[..., 0] - x[..., 1] + 0.5 * x[..., 2])
    I, J, V = orc.example200_triplets(Aloc, cdG, drop_tol=-1.0)
    cG, rG, vG = orc.flush_csc(fesG.ndofs, fesG.ndofs, I, J, V)
    Aref = sp.csc_matrix((vG, rG - 1, cG - 1), shape=(fesG.ndofs,) * 2)
    bref = np.zeros(fesG.ndofs)
    np.add.at(bref, cdG.astype(np.int64).reshape(-1) - 1, bloc.reshape(-1))
    NN = gridG.nnodes
    enG = gridG["EdgeNodes"].astype(np.int64) - 1
    keyG = np.concatenate([np.arange(NN, dtype=np.int64), NN + np.minimum(enG[:, 0], enG[:, 1]) * NN + np.maximum(enG[:, 0], enG[:, 1])])
    order = np.argsort(keyG)
    to_ref = lambda keys: order[np.searchsorted(keyG[order], keys)]  # noqa: E731
    nowned = 0
    for p, l in zip(parts, loc):
        colptr, rowval, nzval, b, _ = l
        A = sp.csc_matrix((nzval, rowval - 1, colptr - 1), shape=(p.l2g.shape[0],) * 2)
        owned = np.nonzero(p.owner == p.rank)[0]
        nowned += owned.shape[0]
        gi = to_ref(p.l2g)
        for c in owned:
            col = A.getcol(c).tocoo()
            ref = Aref.getcol(int(gi[c])).tocoo()
            got = dict(zip(gi[col.row].tolist(), col.data.tolist()))
            want = dict(zip(ref.row.tolist(), ref.data.tolist()))
            assert set(got) == set(want), (p.rank, c)
            for k in want:
                assert abs(got[k] - want[k]) <= 1e-13 * max(1.0, abs(want[k]))
        assert np.abs(b[owned] - bref[gi[owned]]).max() <= 1e-13 * np.abs(bref).max()
    assert nowned == fesG.ndofs

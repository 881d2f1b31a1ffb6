"""Cell partitioning across GPUs and the interface-exchange plan (host side of femb_comm_* / femb_exchange).

The reference is single-process; this is the B200-side extension the north star asks for: the mesh
is partitioned by cells, every rank assembles its own cells into a LOCAL CSC (local dof numbering), a dof on a cut
is owned by exactly one rank, and the non-owners' contributions to its matrix column / vector entry are sent to
the owner (NCCL send/recv of packed values, femb_exchange) and added in ascending peer order.

Local dof sets: the dofs of the rank's own cells (owned or ghost) plus *halo* dofs — rows that a peer contributes
to columns this rank owns (they have no local cell; their pattern entries come from `extra_entries`).
Everything here is O(interface) host work done once per mesh.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .grids import ExtendableGrid


@dataclass
class LocalPart:
    grid: ExtendableGrid   # local nodes (own cells' nodes + halo nodes) and local cells, local 1-based numbering
    l2g: np.ndarray        # int64 [nloc]: global (0-based) id of every local dof
    owner: np.ndarray      # int32 [nloc]: owning rank of every local dof
    rank: int
    nranks: int
    row_nodes: int = 0     # structured slabs: nodes per mesh row (for diagnostics)


@dataclass
class ExchangePlan:
    peers: np.ndarray
    send_ptr: np.ndarray
    send_slots: np.ndarray
    recv_ptr: np.ndarray
    recv_slots: np.ndarray
    sendb_ptr: np.ndarray
    send_b: np.ndarray
    recvb_ptr: np.ndarray
    recv_b: np.ndarray


def slab_partition_2d(n: int, rank: int, nranks: int, rows_per_rank: int | None = None) -> LocalPart:
    """Rank `rank`'s slab of the global `simplexgrid` with `n` quads per row and `nranks * rows_per_rank` quad rows
    (weak scaling: every rank gets `rows_per_rank` rows, default n; global domain [0,1] x [0, nranks*rows/n]).
    Node rows shared by two slabs are owned by the lower rank; the upper rank's first interior node row is the
    lower rank's halo."""
    m = n if rows_per_rank is None else rows_per_rank
    y0, y1 = rank * m, (rank + 1) * m
    top_halo = 1 if rank < nranks - 1 else 0
    nx = n + 1
    rows = np.arange(y0, y1 + 1 + top_halo, dtype=np.int64)
    X = np.linspace(0.0, 1.0, nx)
    Y = rows / float(n)
    coords = np.empty((nx * rows.shape[0], 2))
    coords[:, 0] = np.tile(X, rows.shape[0])
    coords[:, 1] = np.repeat(Y, nx)
    ix = np.arange(1, nx, dtype=np.int64)
    iy = np.arange(1, m + 1, dtype=np.int64)
    ip = (ix[None, :] + (iy[:, None] - 1) * nx).reshape(-1)
    p00, p10, p01, p11 = ip, ip + 1, ip + nx, ip + nx + 1
    cn = np.empty((ip.shape[0], 2, 3), dtype=np.int32)
    cn[:, 0, 0], cn[:, 0, 1], cn[:, 0, 2] = p00, p10, p11   # same split as grids.simplexgrid
    cn[:, 1, 0], cn[:, 1, 1], cn[:, 1, 2] = p11, p01, p00
    grid = ExtendableGrid(coords, cn.reshape(-1, 3), "Triangle2D")
    l2g = (np.arange(nx, dtype=np.int64)[None, :] + rows[:, None] * nx).reshape(-1)
    own_row = np.full(rows.shape[0], rank, dtype=np.int32)
    if rank > 0:
        own_row[0] = rank - 1
    if top_halo:
        own_row[-1] = rank + 1
    owner = np.repeat(own_row, nx).astype(np.int32)
    return LocalPart(grid, l2g, owner, rank, nranks, nx)


# ---------------------------------------------------------------------------------------------
def column_slots(colptr: np.ndarray, cols: np.ndarray):
    """0-based nz positions of the given 0-based columns (concatenated, column-major) and the column lengths."""
    lens = (colptr[cols + 1] - colptr[cols]).astype(np.int64)
    if cols.size == 0 or lens.sum() == 0:
        return np.zeros(0, np.int64), lens
    starts = colptr[cols] - 1
    off = np.cumsum(lens) - lens
    slots = np.arange(lens.sum(), dtype=np.int64) - np.repeat(off, lens) + np.repeat(starts, lens)
    return slots, lens


def ghost_columns(part: LocalPart, colptr: np.ndarray, rowval=None, fetch=None):
    """What this rank sends: per owner rank p, the slots of its non-empty ghost columns and their global (row, col),
    plus the ghost dofs themselves (vector entries).  -> {p: dict(slots, grow, gcol, bdof, gb)}.
    `rowval` is the full local pattern, or `fetch(slots) -> rowval[slots]` reads only the interface columns from
    the device (femb_pattern_get_slots)."""
    if fetch is None:
        fetch = lambda sl: rowval[sl]  # noqa: E731
    out = {}
    lens = np.diff(colptr)
    ghosts = np.nonzero((part.owner != part.rank) & (lens > 0))[0]
    for p in np.unique(part.owner[ghosts]):
        g = ghosts[part.owner[ghosts] == p]
        slots, glens = column_slots(colptr, g)
        cols = np.repeat(g, glens)
        out[int(p)] = dict(slots=slots, grow=part.l2g[fetch(slots) - 1], gcol=part.l2g[cols], bdof=g.astype(np.int64), gb=part.l2g[g])
    return out


def ghost_pattern_from_cells(part: LocalPart, celldofs: np.ndarray):
    """Structural ghost-column pattern straight from CellDofs on the host (before any device work): for every ghost
    dof g the rows = all dofs of the local cells containing g.  Same dict layout as ghost_columns (no slots)."""
    out = {}
    is_ghost = part.owner != part.rank
    cd = celldofs.astype(np.int64) - 1
    cells = np.nonzero(is_ghost[cd].any(axis=1))[0]
    if cells.size == 0:
        return out
    sub = cd[cells]  # (n, nd)
    nd = sub.shape[1]
    rows = np.repeat(sub[:, :, None], nd, axis=2).reshape(-1)
    cols = np.repeat(sub[:, None, :], nd, axis=1).reshape(-1)
    keep = is_ghost[cols]
    rows, cols = rows[keep], cols[keep]
    key = np.unique(cols * part.l2g.shape[0] + rows)
    cols, rows = key // part.l2g.shape[0], key % part.l2g.shape[0]
    for p in np.unique(part.owner[cols]):
        m = part.owner[cols] == p
        g = np.unique(cols[m])
        out[int(p)] = dict(slots=None, grow=part.l2g[rows[m]], gcol=part.l2g[cols[m]], bdof=g.astype(np.int64), gb=part.l2g[g])
    return out


def add_halo(part: LocalPart, sent_by_all: list) -> int:
    """Rows that peers contribute to columns owned here but that are not local dofs yet become halo dofs, appended to
    the local numbering (owner = the sending rank).  Returns how many were added (the caller bumps FESpace.ndofs)."""
    new, own = [], []
    known = np.sort(part.l2g)
    for r, sent in enumerate(sent_by_all):
        if r == part.rank or part.rank not in sent:
            continue
        g = np.unique(sent[part.rank]["grow"])
        pos = np.minimum(np.searchsorted(known, g), known.shape[0] - 1)
        miss = g[known[pos] != g]
        new.append(miss)
        own.append(np.full(miss.shape[0], r, dtype=np.int32))
    if not new or sum(a.shape[0] for a in new) == 0:
        return 0
    new, own = np.concatenate(new), np.concatenate(own)
    new, first = np.unique(new, return_index=True)
    part.l2g = np.concatenate([part.l2g, new])
    part.owner = np.concatenate([part.owner, own[first]])
    return int(new.shape[0])


def _g2l(part: LocalPart, gids: np.ndarray) -> np.ndarray:
    order = np.argsort(part.l2g, kind="stable")
    pos = np.searchsorted(part.l2g[order], gids)
    pos = np.minimum(pos, order.shape[0] - 1)
    loc = order[pos]
    if not np.array_equal(part.l2g[loc], gids):
        raise ValueError("a peer contributes to a dof that is not in this rank's local (own + halo) dof set")
    return loc


def extra_entries(part: LocalPart, sent_by_all: list):
    """Pattern entries (1-based local row, col) that peers will contribute to columns owned here.
    `sent_by_all[r]` = ghost_columns(...) of rank r (e.g. from dist.all_gather_object)."""
    rows, cols = [], []
    for r, sent in enumerate(sent_by_all):
        if r == part.rank or part.rank not in sent:
            continue
        s = sent[part.rank]
        rows.append(_g2l(part, s["grow"]) + 1)
        cols.append(_g2l(part, s["gcol"]) + 1)
    if not rows:
        return np.zeros(0, np.int64), np.zeros(0, np.int64)
    return np.concatenate(rows).astype(np.int64), np.concatenate(cols).astype(np.int64)


def find_slots(colptr, rowval, lrow, lcol):
    """0-based nz position of 0-based (lrow, lcol) in a Julia-style CSC; raises if absent."""
    n_rows = int(rowval.max()) + 1 if rowval.size else 1
    cols = np.repeat(np.arange(colptr.shape[0] - 1, dtype=np.int64), np.diff(colptr))
    key_all = cols * n_rows + (rowval - 1)
    key = lcol.astype(np.int64) * n_rows + lrow.astype(np.int64)
    pos = np.searchsorted(key_all, key)
    pos = np.minimum(pos, max(key_all.shape[0] - 1, 0))
    if key.size and (key_all.size == 0 or not np.array_equal(key_all[pos], key)):
        raise ValueError("a received contribution has no slot in the local pattern (missing femb_pattern_extra?)")
    return pos.astype(np.int64)


def exchange_plan(part: LocalPart, colptr, rowval, sent_by_all: list, fetch=None) -> ExchangePlan:
    """Slot lists for femb_comm_set_exchange from the CURRENT local pattern; both sides list the same entries in
    the same order (the sender's column-major slot order).  `rowval` full pattern or `fetch(slots)` (device)."""
    mine = sent_by_all[part.rank]
    recv_from = {r: sent[part.rank] for r, sent in enumerate(sent_by_all) if r != part.rank and part.rank in sent}
    peers = sorted(set(mine.keys()) | set(recv_from.keys()))
    sp, ss, rp, rs, sbp, sb, rbp, rb = [0], [], [0], [], [0], [], [0], []
    for p in peers:
        if p in mine:
            ss.append(mine[p]["slots"])
            sb.append(mine[p]["bdof"])
        if p in recv_from:
            s = recv_from[p]
            lrow, lcol = _g2l(part, s["grow"]), _g2l(part, s["gcol"])
            if fetch is None:
                rs.append(find_slots(colptr, rowval, lrow, lcol))
            else:  # only the columns that receive something are read back from the device
                ucols = np.unique(lcol)
                slots, lens = column_slots(colptr, ucols)
                rv = fetch(slots) - 1
                key_all = np.repeat(ucols, lens) * part.l2g.shape[0] + rv  # ascending: columns ascending, rows ascending within
                key = lcol.astype(np.int64) * part.l2g.shape[0] + lrow
                pos = np.minimum(np.searchsorted(key_all, key), max(key_all.shape[0] - 1, 0))
                if key.size and (key_all.size == 0 or not np.array_equal(key_all[pos], key)):
                    raise ValueError("a received contribution has no slot in the local pattern (missing femb_pattern_extra?)")
                rs.append(slots[pos])
            rb.append(_g2l(part, s["gb"]).astype(np.int64))
        sp.append(sum(a.shape[0] for a in ss))
        rp.append(sum(a.shape[0] for a in rs))
        sbp.append(sum(a.shape[0] for a in sb))
        rbp.append(sum(a.shape[0] for a in rb))
    cat = lambda xs: np.concatenate(xs).astype(np.int64) if xs else np.zeros(0, np.int64)  # noqa: E731
    i64 = lambda x: np.asarray(x, dtype=np.int64)  # noqa: E731
    return ExchangePlan(np.asarray(peers, dtype=np.int32), i64(sp), cat(ss), i64(rp), cat(rs), i64(sbp), cat(sb), i64(rbp), cat(rb))


def slab_partition_3d(m: int, rank: int, nranks: int, layers: int | None = None, order: int = 2, device: int | None = None):
    """Rank `rank`'s z-slab of the Kuhn-split cube grid (m x m quads per layer, `layers` cube layers per rank, default m)
    with the dofs of H1P1 / H1P2{1,3}: nodes, then (order 2) edges, local numbering of the local grid.  Global dof ids:
    node id, or NN + min_node * NN + max_node for an edge (unique, not dense).  Dofs lying entirely in the bottom plane
    of a slab are owned by the rank below.  Halo dofs are discovered from the peers (add_halo).  `device`: enumerate the
    edges on that GPU (femb_grid_entities) instead of in numpy.  -> (LocalPart, grid)"""
    from .grids import simplexgrid

    L = m if layers is None else layers
    z0 = rank * L
    X = np.linspace(0.0, 1.0, m + 1)
    Z = np.arange(z0, z0 + L + 1) / float(m)
    grid = simplexgrid(X, X, Z)
    if device is not None and order == 2:
        grid.instantiate_on_device(device, components=("edges",))  # xgrid[EdgeNodes] from the device enumeration, not the numpy one
    nxy = (m + 1) * (m + 1)
    NN = nxy * (nranks * L + 1)
    gnode = np.arange(grid.nnodes, dtype=np.int64) + z0 * nxy  # local lexicographic numbering is the global one shifted
    bottom = np.arange(grid.nnodes) < nxy
    l2g, owner = [gnode], [np.where(bottom & (rank > 0), rank - 1, rank).astype(np.int32)]
    if order == 2:
        en = grid["EdgeNodes"].astype(np.int64) - 1
        ga, gb = gnode[en[:, 0]], gnode[en[:, 1]]
        l2g.append(NN + np.minimum(ga, gb) * NN + np.maximum(ga, gb))
        eb = bottom[en[:, 0]] & bottom[en[:, 1]]
        owner.append(np.where(eb & (rank > 0), rank - 1, rank).astype(np.int32))
    return LocalPart(grid, np.concatenate(l2g), np.concatenate(owner), rank, nranks, m + 1), grid


def allgather(obj, nranks):
    """dist.all_gather_object wrapper (works with gloo and nccl process groups)."""
    import torch.distributed as dist

    out = [None] * nranks
    dist.all_gather_object(out, obj)
    return out

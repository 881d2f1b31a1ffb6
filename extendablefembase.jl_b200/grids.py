"""Synthetic simplex grids + the grid components the assembly path consumes.

This is the host-side stand-in for the un-vendored dependency ExtendableGrids.jl
(SURVEY.md rows E1, Appendix A.5): in production these arrays arrive *as data*
from Julia (``xgrid[Coordinates]``, ``xgrid[CellNodes]``, ``xgrid[CellVolumes]``,
``xgrid[CellFaces]``, ``xgrid[CellFaceSigns]`` ... — call sites
``/root/reference/src/dofmaps.jl:448-458``, ``src/feevaluator.jl:94-97``,
``src/fedefs/hdiv_rt0.jl:115``, ``src/fedefs/hdiv_bdm1.jl:201,235``).  Here they
are synthesised so that tests and the benchmark have inputs of the right shape.

Conventions (all arrays are column-major ``k x nitems`` like Julia, stored as
numpy arrays of shape ``(nitems, k)`` C-contiguous, i.e. the same bytes):

* indices are **1-based** ``int32`` exactly as ExtendableGrids stores them;
* faces / edges are numbered in order of first appearance while looping cells
  ``1..ncells`` and local faces / edges ``1..k``;
* ``FaceNodes`` of a face = node order of the local face of the creating cell;
* ``CellFaceSigns`` = +1 for the creating cell, -1 for the neighbour;
* ``CellFaceOrientations`` in ``{1,2,3,4}`` (3-D): 1 = same node sequence as the
  global face, 2..4 = the three reversed sequences (see ``_face_orientation``).

[EXT — recollection of ExtendableGrids; see DESIGN.md "unpinned conventions"].
"""
from __future__ import annotations

import numpy as np

# local topology tables (1-based, as in ExtendableGrids / quoted in
# /root/reference/src/fedefs/h1_p3.jl:167-168 and hdiv_bdm1.jl:143,156,169,182)
LOCAL_CELLFACENODES = {
    "Triangle2D": np.array([[1, 2], [2, 3], [3, 1]], dtype=np.int32),
    "Tetrahedron3D": np.array([[1, 3, 2], [1, 2, 4], [2, 3, 4], [1, 4, 3]], dtype=np.int32),
}
LOCAL_CELLEDGENODES = {
    "Triangle2D": np.array([[1, 2], [2, 3], [3, 1]], dtype=np.int32),
    "Tetrahedron3D": np.array([[1, 2], [1, 3], [1, 4], [2, 3], [2, 4], [3, 4]], dtype=np.int32),
}


class ExtendableGrid:
    """Minimal mirror of ``ExtendableGrids.ExtendableGrid``: a dict of lazily
    instantiated components, addressed by name (``grid["CellNodes"]``)."""

    def __init__(self, coords: np.ndarray, cellnodes: np.ndarray, geometry: str):
        coords = np.ascontiguousarray(coords, dtype=np.float64)
        cellnodes = np.ascontiguousarray(cellnodes, dtype=np.int32)
        self.geometry = geometry
        self.dim = coords.shape[1]
        self._c = {"Coordinates": coords, "CellNodes": cellnodes}

    # -- component access ------------------------------------------------
    def __getitem__(self, name: str):
        if name not in self._c:
            self._instantiate(name)
        return self._c[name]

    def __setitem__(self, name: str, value):
        self._c[name] = value

    def __contains__(self, name):
        return name in self._c

    @property
    def nnodes(self) -> int:
        return self._c["Coordinates"].shape[0]

    @property
    def ncells(self) -> int:
        return self._c["CellNodes"].shape[0]

    @property
    def nfaces(self) -> int:
        return self["FaceNodes"].shape[0]

    @property
    def nedges(self) -> int:
        return self["EdgeNodes"].shape[0]

    def num_cells(self) -> int:
        return self.ncells

    # -- lazy components -------------------------------------------------
    def _instantiate(self, name: str):
        if name == "CellVolumes":
            self._c[name] = cell_volumes(self._c["Coordinates"], self._c["CellNodes"])
        elif name == "UniqueCellGeometries":
            self._c[name] = [self.geometry]
        elif name in ("FaceNodes", "CellFaces", "CellFaceSigns", "CellFaceOrientations", "FaceCells"):
            self._build_faces()
        elif name in ("EdgeNodes", "CellEdges", "CellEdgeSigns"):
            self._build_edges()
        elif name in ("BFaceFaces", "BFaceNodes"):
            fc = self["FaceCells"]
            bf = np.nonzero(fc[:, 1] == 0)[0].astype(np.int32) + 1
            self._c["BFaceFaces"] = bf
            self._c["BFaceNodes"] = self["FaceNodes"][bf - 1]
        else:
            raise KeyError(f"grid component {name!r} not available")

    def instantiate_on_device(self, device: int = 0, components=("faces", "edges")):
        """Enumerate faces / edges with the CUDA library (femb_grid_entities, csrc/femb_grid.cu) instead of the numpy
        routines below and cache the same grid components: the device counterpart of the reference's lazy
        ``xgrid[CellFaces]`` ... instantiation (timed at Example200_LowLevelPoisson.jl:55)."""
        from . import _lib

        ctx = _lib.Context(device)
        try:
            ctx.mesh_set(self._c["Coordinates"], self._c["CellNodes"], None)
            if "faces" in components or self.dim == 2:
                r = ctx.grid_entities_get(_lib.ENTITY_FACES)
                self._c["CellFaces"], self._c["FaceNodes"], self._c["CellFaceSigns"] = r["cell_entities"], r["entity_nodes"], r["signs"]
                self._c["CellFaceOrientations"], self._c["FaceCells"] = r["orientations"], r["facecells"]
                if self.dim == 2:
                    self._c["EdgeNodes"], self._c["CellEdges"], self._c["CellEdgeSigns"] = r["entity_nodes"], r["cell_entities"], r["signs"]
            if "edges" in components and self.dim == 3:
                r = ctx.grid_entities_get(_lib.ENTITY_EDGES)
                self._c["CellEdges"], self._c["EdgeNodes"], self._c["CellEdgeSigns"] = r["cell_entities"], r["entity_nodes"], r["signs"]
        finally:
            ctx.close()
        return self

    def _build_faces(self):
        lfn = LOCAL_CELLFACENODES[self.geometry] - 1
        cn = self._c["CellNodes"]
        nc, nf = cn.shape[0], lfn.shape[0]
        fnodes = cn[:, lfn]  # (nc, nf, k) node ids (1-based) in local face order
        flat = fnodes.reshape(nc * nf, -1)
        ids, first = _first_appearance_ids(np.sort(flat, axis=1))
        nfaces = first.shape[0]
        cellfaces = (ids + 1).astype(np.int32).reshape(nc, nf)
        facenodes = np.ascontiguousarray(flat[first]).astype(np.int32)
        creator = np.zeros(nc * nf, dtype=bool)
        creator[first] = True
        signs = np.where(creator, 1, -1).astype(np.int32).reshape(nc, nf)
        # FaceCells: (creating cell, neighbour or 0)
        facecells = np.zeros((nfaces, 2), dtype=np.int32)
        cell_of = (np.arange(nc * nf) // nf + 1).astype(np.int32)
        facecells[ids[creator], 0] = cell_of[creator]
        facecells[ids[~creator], 1] = cell_of[~creator]
        self._c["FaceNodes"] = facenodes
        self._c["CellFaces"] = cellfaces
        self._c["CellFaceSigns"] = signs
        self._c["FaceCells"] = facecells
        if self.dim == 3:
            gl = facenodes[ids]  # global node order of each (cell, local face)
            self._c["CellFaceOrientations"] = _face_orientation(flat, gl).reshape(nc, nf)
        else:
            # 2-D: orientation 1 for the creator, 2 for the neighbour (reversed edge)
            self._c["CellFaceOrientations"] = np.where(creator, 1, 2).astype(np.int32).reshape(nc, nf)

    def _build_edges(self):
        if self.dim == 2:
            self._build_faces()
            self._c["EdgeNodes"] = self._c["FaceNodes"]
            self._c["CellEdges"] = self._c["CellFaces"]
            self._c["CellEdgeSigns"] = self._c["CellFaceSigns"]
            return
        len_ = LOCAL_CELLEDGENODES[self.geometry] - 1
        cn = self._c["CellNodes"]
        nc, ne = cn.shape[0], len_.shape[0]
        enodes = cn[:, len_].reshape(nc * ne, 2)
        ids, first = _first_appearance_ids(np.sort(enodes, axis=1))
        self._c["CellEdges"] = (ids + 1).astype(np.int32).reshape(nc, ne)
        edgenodes = np.ascontiguousarray(enodes[first]).astype(np.int32)
        self._c["EdgeNodes"] = edgenodes
        same = enodes[:, 0] == edgenodes[ids, 0]
        self._c["CellEdgeSigns"] = np.where(same, 1, -1).astype(np.int32).reshape(nc, ne)


def _first_appearance_ids(sorted_rows: np.ndarray):
    """ids[i] = 0-based rank (by first appearance) of the entity in row i;
    first[id] = row that created it."""
    nrow, k = sorted_rows.shape
    nmax = int(sorted_rows.max()) + 1
    key = sorted_rows[:, 0].astype(np.int64)
    for j in range(1, k):
        if float(key.max() + 1) * nmax >= 2.0 ** 62:  # 3 nodes of a large mesh: compress the partial key to dense ids first
            key = np.unique(key, return_inverse=True)[1].astype(np.int64).reshape(-1)
        key = key * nmax + sorted_rows[:, j].astype(np.int64)
    _, first_idx, inv = np.unique(key, return_index=True, return_inverse=True)
    order = np.argsort(first_idx, kind="stable")  # unique-id -> appearance rank
    rank = np.empty_like(order)
    rank[order] = np.arange(order.shape[0])
    return rank[inv.reshape(-1)], first_idx[order]


def _face_orientation(local: np.ndarray, glob: np.ndarray) -> np.ndarray:
    """Orientation code of a local triangular face (node triple ``local`` in the
    cell's outward order) relative to the global face's node triple ``glob``.

    1: local == glob (creating cell).  The neighbour sees the face reversed:
    2: glob == (l3, l2, l1), 3: glob == (l2, l1, l3), 4: glob == (l1, l3, l2).
    The numbering of 2..4 is the ONLY one for which the subset/coefficient
    tables of HDIVBDM1{3} (/root/reference/src/fedefs/hdiv_bdm1.jl:215-249) give a
    basis dual to the face functionals of hdiv_bdm1.jl:47-54 (all 4 local faces
    agree); pinned by tests/test_oracle_pins.py (test_bdm1_orientation_codes) and
    the interpolation round trip there.
    """
    l1, l2, l3 = local[:, 0], local[:, 1], local[:, 2]
    g1, g2, g3 = glob[:, 0], glob[:, 1], glob[:, 2]
    out = np.zeros(local.shape[0], dtype=np.int32)
    out[(g1 == l1) & (g2 == l2) & (g3 == l3)] = 1
    out[(g1 == l3) & (g2 == l2) & (g3 == l1)] = 2
    out[(g1 == l2) & (g2 == l1) & (g3 == l3)] = 3
    out[(g1 == l1) & (g2 == l3) & (g3 == l2)] = 4
    if (out == 0).any():
        raise ValueError("inconsistent face orientation (non-manifold or same-handed neighbour faces)")
    return out


def cell_volumes(coords: np.ndarray, cellnodes: np.ndarray, chunk: int = 1 << 22) -> np.ndarray:
    """|T| per simplex cell = |det A| / d!  ([EXT E1] ``xgrid[CellVolumes]``); chunked so that a 64M-cell
    grid does not need multi-GB temporaries."""
    d = coords.shape[1]
    nc = cellnodes.shape[0]
    out = np.empty(nc, dtype=np.float64)
    for c0 in range(0, nc, chunk):
        x = coords[cellnodes[c0:c0 + chunk] - 1]  # (n, d+1, d)
        a = x[:, 1:, :] - x[:, :1, :]
        if d == 2:
            det = a[:, 0, 0] * a[:, 1, 1] - a[:, 0, 1] * a[:, 1, 0]
            out[c0:c0 + chunk] = np.abs(det) / 2.0
        else:
            det = (
                a[:, 0, 0] * (a[:, 1, 1] * a[:, 2, 2] - a[:, 1, 2] * a[:, 2, 1])
                - a[:, 0, 1] * (a[:, 1, 0] * a[:, 2, 2] - a[:, 1, 2] * a[:, 2, 0])
                + a[:, 0, 2] * (a[:, 1, 0] * a[:, 2, 1] - a[:, 1, 1] * a[:, 2, 0])
            )
            out[c0:c0 + chunk] = np.abs(det) / 6.0
    return out


# ----------------------------------------------------------------------
# generators
# ----------------------------------------------------------------------
def simplexgrid(X, Y=None, Z=None) -> ExtendableGrid:
    """Tensor-product simplex grid; restates ExtendableGrids ``simplexgrid(X,Y[,Z])``
    as used at /root/reference/examples/Example200_LowLevelPoisson.jl:53 and
    Example210_LowLevelNavierStokes.jl:75 ([EXT]).

    2-D: nodes ``ip = ix + (iy-1)*nx``; per quad two triangles ``(p00,p10,p11)``,
    ``(p11,p01,p00)``.  3-D: Kuhn/Freudenthal split of each cube into 6 positively
    oriented tetrahedra sharing the main diagonal.
    """
    X = np.asarray(X, dtype=np.float64)
    if Y is None:
        raise ValueError("1-D grids are outside the hot path")
    Y = np.asarray(Y, dtype=np.float64)
    nx, ny = X.shape[0], Y.shape[0]
    if Z is None:
        coords = np.empty((nx * ny, 2))
        coords[:, 0] = np.tile(X, ny)
        coords[:, 1] = np.repeat(Y, nx)
        ix = np.arange(1, nx, dtype=np.int64)
        iy = np.arange(1, ny, dtype=np.int64)
        ip = (ix[None, :] + (iy[:, None] - 1) * nx).reshape(-1)
        p00, p10, p01, p11 = ip, ip + 1, ip + nx, ip + nx + 1
        cn = np.empty((ip.shape[0], 2, 3), dtype=np.int32)
        cn[:, 0, 0], cn[:, 0, 1], cn[:, 0, 2] = p00, p10, p11
        cn[:, 1, 0], cn[:, 1, 1], cn[:, 1, 2] = p11, p01, p00
        return ExtendableGrid(coords, cn.reshape(-1, 3), "Triangle2D")
    Z = np.asarray(Z, dtype=np.float64)
    nz = Z.shape[0]
    coords = np.empty((nx * ny * nz, 3))
    coords[:, 0] = np.tile(X, ny * nz)
    coords[:, 1] = np.tile(np.repeat(Y, nx), nz)
    coords[:, 2] = np.repeat(Z, nx * ny)
    ix = np.arange(1, nx, dtype=np.int64)
    iy = np.arange(1, ny, dtype=np.int64)
    iz = np.arange(1, nz, dtype=np.int64)
    ip = (ix[None, None, :] + (iy[None, :, None] - 1) * nx + (iz[:, None, None] - 1) * nx * ny).reshape(-1)
    step = (1, nx, nx * ny)
    # Kuhn: one tet per permutation (a,b,c) of the axes: v0, v0+e_a, v0+e_a+e_b, v0+e_a+e_b+e_c;
    # odd permutations get two nodes swapped so that det(A) > 0 for all six.
    perms = [(0, 1, 2), (0, 2, 1), (1, 0, 2), (1, 2, 0), (2, 0, 1), (2, 1, 0)]
    cn = np.empty((ip.shape[0], 6, 4), dtype=np.int32)
    for t, (a, b, c) in enumerate(perms):
        v0, v1 = ip, ip + step[a]
        v2 = v1 + step[b]
        v3 = v2 + step[c]
        m = np.zeros((3, 3))
        m[a, 0] = 1
        m[a, 1] = m[b, 1] = 1
        m[:, 2] = 1
        if np.linalg.det(m) > 0:
            cn[:, t, 0], cn[:, t, 1], cn[:, t, 2], cn[:, t, 3] = v0, v1, v2, v3
        else:
            cn[:, t, 0], cn[:, t, 1], cn[:, t, 2], cn[:, t, 3] = v0, v2, v1, v3
    return ExtendableGrid(coords, cn.reshape(-1, 4), "Tetrahedron3D")


def grid_triangle(coords) -> ExtendableGrid:
    """Single triangle (cf. ``grid_triangle`` used at /root/reference/test/test_operators.jl:83)."""
    return ExtendableGrid(np.asarray(coords, dtype=np.float64).reshape(3, 2), np.array([[1, 2, 3]], dtype=np.int32), "Triangle2D")


def grid_tetrahedron(coords) -> ExtendableGrid:
    return ExtendableGrid(np.asarray(coords, dtype=np.float64).reshape(4, 3), np.array([[1, 2, 3, 4]], dtype=np.int32), "Tetrahedron3D")


def jitter(grid: ExtendableGrid, amplitude: float = 0.25, seed: int = 20261019) -> ExtendableGrid:
    """Perturb interior nodes of a tensor grid by U(-a*h, a*h) (SURVEY.md §8(d):
    removes exact zeros so the P1 pattern is the full 7-point stencil)."""
    c = grid["Coordinates"].copy()
    lo, hi = c.min(axis=0), c.max(axis=0)
    interior = np.all((c > lo + 1e-12) & (c < hi - 1e-12), axis=1)
    h = ((hi - lo).prod() / max(grid.ncells, 1) * (2 if grid.dim == 2 else 6)) ** (1.0 / grid.dim)
    rng = np.random.default_rng(seed)
    c[interior] += rng.uniform(-amplitude * h, amplitude * h, size=(int(interior.sum()), grid.dim))
    return ExtendableGrid(c, grid["CellNodes"], grid.geometry)


def slab(grid_n: int, rank: int, nranks: int):
    """Row range ``[y0, y1)`` of quads/cubes owned by ``rank`` when the slowest mesh
    index is split into contiguous slabs."""
    base, rem = divmod(grid_n, nranks)
    y0 = rank * base + min(rank, rem)
    return y0, y0 + base + (1 if rank < rem else 0)

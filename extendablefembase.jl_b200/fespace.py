"""FESpace and the CellDofs dofmap (host side).

Mirrors ``FESpace{FEType}(xgrid)`` (/root/reference/src/finiteelements.jl:91-100,
:193-248), ``count_ndofs`` (:395-449) and the pattern-driven CellDofs generator
``init_dofmap_from_pattern!`` (/root/reference/src/dofmaps.jl:263-301 parser,
:419-627 generator).  The CellDofs layout is an *input contract* of the CUDA
library: ``int32``, 1-based, cell-contiguous (``nd x ncells`` column-major), local
order = component-major, within a component the pattern segments in order.
"""
from __future__ import annotations

import re

import numpy as np

from .fedefs import FEType, _NEDGES, _NFACES, _NNODES

_SEG = re.compile(r"([NnFfEeIiCc])(\d+)")


def parse_pattern(pattern: str):
    """dofmaps.jl:284-301 -> list of (type_char_lower, each_component, ndofs)."""
    return [(m.group(1).lower(), m.group(1).isupper(), int(m.group(2))) for m in _SEG.finditer(pattern)]


class FESpace:
    """``fe_space["CellDofs"]`` builds the dofmap lazily like ``fe_space[CellDofs]``
    (finiteelements.jl:376-388)."""

    def __init__(self, fetype: FEType, xgrid, name: str | None = None):
        self.fetype = fetype
        self.xgrid = xgrid
        self.name = name or repr(fetype)
        self._maps: dict = {}
        self.ndofs, self.coffset = count_ndofs(xgrid, fetype)

    def eltype(self):
        return self.fetype

    def __getitem__(self, key: str):
        if key not in self._maps:
            if key != "CellDofs":
                raise KeyError(f"dofmap {key!r} is not on the assembly path")
            self._maps[key] = init_celldofs_from_pattern(self)
        return self._maps[key]

    def __repr__(self):
        return f"FESpace({self.fetype}, ndofs={self.ndofs})"


def count_ndofs(xgrid, fetype: FEType):
    """finiteelements.jl:395-449 (single cell geometry, unbroken)."""
    g = xgrid.geometry
    nc = fetype.ncomponents
    total = coffset = 0
    for t, each, q in parse_pattern(fetype.get_dofmap_pattern(g)):
        mult = nc if each else 1
        if t == "n":
            n = xgrid.nnodes
        elif t == "f":
            n = xgrid.nfaces
        elif t == "e":
            n = xgrid.nedges
        elif t in ("i", "c"):
            n = xgrid.ncells
        total += n * q * mult
        if each:
            coffset += n * q
    return total, coffset


def init_celldofs_from_pattern(fes: FESpace) -> np.ndarray:
    """dofmaps.jl:419-627 restated with array operations.  Returns ``(ncells, nd)``
    int32 (same bytes as Julia's ``nd x ncells`` colentries)."""
    xgrid, fetype = fes.xgrid, fes.fetype
    g = xgrid.geometry
    segs = parse_pattern(fetype.get_dofmap_pattern(g))
    ncomp = fetype.ncomponents
    nnodes, ncells = xgrid.nnodes, xgrid.ncells
    cells = np.arange(1, ncells + 1, dtype=np.int64)
    cols = []

    def emit(items, nitems, q, offset):
        # for n in items, for m in 1:q : items[n] + offset + (m-1)*nitems   (dofmaps.jl:527-534)
        it = items.astype(np.int64)
        block = it[:, :, None] + offset + np.arange(q, dtype=np.int64)[None, None, :] * nitems
        cols.append(block.reshape(ncells, -1))

    def run(each_component: bool, offset0: int):
        offset = offset0
        for t, each, q in segs:
            if each != each_component:
                continue
            if t == "n":
                emit(xgrid["CellNodes"], nnodes, q, offset)
                offset += nnodes * q
            elif t == "f":
                emit(xgrid["CellFaces"], xgrid.nfaces, q, offset)
                offset += xgrid.nfaces * q
            elif t == "e":
                emit(xgrid["CellEdges"], xgrid.nedges, q, offset)
                offset += xgrid.nedges * q
            elif t in ("i", "c"):
                # for m in 1:q: item + offset; offset += nitems   (dofmaps.jl:551-557)
                emit(cells[:, None], ncells, q, offset)
                offset += ncells * q

    for c in range(ncomp):
        run(True, c * fes.coffset)
    run(False, ncomp * fes.coffset)
    out = np.concatenate(cols, axis=1)
    assert out.shape[1] == fetype.get_ndofs(g), (out.shape, fetype.get_ndofs(g))
    assert out.max() <= fes.ndofs
    return np.ascontiguousarray(out, dtype=np.int32)


def boundarydofs(fes: FESpace) -> np.ndarray:
    """Sorted unique dofs on boundary faces (dofmaps.jl:871-899) for H1 elements:
    derived from the cells' local face -> local dof incidence."""
    xgrid, fetype = fes.xgrid, fes.fetype
    g = xgrid.geometry
    if fetype.family != "H1":
        raise NotImplementedError("boundarydofs is provided for H1 spaces (Dirichlet penalisation of Example200/210)")
    from .grids import LOCAL_CELLFACENODES, LOCAL_CELLEDGENODES

    celldofs = fes["CellDofs"]
    fc = xgrid["FaceCells"]
    bfaces = np.nonzero(fc[:, 1] == 0)[0]
    cf = xgrid["CellFaces"]
    cellof = fc[bfaces, 0] - 1
    lface = np.argmax(cf[cellof] == (bfaces[:, None] + 1), axis=1)
    nd_s = fetype.get_ndofs(g) // fetype.ncomponents
    segs = parse_pattern(fetype.get_dofmap_pattern(g))
    lfn = LOCAL_CELLFACENODES[g] - 1
    len_ = LOCAL_CELLEDGENODES[g] - 1
    nloc_faces = lfn.shape[0]
    # local scalar dofs attached to each local face
    local = [[] for _ in range(nloc_faces)]
    pos = 0
    for t, each, q in segs:
        assert each
        if t == "n":
            for f in range(nloc_faces):
                local[f] += [pos + n * q + m for n in lfn[f] for m in range(q)]
            pos += _NNODES[g] * q
        elif t == "f":
            for f in range(nloc_faces):
                local[f] += [pos + f * q + m for m in range(q)]
            pos += _NFACES[g] * q
        elif t == "e":
            for f in range(nloc_faces):
                fn = set(lfn[f].tolist())
                for e in range(len_.shape[0]):
                    if set(len_[e].tolist()) <= fn:
                        local[f] += [pos + e * q + m for m in range(q)]
            pos += _NEDGES[g] * q
        elif t == "i":
            pos += q
    out = []
    for f in range(nloc_faces):
        sel = lface == f
        if not sel.any():
            continue
        idx = np.array([c * nd_s + j for c in range(fetype.ncomponents) for j in local[f]], dtype=np.int64)
        out.append(celldofs[cellof[sel]][:, idx].reshape(-1))
    return np.unique(np.concatenate(out)).astype(np.int64)

"""Host grid components (grids.py): first-appearance numbering of faces / edges, also for node ids whose packed
3-tuple would overflow 64 bits (meshes beyond ~2M nodes)."""
import numpy as np

import femb_b200 as fb
from femb_b200 import grids


def _brute(rows):
    seen, ids, first = {}, [], []
    for i, r in enumerate(map(tuple, rows)):
        if r not in seen:
            seen[r] = len(seen)
            first.append(i)
        ids.append(seen[r])
    return np.array(ids), np.array(first)


def test_first_appearance_ids_small_and_huge_node_ids():
    rng = np.random.default_rng(5)
    for nmax in (50, 3_000_000_000 // 1000, 2_000_000_000):
        pool = np.sort(rng.integers(1, nmax, size=(40, 3)), axis=1)
        rows = pool[rng.integers(0, 40, size=300)]
        ids, first = grids._first_appearance_ids(rows)
        bi, bf = _brute(rows)
        assert np.array_equal(ids, bi) and np.array_equal(first, bf)


def test_face_and_edge_counts_euler():
    X = np.linspace(0, 1, 4)
    g = fb.simplexgrid(X, X, X)
    # Euler characteristic of a ball: V - E + F - C = 1
    assert g.nnodes - g.nedges + g.nfaces - g.ncells == 1
    g2 = fb.simplexgrid(X, X)
    assert g2.nnodes - g2.nfaces + g2.ncells == 1


# hand-derived topology of the unit square split like simplexgrid (cells (1,2,4), (4,3,1)): first-appearance numbering,
# FaceNodes in the creating cell's local order, +1 for the creator / -1 for the neighbour, FaceCells (creator, neighbour)
HAND_SQUARE = dict(
    cellnodes=np.array([[1, 2, 4], [4, 3, 1]], dtype=np.int32),
    CellFaces=np.array([[1, 2, 3], [4, 5, 3]]), FaceNodes=np.array([[1, 2], [2, 4], [4, 1], [4, 3], [3, 1]]),
    CellFaceSigns=np.array([[1, 1, 1], [1, 1, -1]]), FaceCells=np.array([[1, 0], [1, 0], [1, 2], [2, 0], [2, 0]]),
    CellFaceOrientations=np.array([[1, 1, 1], [1, 1, 2]]),
)
# two positively oriented tetrahedra glued along the face {2,3,4}: cell 1 = (1,2,3,4), cell 2 = (2,3,4,5)
HAND_TETS = dict(
    cellnodes=np.array([[1, 2, 3, 4], [2, 3, 4, 5]], dtype=np.int32),
    # local faces (1,3,2), (1,2,4), (2,3,4), (1,4,3):  cell 1 -> (1,3,2) (1,2,4) (2,3,4) (1,4,3);  cell 2 -> (2,4,3) (2,3,5) (3,4,5) (2,5,4)
    CellFaces=np.array([[1, 2, 3, 4], [3, 5, 6, 7]]),
    FaceNodes=np.array([[1, 3, 2], [1, 2, 4], [2, 3, 4], [1, 4, 3], [2, 3, 5], [3, 4, 5], [2, 5, 4]]),
    CellFaceSigns=np.array([[1, 1, 1, 1], [-1, 1, 1, 1]]),
    # the neighbour reads the glued face as (2,4,3) = (l1,l3,l2) of the global (2,3,4): orientation code 4 (hdiv_bdm1.jl:238-246)
    CellFaceOrientations=np.array([[1, 1, 1, 1], [4, 1, 1, 1]]),
    FaceCells=np.array([[1, 0], [1, 0], [1, 2], [1, 0], [2, 0], [2, 0], [2, 0]]),
    # local edges (1,2) (1,3) (1,4) (2,3) (2,4) (3,4): cell 2 -> (2,3) (2,4) (2,5) (3,4) (3,5) (4,5)
    CellEdges=np.array([[1, 2, 3, 4, 5, 6], [4, 5, 7, 6, 8, 9]]),
    EdgeNodes=np.array([[1, 2], [1, 3], [1, 4], [2, 3], [2, 4], [3, 4], [2, 5], [3, 5], [4, 5]]),
    CellEdgeSigns=np.array([[1, 1, 1, 1, 1, 1], [1, 1, 1, 1, 1, 1]]),
)


def hand_grids():
    sq = fb.ExtendableGrid(np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0], [1.0, 1.0]]), HAND_SQUARE["cellnodes"], "Triangle2D")
    tc = np.array([[0.0, 0.0, 0.0], [1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0], [1.0, 1.0, 1.0]])
    tets = fb.ExtendableGrid(tc, HAND_TETS["cellnodes"], "Tetrahedron3D")
    return (sq, HAND_SQUARE), (tets, HAND_TETS)


def test_hand_derived_topology_host():
    for grid, hand in hand_grids():
        for name, want in hand.items():
            if name != "cellnodes":
                assert np.array_equal(grid[name], want), name
    tets = hand_grids()[1][0]
    c = tets["Coordinates"][tets["CellNodes"] - 1]
    det = np.linalg.det(np.transpose(c[:, 1:, :] - c[:, :1, :], (0, 2, 1)))
    assert (det > 0).all()  # both cells positively oriented, as simplexgrid produces them


def test_hand_derived_boundarydofs():
    """boundarydofs (dofmaps.jl:871-899) on the two-triangle square: all four nodes; of the five faces only the diagonal
    (face 3) is interior.  P2: nodes + boundary faces; vector-valued: the same per component, shifted by coffset."""
    from femb_b200.fespace import boundarydofs

    sq = hand_grids()[0][0]
    assert boundarydofs(fb.FESpace(fb.H1Pk(1, 2, 1), sq)).tolist() == [1, 2, 3, 4]
    assert boundarydofs(fb.FESpace(fb.H1Pk(1, 2, 2), sq)).tolist() == [1, 2, 3, 4, 4 + 1, 4 + 2, 4 + 4, 4 + 5]
    fes2 = fb.FESpace(fb.H1Pk(2, 2, 2), sq)
    assert fes2.coffset == 9
    assert boundarydofs(fes2).tolist() == [1, 2, 3, 4, 5, 6, 8, 9] + [9 + d for d in (1, 2, 3, 4, 5, 6, 8, 9)]
    # P3: two dofs per face (F2) and one interior dof per cell (I1): ndofs = 4 + 2*5 + 2
    fes3 = fb.FESpace(fb.H1Pk(1, 2, 3), sq)
    assert fes3.ndofs == 16
    want = [1, 2, 3, 4] + sorted(4 + f + m * 5 for f in (1, 2, 4, 5) for m in (0, 1))
    assert boundarydofs(fes3).tolist() == want
    # two glued tetrahedra: every node and every edge lies on the boundary, only the glued face is interior
    tets = hand_grids()[1][0]
    assert boundarydofs(fb.FESpace(fb.H1P2(1, 3), tets)).tolist() == list(range(1, 5 + 9 + 1))


def test_host_topology_matches_the_oracle_loop():
    """grids.py (sort-based, vectorised) against the literal first-appearance loop of the oracle on shuffled meshes"""
    import femb_oracle as orc

    rng = np.random.default_rng(12)
    X = np.linspace(0, 1, 5)
    for grid in (fb.simplexgrid(X, X), fb.simplexgrid(X[:4], X[:4], X[:4])):
        cn = grid["CellNodes"][rng.permutation(grid.ncells)]
        g = fb.ExtendableGrid(grid["Coordinates"], cn, grid.geometry)
        f = orc.enumerate_entities(g.geometry, cn, "faces")
        assert np.array_equal(g["CellFaces"], f["cell_entities"]) and np.array_equal(g["FaceNodes"], f["entity_nodes"])
        assert np.array_equal(g["CellFaceSigns"], f["signs"]) and np.array_equal(g["FaceCells"], f["facecells"])
        assert np.array_equal(g["CellFaceOrientations"], f["orientations"]) and (f["orientations"] > 0).all()
        e = orc.enumerate_entities(g.geometry, cn, "edges")
        assert np.array_equal(g["CellEdges"], e["cell_entities"]) and np.array_equal(g["EdgeNodes"], e["entity_nodes"])
        assert np.array_equal(g["CellEdgeSigns"], e["signs"])
    for grid, hand in hand_grids():
        f = orc.enumerate_entities(grid.geometry, hand["cellnodes"], "faces")
        assert np.array_equal(f["cell_entities"], hand["CellFaces"]) and np.array_equal(f["entity_nodes"], hand["FaceNodes"])
        assert np.array_equal(f["orientations"], hand["CellFaceOrientations"])


def test_boundarydofs_geometric_criterion():
    """Independent check of boundarydofs: on the unit square / cube a P2 dof is a boundary dof iff its Lagrange point
    (node or edge midpoint) lies on the boundary of the domain."""
    from femb_b200.fespace import boundarydofs

    X = np.linspace(0, 1, 5)
    for grid in (fb.simplexgrid(X, X), fb.simplexgrid(X[:4] / X[3], X[:4] / X[3], X[:4] / X[3])):
        c = grid["Coordinates"]
        en = grid["EdgeNodes"]
        pts = np.concatenate([c, 0.5 * (c[en[:, 0] - 1] + c[en[:, 1] - 1])])
        on_bnd = ((np.abs(pts) < 1e-14) | (np.abs(pts - 1.0) < 1e-14)).any(axis=1)
        ft = fb.H1Pk(1, 2, 2) if grid.dim == 2 else fb.H1P2(1, 3)
        fes = fb.FESpace(ft, grid)
        assert fes.ndofs == pts.shape[0]
        assert np.array_equal(boundarydofs(fes), np.nonzero(on_bnd)[0] + 1)

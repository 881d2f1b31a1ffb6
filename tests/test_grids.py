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

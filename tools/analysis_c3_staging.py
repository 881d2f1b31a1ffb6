#!/usr/bin/env python
"""CPU-side feasibility numbers for staging the per-cell coefficient records of the P2-3D gather in shared memory
(DESIGN.md §9): for every CTA of the owner-computes gather (128 consecutive columns; vertex columns: 32 per CTA) count the
(column, cell) visits and the DISTINCT cells behind them.  visits / distinct = how often a staged record would be reused;
distinct * 80 B = the shared memory the staging needs."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import femb_b200 as fb  # noqa: E402  (host-side grid / dofmap generators only)


def main(m=24):
    X = np.linspace(0, 1, m + 1)
    grid = fb.simplexgrid(X, X, X)
    fes = fb.FESpace(fb.H1P2(1, 3), grid)
    cd = fes["CellDofs"].astype(np.int64) - 1  # (ncells, 10)
    ncells = cd.shape[0]
    dof = cd.reshape(-1)
    cell = np.repeat(np.arange(ncells), 10)
    order = np.argsort(dof, kind="stable")
    dof, cell = dof[order], cell[order]
    nn = grid.nnodes
    out = {}
    for name, lo, hi, cols in (("vertex columns (SPLIT 4, 32 per CTA)", 0, nn, 32), ("edge columns (SPLIT 1, 128 per CTA)", nn, fes.ndofs, 128)):
        sel = (dof >= lo) & (dof < hi)
        cta = (dof[sel] - lo) // cols
        c = cell[sel]
        visits = np.bincount(cta)
        pair = np.unique(cta * ncells + c)
        distinct = np.bincount(pair // ncells, minlength=visits.shape[0])
        out[name] = dict(ctas=int(visits.shape[0]), visits_per_cta=float(visits.mean()), distinct_cells_per_cta=float(distinct.mean()),
                         reuse=float(visits.sum() / distinct.sum()), smem_kb=float(distinct.mean() * 80 / 1024), smem_kb_max=float(distinct.max() * 80 / 1024))
    for k, v in out.items():
        print(k, {a: round(b, 2) if isinstance(b, float) else b for a, b in v.items()})


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 24)

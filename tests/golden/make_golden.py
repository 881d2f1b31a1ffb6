#!/usr/bin/env python
"""Generates tests/golden/*.npz.

Two kinds of fixtures:
  * hand_*   : values derived by hand from the reference's formulas (the non-reference triangle of
               /root/reference/test/test_operators.jl:83, the right-isosceles structured cell) — true known answers;
  * oracle_* : outputs of the pinned CPU oracle (oracle/femb_oracle.py) on tiny seeded inputs.  The reference itself
               cannot run in this image (pure Julia, no julia binary), so these are regression vectors for the oracle and
               the device path, NOT outputs of the reference.
Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import femb_b200 as fb  # noqa: E402  (host-side grid / dofmap generators only; no device work)
import femb_oracle as orc  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    # ---- hand-derived ------------------------------------------------------------------------
    np.savez(os.path.join(HERE, "hand_appendix_c.npz"),
             tri_coords=np.array([[-1.0, 0.0], [1.0, 0.0], [0.0, 1.0]]),
             tri_p1_stiffness=np.array([[0.5, 0, -0.5], [0, 0.5, -0.5], [-0.5, -0.5, 1.0]]),
             tri_jacobian=np.array([[2.0, 1.0], [0.0, 1.0]]), tri_det=2.0, tri_inv_t=np.array([[0.5, 0.0], [-0.5, 1.0]]),
             iso_p1_stiffness=0.5 * np.array([[1.0, -1, 0], [-1, 2, -1], [0, -1, 1]]),
             iso_p2_stiffness=np.array([[1 / 2, 1 / 6, 0, -2 / 3, 0, 0], [1 / 6, 1, 1 / 6, -2 / 3, -2 / 3, 0], [0, 1 / 6, 1 / 2, 0, -2 / 3, 0],
                                        [-2 / 3, -2 / 3, 0, 8 / 3, 0, -4 / 3], [0, -2 / 3, -2 / 3, 0, 8 / 3, -4 / 3], [0, 0, 0, -4 / 3, -4 / 3, 8 / 3]]),
             p2_curl2_at_midpoint=1.0 / 3.0, p2_tet_curl3=np.array([0.5 - 1.5, 0.0, 0.0]))
    # ---- oracle regression vectors -----------------------------------------------------------
    rhs = lambda x: x[..., 0] - x[..., 1]  # noqa: E731  Example200:38
    for tag, order, n, jit in (("p1_n3", 1, 3, False), ("p2_n3", 2, 3, False), ("p1_n4_jit", 1, 4, True), ("p2_n4_jit", 2, 4, True)):
        X = np.linspace(0, 1, n + 1)
        grid = fb.simplexgrid(X, X)
        if jit:
            grid = fb.jitter(grid)
        fes = fb.FESpace(fb.H1Pk(1, 2, order), grid)
        el = orc.Element("H1Pk", 1, order)
        colptr, rowval, nzval, b = orc.example200_assemble(el, "Triangle2D", grid["Coordinates"], grid["CellNodes"], grid["CellVolumes"], fes["CellDofs"],
                                                           fes.ndofs, 1.3, rhs)
        np.savez(os.path.join(HERE, f"oracle_example200_{tag}.npz"), coords=grid["Coordinates"], cellnodes=grid["CellNodes"], celldofs=fes["CellDofs"],
                 mu=1.3, colptr=colptr, rowval=rowval, nzval=nzval, b=b, order=order, jitter=jit, n=n)
    n = 2
    X = np.linspace(0, 1, n + 1)
    grid = fb.jitter(fb.simplexgrid(X, X), 0.2)
    FESu, FESp = fb.FESpace(fb.H1Pk(2, 2, 2), grid), fb.FESpace(fb.H1Pk(1, 2, 1), grid)
    sol = np.random.default_rng(210).uniform(-1, 1, FESu.ndofs + FESp.ndofs)
    colptr, rowval, nzval, b = orc.example210_assemble(grid["Coordinates"], grid["CellNodes"], grid["CellVolumes"], FESu["CellDofs"], FESp["CellDofs"],
                                                       FESu.ndofs, FESp.ndofs, sol, 1e-2, 0.0, True, True)
    np.savez(os.path.join(HERE, "oracle_example210_n2_jit.npz"), coords=grid["Coordinates"], cellnodes=grid["CellNodes"], celldofs_u=FESu["CellDofs"],
             celldofs_p=FESp["CellDofs"], sol=sol, mu=1e-2, colptr=colptr, rowval=rowval, nzval=nzval, b=b)
    print("written:", sorted(f for f in os.listdir(HERE) if f.endswith(".npz")))


if __name__ == "__main__":
    main()

"""Example200_LowLevelPoisson with the batched B200 `assemble!` (mirror of
/root/reference/examples/Example200_LowLevelPoisson.jl:32-101): assemble on the device, penalise the boundary dofs on the
host exactly like the reference (`A[dof,dof] = 1e60`, `b[dof] = 0`, Example200:86-93), solve on the host."""
from __future__ import annotations

import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import femb_b200 as fb  # noqa: E402


def solve_poisson_lowlevel(fe_space, mu, rhs):
    import scipy.sparse.linalg as spla

    solution = fb.FEVector(fe_space)
    stiffness_matrix = fb.FEMatrix(fe_space, fe_space)
    rhs_vector = fb.FEVector(fe_space)
    t0 = time.time()
    fb.assemble_poisson(stiffness_matrix.entries, rhs_vector.entries, fe_space, rhs, mu)
    bdofs = fb.boundarydofs(fe_space)
    for dof in bdofs:  # fix boundary dofs (Example200:86-93)
        stiffness_matrix.entries[int(dof), int(dof)] = 1.0e60
        rhs_vector.entries[dof - 1] = 0.0
    time_assembly = time.time() - t0
    t0 = time.time()
    solution.entries[:] = spla.spsolve(stiffness_matrix.entries.to_scipy().tocsc(), rhs_vector.entries)
    return solution, time_assembly, time.time() - t0


def main(maxnref=5, order=2, mu=1.0, rhs=None):
    rhs = rhs or fb.AffineFunction([0.0], [[1.0, -1.0]])  # x -> x[1] - x[2]
    FEType = fb.H1Pk(1, 2, order)
    out = []
    for level in range(1, maxnref + 1):
        X = np.linspace(0, 1, 2 ** level + 1)
        xgrid = fb.simplexgrid(X, X)
        fe_space = fb.FESpace(FEType, xgrid)
        solution, ta, ts = solve_poisson_lowlevel(fe_space, mu, rhs)
        print(f"LEVEL = {level}, ndofs = {fe_space.ndofs}, assembly {ta:.3f}s, solve {ts:.3f}s, max|u| = {np.abs(solution.entries).max():.6e}")
        out.append(solution)
    return out


if __name__ == "__main__":
    main()

"""Example210_LowLevelNavierStokes with the batched B200 assembly (mirror of
/root/reference/examples/Example210_LowLevelNavierStokes.jl:67-216): Taylor-Hood P2-P1, Newton iteration with the
linear part assembled once and the linearised convection re-assembled every iteration on the device; boundary data,
penalisation, the linear solves and the error computation stay on the host as in the reference."""
from __future__ import annotations

import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import femb_b200 as fb  # noqa: E402

MU = 1.0e-2


def u_exact(x, t=0.0):
    e = np.exp(-8.0 * np.pi * np.pi * MU * t)
    return np.stack([e * np.sin(2 * np.pi * x[..., 0]) * np.sin(2 * np.pi * x[..., 1]), e * np.cos(2 * np.pi * x[..., 0]) * np.cos(2 * np.pi * x[..., 1])], axis=-1)


def p_exact(x, t=0.0):
    return np.exp(-16 * np.pi * np.pi * MU * t) * (np.cos(4 * np.pi * x[..., 0]) - np.cos(4 * np.pi * x[..., 1])) / 4


def p2_nodal_points(xgrid):
    c, fn = xgrid["Coordinates"], xgrid["FaceNodes"]
    return np.concatenate([c, 0.5 * (c[fn[:, 0] - 1] + c[fn[:, 1] - 1])])


def p2_interpolate(xgrid, u, teval):
    """interpolate!(u_init[1], u!; time) of the reference for H1Pk{2,2,2} on the host (vectorised): node values and edge dofs
    that preserve the edge mean (MomentInterpolator of order 0: (mean - (u(a) + u(b)) / 6) * 3/2, 3-point Gauss rule) —
    the same operator femb_interpolate applies on the device.  -> (npoints, 2)"""
    from femb_b200.quadrature import gauss_jacobi_01

    c, fn = xgrid["Coordinates"], xgrid["FaceNodes"]
    a, b = c[np.minimum(fn[:, 0], fn[:, 1]) - 1], c[np.maximum(fn[:, 0], fn[:, 1]) - 1]
    un = u(c, teval)
    ua, ub = u(a, teval), u(b, teval)
    qx, qw = gauss_jacobi_01(3)
    mean = sum(w * u(a + x * (b - a), teval) for x, w in zip(qx, qw))
    return np.concatenate([un, (mean - (ua + ub) / 6.0) * 1.5])


def solve_stokes_lowlevel(FES, teval=0.0, maxits=20):
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla

    FESu, FESp = FES
    sol, A, b = fb.FEVector(FES), fb.FEMatrix(FES), fb.FEVector(FES)
    update_system = fb.prepare_assembly(A, b, FESu, FESp, sol, teval=teval, mu=MU)
    update_system(True, False)
    Alin, blin = A.entries.nzval.copy(), b.entries.copy()  # keep the linear part (Example210:164-165)
    # boundary data: nodal interpolation of the exact velocity (Example210:168-173)
    u_init = np.zeros_like(sol.entries)
    ue = p2_interpolate(FESu.xgrid, u_exact, teval)
    u_init[: FESu.ndofs] = np.concatenate([ue[:, 0], ue[:, 1]])
    fixed = np.concatenate([fb.boundarydofs(FESu), [FESu.ndofs + 1]])  # + one pressure dof
    E = A.entries
    n = E.shape[0]
    fx = fixed - 1
    free = np.setdiff1d(np.arange(n), fx)

    def solve(nz, rhs):
        """The reference penalises the fixed dofs with 1e60 and calls `\\` (Example210:197-200); scipy's SuperLU does not
        digest that scaling, so the host solve here eliminates the fixed dofs instead — the limit of the penalty system."""
        M = sp.csr_matrix(sp.csc_matrix((nz, E.rowval - 1, E.colptr - 1), shape=E.shape))
        x = np.zeros(n)
        x[fx] = u_init[fx]
        x[free] = spla.spsolve(M[free][:, free].tocsc(), rhs[free] - M[free][:, fx] @ x[fx])
        return M, x

    nlres = np.inf
    for it in range(1, maxits + 1):
        M, x = solve(E.nzval, b.entries)
        sol.entries[:] = x
        res = M @ sol.entries - b.entries
        res[fx] = 0
        linres = np.linalg.norm(res)
        E.nzval[:] = 0.0  # fill!(A.entries.cscmatrix.nzval, 0); fill!(b.entries, 0)   (Example210:187-188)
        b.entries[:] = 0.0
        update_system(False, True)  # Newton-linearised convection around the current solution, on the device
        E.nzval += Alin
        b.entries += blin
        M = sp.csr_matrix(sp.csc_matrix((E.nzval, E.rowval - 1, E.colptr - 1), shape=E.shape))
        res = M @ sol.entries - b.entries
        res[fx] = 0
        nlres = np.linalg.norm(res)
        print(f"ITERATION {it}: linear residual = {linres:.3e}, nonlinear residual = {nlres:.3e}")
        if nlres < max(1.0e-12, 20 * linres):
            break
    return sol, nlres, it


def solve_stokes_device(FES, teval=0.0, maxits=20):
    """The same Newton loop with everything but the linear solve on the device: the linear part is kept there
    (femb_system_save = `Alin = deepcopy(A)`), each iteration re-assembles the convection, adds the saved part, applies
    the 1e60 penalisation to the boundary dofs enumerated on the device and evaluates both residual norms there
    (Example210:175-212).  Only the penalised system travels to the host, for the linear solve."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    from femb_b200 import _lib
    from femb_b200.fespace import parse_pattern

    FESu, FESp = FES
    sol, A, b = fb.FEVector(FES), fb.FEMatrix(FES), fb.FEVector(FES)
    update_system = fb.prepare_assembly(A, b, FESu, FESp, sol, teval=teval, mu=MU)
    update_system(True, False)  # builds the pattern, assembles the linear part (device copy stays in the session)
    S = update_system.session
    ctx = S.ctx
    ctx.system_save()
    E = A.entries
    n = E.shape[0]
    segs = parse_pattern(FESu.fetype.get_dofmap_pattern(FESu.xgrid.geometry))
    nb = ctx.boundary_dofs(0, segs, download=False)
    fx = np.concatenate([ctx.boundary_dofs(0, segs), [FESu.ndofs + 1]]) - 1
    # boundary data: interpolate!(u_init[1], u!; time) restricted to the boundary dofs, ON THE DEVICE (femb_interpolate: node
    # values + edge-mean preserving edge dofs with the reference's 3-point Gauss rule); it stays there for dirichlet_apply,
    # the copy is only what the host solve below eliminates with
    from femb_b200.quadrature import gauss_jacobi_01

    qx, qw = gauss_jacobi_01(3)
    u_init = np.zeros(n)
    u_init[: FESu.ndofs] = ctx.interpolate(0, segs, FESu.ndofs, _lib.FUNC_EX210_U, np.array([MU, teval]), qx, qw, only_boundary=True)
    assert fx.shape[0] == nb + 1
    free = np.setdiff1d(np.arange(n), fx)
    pin = np.array([FESu.ndofs + 1])
    egu, eiu, eip = S.evaluator(0, "Gradient"), S.evaluator(0, "Identity"), S.evaluator(1, "Identity")

    def host_solve():
        """the linear solve of the penalised system = elimination of the fixed dofs (see solve_stokes_lowlevel)."""
        M = sp.csr_matrix(sp.csc_matrix((ctx.matrix_get(), E.rowval - 1, E.colptr - 1), shape=E.shape))
        rhs = ctx.vector_get(n)
        x = np.zeros(n)
        x[fx] = u_init[fx]
        x[free] = spla.spsolve(M[free][:, free].tocsc(), rhs[free] - M[free][:, fx] @ x[fx])
        return x

    ctx.dirichlet_apply(True, 0, pin, None, 1.0e60)  # device-resident boundary data; also registers the fixed rows the residuals skip
    nlres = np.inf
    for it in range(1, maxits + 1):
        sol.entries[:] = host_solve()
        linres = ctx.residual(sol.entries, zero_fixed=True)  # uploads sol: it is also the linearisation point below
        ctx.matrix_zero()
        ctx.vector_zero()
        ctx.assemble_navierstokes(egu._eid, eiu._eid, eip._eid, MU, False, True, _lib.RHS_NONE, None)
        ctx.system_add_saved()
        ctx.dirichlet_apply(True, 0, pin, None, 1.0e60)
        nlres = ctx.residual(None, zero_fixed=True)
        print(f"ITERATION {it}: linear residual = {linres:.3e}, nonlinear residual = {nlres:.3e}")
        if nlres < max(1.0e-12, 20 * linres):
            break
    return sol, nlres, it


def l2_errors(sol, FESu, FESp, order=2):
    """Example210:88-96: subtract the integral mean of the pressure (compute_error(sol[2], nothing, order, 1)), then the
    VertexRule l2 errors of velocity and pressure — compute_error runs on the device (femb_eval_error)."""
    nu = FESu.ndofs
    pmean = fb.compute_error(FESp, sol.entries[nu:], None, order, 1).sum()
    sol.entries[nu:] -= pmean
    err_u = float(np.sqrt(fb.compute_error(FESu, sol.entries[:nu], u_exact, 2).sum()))
    err_p = float(np.sqrt(fb.compute_error(FESp, sol.entries[nu:], lambda x: p_exact(x)[..., None], 2).sum()))
    return err_u, err_p


def main(nref=5, order=2, device_loop=False):
    X = np.linspace(0, 1, 2 ** nref + 1)
    xgrid = fb.simplexgrid(X, X)
    FES = [fb.FESpace(fb.H1Pk(2, 2, order), xgrid, "velocity space"), fb.FESpace(fb.H1Pk(1, 2, order - 1), xgrid, "pressure space")]
    sol, nlres, its = (solve_stokes_device if device_loop else solve_stokes_lowlevel)(FES)
    err, err_p = l2_errors(sol, *FES, order=order)
    print(f"l2 error velo = {err:.4e}, l2 error pressure = {err_p:.4e} after {its} Newton iterations")
    return sol, nlres, its, err


if __name__ == "__main__":
    main(device_loop="--device-loop" in sys.argv)

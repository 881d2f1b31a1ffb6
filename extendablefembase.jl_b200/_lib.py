"""ctypes binding of libfemb.so — one Python stub per ``ccall`` of julia/FEMBaseB200.jl.

The product path fails loudly: if the shared library is missing or no B200 is usable,
``load()`` / ``Context()`` raise; nothing here computes on the CPU.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIBPATH = os.environ.get("FEMB_LIB") or os.path.join(HERE, "libfemb.so")  # FEMB_LIB: alternative build (tuning experiments)

# enums of include/femb.h
FAMILY = {"H1": 0, "Hdiv": 1, "L2": 2, "Hcurl": 3}
OP_IDENTITY, OP_GRADIENT, OP_DIVERGENCE = 0, 1, 2
RHS_NONE, RHS_AFFINE, RHS_QPVALUES, RHS_EX210 = 0, 1, 2, 3
FUNC_AFFINE, FUNC_EX210_U = 1, 2
BIL_TRANSPOSE_COPY, BIL_SYMMETRIC_UPPER, BIL_NO_VOLUME = 1, 2, 4
SCATTER_ATOMIC, SCATTER_COLORED, SCATTER_GATHER = 0, 1, 2
ENTITY_FACES, ENTITY_EDGES = 0, 1

_i32, _i64, _f64 = C.c_int32, C.c_int64, C.c_double
_p = C.c_void_p

# name -> (restype, argtypes): every symbol include/femb.h declares
SIGNATURES = {
    "femb_version": (_i32, []),
    "femb_create": (_i32, [_i32, C.POINTER(_p)]),
    "femb_destroy": (_i32, [_p]),
    "femb_last_error": (C.c_char_p, [_p]),
    "femb_sync": (_i32, [_p]),
    "femb_host_register": (_i32, [_p, _p, _i64]),
    "femb_host_unregister": (_i32, [_p, _p]),
    "femb_mesh_set": (_i32, [_p, _i32, _i64, _p, _i32, _i64, _p, _p]),
    "femb_space_set": (_i32, [_p, _i32, _i32, _i32, _i64, _i32, _i32, _i32, _p, _p, _p]),
    "femb_grid_entities": (_i32, [_p, _i32, C.POINTER(_i64)]),
    "femb_grid_entities_get": (_i32, [_p, _i32, _p, _p, _p, _p, _p]),
    "femb_space_set_from_pattern": (_i32, [_p, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _p, C.POINTER(_i64), C.POINTER(_i64)]),
    "femb_space_get_celldofs": (_i32, [_p, _i32, _p]),
    "femb_boundary_dofs": (_i32, [_p, _i32, _i32, _p, C.POINTER(_i64)]),
    "femb_boundary_dofs_get": (_i32, [_p, _p]),
    "femb_dirichlet_apply": (_i32, [_p, _i32, _i64, _i64, _p, _p, _f64]),
    "femb_interpolate": (_i32, [_p, _i32, _i32, _p, _i64, _i32, _p, _i32, _p, _p, _i32, _p]),
    "femb_residual": (_i32, [_p, _p, _i32, _p, C.POINTER(_f64)]),
    "femb_eval_error": (_i32, [_p, _i32, _p, _p, _i32, _p]),
    "femb_block_matmul": (_i32, [_p, _i32, _i32, _p, _f64, _i32, _p]),
    "femb_lrmatmul": (_i32, [_p, _p, _p, _p, _p, _f64, C.POINTER(_f64)]),
    "femb_system_save": (_i32, [_p]),
    "femb_system_add_saved": (_i32, [_p]),
    "femb_quadrature_set": (_i32, [_p, _i32, _p, _p]),
    "femb_evaluator_set": (_i32, [_p, _i32, _i32, _i32, _p, _p]),
    "femb_evaluator_eval": (_i32, [_p, _i32, _i64, _i64, _p]),
    "femb_qp_coords": (_i32, [_p, _p]),
    "femb_matrix_layout": (_i32, [_p, _i32, _p, _i32, _p]),
    "femb_pattern_build": (_i32, [_p, _i32, _p, _p, _i64, _p, C.POINTER(_i64)]),
    "femb_pattern_adopt": (_i32, [_p, _i64, _i64, _i64, _p, _p, _i32, _p, _p]),
    "femb_pattern_get": (_i32, [_p, _p, _p]),
    "femb_pattern_get_slots": (_i32, [_p, _i64, _p, _p]),
    "femb_pattern_compress": (_i32, [_p, C.POINTER(_i64)]),
    "femb_pattern_track": (_i32, [_p, _i32]),
    "femb_matrix_zero": (_i32, [_p]),
    "femb_vector_zero": (_i32, [_p]),
    "femb_matrix_get": (_i32, [_p, _p]),
    "femb_matrix_add_to": (_i32, [_p, _p]),
    "femb_vector_get": (_i32, [_p, _p]),
    "femb_matrix_get_range": (_i32, [_p, _i64, _i64, _p]),
    "femb_vector_get_range": (_i32, [_p, _i64, _i64, _p]),
    "femb_vector_set": (_i32, [_p, _i32, _p]),
    "femb_set_scatter": (_i32, [_p, _i32]),
    "femb_assemble_bilinear": (_i32, [_p, _i32, _i32, _i32, _i32, _f64, _i32, _f64]),
    "femb_assemble_hdiv_mixed": (_i32, [_p, _i32, _i32, _i32, _f64, _f64]),
    "femb_assemble_linear": (_i32, [_p, _i32, _i32, _f64, _i32, _p]),
    "femb_assemble_poisson": (_i32, [_p, _i32, _i32, _f64, _i32, _p, _f64]),
    "femb_assemble_poisson_host": (_i32, [_p, _i32, _i32, _f64, _i32, _p, _f64, _p, _p, _p, _p, _p, _i32]),
    "femb_assemble_navierstokes": (_i32, [_p, _i32, _i32, _i32, _f64, _i32, _i32, _i32, _p]),
    "femb_last_kernel_ms": (_i32, [_p, C.POINTER(C.c_float), C.POINTER(_i32)]),
    "femb_launch_count": (_i64, [_p]),
    "femb_device_ptr": (_p, [_p, _i32]),
    "femb_region_begin": (_i32, [_p]),
    "femb_region_end": (_i32, [_p, C.POINTER(C.c_float)]),
    "femb_comm_unique_id": (_i32, [C.c_char_p]),
    "femb_comm_init": (_i32, [_p, _i32, _i32, C.c_char_p]),
    "femb_comm_set_exchange": (_i32, [_p, _i32, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "femb_pattern_extra": (_i32, [_p, _i64, _p, _p]),
    "femb_exchange": (_i32, [_p]),
    "femb_set_fused_exchange": (_i32, [_p, _i32]),
    "femb_comm_destroy": (_i32, [_p]),
}

_lib = None


class FembError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(f"libfemb status {status}: {msg}")
        self.status = status


def load(path: str | None = None) -> C.CDLL:
    """dlopen libfemb.so and bind every declared symbol (raises if one is missing)."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIBPATH
    if not os.path.exists(p):
        raise FileNotFoundError(f"{p} not found — run `python __graft_entry__.py build` (no CPU fallback exists)")
    lib = C.CDLL(p)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype, fn.argtypes = res, args
    if path is None:
        _lib = lib
    return lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(_p)


def _arr(a, dtype):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=dtype)
    return a


class Context:
    """Owns one ``femb_ctx`` (one CUDA device)."""

    def __init__(self, device: int = 0):
        self.lib = load()
        h = _p()
        st = self.lib.femb_create(device, C.byref(h))
        if st != 0:
            raise FembError(st, self.lib.femb_last_error(None).decode())
        self.h = h
        self._keep = []

    def close(self):
        if getattr(self, "h", None):
            self.lib.femb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _ck(self, st):
        if st != 0:
            raise FembError(st, self.lib.femb_last_error(self.h).decode())

    # ---- registration ---------------------------------------------------
    def mesh_set(self, coords, cellnodes, cellvolumes=None):
        coords = _arr(coords, np.float64)
        cellnodes = _arr(cellnodes, np.int32)
        cellvolumes = _arr(cellvolumes, np.float64)
        nn, dim = coords.shape
        nc, nnpc = cellnodes.shape
        self._ck(self.lib.femb_mesh_set(self.h, dim, nn, _ptr(coords), nnpc, nc, _ptr(cellnodes), _ptr(cellvolumes)))
        self.dim, self.ncells = dim, nc

    def space_set(self, sid, family, ncomp, ndofs, nd, nd_all, handler, celldofs=None, signs=None, orient=None):
        celldofs, signs, orient = _arr(celldofs, np.int32), _arr(signs, np.int32), _arr(orient, np.int32)
        self._ck(self.lib.femb_space_set(self.h, sid, family, ncomp, ndofs, nd, nd_all, handler, _ptr(celldofs), _ptr(signs), _ptr(orient)))

    # ---- grid components / dofmaps generated on the device ---------------
    def grid_entities(self, kind):
        """Enumerate the faces (ENTITY_FACES) or edges (ENTITY_EDGES) of the registered mesh on the device -> count."""
        n = _i64(0)
        self._ck(self.lib.femb_grid_entities(self.h, kind, C.byref(n)))
        return int(n.value)

    def grid_entities_get(self, kind):
        """-> dict with cell_entities (ncells, per_cell), entity_nodes (n, k), signs, and for faces orientations, facecells."""
        n = self.grid_entities(kind)
        faces = kind == ENTITY_FACES or self.dim == 2
        per = self.dim + 1 if faces else 6
        k = 3 if (faces and self.dim == 3) else 2
        out = {"cell_entities": np.empty((self.ncells, per), np.int32), "entity_nodes": np.empty((n, k), np.int32),
               "signs": np.empty((self.ncells, per), np.int32)}
        if faces:
            out["orientations"] = np.empty((self.ncells, per), np.int32)
            out["facecells"] = np.empty((n, 2), np.int32)
        self._ck(self.lib.femb_grid_entities_get(self.h, kind, _ptr(out["cell_entities"]), _ptr(out["entity_nodes"]), _ptr(out["signs"]),
                                                 _ptr(out.get("orientations")), _ptr(out.get("facecells"))))
        return out

    def space_set_from_pattern(self, sid, family, ncomp, nd, nd_all, handler, segs):
        """segs: list of (type, each_component, q) with type in 'nfeic' -> (ndofs, coffset); CellDofs stay on the device."""
        code = {"n": 0, "f": 1, "e": 2, "i": 3, "c": 3}
        flat = np.array([[code[t], int(each), q] for t, each, q in segs], dtype=np.int32)
        ndofs, coffset = _i64(0), _i64(0)
        self._ck(self.lib.femb_space_set_from_pattern(self.h, sid, family, ncomp, nd, nd_all, handler, flat.shape[0], _ptr(flat), C.byref(ndofs), C.byref(coffset)))
        return int(ndofs.value), int(coffset.value)

    def space_get_celldofs(self, sid, nd):
        out = np.empty((self.ncells, nd), np.int32)
        self._ck(self.lib.femb_space_get_celldofs(self.h, sid, _ptr(out)))
        return out

    # ---- after the assembly: Dirichlet rows, residual ---------------------
    _SEG_CODE = {"n": 0, "f": 1, "e": 2, "i": 3, "c": 3}

    def boundary_dofs(self, sid, segs, download=True):
        """boundarydofs(FES) on the device -> sorted unique 1-based dofs (int64 like the host mirror), or the count."""
        flat = np.array([[self._SEG_CODE[t], int(each), q] for t, each, q in segs], dtype=np.int32)
        n = _i64(0)
        self._ck(self.lib.femb_boundary_dofs(self.h, sid, flat.shape[0], _ptr(flat), C.byref(n)))
        if not download:
            return int(n.value)
        out = np.empty(int(n.value), np.int32)
        self._ck(self.lib.femb_boundary_dofs_get(self.h, _ptr(out)))
        return out.astype(np.int64)

    def dirichlet_apply(self, use_boundary_list=True, dof_offset=0, extra_dofs=None, values=None, penalty=1.0e60):
        extra = _arr(extra_dofs, np.int64)
        values = _arr(values, np.float64)
        self._ck(self.lib.femb_dirichlet_apply(self.h, int(use_boundary_list), dof_offset, 0 if extra is None else extra.shape[0], _ptr(extra), _ptr(values), penalty))

    def interpolate(self, sid, segs, ndofs, func_kind, params, qx=None, qw=None, only_boundary=False, dof_offset=0, download=True):
        """interpolate!(target, FES, u) for nodal H1 spaces (node dofs + at most one dof per edge) on the device; the result
        stays there as the boundary data of dirichlet_apply(values=None).  qx/qw: rule on [0,1] for the edge means
        (None: midpoint values).  func_kind: FUNC_AFFINE / FUNC_EX210_U."""
        flat = np.array([[self._SEG_CODE[t], int(each), q] for t, each, q in segs], dtype=np.int32)
        params = _arr(params, np.float64)
        qx, qw = _arr(qx, np.float64), _arr(qw, np.float64)
        out = np.empty(ndofs, np.float64) if download else None
        self._ck(self.lib.femb_interpolate(self.h, sid, flat.shape[0], _ptr(flat), dof_offset, func_kind, _ptr(params), 0 if qx is None else qx.shape[0],
                                           _ptr(qx), _ptr(qw), int(only_boundary), _ptr(out)))
        return out

    def eval_error(self, eid, coefficients, exact_qp, p, resultdim):
        """compute_error: -> (ncells, resultdim) array of sum_qp (uh - u)^p |T| w_qp."""
        coefficients, exact_qp = _arr(coefficients, np.float64), _arr(exact_qp, np.float64)
        out = np.empty((self.ncells, resultdim), np.float64)
        self._ck(self.lib.femb_eval_error(self.h, eid, _ptr(coefficients), _ptr(exact_qp), p, _ptr(out)))
        return out

    def block_matmul(self, a, b, brow=-1, bcol=-1, factor=1.0, transposed=False):
        """addblock_matmul!(a, B[brow, bcol], b; factor, transposed): a (float64, contiguous) is updated in place."""
        assert a.dtype == np.float64 and a.flags.c_contiguous
        b = _arr(b, np.float64)
        self._ck(self.lib.femb_block_matmul(self.h, brow, bcol, _ptr(b), factor, int(transposed), _ptr(a)))
        return a

    def lrmatmul(self, a1, b1, a2=None, b2=None, factor=1.0):
        """lrmatmul(a, A, b) / ldrdmatmul(a1, a2, A, b1, b2): (a1 - a2)' A (b1 - b2) * factor."""
        a1, a2, b1, b2 = (_arr(v, np.float64) for v in (a1, a2, b1, b2))
        out = _f64(0.0)
        self._ck(self.lib.femb_lrmatmul(self.h, _ptr(a1), _ptr(a2), _ptr(b1), _ptr(b2), factor, C.byref(out)))
        return float(out.value)

    def system_save(self):
        self._ck(self.lib.femb_system_save(self.h))

    def system_add_saved(self):
        self._ck(self.lib.femb_system_add_saved(self.h))

    def residual(self, x=None, zero_fixed=True, want_vector=False, n=None):
        """-> ||A x - b||_2 (and the vector when want_vector)."""
        x = _arr(x, np.float64)
        res = np.empty(n, np.float64) if want_vector else None
        nrm = _f64(0.0)
        self._ck(self.lib.femb_residual(self.h, _ptr(x), int(zero_fixed), _ptr(res), C.byref(nrm)))
        return (float(nrm.value), res) if want_vector else float(nrm.value)

    def quadrature_set(self, xref, w):
        xref, w = _arr(xref, np.float64), _arr(w, np.float64)
        self._ck(self.lib.femb_quadrature_set(self.h, w.shape[0], _ptr(xref), _ptr(w)))

    def evaluator_set(self, eid, sid, op, refvals=None, refderivs=None):
        refvals, refderivs = _arr(refvals, np.float64), _arr(refderivs, np.float64)
        self._ck(self.lib.femb_evaluator_set(self.h, eid, sid, op, _ptr(refvals), _ptr(refderivs)))

    def evaluator_eval(self, eid, c0, c1, nqp, nd, resultdim):
        out = np.empty((c1 - c0, nqp, nd, resultdim), dtype=np.float64)
        self._ck(self.lib.femb_evaluator_eval(self.h, eid, c0, c1, _ptr(out)))
        return out

    def qp_coords(self, nqp):
        out = np.empty((self.ncells, nqp, self.dim), dtype=np.float64)
        self._ck(self.lib.femb_qp_coords(self.h, _ptr(out)))
        return out

    # ---- pattern --------------------------------------------------------
    def matrix_layout(self, row_spaces, col_spaces):
        r, c = np.asarray(row_spaces, np.int32), np.asarray(col_spaces, np.int32)
        self._ck(self.lib.femb_matrix_layout(self.h, r.shape[0], _ptr(r), c.shape[0], _ptr(c)))
        self._nrs, self._ncs = r.shape[0], c.shape[0]

    def pattern_build(self, blocks, extra_diag=None) -> int:
        br = np.asarray([b[0] for b in blocks], np.int32)
        bc = np.asarray([b[1] for b in blocks], np.int32)
        ex = None if extra_diag is None else np.ascontiguousarray(extra_diag, dtype=np.int64)
        nnz = _i64(0)
        self._ck(self.lib.femb_pattern_build(self.h, br.shape[0], _ptr(br), _ptr(bc), 0 if ex is None else ex.shape[0], _ptr(ex), C.byref(nnz)))
        self.nnz = nnz.value
        return self.nnz

    def pattern_extra(self, rows, cols):
        """Extra structural entries (1-based) merged into the next pattern_build (multi-GPU halo rows)."""
        rows, cols = _arr(rows, np.int64), _arr(cols, np.int64)
        self._ck(self.lib.femb_pattern_extra(self.h, rows.shape[0], _ptr(rows), _ptr(cols)))

    def pattern_adopt(self, nrows, ncols, colptr, rowval, blocks):
        colptr, rowval = _arr(colptr, np.int64), _arr(rowval, np.int64)
        br = np.asarray([b[0] for b in blocks], np.int32)
        bc = np.asarray([b[1] for b in blocks], np.int32)
        self.nnz = int(colptr[-1] - 1)
        self._ck(self.lib.femb_pattern_adopt(self.h, nrows, ncols, self.nnz, _ptr(colptr), _ptr(rowval), br.shape[0], _ptr(br), _ptr(bc)))

    def pattern_get(self, ncols):
        colptr = np.empty(ncols + 1, dtype=np.int64)
        rowval = np.empty(self.nnz, dtype=np.int64)
        self._ck(self.lib.femb_pattern_get(self.h, _ptr(colptr), _ptr(rowval)))
        return colptr, rowval

    def pattern_get_colptr(self, ncols):
        colptr = np.empty(ncols + 1, dtype=np.int64)
        self._ck(self.lib.femb_pattern_get(self.h, _ptr(colptr), None))
        return colptr

    def pattern_get_slots(self, slots):
        """rowval (1-based) at the given 0-based nz positions"""
        slots = _arr(slots, np.int64)
        out = np.empty(slots.shape[0], dtype=np.int64)
        self._ck(self.lib.femb_pattern_get_slots(self.h, slots.shape[0], _ptr(slots), _ptr(out)))
        return out

    def pattern_track(self, enable=True):
        self._ck(self.lib.femb_pattern_track(self.h, 1 if enable else 0))

    def pattern_compress(self) -> int:
        nnz = _i64(0)
        self._ck(self.lib.femb_pattern_compress(self.h, C.byref(nnz)))
        self.nnz = nnz.value
        return self.nnz

    # ---- values ---------------------------------------------------------
    def matrix_zero(self):
        self._ck(self.lib.femb_matrix_zero(self.h))

    def vector_zero(self):
        self._ck(self.lib.femb_vector_zero(self.h))

    def matrix_get(self, out=None):
        out = np.empty(self.nnz, dtype=np.float64) if out is None else out
        self._ck(self.lib.femb_matrix_get(self.h, _ptr(out)))
        return out

    def matrix_add_to(self, inout):
        self._ck(self.lib.femb_matrix_add_to(self.h, _ptr(inout)))

    def vector_get(self, n=None, out=None):
        out = np.empty(n, dtype=np.float64) if out is None else out
        self._ck(self.lib.femb_vector_get(self.h, _ptr(out)))
        return out

    def matrix_get_range(self, first, count, out):
        """out must be a contiguous float64 array (view) of at least `count` elements"""
        self._ck(self.lib.femb_matrix_get_range(self.h, first, count, _ptr(out)))

    def vector_get_range(self, first, count, out):
        self._ck(self.lib.femb_vector_get_range(self.h, first, count, _ptr(out)))

    def vector_set(self, which, v):
        v = _arr(v, np.float64)
        self._ck(self.lib.femb_vector_set(self.h, which, _ptr(v)))

    def set_scatter(self, policy):
        self._ck(self.lib.femb_set_scatter(self.h, policy))

    def sync(self):
        self._ck(self.lib.femb_sync(self.h))

    def host_register(self, a: np.ndarray):
        self._ck(self.lib.femb_host_register(self.h, _ptr(a), a.nbytes))

    def host_unregister(self, a: np.ndarray):
        self._ck(self.lib.femb_host_unregister(self.h, _ptr(a)))

    # ---- assembly -------------------------------------------------------
    def assemble_bilinear(self, er, ec, brow, bcol, factor=1.0, flags=0, drop_tol=0.0):
        self._ck(self.lib.femb_assemble_bilinear(self.h, er, ec, brow, bcol, factor, flags, drop_tol))

    def assemble_hdiv_mixed(self, e_id, e_div, e_q, factor_mass=1.0, factor_div=1.0):
        self._ck(self.lib.femb_assemble_hdiv_mixed(self.h, e_id, e_div, e_q, factor_mass, factor_div))

    def assemble_linear(self, ev, brow, factor, rhs_kind, data):
        data = _arr(data, np.float64)
        self._ck(self.lib.femb_assemble_linear(self.h, ev, brow, factor, rhs_kind, _ptr(data)))

    def assemble_poisson(self, eg, ei, mu, rhs_kind, data, drop_tol):
        data = _arr(data, np.float64)
        self._ck(self.lib.femb_assemble_poisson(self.h, eg, ei, mu, rhs_kind, _ptr(data), drop_tol))

    def assemble_poisson_host(self, eg, ei, mu, rhs_kind, data, drop_tol, coords, cellnodes, cellvolumes, nzval_out, b_out, nchunks=16):
        """Streamed host-to-host assembly (H2D / kernel / D2H overlap); arrays must be C-contiguous of the right dtype."""
        data = _arr(data, np.float64)
        for a, dt in ((coords, np.float64), (cellnodes, np.int32), (nzval_out, np.float64)):
            assert a is None or (a.dtype == dt and a.flags.c_contiguous)
        self._ck(self.lib.femb_assemble_poisson_host(self.h, eg, ei, mu, rhs_kind, _ptr(data), drop_tol, _ptr(coords), _ptr(cellnodes), _ptr(cellvolumes),
                                                     _ptr(nzval_out), _ptr(b_out), nchunks))

    def assemble_navierstokes(self, egu, eiu, eip, mu, linear, nonlinear, rhs_kind, data):
        data = _arr(data, np.float64)
        self._ck(self.lib.femb_assemble_navierstokes(self.h, egu, eiu, eip, mu, int(linear), int(nonlinear), rhs_kind, _ptr(data)))

    # ---- multi-GPU -------------------------------------------------------
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        st = load().femb_comm_unique_id(buf)
        if st != 0:
            raise FembError(st, "ncclGetUniqueId failed (is libnccl.so.2 loadable?)")
        return buf.raw

    def comm_init(self, nranks: int, rank: int, uid: bytes):
        self._ck(self.lib.femb_comm_init(self.h, nranks, rank, C.create_string_buffer(uid, 128)))

    def comm_set_exchange(self, plan):
        """plan: ExchangePlan of partition.py (0-based slot / vector index lists per peer)."""
        a = lambda x: np.ascontiguousarray(x, dtype=np.int64)  # noqa: E731
        peers = np.ascontiguousarray(plan.peers, dtype=np.int32)
        arrs = [a(plan.send_ptr), a(plan.send_slots), a(plan.recv_ptr), a(plan.recv_slots), a(plan.sendb_ptr), a(plan.send_b), a(plan.recvb_ptr), a(plan.recv_b)]
        self._ck(self.lib.femb_comm_set_exchange(self.h, peers.shape[0], _ptr(peers), *[_ptr(x) for x in arrs]))

    def set_fused_exchange(self, enable=True):
        self._ck(self.lib.femb_set_fused_exchange(self.h, 1 if enable else 0))

    def exchange(self):
        self._ck(self.lib.femb_exchange(self.h))

    def comm_destroy(self):
        self._ck(self.lib.femb_comm_destroy(self.h))

    def last_kernel_ms(self):
        ms, n = C.c_float(0), _i32(0)
        self._ck(self.lib.femb_last_kernel_ms(self.h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def region_begin(self):
        self._ck(self.lib.femb_region_begin(self.h))

    def region_end(self) -> float:
        ms = C.c_float(0)
        self._ck(self.lib.femb_region_end(self.h, C.byref(ms)))
        return ms.value

    def launch_count(self) -> int:
        return int(self.lib.femb_launch_count(self.h))

    def device_ptr(self, which) -> int:
        return self.lib.femb_device_ptr(self.h, which) or 0

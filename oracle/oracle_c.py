"""TEST INFRASTRUCTURE — ctypes loader of oracle/libfemb_oracle.so (C restatement of the reference's serial
Example200 loop, femb_oracle_c.c).  Importable only from tests/, __graft_entry__.smoke() and bench.py's CPU legs."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def load():
    global _lib
    if _lib is None:
        so = os.path.join(HERE, "libfemb_oracle.so")
        src = os.path.join(HERE, "femb_oracle_c.c")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            r = subprocess.run(["make", "-s", "-C", HERE], capture_output=True, text=True)
            if r.returncode != 0 and not os.path.exists(so):
                raise RuntimeError("building oracle/libfemb_oracle.so failed:\n" + r.stdout + r.stderr)
        _lib = C.CDLL(so)
        _lib.femb_oracle_example200.restype = C.c_int64
        _lib.femb_oracle_max_threads.restype = C.c_int
    return _lib


def max_threads() -> int:
    return int(load().femb_oracle_max_threads())


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def example200(coords, cellnodes, cellvol, celldofs, w, xref, refvals, refderivs, mu, rhs, drop_tol, colptr, rowval, nzval, b,
               nthreads=1, cell_begin=0, cell_end=None):
    """Adds into (nzval, b) in place; returns the number of contributions without a slot."""
    lib = load()
    dim = coords.shape[1]
    ncells, nd = celldofs.shape
    nqp = w.shape[0]
    cell_end = ncells if cell_end is None else cell_end
    refvals = np.ascontiguousarray(refvals.reshape(nqp, nd), dtype=np.float64)
    refderivs = np.ascontiguousarray(refderivs.reshape(nqp, nd, dim), dtype=np.float64)
    rhs = None if rhs is None else np.ascontiguousarray(rhs, dtype=np.float64)
    for a, dt in ((coords, np.float64), (cellnodes, np.int32), (cellvol, np.float64), (celldofs, np.int32), (colptr, np.int64), (rowval, np.int64),
                  (nzval, np.float64)):
        assert a.dtype == dt and a.flags.c_contiguous
    return int(lib.femb_oracle_example200(
        C.c_int(dim), C.c_int64(ncells), _p(coords), _p(cellnodes), _p(cellvol), _p(celldofs), C.c_int(nd), C.c_int(nqp), _p(w), _p(xref),
        _p(refvals), _p(refderivs), C.c_double(mu), _p(rhs), C.c_double(drop_tol), _p(colptr), _p(rowval), _p(nzval), _p(b),
        C.c_int(nthreads), C.c_int64(cell_begin), C.c_int64(cell_end)))

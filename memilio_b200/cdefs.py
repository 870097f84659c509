"""ctypes mirrors of the POD structs in include/mio_cuda.h (kept field-for-field in the header's order)."""
from __future__ import annotations

import ctypes as C

import numpy as np

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)
c_int64_p = C.POINTER(C.c_int64)


class IntegratorDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("reserved", C.c_int32), ("abs_tol", C.c_double), ("rel_tol", C.c_double),
                ("dt_min", C.c_double), ("dt_max", C.c_double)]


class GraphDesc(C.Structure):
    _fields_ = [("model", C.c_int32), ("sim_kind", C.c_int32), ("num_groups", C.c_int32), ("num_nodes", C.c_int32),
                ("num_edges", C.c_int32), ("num_psets", C.c_int32), ("num_ctables", C.c_int32),
                ("ct_max_loc", C.c_int32), ("ct_max_pts", C.c_int32), ("mob_max_pts", C.c_int32),
                ("max_runs", C.c_int32), ("record_daily", C.c_int32), ("max_days", C.c_int32),
                ("traj_capacity", C.c_int32),
                ("edge_start", c_int32_p), ("edge_end", c_int32_p), ("edge_coeff", c_double_p),
                ("node_pset", c_int32_p), ("node_ctable", c_int32_p),
                ("integrator", IntegratorDesc), ("t0", C.c_double), ("dt_graph", C.c_double),
                ("vacc_days", C.c_int32), ("flow_simulation", C.c_int32), ("vacc_partial", c_double_p),
                ("vacc_full", c_double_p)]


class RunData(C.Structure):
    _fields_ = [("psets", c_double_p), ("ctables", c_double_p), ("node_local", c_double_p), ("y0", c_double_p),
                ("dt0", c_double_p), ("mobtables", c_double_p)]


def dptr(a):
    """double* of a C-contiguous float64 array (None -> NULL)."""
    if a is None:
        return C.cast(None, c_double_p)
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(c_double_p)


def iptr(a):
    if a is None:
        return C.cast(None, c_int32_p)
    assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(c_int32_p)


def lptr(a):
    if a is None:
        return C.cast(None, c_int64_p)
    assert a.dtype == np.int64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(c_int64_p)


def integrator_desc(kind=0, abs_tol=1e-10, rel_tol=1e-5, dt_min=None, dt_max=None) -> IntegratorDesc:
    fi = np.finfo(np.float64)
    return IntegratorDesc(kind, 0, abs_tol, rel_tol, float(fi.tiny) if dt_min is None else dt_min,
                          float(fi.max) if dt_max is None else dt_max)

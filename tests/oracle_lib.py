"""ctypes loader for the CPU oracle (oracle/, test infrastructure only)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from memilio_b200 import cdefs, flat
from memilio_b200.problem import GraphProblem

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
_LIB = None


def build():
    subprocess.run(["make", "-s", "-C", ORACLE_DIR], check=True)
    return os.path.join(ORACLE_DIR, "_build", "liboracle.so")


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(ORACLE_DIR, "_build", "liboracle.so")
        src = [os.path.join(ORACLE_DIR, f) for f in os.listdir(ORACLE_DIR) if f.endswith((".cpp", ".hpp"))]
        if not os.path.exists(path) or any(os.path.getmtime(s) > os.path.getmtime(path) for s in src):
            build()
        _LIB = C.CDLL(path)
        _LIB.orc_simulate.restype = C.c_int
        _LIB.orc_graph_simulate.restype = C.c_int
        _LIB.orc_dampings_cache.restype = C.c_int
    return _LIB


def dampings_cache(rows, cols, times, levels, types, mats):
    K = len(times)
    times = np.ascontiguousarray(times, dtype=np.float64)
    levels = np.ascontiguousarray(levels, dtype=np.int32)
    types = np.ascontiguousarray(types, dtype=np.int32)
    mats = np.ascontiguousarray(mats, dtype=np.float64).reshape(K, rows * cols)
    ot = np.zeros(max(K, 1))
    om = np.zeros((max(K, 1), rows * cols))
    n = lib().orc_dampings_cache(rows, cols, K, cdefs.dptr(times), cdefs.iptr(levels), cdefs.iptr(types),
                                 cdefs.dptr(mats), cdefs.dptr(ot), cdefs.dptr(om))
    return ot[:n].copy(), om[:n].reshape(n, rows, cols).copy()


def contact_matrix_at(ct, G, max_loc, max_pts, t):
    out = np.zeros((G, G))
    ct = np.ascontiguousarray(ct, dtype=np.float64)
    lib().orc_contact_matrix_at(cdefs.dptr(ct), G, max_loc, max_pts, C.c_double(t), cdefs.dptr(out))
    return out


def osecir_rhs(G, pset, local, ct, max_loc, max_pts, pop, y, t):
    out = np.zeros(G * 10)
    args = [np.ascontiguousarray(a, dtype=np.float64) for a in (pset, local, ct, pop, y)]
    lib().orc_osecir_rhs(G, cdefs.dptr(args[0]), cdefs.dptr(args[1]), cdefs.dptr(args[2]), max_loc, max_pts,
                         cdefs.dptr(args[3]), cdefs.dptr(args[4]), C.c_double(t), cdefs.dptr(out))
    return out


def simulate(G, pset, local, ct, max_loc, max_pts, y0, t0, tmax, dt, integrator=None, capacity=4096,
             model=flat.MODEL_OSECIR):
    """mio::simulate on the oracle. Returns (times[npts], values[npts][n], final dt, counters[3])."""
    integ = integrator if integrator is not None else cdefs.integrator_desc()
    n = G * flat.num_compartments(model)
    args = [np.ascontiguousarray(a, dtype=np.float64) for a in (pset, local, ct, y0)]
    ot = np.zeros(capacity)
    oy = np.zeros((capacity, n))
    odt = np.zeros(1)
    cnt = np.zeros(3, dtype=np.int64)
    npts = lib().orc_simulate(model, G, cdefs.dptr(args[0]), cdefs.dptr(args[1]), cdefs.dptr(args[2]), max_loc,
                              max_pts, cdefs.dptr(args[3]), C.c_double(t0), C.c_double(tmax), C.c_double(dt),
                              C.byref(integ), capacity, cdefs.dptr(ot), cdefs.dptr(oy), cdefs.dptr(odt),
                              cdefs.lptr(cnt))
    assert 0 < npts <= capacity
    return ot[:npts].copy(), oy[:npts].copy(), float(odt[0]), cnt


def simulate_vvs(G, pset, local, ct, max_loc, max_pts, y0, t0, tmax, dt, integrator=None, capacity=4096,
                 sim_kind=flat.SIM_MODEL, vacc_partial=None, vacc_full=None):
    """osecirvvs::simulate on the oracle. Returns (times[npts], values[npts][n], final dt, counters[3])."""
    integ = integrator if integrator is not None else cdefs.integrator_desc()
    n = G * flat.OSECIRVVS_NCOMP
    args = [np.ascontiguousarray(a, dtype=np.float64) for a in (pset, local, ct, y0)]
    vp = None if vacc_partial is None else np.ascontiguousarray(vacc_partial, dtype=np.float64).reshape(G, -1)
    vf = None if vacc_full is None else np.ascontiguousarray(vacc_full, dtype=np.float64).reshape(G, -1)
    days = 0 if vp is None else vp.shape[1]
    ot = np.zeros(capacity)
    oy = np.zeros((capacity, n))
    odt = np.zeros(1)
    cnt = np.zeros(3, dtype=np.int64)
    L = lib()
    L.orc_simulate_vvs.restype = C.c_int
    npts = L.orc_simulate_vvs(sim_kind, G, cdefs.dptr(args[0]), cdefs.dptr(args[1]), cdefs.dptr(args[2]), max_loc,
                              max_pts, cdefs.dptr(args[3]), days, cdefs.dptr(vp), cdefs.dptr(vf), C.c_double(t0),
                              C.c_double(tmax), C.c_double(dt), C.byref(integ), capacity, cdefs.dptr(ot),
                              cdefs.dptr(oy), cdefs.dptr(odt), cdefs.lptr(cnt))
    assert 0 < npts <= capacity
    return ot[:npts].copy(), oy[:npts].copy(), float(odt[0]), cnt


def graph_simulate(p: GraphProblem, run: int, t_max: float, halfsteps: int = 0):
    """GraphSimulation::advance(t_max) of run `run` on the oracle.
    Returns dict(state[N][n], dt[N], counters[N][4], steps, halfsteps[steps][N][n])."""
    d = p.desc(max_runs=1)
    rd = p.run_data(run, 1)
    N, n = p.num_nodes, p.n
    state = np.zeros((N, n))
    dt = np.zeros(N)
    cnt = np.zeros((N, 4), dtype=np.int64)
    hs = np.zeros((max(halfsteps, 1), N, n))
    steps = lib().orc_graph_simulate(C.byref(d), C.byref(rd), C.c_double(t_max), cdefs.dptr(state), cdefs.dptr(dt),
                                     cdefs.lptr(cnt), cdefs.dptr(hs) if halfsteps else None, halfsteps)
    assert steps >= 0
    return dict(state=state, dt=dt, counters=cnt, steps=steps, halfsteps=hs[:min(steps, halfsteps)])


def edge_apply_sequence(G, pset, local, ct, max_loc, max_pts, y0, coeff, t, dt):
    n = G * 10
    out = np.zeros((3, 2, n))
    args = [np.ascontiguousarray(a, dtype=np.float64) for a in (pset, local, ct, y0, coeff)]
    lib().orc_edge_apply_sequence(G, cdefs.dptr(args[0]), cdefs.dptr(args[1]), cdefs.dptr(args[2]), max_loc, max_pts,
                                  cdefs.dptr(args[3]), cdefs.dptr(args[4]), C.c_double(t), C.c_double(dt),
                                  cdefs.dptr(out))
    return out

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


def graph_simulate_daily(p: GraphProblem, run: int, t_max: float, num_days: int):
    """orc_graph_simulate + interpolate_simulation_result(graph): dict(daily[N][num_days][n], state, counters)."""
    d = p.desc(max_runs=1)
    rd = p.run_data(run, 1)
    N, n = p.num_nodes, p.n
    daily = np.zeros((N, num_days, n))
    state = np.zeros((N, n))
    cnt = np.zeros((N, 4), dtype=np.int64)
    L = lib()
    L.orc_graph_simulate_daily.restype = C.c_int
    steps = L.orc_graph_simulate_daily(C.byref(d), C.byref(rd), C.c_double(t_max), C.c_int(num_days), cdefs.dptr(daily),
                                       cdefs.dptr(state), cdefs.lptr(cnt))
    assert steps >= 0, steps
    return dict(daily=daily, state=state, counters=cnt, steps=steps)


def graph_simulate_mobility(p: GraphProblem, run: int, t_max: float, save_groups, cap: int = 4096):
    """get_mobility_results() of every edge (all edges built with saved_compartment_indices = save_groups).
    Returns (npts[E], times[E][cap], values[E][cap][len(save_groups)+1], state[N][n])."""
    d = p.desc(max_runs=1)
    rd = p.run_data(run, 1)
    E, K = p.num_edges, len(save_groups) + 1
    ptr = np.zeros(len(save_groups) + 1, dtype=np.int32)
    idx = []
    for i, g in enumerate(save_groups):
        idx += [int(x) for x in g]
        ptr[i + 1] = len(idx)
    idx = np.ascontiguousarray(idx, dtype=np.int32)
    npts = np.zeros(E, dtype=np.int32)
    times = np.zeros((E, cap))
    values = np.zeros((E, cap, K))
    state = np.zeros((p.num_nodes, p.n))
    L = lib()
    L.orc_graph_simulate_mobility.restype = C.c_int
    steps = L.orc_graph_simulate_mobility(C.byref(d), C.byref(rd), C.c_double(t_max), C.c_int(len(save_groups)),
                                          cdefs.iptr(ptr), cdefs.iptr(idx), C.c_int(cap), cdefs.iptr(npts),
                                          cdefs.dptr(times), cdefs.dptr(values), cdefs.dptr(state))
    assert steps >= 0
    return npts, times, values, state


def interpolate(times, values, tps=None, abs_tol: float = -1.0):
    """interpolate_simulation_result (analyze_result.h:186-265) of one time series."""
    times = np.ascontiguousarray(times, dtype=np.float64)
    values = np.ascontiguousarray(values, dtype=np.float64).reshape(len(times), -1)
    n = values.shape[1]
    cap = (len(tps) if tps is not None else int(np.floor(times[-1]) - np.ceil(times[0])) + 3)
    ot = np.zeros(cap)
    ov = np.zeros((cap, n))
    tp = None if tps is None else np.ascontiguousarray(tps, dtype=np.float64)
    L = lib()
    L.orc_interpolate.restype = C.c_int
    k = L.orc_interpolate(cdefs.dptr(times), cdefs.dptr(values), len(times), n, cdefs.dptr(tp),
                          0 if tp is None else len(tp), C.c_double(abs_tol), cap, cdefs.dptr(ot), cdefs.dptr(ov))
    assert 0 <= k <= cap
    return ot[:k].copy(), ov[:k].copy()


def ensemble_reduce(data, p=(), want_sum=False, want_mean=False):
    """sum_nodes / ensemble_mean / ensemble_percentile (analyze_result.h:267-343) of data[runs][nodes][T][n]."""
    data = np.ascontiguousarray(data, dtype=np.float64)
    runs, nodes, T, n = data.shape
    pa = np.ascontiguousarray(p, dtype=np.float64)
    out_sum = np.zeros((runs, T, n)) if want_sum else None
    out_mean = np.zeros((nodes, T, n)) if want_mean else None
    out_pct = np.zeros((max(len(pa), 1), nodes, T, n))
    lib().orc_ensemble_reduce(cdefs.dptr(data), runs, nodes, T, n, len(pa), cdefs.dptr(pa), cdefs.dptr(out_sum),
                              cdefs.dptr(out_mean), cdefs.dptr(out_pct))
    return dict(sum=out_sum, mean=out_mean, percentile=out_pct[:len(pa)])


def study_run_block(ps, problem: GraphProblem, run0: int, count: int, t_max: float, num_days: int,
                    want_daily: bool = False):
    """ParameterStudy::run_impl for runs [run0, run0+count) on the oracle: draw_sample + advance + interpolate per run.
    `ps` is a memilio_b200.study.ParameterStudy (its distributions and seeds), `problem` any GraphProblem sampled from
    it (graph topology and capacities). Returns dict(state[count][N][n], counters[count][N][4], checksum, daily?)."""
    from memilio_b200 import study
    sd, keep = study.build_desc(ps.graph, ps.seeds, ps.rng_counter, ps.ct_max_pts, ps.mob_max_pts)
    gd = problem.desc(max_runs=1)
    N, n = problem.num_nodes, problem.n
    state = np.zeros((count, N, n))
    cnt = np.zeros((count, N, 4), dtype=np.int64)
    daily = np.zeros((count, N, num_days, n)) if want_daily else None
    chk = C.c_double(0.0)
    L = lib()
    L.orc_study_run_block.restype = C.c_int
    rc = L.orc_study_run_block(C.byref(sd), C.byref(gd), C.c_int64(run0), C.c_int(count), C.c_double(t_max),
                               C.c_int(num_days), cdefs.dptr(daily), cdefs.dptr(state), cdefs.lptr(cnt), C.byref(chk))
    assert rc == 0, rc
    del keep
    return dict(state=state, counters=cnt, checksum=chk.value, daily=daily)


def sample_problem(ps, run0: int, count: int) -> GraphProblem:
    """The runs [run0, run0+count) of the study `ps` sampled by the ORACLE's literal draw_sample(graph)
    (oracle/mio_oracle_sampling.hpp), as a GraphProblem. Graphs without school holidays (one contact table)."""
    from memilio_b200 import study
    g = ps.graph
    assert len(g.holiday_sets) == 0
    d, keep = study.build_desc(g, ps.seeds, ps.rng_counter, ps.ct_max_pts, ps.mob_max_pts)
    G, N, n = g.num_groups, g.num_nodes, g.n
    psets = np.zeros((count, 1, flat.pset_size(g.model, G)))
    ctables = np.zeros((count, 1, flat.ctable_size(G, g.num_locations, ps.ct_max_pts)))
    local = np.zeros((count, N, 2))
    y0 = np.zeros((count, N, n))
    mob = np.zeros((count, flat.mobtable_size(n, ps.mob_max_pts))) if ps.mob_max_pts > 0 else None
    L = lib()
    L.orc_study_sample_runs.restype = C.c_int
    rc = L.orc_study_sample_runs(C.byref(d), C.c_int64(run0), C.c_int(count), cdefs.dptr(psets), cdefs.dptr(ctables),
                                 cdefs.dptr(local), cdefs.dptr(y0), cdefs.dptr(mob), None)
    assert rc == 0
    del keep
    return GraphProblem(model=g.model, num_groups=G, num_nodes=N, edge_start=g.edge_start, edge_end=g.edge_end,
                        edge_coeff=g.edge_coeff, psets=psets, ctables=ctables, node_local=local, y0=y0,
                        ct_max_loc=g.num_locations, ct_max_pts=ps.ct_max_pts, mobtables=mob,
                        mob_max_pts=ps.mob_max_pts, integrator=ps.integrator, t0=ps.t0, dt_graph=ps.dt,
                        sim_kind=g.sim_kind, vacc_partial=g.vacc_partial, vacc_full=g.vacc_full,
                        flow_simulation=int(getattr(ps, "flow_simulation", 0)))


def simulate_flows(G, pset, local, ct, max_loc, max_pts, y0, t0, tmax, dt, integrator=None, capacity=4096):
    """mio::simulate_flows (FlowSimulation<osecir::Model>) on the oracle.
    Returns (times[npts], populations[npts][n], flows[npts][G*15], final dt, counters[3])."""
    return simulate_flows_stepwise(G, pset, local, ct, max_loc, max_pts, y0, t0, [tmax], dt, integrator, capacity)


def simulate_flows_stepwise(G, pset, local, ct, max_loc, max_pts, y0, t0, tmaxs, dt, integrator=None, capacity=4096):
    """One FlowSimulation advanced to every entry of tmaxs in turn (one advance call each)."""
    integ = integrator if integrator is not None else cdefs.integrator_desc()
    args = [np.ascontiguousarray(a, dtype=np.float64) for a in (pset, local, ct, y0)]
    tm = np.ascontiguousarray(tmaxs, dtype=np.float64)
    ot = np.zeros(capacity)
    oy = np.zeros((capacity, G * 10))
    of = np.zeros((capacity, G * 15))
    odt = np.zeros(1)
    cnt = np.zeros(3, dtype=np.int64)
    L = lib()
    L.orc_simulate_flows_stepwise.restype = C.c_int
    npts = L.orc_simulate_flows_stepwise(G, cdefs.dptr(args[0]), cdefs.dptr(args[1]), cdefs.dptr(args[2]), max_loc,
                                         max_pts, cdefs.dptr(args[3]), C.c_double(t0), len(tm), cdefs.dptr(tm),
                                         C.c_double(dt), C.byref(integ), capacity, cdefs.dptr(ot), cdefs.dptr(oy),
                                         cdefs.dptr(of), cdefs.dptr(odt), cdefs.lptr(cnt))
    assert 0 < npts <= capacity
    return ot[:npts].copy(), oy[:npts].copy(), of[:npts].copy(), float(odt[0]), cnt


def graph_simulate_flows(p: GraphProblem, run: int, t_max: float):
    """GraphSimulation::advance of a graph of FlowSimulation nodes (p.flow_simulation): dict(state, flows, counters)."""
    d = p.desc(max_runs=1)
    d.flow_simulation = 1
    rd = p.run_data(run, 1)
    N, n, G = p.num_nodes, p.n, p.num_groups
    state = np.zeros((N, n))
    flows = np.zeros((N, G * 15))
    cnt = np.zeros((N, 4), dtype=np.int64)
    L = lib()
    L.orc_graph_simulate_flows.restype = C.c_int
    steps = L.orc_graph_simulate_flows(C.byref(d), C.byref(rd), C.c_double(t_max), cdefs.dptr(state), cdefs.dptr(flows),
                                       cdefs.lptr(cnt))
    assert steps >= 0
    return dict(state=state, flows=flows, counters=cnt, steps=steps)


# ---- dynamic NPIs (oracle/mio_oracle_npis.hpp) -------------------------------------------------------------------------
class NpiDesc(C.Structure):
    _fields_ = [("num_thresholds", C.c_int32), ("thresholds", cdefs.c_double_p), ("npi_ptr", cdefs.c_int32_p),
                ("npi_value", cdefs.c_double_p), ("npi_level", cdefs.c_int32_p), ("npi_type", cdefs.c_int32_p),
                ("idx_ptr", cdefs.c_int32_p), ("npi_matrix_idx", cdefs.c_int32_p), ("npi_weights", cdefs.c_double_p),
                ("num_weights", C.c_int32), ("reserved", C.c_int32), ("duration", C.c_double), ("delay", C.c_double),
                ("base", C.c_double), ("directive_begin", C.c_double), ("directive_end", C.c_double)]


def npi_desc(thresholds, duration=14.0, delay=0.0, base=1.0, directive_begin=None, directive_end=None):
    """thresholds: list of (value, [dict(value, level, type, matrix_indices, group_weights), ...])"""
    fi = np.finfo(np.float64)
    keep = dict(th=np.array([t for t, _ in thresholds], dtype=np.float64), ptr=[0], val=[], lvl=[], typ=[], iptr=[0],
                idx=[], w=[])
    for _, lst in thresholds:
        for s in lst:
            keep["val"].append(s["value"])
            keep["lvl"].append(s["level"])
            keep["typ"].append(s["type"])
            keep["idx"] += list(s["matrix_indices"])
            keep["iptr"].append(len(keep["idx"]))
            keep["w"].append(np.asarray(s["group_weights"], dtype=np.float64))
        keep["ptr"].append(len(keep["val"]))
    nw = len(keep["w"][0]) if keep["w"] else 1
    arr = dict(th=keep["th"], ptr=np.array(keep["ptr"], np.int32), val=np.array(keep["val"] or [0.0], np.float64),
               lvl=np.array(keep["lvl"] or [0], np.int32), typ=np.array(keep["typ"] or [0], np.int32),
               iptr=np.array(keep["iptr"], np.int32), idx=np.array(keep["idx"] or [0], np.int32),
               w=np.ascontiguousarray(np.stack(keep["w"]) if keep["w"] else np.zeros((1, 1))))
    if len(arr["th"]) == 0:
        arr["th"] = np.zeros(1)
    d = NpiDesc(len(thresholds), cdefs.dptr(arr["th"]), cdefs.iptr(arr["ptr"]), cdefs.dptr(arr["val"]),
                cdefs.iptr(arr["lvl"]), cdefs.iptr(arr["typ"]), cdefs.iptr(arr["iptr"]), cdefs.iptr(arr["idx"]),
                cdefs.dptr(arr["w"]), nw, 0, duration, delay, base,
                float(fi.min) if directive_begin is None else directive_begin,
                float(fi.max) if directive_end is None else directive_end)
    d._keep = arr
    return d


def _damping_lists(mats):
    """mats: per matrix a list of (value, level, type, time)"""
    counts = np.array([len(m) for m in mats], np.int32)
    flatl = [d for m in mats for d in m]
    times = np.array([d[3] for d in flatl] or [0.0], np.float64)
    levels = np.array([d[1] for d in flatl] or [0], np.int32)
    types = np.array([d[2] for d in flatl] or [0], np.int32)
    values = np.array([d[0] for d in flatl] or [0.0], np.float64)
    return counts, times, levels, types, values


def npi_implement(mats, rows, calls, capacity=64):
    """implement_dynamic_npis on a group of rows x 1 expressions with constant dampings; calls: list of
    (npi sampling dict, begin, end). Returns per matrix a list of (time, level, type, coeffs[rows])."""
    counts, times, levels, types, values = _damping_lists(mats)
    dl = [npi_desc([(0.0, [c[0]])]) for c in calls]  # (kept alive: the array below holds plain copies of the structs)
    descs = (NpiDesc * len(calls))(*dl)
    begins = np.array([c[1] for c in calls], np.float64)
    ends = np.array([c[2] for c in calls], np.float64)
    oc = np.zeros(len(mats), np.int32)
    ot = np.zeros(capacity)
    ol = np.zeros(capacity, np.int32)
    oy = np.zeros(capacity, np.int32)
    ov = np.zeros((capacity, rows))
    L = lib()
    n = L.orc_npi_implement(len(mats), rows, cdefs.iptr(counts), cdefs.dptr(times), cdefs.iptr(levels),
                            cdefs.iptr(types), cdefs.dptr(values), len(calls), descs, cdefs.dptr(begins),
                            cdefs.dptr(ends), cdefs.iptr(oc), cdefs.dptr(ot), cdefs.iptr(ol), cdefs.iptr(oy),
                            cdefs.dptr(ov), capacity)
    assert n >= 0
    out, k = [], 0
    for m in range(len(mats)):
        out.append([(ot[k + i], int(ol[k + i]), int(oy[k + i]), ov[k + i].copy()) for i in range(oc[m])])
        k += oc[m]
    return out


def npi_active_damping(dampings, lvl, typ, t):
    counts, times, levels, types, values = _damping_lists([dampings])
    L = lib()
    L.orc_npi_active_damping.restype = C.c_double
    return L.orc_npi_active_damping(int(counts[0]), cdefs.dptr(times), cdefs.iptr(levels), cdefs.iptr(types),
                                    cdefs.dptr(values), lvl, typ, C.c_double(t))


def npi_damping_indices(dampings, lvl, typ, begin, end):
    counts, times, levels, types, values = _damping_lists([dampings])
    out = np.zeros(max(1, len(dampings)), np.int32)
    n = lib().orc_npi_damping_indices(int(counts[0]), cdefs.dptr(times), cdefs.iptr(levels), cdefs.iptr(types),
                                      cdefs.dptr(values), lvl, typ, C.c_double(begin), C.c_double(end), cdefs.iptr(out))
    return [int(x) for x in out[:n]]


def npi_simulate(G, pset, local, ct, max_loc, max_pts, y0, t0, tmaxs, dt, npi, integrator=None, frozen=False,
                 static=(), query=(), capacity=8192):
    """osecir::Simulation with DynamicNPIsInfectedSymptoms on the oracle. static: predefined dampings as
    (location, time, level, type, matrix[G][G]). Returns dict(t, y, contact[nq][G][G], num_dampings[max_loc],
    counters[4] = accepted, rejected, forced, NPIs implemented)."""
    integ = integrator if integrator is not None else cdefs.integrator_desc()
    args = [np.ascontiguousarray(a, dtype=np.float64) for a in (pset, local, ct, y0)]
    tm = np.ascontiguousarray(tmaxs, dtype=np.float64)
    sl = np.array([s[0] for s in static] or [0], np.int32)
    stt = np.array([s[1] for s in static] or [0.0], np.float64)
    slv = np.array([s[2] for s in static] or [0], np.int32)
    sty = np.array([s[3] for s in static] or [0], np.int32)
    sm = np.ascontiguousarray(np.stack([np.asarray(s[4], dtype=np.float64).reshape(G, G) for s in static])
                              if static else np.zeros((1, G, G)))
    q = np.ascontiguousarray(query if len(query) else [0.0], dtype=np.float64)
    ocm = np.zeros((len(q), G, G))
    ond = np.zeros(max_loc, np.int32)
    ot = np.zeros(capacity)
    oy = np.zeros((capacity, G * 10))
    cnt = np.zeros(4, dtype=np.int64)
    L = lib()
    n = L.orc_npi_simulate(G, cdefs.dptr(args[0]), cdefs.dptr(args[1]), cdefs.dptr(args[2]), max_loc, max_pts,
                           cdefs.dptr(args[3]), C.c_double(t0), len(tm), cdefs.dptr(tm), C.c_double(dt),
                           C.byref(integ), C.byref(npi), int(frozen), len(static), cdefs.iptr(sl), cdefs.dptr(stt),
                           cdefs.iptr(slv), cdefs.iptr(sty), cdefs.dptr(sm), len(query), cdefs.dptr(q),
                           cdefs.dptr(ocm), cdefs.iptr(ond), capacity, cdefs.dptr(ot), cdefs.dptr(oy), cdefs.lptr(cnt))
    assert 0 < n <= capacity
    return dict(t=ot[:n].copy(), y=oy[:n].copy(), contact=ocm[:len(query)], num_dampings=ond, counters=cnt)


def npi_graph_simulate(p: GraphProblem, run: int, t_max: float, node_npis, halfsteps: int = 0, query=()):
    """orc_graph_simulate with dynamic NPIs on the nodes; node_npis: one npi_desc per node. Returns dict(state,
    halfsteps[halfsteps][N][n], contact[N][nq][G][G], implemented[N])."""
    d = p.desc(max_runs=1)
    rd = p.run_data(run, 1)
    N, n, G = p.num_nodes, p.n, p.num_groups
    descs = (NpiDesc * N)(*node_npis)
    state = np.zeros((N, n))
    hs = np.zeros((max(halfsteps, 1), N, n))
    q = np.ascontiguousarray(query if len(query) else [0.0], dtype=np.float64)
    cm = np.zeros((N, len(q), G, G))
    impl = np.zeros(N, np.int32)
    steps = lib().orc_npi_graph_simulate(C.byref(d), C.byref(rd), descs, C.c_double(t_max), cdefs.dptr(state),
                                         cdefs.dptr(hs), halfsteps, len(query), cdefs.dptr(q), cdefs.dptr(cm),
                                         cdefs.iptr(impl))
    assert steps >= 0
    return dict(state=state, halfsteps=hs[:halfsteps], contact=cm[:, :len(query)], implemented=impl, steps=steps)

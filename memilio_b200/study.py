"""ParameterStudy on the B200 backend (host-side mirror of cpp/memilio/compartments/parameter_studies.h:47-374).

Sampling (`osecir::draw_sample(graph)`, cpp/models/ode_secir/parameter_space.h:207-251) runs in the C library
(csrc/study.cpp, bit-exact with the reference's libstdc++ sampling); runs are processed in batches on the GPU.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

from . import cdefs, flat
from .problem import GraphProblem

DIST_NONE, DIST_UNIFORM, DIST_NORMAL, DIST_LOGNORMAL, DIST_EXPONENTIAL, DIST_CONSTANT = range(6)
MAX_GROUPS = 8


class UncertainValue(C.Structure):
    """mio::UncertainValue (cpp/memilio/utils/uncertain_value.h:45-258) as a POD."""
    _fields_ = [("kind", C.c_int32), ("reserved", C.c_int32), ("value", C.c_double), ("a", C.c_double),
                ("b", C.c_double), ("lower", C.c_double), ("upper", C.c_double), ("quantile", C.c_double)]


class DampingSampling(C.Structure):
    _fields_ = [("value", UncertainValue), ("level", C.c_int32), ("type", C.c_int32), ("time", C.c_double),
                ("matrix_mask", C.c_uint64), ("group_weights", C.c_double * MAX_GROUPS)]


class StudyDesc(C.Structure):
    _fields_ = [("model", C.c_int32), ("num_groups", C.c_int32), ("num_nodes", C.c_int32),
                ("num_locations", C.c_int32), ("num_dampings", C.c_int32), ("num_holiday_sets", C.c_int32),
                ("ct_max_pts", C.c_int32), ("mob_max_pts", C.c_int32), ("mobility_location", C.c_int32),
                ("num_seeds", C.c_int32), ("seeds", C.c_uint32 * 8), ("rng_counter", C.c_uint64),
                ("header", C.POINTER(UncertainValue)), ("age_params", C.POINTER(UncertainValue)),
                ("node_local", C.POINTER(UncertainValue)), ("populations", C.POINTER(UncertainValue)),
                ("baseline", cdefs.c_double_p), ("minimum", cdefs.c_double_p),
                ("dampings", C.POINTER(DampingSampling)), ("school_holiday_damping", C.POINTER(DampingSampling)),
                ("node_holiday_set", cdefs.c_int32_p), ("holiday_ptr", cdefs.c_int32_p),
                ("holiday_periods", cdefs.c_double_p)]


UV_DTYPE = np.dtype([("kind", np.int32), ("reserved", np.int32), ("value", np.float64), ("a", np.float64),
                     ("b", np.float64), ("lower", np.float64), ("upper", np.float64), ("quantile", np.float64)])
assert UV_DTYPE.itemsize == C.sizeof(UncertainValue)


def uv_array(shape) -> np.ndarray:
    return np.zeros(shape, dtype=UV_DTYPE)


def set_value(arr, idx, value):
    arr[idx]["kind"] = DIST_NONE
    arr[idx]["value"] = value


def set_uniform(arr, idx, lo, hi):
    """assign_uniform_distribution, cpp/examples/ode_secir_parameter_study_graph.cpp:45-49"""
    arr[idx]["kind"] = DIST_UNIFORM
    arr[idx]["value"] = 0.5 * (np.asarray(hi) + np.asarray(lo))
    arr[idx]["a"] = lo
    arr[idx]["b"] = hi


def set_normal(arr, idx, lower, upper, mean, stddev, quantile=2.5758):
    arr[idx]["kind"] = DIST_NORMAL
    arr[idx]["value"] = mean
    arr[idx]["a"] = mean
    arr[idx]["b"] = stddev
    arr[idx]["lower"] = lower
    arr[idx]["upper"] = upper
    arr[idx]["quantile"] = quantile


@dataclasses.dataclass
class Damping:
    """DampingSampling, cpp/memilio/epidemiology/damping_sampling.h:40-120"""
    value: np.ndarray  # one-element UV array
    time: float
    level: int = 0
    type: int = 0
    matrices: Sequence[int] = (0,)
    group_weights: Optional[Sequence[float]] = None


@dataclasses.dataclass
class StudyGraph:
    """The parameter graph a ParameterStudy samples from: `mio::Graph<osecir::Model, MobilityParameters>` with
    UncertainValues (cpp/memilio/mobility/graph.h:145-300)."""
    num_groups: int
    num_nodes: int
    header: np.ndarray  # UV[8]
    age_params: np.ndarray  # UV[G][16]
    node_local: np.ndarray  # UV[N][2]
    populations: np.ndarray  # UV[N][n]
    baseline: np.ndarray  # [L][G][G]
    minimum: Optional[np.ndarray]
    edge_start: np.ndarray
    edge_end: np.ndarray
    edge_coeff: np.ndarray
    dampings: List[Damping] = dataclasses.field(default_factory=list)
    school_holiday_damping: Optional[Damping] = None
    node_holiday_set: Optional[np.ndarray] = None
    holiday_sets: Sequence[Sequence[Tuple[float, float]]] = ()
    mobility_location: int = 0
    damp_mobility: bool = True
    model: int = flat.MODEL_OSECIR
    # osecirvvs: node type and the run-invariant cumulative vaccination arrays [N][G][days]
    sim_kind: int = flat.SIM_MODEL
    vacc_partial: Optional[np.ndarray] = None
    vacc_full: Optional[np.ndarray] = None

    @property
    def n(self):
        return self.num_groups * flat.num_compartments(self.model)

    @property
    def num_locations(self):
        return int(np.asarray(self.baseline).shape[0])


def _lib():
    from . import capi
    L = capi.lib()
    if not getattr(L, "_study_bound", False):
        L.mio_study_create.argtypes = [C.POINTER(StudyDesc), C.POINTER(C.c_void_p)]
        L.mio_study_destroy.argtypes = [C.c_void_p]
        L.mio_study_destroy.restype = None
        L.mio_study_num_ctables.argtypes = [C.c_void_p]
        L.mio_study_node_ctable.argtypes = [C.c_void_p, cdefs.c_int32_p]
        L.mio_study_sample_runs.argtypes = [C.c_void_p, C.c_int64, C.c_int32, cdefs.c_double_p, cdefs.c_double_p,
                                            cdefs.c_double_p, cdefs.c_double_p, cdefs.c_double_p,
                                            C.POINTER(C.c_uint64)]
        L.mio_study_distribute_runs.argtypes = [C.c_int64, C.c_int32, C.c_int32, C.POINTER(C.c_int64),
                                                C.POINTER(C.c_int64)]
        L.mio_study_rng_raw.argtypes = [C.POINTER(C.c_uint32), C.c_int32, C.c_uint64, C.c_int32,
                                        C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.mio_damping_cache_build.argtypes = [C.c_int32, C.c_int32, C.c_int32, cdefs.c_double_p, cdefs.c_int32_p,
                                              cdefs.c_int32_p, cdefs.c_double_p, C.POINTER(C.c_int32),
                                              cdefs.c_double_p, cdefs.c_double_p]
        L._study_bound = True
    return L


def _check(rc):
    from . import capi
    capi._check(rc)


def _ds_struct(d: Damping, G: int) -> DampingSampling:
    s = DampingSampling()
    C.memmove(C.byref(s.value), d.value.ctypes.data, C.sizeof(UncertainValue))
    s.level, s.type, s.time = d.level, d.type, d.time
    mask = 0
    for l in d.matrices:
        mask |= 1 << int(l)
    s.matrix_mask = mask
    w = np.ones(G) if d.group_weights is None else np.asarray(d.group_weights, dtype=np.float64)
    for i in range(G):
        s.group_weights[i] = w[i]
    return s


def build_desc(g: StudyGraph, seeds: Sequence[int], rng_counter: int, ct_max_pts: int, mob_max_pts: int):
    """(StudyDesc, keep-alive list) for the C library and the oracle."""
    G, N = g.num_groups, g.num_nodes
    d = StudyDesc()
    d.model, d.num_groups, d.num_nodes, d.num_locations = g.model, G, N, g.num_locations
    d.num_dampings = len(g.dampings)
    d.num_holiday_sets = len(g.holiday_sets)
    d.ct_max_pts, d.mob_max_pts, d.mobility_location = ct_max_pts, mob_max_pts, g.mobility_location
    d.num_seeds = len(seeds)
    for i, sd in enumerate(seeds):
        d.seeds[i] = int(sd)
    d.rng_counter = int(rng_counter)
    keep = []

    def uvp(a, shape):
        a = np.ascontiguousarray(a, dtype=UV_DTYPE).reshape(shape)
        keep.append(a)
        return a.ctypes.data_as(C.POINTER(UncertainValue))

    d.header = uvp(g.header, (flat.PSET_HEADER,))
    d.age_params = uvp(g.age_params, (G, flat.age_stride(g.model)))
    d.node_local = uvp(g.node_local, (N, 2))
    d.populations = uvp(g.populations, (N, g.n))
    base = np.ascontiguousarray(g.baseline, dtype=np.float64).reshape(g.num_locations, G, G)
    keep.append(base)
    d.baseline = cdefs.dptr(base)
    if g.minimum is not None:
        mn = np.ascontiguousarray(g.minimum, dtype=np.float64).reshape(g.num_locations, G, G)
        keep.append(mn)
        d.minimum = cdefs.dptr(mn)
    if g.dampings:
        arr = (DampingSampling * len(g.dampings))(*[_ds_struct(x, G) for x in g.dampings])
        keep.append(arr)
        d.dampings = C.cast(arr, C.POINTER(DampingSampling))
    if g.school_holiday_damping is not None:
        hs = _ds_struct(g.school_holiday_damping, G)
        keep.append(hs)
        d.school_holiday_damping = C.pointer(hs)
    if len(g.holiday_sets) > 0:
        ptr = np.zeros(len(g.holiday_sets) + 1, dtype=np.int32)
        periods = []
        for k, hs_ in enumerate(g.holiday_sets):
            periods += [list(p) for p in hs_]
            ptr[k + 1] = len(periods)
        periods = np.ascontiguousarray(np.asarray(periods, dtype=np.float64).reshape(-1, 2))
        ns = np.ascontiguousarray(g.node_holiday_set, dtype=np.int32)
        keep += [ptr, periods, ns]
        d.holiday_ptr, d.holiday_periods, d.node_holiday_set = cdefs.iptr(ptr), cdefs.dptr(periods), cdefs.iptr(ns)
    return d, keep


def distribute_runs(num_runs: int, num_procs: int, rank: int) -> Tuple[int, int]:
    """ParameterStudy::distribute_runs (parameter_studies.h:354-367): (first run, count) of `rank`."""
    a, b = C.c_int64(), C.c_int64()
    _check(_lib().mio_study_distribute_runs(num_runs, num_procs, rank, C.byref(a), C.byref(b)))
    return a.value, b.value


def rng_raw(seeds, counter, count):
    s = (C.c_uint32 * len(seeds))(*[int(x) for x in seeds])
    out = (C.c_uint64 * count)()
    key = C.c_uint64()
    _check(_lib().mio_study_rng_raw(s, len(seeds), C.c_uint64(counter), count, out, C.byref(key)))
    return np.array(out[:], dtype=np.uint64), int(key.value)


def damping_cache(rows, cols, times, levels, types, mats):
    K = len(times)
    times = np.ascontiguousarray(times, dtype=np.float64)
    levels = np.ascontiguousarray(levels, dtype=np.int32)
    types = np.ascontiguousarray(types, dtype=np.int32)
    mats = np.ascontiguousarray(mats, dtype=np.float64).reshape(K, rows * cols)
    ot = np.zeros(max(K, 1))
    om = np.zeros((max(K, 1), rows * cols))
    n = C.c_int32()
    _check(_lib().mio_damping_cache_build(rows, cols, K, cdefs.dptr(times), cdefs.iptr(levels), cdefs.iptr(types),
                                          cdefs.dptr(mats), C.byref(n), cdefs.dptr(ot), cdefs.dptr(om)))
    return ot[:n.value].copy(), om[:n.value].reshape(n.value, rows, cols).copy()


class _RecordShape:
    """N, El, n of an ensemble record (a rank without runs has no record but takes part in the reduction)."""

    def __init__(self, N, El, n):
        self.N, self.El, self.n = N, El, n


class ParameterStudy:
    """`mio::ParameterStudy(params_graph, t0, tmax, dt, num_runs)` (parameter_studies.h:84) for a Graph-ODE-SECIR
    study. `run` keeps the reference's semantics: run r is sampled with RNG counter `counter + (r << 32)`, advanced
    to tmax, then handed to `process_result`; after `run` the counter moves on by `num_runs << 32`."""

    def __init__(self, graph: StudyGraph, t0: float, tmax: float, dt: float, num_runs: int,
                 seeds: Sequence[int] = (1, 2, 3, 4, 5, 6), ct_max_pts: int = 4, mob_max_pts: Optional[int] = None,
                 integrator: Optional[cdefs.IntegratorDesc] = None):
        self.graph, self.t0, self.tmax, self.dt, self.num_runs = graph, t0, tmax, dt, num_runs
        self.seeds = list(seeds)
        self.rng_counter = 0
        self.ct_max_pts = ct_max_pts
        if mob_max_pts is None:
            # the osecirvvs draw_sample(graph) copies the edges unchanged (ode_secirvvs/parameter_space.h:205-211)
            mob_max_pts = ct_max_pts if (graph.damp_mobility and graph.dampings and
                                         graph.model == flat.MODEL_OSECIR) else 0
        self.mob_max_pts = mob_max_pts
        self.integrator = integrator if integrator is not None else cdefs.integrator_desc()
        self._h = None
        self._keep = None

    # -- sampling ---------------------------------------------------------------------------------------------
    def _study(self):
        if self._h is None:
            d, keep = build_desc(self.graph, self.seeds, self.rng_counter, self.ct_max_pts, self.mob_max_pts)
            h = C.c_void_p()
            _check(_lib().mio_study_create(C.byref(d), C.byref(h)))
            self._h, self._keep = h, keep
        return self._h

    def _invalidate(self):
        if self._h is not None:
            _lib().mio_study_destroy(self._h)
            self._h = None

    def seed(self, seeds):
        self.seeds = list(seeds)
        self._invalidate()

    def sample_buffers(self, count: int, pinned: bool = False):
        """Host buffers for `sample(..., out=)`; pinned (page locked, via torch) when asked for and available."""
        h = self._study()
        g = self.graph
        G, N, n = g.num_groups, g.num_nodes, g.n
        nct = _lib().mio_study_num_ctables(h)
        shapes = {"psets": (count, 1, flat.pset_size(g.model, G)),
                  "ctables": (count, nct, flat.ctable_size(G, g.num_locations, self.ct_max_pts)),
                  "node_local": (count, N, 2), "y0": (count, N, n)}
        if self.mob_max_pts > 0:
            shapes["mobtables"] = (count, flat.mobtable_size(n, self.mob_max_pts))
        out = {}
        for k, shp in shapes.items():
            if pinned:
                import torch
                t = torch.zeros(shp, dtype=torch.float64, pin_memory=True)
                out[k] = t.numpy()
                out["_pin_" + k] = t  # keeps the page-locked allocation alive
            else:
                out[k] = np.zeros(shp)
        return out

    def sample(self, run0: int, count: int, out=None) -> GraphProblem:
        """The sampled graphs of runs [run0, run0+count) as flat run data (what create_simulation would build).
        `out`: buffers from sample_buffers (at least `count` runs) to sample into instead of fresh arrays."""
        h = self._study()
        g = self.graph
        G, N, n = g.num_groups, g.num_nodes, g.n
        nct = _lib().mio_study_num_ctables(h)
        node_ct = np.zeros(N, dtype=np.int32)
        _check(_lib().mio_study_node_ctable(h, cdefs.iptr(node_ct)))
        if out is None:
            out = self.sample_buffers(count)
        psets, ctables, local, y0 = (out[k][:count] for k in ("psets", "ctables", "node_local", "y0"))
        mob = out["mobtables"][:count] if self.mob_max_pts > 0 else None
        draws = (C.c_uint64 * max(count, 1))()
        _check(_lib().mio_study_sample_runs(h, run0, count, cdefs.dptr(psets), cdefs.dptr(ctables), cdefs.dptr(local),
                                            cdefs.dptr(y0), cdefs.dptr(mob), draws))
        self.last_draws = np.array(draws[:count], dtype=np.uint64)
        return GraphProblem(model=g.model, num_groups=G, num_nodes=N, edge_start=g.edge_start, edge_end=g.edge_end,
                            edge_coeff=g.edge_coeff, psets=psets, ctables=ctables, node_local=local, y0=y0,
                            ct_max_loc=g.num_locations, ct_max_pts=self.ct_max_pts,
                            node_ctable=node_ct if nct > 1 else None, mobtables=mob, mob_max_pts=self.mob_max_pts,
                            integrator=self.integrator, t0=self.t0, dt_graph=self.dt, sim_kind=g.sim_kind,
                            vacc_partial=g.vacc_partial, vacc_full=g.vacc_full,
                            flow_simulation=int(getattr(self, "flow_simulation", 0)))

    # -- run ----------------------------------------------------------------------------------------------------
    def run(self, process_result: Optional[Callable] = None, batch: int = 1024, device: int = 0, rank: int = 0,
            world: int = 1, record_daily: bool = True):
        """ParameterStudy::run (parameter_studies.h:158-217): this rank's block of runs (distribute_runs), in batches.
        `process_result(run_idx, daily[N][days][n] or state[N][n])` is called per run in run order; its results are
        returned as a list (like the ensemble_result vector)."""
        from . import capi
        run0, count = distribute_runs(self.num_runs, world, rank)
        out = []
        sim = None
        days = int(np.floor(self.tmax + 1e-14) - np.ceil(self.t0 - 1e-14)) + 1
        done = 0
        while done < count:
            b = min(batch, count - done)
            p = self.sample(run0 + done, b)
            if sim is None:
                sim = capi.GraphSimulation(p, device=device, max_runs=min(batch, count), record_daily=record_daily,
                                           max_days=days, upload=False)
            sim.reset()
            sim.set_active_runs(b)
            sim.set_runs(0, p, 0, b)
            sim.advance(self.tmax)
            res = sim.daily(0, b) if record_daily else sim.state(0, b)
            for r in range(b):
                out.append(process_result(run0 + done + r, res[r]) if process_result else res[r])
            done += b
        if sim is not None:
            sim.close()
        # parameter_studies.h:339: future calls of run() produce different parameters
        self.rng_counter += self.num_runs << 32
        self._invalidate()
        return out

    def run_ensemble(self, p: Optional[Sequence[float]] = None, mean: bool = True, num_batches: int = 2,
                     device: int = 0, rank: int = 0, world: int = 1, mode: Optional[int] = None, pinned: bool = True,
                     day_block: Optional[int] = None):
        """ParameterStudy::run with the ensemble post-processing of save_results done on the device
        (parameter_studies.h:158-217 + data/analyze_result.h:289-343): this rank's block of runs (distribute_runs) is
        sampled, uploaded and advanced in `num_batches` batches, the daily results are recorded on the GPU and reduced
        to percentiles / mean at the end (exchange over torch.distributed when world > 1; result on rank 0).
        Pipeline: two host threads, each with its own simulation handle and CUDA stream, sample (C library, several
        threads), upload and advance alternate batches, so that uploads and the latency bound kernels of one batch
        overlap the other batch's work on the GPU. Returns (percentiles [len(p)][columns], mean [columns]).
        `day_block`: the ensemble record is sized for that many days only; all runs of the rank stay resident (one
        batch) and are advanced and reduced one day window at a time (studies whose full record does not fit the
        device, e.g. ODE-SECIRVVS with 12 500 runs per GPU: 110 GB)."""
        from concurrent.futures import ThreadPoolExecutor
        from . import capi, ensemble
        if p is None:
            p = ensemble.DEFAULT_PERCENTILES
        if mode is None:
            mode = ensemble.ENS_AGE_SUM
        if world > 1:
            # several ranks share the host: every rank samples with its share of the cores instead of all of them
            import os
            if "MIO_SAMPLE_THREADS" not in os.environ:
                os.environ["MIO_SAMPLE_THREADS"] = str(max(2, (os.cpu_count() or 2) // world))
        if day_block is not None:
            return self._run_ensemble_windows(p, mean, device, rank, world, mode, pinned, int(day_block))
        run0, count = distribute_runs(self.num_runs, world, rank)
        days = int(np.floor(self.tmax + 1e-14) - np.ceil(self.t0 - 1e-14)) + 1
        g = self.graph
        if count == 0:
            # more ranks than runs (distribute_runs gives this rank nothing): no sampling, no simulation, but the
            # rank still takes part in the collectives of the reduction with an empty run segment
            import torch
            El = flat.num_compartments(g.model) if mode == ensemble.ENS_AGE_SUM else g.n
            ncol = g.num_nodes * days * El + days * g.n
            dev = torch.device("cuda", device) if torch.cuda.is_available() else torch.device("cpu")
            out = ensemble.reduce_distributed(torch.empty((ncol, 0), dtype=torch.float64, device=dev), 0, rank, world,
                                              p, mean)
            self.rng_counter += self.num_runs << 32
            self._invalidate()
            return out
        nb = max(1, min(num_batches, count))
        bounds = [(i * count // nb, (i + 1) * count // nb) for i in range(nb)]
        import os as _os
        if (world > 1 or _os.environ.get("MIO_E2E_RAMP") == "1") and nb >= 2 and count >= 1000:
            # several ranks share the host's cores: sampling is 3-4 times slower per rank and the first batch's sampling
            # is not hidden behind any GPU work, so a small batch goes first (10 %, the rest in equal parts)
            first = count // 10
            rest = count - first
            bounds = [(0, first)] + [(first + i * rest // nb, first + (i + 1) * rest // nb) for i in range(nb)]
            nb += 1

        bmax = max(b - a for a, b in bounds)
        key = (count, bmax, days, mode, device, pinned)
        cache = getattr(self, "_ens", None)
        if cache is None or cache["key"] != key:
            self._drop_ensemble_cache()
            bufs = [self.sample_buffers(bmax, pinned) for _ in range(2)]
            first = self.sample(run0 + bounds[0][0], bounds[0][1] - bounds[0][0], out=bufs[0])
            sims = [capi.GraphSimulation(first, device=device, max_runs=bmax, upload=False) for _ in range(2)]
            rec = ensemble.EnsembleRecord(g.num_groups, g.num_nodes, days, capacity_runs=count, mode=mode, device=device,
                                          model=g.model)
            cache = self._ens = {"key": key, "bufs": bufs, "sims": sims, "rec": rec,
                                 "pool": ThreadPoolExecutor(max_workers=2)}
        bufs, sims, rec, pool = cache["bufs"], cache["sims"], cache["rec"], cache["pool"]
        h2d = 0

        import threading
        sample_lock = threading.Lock()  # one sampling call at a time: it uses all host cores itself

        def prepare(i):
            a, b = bounds[i]
            with sample_lock:
                pb = self.sample(run0 + a, b - a, out=bufs[i % 2])
            sim = sims[i % 2]
            sim.reset()
            sim.set_active_runs(b - a)
            rec.attach(sim, a)
            sim.set_runs(0, pb, 0, b - a)
            return sim, sum(getattr(pb, k).nbytes for k in ("psets", "ctables", "node_local", "y0"))

        import os
        import time
        trace = os.environ.get("MIO_CUDA_TRACE") is not None
        t_start = time.perf_counter()

        # two host threads, one per simulation handle (own CUDA stream): each samples, uploads and advances its
        # batches; the kernels of the two streams overlap on the GPU (the order dependent replay of early return
        # steps is latency bound and leaves most of the GPU idle on its own)
        def lane(w):
            total = 0
            for i in range(w, nb, 2):
                sim, nbytes = prepare(i)
                total += nbytes
                sim.advance(self.tmax)
            return total

        futs = [pool.submit(lane, w) for w in range(min(2, nb))]
        h2d = sum(f.result() for f in futs)
        t_wait, t_adv = 0.0, time.perf_counter() - t_start
        self.last_h2d_bytes = h2d
        t2 = time.perf_counter()
        pct, mn = ensemble.reduce_distributed(rec.as_torch(), count, rank, world, p, mean, rec=rec)
        if trace:
            import sys
            print(f"[run_ensemble] waiting for sampling+upload {1e3 * t_wait:.1f} ms, advance {1e3 * t_adv:.1f} ms, "
                  f"reduce {1e3 * (time.perf_counter() - t2):.1f} ms, total {1e3 * (time.perf_counter() - t_start):.1f} ms",
                  file=sys.stderr)
        # parameter_studies.h:339: future calls of run() produce different parameters
        self.rng_counter += self.num_runs << 32
        self._invalidate()
        return pct, mn

    def _run_ensemble_windows(self, p, mean, device, rank, world, mode, pinned, day_block):
        """run_ensemble with a windowed record: one batch, advance + reduce per window of `day_block` days."""
        from . import capi, ensemble
        run0, count = distribute_runs(self.num_runs, world, rank)
        days = int(np.floor(self.tmax + 1e-14) - np.ceil(self.t0 - 1e-14)) + 1
        g = self.graph
        G, N, n = g.num_groups, g.num_nodes, g.n
        El = flat.num_compartments(g.model) if mode == ensemble.ENS_AGE_SUM else n
        blk = max(1, min(day_block, days))
        key = ("windows", count, blk, days, mode, device, pinned)
        cache = getattr(self, "_ens", None)
        if count > 0 and (cache is None or cache["key"] != key):
            self._drop_ensemble_cache()
            bufs = [self.sample_buffers(count, pinned)]
            first = self.sample(run0, count, out=bufs[0])
            sims = [capi.GraphSimulation(first, device=device, max_runs=count, upload=False)]
            rec = ensemble.EnsembleRecord(G, N, blk, capacity_runs=count, mode=mode, device=device, model=g.model)
            cache = self._ens = {"key": key, "bufs": bufs, "sims": sims, "rec": rec, "pool": None, "sampled": True}
        if count > 0:
            sim, rec = cache["sims"][0], cache["rec"]
            pb = first if cache.pop("sampled", False) else self.sample(run0, count, out=cache["bufs"][0])
            sim.reset()
            sim.set_active_runs(count)
            rec.set_window(0)
            rec.attach(sim, 0)
            sim.set_runs(0, pb, 0, count)
            self.last_h2d_bytes = sum(getattr(pb, kk).nbytes for kk in ("psets", "ctables", "node_local", "y0"))
            local = rec.as_torch()
        else:
            import torch
            sim = None
            dev = torch.device("cuda", device) if torch.cuda.is_available() else torch.device("cpu")
            rec = _RecordShape(N, El, n)
            local = torch.empty((N * blk * El + blk * n, 0), dtype=torch.float64, device=dev)
        full = ensemble.windowed_advance_reduce(sim, rec, local, count, self.t0, self.tmax, days, blk, rank, world, p,
                                                mean)
        self.rng_counter += self.num_runs << 32
        self._invalidate()
        if rank != 0:
            return None, None
        return full[:len(p)], (full[len(p)] if mean else None)

    def _drop_ensemble_cache(self):
        cache = getattr(self, "_ens", None)
        if cache is not None:
            if cache["pool"] is not None:
                cache["pool"].shutdown(wait=True)
            for sim in cache["sims"]:
                sim.close()
            cache["rec"].close()
            self._ens = None

    def close(self):
        self._drop_ensemble_cache()
        self._invalidate()

    def __del__(self):
        try:
            self._invalidate()
        except Exception:
            pass

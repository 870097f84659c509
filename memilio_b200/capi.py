"""ctypes binding of memilio_b200/csrc/libmio_cuda.so (the C-ABI of include/mio_cuda.h).

There is no fallback: if the library is missing or no B200 is present every call raises MioCudaError.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

from . import cdefs, flat
from .problem import GraphProblem

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MIO_CUDA_LIB") or os.path.join(_HERE, "csrc", "libmio_cuda.so")  # tuning variants
_LIB = None

EXPORTS = [
    "mio_cuda_last_error_string", "mio_cuda_version", "mio_cuda_device_count", "mio_num_compartments",
    "mio_pset_age_stride", "mio_pset_size", "mio_ctable_size", "mio_mobtable_size", "mio_cuda_graph_create",
    "mio_cuda_graph_destroy", "mio_cuda_reset", "mio_cuda_set_runs", "mio_cuda_set_active_runs", "mio_cuda_advance", "mio_cuda_get_time",
    "mio_cuda_get_state", "mio_cuda_get_node_dt", "mio_cuda_get_counters", "mio_cuda_get_daily",
    "mio_cuda_get_trajectory", "mio_cuda_get_stats", "mio_cuda_get_stream", "mio_cuda_set_profiling",
    "mio_cuda_get_profile", "mio_cuda_device_state_ptr", "mio_cuda_measure_fp64_peak",
    "mio_study_create", "mio_study_destroy", "mio_study_num_ctables", "mio_study_node_ctable",
    "mio_study_sample_runs", "mio_study_distribute_runs", "mio_study_rng_raw", "mio_damping_cache_build",
    "mio_cuda_ensemble_create", "mio_cuda_ensemble_destroy", "mio_cuda_attach_ensemble", "mio_cuda_ensemble_info",
    "mio_cuda_ensemble_reduce", "mio_cuda_select_percentiles", "mio_cuda_reduce_host_columns",
    "mio_cuda_ensemble_set_window", "mio_cuda_comm_get_unique_id", "mio_cuda_comm_create", "mio_cuda_comm_destroy",
    "mio_cuda_ensemble_reduce_sharded", "mio_cuda_enable_mobility_results", "mio_cuda_get_mobility_results",
    "mio_cuda_get_flows", "mio_cuda_get_flow_trajectory", "mio_cuda_set_flows", "mio_cuda_set_contact_tables",
]
KERNEL_CLASSES = ["integrate", "depart", "return", "return_exact", "record", "ensemble"]


class MioCudaError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"mio_cuda status {code}: {msg}")
        self.code = code


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise MioCudaError(1, f"{LIB_PATH} not built (run `python -m memilio_b200.build`); there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        L.mio_cuda_last_error_string.restype = C.c_char_p
        for name in ("mio_pset_size", "mio_ctable_size", "mio_mobtable_size"):
            getattr(L, name).restype = C.c_int64
        L.mio_cuda_graph_create.argtypes = [C.POINTER(cdefs.GraphDesc), C.c_int, C.POINTER(C.c_void_p)]
        L.mio_cuda_graph_destroy.argtypes = [C.c_void_p]
        L.mio_cuda_graph_destroy.restype = None
        L.mio_cuda_set_runs.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(cdefs.RunData)]
        L.mio_cuda_set_active_runs.argtypes = [C.c_void_p, C.c_int32]
        L.mio_cuda_reset.argtypes = [C.c_void_p]
        L.mio_cuda_advance.argtypes = [C.c_void_p, C.c_double]
        L.mio_cuda_get_time.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        L.mio_cuda_get_state.argtypes = [C.c_void_p, C.c_int32, C.c_int32, cdefs.c_double_p]
        L.mio_cuda_get_node_dt.argtypes = [C.c_void_p, C.c_int32, C.c_int32, cdefs.c_double_p]
        L.mio_cuda_get_counters.argtypes = [C.c_void_p, C.c_int32, C.c_int32, cdefs.c_int64_p]
        L.mio_cuda_get_daily.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_int32), cdefs.c_double_p,
                                         C.c_int64]
        L.mio_cuda_get_trajectory.argtypes = [C.c_void_p, C.c_int32, C.c_int32, cdefs.c_int32_p, cdefs.c_double_p,
                                              cdefs.c_double_p]
        L.mio_cuda_get_stats.argtypes = [C.c_void_p, cdefs.c_int64_p, cdefs.c_int64_p]
        L.mio_cuda_get_stream.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
        L.mio_cuda_set_profiling.argtypes = [C.c_void_p, C.c_int]
        L.mio_cuda_get_profile.argtypes = [C.c_void_p, cdefs.c_double_p, cdefs.c_int64_p]
        L.mio_cuda_device_state_ptr.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]
        L.mio_cuda_measure_fp64_peak.argtypes = [C.c_int, C.c_double, C.POINTER(C.c_double)]
        _LIB = L
    return _LIB


def _check(rc):
    if rc != 0:
        raise MioCudaError(rc, lib().mio_cuda_last_error_string().decode())


class GraphSimulation:
    """`mio::make_mobility_sim(t0, dt, graph)` for an ensemble of runs on one GPU
    (cpp/memilio/mobility/metapopulation_mobility_instant.h:656-678, graph_simulation.h:95-121)."""

    def __init__(self, problem: GraphProblem, device: int = 0, max_runs: Optional[int] = None, record_daily=False,
                 max_days: int = 0, traj_capacity: int = 0, upload=True):
        self.p = problem
        self._h = C.c_void_p()
        self.max_runs = problem.num_runs if max_runs is None else max_runs
        self.traj_capacity = traj_capacity
        self.max_days = max_days
        d = problem.desc(self.max_runs, int(record_daily), max_days, traj_capacity)
        _check(lib().mio_cuda_graph_create(C.byref(d), device, C.byref(self._h)))
        if upload:
            self.set_runs(0, problem, 0, min(problem.num_runs, self.max_runs))

    def set_runs(self, slot0: int, problem: GraphProblem, run0: int, count: int):
        rd = problem.run_data(run0, count)
        _check(lib().mio_cuda_set_runs(self._h, slot0, count, C.byref(rd)))

    def reset(self):
        _check(lib().mio_cuda_reset(self._h))

    def set_active_runs(self, count: int):
        _check(lib().mio_cuda_set_active_runs(self._h, count))

    def advance(self, t_max: float):
        _check(lib().mio_cuda_advance(self._h, float(t_max)))

    @property
    def t(self) -> float:
        t = C.c_double()
        _check(lib().mio_cuda_get_time(self._h, C.byref(t)))
        return t.value

    def state(self, slot0=0, count=None) -> np.ndarray:
        count = self.max_runs - slot0 if count is None else count
        out = np.zeros((count, self.p.num_nodes, self.p.n))
        _check(lib().mio_cuda_get_state(self._h, slot0, count, cdefs.dptr(out)))
        return out

    def node_dt(self, slot0=0, count=None) -> np.ndarray:
        count = self.max_runs - slot0 if count is None else count
        out = np.zeros((count, self.p.num_nodes))
        _check(lib().mio_cuda_get_node_dt(self._h, slot0, count, cdefs.dptr(out)))
        return out

    def counters(self, slot0=0, count=None) -> np.ndarray:
        count = self.max_runs - slot0 if count is None else count
        out = np.zeros((count, self.p.num_nodes, flat.NUM_COUNTERS), dtype=np.int64)
        _check(lib().mio_cuda_get_counters(self._h, slot0, count, cdefs.lptr(out)))
        return out

    def daily(self, slot0=0, count=None) -> np.ndarray:
        count = self.max_runs - slot0 if count is None else count
        out = np.zeros((count, self.p.num_nodes, self.max_days, self.p.n))
        nd = C.c_int32()
        _check(lib().mio_cuda_get_daily(self._h, slot0, count, C.byref(nd), cdefs.dptr(out), out.size))
        return out.reshape(-1)[:count * self.p.num_nodes * nd.value * self.p.n].reshape(
            count, self.p.num_nodes, nd.value, self.p.n).copy()

    def trajectory(self, slot0=0, count=None):
        """(npts[count][N], times[count][N][cap], values[count][N][cap][n])"""
        count = self.max_runs - slot0 if count is None else count
        N, cap, n = self.p.num_nodes, self.traj_capacity, self.p.n
        npts = np.zeros((count, N), dtype=np.int32)
        times = np.zeros((count, N, cap))
        values = np.zeros((count, N, cap, n))
        _check(lib().mio_cuda_get_trajectory(self._h, slot0, count, cdefs.iptr(npts), cdefs.dptr(times),
                                             cdefs.dptr(values)))
        return npts, times, values

    def flows(self, slot0=0, count=None) -> np.ndarray:
        """FlowSimulation::get_flows().get_last_value() of every node: [count][N][G*15]"""
        count = self.max_runs - slot0 if count is None else count
        out = np.zeros((count, self.p.num_nodes, self.p.num_groups * 15))
        L = lib()
        L.mio_cuda_get_flows.argtypes = [C.c_void_p, C.c_int32, C.c_int32, cdefs.c_double_p]
        _check(L.mio_cuda_get_flows(self._h, slot0, count, cdefs.dptr(out)))
        return out

    def flow_trajectory(self, slot0=0, count=None) -> np.ndarray:
        """FlowSimulation::get_flows() as a series: [count][N][cap][G*15] (time points: see trajectory())"""
        count = self.max_runs - slot0 if count is None else count
        out = np.zeros((count, self.p.num_nodes, self.traj_capacity, self.p.num_groups * 15))
        L = lib()
        L.mio_cuda_get_flow_trajectory.argtypes = [C.c_void_p, C.c_int32, C.c_int32, cdefs.c_double_p]
        _check(L.mio_cuda_get_flow_trajectory(self._h, slot0, count, cdefs.dptr(out)))
        return out

    def enable_mobility_results(self, save_groups, capacity: int):
        """MobilityEdge(coeffs, save_indices) for every edge: record get_mobility_results()."""
        ptr = np.zeros(len(save_groups) + 1, dtype=np.int32)
        idx = []
        for i, g in enumerate(save_groups):
            idx += [int(x) for x in g]
            ptr[i + 1] = len(idx)
        idx = np.ascontiguousarray(idx if idx else [0], dtype=np.int32)
        self._mob_K, self._mob_cap = len(save_groups), capacity
        L = lib()
        L.mio_cuda_enable_mobility_results.argtypes = [C.c_void_p, C.c_int32, cdefs.c_int32_p, cdefs.c_int32_p, C.c_int32]
        _check(L.mio_cuda_enable_mobility_results(self._h, len(save_groups), cdefs.iptr(ptr), cdefs.iptr(idx), capacity))

    def mobility_results(self, slot0=0, count=None):
        """(npts[count][E], times[capacity], values[count][E][capacity][K+1])"""
        count = self.max_runs - slot0 if count is None else count
        E = self.p.num_edges
        npts = np.zeros((count, E), dtype=np.int32)
        times = np.zeros(self._mob_cap)
        values = np.zeros((count, E, self._mob_cap, self._mob_K + 1))
        L = lib()
        L.mio_cuda_get_mobility_results.argtypes = [C.c_void_p, C.c_int32, C.c_int32, cdefs.c_int32_p, cdefs.c_double_p,
                                                    cdefs.c_double_p]
        _check(L.mio_cuda_get_mobility_results(self._h, slot0, count, cdefs.iptr(npts), cdefs.dptr(times),
                                               cdefs.dptr(values)))
        return npts, times, values

    def stats(self):
        a = np.zeros(1, dtype=np.int64)
        b = np.zeros(1, dtype=np.int64)
        _check(lib().mio_cuda_get_stats(self._h, cdefs.lptr(a), cdefs.lptr(b)))
        return dict(launches=int(a[0]), exact_replays=int(b[0]))

    def stream_ptr(self) -> int:
        """cudaStream_t of the handle (wrap with torch.cuda.ExternalStream to time with torch events)."""
        sp = C.c_void_p()
        _check(lib().mio_cuda_get_stream(self._h, C.byref(sp)))
        return int(sp.value or 0)

    def set_profiling(self, on: bool):
        _check(lib().mio_cuda_set_profiling(self._h, int(on)))

    def profile(self):
        """{kernel class: (device ms, launches)} accumulated since set_profiling(True)."""
        ms = np.zeros(len(KERNEL_CLASSES))
        ln = np.zeros(len(KERNEL_CLASSES), dtype=np.int64)
        _check(lib().mio_cuda_get_profile(self._h, cdefs.dptr(ms), cdefs.lptr(ln)))
        return {k: (float(ms[i]), int(ln[i])) for i, k in enumerate(KERNEL_CLASSES)}

    def device_state_ptr(self):
        ptr = C.c_void_p()
        nd = C.c_int64()
        _check(lib().mio_cuda_device_state_ptr(self._h, C.byref(ptr), C.byref(nd)))
        return int(ptr.value), int(nd.value)

    def close(self):
        if self._h:
            lib().mio_cuda_graph_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def measure_fp64_peak(device: int = 0, seconds: float = 0.5) -> float:
    """Measured FP64 FMA throughput in TFLOP/s (DFMA chains on all SMs)."""
    out = C.c_double()
    _check(lib().mio_cuda_measure_fp64_peak(device, seconds, C.byref(out)))
    return out.value


def make_mobility_sim(problem: GraphProblem, **kw) -> GraphSimulation:
    return GraphSimulation(problem, **kw)

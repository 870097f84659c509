"""GraphProblem: numpy-side container of everything the C-ABI of include/mio_cuda.h needs for a graph ensemble.

It is the flat image of a `mio::Graph<SimulationNode<Sim>, MobilityEdge>` (cpp/memilio/mobility/graph.h:145-300) plus
the per-run sampled data that `ParameterStudy::run` would hand to `make_sampled_graph_simulation`
(cpp/memilio/compartments/parameter_studies.h:383-399).
"""
from __future__ import annotations

import dataclasses
from typing import Optional

import numpy as np

from . import cdefs, flat


@dataclasses.dataclass
class GraphProblem:
    model: int
    num_groups: int
    num_nodes: int
    edge_start: np.ndarray  # [E] int32, reference edge order (sorted by (start, end))
    edge_end: np.ndarray  # [E] int32
    edge_coeff: np.ndarray  # [E][n]
    psets: np.ndarray  # [R][P][psz]
    ctables: np.ndarray  # [R][Ct][csz]
    node_local: np.ndarray  # [R][N][2]
    y0: np.ndarray  # [R][N][n]
    ct_max_loc: int = 1
    ct_max_pts: int = 4
    node_pset: Optional[np.ndarray] = None  # [N] int32
    node_ctable: Optional[np.ndarray] = None  # [N] int32
    dt0: Optional[np.ndarray] = None  # [R][N]
    mobtables: Optional[np.ndarray] = None  # [R][msz]
    mob_max_pts: int = 0
    integrator: cdefs.IntegratorDesc = dataclasses.field(default_factory=cdefs.integrator_desc)
    t0: float = 0.0
    dt_graph: float = 0.5
    sim_kind: int = flat.SIM_MODEL
    vacc_partial: Optional[np.ndarray] = None  # osecirvvs: [N][G][days] cumulative DailyPartialVaccinations
    vacc_full: Optional[np.ndarray] = None  # osecirvvs: [N][G][days] cumulative DailyFullVaccinations
    flow_simulation: int = 0  # 1: the nodes are mio::FlowSimulation (compartments/flow_simulation.h:37-97), osecir only

    def __post_init__(self):
        self.edge_start = np.ascontiguousarray(self.edge_start, dtype=np.int32)
        self.edge_end = np.ascontiguousarray(self.edge_end, dtype=np.int32)
        n = self.n
        self.edge_coeff = np.ascontiguousarray(self.edge_coeff, dtype=np.float64).reshape(self.num_edges, n)
        order = np.lexsort((self.edge_end, self.edge_start))
        if not np.array_equal(order, np.arange(self.num_edges)):
            raise ValueError("edges must be sorted by (start, end) like Graph::add_edge keeps them (graph.h:233-241)")
        for name in ("psets", "ctables", "node_local", "y0"):
            setattr(self, name, np.ascontiguousarray(getattr(self, name), dtype=np.float64))
        R = self.num_runs
        assert self.psets.ndim == 3 and self.psets.shape[2] == flat.pset_size(self.model, self.num_groups)
        assert self.ctables.ndim == 3 and self.ctables.shape[0] == R
        assert self.ctables.shape[2] == flat.ctable_size(self.num_groups, self.ct_max_loc, self.ct_max_pts)
        assert self.node_local.shape == (R, self.num_nodes, flat.NODE_LOCAL)
        assert self.y0.shape == (R, self.num_nodes, n)
        if self.node_pset is not None:
            self.node_pset = np.ascontiguousarray(self.node_pset, dtype=np.int32)
        if self.node_ctable is not None:
            self.node_ctable = np.ascontiguousarray(self.node_ctable, dtype=np.int32)
        if self.dt0 is not None:
            self.dt0 = np.ascontiguousarray(self.dt0, dtype=np.float64).reshape(R, self.num_nodes)
        if self.mobtables is not None:
            self.mobtables = np.ascontiguousarray(self.mobtables, dtype=np.float64).reshape(R, -1)
            assert self.mobtables.shape[1] == flat.mobtable_size(n, self.mob_max_pts)
        for name in ("vacc_partial", "vacc_full"):
            if getattr(self, name) is not None:
                setattr(self, name, np.ascontiguousarray(getattr(self, name), dtype=np.float64).reshape(
                    self.num_nodes, self.num_groups, -1))

    @property
    def n(self) -> int:
        return self.num_groups * flat.num_compartments(self.model)

    @property
    def num_edges(self) -> int:
        return int(self.edge_start.shape[0])

    @property
    def num_runs(self) -> int:
        return int(self.y0.shape[0])

    @property
    def num_psets(self) -> int:
        return int(self.psets.shape[1])

    @property
    def num_ctables(self) -> int:
        return int(self.ctables.shape[1])

    def desc(self, max_runs: Optional[int] = None, record_daily: int = 0, max_days: int = 0,
             traj_capacity: int = 0) -> cdefs.GraphDesc:
        d = cdefs.GraphDesc()
        d.model = self.model
        d.sim_kind = self.sim_kind
        d.num_groups = self.num_groups
        d.num_nodes = self.num_nodes
        d.num_edges = self.num_edges
        d.num_psets = self.num_psets
        d.num_ctables = self.num_ctables
        d.ct_max_loc = self.ct_max_loc
        d.ct_max_pts = self.ct_max_pts
        d.mob_max_pts = self.mob_max_pts if self.mobtables is not None else 0
        d.max_runs = self.num_runs if max_runs is None else max_runs
        d.record_daily = record_daily
        d.max_days = max_days
        d.traj_capacity = traj_capacity
        d.edge_start = cdefs.iptr(self.edge_start)
        d.edge_end = cdefs.iptr(self.edge_end)
        d.edge_coeff = cdefs.dptr(self.edge_coeff)
        d.node_pset = cdefs.iptr(self.node_pset)
        d.node_ctable = cdefs.iptr(self.node_ctable)
        d.integrator = self.integrator
        d.t0 = self.t0
        d.dt_graph = self.dt_graph
        d.flow_simulation = self.flow_simulation
        if self.vacc_partial is not None:
            d.vacc_days = int(self.vacc_partial.shape[2])
            d.vacc_partial = cdefs.dptr(self.vacc_partial)
            d.vacc_full = cdefs.dptr(self.vacc_full)
        return d

    def run_data(self, run0: int = 0, count: Optional[int] = None) -> cdefs.RunData:
        count = self.num_runs - run0 if count is None else count
        sl = slice(run0, run0 + count)
        rd = cdefs.RunData()
        # slices of C-contiguous arrays along axis 0 are C-contiguous views; keep them alive on the struct
        keep = [self.psets[sl], self.ctables[sl], self.node_local[sl], self.y0[sl],
                None if self.dt0 is None else self.dt0[sl], None if self.mobtables is None else self.mobtables[sl]]
        rd.psets, rd.ctables, rd.node_local, rd.y0, rd.dt0, rd.mobtables = [cdefs.dptr(a) for a in keep]
        rd._keep = keep
        return rd


def single_node_problem(G, pset, local, ctable, y0, *, ct_max_loc, ct_max_pts, integrator=None, t0=0.0, dt0=0.1,
                        dt_graph=float("inf"), model=flat.MODEL_OSECIR, sim_kind=flat.SIM_MODEL, vacc_partial=None,
                        vacc_full=None) -> GraphProblem:
    """A 1-node, 0-edge graph: `mio::simulate` expressed as a graph run (make_no_mobility_sim, instant.h:700-712)."""
    n = G * flat.num_compartments(model)
    return GraphProblem(model=model, num_groups=G, num_nodes=1, edge_start=np.zeros(0, np.int32),
                        edge_end=np.zeros(0, np.int32), edge_coeff=np.zeros((0, n)),
                        psets=np.asarray(pset, dtype=np.float64).reshape(1, 1, -1),
                        ctables=np.asarray(ctable, dtype=np.float64).reshape(1, 1, -1),
                        node_local=np.asarray(local, dtype=np.float64).reshape(1, 1, 2),
                        y0=np.asarray(y0, dtype=np.float64).reshape(1, 1, n), ct_max_loc=ct_max_loc,
                        ct_max_pts=ct_max_pts, dt0=np.full((1, 1), dt0),
                        integrator=integrator if integrator is not None else cdefs.integrator_desc(), t0=t0,
                        dt_graph=dt_graph, sim_kind=sim_kind, vacc_partial=vacc_partial, vacc_full=vacc_full)

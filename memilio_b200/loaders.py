"""Loaders: data files and parameter graphs -> the flat descriptors of include/mio_cuda.h.

Host-side mirror of the reference's graph construction helpers:
  * read_mobility_plain      cpp/memilio/io/mobility_io.cpp:123-159 (square matrix, one row per line, blank separated)
  * read_contact_matrices    the `baseline_<location>.txt` / `minimum_<location>.txt` files of data/Germany/contacts as
                             read by the simulation scripts (cpp/simulations/*.cpp set_contact_matrices)
  * set_edges                cpp/memilio/mobility/graph.h:402-446 (commuter matrix -> MobilityCoefficientGroup baseline
                             of every county pair above the 4e-5 threshold, mobile compartments x commuting weights)
  * set_node_uncertainty     cpp/memilio/mobility/graph.h:346-380 (TestAndTraceCapacity = factor * population with a
                             uniform 0.8..1.2 distribution, +-10 % uniform uncertainty on every compartment)
The result is a study.StudyGraph (the parameter graph a ParameterStudy samples from). Reading the RKI / DIVI / population
JSON files (set_nodes' read_func) is outside the hot path's scope: initial populations come in as an array.
"""
from __future__ import annotations

import os
from typing import Iterable, Optional, Sequence

import numpy as np

from . import flat, study

CONTACT_LOCATIONS = ("home", "school_pf_eig", "work", "other")  # ContactLocation enum of the simulation scripts


def read_mobility_plain(path: str) -> np.ndarray:
    """The commuter matrix of a `*.txt` file: absolute numbers of commuters from county i (row) to county j."""
    rows = []
    with open(path) as f:
        for line in f:
            if line.strip() == "" and not rows:
                continue
            rows.append([float(w) for w in line.split(" ") if w.strip() != ""])
    if not rows:
        return np.zeros((0, 0))
    n = len(rows)
    for r in rows:
        if len(r) != n:
            raise ValueError(f"{path}: Not a square matrix.")  # StatusCode::InvalidFileFormat
    return np.asarray(rows, dtype=np.float64)


def read_contact_matrices(directory: str, locations: Sequence[str] = CONTACT_LOCATIONS):
    """(baseline[L][G][G], minimum[L][G][G]) from baseline_<loc>.txt / minimum_<loc>.txt."""
    base, mini = [], []
    for loc in locations:
        b = read_mobility_plain(os.path.join(directory, f"baseline_{loc}.txt"))
        m = read_mobility_plain(os.path.join(directory, f"minimum_{loc}.txt"))
        if b.shape != m.shape:
            raise ValueError(f"contact matrices of location {loc} differ in size")
        base.append(b)
        mini.append(m)
    return np.stack(base), np.stack(mini)


def set_edges(populations: np.ndarray, mobility: np.ndarray, mobile_compartments: Iterable[int],
              commuting_weights: Optional[Sequence[float]] = None, num_compartments: int = 10,
              threshold: float = 4e-5):
    """graph.h:402-446. populations[N][G][C] (values), mobility[N][N] absolute commuters. Returns (edge_start,
    edge_end, edge_coeff[E][G*C]) in the reference's edge order (sorted by (start, end))."""
    pop = np.asarray(populations, dtype=np.float64)
    N, G, C = pop.shape
    assert C == num_compartments
    mob = np.asarray(mobility, dtype=np.float64)
    if mob.shape != (N, N):
        raise ValueError("Mobility matrices do not have the correct size. You may need to run "
                         "transformMobilitydata.py from pycode memilio epidata package.")
    w = np.ones(G) if commuting_weights is None or len(commuting_weights) == 0 else np.asarray(commuting_weights, float)
    mobile = list(mobile_compartments)
    es, ee, coeffs = [], [], []
    for i in range(N):
        working_population = 0.0
        for age in range(G):  # get_group_total sums the compartments in index order
            total = 0.0
            for c in range(C):
                total += pop[i, age, c]
            working_population += total * w[age]
        for j in range(N):
            commuter_coeff_ij = mob[i, j] / working_population
            if commuter_coeff_ij > threshold:
                v = np.zeros(G * C)
                for age in range(G):
                    for c in mobile:
                        v[age * C + c] = commuter_coeff_ij * w[age]
                es.append(i)
                ee.append(j)
                coeffs.append(v)
    return (np.asarray(es, dtype=np.int32), np.asarray(ee, dtype=np.int32),
            np.asarray(coeffs, dtype=np.float64).reshape(len(es), G * C))


def set_node_uncertainty(populations: np.ndarray, tnt_capacity_factor: float):
    """graph.h:346-380 for every node: (node_local UV[N][2], populations UV[N][G*C]). ICUCapacity stays a plain value
    (the caller's), TestAndTraceCapacity ~ U(0.8, 1.2) * factor * population, compartments ~ U(0.9, 1.1) * value."""
    pop = np.asarray(populations, dtype=np.float64)
    N = pop.shape[0]
    flat_pop = pop.reshape(N, -1)
    local = study.uv_array((N, 2))
    pops = study.uv_array(flat_pop.shape)
    for k in range(N):
        total = 0.0
        for v in flat_pop[k]:
            total += v
        cap = total * tnt_capacity_factor
        study.set_uniform(local, (k, 1), 0.8 * cap, 1.2 * cap)
        local[k, 1]["value"] = 0.5 * (1.2 * cap + 0.8 * cap)
        for j, v in enumerate(flat_pop[k]):
            study.set_uniform(pops, (k, j), 0.9 * v, 1.1 * v)
            pops[k, j]["value"] = 0.5 * (1.1 * v + 0.9 * v)
    return local, pops


def study_graph_from_files(populations: np.ndarray, mobility_file: str, contacts_dir: str, header: np.ndarray,
                           age_params: np.ndarray, icu_capacity: np.ndarray, tnt_capacity_factor: float,
                           mobile_compartments=(flat.S, flat.E, flat.INS, flat.ISY, flat.R),
                           commuting_weights=(0.0, 0.0, 1.0, 1.0, 0.33, 0.0), dampings=(),
                           locations: Sequence[str] = CONTACT_LOCATIONS) -> study.StudyGraph:
    """A Germany-style parameter graph from data files: contact matrices per location, commuter matrix -> edges,
    node uncertainty (the mobility baseline sits on ContactLocation::Work like set_edges puts it)."""
    pop = np.asarray(populations, dtype=np.float64)
    N, G, C = pop.shape
    baseline, minimum = read_contact_matrices(contacts_dir, locations)
    if baseline.shape[1] != G:
        raise ValueError("contact matrices and populations differ in the number of age groups")
    es, ee, coeff = set_edges(pop, read_mobility_plain(mobility_file), mobile_compartments, commuting_weights, C)
    local, pops = set_node_uncertainty(pop, tnt_capacity_factor)
    for k in range(N):
        study.set_value(local, (k, 0), float(np.asarray(icu_capacity).reshape(-1)[k % np.asarray(icu_capacity).size]))
    return study.StudyGraph(num_groups=G, num_nodes=N, header=header, age_params=age_params, node_local=local,
                            populations=pops, baseline=baseline, minimum=minimum, edge_start=es, edge_end=ee,
                            edge_coeff=coeff, dampings=list(dampings),
                            mobility_location=list(locations).index("work") if "work" in locations else 0)

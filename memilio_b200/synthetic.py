"""Synthetic Germany-shaped county graph (BASELINE.json configs[2..4], SURVEY.md section 8d "Config 3").

No mobility or population data ships with the reference (data/README.md), so the graph is generated: log-normal
county populations, gravity-law commuter matrix, the reference's own edge threshold and coefficient formula
(`set_edges`, cpp/memilio/mobility/graph.h:402-440: coefficient = commuters / working population * age weight for the
mobile compartments, edge kept iff commuter share > 4e-5).
"""
from __future__ import annotations

from typing import Optional

import numpy as np

from . import cdefs, flat
from .flat import ContactLocation
from .problem import GraphProblem

AGE_SHARES = np.array([0.047, 0.092, 0.225, 0.345, 0.215, 0.076])
COMMUTING_WEIGHTS = np.array([0.0, 0.0, 1.0, 1.0, 0.33, 0.0])  # pycode/.../bindings/models/osecir.cpp:194
MOBILE_COMPARTMENTS = (flat.S, flat.E, flat.INS, flat.ISY, flat.R)  # pycode/.../bindings/models/osecir.cpp:191-193


def county_graph(N: int = 400, seed: int = 20261017, max_out_degree: int = 25, G: int = 6,
                 commuter_share: float = 0.3):
    """Returns (population[N][G], edge_start, edge_end, edge_coeff[E][G*10]) in reference edge order."""
    rng = np.random.default_rng(seed)
    pop_total = np.round(np.exp(rng.normal(11.9, 0.75, N)))
    shares = AGE_SHARES[:G] / AGE_SHARES[:G].sum()
    weights = COMMUTING_WEIGHTS[:G] if G == 6 else np.ones(G)
    pop = pop_total[:, None] * shares[None, :]
    xy = rng.uniform(0, 1, (N, 2)) * np.array([640.0, 870.0])
    d = np.linalg.norm(xy[:, None, :] - xy[None, :, :], axis=2)
    np.fill_diagonal(d, np.inf)
    working = (pop * weights[None, :]).sum(axis=1)
    grav = pop_total[:, None] * pop_total[None, :] / np.maximum(d, 5.0) ** 2
    m = grav / grav.sum(axis=1, keepdims=True) * (commuter_share * working)[:, None]  # absolute commuters i -> j
    coeff_ij = m / working[:, None]
    keep = coeff_ij > 4e-5
    if max_out_degree is not None:
        for i in range(N):
            idx = np.flatnonzero(keep[i])
            if idx.size > max_out_degree:
                drop = idx[np.argsort(coeff_ij[i, idx])[:-max_out_degree]]
                keep[i, drop] = False
    es, ee = np.nonzero(keep)  # row-major => sorted by (start, end)
    n = G * 10
    coeff = np.zeros((es.size, n))
    for a in range(G):
        for c in MOBILE_COMPARTMENTS:
            coeff[:, a * 10 + c] = coeff_ij[es, ee] * weights[a]
    return pop, es.astype(np.int32), ee.astype(np.int32), coeff


def covid_midpoint_pset(G: int = 6, seasonality: float = 0.2) -> np.ndarray:
    """Midpoints of the uniform ranges of set_covid_parameters
    (cpp/examples/ode_secir_parameter_study_graph.cpp:91-159)."""
    mid = lambda lo, hi: 0.5 * (np.asarray(hi, dtype=np.float64) + np.asarray(lo, dtype=np.float64))
    six = lambda a: np.asarray(a, dtype=np.float64)[:G] if G <= 6 else np.resize(a, G)
    return flat.osecir_pset(
        G, StartDay=0.0, Seasonality=seasonality,
        TimeExposed=mid(2.67, 4.0), TimeInfectedNoSymptoms=mid(1.2, 2.53),
        TimeInfectedSymptoms=six(mid([5.6255, 5.6255, 5.6646, 5.5631, 5.501, 5.465],
                                     [8.427, 8.427, 8.4684, 8.3139, 8.169, 8.085])),
        TimeInfectedSevere=six(mid([3.925, 3.925, 4.85, 6.4, 7.2, 9.], [6.075, 6.075, 7., 8.7, 9.8, 13.])),
        TimeInfectedCritical=six(mid([4.95, 4.95, 4.86, 14.14, 14.4, 10.], [8.95, 8.95, 8.86, 20.58, 19.8, 13.2])),
        TransmissionProbabilityOnContact=six(mid([0.02, 0.05, 0.05, 0.05, 0.08, 0.15],
                                                 [0.04, 0.07, 0.07, 0.07, 0.10, 0.20])),
        RelativeTransmissionNoSymptoms=1.0, RiskOfInfectionFromSymptomatic=mid(0.1, 0.3),
        MaxRiskOfInfectionFromSymptomatic=mid(0.3, 0.5),
        RecoveredPerInfectedNoSymptoms=six(mid([0.2, 0.2, 0.15, 0.15, 0.15, 0.15], [0.3, 0.3, 0.25, 0.25, 0.25, 0.25])),
        SeverePerInfectedSymptoms=six(mid([0.006, 0.006, 0.015, 0.049, 0.15, 0.20],
                                          [0.009, 0.009, 0.023, 0.074, 0.18, 0.25])),
        CriticalPerSevere=six(mid([0.05, 0.05, 0.05, 0.10, 0.25, 0.35], [0.10, 0.10, 0.10, 0.20, 0.35, 0.45])),
        DeathsPerCritical=six(mid([0.00, 0.00, 0.10, 0.10, 0.30, 0.5], [0.10, 0.10, 0.18, 0.18, 0.50, 0.7])))


def county_problem(N: int = 400, runs: int = 1, seed: int = 20261017, G: int = 6, max_out_degree: int = 25,
                   n_seeded: int = 10, contact_damping: Optional[float] = 0.6, damping_day: float = 30.0,
                   cont_freq: float = 10.0, perturb: float = 0.0, tnt_capacity_factor: float = 7.5e-5,
                   integrator: Optional[cdefs.IntegratorDesc] = None, dt_graph: float = 0.5) -> GraphProblem:
    """Config 3 (runs=1) / config 4 shaped inputs. `perturb` > 0 draws run-specific parameters with numpy (uniform
    +-perturb around the midpoints) -- used by parity tests that only need *some* per-run variation; the bit-exact
    MEmilio sampling path is memilio_b200.sampling."""
    pop, es, ee, coeff = county_graph(N, seed, max_out_degree, G)
    n = G * 10
    rng = np.random.default_rng(seed + 1)
    y0 = np.zeros((N, G, 10))
    y0[:, :, flat.S] = pop
    seeded = rng.choice(N, size=min(n_seeded, N), replace=False)
    for k in seeded:
        e0 = 1e-4 * pop[k]
        y0[k, :, flat.E] = e0
        y0[k, :, flat.S] -= e0
    local = np.zeros((N, 2))
    local[:, 0] = flat.DBL_MAX  # ICUCapacity default
    local[:, 1] = pop.sum(axis=1) * tnt_capacity_factor
    pset = covid_midpoint_pset(G)
    base = np.full((G, G), cont_freq / G)
    locs = [ContactLocation(base, None, [damping_day], [np.full((G, G), contact_damping)])] \
        if contact_damping is not None else [ContactLocation(base)]
    ct = flat.contact_table(G, locs, 1, 2)
    psets = np.broadcast_to(pset, (runs, 1, pset.size)).copy()
    ctables = np.broadcast_to(ct, (runs, 1, ct.size)).copy()
    locals_ = np.broadcast_to(local, (runs, N, 2)).copy()
    y0s = np.broadcast_to(y0.reshape(N, n), (runs, N, n)).copy()
    if perturb > 0.0:
        prng = np.random.default_rng(seed + 2)
        rec = psets[:, 0, flat.PSET_HEADER:].reshape(runs, G, flat.OSECIR_AGE_STRIDE)
        rec[:, :, :14] *= prng.uniform(1 - perturb, 1 + perturb, (runs, G, 14))
        psets[:, 0, flat.PH["Seasonality"]] = prng.uniform(0.1, 0.3, runs)
        locals_[:, :, 1] *= prng.uniform(0.8, 1.2, (runs, N))
        y0s *= prng.uniform(0.9, 1.1, (runs, N, n))
        if contact_damping is not None:
            for r in range(runs):
                d = prng.uniform(0.4, 0.8)
                ctables[r, 0] = flat.contact_table(
                    G, [ContactLocation(base, None, [damping_day], [np.full((G, G), d)])], 1, 2)
    return GraphProblem(model=flat.MODEL_OSECIR, num_groups=G, num_nodes=N, edge_start=es, edge_end=ee,
                        edge_coeff=coeff, psets=psets, ctables=ctables, node_local=locals_, y0=y0s, ct_max_loc=1,
                        ct_max_pts=2, integrator=integrator if integrator is not None else cdefs.integrator_desc(),
                        t0=0.0, dt_graph=dt_graph)


# pycode/memilio-simulation/memilio/simulation/bindings/models/osecirvvs.cpp:220-231
VVS_MOBILE_COMPARTMENTS = (flat.V_SN, flat.V_EN, flat.V_INSN, flat.V_ISYN, flat.V_SII, flat.V_SPI, flat.V_EPI,
                           flat.V_INSPI, flat.V_ISYPI, flat.V_EII, flat.V_INSII, flat.V_ISYII)
VVS_INFECTED = (flat.V_EN, flat.V_EPI, flat.V_EII, flat.V_INSN, flat.V_INSNC, flat.V_INSPI, flat.V_INSPIC,
                flat.V_INSII, flat.V_INSIIC, flat.V_ISYN, flat.V_ISYNC, flat.V_ISYPI, flat.V_ISYPIC, flat.V_ISYII,
                flat.V_ISYIIC, flat.V_ISEVN, flat.V_ISEVII, flat.V_ISEVPI, flat.V_ICRN, flat.V_ICRPI, flat.V_ICRII)


def benchmark_vvs_pset(G: int = 6) -> np.ndarray:
    """Parameters of cpp/benchmarks/graph_simulation.cpp:31-112 (create_model)."""
    return flat.osecirvvs_pset(
        G, Seasonality=0.2, TimeExposed=3.33, TimeInfectedNoSymptoms=1.87, TimeInfectedSymptoms=7.0,
        TimeInfectedSevere=6.0, TimeInfectedCritical=7.0, TransmissionProbabilityOnContact=0.15,
        RelativeTransmissionNoSymptoms=0.5, RiskOfInfectionFromSymptomatic=0.0, MaxRiskOfInfectionFromSymptomatic=0.4,
        RecoveredPerInfectedNoSymptoms=0.2, SeverePerInfectedSymptoms=0.1, CriticalPerSevere=0.1,
        DeathsPerCritical=0.1, ReducExposedPartialImmunity=0.8, ReducExposedImprovedImmunity=0.331,
        ReducInfectedSymptomsPartialImmunity=0.65, ReducInfectedSymptomsImprovedImmunity=0.243,
        ReducInfectedSevereCriticalDeadPartialImmunity=0.1, ReducInfectedSevereCriticalDeadImprovedImmunity=0.091,
        ReducTimeInfectedMild=0.9)


def county_problem_vvs(N: int = 400, runs: int = 1, seed: int = 20261017, G: int = 6, max_out_degree: int = 25,
                       days: int = 100, sim_kind: int = flat.SIM_MODEL, perturb: float = 0.0,
                       damping_day: float = 5.0, daily_vaccinations: float = 5.0, init_per_10k: float = 20.0,
                       icu_per_10k: float = 100.0 / 6, start_day_new_variant: Optional[float] = None,
                       integrator: Optional[cdefs.IntegratorDesc] = None, dt_graph: float = 0.5) -> GraphProblem:
    """BASELINE config 5 shaped inputs: the synthetic county graph with ODE-SECIRVVS nodes carrying the parameters and
    the initial-value pattern of the reference's graph benchmark (cpp/benchmarks/graph_simulation.cpp:31-112), scaled
    from its 10 000 people per age group to the county's age groups; cumulative vaccination arrays 5*(d+1)
    (SURVEY.md 8d config 5). sim_kind SIM_MODEL = osecirvvs::Simulation nodes, SIM_BASE = the plain mio::Simulation
    the reference benchmark instantiates."""
    pop, es, ee, coeff10 = county_graph(N, seed, max_out_degree, G)
    C = flat.OSECIRVVS_NCOMP
    n = G * C
    coeff = np.zeros((es.size, G, C))
    w_age = coeff10.reshape(es.size, G, 10)[:, :, flat.S]
    for c in VVS_MOBILE_COMPARTMENTS:
        coeff[:, :, c] = w_age
    y0 = np.zeros((N, G, C))
    scale = pop / 10000.0
    for c in VVS_INFECTED:
        y0[:, :, c] = init_per_10k * scale
    y0[:, :, flat.V_SII] = pop / 3
    y0[:, :, flat.V_SPI] = pop / 3
    y0[:, :, flat.V_SN] = pop - y0.sum(axis=2)
    local = np.zeros((N, 2))
    local[:, 0] = icu_per_10k * pop.sum(axis=1) / 10000.0
    local[:, 1] = 0.0143 * pop.sum(axis=1) / 10000.0
    kw = {} if start_day_new_variant is None else dict(StartDayNewVariant=start_day_new_variant,
                                                        InfectiousnessNewVariant=1.4)
    pset = benchmark_vvs_pset(G)
    if kw:
        pset[flat.PHV["StartDayNewVariant"]] = kw["StartDayNewVariant"]
        pset[flat.PSET_HEADER:].reshape(G, flat.OSECIRVVS_AGE_STRIDE)[:, flat.PV["InfectiousnessNewVariant"]] = 1.4
    base = np.full((G, G), 0.5)
    np.fill_diagonal(base, 5.0)
    ct = flat.contact_table(G, [ContactLocation(base, None, [damping_day], [np.full((G, G), 0.3)])], 1, 2)
    psets = np.broadcast_to(pset, (runs, 1, pset.size)).copy()
    ctables = np.broadcast_to(ct, (runs, 1, ct.size)).copy()
    locals_ = np.broadcast_to(local, (runs, N, 2)).copy()
    y0s = np.broadcast_to(y0.reshape(N, n), (runs, N, n)).copy()
    if perturb > 0.0:
        prng = np.random.default_rng(seed + 2)
        rec = psets[:, 0, flat.PSET_HEADER:].reshape(runs, G, flat.OSECIRVVS_AGE_STRIDE)
        rec[:, :, :21] *= prng.uniform(1 - perturb, 1 + perturb, (runs, G, 21))
        psets[:, 0, flat.PHV["Seasonality"]] = prng.uniform(0.1, 0.3, runs)
        locals_ *= prng.uniform(0.8, 1.2, (runs, N, 2))
        y0s *= prng.uniform(0.9, 1.1, (runs, N, n))
        for r in range(runs):
            flat.osecirvvs_apply_constraints(psets[r, 0], G)
    vacc_days = days + 2
    vacc = np.broadcast_to(daily_vaccinations * (np.arange(vacc_days) + 1.0), (N, G, vacc_days)).copy()
    vacc *= scale[:, :, None]
    return GraphProblem(model=flat.MODEL_OSECIRVVS, num_groups=G, num_nodes=N, edge_start=es, edge_end=ee,
                        edge_coeff=coeff.reshape(es.size, n), psets=psets, ctables=ctables, node_local=locals_, y0=y0s,
                        ct_max_loc=1, ct_max_pts=2,
                        integrator=integrator if integrator is not None else cdefs.integrator_desc(), t0=0.0,
                        dt_graph=dt_graph, sim_kind=sim_kind, vacc_partial=vacc, vacc_full=vacc.copy())


# uniform ranges of set_covid_parameters, cpp/examples/ode_secir_parameter_study_graph.cpp:91-159
COVID_UNIFORM = dict(
    TimeExposed=(2.67, 4.0), TimeInfectedNoSymptoms=(1.2, 2.53),
    TimeInfectedSymptoms=([5.6255, 5.6255, 5.6646, 5.5631, 5.501, 5.465], [8.427, 8.427, 8.4684, 8.3139, 8.169, 8.085]),
    TimeInfectedSevere=([3.925, 3.925, 4.85, 6.4, 7.2, 9.], [6.075, 6.075, 7., 8.7, 9.8, 13.]),
    TimeInfectedCritical=([4.95, 4.95, 4.86, 14.14, 14.4, 10.], [8.95, 8.95, 8.86, 20.58, 19.8, 13.2]),
    TransmissionProbabilityOnContact=([0.02, 0.05, 0.05, 0.05, 0.08, 0.15], [0.04, 0.07, 0.07, 0.07, 0.10, 0.20]),
    RelativeTransmissionNoSymptoms=(1.0, 1.0),
    RiskOfInfectionFromSymptomatic=(0.1, 0.3), MaxRiskOfInfectionFromSymptomatic=(0.3, 0.5),
    RecoveredPerInfectedNoSymptoms=([0.2, 0.2, 0.15, 0.15, 0.15, 0.15], [0.3, 0.3, 0.25, 0.25, 0.25, 0.25]),
    SeverePerInfectedSymptoms=([0.006, 0.006, 0.015, 0.049, 0.15, 0.20], [0.009, 0.009, 0.023, 0.074, 0.18, 0.25]),
    CriticalPerSevere=([0.05, 0.05, 0.05, 0.10, 0.25, 0.35], [0.10, 0.10, 0.10, 0.20, 0.35, 0.45]),
    DeathsPerCritical=([0.00, 0.00, 0.10, 0.10, 0.30, 0.5], [0.10, 0.10, 0.18, 0.18, 0.50, 0.7]),
)


def covid_study_graph(N: int = 400, seed: int = 20261017, G: int = 6, max_out_degree: int = 25, n_seeded: int = 10,
                      cont_freq: float = 10.0, damping_day: Optional[float] = 30.0, pop_uncertainty: float = 0.1,
                      tnt_capacity_factor: float = 7.5e-5):
    """Config 4: the synthetic county graph as a *parameter graph* with the uniform distributions of
    set_covid_parameters, +-10 % uniform uncertainty on the seeded compartments and 0.8-1.2 on the test-and-trace
    capacity (cpp/memilio/mobility/graph.h:350-351,376-378), one sampled contact damping U(0.4, 0.8) at
    `damping_day` that also damps the mobility coefficients (parameter_space.h:241-248)."""
    from . import study as st
    pop, es, ee, coeff = county_graph(N, seed, max_out_degree, G)
    n = G * 10
    header = st.uv_array(flat.PSET_HEADER)
    st.set_value(header, flat.PH["StartDay"], 0.0)
    st.set_uniform(header, flat.PH["Seasonality"], 0.1, 0.3)
    st.set_value(header, flat.PH["TestAndTraceCapacityMaxRisk"], 5.0)
    age = st.uv_array((G, flat.OSECIR_AGE_STRIDE))
    for name, slot in flat.PA.items():
        age[:, slot]["value"] = flat.PA_DEFAULT[name]
    for name, (lo, hi) in COVID_UNIFORM.items():
        lo = np.broadcast_to(np.asarray(lo, dtype=np.float64)[:G] if np.ndim(lo) else lo, (G,))
        hi = np.broadcast_to(np.asarray(hi, dtype=np.float64)[:G] if np.ndim(hi) else hi, (G,))
        for i in range(G):
            st.set_uniform(age, (i, flat.PA[name]), lo[i], hi[i])
    rng = np.random.default_rng(seed + 1)
    y0 = np.zeros((N, G, 10))
    y0[:, :, flat.S] = pop
    seeded = rng.choice(N, size=min(n_seeded, N), replace=False)
    for k in seeded:
        e0 = 1e-4 * pop[k]
        y0[k, :, flat.E] = e0
        y0[k, :, flat.S] -= e0
    pops = st.uv_array((N, n))
    pops["value"] = y0.reshape(N, n)
    if pop_uncertainty > 0:
        for k in seeded:
            for i in range(G):
                v = y0[k, i, flat.E]
                st.set_uniform(pops, (k, i * 10 + flat.E), (1 - pop_uncertainty) * v, (1 + pop_uncertainty) * v)
    local = st.uv_array((N, 2))
    st.set_value(local, (slice(None), 0), flat.DBL_MAX)
    tnt = pop.sum(axis=1) * tnt_capacity_factor
    for k in range(N):
        st.set_uniform(local, (k, 1), 0.8 * tnt[k], 1.2 * tnt[k])
    dampings = []
    if damping_day is not None:
        v = st.uv_array(1)
        st.set_uniform(v, 0, 0.4, 0.8)
        dampings.append(st.Damping(value=v, time=damping_day, level=0, type=0, matrices=(0,)))
    return st.StudyGraph(num_groups=G, num_nodes=N, header=header, age_params=age, node_local=local,
                         populations=pops, baseline=np.full((1, G, G), cont_freq / G), minimum=None, edge_start=es,
                         edge_end=ee, edge_coeff=coeff, dampings=dampings)


def vvs_study_graph(N: int = 400, seed: int = 20261017, G: int = 6, max_out_degree: int = 25, days: int = 100,
                    sim_kind: int = flat.SIM_MODEL, rel: float = 0.1, damping_day: Optional[float] = 5.0,
                    daily_vaccinations: float = 5.0, init_per_10k: float = 20.0, icu_per_10k: float = 100.0 / 6):
    """Config 5 as a *parameter graph* for `osecirvvs::draw_sample(graph)`
    (cpp/models/ode_secirvvs/parameter_space.h:166-214): the county graph of county_problem_vvs with the reference
    benchmark's parameters (cpp/benchmarks/graph_simulation.cpp:31-112) as the centres of uniform distributions of
    relative half-width `rel` on every sampled parameter, Seasonality U(0.1, 0.3), +-10 % on the initially infected
    compartments and 0.8-1.2 on the test-and-trace capacity, and one sampled contact damping U(0.2, 0.4)."""
    from . import study as st
    base = county_problem_vvs(N=N, runs=1, seed=seed, G=G, max_out_degree=max_out_degree, days=days, sim_kind=sim_kind,
                              perturb=0.0, daily_vaccinations=daily_vaccinations, init_per_10k=init_per_10k,
                              icu_per_10k=icu_per_10k)
    C = flat.OSECIRVVS_NCOMP
    n = G * C
    pset = base.psets[0, 0]
    header = st.uv_array(flat.PSET_HEADER)
    header["value"] = pset[:flat.PSET_HEADER]
    st.set_uniform(header, flat.PHV["Seasonality"], 0.1, 0.3)
    age = st.uv_array((G, flat.OSECIRVVS_AGE_STRIDE))
    age["value"] = pset[flat.PSET_HEADER:].reshape(G, flat.OSECIRVVS_AGE_STRIDE)
    sampled = ["TimeExposed", "TimeInfectedNoSymptoms", "TimeInfectedSymptoms", "TimeInfectedSevere",
               "TimeInfectedCritical", "TransmissionProbabilityOnContact", "RelativeTransmissionNoSymptoms",
               "RecoveredPerInfectedNoSymptoms", "MaxRiskOfInfectionFromSymptomatic", "SeverePerInfectedSymptoms",
               "CriticalPerSevere", "DeathsPerCritical", "ReducExposedPartialImmunity", "ReducExposedImprovedImmunity",
               "ReducInfectedSymptomsPartialImmunity", "ReducInfectedSymptomsImprovedImmunity",
               "ReducInfectedSevereCriticalDeadPartialImmunity", "ReducInfectedSevereCriticalDeadImprovedImmunity",
               "ReducTimeInfectedMild"]
    for name in sampled:
        slot = flat.PV[name]
        for i in range(G):
            v = float(age[i, slot]["value"])
            if v > 0:
                st.set_uniform(age, (i, slot), (1 - rel) * v, (1 + rel) * v)
    pops = st.uv_array((N, n))
    pops["value"] = base.y0[0]
    y0 = base.y0[0].reshape(N, G, C)
    for k in range(N):
        for i in range(G):
            for c in VVS_INFECTED:
                v = y0[k, i, c]
                if v > 0:
                    st.set_uniform(pops, (k, i * C + c), 0.9 * v, 1.1 * v)
    local = st.uv_array((N, 2))
    local["value"] = base.node_local[0]
    for k in range(N):
        t = float(base.node_local[0, k, 1])
        st.set_uniform(local, (k, 1), 0.8 * t, 1.2 * t)
    baseline = np.full((1, G, G), 0.5)
    for i in range(G):
        baseline[0, i, i] = 5.0
    dampings = []
    if damping_day is not None:
        v = st.uv_array(1)
        st.set_uniform(v, 0, 0.2, 0.4)
        dampings.append(st.Damping(value=v, time=damping_day, level=0, type=0, matrices=(0,)))
    return st.StudyGraph(num_groups=G, num_nodes=N, header=header, age_params=age, node_local=local, populations=pops,
                         baseline=baseline, minimum=None, edge_start=base.edge_start, edge_end=base.edge_end,
                         edge_coeff=base.edge_coeff, dampings=dampings, model=flat.MODEL_OSECIRVVS, sim_kind=sim_kind,
                         vacc_partial=base.vacc_partial, vacc_full=base.vacc_full)

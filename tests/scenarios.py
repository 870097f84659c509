"""Scenario builders shared by the oracle tests (CPU) and the parity tests (GPU).

Each builder restates the *setup* of a reference test or example (cited), producing the flat descriptors of
include/mio_cuda.h. Nothing here reads /root/reference at run time.
"""
from __future__ import annotations

import os

import numpy as np

from memilio_b200 import cdefs, flat
from memilio_b200.flat import ContactLocation
from memilio_b200.problem import GraphProblem, single_node_problem

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_csv(name):
    """cpp/tests/load_test_data.h:38-70: space separated numbers, '#' lines are comments (also ' # t S E ...')."""
    rows = []
    with open(os.path.join(GOLDEN, name)) as f:
        for line in f:
            s = line.strip()
            if not s or s.startswith("#") or s[0].isalpha():  # comment or a header line of column names
                continue
            rows.append([float(w) for w in s.split()])
    return np.asarray(rows)


def secihurd_setup(G, crit_per_sev, constraints_before_contacts=True):
    """Population/parameter setup shared by cpp/tests/test_odesecir.cpp:39-90, test_odesecir_ageres.cpp:27-80,113-165
    and pycode/memilio-simulation/tests/test_osecir.py:38-88, test_osecir_ageres.py:36-100."""
    nb_total, nb_exp, nb_inf, nb_car, nb_hosp, nb_icu, nb_rec, nb_dead = 10000, 100, 50, 50, 20, 10, 10, 0
    fact = 1.0 / G
    y0 = np.zeros((G, 10))
    for i in range(G):
        if G == 1:
            vals = {flat.E: nb_exp, flat.INS: nb_car, flat.ISY: nb_inf, flat.ISEV: nb_hosp, flat.ICR: nb_icu,
                    flat.R: nb_rec, flat.D: nb_dead}
        else:
            vals = {flat.E: fact * nb_exp, flat.INS: fact * nb_car, flat.ISY: fact * nb_inf,
                    flat.ISEV: fact * nb_hosp, flat.ICR: fact * nb_icu, flat.R: fact * nb_rec,
                    flat.D: fact * nb_dead}
        for c, v in vals.items():
            y0[i, c] = v
        # set_difference_from_(group_)total, populations.h:194-205,240-249: accumulate in index order, subtract own
        cur = 0.0
        for c in range(10):
            cur = cur + y0[i, c]
        cur -= y0[i, flat.S]
        y0[i, flat.S] = (nb_total if G == 1 else fact * nb_total) - cur
    pset = flat.osecir_pset(G, StartDay=60, Seasonality=0.2, TimeExposed=3.2, TimeInfectedNoSymptoms=2.0,
                            TimeInfectedSymptoms=5.8, TimeInfectedSevere=9.5, TimeInfectedCritical=7.1,
                            TransmissionProbabilityOnContact=0.05, RelativeTransmissionNoSymptoms=0.7,
                            RecoveredPerInfectedNoSymptoms=0.09, RiskOfInfectionFromSymptomatic=0.25,
                            MaxRiskOfInfectionFromSymptomatic=0.45, SeverePerInfectedSymptoms=0.2,
                            CriticalPerSevere=crit_per_sev, DeathsPerCritical=0.3)
    local = np.array([flat.DBL_MAX, 35.0])  # ICUCapacity default, TestAndTraceCapacity = 35
    flat.osecir_apply_constraints(pset, G, local)
    cont_freq = 10
    base = np.full((G, G), fact * cont_freq if G > 1 else float(cont_freq))
    # contact_matrix[0].add_damping(0.7, SimulationTime(30.)) -> cache = [(0.7, t=30)]
    ct = flat.contact_table(G, [ContactLocation(base, None, [30.0], [np.full((G, G), 0.7)])], 1, 2)
    return dict(G=G, pset=pset, local=local, ct=ct, ct_max_loc=1, ct_max_pts=2, y0=y0.reshape(-1))


GOLDEN_CASES = {
    # name: (csv, G, CriticalPerSevere, integrator kwargs, dt0, tolerance, age-summed?)
    # TestOdeSecir.compareWithPreviousRun, cpp/tests/test_odesecir.cpp:39-112
    "cpp_rkf45_g1": ("secihurd-compare.csv", 1, 0.3,
                     dict(kind=flat.INTEGRATOR_RKF45, abs_tol=1e-1, rel_tol=1e-3, dt_min=0.3, dt_max=1.0), 0.3, 1e-10),
    # TestOdeSecir.compareAgeResWithPreviousRun, cpp/tests/test_odesecir_ageres.cpp:27-111
    "cpp_rkf45_g3": ("secihurd-compare.csv", 3, 0.3,
                     dict(kind=flat.INTEGRATOR_RKF45, abs_tol=1e-1, rel_tol=1e-4, dt_min=0.3, dt_max=1.0), 0.3, 1e-10),
    # TestOdeSecir.compareAgeResWithPreviousRunCashKarp, cpp/tests/test_odesecir_ageres.cpp:113-189
    "cpp_ck54_g3": ("secihurd-compare-cashkarp.csv", 3, 0.25,
                    dict(kind=flat.INTEGRATOR_CK54, abs_tol=1e-1, rel_tol=1e-4, dt_min=0.3, dt_max=1.0), 0.3, 1e-10),
    # test_osecir.py::test_compare_with_cpp, pycode/memilio-simulation/tests/test_osecir.py:122-150
    "py_default_g1": ("ode-secihurd-compare.csv", 1, 0.25, dict(kind=flat.INTEGRATOR_CK54), 0.1, 1e-9),
    # test_osecir_ageres.py::test_compare_with_cpp, pycode/memilio-simulation/tests/test_osecir_ageres.py:117-150
    "py_default_g3": ("ode-secihurd-ageres-compare.csv", 3, 0.25, dict(kind=flat.INTEGRATOR_CK54), 0.1, 1e-10),
}


def golden_case(name):
    csv, G, cps, ikw, dt0, tol = GOLDEN_CASES[name]
    s = secihurd_setup(G, cps)
    s.update(csv=load_csv(csv), integrator=cdefs.integrator_desc(**ikw), dt0=dt0, tol=tol, t0=0.0, tmax=50.0)
    return s


def golden_problem(name) -> GraphProblem:
    s = golden_case(name)
    return single_node_problem(s["G"], s["pset"], s["local"], s["ct"], s["y0"], ct_max_loc=s["ct_max_loc"],
                               ct_max_pts=s["ct_max_pts"], integrator=s["integrator"], t0=s["t0"], dt0=s["dt0"])


def vvs_golden_case():
    """pycode/memilio-simulation/tests/test_osecirvvs.py:38-136,170-196: osecirvvs, 1 age group, default integrator,
    tmax 50 -> ode-secirvvs-compare.csv."""
    G = 1
    y0 = np.zeros(flat.OSECIRVVS_NCOMP)
    y0[flat.V_EN], y0[flat.V_INSN], y0[flat.V_ISYN], y0[flat.V_ISEVN], y0[flat.V_ICRN] = 100, 50, 50, 20, 10
    y0[flat.V_SII] = 10
    cur = 0.0
    for c in range(flat.OSECIRVVS_NCOMP):  # set_difference_from_group_total, populations.h:194-205
        cur = cur + y0[c]
    cur -= y0[flat.V_SN]
    y0[flat.V_SN] = 10000 - cur
    pset = flat.osecirvvs_pset(G, Seasonality=0.2, TimeExposed=3.2, TimeInfectedNoSymptoms=2.0,
                               TimeInfectedSymptoms=5.0, TimeInfectedSevere=10.0, TimeInfectedCritical=8.0,
                               TransmissionProbabilityOnContact=0.05, RelativeTransmissionNoSymptoms=1,
                               RecoveredPerInfectedNoSymptoms=0.09, RiskOfInfectionFromSymptomatic=0.25,
                               SeverePerInfectedSymptoms=0.2, CriticalPerSevere=0.25, DeathsPerCritical=0.3,
                               ReducExposedPartialImmunity=1.0, ReducExposedImprovedImmunity=1.0,
                               ReducInfectedSymptomsPartialImmunity=1.0, ReducInfectedSymptomsImprovedImmunity=0,
                               ReducInfectedSevereCriticalDeadPartialImmunity=0,
                               ReducInfectedSevereCriticalDeadImprovedImmunity=0, ReducTimeInfectedMild=1)
    local = np.array([10000.0, 10000.0])
    flat.osecirvvs_apply_constraints(pset, G, local)
    ct = flat.contact_table(G, [ContactLocation(np.full((1, 1), 10.0), None, [30.0], [np.full((1, 1), 0.7)])], 1, 2)
    vacc = np.zeros((G, 1000))
    return dict(G=G, pset=pset, local=local, ct=ct, ct_max_loc=1, ct_max_pts=2, y0=y0, vacc_partial=vacc,
                vacc_full=vacc.copy(), csv=load_csv("ode-secirvvs-compare.csv"), t0=0.0, tmax=50.0, dt0=0.1, tol=1e-10)


def example_ode_secirvvs():
    """cpp/examples/ode_secirvvs.cpp:24-136: osecirvvs, 1 age group, tmax 30, cumulative vaccinations 10 per day."""
    G = 1
    y0 = np.zeros(flat.OSECIRVVS_NCOMP)
    init = dict(V_EN=10, V_EII=11, V_EPI=12, V_INSN=13, V_INSNC=13, V_INSPI=14, V_INSPIC=14, V_INSII=15, V_INSIIC=15,
                V_ISYN=5, V_ISYNC=5, V_ISYPI=6, V_ISYPIC=6, V_ISYII=7, V_ISYIIC=7, V_ISEVN=8, V_ISEVII=1, V_ISEVPI=2,
                V_ICRN=3, V_ICRPI=4, V_ICRII=5, V_SII=6, V_SPI=7)
    for k, v in init.items():
        y0[getattr(flat, k)] = v
    cur = 0.0
    for c in range(flat.OSECIRVVS_NCOMP):  # set_difference_from_group_total, populations.h:194-205
        cur = cur + y0[c]
    cur -= y0[flat.V_SN]
    y0[flat.V_SN] = 1000 - cur
    pset = flat.osecirvvs_pset(G, Seasonality=0.2, TimeExposed=3.33, TimeInfectedNoSymptoms=1.87,
                               TimeInfectedSymptoms=7, TimeInfectedSevere=6, TimeInfectedCritical=7,
                               TransmissionProbabilityOnContact=0.15, RelativeTransmissionNoSymptoms=0.5,
                               RiskOfInfectionFromSymptomatic=0.0, MaxRiskOfInfectionFromSymptomatic=0.4,
                               RecoveredPerInfectedNoSymptoms=0.2, SeverePerInfectedSymptoms=0.1, CriticalPerSevere=0.1,
                               DeathsPerCritical=0.1, ReducExposedPartialImmunity=0.8,
                               ReducExposedImprovedImmunity=0.331, ReducInfectedSymptomsPartialImmunity=0.65,
                               ReducInfectedSymptomsImprovedImmunity=0.243,
                               ReducInfectedSevereCriticalDeadPartialImmunity=0.1,
                               ReducInfectedSevereCriticalDeadImprovedImmunity=0.091, ReducTimeInfectedMild=0.9)
    local = np.array([100.0, 0.0143])
    flat.osecirvvs_apply_constraints(pset, G, local)
    ct = flat.contact_table(G, [ContactLocation(np.full((1, 1), 5.0), None, [5.0], [np.full((1, 1), 0.3)])], 1, 2)
    vacc = (10.0 * np.arange(31))[None, :].copy()
    return dict(G=G, pset=pset, local=local, ct=ct, ct_max_loc=1, ct_max_pts=2, y0=y0, vacc_partial=vacc,
                vacc_full=vacc.copy(), t0=0.0, tmax=30.0, dt0=0.1)


def age_sum(values, G):
    return values.reshape(values.shape[0], G, 10).sum(axis=1) if G > 1 else values


# ---------------------------------------------------------------------------------------------------------------
def example_ode_secir_graph(runs=1) -> GraphProblem:
    """cpp/examples/ode_secir_graph.cpp:29-112: G=1, two nodes, two edges with coefficient 0.1, dt 0.5, damping 0.7
    at t=15 in node 0 only, TestAndTraceCapacity 35, Seasonality 0.2."""
    G = 1
    pset = flat.osecir_pset(G, StartDay=0, Seasonality=0.2, TimeExposed=3.2, TimeInfectedNoSymptoms=2.0,
                            TimeInfectedSymptoms=5.8, TimeInfectedSevere=9.5, TimeInfectedCritical=7.1,
                            TransmissionProbabilityOnContact=0.1, RelativeTransmissionNoSymptoms=0.7,
                            RecoveredPerInfectedNoSymptoms=0.09, RiskOfInfectionFromSymptomatic=0.25,
                            MaxRiskOfInfectionFromSymptomatic=0.45, SeverePerInfectedSymptoms=0.2,
                            CriticalPerSevere=0.25, DeathsPerCritical=0.3)
    base = np.full((1, 1), 10.0)
    ct0 = flat.contact_table(G, [ContactLocation(base, None, [15.0], [np.full((1, 1), 0.7)])], 1, 2)
    ct1 = flat.contact_table(G, [ContactLocation(base)], 1, 2)
    y0 = np.zeros((2, 10))
    y0[0, flat.S], y0[0, flat.E] = 9990, 100
    y0[1, flat.S] = 10000
    local = np.array([[flat.DBL_MAX, 35.0], [flat.DBL_MAX, 35.0]])
    rep = lambda a: np.broadcast_to(a, (runs,) + a.shape).copy()
    return GraphProblem(model=flat.MODEL_OSECIR, num_groups=G, num_nodes=2, edge_start=[0, 1], edge_end=[1, 0],
                        edge_coeff=np.full((2, 10), 0.1), psets=rep(pset[None, :]),
                        ctables=rep(np.stack([ct0, ct1])), node_local=rep(local), y0=rep(y0), ct_max_loc=1,
                        ct_max_pts=2, node_ctable=np.array([0, 1], dtype=np.int32), dt0=np.full((runs, 2), 0.1),
                        t0=0.0, dt_graph=0.5)


# uniform ranges of set_covid_parameters, cpp/examples/ode_secir_parameter_study_graph.cpp:91-159 (min, max per age)
def _rng(mn, mx):
    return list(zip(mn, mx))


COVID_RANGES = dict(
    TimeExposed=(2.67, 4.0), TimeInfectedNoSymptoms=(1.2, 2.53),
    TimeInfectedSymptoms=_rng([5.6255, 5.6255, 5.6646, 5.5631, 5.501, 5.465],
                              [8.427, 8.427, 8.4684, 8.3139, 8.169, 8.085]),
    TimeInfectedSevere=_rng([3.925, 3.925, 4.85, 6.4, 7.2, 9.], [6.075, 6.075, 7., 8.7, 9.8, 13.]),
    TimeInfectedCritical=_rng([4.95, 4.95, 4.86, 14.14, 14.4, 10.], [8.95, 8.95, 8.86, 20.58, 19.8, 13.2]),
    TransmissionProbabilityOnContact=_rng([0.02, 0.05, 0.05, 0.05, 0.08, 0.15], [0.04, 0.07, 0.07, 0.07, 0.10, 0.20]),
    RelativeTransmissionNoSymptoms=(1.0, 1.0),
    RiskOfInfectionFromSymptomatic=(0.1, 0.3), MaxRiskOfInfectionFromSymptomatic=(0.3, 0.5),
    RecoveredPerInfectedNoSymptoms=_rng([0.2, 0.2, 0.15, 0.15, 0.15, 0.15], [0.3, 0.3, 0.25, 0.25, 0.25, 0.25]),
    SeverePerInfectedSymptoms=_rng([0.006, 0.006, 0.015, 0.049, 0.15, 0.20], [0.009, 0.009, 0.023, 0.074, 0.18, 0.25]),
    CriticalPerSevere=_rng([0.05, 0.05, 0.05, 0.10, 0.25, 0.35], [0.10, 0.10, 0.10, 0.20, 0.35, 0.45]),
    DeathsPerCritical=_rng([0.00, 0.00, 0.10, 0.10, 0.30, 0.5], [0.10, 0.10, 0.18, 0.18, 0.50, 0.7]),
)
SEASONALITY_RANGE = (0.1, 0.3)


def covid_range(name, G=6):
    """(min[G], max[G]) arrays for one parameter tag."""
    r = COVID_RANGES[name]
    if isinstance(r, tuple):
        return np.full(G, r[0]), np.full(G, r[1])
    a = np.asarray(r)
    return a[:, 0].copy(), a[:, 1].copy()


def damping_flow_case(damping=None):
    """TestOdeSecir.testDamping (cpp/tests/test_odesecir.cpp:604-676) and cpp/examples/ode_secir_contact_changes.cpp:
    one age group, 10 of 1000 InfectedSymptoms, contact frequency 10, transmission probability and risk of infection
    from symptomatic 1, every other parameter at its default; optionally one damping of the contact matrix at t = 0."""
    G = 1
    y0 = np.zeros(10)
    y0[flat.ISY] = 10.0
    y0[flat.S] = 1000.0 - 10.0
    pset = flat.osecir_pset(G, TransmissionProbabilityOnContact=1.0, RiskOfInfectionFromSymptomatic=1.0)
    local = np.array([flat.DBL_MAX, flat.DBL_MAX])
    base = np.full((1, 1), 10.0)
    times, mats = ([], []) if damping is None else ([0.0], [np.full((1, 1), float(damping))])
    ct = flat.contact_table(G, [ContactLocation(base, None, times, mats)], 1, 2)
    return dict(G=G, pset=pset, local=local, ct=ct, ct_max_loc=1, ct_max_pts=2, y0=y0)

"""Flat descriptor builders for the C-ABI of include/mio_cuda.h (numpy, host side).

The layouts mirror the `MIO_P*` constants of the header one to one:
parameter sets follow the tags of cpp/models/ode_secir/parameters.h:47-397, contact tables are the evaluated
damping cache of a ContactMatrixGroup (cpp/memilio/epidemiology/contact_matrix.h:271-400, damping.h:571-600).
"""
from __future__ import annotations

import dataclasses
from typing import List, Optional, Sequence

import numpy as np

MODEL_OSECIR = 0
MODEL_OSECIRVVS = 1
INTEGRATOR_CK54 = 0
INTEGRATOR_RKF45 = 1
INTEGRATOR_EULER = 2
SIM_BASE = 0
SIM_MODEL = 1

PSET_HEADER = 8
OSECIR_AGE_STRIDE = 16
OSECIR_NCOMP = 10
OSECIRVVS_NCOMP = 27
NODE_LOCAL = 2
NUM_COUNTERS = 4

DBL_MAX = float(np.finfo(np.float64).max)
DBL_MIN = float(np.finfo(np.float64).tiny)

# header slots
PH = dict(StartDay=0, Seasonality=1, TestAndTraceCapacityMaxRisk=2, commuter_nondetection=3,
          start_commuter_detection=4, end_commuter_detection=5)
# osecir age record slots, order of the tags in ode_secir/parameters.h
PA = dict(TimeExposed=0, TimeInfectedNoSymptoms=1, TimeInfectedSymptoms=2, TimeInfectedSevere=3,
          TimeInfectedCritical=4, TransmissionProbabilityOnContact=5, RelativeTransmissionNoSymptoms=6,
          RecoveredPerInfectedNoSymptoms=7, RiskOfInfectionFromSymptomatic=8,
          MaxRiskOfInfectionFromSymptomatic=9, SeverePerInfectedSymptoms=10, CriticalPerSevere=11,
          DeathsPerSevere=12, DeathsPerCritical=13)
# defaults, ode_secir/parameters.h get_default()
PA_DEFAULT = dict(TimeExposed=1., TimeInfectedNoSymptoms=1., TimeInfectedSymptoms=1., TimeInfectedSevere=1.,
                  TimeInfectedCritical=1., TransmissionProbabilityOnContact=1., RelativeTransmissionNoSymptoms=1.,
                  RecoveredPerInfectedNoSymptoms=0., RiskOfInfectionFromSymptomatic=1.,
                  MaxRiskOfInfectionFromSymptomatic=1., SeverePerInfectedSymptoms=0., CriticalPerSevere=0.,
                  DeathsPerSevere=0., DeathsPerCritical=0.)

# InfectionState, ode_secir/infection_state.h:33-47
S, E, INS, INSC, ISY, ISYC, ISEV, ICR, R, D = range(10)


def num_compartments(model: int) -> int:
    return OSECIR_NCOMP if model == MODEL_OSECIR else OSECIRVVS_NCOMP


def pset_size(model: int, G: int) -> int:
    assert model == MODEL_OSECIR
    return PSET_HEADER + OSECIR_AGE_STRIDE * G


def ctable_size(G: int, max_loc: int, max_pts: int) -> int:
    return 1 + max_loc + max_loc * (max_pts + (2 + max_pts) * G * G)


def mobtable_size(n: int, mob_max_pts: int) -> int:
    return 1 + mob_max_pts + mob_max_pts * n


def osecir_pset(G: int, **kw) -> np.ndarray:
    """Build one osecir parameter set. Scalars by header name, per-age values as scalar or sequence of length G."""
    p = np.zeros(pset_size(MODEL_OSECIR, G))
    p[PH["TestAndTraceCapacityMaxRisk"]] = 5.0
    rec = p[PSET_HEADER:].reshape(G, OSECIR_AGE_STRIDE)
    for name, slot in PA.items():
        rec[:, slot] = PA_DEFAULT[name]
    for name, val in kw.items():
        if name in PH:
            p[PH[name]] = float(val)
        elif name in PA:
            rec[:, PA[name]] = np.broadcast_to(np.asarray(val, dtype=np.float64), (G,))
        else:
            raise KeyError(name)
    return p


def osecir_apply_constraints(p: np.ndarray, G: int, local: Optional[np.ndarray] = None) -> bool:
    """Parameters::apply_constraints, ode_secir/parameters.h:465-615, on a flat parameter set (in place)."""
    corrected = False
    if p[PH["Seasonality"]] < 0.0 or p[PH["Seasonality"]] > 0.5:
        p[PH["Seasonality"]] = 0.0
        corrected = True
    if local is not None:
        for k in range(2):
            if local[k] < 0.0:
                local[k] = 0.0
                corrected = True
    if p[PH["TestAndTraceCapacityMaxRisk"]] < 0.0:
        p[PH["TestAndTraceCapacityMaxRisk"]] = 0.0
        corrected = True
    rec = p[PSET_HEADER:].reshape(G, OSECIR_AGE_STRIDE)
    for i in range(G):
        for name in ("TimeExposed", "TimeInfectedNoSymptoms", "TimeInfectedSymptoms", "TimeInfectedSevere",
                     "TimeInfectedCritical"):
            if rec[i, PA[name]] < 0.1:
                rec[i, PA[name]] = 0.1
                corrected = True
        s = PA["TransmissionProbabilityOnContact"]
        if rec[i, s] < 0.0 or rec[i, s] > 1.0:
            rec[i, s] = 0.0
            corrected = True
        s = PA["RelativeTransmissionNoSymptoms"]
        if rec[i, s] < 0.0:
            rec[i, s] = 0.0
            corrected = True
        for name in ("RecoveredPerInfectedNoSymptoms", "RiskOfInfectionFromSymptomatic",
                     "SeverePerInfectedSymptoms", "CriticalPerSevere", "DeathsPerSevere"):
            s = PA[name]
            if rec[i, s] < 0.0 or rec[i, s] > 1.0:
                rec[i, s] = 0.0
                corrected = True
        if rec[i, PA["CriticalPerSevere"]] + rec[i, PA["DeathsPerSevere"]] > 1.0:
            rec[i, PA["DeathsPerSevere"]] = 0.0
            corrected = True
        s = PA["DeathsPerCritical"]
        if rec[i, s] < 0.0 or rec[i, s] > 1.0:
            rec[i, s] = 0.0
            corrected = True
    return corrected


@dataclasses.dataclass
class ContactLocation:
    """One ContactMatrix: baseline, minimum and the evaluated damping cache (change times, accumulated dampings)."""
    baseline: np.ndarray
    minimum: Optional[np.ndarray] = None
    times: Sequence[float] = ()
    dampings: Sequence[np.ndarray] = ()


def contact_table(G: int, locations: List[ContactLocation], max_loc: int, max_pts: int) -> np.ndarray:
    ct = np.zeros(ctable_size(G, max_loc, max_pts))
    assert 1 <= len(locations) <= max_loc
    ct[0] = len(locations)
    stride = max_pts + (2 + max_pts) * G * G
    for l, loc in enumerate(locations):
        K = len(loc.times)
        assert K <= max_pts and K == len(loc.dampings)
        ct[1 + l] = K
        base = 1 + max_loc + l * stride
        ct[base:base + K] = np.asarray(loc.times, dtype=np.float64)
        o = base + max_pts
        ct[o:o + G * G] = np.asarray(loc.baseline, dtype=np.float64).reshape(-1)
        o += G * G
        if loc.minimum is not None:
            ct[o:o + G * G] = np.asarray(loc.minimum, dtype=np.float64).reshape(-1)
        o += G * G
        for k in range(K):
            ct[o + k * G * G:o + (k + 1) * G * G] = np.broadcast_to(
                np.asarray(loc.dampings[k], dtype=np.float64), (G, G)).reshape(-1)
    return ct


def mobility_table(n: int, times: Sequence[float], dampings: Sequence[np.ndarray], mob_max_pts: int) -> np.ndarray:
    mt = np.zeros(mobtable_size(n, mob_max_pts))
    K = len(times)
    assert K <= mob_max_pts
    mt[0] = K
    mt[1:1 + K] = np.asarray(times, dtype=np.float64)
    for k in range(K):
        o = 1 + mob_max_pts + k * n
        mt[o:o + n] = np.broadcast_to(np.asarray(dampings[k], dtype=np.float64), (n,))
    return mt

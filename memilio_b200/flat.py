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

# osecirvvs (cpp/models/ode_secirvvs/parameters.h): header slot 2 = TestAndTraceCapacityMaxRiskSymptoms
OSECIRVVS_AGE_STRIDE = 32
PHV = dict(PH, TestAndTraceCapacityMaxRiskSymptoms=2, TestAndTraceCapacityMaxRiskNoSymptoms=6, StartDayNewVariant=7)
del PHV["TestAndTraceCapacityMaxRisk"]
PV = dict(PA, ReducExposedPartialImmunity=14, ReducExposedImprovedImmunity=15, ReducInfectedSymptomsPartialImmunity=16,
          ReducInfectedSymptomsImprovedImmunity=17, ReducInfectedSevereCriticalDeadPartialImmunity=18,
          ReducInfectedSevereCriticalDeadImprovedImmunity=19, ReducTimeInfectedMild=20, InfectiousnessNewVariant=21,
          VaccinationGap=22, DaysUntilEffectivePartialImmunity=23, DaysUntilEffectiveImprovedImmunity=24)
PV_DEFAULT = dict(PA_DEFAULT, ReducExposedPartialImmunity=1., ReducExposedImprovedImmunity=1.,
                  ReducInfectedSymptomsPartialImmunity=1., ReducInfectedSymptomsImprovedImmunity=1.,
                  ReducInfectedSevereCriticalDeadPartialImmunity=1.,
                  ReducInfectedSevereCriticalDeadImprovedImmunity=1., ReducTimeInfectedMild=0.5,
                  InfectiousnessNewVariant=1., VaccinationGap=49., DaysUntilEffectivePartialImmunity=14.,
                  DaysUntilEffectiveImprovedImmunity=7.)
# osecirvvs InfectionState, ode_secirvvs/infection_state.h:36-67
(V_SN, V_SPI, V_EN, V_EPI, V_EII, V_INSN, V_INSPI, V_INSII, V_INSNC, V_INSPIC, V_INSIIC, V_ISYN, V_ISYPI, V_ISYII,
 V_ISYNC, V_ISYPIC, V_ISYIIC, V_ISEVN, V_ISEVPI, V_ISEVII, V_ICRN, V_ICRPI, V_ICRII, V_SII, V_DN, V_DPI,
 V_DII) = range(27)

# InfectionState, ode_secir/infection_state.h:33-47
S, E, INS, INSC, ISY, ISYC, ISEV, ICR, R, D = range(10)


def num_compartments(model: int) -> int:
    return OSECIR_NCOMP if model == MODEL_OSECIR else OSECIRVVS_NCOMP


def age_stride(model: int) -> int:
    return OSECIR_AGE_STRIDE if model == MODEL_OSECIR else OSECIRVVS_AGE_STRIDE


def pset_size(model: int, G: int) -> int:
    return PSET_HEADER + age_stride(model) * G


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


def osecirvvs_pset(G: int, **kw) -> np.ndarray:
    """One osecirvvs parameter set (defaults of ode_secirvvs/parameters.h get_default())."""
    p = np.zeros(pset_size(MODEL_OSECIRVVS, G))
    p[PHV["TestAndTraceCapacityMaxRiskSymptoms"]] = 15.0
    p[PHV["TestAndTraceCapacityMaxRiskNoSymptoms"]] = 2.0
    p[PHV["StartDayNewVariant"]] = DBL_MAX
    rec = p[PSET_HEADER:].reshape(G, OSECIRVVS_AGE_STRIDE)
    for name, slot in PV.items():
        rec[:, slot] = PV_DEFAULT[name]
    for name, val in kw.items():
        if name in PHV:
            p[PHV[name]] = float(val)
        elif name in PV:
            rec[:, PV[name]] = np.broadcast_to(np.asarray(val, dtype=np.float64), (G,))
        else:
            raise KeyError(name)
    return p


def osecirvvs_apply_constraints(p: np.ndarray, G: int, local: Optional[np.ndarray] = None) -> bool:
    """osecirvvs Parameters::apply_constraints, ode_secirvvs/parameters.h:705-946, on a flat parameter set."""
    corrected = False

    def fix(arr, idx, bad, to):
        nonlocal corrected
        if bad:
            arr[idx] = to
            corrected = True
    fix(p, PHV["Seasonality"], p[PHV["Seasonality"]] < 0.0 or p[PHV["Seasonality"]] > 0.5, 0.0)
    if local is not None:
        for k in range(2):
            fix(local, k, local[k] < 0.0, 0.0)
    for name in ("TestAndTraceCapacityMaxRiskSymptoms", "TestAndTraceCapacityMaxRiskNoSymptoms"):
        fix(p, PHV[name], p[PHV[name]] < 0.0, 0.0)
    rec = p[PSET_HEADER:].reshape(G, OSECIRVVS_AGE_STRIDE)
    for i in range(G):
        r = rec[i]
        for name in ("TimeExposed", "TimeInfectedNoSymptoms", "TimeInfectedSymptoms", "TimeInfectedSevere",
                     "TimeInfectedCritical"):
            fix(r, PV[name], r[PV[name]] < 0.1, 0.1)
        s = PV["TransmissionProbabilityOnContact"]
        fix(r, s, r[s] < 0.0 or r[s] > 1.0, 0.0)
        s = PV["RelativeTransmissionNoSymptoms"]
        fix(r, s, r[s] < 0.0, 0.0)
        for name in ("RecoveredPerInfectedNoSymptoms", "RiskOfInfectionFromSymptomatic", "SeverePerInfectedSymptoms",
                     "CriticalPerSevere", "DeathsPerSevere"):
            s = PV[name]
            fix(r, s, r[s] < 0.0 or r[s] > 1.0, 0.0)
        fix(r, PV["DeathsPerSevere"], r[PV["CriticalPerSevere"]] + r[PV["DeathsPerSevere"]] > 1.0, 0.0)
        s = PV["DeathsPerCritical"]
        fix(r, s, r[s] < 0.0 or r[s] > 1.0, 0.0)
        for name in ("DaysUntilEffectivePartialImmunity", "DaysUntilEffectiveImprovedImmunity", "VaccinationGap"):
            fix(r, PV[name], r[PV[name]] < 0.0, 0.0)
        for name in ("ReducExposedPartialImmunity", "ReducExposedImprovedImmunity",
                     "ReducInfectedSymptomsPartialImmunity", "ReducInfectedSymptomsImprovedImmunity",
                     "ReducInfectedSevereCriticalDeadPartialImmunity",
                     "ReducInfectedSevereCriticalDeadImprovedImmunity", "ReducTimeInfectedMild"):
            fix(r, PV[name], r[PV[name]] <= 0.0 or r[PV[name]] > 1.0, 1.0)
        fix(r, PV["InfectiousnessNewVariant"], r[PV["InfectiousnessNewVariant"]] < 0.0, 1.0)
    return corrected


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

// mio_oracle_vvs.hpp -- CPU ORACLE (test infrastructure, NOT product code): ODE-SECIRVVS.
//
// Literal restatement of cpp/models/ode_secirvvs/model.h:40-88 (flows), :111-520 (get_flows), :599-771
// (Simulation::apply_variant / apply_vaccination / advance without dynamic NPIs), :846-940 (get_mobility_factors,
// test_commuters), cpp/models/ode_secirvvs/parameters.h:43-625 (defaults), :705-946 (apply_constraints).
// PINNED by the reference's golden vector pycode/memilio-simulation/tests/data/ode-secirvvs-compare.csv
// (test_osecirvvs.py:170-196, G = 1, default integrator, day-chunk advance; 1e-10).
// Graph-level SECIRVVS trajectories: PARITY UNPINNED by the reference (no golden data).
#pragma once

#include "mio_oracle.hpp"

namespace orc
{
namespace osecirvvs
{

enum : int
{
    SN = 0,
    SPI,
    EN,
    EPI,
    EII,
    INSN,
    INSPI,
    INSII,
    INSNC,
    INSPIC,
    INSIIC,
    ISyN,
    ISyPI,
    ISyII,
    ISyNC,
    ISyPIC,
    ISyIIC,
    ISevN,
    ISevPI,
    ISevII,
    ICrN,
    ICrPI,
    ICrII,
    SII,
    DN,
    DPI,
    DII,
    NC
};
constexpr int NF = 45;
// flow list of model.h:40-88 in TypeList order
static const int FLOW_SRC[NF] = {
    SN,  EN,  INSN,  INSN,  INSNC,  INSNC,  ISyN,  ISyN,  ISyNC,  ISyNC,  ISevN,  ISevN,  ISevN,  ICrN,  ICrN,
    SPI, EPI, INSPI, INSPI, INSPIC, INSPIC, ISyPI, ISyPI, ISyPIC, ISyPIC, ISevPI, ISevPI, ISevPI, ICrPI, ICrPI,
    SII, EII, INSII, INSII, INSIIC, INSIIC, ISyII, ISyII, ISyIIC, ISyIIC, ISevII, ISevII, ISevII, ICrII, ICrII};
static const int FLOW_DST[NF] = {
    EN,  INSN,  ISyN,  SII, ISyNC,  SII, ISevN,  SII, ISevN,  SII, ICrN,  SII, DN,  DN,  SII,
    EPI, INSPI, ISyPI, SII, ISyPIC, SII, ISevPI, SII, ISevPI, SII, ICrPI, SII, DPI, DPI, SII,
    EII, INSII, ISyII, SII, ISyIIC, SII, ISevII, SII, ISevII, SII, ICrII, SII, DII, DII, SII};
// index of a flow inside one immunity layer's block of 15
enum : int
{
    fS_E = 0,
    fE_INS,
    fINS_ISy,
    fINS_R,
    fINSC_ISyC,
    fINSC_R,
    fISy_ISev,
    fISy_R,
    fISyC_ISev,
    fISyC_R,
    fISev_ICr,
    fISev_R,
    fISev_D,
    fICr_D,
    fICr_R
};

struct Params {
    int G                                        = 1;
    double StartDay                              = 0.0;
    double StartDayNewVariant                    = std::numeric_limits<double>::max();
    double Seasonality                           = 0.0;
    double ICUCapacity                           = std::numeric_limits<double>::max();
    double TestAndTraceCapacity                  = std::numeric_limits<double>::max();
    double TestAndTraceCapacityMaxRiskNoSymptoms = 2.0;
    double TestAndTraceCapacityMaxRiskSymptoms   = 15.0;
    DampingExprGroup ContactPatterns;
    Vec TimeExposed, TimeInfectedNoSymptoms, TimeInfectedSymptoms, TimeInfectedSevere, TimeInfectedCritical,
        TransmissionProbabilityOnContact, RelativeTransmissionNoSymptoms, RecoveredPerInfectedNoSymptoms,
        RiskOfInfectionFromSymptomatic, MaxRiskOfInfectionFromSymptomatic, SeverePerInfectedSymptoms,
        CriticalPerSevere, DeathsPerSevere, DeathsPerCritical, VaccinationGap, DaysUntilEffectivePartialImmunity,
        DaysUntilEffectiveImprovedImmunity, ReducExposedPartialImmunity, ReducExposedImprovedImmunity,
        ReducInfectedSymptomsPartialImmunity, ReducInfectedSymptomsImprovedImmunity,
        ReducInfectedSevereCriticalDeadPartialImmunity, ReducInfectedSevereCriticalDeadImprovedImmunity,
        ReducTimeInfectedMild, InfectiousnessNewVariant;
    // [age][day], cumulative (parameters.h:470-497)
    std::vector<Vec> DailyPartialVaccinations, DailyFullVaccinations;
    double commuter_nondetection = 0.0, start_commuter_detection = 0.0, end_commuter_detection = 0.0;

    explicit Params(int g = 1)
        : G(g)
    {
        auto a = [g](double v) {
            return Vec(size_t(g), v);
        };
        TimeExposed = TimeInfectedNoSymptoms = TimeInfectedSymptoms = TimeInfectedSevere = TimeInfectedCritical = a(1.);
        TransmissionProbabilityOnContact = RelativeTransmissionNoSymptoms = a(1.);
        RecoveredPerInfectedNoSymptoms                                   = a(0.);
        RiskOfInfectionFromSymptomatic = MaxRiskOfInfectionFromSymptomatic = a(1.);
        SeverePerInfectedSymptoms = CriticalPerSevere = DeathsPerSevere = DeathsPerCritical = a(0.);
        VaccinationGap                     = a(49.0);
        DaysUntilEffectivePartialImmunity  = a(14.0);
        DaysUntilEffectiveImprovedImmunity = a(7.0);
        ReducExposedPartialImmunity = ReducExposedImprovedImmunity = ReducInfectedSymptomsPartialImmunity =
            ReducInfectedSymptomsImprovedImmunity = ReducInfectedSevereCriticalDeadPartialImmunity =
                ReducInfectedSevereCriticalDeadImprovedImmunity = a(1.);
        ReducTimeInfectedMild    = a(0.5);
        InfectiousnessNewVariant = a(1.0);
        DailyPartialVaccinations.assign(size_t(g), Vec());
        DailyFullVaccinations.assign(size_t(g), Vec());
        ContactPatterns.m.emplace_back(Mat(g, g, 0.0));
    }

    // parameters.h:705-946
    bool apply_constraints()
    {
        bool corrected = false;
        auto fix       = [&](double& v, bool bad, double to) {
            if (bad) {
                v         = to;
                corrected = true;
            }
        };
        fix(Seasonality, Seasonality < 0.0 || Seasonality > 0.5, 0);
        fix(ICUCapacity, ICUCapacity < 0.0, 0);
        fix(TestAndTraceCapacity, TestAndTraceCapacity < 0.0, 0);
        fix(TestAndTraceCapacityMaxRiskSymptoms, TestAndTraceCapacityMaxRiskSymptoms < 0.0, 0);
        fix(TestAndTraceCapacityMaxRiskNoSymptoms, TestAndTraceCapacityMaxRiskNoSymptoms < 0.0, 0);
        const double tol_times = 1e-1;
        for (int i = 0; i < G; ++i) {
            fix(TimeExposed[i], TimeExposed[i] < tol_times, tol_times);
            fix(TimeInfectedNoSymptoms[i], TimeInfectedNoSymptoms[i] < tol_times, tol_times);
            fix(TimeInfectedSymptoms[i], TimeInfectedSymptoms[i] < tol_times, tol_times);
            fix(TimeInfectedSevere[i], TimeInfectedSevere[i] < tol_times, tol_times);
            fix(TimeInfectedCritical[i], TimeInfectedCritical[i] < tol_times, tol_times);
            fix(TransmissionProbabilityOnContact[i],
                TransmissionProbabilityOnContact[i] < 0.0 || TransmissionProbabilityOnContact[i] > 1.0, 0.0);
            fix(RelativeTransmissionNoSymptoms[i], RelativeTransmissionNoSymptoms[i] < 0.0, 0);
            fix(RecoveredPerInfectedNoSymptoms[i],
                RecoveredPerInfectedNoSymptoms[i] < 0.0 || RecoveredPerInfectedNoSymptoms[i] > 1.0, 0);
            fix(RiskOfInfectionFromSymptomatic[i],
                RiskOfInfectionFromSymptomatic[i] < 0.0 || RiskOfInfectionFromSymptomatic[i] > 1.0, 0);
            fix(SeverePerInfectedSymptoms[i], SeverePerInfectedSymptoms[i] < 0.0 || SeverePerInfectedSymptoms[i] > 1.0,
                0);
            fix(CriticalPerSevere[i], CriticalPerSevere[i] < 0.0 || CriticalPerSevere[i] > 1.0, 0);
            fix(DeathsPerSevere[i], DeathsPerSevere[i] < 0.0 || DeathsPerSevere[i] > 1.0, 0);
            fix(DeathsPerSevere[i], CriticalPerSevere[i] + DeathsPerSevere[i] > 1.0, 0);
            fix(DeathsPerCritical[i], DeathsPerCritical[i] < 0.0 || DeathsPerCritical[i] > 1.0, 0);
            fix(DaysUntilEffectivePartialImmunity[i], DaysUntilEffectivePartialImmunity[i] < 0.0, 0);
            fix(DaysUntilEffectiveImprovedImmunity[i], DaysUntilEffectiveImprovedImmunity[i] < 0.0, 0);
            for (Vec* r : {&ReducExposedPartialImmunity, &ReducExposedImprovedImmunity,
                           &ReducInfectedSymptomsPartialImmunity, &ReducInfectedSymptomsImprovedImmunity,
                           &ReducInfectedSevereCriticalDeadPartialImmunity,
                           &ReducInfectedSevereCriticalDeadImprovedImmunity, &ReducTimeInfectedMild}) {
                fix((*r)[i], (*r)[i] <= 0.0 || (*r)[i] > 1.0, 1.0);
            }
            fix(InfectiousnessNewVariant[i], InfectiousnessNewVariant[i] < 0.0, 1.0);
            fix(VaccinationGap[i], VaccinationGap[i] < 0.0, 0);
        }
        return corrected;
    }
};

struct Model {
    Params parameters;
    Vec populations; // flat, idx = age*NC + compartment
    // false: the node type is the plain mio::Simulation<Model> (as in cpp/benchmarks/graph_simulation.cpp:121-126):
    // the ADL hooks fall back to unit mobility factors and no commuter testing (instant.h:513-518,537-541) and
    // advance() neither vaccinates nor applies the variant
    bool model_sim = true;
    explicit Model(int g = 1)
        : parameters(g)
        , populations(size_t(g) * NC, 0.0)
    {
    }
    int num_compartments() const
    {
        return NC;
    }

    // model.h:111-520
    void get_flows(const Vec& pop, const Vec& y, double t, Vec& flows) const
    {
        const Params& params = parameters;
        const int n_agegroups = params.G;
        double icu_occupancy           = 0.0;
        double test_and_trace_required = 0.0;
        for (int i = 0; i < n_agegroups; ++i) {
            const int o = i * NC;
            test_and_trace_required += (1 - params.RecoveredPerInfectedNoSymptoms[i]) /
                                       params.TimeInfectedNoSymptoms[i] *
                                       (pop[o + INSN] + pop[o + INSPI] + pop[o + INSII] + pop[o + INSNC] +
                                        pop[o + INSPIC] + pop[o + INSIIC]);
            icu_occupancy += pop[o + ICrN] + pop[o + ICrPI] + pop[o + ICrII];
        }
        for (int i = 0; i < n_agegroups; i++) {
            const int o = i * NC;
            const int SNi = o + SN, ENi = o + EN, INSNi = o + INSN, ISyNi = o + ISyN, ISevNi = o + ISevN,
                      ICrNi = o + ICrN, INSNCi = o + INSNC, ISyNCi = o + ISyNC, SPIi = o + SPI, EPIi = o + EPI,
                      INSPIi = o + INSPI, ISyPIi = o + ISyPI, ISevPIi = o + ISevPI, ICrPIi = o + ICrPI,
                      INSPICi = o + INSPIC, ISyPICi = o + ISyPIC, EIIi = o + EII, INSIIi = o + INSII,
                      ISyIIi = o + ISyII, ISevIIi = o + ISevII, ICrIIi = o + ICrII, INSIICi = o + INSIIC,
                      ISyIICi = o + ISyIIC, SIIi = o + SII;
            double* fN  = flows.data() + size_t(i) * NF;
            double* fPI = fN + 15;
            double* fII = fN + 30;

            double reducExposedPartialImmunity           = params.ReducExposedPartialImmunity[i];
            double reducExposedImprovedImmunity          = params.ReducExposedImprovedImmunity[i];
            double reducInfectedSymptomsPartialImmunity  = params.ReducInfectedSymptomsPartialImmunity[i];
            double reducInfectedSymptomsImprovedImmunity = params.ReducInfectedSymptomsImprovedImmunity[i];
            double reducInfectedSevereCriticalDeadPartialImmunity =
                params.ReducInfectedSevereCriticalDeadPartialImmunity[i];
            double reducInfectedSevereCriticalDeadImprovedImmunity =
                params.ReducInfectedSevereCriticalDeadImprovedImmunity[i];
            double reducTimeInfectedMild = params.ReducTimeInfectedMild[i];

            auto riskFromInfectedSymptomatic = smoother_cosine(
                test_and_trace_required, params.TestAndTraceCapacity,
                params.TestAndTraceCapacity * params.TestAndTraceCapacityMaxRiskSymptoms,
                params.RiskOfInfectionFromSymptomatic[i], params.MaxRiskOfInfectionFromSymptomatic[i]);
            auto riskFromInfectedNoSymptoms =
                smoother_cosine(test_and_trace_required, params.TestAndTraceCapacity,
                                params.TestAndTraceCapacity * params.TestAndTraceCapacityMaxRiskNoSymptoms,
                                params.RelativeTransmissionNoSymptoms[i], 1.0);

            for (int j = 0; j < n_agegroups; j++) {
                const int q = j * NC;
                double season_val =
                    (1 + params.Seasonality *
                             std::sin(std::numbers::pi_v<double> * ((params.StartDay + t) / 182.5 + 0.5)));
                double cont_freq_eff = season_val * params.ContactPatterns.at(t, i, j);
                double Nj = pop[q + SN] + pop[q + EN] + pop[q + INSN] + pop[q + ISyN] + pop[q + ISevN] +
                            pop[q + ICrN] + pop[q + INSNC] + pop[q + ISyNC] + pop[q + SPI] + pop[q + EPI] +
                            pop[q + INSPI] + pop[q + ISyPI] + pop[q + ISevPI] + pop[q + ICrPI] + pop[q + INSPIC] +
                            pop[q + ISyPIC] + pop[q + SII] + pop[q + EII] + pop[q + INSII] + pop[q + ISyII] +
                            pop[q + ISevII] + pop[q + ICrII] + pop[q + INSIIC] + pop[q + ISyIIC];
                const double divNj = (Nj < 1e-12) ? 0.0 : 1.0 / Nj;
                double ext_inf_force_dummy =
                    cont_freq_eff * divNj * params.TransmissionProbabilityOnContact[i] *
                    (riskFromInfectedNoSymptoms * (pop[q + INSN] + pop[q + INSPI] + pop[q + INSII]) +
                     riskFromInfectedSymptomatic * (pop[q + ISyN] + pop[q + ISyPI] + pop[q + ISyII]));
                double dummy_SN  = y[SNi] * ext_inf_force_dummy;
                double dummy_SPI = y[SPIi] * reducExposedPartialImmunity * ext_inf_force_dummy;
                double dummy_SII = y[SIIi] * reducExposedImprovedImmunity * ext_inf_force_dummy;
                fN[fS_E] += dummy_SN;
                fPI[fS_E] += dummy_SPI;
                fII[fS_E] += dummy_SII;
            }

            double criticalPerSevereAdjusted = smoother_cosine(icu_occupancy, 0.90 * params.ICUCapacity,
                                                               params.ICUCapacity, params.CriticalPerSevere[i], 0);
            double deathsPerSevereAdjusted   = params.CriticalPerSevere[i] - criticalPerSevereAdjusted;

            /**** path of immune-naive ***/
            fN[fE_INS] += y[ENi] / params.TimeExposed[i];
            fN[fINS_R] = params.RecoveredPerInfectedNoSymptoms[i] / params.TimeInfectedNoSymptoms[i] * y[INSNi];
            fN[fINS_ISy] =
                (1 - params.RecoveredPerInfectedNoSymptoms[i]) / params.TimeInfectedNoSymptoms[i] * y[INSNi];
            fN[fINSC_ISyC] =
                (1 - params.RecoveredPerInfectedNoSymptoms[i]) / params.TimeInfectedNoSymptoms[i] * y[INSNCi];
            fN[fINSC_R] = params.RecoveredPerInfectedNoSymptoms[i] / params.TimeInfectedNoSymptoms[i] * y[INSNCi];
            fN[fISy_ISev] = params.SeverePerInfectedSymptoms[i] / params.TimeInfectedSymptoms[i] * y[ISyNi];
            fN[fISy_R]    = (1 - params.SeverePerInfectedSymptoms[i]) / params.TimeInfectedSymptoms[i] * y[ISyNi];
            fN[fISyC_ISev] = params.SeverePerInfectedSymptoms[i] / params.TimeInfectedSymptoms[i] * y[ISyNCi];
            fN[fISyC_R] = (1 - params.SeverePerInfectedSymptoms[i]) / params.TimeInfectedSymptoms[i] * y[ISyNCi];
            fN[fISev_ICr] = criticalPerSevereAdjusted / params.TimeInfectedSevere[i] * y[ISevNi];
            fN[fISev_R]   = (1 - params.CriticalPerSevere[i] - params.DeathsPerSevere[i]) /
                          params.TimeInfectedSevere[i] * y[ISevNi];
            fN[fISev_D] =
                (params.DeathsPerSevere[i] + deathsPerSevereAdjusted) / params.TimeInfectedSevere[i] * y[ISevNi];
            fN[fICr_D] = params.DeathsPerCritical[i] / params.TimeInfectedCritical[i] * y[ICrNi];
            fN[fICr_R] = (1 - params.DeathsPerCritical[i]) / params.TimeInfectedCritical[i] * y[ICrNi];

            /**** path of partially immune ***/
            fPI[fE_INS] += y[EPIi] / params.TimeExposed[i];
            fPI[fINS_R] = (1 - (reducInfectedSymptomsPartialImmunity / reducExposedPartialImmunity) *
                                   (1 - params.RecoveredPerInfectedNoSymptoms[i])) /
                          (params.TimeInfectedNoSymptoms[i] * reducTimeInfectedMild) * y[INSPIi];
            fPI[fINS_ISy] = (reducInfectedSymptomsPartialImmunity / reducExposedPartialImmunity) *
                            (1 - params.RecoveredPerInfectedNoSymptoms[i]) /
                            (params.TimeInfectedNoSymptoms[i] * reducTimeInfectedMild) * y[INSPIi];
            fPI[fINSC_ISyC] = (reducInfectedSymptomsPartialImmunity / reducExposedPartialImmunity) *
                              (1 - params.RecoveredPerInfectedNoSymptoms[i]) /
                              (params.TimeInfectedNoSymptoms[i] * reducTimeInfectedMild) * y[INSPICi];
            fPI[fINSC_R] = (1 - (reducInfectedSymptomsPartialImmunity / reducExposedPartialImmunity) *
                                    (1 - params.RecoveredPerInfectedNoSymptoms[i])) /
                           (params.TimeInfectedNoSymptoms[i] * reducTimeInfectedMild) * y[INSPICi];
            fPI[fISy_ISev] = reducInfectedSevereCriticalDeadPartialImmunity / reducInfectedSymptomsPartialImmunity *
                             params.SeverePerInfectedSymptoms[i] /
                             (params.TimeInfectedSymptoms[i] * reducTimeInfectedMild) * y[ISyPIi];
            fPI[fISy_R] =
                (1 - (reducInfectedSevereCriticalDeadPartialImmunity / reducInfectedSymptomsPartialImmunity) *
                         params.SeverePerInfectedSymptoms[i]) /
                (params.TimeInfectedSymptoms[i] * reducTimeInfectedMild) * y[ISyPIi];
            fPI[fISyC_ISev] = reducInfectedSevereCriticalDeadPartialImmunity / reducInfectedSymptomsPartialImmunity *
                              params.SeverePerInfectedSymptoms[i] /
                              (params.TimeInfectedSymptoms[i] * reducTimeInfectedMild) * y[ISyPICi];
            fPI[fISyC_R] =
                (1 - (reducInfectedSevereCriticalDeadPartialImmunity / reducInfectedSymptomsPartialImmunity) *
                         params.SeverePerInfectedSymptoms[i]) /
                (params.TimeInfectedSymptoms[i] * reducTimeInfectedMild) * y[ISyPICi];
            fPI[fISev_ICr] = reducInfectedSevereCriticalDeadPartialImmunity /
                             reducInfectedSevereCriticalDeadPartialImmunity * criticalPerSevereAdjusted /
                             params.TimeInfectedSevere[i] * y[ISevPIi];
            fPI[fISev_R] =
                (1 - (reducInfectedSevereCriticalDeadPartialImmunity / reducInfectedSevereCriticalDeadPartialImmunity) *
                         (params.CriticalPerSevere[i] + params.DeathsPerSevere[i])) /
                params.TimeInfectedSevere[i] * y[ISevPIi];
            fPI[fISev_D] =
                (reducInfectedSevereCriticalDeadPartialImmunity / reducInfectedSevereCriticalDeadPartialImmunity) *
                (params.DeathsPerSevere[i] + deathsPerSevereAdjusted) / params.TimeInfectedSevere[i] * y[ISevPIi];
            fPI[fICr_D] =
                (reducInfectedSevereCriticalDeadPartialImmunity / reducInfectedSevereCriticalDeadPartialImmunity) *
                params.DeathsPerCritical[i] / params.TimeInfectedCritical[i] * y[ICrPIi];
            fPI[fICr_R] =
                (1 - (reducInfectedSevereCriticalDeadPartialImmunity / reducInfectedSevereCriticalDeadPartialImmunity) *
                         params.DeathsPerCritical[i]) /
                params.TimeInfectedCritical[i] * y[ICrPIi];

            /**** path of improved immunity ***/
            fII[fE_INS] += y[EIIi] / params.TimeExposed[i];
            fII[fINS_R] = (1 - (reducInfectedSymptomsImprovedImmunity / reducExposedImprovedImmunity) *
                                   (1 - params.RecoveredPerInfectedNoSymptoms[i])) /
                          (params.TimeInfectedNoSymptoms[i] * reducTimeInfectedMild) * y[INSIIi];
            fII[fINS_ISy] = (reducInfectedSymptomsImprovedImmunity / reducExposedImprovedImmunity) *
                            (1 - params.RecoveredPerInfectedNoSymptoms[i]) /
                            (params.TimeInfectedNoSymptoms[i] * reducTimeInfectedMild) * y[INSIIi];
            fII[fINSC_ISyC] = (reducInfectedSymptomsImprovedImmunity / reducExposedImprovedImmunity) *
                              (1 - params.RecoveredPerInfectedNoSymptoms[i]) /
                              (params.TimeInfectedNoSymptoms[i] * reducTimeInfectedMild) * y[INSIICi];
            fII[fINSC_R] = (1 - (reducInfectedSymptomsImprovedImmunity / reducExposedImprovedImmunity) *
                                    (1 - params.RecoveredPerInfectedNoSymptoms[i])) /
                           (params.TimeInfectedNoSymptoms[i] * reducTimeInfectedMild) * y[INSIICi];
            fII[fISy_ISev] = reducInfectedSevereCriticalDeadImprovedImmunity / reducInfectedSymptomsImprovedImmunity *
                             params.SeverePerInfectedSymptoms[i] /
                             (params.TimeInfectedSymptoms[i] * reducTimeInfectedMild) * y[ISyIIi];
            fII[fISy_R] =
                (1 - (reducInfectedSevereCriticalDeadImprovedImmunity / reducInfectedSymptomsImprovedImmunity) *
                         params.SeverePerInfectedSymptoms[i]) /
                (params.TimeInfectedSymptoms[i] * reducTimeInfectedMild) * y[ISyIIi];
            fII[fISyC_ISev] = reducInfectedSevereCriticalDeadImprovedImmunity / reducInfectedSymptomsImprovedImmunity *
                              params.SeverePerInfectedSymptoms[i] /
                              (params.TimeInfectedSymptoms[i] * reducTimeInfectedMild) * y[ISyIICi];
            fII[fISyC_R] =
                (1 - (reducInfectedSevereCriticalDeadImprovedImmunity / reducInfectedSymptomsImprovedImmunity) *
                         params.SeverePerInfectedSymptoms[i]) /
                (params.TimeInfectedSymptoms[i] * reducTimeInfectedMild) * y[ISyIICi];
            fII[fISev_ICr] = reducInfectedSevereCriticalDeadImprovedImmunity /
                             reducInfectedSevereCriticalDeadImprovedImmunity * criticalPerSevereAdjusted /
                             params.TimeInfectedSevere[i] * y[ISevIIi];
            fII[fISev_R] = (1 - (reducInfectedSevereCriticalDeadImprovedImmunity /
                                 reducInfectedSevereCriticalDeadImprovedImmunity) *
                                    (params.CriticalPerSevere[i] + params.DeathsPerSevere[i])) /
                           params.TimeInfectedSevere[i] * y[ISevIIi];
            fII[fISev_D] =
                (reducInfectedSevereCriticalDeadImprovedImmunity / reducInfectedSevereCriticalDeadImprovedImmunity) *
                (params.DeathsPerSevere[i] + deathsPerSevereAdjusted) / params.TimeInfectedSevere[i] * y[ISevIIi];
            fII[fICr_D] =
                (reducInfectedSevereCriticalDeadImprovedImmunity / reducInfectedSevereCriticalDeadImprovedImmunity) *
                params.DeathsPerCritical[i] / params.TimeInfectedCritical[i] * y[ICrIIi];
            fII[fICr_R] = (1 - (reducInfectedSevereCriticalDeadImprovedImmunity /
                                reducInfectedSevereCriticalDeadImprovedImmunity) *
                                   params.DeathsPerCritical[i]) /
                          params.TimeInfectedCritical[i] * y[ICrIIi];
        }
    }

    mutable Vec m_flow_values;
    void get_derivatives(const Vec& pop, const Vec& y, double t, Vec& dydt) const
    {
        const int G = parameters.G;
        m_flow_values.assign(size_t(G) * NF, 0.0);
        get_flows(pop, y, t, m_flow_values);
        std::fill(dydt.begin(), dydt.end(), 0.0);
        for (int i = 0; i < G; ++i) {
            for (int k = 0; k < NF; ++k) {
                dydt[i * NC + FLOW_SRC[k]] -= m_flow_values[i * NF + k];
                dydt[i * NC + FLOW_DST[k]] += m_flow_values[i * NF + k];
            }
        }
    }
    void eval_right_hand_side(const Vec& pop, const Vec& y, double t, Vec& dydt) const
    {
        std::fill(dydt.begin(), dydt.end(), 0.0);
        get_derivatives(pop, y, t, dydt);
    }

    // model.h:846-885 (used when the node type is osecirvvs::Simulation)
    Vec get_mobility_factors(double /*t*/, const Vec& y) const
    {
        if (!model_sim) {
            return Vec(y.size(), 1.0);
        }
        const Params& p                = parameters;
        double test_and_trace_required = 0.0;
        for (int i = 0; i < p.G; ++i) {
            const double y_INS = y[i * NC + INSN] + y[i * NC + INSPI] + y[i * NC + INSII];
            test_and_trace_required += (1 - p.RecoveredPerInfectedNoSymptoms[i]) / p.TimeInfectedNoSymptoms[i] * y_INS;
        }
        Vec factors(y.size(), 1.0);
        for (int i = 0; i < p.G; ++i) {
            const double r = smoother_cosine(test_and_trace_required, p.TestAndTraceCapacity,
                                             p.TestAndTraceCapacity * p.TestAndTraceCapacityMaxRiskSymptoms,
                                             p.RiskOfInfectionFromSymptomatic[i],
                                             p.MaxRiskOfInfectionFromSymptomatic[i]);
            factors[i * NC + ISyN]  = r;
            factors[i * NC + ISyPI] = r;
            factors[i * NC + ISyII] = r;
        }
        return factors;
    }

    // model.h:887-940
    void test_commuters(Vec& last_value, Vec& mobile_population, double time) const
    {
        if (!model_sim) {
            return;
        }
        const Params& p     = parameters;
        double nondetection = 1.0;
        if (time >= p.start_commuter_detection && time < p.end_commuter_detection) {
            nondetection = p.commuter_nondetection;
        }
        for (int i = 0; i < p.G; ++i) {
            const int o = i * NC;
            const int pairs[6][2] = {{ISyN, ISyNC},   {INSN, INSNC},   {ISyPI, ISyPIC},
                                     {INSPI, INSPIC}, {ISyII, ISyIIC}, {INSII, INSIIC}};
            for (auto& pr : pairs) {
                last_value[o + pr[0]] -= mobile_population[o + pr[0]] * (1 - nondetection);
                last_value[o + pr[1]] += mobile_population[o + pr[0]] * (1 - nondetection);
            }
            for (int c : {ISyN, INSN, ISyPI, INSPI, ISyII, INSII}) {
                mobile_population[o + c] *= nondetection;
            }
        }
    }
};

} // namespace osecirvvs

// osecirvvs::Simulation, model.h:568-771 (dynamic NPIs empty). `model_sim == false` gives the plain mio::Simulation
// the reference's graph benchmark instantiates (cpp/benchmarks/graph_simulation.cpp:121-126): no vaccination, unit
// mobility factors, no commuter testing.
template <>
struct Simulation<osecirvvs::Model> {
    using M = osecirvvs::Model;
    M model;
    SystemIntegrator integrator;
    TimeSeries result;
    double dt;
    bool model_sim = true;
    Simulation(const M& m, double t0 = 0., double dt_ = 0.1)
        : model(m)
        , integrator(std::make_unique<CashKarp54Core>())
        , dt(dt_)
        , model_sim(m.model_sim)
    {
        result.add(t0, m.populations);
    }
    void set_integrator_core(std::unique_ptr<IntegratorCore> c)
    {
        integrator.core = std::move(c);
    }
    Vec& base_advance(double tmax)
    {
        return integrator.advance(
            [this](const Vec& y, double t, Vec& dydt) {
                model.eval_right_hand_side(y, y, t, dydt);
            },
            tmax, dt, result);
    }
    // model.h:576-597
    void apply_variant(double t, const Vec& base_infectiousness)
    {
        auto& p = model.parameters;
        if (p.StartDay + t >= p.StartDayNewVariant - 1e-10) {
            const double days_variant      = t - (p.StartDayNewVariant - p.StartDay);
            const double share_new_variant = std::min(1.0, 0.01 * std::pow(2, (1. / 7) * days_variant));
            for (int i = 0; i < p.G; ++i) {
                double new_transmission = (1 - share_new_variant) * base_infectiousness[i] +
                                          share_new_variant * base_infectiousness[i] * p.InfectiousnessNewVariant[i];
                p.TransmissionProbabilityOnContact[i] = new_transmission;
            }
        }
    }
    // model.h:599-649
    void apply_vaccination(double t)
    {
        const size_t t_idx = size_t(std::floor(t));
        auto& p            = model.parameters;
        Vec& last_value    = result.last();
        const int count = osecirvvs::NC, S = osecirvvs::SN, SV = osecirvvs::SPI, R = osecirvvs::SII;
        for (int i = 0; i < p.G; ++i) {
            double first_vacc, full_vacc;
            if (t_idx == 0) {
                first_vacc = p.DailyPartialVaccinations[i][t_idx];
                full_vacc  = p.DailyFullVaccinations[i][t_idx];
            }
            else {
                first_vacc = p.DailyPartialVaccinations[i][t_idx] - p.DailyPartialVaccinations[i][t_idx - 1];
                full_vacc  = p.DailyFullVaccinations[i][t_idx] - p.DailyFullVaccinations[i][t_idx - 1];
            }
            if (last_value[count * i + S] - first_vacc < 0) {
                first_vacc = 0.99 * last_value[count * i + S];
            }
            last_value[count * i + S] -= first_vacc;
            last_value[count * i + SV] += first_vacc;
            if (last_value[count * i + SV] - full_vacc < 0) {
                full_vacc = 0.99 * last_value[count * i + SV];
            }
            last_value[count * i + SV] -= full_vacc;
            last_value[count * i + R] += full_vacc;
        }
    }
    // model.h:651-771
    Vec& advance(double tmax)
    {
        if (!model_sim) {
            return base_advance(tmax);
        }
        Vec base_infectiousness = model.parameters.TransmissionProbabilityOnContact;
        double t                = result.times.back();
        while (t < tmax) {
            if (t == 0) {
                apply_variant(t, base_infectiousness);
            }
            double dt_eff = std::min(1.0, tmax - t);
            base_advance(t + dt_eff);
            if (t + 0.5 + dt_eff - std::floor(t + 0.5) >= 1) {
                apply_vaccination(t + 0.5 + dt_eff);
                apply_variant(t, base_infectiousness);
            }
            t = t + dt_eff;
        }
        model.parameters.TransmissionProbabilityOnContact = base_infectiousness;
        return result.last();
    }
};

} // namespace orc

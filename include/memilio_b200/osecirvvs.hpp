// osecirvvs.hpp -- mio::osecirvvs::Model / Parameters / Simulation / simulate with the reference's names
// (cpp/models/ode_secirvvs/infection_state.h:36-67, parameters.h:42-640, model.h:93-110,651-790), packed into the flat
// descriptors of mio_cuda.h (MIO_MODEL_OSECIRVVS) and run on the GPU. The right-hand side, the day-chunk advance with
// apply_vaccination / apply_variant and the mobility hooks live in csrc/ (integrate.cu, edges.cu).
#pragma once

#include "mio.hpp"

namespace mio
{

// SimulationDay, memilio/epidemiology/simulation_day.h
struct SimulationDay {
    size_t value = 0;
    SimulationDay() = default;
    explicit SimulationDay(size_t v)
        : value(v)
    {
    }
    explicit SimulationDay(int v)
        : value(size_t(v))
    {
    }
    size_t get() const
    {
        return value;
    }
};

namespace osecirvvs
{

enum class InfectionState
{
    SusceptibleNaive = 0,
    SusceptiblePartialImmunity,
    ExposedNaive,
    ExposedPartialImmunity,
    ExposedImprovedImmunity,
    InfectedNoSymptomsNaive,
    InfectedNoSymptomsPartialImmunity,
    InfectedNoSymptomsImprovedImmunity,
    InfectedNoSymptomsNaiveConfirmed,
    InfectedNoSymptomsPartialImmunityConfirmed,
    InfectedNoSymptomsImprovedImmunityConfirmed,
    InfectedSymptomsNaive,
    InfectedSymptomsPartialImmunity,
    InfectedSymptomsImprovedImmunity,
    InfectedSymptomsNaiveConfirmed,
    InfectedSymptomsPartialImmunityConfirmed,
    InfectedSymptomsImprovedImmunityConfirmed,
    InfectedSevereNaive,
    InfectedSeverePartialImmunity,
    InfectedSevereImprovedImmunity,
    InfectedCriticalNaive,
    InfectedCriticalPartialImmunity,
    InfectedCriticalImprovedImmunity,
    SusceptibleImprovedImmunity,
    DeadNaive,
    DeadPartialImmunity,
    DeadImprovedImmunity,
    Count
};

// CustomIndexArray<FP, AgeGroup, SimulationDay> of the cumulative vaccination numbers (parameters.h:469-499)
template <class FP>
class DailyVaccinationArray
{
public:
    struct Index {
        AgeGroup age;
        SimulationDay day;
    };
    explicit DailyVaccinationArray(AgeGroup groups = AgeGroup(0))
        : m_G(groups.value)
    {
    }
    void resize(SimulationDay days)
    {
        std::vector<FP> v(m_G * days.value, FP(0));
        for (size_t g = 0; g < m_G; ++g) {
            for (size_t d = 0; d < std::min(m_days, days.value); ++d) {
                v[g * days.value + d] = m_v[g * m_days + d];
            }
        }
        m_v    = std::move(v);
        m_days = days.value;
    }
    FP& operator[](Index i)
    {
        return m_v[i.age.value * m_days + i.day.value];
    }
    const FP& operator[](Index i) const
    {
        return m_v[i.age.value * m_days + i.day.value];
    }
    template <class Tag>
    auto size() const
    {
        if constexpr (std::is_same_v<Tag, SimulationDay>) {
            return SimulationDay(m_days);
        }
        else {
            return AgeGroup(m_G);
        }
    }
    size_t num_days() const
    {
        return m_days;
    }
    const std::vector<FP>& array() const // [G][days]
    {
        return m_v;
    }

private:
    size_t m_G, m_days = 0;
    std::vector<FP> m_v;
};

// parameter tags, ode_secirvvs/parameters.h:42-640
#define MIO_B200_VVS_SCALAR_TAG(NAME, TYPE, DEFAULT)                                                                   \
    template <class FP = ScalarType>                                                                                   \
    struct NAME {                                                                                                      \
        using Type = TYPE;                                                                                             \
        static Type get_default(AgeGroup)                                                                              \
        {                                                                                                              \
            return Type(DEFAULT);                                                                                      \
        }                                                                                                              \
    };
#define MIO_B200_VVS_AGE_TAG(NAME, DEFAULT)                                                                            \
    template <class FP = ScalarType>                                                                                   \
    struct NAME {                                                                                                      \
        using Type = AgeArray<UncertainValue<FP>>;                                                                     \
        static Type get_default(AgeGroup n)                                                                            \
        {                                                                                                              \
            return Type(n.value, UncertainValue<FP>(DEFAULT));                                                         \
        }                                                                                                              \
    };
MIO_B200_VVS_SCALAR_TAG(StartDay, FP, 0.0)
MIO_B200_VVS_SCALAR_TAG(StartDayNewVariant, FP, std::numeric_limits<FP>::max())
MIO_B200_VVS_SCALAR_TAG(Seasonality, UncertainValue<FP>, 0.0)
MIO_B200_VVS_SCALAR_TAG(ICUCapacity, UncertainValue<FP>, std::numeric_limits<FP>::max())
MIO_B200_VVS_SCALAR_TAG(TestAndTraceCapacity, UncertainValue<FP>, std::numeric_limits<FP>::max())
MIO_B200_VVS_SCALAR_TAG(TestAndTraceCapacityMaxRiskNoSymptoms, UncertainValue<FP>, 2.0)
MIO_B200_VVS_SCALAR_TAG(TestAndTraceCapacityMaxRiskSymptoms, UncertainValue<FP>, 15.0)
MIO_B200_VVS_AGE_TAG(TimeExposed, 1.0)
MIO_B200_VVS_AGE_TAG(TimeInfectedNoSymptoms, 1.0)
MIO_B200_VVS_AGE_TAG(TimeInfectedSymptoms, 1.0)
MIO_B200_VVS_AGE_TAG(TimeInfectedSevere, 1.0)
MIO_B200_VVS_AGE_TAG(TimeInfectedCritical, 1.0)
MIO_B200_VVS_AGE_TAG(TransmissionProbabilityOnContact, 1.0)
MIO_B200_VVS_AGE_TAG(RelativeTransmissionNoSymptoms, 1.0)
MIO_B200_VVS_AGE_TAG(RecoveredPerInfectedNoSymptoms, 0.0)
MIO_B200_VVS_AGE_TAG(RiskOfInfectionFromSymptomatic, 1.0)
MIO_B200_VVS_AGE_TAG(MaxRiskOfInfectionFromSymptomatic, 1.0)
MIO_B200_VVS_AGE_TAG(SeverePerInfectedSymptoms, 0.0)
MIO_B200_VVS_AGE_TAG(CriticalPerSevere, 0.0)
MIO_B200_VVS_AGE_TAG(DeathsPerSevere, 0.0)
MIO_B200_VVS_AGE_TAG(DeathsPerCritical, 0.0)
MIO_B200_VVS_AGE_TAG(VaccinationGap, 49.0)
MIO_B200_VVS_AGE_TAG(DaysUntilEffectivePartialImmunity, 14.0)
MIO_B200_VVS_AGE_TAG(DaysUntilEffectiveImprovedImmunity, 7.0)
MIO_B200_VVS_AGE_TAG(ReducExposedPartialImmunity, 1.0)
MIO_B200_VVS_AGE_TAG(ReducExposedImprovedImmunity, 1.0)
MIO_B200_VVS_AGE_TAG(ReducInfectedSymptomsPartialImmunity, 1.0)
MIO_B200_VVS_AGE_TAG(ReducInfectedSymptomsImprovedImmunity, 1.0)
MIO_B200_VVS_AGE_TAG(ReducInfectedSevereCriticalDeadPartialImmunity, 1.0)
MIO_B200_VVS_AGE_TAG(ReducInfectedSevereCriticalDeadImprovedImmunity, 1.0)
MIO_B200_VVS_AGE_TAG(ReducTimeInfectedMild, 0.5)
#undef MIO_B200_VVS_SCALAR_TAG
#undef MIO_B200_VVS_AGE_TAG
template <class FP = ScalarType>
struct InfectiousnessNewVariant {
    using Type = AgeArray<FP>;
    static Type get_default(AgeGroup n)
    {
        return Type(n.value, FP(1.0));
    }
};
template <class FP = ScalarType>
struct DailyPartialVaccinations {
    using Type = DailyVaccinationArray<FP>;
    static Type get_default(AgeGroup n)
    {
        return Type(n);
    }
};
template <class FP = ScalarType>
struct DailyFullVaccinations {
    using Type = DailyVaccinationArray<FP>;
    static Type get_default(AgeGroup n)
    {
        return Type(n);
    }
};
template <class FP = ScalarType>
struct ContactPatterns {
    using Type = UncertainContactMatrix<FP>;
    static Type get_default(AgeGroup n)
    {
        return Type(1, n.value);
    }
};

template <class FP = ScalarType>
class Parameters
{
    template <class Tag>
    struct Slot {
        typename Tag::Type value;
    };
    using Tuple = std::tuple<
        Slot<StartDay<FP>>, Slot<StartDayNewVariant<FP>>, Slot<Seasonality<FP>>, Slot<ICUCapacity<FP>>,
        Slot<TestAndTraceCapacity<FP>>, Slot<TestAndTraceCapacityMaxRiskNoSymptoms<FP>>,
        Slot<TestAndTraceCapacityMaxRiskSymptoms<FP>>, Slot<ContactPatterns<FP>>, Slot<TimeExposed<FP>>,
        Slot<TimeInfectedNoSymptoms<FP>>, Slot<TimeInfectedSymptoms<FP>>, Slot<TimeInfectedSevere<FP>>,
        Slot<TimeInfectedCritical<FP>>, Slot<TransmissionProbabilityOnContact<FP>>,
        Slot<RelativeTransmissionNoSymptoms<FP>>, Slot<RecoveredPerInfectedNoSymptoms<FP>>,
        Slot<RiskOfInfectionFromSymptomatic<FP>>, Slot<MaxRiskOfInfectionFromSymptomatic<FP>>,
        Slot<SeverePerInfectedSymptoms<FP>>, Slot<CriticalPerSevere<FP>>, Slot<DeathsPerSevere<FP>>,
        Slot<DeathsPerCritical<FP>>, Slot<VaccinationGap<FP>>, Slot<DaysUntilEffectivePartialImmunity<FP>>,
        Slot<DaysUntilEffectiveImprovedImmunity<FP>>, Slot<DailyPartialVaccinations<FP>>,
        Slot<DailyFullVaccinations<FP>>, Slot<ReducExposedPartialImmunity<FP>>, Slot<ReducExposedImprovedImmunity<FP>>,
        Slot<ReducInfectedSymptomsPartialImmunity<FP>>, Slot<ReducInfectedSymptomsImprovedImmunity<FP>>,
        Slot<ReducInfectedSevereCriticalDeadPartialImmunity<FP>>,
        Slot<ReducInfectedSevereCriticalDeadImprovedImmunity<FP>>, Slot<ReducTimeInfectedMild<FP>>,
        Slot<InfectiousnessNewVariant<FP>>>;

public:
    explicit Parameters(AgeGroup num_groups)
        : m_num_groups(num_groups)
        , m_t(make(num_groups, static_cast<Tuple*>(nullptr)))
    {
    }
    AgeGroup get_num_groups() const
    {
        return m_num_groups;
    }
    template <class Tag>
    typename Tag::Type& get()
    {
        return std::get<Slot<Tag>>(m_t).value;
    }
    template <class Tag>
    const typename Tag::Type& get() const
    {
        return std::get<Slot<Tag>>(m_t).value;
    }
    template <class Tag, class V>
    void set(V&& v)
    {
        get<Tag>() = std::forward<V>(v);
    }
    FP& get_commuter_nondetection()
    {
        return m_commuter_nondetection;
    }
    FP get_commuter_nondetection() const
    {
        return m_commuter_nondetection;
    }
    FP& get_start_commuter_detection()
    {
        return m_start_commuter_detection;
    }
    FP get_start_commuter_detection() const
    {
        return m_start_commuter_detection;
    }
    FP& get_end_commuter_detection()
    {
        return m_end_commuter_detection;
    }
    FP get_end_commuter_detection() const
    {
        return m_end_commuter_detection;
    }

    // flat image: MIO_PH_* header, MIO_PA_* / MIO_PV_* age records of mio_cuda.h
    void pack_values(double* pset, double* local) const
    {
        const size_t G = m_num_groups.value;
        std::fill(pset, pset + MIO_PSET_HEADER + G * MIO_OSECIRVVS_AGE_STRIDE, 0.0);
        pset[MIO_PH_START_DAY]                = get<StartDay<FP>>();
        pset[MIO_PH_SEASONALITY]              = get<Seasonality<FP>>().value();
        pset[MIO_PH_TNT_MAX_RISK]             = get<TestAndTraceCapacityMaxRiskSymptoms<FP>>().value();
        pset[MIO_PH_TNT_MAX_RISK_NOSYMP]      = get<TestAndTraceCapacityMaxRiskNoSymptoms<FP>>().value();
        pset[MIO_PH_START_DAY_NEW_VARIANT]    = get<StartDayNewVariant<FP>>();
        pset[MIO_PH_COMMUTER_NONDETECTION]    = m_commuter_nondetection;
        pset[MIO_PH_START_COMMUTER_DETECTION] = m_start_commuter_detection;
        pset[MIO_PH_END_COMMUTER_DETECTION]   = m_end_commuter_detection;
        local[MIO_NL_ICU_CAPACITY]            = get<ICUCapacity<FP>>().value();
        local[MIO_NL_TNT_CAPACITY]            = get<TestAndTraceCapacity<FP>>().value();
        for (size_t i = 0; i < G; ++i) {
            double* a = pset + MIO_PSET_HEADER + i * MIO_OSECIRVVS_AGE_STRIDE;
            const AgeGroup g(i);
            a[MIO_PA_TIME_EXPOSED]          = get<TimeExposed<FP>>()[g].value();
            a[MIO_PA_TIME_INS]              = get<TimeInfectedNoSymptoms<FP>>()[g].value();
            a[MIO_PA_TIME_ISY]              = get<TimeInfectedSymptoms<FP>>()[g].value();
            a[MIO_PA_TIME_ISEV]             = get<TimeInfectedSevere<FP>>()[g].value();
            a[MIO_PA_TIME_ICR]              = get<TimeInfectedCritical<FP>>()[g].value();
            a[MIO_PA_TRANS_PROB]            = get<TransmissionProbabilityOnContact<FP>>()[g].value();
            a[MIO_PA_REL_TRANS_NOSYMP]      = get<RelativeTransmissionNoSymptoms<FP>>()[g].value();
            a[MIO_PA_REC_PER_INS]           = get<RecoveredPerInfectedNoSymptoms<FP>>()[g].value();
            a[MIO_PA_RISK_SYMP]             = get<RiskOfInfectionFromSymptomatic<FP>>()[g].value();
            a[MIO_PA_MAX_RISK_SYMP]         = get<MaxRiskOfInfectionFromSymptomatic<FP>>()[g].value();
            a[MIO_PA_SEV_PER_ISY]           = get<SeverePerInfectedSymptoms<FP>>()[g].value();
            a[MIO_PA_CRIT_PER_SEV]          = get<CriticalPerSevere<FP>>()[g].value();
            a[MIO_PA_DEATHS_PER_SEV]        = get<DeathsPerSevere<FP>>()[g].value();
            a[MIO_PA_DEATHS_PER_CRIT]       = get<DeathsPerCritical<FP>>()[g].value();
            a[MIO_PV_REDUC_EXPOSED_PI]      = get<ReducExposedPartialImmunity<FP>>()[g].value();
            a[MIO_PV_REDUC_EXPOSED_II]      = get<ReducExposedImprovedImmunity<FP>>()[g].value();
            a[MIO_PV_REDUC_INF_SYMP_PI]     = get<ReducInfectedSymptomsPartialImmunity<FP>>()[g].value();
            a[MIO_PV_REDUC_INF_SYMP_II]     = get<ReducInfectedSymptomsImprovedImmunity<FP>>()[g].value();
            a[MIO_PV_REDUC_SEV_CRIT_DEAD_PI] = get<ReducInfectedSevereCriticalDeadPartialImmunity<FP>>()[g].value();
            a[MIO_PV_REDUC_SEV_CRIT_DEAD_II] = get<ReducInfectedSevereCriticalDeadImprovedImmunity<FP>>()[g].value();
            a[MIO_PV_REDUC_TIME_MILD]       = get<ReducTimeInfectedMild<FP>>()[g].value();
            a[MIO_PV_INFECTIOUSNESS_NEW_VARIANT] = get<InfectiousnessNewVariant<FP>>()[g];
            a[MIO_PV_VACCINATION_GAP]            = get<VaccinationGap<FP>>()[g].value();
            a[MIO_PV_DAYS_UNTIL_EFFECTIVE_PI]    = get<DaysUntilEffectivePartialImmunity<FP>>()[g].value();
            a[MIO_PV_DAYS_UNTIL_EFFECTIVE_II]    = get<DaysUntilEffectiveImprovedImmunity<FP>>()[g].value();
        }
    }

    // parameters.h:660-860: out-of-range values are corrected (and reported by the return value)
    bool apply_constraints()
    {
        bool corrected = false;
        auto fix = [&](UncertainValue<FP>& v, bool bad, FP to) {
            if (bad) {
                v         = to;
                corrected = true;
            }
        };
        fix(get<Seasonality<FP>>(), get<Seasonality<FP>>().value() < 0.0 || get<Seasonality<FP>>().value() > 0.5, 0.0);
        fix(get<ICUCapacity<FP>>(), get<ICUCapacity<FP>>().value() < 0.0, 0.0);
        fix(get<TestAndTraceCapacity<FP>>(), get<TestAndTraceCapacity<FP>>().value() < 0.0, 0.0);
        fix(get<TestAndTraceCapacityMaxRiskSymptoms<FP>>(), get<TestAndTraceCapacityMaxRiskSymptoms<FP>>().value() < 0.0,
            0.0);
        fix(get<TestAndTraceCapacityMaxRiskNoSymptoms<FP>>(),
            get<TestAndTraceCapacityMaxRiskNoSymptoms<FP>>().value() < 0.0, 0.0);
        const FP tol_times = 1e-1;
        for (auto i = AgeGroup(0); i < m_num_groups; ++i) {
            auto time = [&](UncertainValue<FP>& v) {
                fix(v, v.value() < tol_times, tol_times);
            };
            auto prob = [&](UncertainValue<FP>& v) {
                fix(v, v.value() < 0.0 || v.value() > 1.0, 0.0);
            };
            auto reduc = [&](UncertainValue<FP>& v) {
                fix(v, v.value() <= 0.0 || v.value() > 1.0, 1.0);
            };
            time(get<TimeExposed<FP>>()[i]);
            time(get<TimeInfectedNoSymptoms<FP>>()[i]);
            time(get<TimeInfectedSymptoms<FP>>()[i]);
            time(get<TimeInfectedSevere<FP>>()[i]);
            time(get<TimeInfectedCritical<FP>>()[i]);
            prob(get<TransmissionProbabilityOnContact<FP>>()[i]);
            fix(get<RelativeTransmissionNoSymptoms<FP>>()[i], get<RelativeTransmissionNoSymptoms<FP>>()[i].value() < 0.0,
                0.0);
            prob(get<RecoveredPerInfectedNoSymptoms<FP>>()[i]);
            prob(get<RiskOfInfectionFromSymptomatic<FP>>()[i]);
            prob(get<SeverePerInfectedSymptoms<FP>>()[i]);
            prob(get<CriticalPerSevere<FP>>()[i]);
            prob(get<DeathsPerCritical<FP>>()[i]);
            fix(get<DaysUntilEffectivePartialImmunity<FP>>()[i],
                get<DaysUntilEffectivePartialImmunity<FP>>()[i].value() < 0.0, 0.0);
            fix(get<DaysUntilEffectiveImprovedImmunity<FP>>()[i],
                get<DaysUntilEffectiveImprovedImmunity<FP>>()[i].value() < 0.0, 0.0);
            reduc(get<ReducExposedPartialImmunity<FP>>()[i]);
            reduc(get<ReducExposedImprovedImmunity<FP>>()[i]);
            reduc(get<ReducInfectedSymptomsPartialImmunity<FP>>()[i]);
            reduc(get<ReducInfectedSymptomsImprovedImmunity<FP>>()[i]);
            reduc(get<ReducInfectedSevereCriticalDeadPartialImmunity<FP>>()[i]);
            reduc(get<ReducInfectedSevereCriticalDeadImprovedImmunity<FP>>()[i]);
            reduc(get<ReducTimeInfectedMild<FP>>()[i]);
            if (get<InfectiousnessNewVariant<FP>>()[i] < 0.0) {
                get<InfectiousnessNewVariant<FP>>()[i] = 1.0;
                corrected                              = true;
            }
            fix(get<VaccinationGap<FP>>()[i], get<VaccinationGap<FP>>()[i].value() < 0.0, 0.0);
        }
        return corrected;
    }
    bool check_constraints() const
    {
        Parameters copy(*this);
        return copy.apply_constraints();
    }

private:
    template <class... S>
    static std::tuple<S...> make(AgeGroup n, std::tuple<S...>*)
    {
        return std::tuple<S...>{S{tag_of<S>::type::get_default(n)}...};
    }
    template <class S>
    struct tag_of;
    template <class Tag>
    struct tag_of<Slot<Tag>> {
        using type = Tag;
    };
    AgeGroup m_num_groups;
    Tuple m_t;
    FP m_commuter_nondetection    = 0.0;
    FP m_start_commuter_detection = 0.0;
    FP m_end_commuter_detection   = 0.0;
};

template <class FP = ScalarType>
class Model
{
public:
    using Compartments = InfectionState;
    using Populations  = mio::Populations<FP, InfectionState>;
    using ParameterSet = Parameters<FP>;
    explicit Model(int num_agegroups)
        : populations(size_t(num_agegroups))
        , parameters(AgeGroup(size_t(num_agegroups)))
    {
    }
    explicit Model(size_t num_agegroups)
        : Model(int(num_agegroups))
    {
    }
    bool apply_constraints()
    {
        return (int(parameters.apply_constraints()) | int(populations.apply_constraints())) != 0;
    }
    bool check_constraints() const
    {
        Model c(*this);
        return c.apply_constraints();
    }
    Eigen::VectorX<FP> get_initial_values() const
    {
        return populations.get_compartments();
    }
    Populations populations;
    ParameterSet parameters;
};

namespace detail
{
constexpr int NC = int(InfectionState::Count);
}

// osecirvvs::Simulation (model.h:651-790): one node, advanced in day chunks with apply_vaccination / apply_variant.
// advance() runs the node on the GPU as a 1-node graph (sim_kind MIO_SIM_MODEL) and appends every accepted integrator
// step to the result. BASE = true gives the plain mio::Simulation<FP, osecirvvs::Model<FP>> (no vaccination, no
// variant, MIO_SIM_BASE) that the reference's graph benchmark instantiates.
template <class FP = ScalarType, bool BASE = false>
class SimulationT
{
public:
    using Model = osecirvvs::Model<FP>;
    SimulationT(const Model& model, FP t0 = 0., FP dt = 0.1)
        : m_model(model)
        , m_t0(t0)
        , m_dt(dt)
        , m_result(t0, model.get_initial_values())
        , m_core(std::make_shared<DefaultIntegratorCore<FP>>())
    {
    }
    void set_integrator_core(std::unique_ptr<OdeIntegratorCore<FP>>&& core)
    {
        m_core = std::move(core);
    }
    const OdeIntegratorCore<FP>& get_integrator_core() const
    {
        return *m_core;
    }
    TimeSeries<FP>& get_result()
    {
        return m_result;
    }
    const TimeSeries<FP>& get_result() const
    {
        return m_result;
    }
    const Model& get_model() const
    {
        return m_model;
    }
    Model& get_model()
    {
        return m_model;
    }
    FP get_dt() const
    {
        return m_dt;
    }
    FP& get_dt()
    {
        return m_dt;
    }
    Eigen::VectorX<FP> advance(FP tmax)
    {
        const int G    = int(m_model.parameters.get_num_groups().value);
        const int n    = G * detail::NC;
        const auto& cm = static_cast<const ContactMatrixGroup<FP>&>(
            m_model.parameters.template get<ContactPatterns<FP>>());
        const int L = int(cm.get_num_matrices()), K = std::max(1, osecir::detail::num_change_points(cm));
        const auto& vp = m_model.parameters.template get<DailyPartialVaccinations<FP>>();
        const auto& vf = m_model.parameters.template get<DailyFullVaccinations<FP>>();
        if (vp.num_days() != vf.num_days()) {
            throw std::runtime_error("osecirvvs::Simulation: DailyPartialVaccinations and DailyFullVaccinations must "
                                     "cover the same days");
        }
        mio_graph_desc d{};
        d.model         = MIO_MODEL_OSECIRVVS;
        d.sim_kind      = BASE ? MIO_SIM_BASE : MIO_SIM_MODEL;
        d.num_groups    = G;
        d.num_nodes     = 1;
        d.num_psets     = 1;
        d.num_ctables   = 1;
        d.ct_max_loc    = L;
        d.ct_max_pts    = K;
        d.max_runs      = 1;
        d.traj_capacity = 1 << 16;
        d.integrator    = m_core->desc();
        d.t0            = m_result.get_last_time();
        d.dt_graph      = std::numeric_limits<double>::infinity();
        d.vacc_days     = int32_t(vp.num_days());
        d.vacc_partial  = vp.array().data();
        d.vacc_full     = vf.array().data();
        const size_t psz = size_t(mio_pset_size(MIO_MODEL_OSECIRVVS, G));
        std::vector<double> pset(psz, 0.0), local(MIO_NODE_LOCAL);
        m_model.parameters.pack_values(pset.data(), local.data());
        std::vector<double> ct = osecir::detail::contact_table(cm, G, L, K);
        std::vector<double> y0(m_result.get_last_value().data(), m_result.get_last_value().data() + n);
        mio_run_data rd{pset.data(), ct.data(), local.data(), y0.data(), &m_dt, nullptr};
        mio_handle* h = nullptr;
        cuda_check(mio_cuda_graph_create(&d, 0, &h));
        try {
            cuda_check(mio_cuda_set_runs(h, 0, 1, &rd));
            cuda_check(mio_cuda_advance(h, tmax));
            int32_t npts     = 0;
            const size_t cap = size_t(d.traj_capacity);
            std::vector<double> times(cap), values(cap * size_t(n));
            cuda_check(mio_cuda_get_trajectory(h, 0, 1, &npts, times.data(), values.data()));
            if (npts > d.traj_capacity) {
                throw std::runtime_error("Simulation::advance: more integrator steps than the trajectory capacity");
            }
            for (int k = 1; k < npts; ++k) { // point 0 is the start value, already in the result
                Eigen::VectorX<FP> v(n);
                std::copy(values.begin() + size_t(k) * n, values.begin() + size_t(k + 1) * n, v.data());
                m_result.add_time_point(times[size_t(k)], v);
            }
            cuda_check(mio_cuda_get_node_dt(h, 0, 1, &m_dt));
        }
        catch (...) {
            mio_cuda_graph_destroy(h);
            throw;
        }
        mio_cuda_graph_destroy(h);
        return m_result.get_last_value();
    }

private:
    Model m_model;
    FP m_t0, m_dt;
    TimeSeries<FP> m_result;
    std::shared_ptr<OdeIntegratorCore<FP>> m_core;
};

template <class FP = ScalarType>
using Simulation = SimulationT<FP, false>;

// osecirvvs::simulate, model.h:792-818
template <class FP = ScalarType>
TimeSeries<FP> simulate(FP t0, FP tmax, FP dt, const Model<FP>& model,
                        std::unique_ptr<OdeIntegratorCore<FP>>&& integrator = nullptr)
{
    model.check_constraints();
    Simulation<FP> sim(model, t0, dt);
    if (integrator) {
        sim.set_integrator_core(std::move(integrator));
    }
    sim.advance(tmax);
    return sim.get_result();
}

} // namespace osecirvvs

namespace detail
{
template <class FP, bool BASE>
struct SimTraits<osecirvvs::SimulationT<FP, BASE>> {
    static constexpr bool flow_simulation = false;
    static bool has_dynamic_npis(const osecirvvs::Model<FP>&) // (not part of this backend for ODE-SECIRVVS)
    {
        return false;
    }
    static bool check_dynamic_npis(osecirvvs::SimulationT<FP, BASE>&, FP, FP)
    {
        return false;
    }
    static constexpr int model_kind = MIO_MODEL_OSECIRVVS;
    static constexpr int sim_kind   = BASE ? MIO_SIM_BASE : MIO_SIM_MODEL;
    static constexpr int NC         = osecirvvs::detail::NC;
    static const ContactMatrixGroup<FP>& contacts(const osecirvvs::Model<FP>& m)
    {
        return static_cast<const ContactMatrixGroup<FP>&>(m.parameters.template get<osecirvvs::ContactPatterns<FP>>());
    }
    static void vaccinations(const osecirvvs::Model<FP>& m, std::vector<double>& partial, std::vector<double>& full,
                             int& days)
    {
        const auto& vp = m.parameters.template get<osecirvvs::DailyPartialVaccinations<FP>>();
        const auto& vf = m.parameters.template get<osecirvvs::DailyFullVaccinations<FP>>();
        if (vp.num_days() != vf.num_days()) {
            throw std::runtime_error("osecirvvs: DailyPartialVaccinations and DailyFullVaccinations must cover the "
                                     "same days");
        }
        days = int(vp.num_days());
        partial.insert(partial.end(), vp.array().begin(), vp.array().end());
        full.insert(full.end(), vf.array().begin(), vf.array().end());
    }
};
} // namespace detail
} // namespace mio

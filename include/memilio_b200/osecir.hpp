// osecir.hpp -- mio::osecir::{InfectionState, Parameters, Model, Simulation, simulate} with the reference's names
// (cpp/models/ode_secir/infection_state.h:33-47, parameters.h:47-460, model.h:61-80,258-411) as plain containers that
// pack into the flat parameter sets / contact tables of include/mio_cuda.h. The right-hand side itself only exists on
// the GPU (csrc/integrate.cu).
#pragma once

#include "mio.hpp"

namespace mio
{

// Populations<FP, AgeGroup, InfectionState> (epidemiology/populations.h): flat index age * Count + state
template <class FP, class State>
class Populations
{
public:
    struct Index {
        AgeGroup age;
        State state;
    };
    explicit Populations(size_t num_groups = 1)
        : m_G(num_groups)
        , m_v(num_groups * size_t(State::Count))
    {
    }
    size_t get_flat_index(Index i) const
    {
        return i.age.value * size_t(State::Count) + size_t(i.state);
    }
    UncertainValue<FP>& operator[](Index i)
    {
        return m_v[get_flat_index(i)];
    }
    const UncertainValue<FP>& operator[](Index i) const
    {
        return m_v[get_flat_index(i)];
    }
    // populations.h:108-112: the entry of index i in a flat vector of compartment values
    FP get_from(const Eigen::VectorX<FP>& y, Index i) const
    {
        return y[Eigen::Index(get_flat_index(i))];
    }
    size_t numel() const
    {
        return m_v.size();
    }
    size_t get_num_compartments() const
    {
        return m_v.size();
    }
    std::vector<UncertainValue<FP>>& array()
    {
        return m_v;
    }
    const std::vector<UncertainValue<FP>>& array() const
    {
        return m_v;
    }
    Eigen::VectorX<FP> get_compartments() const
    {
        Eigen::VectorX<FP> v(Eigen::Index(m_v.size()));
        for (size_t i = 0; i < m_v.size(); ++i) {
            v[Eigen::Index(i)] = m_v[i].value();
        }
        return v;
    }
    // populations.h:136-142
    template <class T = AgeGroup>
    FP get_group_total(AgeGroup g) const
    {
        FP acc = 0.0;
        for (size_t c = 0; c < size_t(State::Count); ++c) {
            acc = acc + m_v[g.value * size_t(State::Count) + c].value();
        }
        return acc;
    }
    FP get_total() const
    {
        FP acc = 0.0;
        for (auto& v : m_v) {
            acc += v.value();
        }
        return acc;
    }
    // populations.h:194-205
    template <class T = AgeGroup>
    void set_difference_from_group_total(Index midx, FP total_group_population)
    {
        FP current = get_group_total(midx.age);
        size_t idx = get_flat_index(midx);
        current -= m_v[idx].value();
        m_v[idx] = total_group_population - current;
    }
    // populations.h:240-249
    void set_difference_from_total(Index midx, FP total_population)
    {
        FP current = get_total();
        size_t idx = get_flat_index(midx);
        current -= m_v[idx].value();
        m_v[idx] = total_population - current;
    }
    // populations.h:216-232
    void set_total(FP value)
    {
        FP current = get_total();
        if (std::fabs(current) < 1e-12) {
            for (auto& v : m_v) {
                v = value / FP(m_v.size());
            }
        }
        else {
            for (auto& v : m_v) {
                v = v.value() * (value / current);
            }
        }
    }
    bool apply_constraints()
    {
        bool corrected = false;
        for (auto& v : m_v) {
            if (v.value() < 0.0) {
                v         = 0.0;
                corrected = true;
            }
        }
        return corrected;
    }

private:
    size_t m_G;
    std::vector<UncertainValue<FP>> m_v;
};

namespace osecir
{

enum class InfectionState
{
    Susceptible,
    Exposed,
    InfectedNoSymptoms,
    InfectedNoSymptomsConfirmed,
    InfectedSymptoms,
    InfectedSymptomsConfirmed,
    InfectedSevere,
    InfectedCritical,
    Recovered,
    Dead,
    Count
};

// parameter tags, ode_secir/parameters.h:47-397. SLOT: >= 0 index into the age record (MIO_PA_*), otherwise scalar.
#define MIO_B200_SCALAR_TAG(NAME, DEFAULT)                                                                             \
    template <class FP = ScalarType>                                                                                   \
    struct NAME {                                                                                                      \
        using Type = UncertainValue<FP>;                                                                               \
        static Type get_default(AgeGroup)                                                                              \
        {                                                                                                              \
            return Type(DEFAULT);                                                                                      \
        }                                                                                                              \
    };
#define MIO_B200_AGE_TAG(NAME, DEFAULT)                                                                                \
    template <class FP = ScalarType>                                                                                   \
    struct NAME {                                                                                                      \
        using Type = AgeArray<UncertainValue<FP>>;                                                                     \
        static Type get_default(AgeGroup n)                                                                            \
        {                                                                                                              \
            return Type(n.value, UncertainValue<FP>(DEFAULT));                                                         \
        }                                                                                                              \
    };
MIO_B200_SCALAR_TAG(StartDay, 0.0)
MIO_B200_SCALAR_TAG(Seasonality, 0.0)
MIO_B200_SCALAR_TAG(ICUCapacity, std::numeric_limits<FP>::max())
MIO_B200_SCALAR_TAG(TestAndTraceCapacity, std::numeric_limits<FP>::max())
MIO_B200_SCALAR_TAG(TestAndTraceCapacityMaxRisk, 5.0)
MIO_B200_AGE_TAG(TimeExposed, 1.0)
MIO_B200_AGE_TAG(TimeInfectedNoSymptoms, 1.0)
MIO_B200_AGE_TAG(TimeInfectedSymptoms, 1.0)
MIO_B200_AGE_TAG(TimeInfectedSevere, 1.0)
MIO_B200_AGE_TAG(TimeInfectedCritical, 1.0)
MIO_B200_AGE_TAG(TransmissionProbabilityOnContact, 1.0)
MIO_B200_AGE_TAG(RelativeTransmissionNoSymptoms, 1.0)
MIO_B200_AGE_TAG(RecoveredPerInfectedNoSymptoms, 0.0)
MIO_B200_AGE_TAG(RiskOfInfectionFromSymptomatic, 1.0)
MIO_B200_AGE_TAG(MaxRiskOfInfectionFromSymptomatic, 1.0)
MIO_B200_AGE_TAG(SeverePerInfectedSymptoms, 0.0)
MIO_B200_AGE_TAG(CriticalPerSevere, 0.0)
MIO_B200_AGE_TAG(DeathsPerSevere, 0.0)
MIO_B200_AGE_TAG(DeathsPerCritical, 0.0)
#undef MIO_B200_SCALAR_TAG
#undef MIO_B200_AGE_TAG
template <class FP = ScalarType>
struct ContactPatterns {
    using Type = UncertainContactMatrix<FP>;
    static Type get_default(AgeGroup n)
    {
        return Type(1, n.value);
    }
};

// DynamicNPIsInfectedSymptoms, parameters.h:395-410: NPIs that follow the relative number of symptomatic infections
template <class FP = ScalarType>
struct DynamicNPIsInfectedSymptoms {
    using Type = DynamicNPIs<FP>;
    static Type get_default(AgeGroup)
    {
        return Type();
    }
};

template <class FP = ScalarType>
class Parameters
{
    template <class Tag>
    struct Slot {
        typename Tag::Type value;
    };
    using Tuple =
        std::tuple<Slot<StartDay<FP>>, Slot<Seasonality<FP>>, Slot<ICUCapacity<FP>>, Slot<TestAndTraceCapacity<FP>>,
                   Slot<TestAndTraceCapacityMaxRisk<FP>>, Slot<ContactPatterns<FP>>, Slot<TimeExposed<FP>>,
                   Slot<TimeInfectedNoSymptoms<FP>>, Slot<TimeInfectedSymptoms<FP>>, Slot<TimeInfectedSevere<FP>>,
                   Slot<TimeInfectedCritical<FP>>, Slot<TransmissionProbabilityOnContact<FP>>,
                   Slot<RelativeTransmissionNoSymptoms<FP>>, Slot<RecoveredPerInfectedNoSymptoms<FP>>,
                   Slot<RiskOfInfectionFromSymptomatic<FP>>, Slot<MaxRiskOfInfectionFromSymptomatic<FP>>,
                   Slot<SeverePerInfectedSymptoms<FP>>, Slot<CriticalPerSevere<FP>>, Slot<DeathsPerSevere<FP>>,
                   Slot<DeathsPerCritical<FP>>, Slot<DynamicNPIsInfectedSymptoms<FP>>>;

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
    // commuter testing, parameters.h:417-460
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

    // flat image (MIO_PH_* / MIO_PA_* of mio_cuda.h): values, and the same slots as distributions
    void pack_values(double* pset, double* local) const
    {
        pack([&](int slot, const UncertainValue<FP>& v) { pset[slot] = v.value(); },
             [&](int slot, const UncertainValue<FP>& v) { local[slot] = v.value(); });
    }
    void pack_distributions(mio_uncertain_value* pset, mio_uncertain_value* local) const
    {
        pack([&](int slot, const UncertainValue<FP>& v) { pset[slot] = v.to_pod(); },
             [&](int slot, const UncertainValue<FP>& v) { local[slot] = v.to_pod(); });
    }
    void unpack_values(const double* pset, const double* local)
    {
        const_cast<const Parameters*>(this)->pack(
            [&](int slot, const UncertainValue<FP>& v) { const_cast<UncertainValue<FP>&>(v) = pset[slot]; },
            [&](int slot, const UncertainValue<FP>& v) { const_cast<UncertainValue<FP>&>(v) = local[slot]; });
        m_commuter_nondetection    = pset[MIO_PH_COMMUTER_NONDETECTION];
        m_start_commuter_detection = pset[MIO_PH_START_COMMUTER_DETECTION];
        m_end_commuter_detection   = pset[MIO_PH_END_COMMUTER_DETECTION];
    }

    // parameters.h:465-615, on the flat image so that the host code and the sampler share one implementation
    bool apply_constraints()
    {
        const size_t G = m_num_groups.value;
        std::vector<double> p(MIO_PSET_HEADER + G * MIO_OSECIR_AGE_STRIDE, 0.0), before;
        double local[MIO_NODE_LOCAL];
        pack_values(p.data(), local);
        before                     = p;
        const double l0 = local[0], l1 = local[1];
        auto clamp0 = [](double& v, bool bad) {
            if (bad) {
                v = 0.0;
            }
        };
        clamp0(p[MIO_PH_SEASONALITY], p[MIO_PH_SEASONALITY] < 0.0 || p[MIO_PH_SEASONALITY] > 0.5);
        clamp0(local[0], local[0] < 0.0);
        clamp0(local[1], local[1] < 0.0);
        clamp0(p[MIO_PH_TNT_MAX_RISK], p[MIO_PH_TNT_MAX_RISK] < 0.0);
        for (size_t i = 0; i < G; ++i) {
            double* a = p.data() + MIO_PSET_HEADER + i * MIO_OSECIR_AGE_STRIDE;
            for (int s : {MIO_PA_TIME_EXPOSED, MIO_PA_TIME_INS, MIO_PA_TIME_ISY, MIO_PA_TIME_ISEV, MIO_PA_TIME_ICR}) {
                if (a[s] < 1e-1) {
                    a[s] = 1e-1;
                }
            }
            clamp0(a[MIO_PA_TRANS_PROB], a[MIO_PA_TRANS_PROB] < 0.0 || a[MIO_PA_TRANS_PROB] > 1.0);
            clamp0(a[MIO_PA_REL_TRANS_NOSYMP], a[MIO_PA_REL_TRANS_NOSYMP] < 0.0);
            for (int s : {MIO_PA_REC_PER_INS, MIO_PA_RISK_SYMP, MIO_PA_SEV_PER_ISY, MIO_PA_CRIT_PER_SEV,
                          MIO_PA_DEATHS_PER_SEV}) {
                clamp0(a[s], a[s] < 0.0 || a[s] > 1.0);
            }
            clamp0(a[MIO_PA_DEATHS_PER_SEV], a[MIO_PA_CRIT_PER_SEV] + a[MIO_PA_DEATHS_PER_SEV] > 1.0);
            clamp0(a[MIO_PA_DEATHS_PER_CRIT], a[MIO_PA_DEATHS_PER_CRIT] < 0.0 || a[MIO_PA_DEATHS_PER_CRIT] > 1.0);
        }
        const bool corrected = p != before || l0 != local[0] || l1 != local[1];
        if (corrected) {
            unpack_values(p.data(), local);
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
    template <class FS, class FL>
    void pack(FS&& scalar, FL&& local) const
    {
        scalar(MIO_PH_START_DAY, get<StartDay<FP>>());
        scalar(MIO_PH_SEASONALITY, get<Seasonality<FP>>());
        scalar(MIO_PH_TNT_MAX_RISK, get<TestAndTraceCapacityMaxRisk<FP>>());
        scalar(MIO_PH_COMMUTER_NONDETECTION, UncertainValue<FP>(m_commuter_nondetection));
        scalar(MIO_PH_START_COMMUTER_DETECTION, UncertainValue<FP>(m_start_commuter_detection));
        scalar(MIO_PH_END_COMMUTER_DETECTION, UncertainValue<FP>(m_end_commuter_detection));
        local(MIO_NL_ICU_CAPACITY, get<ICUCapacity<FP>>());
        local(MIO_NL_TNT_CAPACITY, get<TestAndTraceCapacity<FP>>());
        for (size_t i = 0; i < m_num_groups.value; ++i) {
            const int o = int(MIO_PSET_HEADER + i * MIO_OSECIR_AGE_STRIDE);
            const AgeGroup a(i);
            scalar(o + MIO_PA_TIME_EXPOSED, get<TimeExposed<FP>>()[a]);
            scalar(o + MIO_PA_TIME_INS, get<TimeInfectedNoSymptoms<FP>>()[a]);
            scalar(o + MIO_PA_TIME_ISY, get<TimeInfectedSymptoms<FP>>()[a]);
            scalar(o + MIO_PA_TIME_ISEV, get<TimeInfectedSevere<FP>>()[a]);
            scalar(o + MIO_PA_TIME_ICR, get<TimeInfectedCritical<FP>>()[a]);
            scalar(o + MIO_PA_TRANS_PROB, get<TransmissionProbabilityOnContact<FP>>()[a]);
            scalar(o + MIO_PA_REL_TRANS_NOSYMP, get<RelativeTransmissionNoSymptoms<FP>>()[a]);
            scalar(o + MIO_PA_REC_PER_INS, get<RecoveredPerInfectedNoSymptoms<FP>>()[a]);
            scalar(o + MIO_PA_RISK_SYMP, get<RiskOfInfectionFromSymptomatic<FP>>()[a]);
            scalar(o + MIO_PA_MAX_RISK_SYMP, get<MaxRiskOfInfectionFromSymptomatic<FP>>()[a]);
            scalar(o + MIO_PA_SEV_PER_ISY, get<SeverePerInfectedSymptoms<FP>>()[a]);
            scalar(o + MIO_PA_CRIT_PER_SEV, get<CriticalPerSevere<FP>>()[a]);
            scalar(o + MIO_PA_DEATHS_PER_SEV, get<DeathsPerSevere<FP>>()[a]);
            scalar(o + MIO_PA_DEATHS_PER_CRIT, get<DeathsPerCritical<FP>>()[a]);
        }
    }

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
    // FlowModel::get_flat_flow_index (compartments/flow_model.h:181-192): position of the flow Source -> Target of an
    // age group in the flow vectors of a FlowSimulation; the flows of one age group are contiguous
    template <InfectionState Source, InfectionState Target>
    size_t get_flat_flow_index(AgeGroup age) const
    {
        constexpr int k = flow_position(Source, Target);
        static_assert(k >= 0, "no such flow in the ODE-SECIR model");
        return size_t(age) * size_t(num_flows_per_group) + size_t(k);
    }
    Eigen::VectorX<FP> get_initial_flows() const // flow_model.h:146-149
    {
        Eigen::VectorX<FP> f(Eigen::Index(size_t(parameters.get_num_groups().value) * num_flows_per_group));
        for (Eigen::Index i = 0; i < f.size(); ++i) {
            f[i] = 0;
        }
        return f;
    }
    static constexpr int num_flows_per_group = 15;
    // the model's Flows type list (model.h:45-59) as (source, target) pairs, in its order
    static constexpr int flow_position(InfectionState s, InfectionState t)
    {
        using I = InfectionState;
        constexpr I list[num_flows_per_group][2] = {
            {I::Susceptible, I::Exposed},
            {I::Exposed, I::InfectedNoSymptoms},
            {I::InfectedNoSymptoms, I::InfectedSymptoms},
            {I::InfectedNoSymptoms, I::Recovered},
            {I::InfectedNoSymptomsConfirmed, I::InfectedSymptomsConfirmed},
            {I::InfectedNoSymptomsConfirmed, I::Recovered},
            {I::InfectedSymptoms, I::InfectedSevere},
            {I::InfectedSymptoms, I::Recovered},
            {I::InfectedSymptomsConfirmed, I::InfectedSevere},
            {I::InfectedSymptomsConfirmed, I::Recovered},
            {I::InfectedSevere, I::InfectedCritical},
            {I::InfectedSevere, I::Recovered},
            {I::InfectedSevere, I::Dead},
            {I::InfectedCritical, I::Dead},
            {I::InfectedCritical, I::Recovered}};
        for (int k = 0; k < num_flows_per_group; ++k) {
            if (list[k][0] == s && list[k][1] == t) {
                return k;
            }
        }
        return -1;
    }
    Populations populations;
    ParameterSet parameters;
};

namespace detail
{
constexpr int NC = int(InfectionState::Count);

// evaluated contact table of a ContactMatrixGroup (layout of mio_cuda.h), change points via the library's
// Dampings::update_cache restatement
template <class FP>
std::vector<double> contact_table(const ContactMatrixGroup<FP>& cm, int G, int max_loc, int max_pts)
{
    const size_t csz = size_t(mio_ctable_size(G, max_loc, max_pts));
    std::vector<double> ct(csz, 0.0);
    const int L = int(cm.get_num_matrices());
    if (L > max_loc) {
        throw std::runtime_error("contact table: too many contact locations");
    }
    ct[0]                = L;
    const int64_t stride = max_pts + int64_t(2 + max_pts) * G * G;
    for (int l = 0; l < L; ++l) {
        const auto& m = cm[size_t(l)];
        const int K   = int(m.get_dampings().size());
        const size_t Ku = size_t(K), K1 = size_t(std::max(K, 1)), GG = size_t(G) * size_t(G);
        std::vector<double> times(Ku), mats(Ku * GG), ot(K1), om(K1 * GG);
        std::vector<int32_t> levels(Ku), types(Ku);
        for (int k = 0; k < K; ++k) {
            const auto& d    = m.get_dampings()[size_t(k)];
            times[size_t(k)] = d.time;
            levels[size_t(k)] = d.level;
            types[size_t(k)]  = d.type;
            std::copy(d.m.data(), d.m.data() + G * G, mats.begin() + size_t(k) * G * G);
        }
        int32_t nout = 0;
        cuda_check(mio_damping_cache_build(G, G, K, times.data(), levels.data(), types.data(), mats.data(), &nout,
                                           ot.data(), om.data()));
        if (nout > max_pts) {
            throw std::runtime_error("contact table: too many damping change points");
        }
        ct[size_t(1 + l)] = nout;
        double* p         = ct.data() + 1 + max_loc + l * stride;
        std::copy(ot.begin(), ot.begin() + nout, p);
        std::copy(m.get_baseline().data(), m.get_baseline().data() + G * G, p + max_pts);
        std::copy(m.get_minimum().data(), m.get_minimum().data() + G * G, p + max_pts + G * G);
        std::copy(om.begin(), om.begin() + size_t(nout) * G * G, p + max_pts + 2 * G * G);
    }
    return ct;
}

template <class FP>
int num_change_points(const ContactMatrixGroup<FP>& cm)
{
    size_t k = 0;
    for (size_t l = 0; l < cm.get_num_matrices(); ++l) {
        k = std::max(k, cm[l].get_dampings().size());
    }
    return int(k);
}
} // namespace detail

} // namespace osecir

// mio::Simulation / mio::FlowSimulation (compartments/simulation.h:36-190, flow_simulation.h:37-97): defined for
// osecir::Model below; as BaseT of osecir::Simulation they select what a node integrates
template <class FP, class M>
class Simulation;
template <class FP, class M>
class FlowSimulation;

namespace osecir
{

// osecir::Simulation (model.h:258-367): one node. advance() runs the node on the GPU as a 1-node graph and appends
// every accepted integrator step to the result, like SystemIntegrator::advance does (integrator.h:224).
// BaseT = mio::FlowSimulation<FP, Model<FP>>: the integrands are the cumulative flows (flow_simulation.h:63-92), the
// result is derived from them, get_flows() is their series.
template <class FP = ScalarType, class BaseT = mio::Simulation<FP, osecir::Model<FP>>>
class Simulation
{
public:
    using Model = osecir::Model<FP>;
    static constexpr bool is_flow_simulation = std::is_same_v<BaseT, mio::FlowSimulation<FP, osecir::Model<FP>>>;
    Simulation(const Model& model, FP t0 = 0., FP dt = 0.1)
        : m_model(model)
        , m_t0(t0)
        , m_dt(dt)
        , m_result(t0, model.get_initial_values())
        , m_flow_result(t0, model.get_initial_flows())
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
    // FlowSimulation::get_flows, flow_simulation.h:104-119
    TimeSeries<FP>& get_flows()
    {
        static_assert(is_flow_simulation, "get_flows() needs BaseT = mio::FlowSimulation");
        return m_flow_result;
    }
    const TimeSeries<FP>& get_flows() const
    {
        static_assert(is_flow_simulation, "get_flows() needs BaseT = mio::FlowSimulation");
        return m_flow_result;
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
    // Simulation::advance (model.h:289-367): without dynamic NPIs one integration to tmax; with them chunks of at most
    // a day (ending at the NPI's end), the threshold is looked at before every chunk
    Eigen::VectorX<FP> advance(FP tmax)
    {
        if (m_model.parameters.template get<DynamicNPIsInfectedSymptoms<FP>>().get_thresholds().empty()) {
            return advance_chunk(tmax);
        }
        FP t = m_result.get_last_time();
        while (t < tmax) {
            check_dynamic_npis(t);
            FP dt_eff = std::min<FP>(1.0, tmax - t);
            if (t < m_npi_end) {
                dt_eff = std::min<FP>(dt_eff, m_npi_end - t);
            }
            advance_chunk(t + dt_eff);
            t = t + dt_eff;
        }
        return m_result.get_last_value();
    }
    // The decision taken before a chunk that starts at t (model.h:303-353): is the biggest exceeded threshold stronger
    // than the NPI in force, or has that one expired? Then its dampings enter the contact patterns from t + delay
    // (no delay at t <= 0 and for a renewal at expiry) for `duration` days, inside the directive window; the damping
    // times are one day later than the NPI times so that the smoothing ramp lies in the future. Returns whether the
    // contact patterns changed.
    bool check_dynamic_npis(FP t)
    {
        auto& npis = m_model.parameters.template get<DynamicNPIsInfectedSymptoms<FP>>();
        if (npis.get_thresholds().empty()) {
            return false;
        }
        const FP window_begin = npis.get_directive_begin().get(), window_end = npis.get_directive_end().get();
        auto at_least = [](FP a, FP b) { // floating_point_greater_equal(a, b, 1e-10)
            return !(a - b < -(1e-10 + std::numeric_limits<FP>::min() * std::max(std::abs(a), std::abs(b))));
        };
        if (!at_least(t, window_begin) || !(t < window_end)) {
            return false;
        }
        const FP level = get_infections_relative(*this, t, m_result.get_last_value()) * npis.get_base_value();
        auto hit       = npis.get_max_exceeded_threshold(level);
        const bool expired = at_least(t, m_npi_end);
        if (hit == npis.get_thresholds().end() || !(hit->first > m_npi_threshold || expired)) {
            return false;
        }
        const bool renewal = expired && !(hit->first > m_npi_threshold);
        const FP delay     = (renewal || !(t > 0)) ? FP(0) : npis.get_implementation_delay().get();
        const FP start     = t + delay;
        if (!(start < window_end)) {
            return false;
        }
        const FP end    = std::min<FP>(window_end, start + npis.get_duration().get());
        m_npi_threshold = hit->first;
        m_npi_end       = end;
        const FP damp_from = (!renewal && start > 0) ? start + 1 : start;
        implement_dynamic_npis<FP>(m_model.parameters.template get<ContactPatterns<FP>>().get_cont_freq_mat(),
                                   hit->second, SimulationTime<FP>(damp_from), SimulationTime<FP>(end + 1),
                                   [](const Eigen::VectorX<FP>& g) {
                                       return make_contact_damping_matrix<FP>(g);
                                   });
        return true;
    }
    // end of the dynamic NPI in force (a chunk must not step over it, model.h:355-361)
    FP get_dynamic_npi_end() const
    {
        return m_npi_end;
    }
    Eigen::VectorX<FP> advance_chunk(FP tmax)
    {
        const int G   = int(m_model.parameters.get_num_groups().value);
        const int n   = G * detail::NC;
        const int nf  = G * Model::num_flows_per_group;
        const auto& cm = static_cast<const ContactMatrixGroup<FP>&>(
            m_model.parameters.template get<ContactPatterns<FP>>());
        const int L = int(cm.get_num_matrices()), K = std::max(1, detail::num_change_points(cm));
        mio_graph_desc d{};
        d.model           = MIO_MODEL_OSECIR;
        d.sim_kind        = MIO_SIM_MODEL;
        d.num_groups      = G;
        d.num_nodes       = 1;
        d.num_psets       = 1;
        d.num_ctables     = 1;
        d.ct_max_loc      = L;
        d.ct_max_pts      = K;
        d.max_runs        = 1;
        d.traj_capacity   = 1 << 16;
        d.integrator      = m_core->desc();
        d.t0              = m_result.get_last_time();
        d.dt_graph        = std::numeric_limits<double>::infinity();
        d.flow_simulation = is_flow_simulation ? 1 : 0;
        const size_t psz = size_t(mio_pset_size(MIO_MODEL_OSECIR, G));
        std::vector<double> pset(psz, 0.0), local(MIO_NODE_LOCAL);
        m_model.parameters.pack_values(pset.data(), local.data());
        std::vector<double> ct = detail::contact_table(cm, G, L, K);
        std::vector<double> y0(m_result.get_last_value().data(), m_result.get_last_value().data() + n);
        mio_run_data rd{pset.data(), ct.data(), local.data(), y0.data(), &m_dt, nullptr};
        mio_handle* h = nullptr;
        cuda_check(mio_cuda_graph_create(&d, 0, &h));
        try {
            cuda_check(mio_cuda_set_runs(h, 0, 1, &rd));
            if constexpr (is_flow_simulation) {
                // the integration continues from the last flow time point (error control acts on the flows)
                cuda_check(mio_cuda_set_flows(h, 0, 1, m_flow_result.get_last_value().data()));
            }
            cuda_check(mio_cuda_advance(h, tmax));
            int32_t npts = 0;
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
            if constexpr (is_flow_simulation) {
                std::vector<double> fv(cap * size_t(nf));
                cuda_check(mio_cuda_get_flow_trajectory(h, 0, 1, fv.data()));
                for (int k = 1; k < npts; ++k) {
                    Eigen::VectorX<FP> v(nf);
                    std::copy(fv.begin() + size_t(k) * nf, fv.begin() + size_t(k + 1) * nf, v.data());
                    m_flow_result.add_time_point(times[size_t(k)], v);
                }
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
    TimeSeries<FP> m_result, m_flow_result;
    std::shared_ptr<OdeIntegratorCore<FP>> m_core;
    FP m_npi_threshold = -std::numeric_limits<FP>::max(); // threshold and end of the dynamic NPI in force
    FP m_npi_end       = 0;
};

// get_infections_relative, model.h:413-423: symptomatic infections relative to the model's total population
template <class FP, class BaseT>
FP get_infections_relative(const Simulation<FP, BaseT>& sim, FP /*t*/, const Eigen::VectorX<FP>& y)
{
    FP sum_inf = 0;
    for (auto i = AgeGroup(0); i < sim.get_model().parameters.get_num_groups(); i++) {
        sum_inf += sim.get_model().populations.get_from(y, {i, InfectionState::InfectedSymptoms});
    }
    return sum_inf / sim.get_model().populations.get_total();
}

} // namespace osecir

// the generic simulation types for the ODE-SECIR model (the reference's osecir::Simulation only adds the dynamic NPI
// check to its base): mio::Simulation<FP, osecir::Model<FP>> and mio::FlowSimulation<FP, osecir::Model<FP>>
template <class FP>
class Simulation<FP, osecir::Model<FP>> : public osecir::Simulation<FP>
{
public:
    using osecir::Simulation<FP>::Simulation;
};
template <class FP>
class FlowSimulation<FP, osecir::Model<FP>>
    : public osecir::Simulation<FP, FlowSimulation<FP, osecir::Model<FP>>>
{
public:
    using osecir::Simulation<FP, FlowSimulation<FP, osecir::Model<FP>>>::Simulation;
};

// mio::simulate / mio::simulate_flows, compartments/simulation.h:223-236, flow_simulation.h:110-121
template <class FP, class Model, class Sim = Simulation<FP, Model>>
TimeSeries<FP> simulate(FP t0, FP tmax, FP dt, const Model& model,
                        std::unique_ptr<OdeIntegratorCore<FP>>&& integrator_core = nullptr)
{
    model.check_constraints();
    Sim sim(model, t0, dt);
    if (integrator_core) {
        sim.set_integrator_core(std::move(integrator_core));
    }
    sim.advance(tmax);
    return sim.get_result();
}
template <class FP, class Model, class Sim = FlowSimulation<FP, Model>>
std::vector<TimeSeries<FP>> simulate_flows(FP t0, FP tmax, FP dt, const Model& model,
                                           std::unique_ptr<OdeIntegratorCore<FP>>&& integrator_core = nullptr)
{
    model.check_constraints();
    Sim sim(model, t0, dt);
    if (integrator_core) {
        sim.set_integrator_core(std::move(integrator_core));
    }
    sim.advance(tmax);
    return {sim.get_result(), sim.get_flows()};
}

namespace osecir
{

// osecir::simulate / osecir::simulate_flows, model.h:383-411
template <class FP = ScalarType>
TimeSeries<FP> simulate(FP t0, FP tmax, FP dt, const Model<FP>& model,
                        std::unique_ptr<OdeIntegratorCore<FP>>&& integrator = nullptr)
{
    return mio::simulate<FP, Model<FP>, Simulation<FP>>(t0, tmax, dt, model, std::move(integrator));
}
template <class FP = ScalarType>
std::vector<TimeSeries<FP>> simulate_flows(FP t0, FP tmax, FP dt, const Model<FP>& model,
                                           std::unique_ptr<OdeIntegratorCore<FP>>&& integrator = nullptr)
{
    return mio::simulate_flows<FP, Model<FP>, Simulation<FP, mio::FlowSimulation<FP, Model<FP>>>>(
        t0, tmax, dt, model, std::move(integrator));
}

// draw_sample_infection / draw_sample_demographics / draw_sample(Model&), parameter_space.h:125-205
template <class FP>
void draw_sample_demographics(Model<FP>& model)
{
    model.parameters.template get<ICUCapacity<FP>>().draw_sample();
    model.parameters.template get<TestAndTraceCapacity<FP>>().draw_sample();
    for (auto i = AgeGroup(0); i < model.parameters.get_num_groups(); i++) {
        FP group_total = model.populations.get_group_total(i);
        model.populations[{i, InfectionState::Exposed}].draw_sample();
        model.populations[{i, InfectionState::InfectedNoSymptoms}].draw_sample();
        model.populations[{i, InfectionState::InfectedSymptoms}].draw_sample();
        model.populations[{i, InfectionState::InfectedSevere}].draw_sample();
        model.populations[{i, InfectionState::InfectedCritical}].draw_sample();
        model.populations[{i, InfectionState::Recovered}].draw_sample();
        model.populations.template set_difference_from_group_total<AgeGroup>({i, InfectionState::Susceptible},
                                                                             group_total);
        model.populations.template set_difference_from_group_total<AgeGroup>({i, InfectionState::Susceptible},
                                                                             model.populations.get_group_total(i));
    }
}

template <class FP>
void draw_sample_infection(Model<FP>& model)
{
    auto& p = model.parameters;
    p.template get<Seasonality<FP>>().draw_sample();
    p.template get<TimeExposed<FP>>()[AgeGroup(0)].draw_sample();
    p.template get<TimeInfectedNoSymptoms<FP>>()[AgeGroup(0)].draw_sample();
    p.template get<TimeInfectedSymptoms<FP>>()[AgeGroup(0)].draw_sample();
    p.template get<RelativeTransmissionNoSymptoms<FP>>()[AgeGroup(0)].draw_sample();
    p.template get<RiskOfInfectionFromSymptomatic<FP>>()[AgeGroup(0)].draw_sample();
    p.template get<MaxRiskOfInfectionFromSymptomatic<FP>>()[AgeGroup(0)].draw_sample();
    for (auto i = AgeGroup(0); i < p.get_num_groups(); i++) {
        p.template get<TimeExposed<FP>>()[i]            = p.template get<TimeExposed<FP>>()[AgeGroup(0)];
        p.template get<TimeInfectedNoSymptoms<FP>>()[i] = p.template get<TimeInfectedNoSymptoms<FP>>()[AgeGroup(0)];
        p.template get<RelativeTransmissionNoSymptoms<FP>>()[i] =
            p.template get<RelativeTransmissionNoSymptoms<FP>>()[AgeGroup(0)];
        p.template get<RiskOfInfectionFromSymptomatic<FP>>()[i] =
            p.template get<RiskOfInfectionFromSymptomatic<FP>>()[AgeGroup(0)];
        p.template get<MaxRiskOfInfectionFromSymptomatic<FP>>()[i] =
            p.template get<MaxRiskOfInfectionFromSymptomatic<FP>>()[AgeGroup(0)];
        p.template get<TimeInfectedSymptoms<FP>>()[i].draw_sample();
        p.template get<TimeInfectedSevere<FP>>()[i].draw_sample();
        p.template get<TimeInfectedCritical<FP>>()[i].draw_sample();
        p.template get<TransmissionProbabilityOnContact<FP>>()[i].draw_sample();
        p.template get<RecoveredPerInfectedNoSymptoms<FP>>()[i].draw_sample();
        p.template get<DeathsPerCritical<FP>>()[i].draw_sample();
        p.template get<SeverePerInfectedSymptoms<FP>>()[i].draw_sample();
        p.template get<CriticalPerSevere<FP>>()[i].draw_sample();
    }
}

template <class FP>
void draw_sample(Model<FP>& model)
{
    draw_sample_infection(model);
    draw_sample_demographics(model);
    model.parameters.template get<ContactPatterns<FP>>().draw_sample();
    model.apply_constraints();
}

} // namespace osecir
} // namespace mio

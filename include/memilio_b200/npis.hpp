// npis.hpp -- mio::DynamicNPIs and implement_dynamic_npis with the reference's names.
//
// Dynamic NPIs are control logic between integration chunks: a threshold on the relative number of symptomatic
// infections decides, at most once per day, whether dampings are added to a node's contact patterns
// (cpp/models/ode_secir/model.h:299-353). That decision is taken on the host (here); the integration chunks run on the
// GPU, and a changed contact pattern reaches the device as a new contact table (mio_cuda_set_contact_tables).
//   cpp/memilio/epidemiology/dynamic_npis.h:40-120   DynamicNPIs
//   cpp/memilio/epidemiology/dynamic_npis.h:270-306  get_damping_indices, get_active_damping
//   cpp/memilio/epidemiology/dynamic_npis.h:335-395  implement_dynamic_npis
//   cpp/memilio/epidemiology/damping_sampling.h:282-288 make_contact_damping_matrix
#pragma once

namespace mio
{

template <class FP = ScalarType>
class DynamicNPIs
{
public:
    using Threshold = std::pair<FP, std::vector<DampingSampling<FP>>>;
    // the thresholds stay sorted by value, biggest first; setting an existing value replaces its NPIs
    void set_threshold(FP threshold, const std::vector<DampingSampling<FP>>& dampings)
    {
        auto pos = m_thresholds.begin();
        while (pos != m_thresholds.end() && pos->first > threshold) {
            ++pos;
        }
        if (pos != m_thresholds.end() && pos->first == threshold) {
            pos->second = dampings;
        }
        else {
            m_thresholds.insert(pos, Threshold(threshold, dampings));
        }
    }
    // iterator to the biggest threshold below `value` (end() if there is none)
    auto get_max_exceeded_threshold(FP value)
    {
        auto it = m_thresholds.begin();
        while (it != m_thresholds.end() && !(value > it->first)) {
            ++it;
        }
        return it;
    }
    std::vector<Threshold>& get_thresholds()
    {
        return m_thresholds;
    }
    const std::vector<Threshold>& get_thresholds() const
    {
        return m_thresholds;
    }
    SimulationTime<FP> get_duration() const
    {
        return m_duration;
    }
    void set_duration(SimulationTime<FP> v)
    {
        m_duration = v;
    }
    SimulationTime<FP> get_implementation_delay() const
    {
        return m_delay;
    }
    void set_implementation_delay(SimulationTime<FP> v)
    {
        m_delay = v;
    }
    FP get_base_value() const
    {
        return m_base;
    }
    void set_base_value(FP v)
    {
        m_base = v;
    }
    SimulationTime<FP> get_directive_begin() const
    {
        return m_directive_begin;
    }
    void set_directive_begin(SimulationTime<FP> v)
    {
        m_directive_begin = v;
    }
    SimulationTime<FP> get_directive_end() const
    {
        return m_directive_end;
    }
    void set_directive_end(SimulationTime<FP> v)
    {
        m_directive_end = v;
    }
    void draw_sample()
    {
        for (auto& t : m_thresholds) {
            for (auto& d : t.second) {
                d.draw_sample();
            }
        }
    }

private:
    std::vector<Threshold> m_thresholds;
    SimulationTime<FP> m_duration{14.0}, m_delay{0.0};
    SimulationTime<FP> m_directive_begin{std::numeric_limits<FP>::lowest()};
    SimulationTime<FP> m_directive_end{std::numeric_limits<FP>::max()};
    FP m_base{1.0};
};

// 1 - sqrt((1 - g)(1 - g)^T): the damping matrix of group-wise factors g
template <class FP>
Eigen::MatrixX<FP> make_contact_damping_matrix(const Eigen::VectorX<FP>& g)
{
    const Eigen::Index n = g.size();
    Eigen::MatrixX<FP> m(n, n);
    for (Eigen::Index i = 0; i < n; ++i) {
        for (Eigen::Index j = 0; j < n; ++j) {
            m(i, j) = 1.0 - std::sqrt((1.0 - g[i]) * (1.0 - g[j]));
        }
    }
    return m;
}

// indices of the dampings of (level, type) strictly inside (begin, end)
template <class FP>
std::vector<size_t> get_damping_indices(const ContactMatrix<FP>& expr, DampingLevel lvl, DampingType type,
                                        SimulationTime<FP> begin, SimulationTime<FP> end)
{
    std::vector<size_t> hits;
    const auto& list = expr.get_dampings();
    for (size_t i = 0; i < list.size(); ++i) {
        if (list[i].level == lvl.value && list[i].type == type.value && list[i].time > begin.get() &&
            list[i].time < end.get()) {
            hits.push_back(i);
        }
    }
    return hits;
}

// coefficients of the latest damping of (level, type) at or before t; zero if there is none
template <class FP>
Eigen::MatrixX<FP> get_active_damping(const ContactMatrix<FP>& expr, DampingLevel lvl, DampingType type,
                                      SimulationTime<FP> t)
{
    const auto& list = expr.get_dampings();
    for (size_t i = list.size(); i-- > 0;) {
        if (list[i].level == lvl.value && list[i].type == type.value && list[i].time <= t.get()) {
            return list[i].m;
        }
    }
    return Eigen::MatrixX<FP>::Zero(expr.get_baseline().rows(), expr.get_baseline().cols());
}

namespace detail
{
template <class FP>
Eigen::MatrixX<FP> elementwise_max(const Eigen::MatrixX<FP>& a, const Eigen::MatrixX<FP>& b)
{
    Eigen::MatrixX<FP> r = a;
    for (Eigen::Index k = 0; k < r.size(); ++k) {
        r.data()[k] = std::max(a.data()[k], b.data()[k]);
    }
    return r;
}
} // namespace detail

// An NPI holds its (level, type) slot of the listed matrices at >= its value during [begin, end): the slot is raised at
// `begin`, every damping of the slot that starts inside the window is raised too, and at `end` the slot returns to what
// would have been active then. Afterwards entries that repeat the value of their predecessor in the slot are dropped.
template <class FP, class MakeMatrix>
void implement_dynamic_npis(ContactMatrixGroup<FP>& group, const std::vector<DampingSampling<FP>>& npis,
                            SimulationTime<FP> begin, SimulationTime<FP> end, MakeMatrix&& make_matrix)
{
    for (const auto& npi : npis) {
        const DampingLevel level = npi.get_level();
        const DampingType type   = npi.get_type();
        Eigen::VectorX<FP> weighted = npi.get_group_weights();
        for (Eigen::Index k = 0; k < weighted.size(); ++k) {
            weighted[k] = npi.get_value().value() * weighted[k];
        }
        const Eigen::MatrixX<FP> value = make_matrix(weighted);
        for (size_t mat : npi.get_matrix_indices()) {
            ContactMatrix<FP>& expr              = group[mat];
            const Eigen::MatrixX<FP> at_begin    = get_active_damping(expr, level, type, begin);
            const Eigen::MatrixX<FP> restored    = get_active_damping(expr, level, type, end);
            bool changed                         = false;
            bool raises                          = false;
            for (Eigen::Index k = 0; k < value.size(); ++k) {
                raises = raises || at_begin.data()[k] < value.data()[k];
            }
            if (raises) {
                expr.add_damping(detail::elementwise_max(value, at_begin), level, type, begin);
                changed = true;
            }
            for (size_t i : get_damping_indices(expr, level, type, begin, end)) {
                const auto inside = expr.get_dampings()[i]; // copy: add_damping replaces the entry
                expr.add_damping(detail::elementwise_max(inside.m, value), level, type, SimulationTime<FP>(inside.time));
                changed = true;
            }
            if (changed) {
                expr.add_damping(restored, level, type, end);
            }
        }
    }
    for (auto& expr : group) {
        // from the back; an entry is obsolete if the previous entry of its slot carries the same coefficients
        for (size_t back = 0; back + 1 < expr.get_dampings().size(); ++back) {
            const auto& list = expr.get_dampings();
            const size_t i   = list.size() - 1 - back;
            size_t prev      = i;
            while (prev-- > 0) {
                if (list[prev].level == list[i].level && list[prev].type == list[i].type) {
                    break;
                }
            }
            if (prev != size_t(-1) && list[prev].m == list[i].m) {
                expr.remove_damping(i);
            }
        }
    }
}

} // namespace mio

// mio_oracle_npis.hpp -- CPU restatement of MEmilio's dynamic NPIs. TEST INFRASTRUCTURE, not product
// code: only tests/, __graft_entry__.smoke() and bench.py's CPU arm may use anything under oracle/.
//
// Follows, literally:
//   cpp/memilio/epidemiology/dynamic_npis.h:40-120   DynamicNPIs (thresholds sorted descending, duration, delay, base
//                                                    value, directive window)
//   cpp/memilio/epidemiology/dynamic_npis.h:270-306  get_damping_indices, get_active_damping
//   cpp/memilio/epidemiology/dynamic_npis.h:335-395  implement_dynamic_npis
//   cpp/memilio/epidemiology/damping_sampling.h:282-306 make_contact_damping_matrix, make_mobility_damping_vector
//   cpp/models/ode_secir/model.h:289-367             osecir::Simulation::advance with the dynamic NPI check
//   cpp/models/ode_secir/model.h:413-423             get_infections_relative
// Pinned by the exact cases of cpp/tests/test_dynamic_npis.cpp (tests/test_oracle_npis.py).
#pragma once

#include "mio_oracle.hpp"

namespace orc
{

// math/floating_point.h:115-139
inline bool floating_point_less(double v1, double v2, double abs_tol = 0, double rel_tol = DBL_MIN)
{
    const double diff = v1 - v2;
    return diff < -(abs_tol + rel_tol * std::max(std::abs(v1), std::abs(v2)));
}
inline bool floating_point_greater_equal(double v1, double v2, double abs_tol = 0, double rel_tol = DBL_MIN)
{
    return !floating_point_less(v1, v2, abs_tol, rel_tol);
}

// epidemiology/damping_sampling.h:40-250, the members dynamic NPIs read
struct DampingSamplingO {
    double value = 0.0;
    int level = 0, type = 0;
    double time = 0.0;
    std::vector<int> matrix_indices;
    Vec group_weights;
};

// damping_sampling.h:282-288: 1 - sqrt((1 - g) (1 - g)^T)
inline Mat make_contact_damping_matrix(const Vec& g)
{
    const int n = int(g.size());
    Mat m(n, n);
    for (int i = 0; i < n; ++i) {
        for (int j = 0; j < n; ++j) {
            m(i, j) = 1.0 - std::sqrt((1.0 - g[size_t(i)]) * (1.0 - g[size_t(j)]));
        }
    }
    return m;
}
// damping_sampling.h:298-306: the group's value for every compartment of the group
inline Mat make_mobility_damping_vector(int size, const Vec& g)
{
    Mat m(size, 1);
    const size_t num_compartments = size_t(size) / g.size();
    for (int i = 0; i < size; ++i) {
        m(i, 0) = g[size_t(i) / num_compartments];
    }
    return m;
}

struct DynamicNPIs {
    std::vector<std::pair<double, std::vector<DampingSamplingO>>> thresholds; // sorted by value, descending
    double duration = 14.0, delay = 0.0;
    double directive_begin = std::numeric_limits<double>::lowest(), directive_end = std::numeric_limits<double>::max();
    double base = 1.0;

    // dynamic_npis.h:56-61 (insert_sorted_replace, stl_util.h:82-96)
    void set_threshold(double threshold, const std::vector<DampingSamplingO>& dampings)
    {
        auto pred = [](auto& t1, auto& t2) {
            return t1.first > t2.first;
        };
        auto item = std::make_pair(threshold, dampings);
        auto lb   = std::lower_bound(thresholds.begin(), thresholds.end(), item, pred);
        if (lb != thresholds.end() && !pred(*lb, item) && !pred(item, *lb)) {
            *lb = item;
        }
        else {
            thresholds.insert(lb, item);
        }
    }
    // dynamic_npis.h:70-79: index of the biggest threshold that `value` exceeds, thresholds.size() if none
    size_t get_max_exceeded_threshold(double value) const
    {
        auto it = std::upper_bound(thresholds.begin(), thresholds.end(), value, [](auto& val, auto& t2) {
            return val > t2.first;
        });
        return size_t(it - thresholds.begin());
    }
};

// dynamic_npis.h:270-283 (the time period is an open interval)
inline std::vector<size_t> get_damping_indices(const DampingExpr& e, int lvl, int type, double begin, double end)
{
    std::vector<size_t> indices;
    for (size_t i = 0; i < e.dampings.list.size(); ++i) {
        const auto& d = e.dampings.list[i];
        if (d.level == lvl && d.type == type && d.time > begin && d.time < end) {
            indices.push_back(i);
        }
    }
    return indices;
}
// dynamic_npis.h:294-306
inline Mat get_active_damping(const DampingExpr& e, int lvl, int type, double t)
{
    for (auto it = e.dampings.list.rbegin(); it != e.dampings.list.rend(); ++it) {
        if (it->level == lvl && it->type == type && it->time <= t) {
            return it->m;
        }
    }
    return Mat(e.baseline.r, e.baseline.c, 0.0);
}

// dynamic_npis.h:335-395
template <class MakeMatrix>
void implement_dynamic_npis(DampingExprGroup& group, const std::vector<DampingSamplingO>& npis, double begin, double end,
                            MakeMatrix&& make_matrix)
{
    auto cwise_max = [](const Mat& a, const Mat& b) {
        Mat r = a;
        for (size_t k = 0; k < r.a.size(); ++k) {
            r.a[k] = std::max(a.a[k], b.a[k]);
        }
        return r;
    };
    for (auto& npi : npis) {
        for (auto& mat_idx : npi.matrix_indices) {
            const int type = npi.type, level = npi.level;
            auto& expr     = group.m[size_t(mat_idx)];
            const Mat active     = get_active_damping(expr, level, type, begin);
            const Mat active_end = get_active_damping(expr, level, type, end); // copy: it may be removed or changed
            Vec weighted(npi.group_weights.size());
            for (size_t k = 0; k < weighted.size(); ++k) {
                weighted[k] = npi.value * npi.group_weights[k];
            }
            const Mat value      = make_matrix(weighted);
            bool npi_implemented = false;
            // add begin of npi if not already bigger
            bool any_smaller = false;
            for (size_t k = 0; k < value.a.size(); ++k) {
                any_smaller = any_smaller || active.a[k] < value.a[k];
            }
            if (any_smaller) {
                expr.dampings.add(Damping{begin, level, type, cwise_max(value, active)});
                npi_implemented = true;
            }
            // replace dampings during the new npi
            for (auto i : get_damping_indices(expr, level, type, begin, end)) {
                const Damping d = expr.dampings.list[i];
                expr.dampings.add(Damping{d.time, level, type, cwise_max(d.m, value)});
                npi_implemented = true;
            }
            // add end of npi to restore active dampings if any change was made
            if (npi_implemented) {
                expr.dampings.add(Damping{end, level, type, active_end});
            }
        }
    }
    // remove duplicates that accumulated because of dampings that become active during the time span: a damping is
    // obsolete if the most recent damping of the same type and level has the same value
    for (auto& expr : group.m) {
        auto& list = expr.dampings.list;
        for (int i = 0; i < int(list.size()) - 1; ++i) {
            auto it      = list.rbegin() + i;
            auto it_prev = std::find_if(it + 1, list.rend(), [&di = *it](auto& dj) {
                return di.level == dj.level && di.type == dj.type;
            });
            if (it_prev != list.rend() && it->m.a == it_prev->m.a) {
                list.erase(list.begin() + (long(list.size()) - 1 - i));
                expr.dampings.cache.clear();
                expr.dampings.update_cache();
            }
        }
    }
}

namespace osecir
{

// model.h:413-423: symptomatic infections relative to the model's (initial) total population
inline double get_infections_relative(const Model& model, const Vec& y)
{
    double sum_inf = 0;
    for (int i = 0; i < model.parameters.G; ++i) {
        sum_inf += y[size_t(i * NC + ISy)];
    }
    double total = 0;
    for (double v : model.populations) {
        total += v;
    }
    return sum_inf / total;
}

// The dynamic NPI part of osecir::Simulation (model.h:289-367: DynamicNPIsInfectedSymptoms, m_dynamic_npi) around a
// node's base simulation. `frozen`: the base advance appends the last value unchanged (the MockSimulation of
// cpp/tests/test_dynamic_npis.cpp).
struct NpiController {
    DynamicNPIs npis;
    std::pair<double, double> dynamic_npi = {-std::numeric_limits<double>::max(), 0.0};
    bool frozen                           = false;
    int num_implemented                   = 0;

    void base_advance(orc::Simulation<Model>& sim, double tmax) const
    {
        if (frozen) {
            const Vec v = sim.result.last();
            sim.result.add(tmax, v);
        }
        else {
            sim.advance(tmax);
        }
    }
    Vec& advance(orc::Simulation<Model>& sim, double tmax)
    {
        if (npis.thresholds.empty()) {
            base_advance(sim, tmax);
            return sim.result.last();
        }
        auto& contact_patterns = sim.model.parameters.ContactPatterns;
        double delay_npi_implementation;
        double t = sim.result.times.back();
        while (t < tmax) {
            if (t > 0) {
                delay_npi_implementation = npis.delay;
            }
            else { // DynamicNPIs for t=0 are treated as 'from-start NPIs'
                delay_npi_implementation = 0;
            }
            const double direc_begin = npis.directive_begin, direc_end = npis.directive_end;
            if (floating_point_greater_equal(t, direc_begin, 1e-10) && t < direc_end) {
                const double inf_rel   = get_infections_relative(sim.model, sim.result.last()) * npis.base;
                const size_t exceeded  = npis.get_max_exceeded_threshold(inf_rel);
                const bool npi_expired = floating_point_greater_equal(t, dynamic_npi.second, 1e-10);
                if (exceeded != npis.thresholds.size() &&
                    (npis.thresholds[exceeded].first > dynamic_npi.first || npi_expired)) {
                    const bool is_expiry_renewal =
                        npi_expired && !(npis.thresholds[exceeded].first > dynamic_npi.first);
                    const double effective_delay = is_expiry_renewal ? 0.0 : delay_npi_implementation;
                    if (t + effective_delay < direc_end) {
                        const double t_start = t + effective_delay;
                        const double t_end   = std::min(direc_end, t_start + npis.duration);
                        dynamic_npi          = std::make_pair(npis.thresholds[exceeded].first, t_end);
                        const double t_start_damping = (!is_expiry_renewal && t_start > 0) ? t_start + 1.0 : t_start;
                        const double t_end_damping   = t_end + 1.0;
                        implement_dynamic_npis(contact_patterns, npis.thresholds[exceeded].second, t_start_damping,
                                               t_end_damping, [](const Vec& g) {
                                                   return make_contact_damping_matrix(g);
                                               });
                        ++num_implemented;
                    }
                }
            }
            double dt_eff        = std::min(1.0, tmax - t);
            const double npi_end = dynamic_npi.second;
            if (t < npi_end) {
                dt_eff = std::min(dt_eff, npi_end - t);
            }
            base_advance(sim, t + dt_eff);
            t = t + dt_eff;
        }
        return sim.result.last();
    }
};

struct NpiSimulation {
    orc::Simulation<Model> sim;
    NpiController ctl;
    NpiSimulation(const Model& m, const DynamicNPIs& n, double t0 = 0., double dt = 0.1)
        : sim(m, t0, dt)
    {
        ctl.npis = n;
    }
    Vec& advance(double tmax)
    {
        return ctl.advance(sim, tmax);
    }
};

} // namespace osecir
} // namespace orc

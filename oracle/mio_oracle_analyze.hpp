// mio_oracle_analyze.hpp -- CPU ORACLE (test infrastructure, NOT product code).
//
// Literal restatement of the ensemble post-processing of cpp/memilio/data/analyze_result.h (SciCompMod/memilio @
// 8851d5be): interpolate_simulation_result :186-265 (+ math/interpolation.h:44-47), sum_nodes :267-287,
// ensemble_mean :289-314, ensemble_percentile :316-343. Pinned by the small exact cases of the reference's own tests
// (cpp/tests/test_analyze_result.cpp), see tests/test_oracle_analyze.py.
#pragma once

#include "mio_oracle.hpp"

namespace orc
{

// math/interpolation.h:44-47
inline Vec linear_interpolation(double x_eval, double x_1, double x_2, const Vec& y1, const Vec& y2)
{
    const double weight = (x_eval - x_1) / (x_2 - x_1);
    Vec out(y1.size());
    for (size_t i = 0; i < y1.size(); ++i) {
        out[i] = y1[i] + weight * (y2[i] - y1[i]);
    }
    return out;
}

// analyze_result.h:211-265
inline TimeSeries interpolate_simulation_result(const TimeSeries& simulation_result,
                                                const std::vector<double>& interpolation_times)
{
    TimeSeries interpolated;
    if (interpolation_times.empty()) {
        return interpolated;
    }
    size_t interp_idx = 0;
    if (simulation_result.times[0] >= interpolation_times[0]) {
        interpolated.add(interpolation_times[0], simulation_result.values[0]);
        ++interp_idx;
    }
    for (size_t sim_idx = 0; sim_idx + 1 < simulation_result.n() && interp_idx < interpolation_times.size();) {
        if (simulation_result.times[sim_idx] < interpolation_times[interp_idx] &&
            simulation_result.times[sim_idx + 1] >= interpolation_times[interp_idx]) {
            interpolated.add(interpolation_times[interp_idx],
                             linear_interpolation(interpolation_times[interp_idx], simulation_result.times[sim_idx],
                                                  simulation_result.times[sim_idx + 1],
                                                  simulation_result.values[sim_idx],
                                                  simulation_result.values[sim_idx + 1]));
            ++interp_idx;
        }
        else {
            ++sim_idx;
        }
    }
    if (interp_idx < interpolation_times.size() && simulation_result.last_time() < interpolation_times[interp_idx]) {
        interpolated.add(interpolation_times[interp_idx], simulation_result.last());
    }
    return interpolated;
}

// analyze_result.h:186-209; default abs_tol = 100 * Limits<double>::zero_tolerance() = 1e-10 (:50-51, config.h:72-79)
inline TimeSeries interpolate_simulation_result(const TimeSeries& simulation_result, double abs_tol = 100. * 1e-12)
{
    const double t0    = simulation_result.times[0];
    const double t_max = simulation_result.last_time();
    const int day0     = static_cast<int>(std::ceil(t0 - abs_tol));
    const int day_max  = static_cast<int>(std::floor(t_max + abs_tol));
    std::vector<double> tps(day_max - day0 + 1);
    std::iota(tps.begin(), tps.end(), double(day0));
    return interpolate_simulation_result(simulation_result, tps);
}

using EnsembleResult = std::vector<std::vector<TimeSeries>>; // [run][node]

// analyze_result.h:267-287
inline EnsembleResult sum_nodes(const EnsembleResult& ensemble_result)
{
    const size_t num_runs = ensemble_result.size(), num_nodes = ensemble_result[0].size();
    const size_t num_time_points = ensemble_result[0][0].n(), num_elements = ensemble_result[0][0].values[0].size();
    EnsembleResult sum_result(num_runs, std::vector<TimeSeries>(1));
    for (size_t run = 0; run < num_runs; run++) {
        for (size_t time = 0; time < num_time_points; time++) {
            sum_result[run][0].add(ensemble_result[run][0].times[time], Vec(num_elements, 0.0));
            for (size_t node = 0; node < num_nodes; node++) {
                for (size_t e = 0; e < num_elements; ++e) {
                    sum_result[run][0].values[time][e] += ensemble_result[run][node].values[time][e];
                }
            }
        }
    }
    return sum_result;
}

// analyze_result.h:289-314
inline std::vector<TimeSeries> ensemble_mean(const EnsembleResult& ensemble_result)
{
    const size_t num_runs = ensemble_result.size(), num_nodes = ensemble_result[0].size();
    const size_t num_time_points = ensemble_result[0][0].n(), num_elements = ensemble_result[0][0].values[0].size();
    std::vector<TimeSeries> mean(num_nodes);
    for (size_t node = 0; node < num_nodes; node++) {
        for (size_t time = 0; time < num_time_points; time++) {
            mean[node].add(0.0, Vec(num_elements, 0.0));
        }
    }
    for (size_t run = 0; run < num_runs; run++) {
        for (size_t node = 0; node < num_nodes; node++) {
            for (size_t time = 0; time < num_time_points; time++) {
                mean[node].times[time] = ensemble_result[run][node].times[time];
                for (size_t e = 0; e < num_elements; ++e) {
                    mean[node].values[time][e] += ensemble_result[run][node].values[time][e] / double(num_runs);
                }
            }
        }
    }
    return mean;
}

// analyze_result.h:316-343
inline std::vector<TimeSeries> ensemble_percentile(const EnsembleResult& ensemble_result, double p)
{
    assert(p > 0.0 && p < 1.0);
    const size_t num_runs = ensemble_result.size(), num_nodes = ensemble_result[0].size();
    const size_t num_time_points = ensemble_result[0][0].n(), num_elements = ensemble_result[0][0].values[0].size();
    std::vector<TimeSeries> percentile(num_nodes);
    std::vector<double> single_element_ensemble(num_runs);
    for (size_t node = 0; node < num_nodes; node++) {
        for (size_t time = 0; time < num_time_points; time++) {
            percentile[node].add(ensemble_result[0][node].times[time], Vec(num_elements, 0.0));
            for (size_t elem = 0; elem < num_elements; elem++) {
                std::transform(ensemble_result.begin(), ensemble_result.end(), single_element_ensemble.begin(),
                               [=](auto& run) {
                                   return run[node].values[time][elem];
                               });
                std::sort(single_element_ensemble.begin(), single_element_ensemble.end());
                percentile[node].values[time][elem] = single_element_ensemble[static_cast<size_t>(num_runs * p)];
            }
        }
    }
    return percentile;
}

} // namespace orc

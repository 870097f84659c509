// io.hpp -- the few IO entry points the Graph-ODE examples call (cpp/memilio/io/io.h:49-59,354, io/result_io.h:44-215,
// utils/base_dir.h). save_result writes the reference's HDF5 layout (one group per region id with the datasets Time,
// Group1..GroupG, Total; io/result_io.cpp:33-100) through the library-free writer of h5lite.hpp;
// save_results / save_edges produce the reference's files (results_run<i>.h5, results_run<i>_sum.h5, pXX/Results.h5,
// pXX/Results_sum.h5). The percentile and mean reductions run on the device (mio_cuda_reduce_host_columns).
#pragma once

#include "h5lite.hpp"
#include "study.hpp"

#include <filesystem>
#include <fstream>
#include <optional>

namespace mio
{

enum class StatusCode
{
    OK = 0,
    UnknownError,
    OutOfRange,
    InvalidValue,
    InvalidFileFormat,
    KeyNotFound,
    InvalidType,
    FileNotFound
};

struct IOStatus {
    StatusCode code = StatusCode::OK;
    std::string message;
    bool is_ok() const
    {
        return code == StatusCode::OK;
    }
};

template <class T>
class IOResult
{
public:
    IOResult(T v)
        : m_value(std::move(v))
    {
    }
    IOResult(IOStatus s)
        : m_status(std::move(s))
    {
    }
    explicit operator bool() const
    {
        return m_status.is_ok();
    }
    T& value()
    {
        return *m_value;
    }
    const IOStatus& error() const
    {
        return m_status;
    }

private:
    std::optional<T> m_value;
    IOStatus m_status;
};
template <>
class IOResult<void>
{
public:
    IOResult() = default;
    IOResult(IOStatus s)
        : m_status(std::move(s))
    {
    }
    explicit operator bool() const
    {
        return m_status.is_ok();
    }
    const IOStatus& error() const
    {
        return m_status;
    }

private:
    IOStatus m_status;
};
inline IOResult<void> success()
{
    return IOResult<void>();
}
inline IOStatus failure(StatusCode c, const std::string& msg)
{
    return IOStatus{c, msg};
}

inline std::filesystem::path example_results_dir(const std::string& name)
{
    return std::filesystem::path("example_results") / name;
}
inline std::filesystem::path create_directories_or_exit(const std::filesystem::path& p)
{
    std::error_code ec;
    std::filesystem::create_directories(p, ec);
    if (ec) {
        std::cerr << "Failed to create directory " << p << ": " << ec.message() << "\n";
        std::exit(1);
    }
    return p;
}

namespace detail
{
inline IOResult<void> write_csv(const std::vector<TimeSeries<ScalarType>>& nodes, const std::vector<int>& ids,
                                const std::filesystem::path& file)
{
    std::ofstream out(file);
    if (!out) {
        return failure(StatusCode::FileNotFound, "cannot open " + file.string());
    }
    out.precision(17);
    out << "node,time";
    for (Eigen::Index e = 0; nodes.size() && e < nodes[0].get_num_elements(); ++e) {
        out << ",y" << e;
    }
    out << "\n";
    for (size_t k = 0; k < nodes.size(); ++k) {
        for (Eigen::Index t = 0; t < nodes[k].get_num_time_points(); ++t) {
            out << (k < ids.size() ? ids[k] : int(k)) << "," << nodes[k].get_time(t);
            for (Eigen::Index e = 0; e < nodes[k].get_num_elements(); ++e) {
                out << "," << nodes[k][t][e];
            }
            out << "\n";
        }
    }
    return success();
}
} // namespace detail

// save_result, io/result_io.h:44, io/result_io.cpp:33-100: per region id a group with "Time" [T], "Group<k>" [T][C] for
// every age group and "Total" [T][C] (their sum, accumulated in group order)
inline IOResult<void> save_result(const std::vector<TimeSeries<ScalarType>>& results, const std::vector<int>& ids,
                                  int num_groups, const std::string& filename)
{
    std::vector<h5lite::Group> groups;
    for (size_t r = 0; r < results.size(); ++r) {
        const auto& ts  = results[r];
        const size_t T  = size_t(ts.get_num_time_points());
        const size_t C  = size_t(ts.get_num_elements()) / size_t(num_groups);
        h5lite::Group g;
        g.name = std::to_string(r < ids.size() ? ids[r] : int(r));
        h5lite::Dataset time{"Time", {T}, {}};
        for (size_t t = 0; t < T; ++t) {
            time.values.push_back(ts.get_time(Eigen::Index(t)));
        }
        g.datasets.push_back(std::move(time));
        h5lite::Dataset total{"Total", {T, C}, std::vector<double>(T * C, 0.0)};
        for (int k = 0; k < num_groups; ++k) {
            h5lite::Dataset grp{"Group" + std::to_string(k + 1), {T, C}, std::vector<double>(T * C, 0.0)};
            for (size_t t = 0; t < T; ++t) {
                for (size_t c = 0; c < C; ++c) {
                    const double v        = ts[Eigen::Index(t)][Eigen::Index(size_t(k) * C + c)];
                    grp.values[t * C + c] = v;
                    total.values[t * C + c] += v;
                }
            }
            g.datasets.push_back(std::move(grp));
        }
        g.datasets.push_back(std::move(total));
        groups.push_back(std::move(g));
    }
    if (!h5lite::write_file(filename, groups)) {
        return failure(StatusCode::FileNotFound, "Failed to open the file " + filename);
    }
    return success();
}

// SimulationResult / read_result, io/result_io.h:47-110, io/result_io.cpp:110-230: the groups' series (all age groups
// side by side) and the totals of every region of a result file, regions in ascending id order
class SimulationResult
{
public:
    SimulationResult(const TimeSeries<ScalarType>& groups, const TimeSeries<ScalarType>& totals)
        : m_groups(groups)
        , m_totals(totals)
    {
    }
    const TimeSeries<ScalarType>& get_groups() const
    {
        return m_groups;
    }
    const TimeSeries<ScalarType>& get_totals() const
    {
        return m_totals;
    }

private:
    TimeSeries<ScalarType> m_groups, m_totals;
};

inline IOResult<std::vector<SimulationResult>> read_result(const std::string& filename)
{
    try {
        h5lite::Reader file(filename);
        auto regions = file.list();
        std::sort(regions.begin(), regions.end(), [](const auto& a, const auto& b) {
            return std::stoi(a.name) < std::stoi(b.name);
        });
        std::vector<SimulationResult> out;
        for (auto& region : regions) {
            std::vector<h5lite::Dataset> groups;
            h5lite::Dataset time, total;
            for (auto& m : file.list(&region)) {
                if (m.name == "Time") {
                    time = file.dataset(m);
                }
                else if (m.name == "Total") {
                    total = file.dataset(m);
                }
            }
            for (int k = 1;; ++k) { // Group1, Group2, ... in numerical order
                bool found = false;
                for (auto& m : file.list(&region)) {
                    if (m.name == "Group" + std::to_string(k)) {
                        groups.push_back(file.dataset(m));
                        found = true;
                    }
                }
                if (!found) {
                    break;
                }
            }
            if (time.shape.size() != 1 || total.shape.size() != 2 || total.shape[0] != time.shape[0]) {
                return failure(StatusCode::InvalidFileFormat, "Number of time points does not match.");
            }
            const size_t T = size_t(time.shape[0]), C = size_t(total.shape[1]);
            TimeSeries<ScalarType> gs{Eigen::Index(C * groups.size())}, ts{Eigen::Index(C)};
            for (size_t t = 0; t < T; ++t) {
                Eigen::VectorX<ScalarType> gv{Eigen::Index(C * groups.size())}, tv{Eigen::Index(C)};
                for (size_t k = 0; k < groups.size(); ++k) {
                    if (groups[k].shape != total.shape) {
                        return failure(StatusCode::InvalidFileFormat, "Number of infection states does not match.");
                    }
                    for (size_t c = 0; c < C; ++c) {
                        gv[Eigen::Index(k * C + c)] = groups[k].values[t * C + c];
                    }
                }
                for (size_t c = 0; c < C; ++c) {
                    tv[Eigen::Index(c)] = total.values[t * C + c];
                }
                gs.add_time_point(time.values[t], gv);
                ts.add_time_point(time.values[t], tv);
            }
            out.emplace_back(gs, ts);
        }
        return out;
    }
    catch (const std::exception& e) {
        return failure(StatusCode::InvalidFileFormat, e.what());
    }
}

// save_edges of one run, io/mobility_io.cpp:206-306: one group per start node with "Time", "End<target id>" for every
// edge of that node and "Total" (their sum in edge order); the edges of a start node are consecutive
inline IOResult<void> save_edges(const std::vector<TimeSeries<ScalarType>>& results,
                                 const std::vector<std::pair<int, int>>& ids, const std::string& filename)
{
    std::vector<h5lite::Group> groups;
    size_t e = 0;
    while (e < results.size()) {
        const auto& first = results[e];
        const size_t T    = size_t(first.get_num_time_points());
        if (T == 0) {
            return failure(StatusCode::InvalidValue, "No time points in the TimeSeries for Edge combination " +
                                                         std::to_string(ids[e].first) + " -> " +
                                                         std::to_string(ids[e].second));
        }
        const size_t C = size_t(first.get_num_elements());
        h5lite::Group g;
        g.name = std::to_string(ids[e].first);
        h5lite::Dataset time{"Time", {T}, {}};
        for (size_t t = 0; t < T; ++t) {
            time.values.push_back(first.get_time(Eigen::Index(t)));
        }
        g.datasets.push_back(std::move(time));
        h5lite::Dataset total{"Total", {T, C}, std::vector<double>(T * C, 0.0)};
        const int start = ids[e].first;
        for (; e < results.size() && ids[e].first == start; ++e) {
            h5lite::Dataset d{"End" + std::to_string(ids[e].second), {T, C}, std::vector<double>(T * C, 0.0)};
            const auto& r = results[e];
            for (size_t t = 0; t < std::min(T, size_t(r.get_num_time_points())); ++t) {
                for (size_t c = 0; c < C; ++c) {
                    d.values[t * C + c] = r[Eigen::Index(t)][Eigen::Index(c)];
                    total.values[t * C + c] += d.values[t * C + c];
                }
            }
            g.datasets.push_back(std::move(d));
        }
        g.datasets.push_back(std::move(total));
        groups.push_back(std::move(g));
    }
    if (!h5lite::write_file(filename, groups)) {
        return failure(StatusCode::FileNotFound, "Failed to open the HDF5 file: " + filename);
    }
    return success();
}

// save_results, io/result_io.h:135-202: per-run results (optional) and the p05/p25/p50/p75/p95 percentiles of the
// per-node results and of the node-summed results
template <class Model>
IOResult<void> save_results(const std::vector<std::vector<TimeSeries<ScalarType>>>& ensemble_results,
                            const std::vector<std::vector<Model>>& ensemble_params, const std::vector<int>& county_ids,
                            const std::filesystem::path& result_dir, bool save_single_runs = true,
                            bool save_percentiles = true)
{
    auto ensemble_result_sum = sum_nodes(ensemble_results);
    const int num_groups     = int(size_t(ensemble_params.at(0).at(0).parameters.get_num_groups()));
    if (save_single_runs) {
        for (size_t i = 0; i < ensemble_results.size(); ++i) {
            auto r1 = save_result(ensemble_result_sum[i], {0}, num_groups,
                                  (result_dir / ("results_run" + std::to_string(i) + "_sum.h5")).string());
            auto r2 = save_result(ensemble_results[i], county_ids, num_groups,
                                  (result_dir / ("results_run" + std::to_string(i) + ".h5")).string());
            if (!r1 || !r2) {
                return !r1 ? r1 : r2;
            }
        }
    }
    if (save_percentiles) {
        const std::pair<const char*, ScalarType> ps[] = {{"p05", 0.05}, {"p25", 0.25}, {"p50", 0.5},
                                                          {"p75", 0.75}, {"p95", 0.95}};
        for (auto& [name, p] : ps) {
            auto dir = create_directories_or_exit(result_dir / name);
            auto r1  = save_result(ensemble_percentile(ensemble_result_sum, p), {0}, num_groups,
                                   (dir / "Results_sum.h5").string());
            auto r2  = save_result(ensemble_percentile(ensemble_results, p), county_ids, num_groups,
                                   (dir / "Results.h5").string());
            if (!r1 || !r2) {
                return !r1 ? r1 : r2;
            }
        }
    }
    return success();
}

// save_edges of an ensemble, io/mobility_io.cpp:162-204: Edges_run<i>.h5 per run and pXX/Edges.h5 for the percentiles
inline IOResult<void> save_edges(const std::vector<std::vector<TimeSeries<ScalarType>>>& ensemble_edges,
                                 const std::vector<std::pair<int, int>>& pairs_edges,
                                 const std::filesystem::path& result_dir, bool save_single_runs = true,
                                 bool save_percentiles = true)
{
    if (save_single_runs) {
        for (size_t i = 0; i < ensemble_edges.size(); ++i) {
            auto r = save_edges(ensemble_edges[i], pairs_edges,
                                (result_dir / ("Edges_run" + std::to_string(i) + ".h5")).string());
            if (!r) {
                return r;
            }
        }
    }
    if (save_percentiles) {
        const std::pair<const char*, ScalarType> ps[] = {{"p05", 0.05}, {"p25", 0.25}, {"p50", 0.5},
                                                          {"p75", 0.75}, {"p95", 0.95}};
        for (auto& [name, p] : ps) {
            auto dir = create_directories_or_exit(result_dir / name);
            auto r   = save_edges(ensemble_percentile(ensemble_edges, p), pairs_edges, (dir / "Edges.h5").string());
            if (!r) {
                return r;
            }
        }
    }
    return success();
}

} // namespace mio

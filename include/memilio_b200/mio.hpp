// mio.hpp -- header-only C++ facade of the B200 backend with MEmilio's names and signatures for the Graph-ODE-SECIR
// ensemble path. It packs the objects an existing MEmilio program builds into the flat descriptors of
// include/mio_cuda.h and runs them on the GPU; there is no CPU implementation behind it.
//
// Mirrors (reference paths): mio::Graph cpp/memilio/mobility/graph.h:145-300; SimulationNode / MobilityParameters /
// MobilityEdge / make_mobility_sim cpp/memilio/mobility/metapopulation_mobility_instant.h:40-101,115-260,270-420,
// 656-678; GraphSimulation cpp/memilio/mobility/graph_simulation.h:41-121; osecir::Model / Parameters /
// InfectionState cpp/models/ode_secir/{model.h,parameters.h,infection_state.h}; osecir::draw_sample
// cpp/models/ode_secir/parameter_space.h:125-251; ContactMatrix(Group) cpp/memilio/epidemiology/contact_matrix.h;
// UncertainContactMatrix epidemiology/uncertain_matrix.h; UncertainValue utils/uncertain_value.h;
// ParameterDistribution* utils/parameter_distributions.h; RandomNumberGenerator utils/random_number_generator.h;
// TimeSeries utils/time_series.h; ParameterStudy compartments/parameter_studies.h:47-374;
// interpolate_simulation_result / sum_nodes / ensemble_mean / ensemble_percentile data/analyze_result.h.
//
// Compiles without Eigen/Boost: the small part of Eigen::VectorX / MatrixX the examples touch is provided by
// eigen_shim.hpp unless the real Eigen is included first.
#pragma once

#include "../mio_cuda.h"
#include "eigen_shim.hpp"

#include <algorithm>
#include <array>
#include <cassert>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <functional>
#include <iostream>
#include <limits>
#include <memory>
#include <random>
#include <stdexcept>
#include <string>
#include <tuple>
#include <type_traits>
#include <utility>
#include <vector>

using ScalarType = double;

namespace mio
{

// ---- logging (utils/logging.h): the facade only keeps the level ---------------------------------------------------
enum class LogLevel
{
    trace,
    debug,
    info,
    warn,
    err,
    critical,
    off
};
inline LogLevel& detail_log_level()
{
    static LogLevel l = LogLevel::warn;
    return l;
}
inline void set_log_level(LogLevel l)
{
    detail_log_level() = l;
}
template <class... T>
inline void log_info(const char*, T&&...)
{
}
template <class... T>
inline void log_warning(const char*, T&&...)
{
}

namespace mpi
{
inline bool is_root()
{
    return true;
}
} // namespace mpi

inline void cuda_check(int rc)
{
    if (rc != MIO_OK) {
        throw std::runtime_error(std::string("mio_cuda: ") + mio_cuda_last_error_string());
    }
}

// ---- AgeGroup / indices (epidemiology/age_group.h, utils/index.h) -------------------------------------------------
struct AgeGroup {
    size_t value = 0;
    AgeGroup() = default;
    AgeGroup(size_t v)
        : value(v)
    {
    }
    AgeGroup(int v)
        : value(size_t(v))
    {
    }
    explicit operator size_t() const
    {
        return value;
    }
    size_t get() const
    {
        return value;
    }
    AgeGroup& operator++()
    {
        ++value;
        return *this;
    }
    AgeGroup operator++(int)
    {
        AgeGroup t = *this;
        ++value;
        return t;
    }
    bool operator<(const AgeGroup& o) const
    {
        return value < o.value;
    }
    bool operator==(const AgeGroup& o) const
    {
        return value == o.value;
    }
    bool operator!=(const AgeGroup& o) const
    {
        return value != o.value;
    }
};

template <class FP>
struct SimulationTime {
    FP value;
    explicit SimulationTime(FP v)
        : value(v)
    {
    }
    FP get() const
    {
        return value;
    }
};
struct DampingLevel {
    int value;
    explicit DampingLevel(int v = 0)
        : value(v)
    {
    }
};
struct DampingType {
    int value;
    explicit DampingType(int v = 0)
        : value(v)
    {
    }
};

// ---- random numbers (utils/random_number_generator.h:295-338): generator values come from the C library -----------
class RandomNumberGenerator
{
public:
    using result_type = uint64_t;
    RandomNumberGenerator()
    {
        std::random_device rd;
        seed({rd(), rd(), rd(), rd(), rd(), rd()});
    }
    static constexpr result_type min()
    {
        return 0;
    }
    static constexpr result_type max()
    {
        return std::numeric_limits<result_type>::max();
    }
    result_type operator()()
    {
        uint64_t out = 0;
        cuda_check(mio_study_rng_raw(m_seeds.data(), int32_t(m_seeds.size()), m_counter, 1, &out, nullptr));
        ++m_counter;
        return out;
    }
    void seed(const std::vector<uint32_t>& seeds)
    {
        m_seeds = seeds;
    }
    const std::vector<uint32_t>& get_seeds() const
    {
        return m_seeds;
    }
    uint64_t get_counter() const
    {
        return m_counter;
    }
    void set_counter(uint64_t c)
    {
        m_counter = c;
    }
    void synchronize()
    {
    }

private:
    std::vector<uint32_t> m_seeds;
    uint64_t m_counter = 0;
};

inline RandomNumberGenerator& thread_local_rng()
{
    static thread_local RandomNumberGenerator rng;
    return rng;
}

// ---- parameter distributions (utils/parameter_distributions.h) ----------------------------------------------------
class ParameterDistribution
{
public:
    virtual ~ParameterDistribution()                                 = default;
    virtual ScalarType get_rand_sample(RandomNumberGenerator& rng)   = 0;
    virtual ParameterDistribution* clone() const                     = 0;
    virtual mio_uncertain_value to_pod(ScalarType value) const       = 0;
    ScalarType get_sample(RandomNumberGenerator& rng)
    {
        return get_rand_sample(rng);
    }
};

class ParameterDistributionUniform : public ParameterDistribution
{
public:
    ParameterDistributionUniform(ScalarType lower_bound, ScalarType upper_bound)
        : m_upper_bound(upper_bound)
        , m_lower_bound(lower_bound)
    {
    }
    ScalarType get_lower_bound() const
    {
        return m_lower_bound;
    }
    ScalarType get_upper_bound() const
    {
        return m_upper_bound;
    }
    ScalarType get_rand_sample(RandomNumberGenerator& rng) override
    {
        return std::uniform_real_distribution<ScalarType>(m_lower_bound, m_upper_bound)(rng);
    }
    ParameterDistribution* clone() const override
    {
        return new ParameterDistributionUniform(*this);
    }
    mio_uncertain_value to_pod(ScalarType value) const override
    {
        return mio_uncertain_value{MIO_DIST_UNIFORM, 0, value, m_lower_bound, m_upper_bound, 0, 0, 0};
    }

private:
    ScalarType m_upper_bound, m_lower_bound;
};

class ParameterDistributionNormal : public ParameterDistribution
{
public:
    ParameterDistributionNormal(ScalarType lower_bound, ScalarType upper_bound, ScalarType mean,
                                ScalarType standard_dev, ScalarType quantile = 2.5758)
        : m_mean(mean)
        , m_standard_dev(standard_dev)
        , m_upper_bound(upper_bound)
        , m_lower_bound(lower_bound)
        , m_quantile(quantile)
    {
    }
    ScalarType get_rand_sample(RandomNumberGenerator& rng) override
    {
        if (m_upper_bound == m_lower_bound) {
            return m_lower_bound;
        }
        if (m_mean < m_lower_bound || m_mean > m_upper_bound) {
            m_mean = 0.5 * (m_upper_bound - m_lower_bound);
        }
        if (m_mean + m_standard_dev * m_quantile > m_upper_bound) {
            m_standard_dev = (m_upper_bound - m_mean) / m_quantile;
        }
        if (m_mean - m_standard_dev * m_quantile < m_lower_bound) {
            m_standard_dev = (m_mean - m_lower_bound) / m_quantile;
        }
        int i            = 0;
        int retries      = 10;
        ScalarType rnumb = std::normal_distribution<ScalarType>(m_mean, m_standard_dev)(rng);
        while ((rnumb > m_upper_bound || rnumb < m_lower_bound) && i < retries) {
            rnumb = std::normal_distribution<ScalarType>(m_mean, m_standard_dev)(rng);
            i++;
            if (i == retries) {
                rnumb = rnumb > m_upper_bound ? m_upper_bound : m_lower_bound;
            }
        }
        return rnumb;
    }
    ParameterDistribution* clone() const override
    {
        return new ParameterDistributionNormal(*this);
    }
    mio_uncertain_value to_pod(ScalarType value) const override
    {
        return mio_uncertain_value{MIO_DIST_NORMAL, 0,           value,        m_mean,
                                   m_standard_dev,  m_lower_bound, m_upper_bound, m_quantile};
    }

private:
    ScalarType m_mean, m_standard_dev, m_upper_bound, m_lower_bound, m_quantile;
};

// ---- UncertainValue (utils/uncertain_value.h:45-258) --------------------------------------------------------------
template <class FP = ScalarType>
class UncertainValue
{
public:
    UncertainValue(FP v = 0.)
        : m_value(v)
    {
    }
    UncertainValue(FP v, const ParameterDistribution& dist)
        : m_value(v)
        , m_dist(dist.clone())
    {
    }
    UncertainValue(const UncertainValue& o)
        : m_value(o.m_value)
        , m_dist(o.m_dist ? o.m_dist->clone() : nullptr)
    {
    }
    UncertainValue(UncertainValue&&) = default;
    UncertainValue& operator=(const UncertainValue& o)
    {
        if (this != &o) {
            m_value = o.m_value;
            m_dist.reset(o.m_dist ? o.m_dist->clone() : nullptr);
        }
        return *this;
    }
    UncertainValue& operator=(UncertainValue&&) = default;
    // assigning a scalar keeps the distribution (uncertain_value.h:107-111)
    UncertainValue& operator=(FP v)
    {
        m_value = v;
        return *this;
    }
    operator FP&()
    {
        return m_value;
    }
    operator const FP&() const
    {
        return m_value;
    }
    FP value() const
    {
        return m_value;
    }
    void set_distribution(const ParameterDistribution& dist)
    {
        m_dist.reset(dist.clone());
    }
    const ParameterDistribution* get_distribution() const
    {
        return m_dist.get();
    }
    FP draw_sample()
    {
        if (m_dist) {
            m_value = m_dist->get_sample(thread_local_rng());
        }
        return m_value;
    }
    mio_uncertain_value to_pod() const
    {
        if (m_dist) {
            return m_dist->to_pod(m_value);
        }
        return mio_uncertain_value{MIO_DIST_NONE, 0, m_value, 0, 0, 0, 0, 0};
    }

private:
    FP m_value;
    std::unique_ptr<ParameterDistribution> m_dist;
};

// array of values indexed by AgeGroup (utils/custom_index_array.h): assignment of a scalar sets every element
template <class T>
class AgeArray
{
public:
    AgeArray(size_t n = 0, T v = T())
        : m_v(n, v)
    {
    }
    T& operator[](AgeGroup i)
    {
        return m_v[i.value];
    }
    const T& operator[](AgeGroup i) const
    {
        return m_v[i.value];
    }
    template <class U, class = std::enable_if_t<std::is_arithmetic_v<U>>>
    AgeArray& operator=(U v)
    {
        for (auto& x : m_v) {
            x = ScalarType(v);
        }
        return *this;
    }
    size_t numel() const
    {
        return m_v.size();
    }
    template <class Tag = AgeGroup>
    AgeGroup size() const
    {
        return AgeGroup(m_v.size());
    }
    auto begin()
    {
        return m_v.begin();
    }
    auto end()
    {
        return m_v.end();
    }

private:
    std::vector<T> m_v;
};

template <class T, class... Tags>
using CustomIndexArray = AgeArray<T>; // only AgeGroup-indexed arrays occur on this path

// ---- TimeSeries (utils/time_series.h:57-620): columns [t, y_0 .. y_{n-1}] ------------------------------------------
template <class FP = ScalarType>
class TimeSeries
{
public:
    using Vector = Eigen::VectorX<FP>;
    explicit TimeSeries(Eigen::Index num_elements = 0)
        : m_n(num_elements)
    {
    }
    TimeSeries(FP t, const Vector& v)
        : m_n(v.size())
    {
        add_time_point(t, v);
    }
    static TimeSeries zero(Eigen::Index num_time_points, Eigen::Index num_elements)
    {
        TimeSeries ts(num_elements);
        for (Eigen::Index i = 0; i < num_time_points; ++i) {
            ts.add_time_point(0.0, Vector::Zero(num_elements));
        }
        return ts;
    }
    Vector& add_time_point(FP t, const Vector& v)
    {
        m_t.push_back(t);
        m_y.push_back(v);
        return m_y.back();
    }
    Vector& add_time_point(FP t)
    {
        return add_time_point(t, Vector::Zero(m_n));
    }
    Eigen::Index get_num_time_points() const
    {
        return Eigen::Index(m_t.size());
    }
    Eigen::Index get_num_elements() const
    {
        return m_n;
    }
    FP& get_time(Eigen::Index i)
    {
        return m_t[size_t(i)];
    }
    FP get_time(Eigen::Index i) const
    {
        return m_t[size_t(i)];
    }
    FP get_last_time() const
    {
        return m_t.back();
    }
    Vector& get_value(Eigen::Index i)
    {
        return m_y[size_t(i)];
    }
    const Vector& get_value(Eigen::Index i) const
    {
        return m_y[size_t(i)];
    }
    Vector& operator[](Eigen::Index i)
    {
        return m_y[size_t(i)];
    }
    const Vector& operator[](Eigen::Index i) const
    {
        return m_y[size_t(i)];
    }
    Vector& get_last_value()
    {
        return m_y.back();
    }
    const Vector& get_last_value() const
    {
        return m_y.back();
    }
    const std::vector<FP>& get_times() const
    {
        return m_t;
    }
    void print_table(const std::vector<std::string>& column_labels = {}, size_t width = 16, size_t precision = 5,
                     std::ostream& out = std::cout) const
    {
        auto w = int(width);
        out << std::string(size_t(w > 4 ? w - 4 : 0), ' ') << "Time";
        for (Eigen::Index k = 0; k < m_n; ++k) {
            std::string l = size_t(k) < column_labels.size() ? column_labels[size_t(k)] : "#" + std::to_string(k + 1);
            out << " " << std::string(l.size() < width ? width - l.size() : 0, ' ') << l;
        }
        out << "\n";
        char buf[64];
        for (size_t i = 0; i < m_t.size(); ++i) {
            std::snprintf(buf, sizeof buf, "%*.*f", w, int(precision), m_t[i]);
            out << buf;
            for (Eigen::Index k = 0; k < m_n; ++k) {
                std::snprintf(buf, sizeof buf, " %*.*f", w, int(precision), m_y[i][k]);
                out << buf;
            }
            out << "\n";
        }
    }

private:
    Eigen::Index m_n;
    std::vector<FP> m_t;
    std::vector<Vector> m_y;
};

// ---- contact matrices with dampings (epidemiology/contact_matrix.h, damping.h) -------------------------------------
template <class FP = ScalarType>
class ContactMatrix
{
public:
    struct RawDamping {
        Eigen::MatrixX<FP> m;
        int level, type;
        FP time;
        // the reference's accessors (epidemiology/damping.h:60-120)
        const Eigen::MatrixX<FP>& get_coeffs() const
        {
            return m;
        }
        DampingLevel get_level() const
        {
            return DampingLevel(level);
        }
        DampingType get_type() const
        {
            return DampingType(type);
        }
        SimulationTime<FP> get_time() const
        {
            return SimulationTime<FP>(time);
        }
    };
    ContactMatrix() = default;
    explicit ContactMatrix(const Eigen::MatrixX<FP>& baseline)
        : m_baseline(baseline)
        , m_minimum(Eigen::MatrixX<FP>::Zero(baseline.rows(), baseline.cols()))
    {
    }
    ContactMatrix(const Eigen::MatrixX<FP>& baseline, const Eigen::MatrixX<FP>& minimum)
        : m_baseline(baseline)
        , m_minimum(minimum)
    {
    }
    const Eigen::MatrixX<FP>& get_baseline() const
    {
        return m_baseline;
    }
    Eigen::MatrixX<FP>& get_baseline()
    {
        return m_baseline;
    }
    const Eigen::MatrixX<FP>& get_minimum() const
    {
        return m_minimum;
    }
    Eigen::MatrixX<FP>& get_minimum()
    {
        return m_minimum;
    }
    void add_damping(FP d, SimulationTime<FP> t)
    {
        add_damping(d, DampingLevel(0), DampingType(0), t);
    }
    void add_damping(FP d, DampingLevel l, DampingType ty, SimulationTime<FP> t)
    {
        add_damping(Eigen::MatrixX<FP>::Constant(m_baseline.rows(), m_baseline.cols(), d), l, ty, t);
    }
    void add_damping(const Eigen::MatrixX<FP>& m, SimulationTime<FP> t)
    {
        add_damping(m, DampingLevel(0), DampingType(0), t);
    }
    // the list is kept sorted by (time, type, level); a damping with the same key replaces the old one
    // (Dampings::add_, epidemiology/damping.h:606-617)
    void add_damping(const Eigen::MatrixX<FP>& m, DampingLevel l, DampingType ty, SimulationTime<FP> t)
    {
        RawDamping d{m, l.value, ty.value, t.get()};
        auto before = [](const RawDamping& a, const RawDamping& b) {
            return std::make_tuple(a.time, a.type, a.level) < std::make_tuple(b.time, b.type, b.level);
        };
        auto pos = std::lower_bound(m_dampings.begin(), m_dampings.end(), d, before);
        if (pos != m_dampings.end() && !before(d, *pos)) {
            *pos = d;
        }
        else {
            m_dampings.insert(pos, d);
        }
    }
    void remove_damping(size_t i)
    {
        m_dampings.erase(m_dampings.begin() + std::ptrdiff_t(i));
    }
    void clear_dampings()
    {
        m_dampings.clear();
    }
    // contact_matrix.h:190-194, damping.h:381-396: baseline - D(t) (baseline - minimum), D(t) from the cache of
    // accumulated dampings (built by the library) with the cosine ramp over the day before a change point
    Eigen::MatrixX<FP> get_matrix_at(SimulationTime<FP> t) const
    {
        const Eigen::Index R = m_baseline.rows(), C = m_baseline.cols();
        const size_t K = m_dampings.size(), RC = size_t(R) * size_t(C);
        std::vector<double> times(K), mats(K * RC), ot(std::max<size_t>(K, 1)), om(std::max<size_t>(K, 1) * RC);
        std::vector<int32_t> levels(K), types(K);
        for (size_t k = 0; k < K; ++k) {
            times[k]  = m_dampings[k].time;
            levels[k] = m_dampings[k].level;
            types[k]  = m_dampings[k].type;
            std::copy(m_dampings[k].m.data(), m_dampings[k].m.data() + RC, mats.begin() + std::ptrdiff_t(k * RC));
        }
        int32_t nout = 0;
        cuda_check(mio_damping_cache_build(int32_t(R), int32_t(C), int32_t(K), times.data(), levels.data(), types.data(),
                                           mats.data(), &nout, ot.data(), om.data()));
        int ub = 0; // first change point after t
        while (ub < nout && !(t.get() < ot[size_t(ub)])) {
            ++ub;
        }
        Eigen::MatrixX<FP> out(R, C);
        const double pi = 3.141592653589793238462643383279502884;
        for (size_t j = 0; j < RC; ++j) {
            const double left = ub > 0 ? om[size_t(ub - 1) * RC + j] : 0.0;
            double d          = left;
            if (ub < nout) { // smoother_cosine over [change point - 1, change point], math/smoother.h:32-55
                const double xr = ot[size_t(ub)], xl = xr - 1.0, right = om[size_t(ub) * RC + j];
                if (t.get() > xl) {
                    d = 0.5 * (left - right) * std::cos(pi / (xr - xl) * (t.get() - xl)) + 0.5 * (left + right);
                }
            }
            out.data()[j] = m_baseline.data()[j] - d * (m_baseline.data()[j] - m_minimum.data()[j]);
        }
        return out;
    }
    const std::vector<RawDamping>& get_dampings() const
    {
        return m_dampings;
    }

private:
    Eigen::MatrixX<FP> m_baseline, m_minimum;
    std::vector<RawDamping> m_dampings;
};

template <class FP = ScalarType>
class ContactMatrixGroup
{
public:
    ContactMatrixGroup(size_t num_matrices = 1, size_t num_groups = 1)
        : m_matrices(num_matrices, ContactMatrix<FP>(Eigen::MatrixX<FP>::Zero(Eigen::Index(num_groups),
                                                                               Eigen::Index(num_groups))))
    {
    }
    ContactMatrix<FP>& operator[](size_t i)
    {
        return m_matrices[i];
    }
    const ContactMatrix<FP>& operator[](size_t i) const
    {
        return m_matrices[i];
    }
    size_t get_num_matrices() const
    {
        return m_matrices.size();
    }
    void clear_dampings()
    {
        for (auto& m : m_matrices) {
            m.clear_dampings();
        }
    }
    // contact_matrix.h:392-400: sum over the members
    Eigen::MatrixX<FP> get_matrix_at(SimulationTime<FP> t) const
    {
        Eigen::MatrixX<FP> sum = m_matrices.at(0).get_matrix_at(t);
        for (size_t l = 1; l < m_matrices.size(); ++l) {
            const Eigen::MatrixX<FP> ml = m_matrices[l].get_matrix_at(t);
            for (Eigen::Index j = 0; j < sum.size(); ++j) {
                sum.data()[j] = sum.data()[j] + ml.data()[j];
            }
        }
        return sum;
    }
    auto begin()
    {
        return m_matrices.begin();
    }
    auto end()
    {
        return m_matrices.end();
    }

private:
    std::vector<ContactMatrix<FP>> m_matrices;
};

// DampingSampling, epidemiology/damping_sampling.h:40-120
template <class FP = ScalarType>
class DampingSampling
{
public:
    DampingSampling(const UncertainValue<FP>& value, DampingLevel level, DampingType type, SimulationTime<FP> time,
                    const std::vector<size_t>& matrices, const Eigen::VectorX<FP>& groups)
        : m_value(value)
        , m_level(level)
        , m_type(type)
        , m_time(time)
        , m_matrices(matrices)
        , m_groups(groups)
    {
    }
    UncertainValue<FP>& get_value()
    {
        return m_value;
    }
    const UncertainValue<FP>& get_value() const
    {
        return m_value;
    }
    DampingLevel get_level() const
    {
        return m_level;
    }
    DampingType get_type() const
    {
        return m_type;
    }
    SimulationTime<FP> get_time() const
    {
        return m_time;
    }
    void set_time(SimulationTime<FP> t)
    {
        m_time = t;
    }
    const std::vector<size_t>& get_matrix_indices() const
    {
        return m_matrices;
    }
    const Eigen::VectorX<FP>& get_group_weights() const
    {
        return m_groups;
    }
    FP draw_sample()
    {
        return m_value.draw_sample();
    }
    mio_damping_sampling to_pod() const
    {
        mio_damping_sampling s{};
        s.value = m_value.to_pod();
        s.level = m_level.value;
        s.type  = m_type.value;
        s.time  = m_time.get();
        for (size_t l : m_matrices) {
            s.matrix_mask |= uint64_t(1) << l;
        }
        for (Eigen::Index i = 0; i < m_groups.size() && i < MIO_MAX_GROUPS; ++i) {
            s.group_weights[i] = m_groups[i];
        }
        return s;
    }

private:
    UncertainValue<FP> m_value;
    DampingLevel m_level;
    DampingType m_type;
    SimulationTime<FP> m_time;
    std::vector<size_t> m_matrices;
    Eigen::VectorX<FP> m_groups;
};

// UncertainContactMatrix, epidemiology/uncertain_matrix.h:42-260
template <class FP = ScalarType>
class UncertainContactMatrix
{
public:
    UncertainContactMatrix(size_t num_matrices = 1, size_t num_groups = 1)
        : m_cont_freq(num_matrices, num_groups)
        , m_school_holiday_damping(UncertainValue<FP>(0.0), DampingLevel(0), DampingType(0), SimulationTime<FP>(0.0),
                                   {}, Eigen::VectorX<FP>::Zero(Eigen::Index(num_groups)))
    {
    }
    operator ContactMatrixGroup<FP>&()
    {
        return m_cont_freq;
    }
    operator const ContactMatrixGroup<FP>&() const
    {
        return m_cont_freq;
    }
    ContactMatrixGroup<FP>& get_cont_freq_mat()
    {
        return m_cont_freq;
    }
    const ContactMatrixGroup<FP>& get_cont_freq_mat() const
    {
        return m_cont_freq;
    }
    std::vector<DampingSampling<FP>>& get_dampings()
    {
        return m_dampings;
    }
    const std::vector<DampingSampling<FP>>& get_dampings() const
    {
        return m_dampings;
    }
    DampingSampling<FP>& get_school_holiday_damping()
    {
        return m_school_holiday_damping;
    }
    const DampingSampling<FP>& get_school_holiday_damping() const
    {
        return m_school_holiday_damping;
    }
    std::vector<std::pair<SimulationTime<FP>, SimulationTime<FP>>>& get_school_holidays()
    {
        return m_school_holidays;
    }
    const std::vector<std::pair<SimulationTime<FP>, SimulationTime<FP>>>& get_school_holidays() const
    {
        return m_school_holidays;
    }
    void draw_sample_dampings()
    {
        for (auto& d : m_dampings) {
            d.draw_sample();
        }
        m_school_holiday_damping.draw_sample();
    }
    // make_matrix, uncertain_matrix.h:178-200
    ContactMatrixGroup<FP> make_matrix(bool accum = false)
    {
        if (!accum) {
            m_cont_freq.clear_dampings();
        }
        auto apply = [&](const DampingSampling<FP>& d, FP value, FP time) {
            const Eigen::Index G = d.get_group_weights().size();
            Eigen::MatrixX<FP> m(G, G);
            for (Eigen::Index i = 0; i < G; ++i) {
                for (Eigen::Index j = 0; j < G; ++j) { // make_contact_damping_matrix, damping_sampling.h:282-289
                    m(i, j) = 1.0 - std::sqrt((1.0 - value * d.get_group_weights()[i]) *
                                              (1.0 - value * d.get_group_weights()[j]));
                }
            }
            for (size_t l : d.get_matrix_indices()) {
                m_cont_freq[l].add_damping(m, d.get_level(), d.get_type(), SimulationTime<FP>(time));
            }
        };
        for (auto& d : m_dampings) {
            apply(d, d.get_value().value(), d.get_time().get());
        }
        for (auto& h : m_school_holidays) {
            apply(m_school_holiday_damping, m_school_holiday_damping.get_value().value(), h.first.get());
            apply(m_school_holiday_damping, 0.0, h.second.get());
        }
        return m_cont_freq;
    }
    ContactMatrixGroup<FP> draw_sample(bool accum = false)
    {
        draw_sample_dampings();
        return make_matrix(accum);
    }

private:
    ContactMatrixGroup<FP> m_cont_freq;
    std::vector<DampingSampling<FP>> m_dampings;
    DampingSampling<FP> m_school_holiday_damping;
    std::vector<std::pair<SimulationTime<FP>, SimulationTime<FP>>> m_school_holidays;
};

// ---- integrator cores (math/integrator.h, adapt_rk.h, stepper_wrapper.h, euler.h): descriptors only ----------------
// (class templates over FP like the reference's; the arithmetic on the device is double)
template <class FP = ScalarType>
class OdeIntegratorCore
{
public:
    virtual ~OdeIntegratorCore() = default;
    virtual mio_integrator_desc desc() const = 0;
    virtual std::unique_ptr<OdeIntegratorCore<FP>> clone() const = 0;
};
using IntegratorCore = OdeIntegratorCore<ScalarType>;
template <class FP = ScalarType>
class EulerIntegratorCore : public OdeIntegratorCore<FP>
{
public:
    mio_integrator_desc desc() const override
    {
        return mio_integrator_desc{MIO_INTEGRATOR_EULER, 0, 0, 0, 0, 0};
    }
    std::unique_ptr<OdeIntegratorCore<FP>> clone() const override
    {
        return std::make_unique<EulerIntegratorCore>(*this);
    }
};
template <class FP, int KIND>
class AdaptiveCoreT : public OdeIntegratorCore<FP>
{
public:
    AdaptiveCoreT(FP abs_tol = 1e-10, FP rel_tol = 1e-5, FP dt_min = std::numeric_limits<ScalarType>::min(),
                  FP dt_max = std::numeric_limits<ScalarType>::max())
        : m_abs(abs_tol)
        , m_rel(rel_tol)
        , m_dt_min(dt_min)
        , m_dt_max(dt_max)
    {
    }
    void set_abs_tolerance(FP v)
    {
        m_abs = v;
    }
    void set_rel_tolerance(FP v)
    {
        m_rel = v;
    }
    void set_dt_min(FP v)
    {
        m_dt_min = v;
    }
    void set_dt_max(FP v)
    {
        m_dt_max = v;
    }
    mio_integrator_desc desc() const override
    {
        return mio_integrator_desc{KIND, 0, m_abs, m_rel, m_dt_min, m_dt_max};
    }
    std::unique_ptr<OdeIntegratorCore<FP>> clone() const override
    {
        return std::make_unique<AdaptiveCoreT>(*this);
    }

private:
    FP m_abs, m_rel, m_dt_min, m_dt_max;
};
template <class FP = ScalarType>
using RKIntegratorCore = AdaptiveCoreT<FP, MIO_INTEGRATOR_RKF45>; // math/adapt_rk.h:145-310
template <class FP = ScalarType>
using DefaultIntegratorCore = AdaptiveCoreT<FP, MIO_INTEGRATOR_CK54>; // compartments/simulation.h:33-34 (boost CK54)

} // namespace mio

#include "npis.hpp"
#include "osecir.hpp"
#include "graph.hpp"
#include "study.hpp"
#include "osecirvvs.hpp"
#include "io.hpp"

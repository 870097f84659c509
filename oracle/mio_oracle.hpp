// mio_oracle.hpp -- CPU ORACLE (test infrastructure, NOT product code).
//
// A literal CPU restatement of the MEmilio Graph-ODE ensemble hot path (SciCompMod/memilio @ 8851d5be), written
// from the reference sources, operation order preserved, no Eigen/Boost. Only tests/, __graft_entry__.smoke() and the
// cpu_baseline / --impl reference legs of bench.py may use it. The CUDA product path never calls into this file.
//
// Pinning status (see DESIGN.md "Oracle"):
//   * single-node osecir trajectories (G=1, G=3), both integrator cores: PINNED by the reference's golden CSVs
//     cpp/tests/data/secihurd-compare.csv, secihurd-compare-cashkarp.csv,
//     pycode/memilio-simulation/tests/data/ode-secihurd-compare.csv, ode-secihurd-ageres-compare.csv
//   * MobilityEdge::apply_mobility: pinned by the exact numbers of TestMobility.edgeApplyMobility
//     (cpp/tests/test_mobility.cpp:154-215); graph-level SECIR trajectories: PARITY UNPINNED by the reference
//     (no golden file exists), the literal restatement below is the arbiter.
//   * threefry2x32: pinned by Random123 known-answer vectors; sampled parameter values: PARITY UNPINNED by the
//     reference's tests (bit-exactness comes from using the same libstdc++ distributions).
//
// Third-party arithmetic restated here because it is not under /root/reference:
//   Boost.Odeint 1.84.0 (pin: cpp/thirdparty/CMakeLists.txt:5) controlled_runge_kutta<runge_kutta_cash_karp54>,
//   default_error_checker, default_step_adjuster;  Random123 v1.14.0 threefry2x32 (20 rounds).
#pragma once

#include <algorithm>
#include <cassert>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <functional>
#include <limits>
#include <memory>
#include <numbers>
#include <numeric>
#include <tuple>
#include <utility>
#include <vector>

namespace orc
{

using Vec = std::vector<double>;

// ---------------------------------------------------------------------------------------------------------------
// cpp/memilio/math/smoother.h:32-55
inline double smoother_cosine(double x, double xleft, double xright, double yleft, double yright)
{
    if (x <= xleft) {
        return yleft;
    }
    if (x >= xright) {
        return yright;
    }
    return 0.5 * (yleft - yright) * std::cos(std::numbers::pi_v<double> / (xright - xleft) * (x - xleft)) +
           0.5 * (yleft + yright);
}

// cpp/memilio/math/floating_point.h:46-63
inline bool floating_point_equal(double v1, double v2, double abs_tol = 0, double rel_tol = DBL_MIN)
{
    double diff = std::abs(v1 - v2);
    return diff <= abs_tol || diff <= std::max(std::abs(v1), std::abs(v2)) * rel_tol;
}

// ---------------------------------------------------------------------------------------------------------------
// Dense row-major matrix; column vectors are r x 1.
struct Mat {
    int r = 0, c = 0;
    Vec a;
    Mat() = default;
    Mat(int r_, int c_, double v = 0.0)
        : r(r_)
        , c(c_)
        , a(size_t(r_) * c_, v)
    {
    }
    double& operator()(int i, int j)
    {
        return a[size_t(i) * c + j];
    }
    double operator()(int i, int j) const
    {
        return a[size_t(i) * c + j];
    }
};

// cpp/memilio/epidemiology/damping.h:276-600 (Dampings<FP, D>): a sorted list of (time, type, level, matrix) and the
// cache of accumulated dampings active from each change point on.
struct Damping {
    double time;
    int level;
    int type;
    Mat m;
};

struct Dampings {
    int rows = 0, cols = 0;
    std::vector<Damping> list;
    std::vector<std::pair<Mat, double>> cache; // (accumulated damping, change time), with the two sentinels

    Dampings() = default;
    Dampings(int r, int c)
        : rows(r)
        , cols(c)
    {
        update_cache();
    }

    // damping.h:606-617 add_: sorted by (time, type, level), equal key replaces (stl_util.h:82-96)
    void add(const Damping& d)
    {
        auto key = [](const Damping& x) {
            return std::make_tuple(x.time, x.type, x.level);
        };
        auto lb = std::lower_bound(list.begin(), list.end(), d, [&](const Damping& a, const Damping& b) {
            return key(a) < key(b);
        });
        if (lb != list.end() && key(*lb) == key(d)) {
            *lb = d;
        }
        else {
            list.insert(lb, d);
        }
        cache.clear();
        update_cache();
    }
    void add(double value, int level, int type, double time)
    {
        add(Damping{time, level, type, Mat(rows, cols, value)});
    }
    void clear()
    {
        list.clear();
        cache.clear();
        update_cache();
    }

    // damping.h:571-604 update_cache, :619-660 update_active_dampings, :545-563 inclusive_exclusive_sum
    void update_cache()
    {
        if (!cache.empty()) {
            return;
        }
        cache.emplace_back(Mat(rows, cols, 0.0), std::numeric_limits<double>::lowest());
        struct Active {
            const Mat* m;
            int level;
            int type;
        };
        std::vector<Active> active_by_type;
        std::vector<std::pair<Mat, int>> sum_by_level;
        for (auto& damping : list) {
            auto it_same = std::find_if(active_by_type.begin(), active_by_type.end(), [&](const Active& a) {
                return a.level == damping.level && a.type == damping.type;
            });
            if (it_same != active_by_type.end()) {
                auto& sum_same_level = *std::find_if(sum_by_level.begin(), sum_by_level.end(), [&](auto& s) {
                    return s.second == damping.level;
                });
                for (size_t k = 0; k < damping.m.a.size(); ++k) {
                    sum_same_level.first.a[k] += damping.m.a[k] - it_same->m->a[k];
                }
                it_same->m = &damping.m;
            }
            else {
                active_by_type.push_back(Active{&damping.m, damping.level, damping.type});
                auto it_level = std::find_if(sum_by_level.begin(), sum_by_level.end(), [&](auto& s) {
                    return s.second == damping.level;
                });
                if (it_level != sum_by_level.end()) {
                    for (size_t k = 0; k < damping.m.a.size(); ++k) {
                        it_level->first.a[k] += damping.m.a[k];
                    }
                }
                else {
                    sum_by_level.emplace_back(damping.m, damping.level);
                }
            }
            // inclusive-exclusive sum over levels: sum = sum + m - sum*m
            Mat combined = sum_by_level.front().first;
            for (size_t l = 1; l < sum_by_level.size(); ++l) {
                const Mat& mb = sum_by_level[l].first;
                for (size_t k = 0; k < combined.a.size(); ++k) {
                    double prod   = combined.a[k] * mb.a[k];
                    combined.a[k] = combined.a[k] + mb.a[k] - prod;
                }
            }
            if (floating_point_equal(damping.time, cache.back().second, 1e-15, 1e-15)) {
                cache.back().first = combined;
            }
            else {
                cache.emplace_back(combined, damping.time);
            }
        }
        cache.emplace_back(cache.back().first, std::numeric_limits<double>::max());
    }

    // install an already evaluated cache (change times + accumulated matrices, without sentinels)
    void set_cache(const std::vector<double>& times, const std::vector<Mat>& mats)
    {
        cache.clear();
        cache.emplace_back(Mat(rows, cols, 0.0), std::numeric_limits<double>::lowest());
        for (size_t k = 0; k < times.size(); ++k) {
            cache.emplace_back(mats[k], times[k]);
        }
        cache.emplace_back(cache.back().first, std::numeric_limits<double>::max());
    }

    // damping.h:381-396 get_matrix_at(t)(i,j): upper_bound on the change times + cosine ramp over the day before
    double at(double t, int i, int j) const
    {
        auto ub = std::upper_bound(cache.begin(), cache.end(), t, [](double tt, const std::pair<Mat, double>& e) {
            return tt < e.second;
        });
        return smoother_cosine(t, ub->second - 1.0, ub->second, (ub - 1)->first(i, j), ub->first(i, j));
    }
};

// cpp/memilio/epidemiology/contact_matrix.h:45-200 DampingMatrixExpression
struct DampingExpr {
    Mat baseline, minimum;
    Dampings dampings;
    DampingExpr() = default;
    DampingExpr(const Mat& b, const Mat& m)
        : baseline(b)
        , minimum(m)
        , dampings(b.r, b.c)
    {
    }
    explicit DampingExpr(const Mat& b)
        : DampingExpr(b, Mat(b.r, b.c, 0.0))
    {
    }
    // contact_matrix.h:190-194
    double at(double t, int i, int j) const
    {
        return baseline(i, j) - dampings.at(t, i, j) * (baseline(i, j) - minimum(i, j));
    }
};

// cpp/memilio/epidemiology/contact_matrix.h:271-420 DampingMatrixExpressionGroup
struct DampingExprGroup {
    std::vector<DampingExpr> m;
    // contact_matrix.h:392-400: accumulate over the members, per element
    double at(double t, int i, int j) const
    {
        double s = 0.0;
        for (auto& e : m) {
            s = s + e.at(t, i, j);
        }
        return s;
    }
    int rows() const
    {
        return m[0].baseline.r;
    }
    int cols() const
    {
        return m[0].baseline.c;
    }
};

// ---------------------------------------------------------------------------------------------------------------
// cpp/memilio/utils/time_series.h:57 -- (t, y) columns
struct TimeSeries {
    Vec times;
    std::vector<Vec> values;
    size_t n() const
    {
        return times.size();
    }
    void add(double t, const Vec& v)
    {
        times.push_back(t);
        values.push_back(v);
    }
    Vec& last()
    {
        return values.back();
    }
    const Vec& last() const
    {
        return values.back();
    }
    double last_time() const
    {
        return times.back();
    }
    void remove(size_t i)
    {
        times.erase(times.begin() + i);
        values.erase(values.begin() + i);
    }
};

using DerivFunction = std::function<void(const Vec& y, double t, Vec& dydt)>;

// cpp/memilio/math/integrator.h:46-130 IntegratorCore
struct IntegratorCore {
    double dt_min, dt_max;
    // statistics (not part of the reference; used for the "step counts" parity criterion)
    mutable int64_t n_rejected = 0;
    mutable int64_t n_forced   = 0;
    IntegratorCore(double mn, double mx)
        : dt_min(mn)
        , dt_max(mx)
    {
    }
    virtual ~IntegratorCore() = default;
    virtual bool step(const DerivFunction& f, const Vec& yt, double& t, double& dt, Vec& ytp1) const = 0;
    virtual std::unique_ptr<IntegratorCore> clone() const                                             = 0;
};

// cpp/memilio/math/euler.h:54-61
struct EulerCore : IntegratorCore {
    EulerCore()
        : IntegratorCore(0.0, 0.0)
    {
    }
    bool step(const DerivFunction& f, const Vec& yt, double& t, double& dt, Vec& ytp1) const override
    {
        f(yt, t, ytp1);
        for (size_t i = 0; i < yt.size(); ++i) {
            ytp1[i] = yt[i] + dt * ytp1[i];
        }
        t += dt;
        return true;
    }
    std::unique_ptr<IntegratorCore> clone() const override
    {
        return std::make_unique<EulerCore>(*this);
    }
};

// cpp/memilio/math/stepper_wrapper.h:59-202 over Boost.Odeint 1.84 controlled_runge_kutta<runge_kutta_cash_karp54>
// (boost/numeric/odeint/stepper/controlled_runge_kutta.hpp, runge_kutta_cash_karp54.hpp,
//  explicit_error_generic_rk.hpp, detail/generic_rk_algorithm.hpp) -- restated from the published algorithm.
struct CashKarp54Core : IntegratorCore {
    double abs_tol, rel_tol;
    CashKarp54Core(double abs_tol_ = 1e-10, double rel_tol_ = 1e-5, double mn = DBL_MIN, double mx = DBL_MAX)
        : IntegratorCore(mn, mx)
        , abs_tol(abs_tol_)
        , rel_tol(rel_tol_)
    {
    }
    std::unique_ptr<IntegratorCore> clone() const override
    {
        return std::make_unique<CashKarp54Core>(*this);
    }

    mutable Vec k1, k2, k3, k4, k5, k6, xtmp, xerr;

    // controlled_runge_kutta<..., explicit_error_stepper_tag>::try_step(system, in, t, out, dt)
    bool try_step(const DerivFunction& f, const Vec& x, double& t, Vec& out, double& dt) const
    {
        const size_t n = x.size();
        for (Vec* v : {&k1, &k2, &k3, &k4, &k5, &k6, &xtmp, &xerr}) {
            v->resize(n);
        }
        // m_dxdt = sys(in, t), evaluated on every attempt
        f(x, t, k1);
        // step_adjuster.check_step_size_limit(dt): less_eq_with_sign(dt, max_dt, dt)
        if (dt_max != 0.0 && !(dt - dt_max <= std::numeric_limits<double>::epsilon())) {
            dt = dt_max;
            return false;
        }
        // runge_kutta_cash_karp54 tableau (rk54_ck_coefficients_a1..a5, b, db, c)
        const double c2 = 1.0 / 5.0, c3 = 3.0 / 10.0, c4 = 3.0 / 5.0, c5 = 1.0, c6 = 7.0 / 8.0;
        const double a21 = 1.0 / 5.0;
        const double a31 = 3.0 / 40.0, a32 = 9.0 / 40.0;
        const double a41 = 3.0 / 10.0, a42 = -9.0 / 10.0, a43 = 6.0 / 5.0;
        const double a51 = -11.0 / 54.0, a52 = 5.0 / 2.0, a53 = -70.0 / 27.0, a54 = 35.0 / 27.0;
        const double a61 = 1631.0 / 55296.0, a62 = 175.0 / 512.0, a63 = 575.0 / 13824.0, a64 = 44275.0 / 110592.0,
                     a65 = 253.0 / 4096.0;
        const double b1 = 37.0 / 378.0, b2 = 0.0, b3 = 250.0 / 621.0, b4 = 125.0 / 594.0, b5 = 0.0,
                     b6 = 512.0 / 1771.0;
        const double db1 = b1 - 2825.0 / 27648.0, db2 = 0.0, db3 = b3 - 18575.0 / 48384.0,
                     db4 = b4 - 13525.0 / 55296.0, db5 = -277.0 / 14336.0, db6 = b6 - 1.0 / 4.0;
        // generic_rk_algorithm: x_tmp = 1*x + (a_i1*dt)*k1 + ... (scale_sumN, left to right)
        for (size_t i = 0; i < n; ++i)
            xtmp[i] = 1.0 * x[i] + (a21 * dt) * k1[i];
        f(xtmp, t + c2 * dt, k2);
        for (size_t i = 0; i < n; ++i)
            xtmp[i] = 1.0 * x[i] + (a31 * dt) * k1[i] + (a32 * dt) * k2[i];
        f(xtmp, t + c3 * dt, k3);
        for (size_t i = 0; i < n; ++i)
            xtmp[i] = 1.0 * x[i] + (a41 * dt) * k1[i] + (a42 * dt) * k2[i] + (a43 * dt) * k3[i];
        f(xtmp, t + c4 * dt, k4);
        for (size_t i = 0; i < n; ++i)
            xtmp[i] = 1.0 * x[i] + (a51 * dt) * k1[i] + (a52 * dt) * k2[i] + (a53 * dt) * k3[i] + (a54 * dt) * k4[i];
        f(xtmp, t + c5 * dt, k5);
        for (size_t i = 0; i < n; ++i)
            xtmp[i] = 1.0 * x[i] + (a61 * dt) * k1[i] + (a62 * dt) * k2[i] + (a63 * dt) * k3[i] + (a64 * dt) * k4[i] +
                      (a65 * dt) * k5[i];
        f(xtmp, t + c6 * dt, k6);
        // out = 1*x + (b_i*dt)*k_i  (5th order solution is propagated)
        for (size_t i = 0; i < n; ++i)
            out[i] = 1.0 * x[i] + (b1 * dt) * k1[i] + (b2 * dt) * k2[i] + (b3 * dt) * k3[i] + (b4 * dt) * k4[i] +
                     (b5 * dt) * k5[i] + (b6 * dt) * k6[i];
        // xerr = (dt*db_i)*k_i   (generic_rk_scale_sum_err)
        for (size_t i = 0; i < n; ++i)
            xerr[i] = (dt * db1) * k1[i] + (dt * db2) * k2[i] + (dt * db3) * k3[i] + (dt * db4) * k4[i] +
                      (dt * db5) * k5[i] + (dt * db6) * k6[i];
        // default_error_checker::error: |xerr| / (eps_abs + eps_rel*(a_x*|x| + a_dxdt*dt*|dxdt|)), max norm
        double max_rel_err = 0.0;
        for (size_t i = 0; i < n; ++i) {
            double e    = std::abs(xerr[i]) / (abs_tol + rel_tol * (1.0 * std::abs(x[i]) + (1.0 * dt) * std::abs(k1[i])));
            max_rel_err = std::max(max_rel_err, e);
        }
        if (max_rel_err > 1.0) {
            // default_step_adjuster::decrease_step(dt, err, error_order = 4)
            dt *= std::max(0.9 * std::pow(max_rel_err, -1.0 / (4.0 - 1.0)), 1.0 / 5.0);
            if (dt_max != 0.0) {
                dt = std::min(dt, dt_max);
            }
            return false;
        }
        t += dt;
        // default_step_adjuster::increase_step(dt, err, stepper_order = 5)
        if (max_rel_err < 0.5) {
            max_rel_err = std::max(std::pow(5.0, -5.0), max_rel_err);
            dt *= 9.0 / 10.0 * std::pow(max_rel_err, -1.0 / 5.0);
            if (dt_max != 0.0) {
                dt = std::min(dt, dt_max);
            }
        }
        return true;
    }

    // stepper_wrapper.h:112-163
    bool step(const DerivFunction& f, const Vec& yt, double& t, double& dt, Vec& ytp1) const override
    {
        bool success     = false;
        bool is_dt_valid = true;
        while (!success && is_dt_valid) {
            if (dt < dt_min) {
                is_dt_valid = false;
                dt          = dt_min;
            }
            success = try_step(f, yt, t, ytp1, dt);
            if (!success) {
                ++n_rejected;
            }
        }
        dt = std::max(dt, dt_min);
        if (!success) {
            t += dt_min;
            ++n_forced;
            --n_rejected; // the last failed attempt is the forced step, not a retry
            return false;
        }
        return true;
    }
};

// cpp/memilio/math/adapt_rk.h:145-310 RKIntegratorCore (Runge-Kutta-Fehlberg 4(5), tableau :71-98, :120-137)
struct RKF45Core : IntegratorCore {
    double abs_tol, rel_tol;
    RKF45Core(double abs_tol_ = 1e-10, double rel_tol_ = 1e-5, double mn = DBL_MIN, double mx = DBL_MAX)
        : IntegratorCore(mn, mx)
        , abs_tol(abs_tol_)
        , rel_tol(rel_tol_)
    {
    }
    std::unique_ptr<IntegratorCore> clone() const override
    {
        return std::make_unique<RKF45Core>(*this);
    }
    mutable std::vector<Vec> kt;
    mutable Vec yt_eval;

    bool step(const DerivFunction& f, const Vec& yt, double& t, double& dt, Vec& ytp1) const override
    {
        static const double tab[5][6] = {{0.25, 0.25, 0, 0, 0, 0},
                                         {3 / 8.0, 3 / 32.0, 9 / 32.0, 0, 0, 0},
                                         {12 / 13.0, 1932 / 2197.0, -7200 / 2197.0, 7296 / 2197.0, 0, 0},
                                         {1.0, 439 / 216.0, -8.0, 3680 / 513.0, -845 / 4104.0, 0},
                                         {0.5, -8 / 27.0, 2.0, -3544 / 2565.0, 1859 / 4104.0, -11 / 40.0}};
        static const double low[6]    = {25 / 216.0, 0.0, 1408 / 2565.0, 2197 / 4104.0, -0.2, 0.0};
        static const double high[6]   = {16 / 135.0, 0.0, 6656 / 12825.0, 28561 / 56430.0, -9 / 50.0, 2 / 55.0};
        const size_t n                = yt.size();
        dt                            = std::min(dt, dt_max);
        bool converged                = false;
        bool dt_is_invalid            = false;
        kt.assign(6, Vec(n));
        yt_eval = yt;
        Vec err(n), eps(n);
        while (!converged && !dt_is_invalid) {
            if (dt < dt_min) {
                dt_is_invalid = true;
                dt            = dt_min;
            }
            f(yt_eval, t, kt[0]);
            for (int i = 1; i < 6; ++i) {
                double t_eval = t;
                t_eval += tab[i - 1][0] * dt;
                ytp1 = yt_eval;
                for (int k = 1; k < i + 1; ++k) {
                    const double w = dt * tab[i - 1][k];
                    for (size_t j = 0; j < n; ++j) {
                        ytp1[j] += w * kt[k - 1][j];
                    }
                }
                f(ytp1, t_eval, kt[i]);
            }
            // ytp1 = yt + dt * (K * low); error = dt * |K * (high - low)|
            for (size_t j = 0; j < n; ++j) {
                double sl = 0.0, se = 0.0;
                for (int k = 0; k < 6; ++k) {
                    sl += kt[k][j] * low[k];
                    se += kt[k][j] * (high[k] - low[k]);
                }
                ytp1[j] = yt_eval[j] + dt * sl;
                err[j]  = dt * std::abs(se);
                eps[j]  = abs_tol + std::abs(ytp1[j]) * rel_tol;
            }
            converged = true;
            for (size_t j = 0; j < n; ++j) {
                converged = converged && (err[j] <= eps[j]);
            }
            if (converged || dt_is_invalid) {
                t += dt;
            }
            else {
                ++n_rejected;
            }
            double min_ratio = std::numeric_limits<double>::infinity();
            for (size_t j = 0; j < n; ++j) {
                double q = eps[j] / err[j];
                // Eigen minCoeff propagates like std::min on non-NaN input
                if (q < min_ratio) {
                    min_ratio = q;
                }
            }
            double dt_new = dt * std::pow(min_ratio, 1. / 5);
            dt_new *= 0.9;
            dt = std::min(dt_new, dt_max);
        }
        dt = std::max(dt, dt_min);
        if (!converged) {
            ++n_forced;
        }
        return converged;
    }
};

// cpp/memilio/math/integrator.h:184-249 SystemIntegrator::advance
struct SystemIntegrator {
    std::unique_ptr<IntegratorCore> core;
    int64_t n_steps = 0;
    explicit SystemIntegrator(std::unique_ptr<IntegratorCore> c)
        : core(std::move(c))
    {
    }
    SystemIntegrator(const SystemIntegrator& o)
        : core(o.core->clone())
        , n_steps(o.n_steps)
    {
    }
    SystemIntegrator& operator=(const SystemIntegrator& o)
    {
        core    = o.core->clone();
        n_steps = o.n_steps;
        return *this;
    }

    Vec& advance(const DerivFunction& f, const double tmax, double& dt, TimeSeries& results)
    {
        const double t0  = results.last_time();
        const double eps = 1e-12; // Limits<double>::zero_tolerance, config.h:72-79
        assert(tmax > t0);
        assert(dt > 0);
        double dt_restore     = 0.0;
        double dt_min_restore = core->dt_min;
        double t              = t0;
        for (size_t i = results.n() - 1; std::fabs((tmax - t) / (tmax - t0)) > eps; ++i) {
            if (dt > tmax - t) {
                dt_restore   = dt;
                dt           = tmax - t;
                core->dt_min = std::min(tmax - t, core->dt_min);
            }
            results.add(0.0, Vec(results.values[i].size()));
            Vec yt = results.values[i]; // results[i] may be reallocated by add(); Eigen::Ref semantic is a view
            core->step(f, yt, t, dt, results.values[i + 1]);
            results.times.back() = t;
            ++n_steps;
        }
        core->dt_min = dt_min_restore;
        dt           = std::max(dt, dt_restore);
        return results.last();
    }
};

// ---------------------------------------------------------------------------------------------------------------
// cpp/models/ode_secir: infection_state.h:33-47, model.h:45-59
namespace osecir
{
enum : int
{
    S = 0,
    E,
    INS,
    INSC,
    ISy,
    ISyC,
    ISev,
    ICr,
    R,
    D,
    NC
};
// flow order of the TypeList (model.h:45-59): (source, target)
static const int FLOW_SRC[15] = {S, E, INS, INS, INSC, INSC, ISy, ISy, ISyC, ISyC, ISev, ISev, ISev, ICr, ICr};
static const int FLOW_DST[15] = {E, INS, ISy, R, ISyC, R, ISev, R, ISev, R, ICr, R, D, D, R};
enum : int
{
    F_S_E = 0,
    F_E_INS,
    F_INS_ISy,
    F_INS_R,
    F_INSC_ISyC,
    F_INSC_R,
    F_ISy_ISev,
    F_ISy_R,
    F_ISyC_ISev,
    F_ISyC_R,
    F_ISev_ICr,
    F_ISev_R,
    F_ISev_D,
    F_ICr_D,
    F_ICr_R,
    NF
};

// cpp/models/ode_secir/parameters.h:47-397
struct Params {
    int G = 1;
    double StartDay                    = 0.0;
    double Seasonality                 = 0.0;
    double ICUCapacity                 = std::numeric_limits<double>::max();
    double TestAndTraceCapacity        = std::numeric_limits<double>::max();
    double TestAndTraceCapacityMaxRisk = 5.0;
    double commuter_nondetection       = 0.0;
    double start_commuter_detection    = 0.0;
    double end_commuter_detection      = 0.0;
    Vec TimeExposed, TimeInfectedNoSymptoms, TimeInfectedSymptoms, TimeInfectedSevere, TimeInfectedCritical,
        TransmissionProbabilityOnContact, RelativeTransmissionNoSymptoms, RecoveredPerInfectedNoSymptoms,
        RiskOfInfectionFromSymptomatic, MaxRiskOfInfectionFromSymptomatic, SeverePerInfectedSymptoms,
        CriticalPerSevere, DeathsPerSevere, DeathsPerCritical;
    DampingExprGroup ContactPatterns;

    explicit Params(int g = 1)
        : G(g)
        , TimeExposed(g, 1.)
        , TimeInfectedNoSymptoms(g, 1.)
        , TimeInfectedSymptoms(g, 1.)
        , TimeInfectedSevere(g, 1.)
        , TimeInfectedCritical(g, 1.)
        , TransmissionProbabilityOnContact(g, 1.)
        , RelativeTransmissionNoSymptoms(g, 1.)
        , RecoveredPerInfectedNoSymptoms(g, 0.)
        , RiskOfInfectionFromSymptomatic(g, 1.)
        , MaxRiskOfInfectionFromSymptomatic(g, 1.)
        , SeverePerInfectedSymptoms(g, 0.)
        , CriticalPerSevere(g, 0.)
        , DeathsPerSevere(g, 0.)
        , DeathsPerCritical(g, 0.)
    {
        ContactPatterns.m.emplace_back(Mat(g, g, 0.0));
    }

    // parameters.h:465-615
    bool apply_constraints()
    {
        const double tol_times = 1e-1;
        bool corrected         = false;
        if (Seasonality < 0.0 || Seasonality > 0.5) {
            Seasonality = 0;
            corrected   = true;
        }
        if (ICUCapacity < 0.0) {
            ICUCapacity = 0;
            corrected   = true;
        }
        if (TestAndTraceCapacity < 0.0) {
            TestAndTraceCapacity = 0;
            corrected            = true;
        }
        if (TestAndTraceCapacityMaxRisk < 0.0) {
            TestAndTraceCapacityMaxRisk = 0;
            corrected                   = true;
        }
        for (int i = 0; i < G; ++i) {
            for (Vec* tv : {&TimeExposed, &TimeInfectedNoSymptoms, &TimeInfectedSymptoms, &TimeInfectedSevere,
                            &TimeInfectedCritical}) {
                if ((*tv)[i] < tol_times) {
                    (*tv)[i]  = tol_times;
                    corrected = true;
                }
            }
            if (TransmissionProbabilityOnContact[i] < 0.0 || TransmissionProbabilityOnContact[i] > 1.0) {
                TransmissionProbabilityOnContact[i] = 0.0;
                corrected                           = true;
            }
            if (RelativeTransmissionNoSymptoms[i] < 0.0) {
                RelativeTransmissionNoSymptoms[i] = 0;
                corrected                         = true;
            }
            for (Vec* pv : {&RecoveredPerInfectedNoSymptoms, &RiskOfInfectionFromSymptomatic,
                            &SeverePerInfectedSymptoms, &CriticalPerSevere, &DeathsPerSevere}) {
                if ((*pv)[i] < 0.0 || (*pv)[i] > 1.0) {
                    (*pv)[i]  = 0;
                    corrected = true;
                }
            }
            if (CriticalPerSevere[i] + DeathsPerSevere[i] > 1.0) {
                DeathsPerSevere[i] = 0;
                corrected          = true;
            }
            if (DeathsPerCritical[i] < 0.0 || DeathsPerCritical[i] > 1.0) {
                DeathsPerCritical[i] = 0;
                corrected            = true;
            }
        }
        return corrected;
    }
};

struct Model {
    Params parameters;
    Vec populations; // flat, idx = age*NC + compartment
    explicit Model(int g = 1)
        : parameters(g)
        , populations(size_t(g) * NC, 0.0)
    {
    }
    int num_compartments() const
    {
        return NC;
    }

    // model.h:82-214
    void get_flows(const Vec& pop, const Vec& y, double t, Vec& flows) const
    {
        const Params& params = parameters;
        const int n_agegroups = params.G;
        double icu_occupancy = 0.0;
        double test_and_trace_required = 0.0;
        for (int i = 0; i < n_agegroups; ++i) {
            test_and_trace_required += (1 - params.RecoveredPerInfectedNoSymptoms[i]) /
                                       params.TimeInfectedNoSymptoms[i] * pop[i * NC + INS];
            icu_occupancy += pop[i * NC + ICr];
        }
        for (int i = 0; i < n_agegroups; ++i) {
            const int Si = i * NC + S, Ei = i * NC + E, INSi = i * NC + INS, INSCi = i * NC + INSC,
                      ISyi = i * NC + ISy, ISyCi = i * NC + ISyC, ISevi = i * NC + ISev, ICri = i * NC + ICr;
            for (int j = 0; j < n_agegroups; ++j) {
                const int Sj = j * NC + S, Ej = j * NC + E, INSj = j * NC + INS, ISyj = j * NC + ISy,
                          ISevj = j * NC + ISev, ICrj = j * NC + ICr, Rj = j * NC + R;
                double riskFromInfectedSymptomatic = smoother_cosine(
                    test_and_trace_required, params.TestAndTraceCapacity,
                    params.TestAndTraceCapacity * params.TestAndTraceCapacityMaxRisk,
                    params.RiskOfInfectionFromSymptomatic[j], params.MaxRiskOfInfectionFromSymptomatic[j]);
                double season_val =
                    (1 + params.Seasonality *
                             std::sin(std::numbers::pi_v<double> * ((params.StartDay + t) / 182.5 + 0.5)));
                double cont_freq_eff = season_val * params.ContactPatterns.at(t, i, j);
                double Nj = pop[Sj] + pop[Ej] + pop[INSj] + pop[ISyj] + pop[ISevj] + pop[ICrj] + pop[Rj];
                const double divNj = (Nj < 1e-12) ? 0.0 : 1.0 / Nj;
                double dummy_S     = y[Si] * cont_freq_eff * divNj * params.TransmissionProbabilityOnContact[i] *
                                 (params.RelativeTransmissionNoSymptoms[j] * pop[INSj] +
                                  riskFromInfectedSymptomatic * pop[ISyj]);
                flows[i * NF + F_S_E] += dummy_S;
            }
            double criticalPerSevereAdjusted =
                smoother_cosine(icu_occupancy, 0.90 * params.ICUCapacity, params.ICUCapacity,
                                params.CriticalPerSevere[i], 0.0);
            double deathsPerSevereAdjusted = params.CriticalPerSevere[i] - criticalPerSevereAdjusted;

            flows[i * NF + F_E_INS] = (1.0 / params.TimeExposed[i]) * y[Ei];
            flows[i * NF + F_INS_ISy] =
                (1 - params.RecoveredPerInfectedNoSymptoms[i]) * (1 / params.TimeInfectedNoSymptoms[i]) * y[INSi];
            flows[i * NF + F_INS_R] =
                params.RecoveredPerInfectedNoSymptoms[i] * (1.0 / params.TimeInfectedNoSymptoms[i]) * y[INSi];
            flows[i * NF + F_INSC_ISyC] = (1.0 - params.RecoveredPerInfectedNoSymptoms[i]) *
                                          (1.0 / params.TimeInfectedNoSymptoms[i]) * y[INSCi];
            flows[i * NF + F_INSC_R] =
                params.RecoveredPerInfectedNoSymptoms[i] * (1.0 / params.TimeInfectedNoSymptoms[i]) * y[INSCi];
            flows[i * NF + F_ISy_ISev] =
                params.SeverePerInfectedSymptoms[i] / params.TimeInfectedSymptoms[i] * y[ISyi];
            flows[i * NF + F_ISy_R] =
                (1 - params.SeverePerInfectedSymptoms[i]) / params.TimeInfectedSymptoms[i] * y[ISyi];
            flows[i * NF + F_ISyC_ISev] =
                params.SeverePerInfectedSymptoms[i] / params.TimeInfectedSymptoms[i] * y[ISyCi];
            flows[i * NF + F_ISyC_R] =
                (1.0 - params.SeverePerInfectedSymptoms[i]) / params.TimeInfectedSymptoms[i] * y[ISyCi];
            flows[i * NF + F_ISev_ICr] = criticalPerSevereAdjusted / params.TimeInfectedSevere[i] * y[ISevi];
            flows[i * NF + F_ISev_R] = (1.0 - params.CriticalPerSevere[i] - params.DeathsPerSevere[i]) /
                                       params.TimeInfectedSevere[i] * y[ISevi];
            flows[i * NF + F_ISev_D] =
                (params.DeathsPerSevere[i] + deathsPerSevereAdjusted) / params.TimeInfectedSevere[i] * y[ISevi];
            flows[i * NF + F_ICr_D] = params.DeathsPerCritical[i] / params.TimeInfectedCritical[i] * y[ICri];
            flows[i * NF + F_ICr_R] =
                (1.0 - params.DeathsPerCritical[i]) / params.TimeInfectedCritical[i] * y[ICri];
        }
    }

    // flow_model.h:112-144,219-235 + compartmental_model.h:104-109
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
    // FlowModel::get_derivatives(flows, dydt), flow_model.h:112-125, and get_initial_flows :146-149
    int num_flows() const
    {
        return parameters.G * NF;
    }
    void derivatives_from_flows(const Vec& flows, Vec& dydt) const
    {
        std::fill(dydt.begin(), dydt.end(), 0.0);
        for (int i = 0; i < parameters.G; ++i) {
            for (int k = 0; k < NF; ++k) {
                dydt[i * NC + FLOW_SRC[k]] -= flows[i * NF + k];
                dydt[i * NC + FLOW_DST[k]] += flows[i * NF + k];
            }
        }
    }

    // model.h:687-715 get_mobility_factors
    Vec get_mobility_factors(double /*t*/, const Vec& y) const
    {
        const Params& p = parameters;
        double test_and_trace_required = 0.0;
        for (int i = 0; i < p.G; ++i) {
            test_and_trace_required +=
                (1 - p.RecoveredPerInfectedNoSymptoms[i]) / p.TimeInfectedNoSymptoms[i] * y[i * NC + INS];
        }
        Vec factors(y.size(), 1.0);
        for (int i = 0; i < p.G; ++i) {
            factors[i * NC + ISy] = smoother_cosine(test_and_trace_required, p.TestAndTraceCapacity,
                                                    p.TestAndTraceCapacity * p.TestAndTraceCapacityMaxRisk,
                                                    p.RiskOfInfectionFromSymptomatic[i],
                                                    p.MaxRiskOfInfectionFromSymptomatic[i]);
        }
        return factors;
    }

    // model.h:717-742 test_commuters (acts on the node's current value and on the mobile population)
    void test_commuters(Vec& last_value, Vec& mobile_population, double time) const
    {
        const Params& p     = parameters;
        double nondetection = 1.0;
        if (time >= p.start_commuter_detection && time < p.end_commuter_detection) {
            nondetection = p.commuter_nondetection;
        }
        for (int i = 0; i < p.G; ++i) {
            const int INSi = i * NC + INS, INSCi = i * NC + INSC, ISyi = i * NC + ISy, ISyCi = i * NC + ISyC;
            last_value[INSi] -= mobile_population[INSi] * (1 - nondetection);
            last_value[INSCi] += mobile_population[INSi] * (1 - nondetection);
            last_value[ISyi] -= mobile_population[ISyi] * (1 - nondetection);
            last_value[ISyCi] += mobile_population[ISyi] * (1 - nondetection);
            mobile_population[ISyi] *= nondetection;
            mobile_population[INSi] *= nondetection;
        }
    }
};
} // namespace osecir

// ---------------------------------------------------------------------------------------------------------------
// cpp/memilio/compartments/simulation.h:41-74 + simulation_base.h:40-176 (+ osecir::Simulation model.h:289-297 when
// no dynamic NPIs are set: passthrough)
template <class M>
struct Simulation {
    M model;
    SystemIntegrator integrator;
    TimeSeries result;
    double dt;
    // mio::FlowSimulation (compartments/flow_simulation.h:37-97, flow_simulation_base.h:40-120) instead of
    // mio::Simulation: the integrands are the cumulative flows, the populations are derived from them
    bool flow_sim = false;
    TimeSeries flow_result;
    Simulation(const M& m, double t0 = 0., double dt_ = 0.1)
        : model(m)
        , integrator(std::make_unique<CashKarp54Core>())
        , dt(dt_)
    {
        result.add(t0, m.populations);
    }
    void set_integrator_core(std::unique_ptr<IntegratorCore> c)
    {
        integrator.core = std::move(c);
    }
    void make_flow_simulation()
    {
        flow_sim = true;
        flow_result.add(result.times[0], Vec(size_t(model.num_flows()), 0.0)); // get_initial_flows: zero
    }
    Vec& advance_flows(double tmax)
    {
        // FlowSimulation::advance, flow_simulation.h:63-92
        assert(flow_result.n() == result.n());
        Vec pop(result.last().size()), diff(size_t(model.num_flows()));
        const size_t base = result.n() - 1; // the last available time point of the population result
        integrator.advance(
            [&](const Vec& flows, double t, Vec& dflows_dt) {
                for (size_t j = 0; j < diff.size(); ++j) {
                    diff[j] = flows[j] - flow_result.values[base][j];
                }
                model.derivatives_from_flows(diff, pop); // note: overwrites values in pop
                for (size_t j = 0; j < pop.size(); ++j) {
                    pop[j] += result.values[base][j];
                }
                std::fill(dflows_dt.begin(), dflows_dt.end(), 0.0);
                model.get_flows(pop, pop, t, dflows_dt);
            },
            tmax, dt, flow_result);
        // compute_population_results, flow_simulation_base.h:84-97
        const size_t last_tp = result.n() - 1;
        for (size_t i = result.n(); i < flow_result.n(); ++i) {
            for (size_t j = 0; j < diff.size(); ++j) {
                diff[j] = flow_result.values[i][j] - flow_result.values[last_tp][j];
            }
            Vec v(result.values[last_tp].size());
            model.derivatives_from_flows(diff, v);
            for (size_t j = 0; j < v.size(); ++j) {
                v[j] += result.values[last_tp][j];
            }
            result.add(flow_result.times[i], v);
        }
        return result.last();
    }
    Vec& advance(double tmax)
    {
        if constexpr (requires(const M& m) { m.num_flows(); }) {
            if (flow_sim) {
                return advance_flows(tmax);
            }
        }
        return integrator.advance(
            [this](const Vec& y, double t, Vec& dydt) {
                model.eval_right_hand_side(y, y, t, dydt);
            },
            tmax, dt, result);
    }
};

// ---------------------------------------------------------------------------------------------------------------
// cpp/memilio/mobility/metapopulation_mobility_instant.h:40-101 SimulationNode
template <class M>
struct SimulationNode {
    Simulation<M> sim;
    Vec last_state;
    double t0;
    SimulationNode(const M& m, double t0_, double dt = 0.1)
        : sim(m, t0_, dt)
        , last_state(sim.result.last())
        , t0(t0_)
    {
    }
    void advance(double t, double dt)
    {
        sim.advance(t + dt);
        last_state = sim.result.last();
    }
};

// MobilityCoefficientGroup (instant.h:103-110): DampingMatrixExpressionGroup of column vectors
struct MobilityEdgeStats {
    int64_t underflows = 0;
};

// cpp/memilio/mobility/metapopulation_mobility_instant.h:222-440, 546-623 MobilityEdge
struct MobilityEdge {
    DampingExprGroup coefficients; // n x 1
    TimeSeries mobile_population;
    Vec return_times;
    bool return_mobile_population = false;
    std::vector<std::vector<size_t>> saved_compartment_indices;
    TimeSeries mobility_results;
    MobilityEdgeStats stats;

    explicit MobilityEdge(const Vec& coeffs, std::vector<std::vector<size_t>> save_indices = {})
        : saved_compartment_indices(std::move(save_indices))
    {
        Mat b(int(coeffs.size()), 1);
        b.a = coeffs;
        coefficients.m.emplace_back(b);
    }

    Vec coeffs_at(double t) const
    {
        const int n = coefficients.rows();
        Vec c(n);
        for (int i = 0; i < n; ++i) {
            c[i] = coefficients.at(t, i, 0);
        }
        return c;
    }

    // instant.h:421-445
    void add_mobility_result_time_point(double t)
    {
        const size_t k = saved_compartment_indices.size();
        if (k > 0) {
            const Vec& last_value = mobile_population.last();
            Vec condensed(k + 1, 0.0);
            for (size_t i = 0; i < k; ++i) {
                double sum = 0.0;
                for (auto index : saved_compartment_indices[i]) {
                    sum += last_value[index];
                }
                condensed[i] = sum;
            }
            condensed[k] = std::accumulate(last_value.begin(), last_value.end(), 0.0);
            mobility_results.add(t, condensed);
        }
    }

    // instant.h:546-623 (dynamic NPIs on edges :549-566 are out of scope: thresholds empty)
    template <class M>
    void apply_mobility(double t, double dt, SimulationNode<M>& node_from, SimulationNode<M>& node_to)
    {
        // returns
        for (int i = int(return_times.size()) - 1; i >= 0; --i) {
            if (return_times[i] <= t) {
                // find_value_reverse(node_to.result, t_depart, 1e-10, 1e-10), time_series.h:905-916
                const double t_dep = mobile_population.times[i];
                const Vec* v0      = nullptr;
                for (int k = int(node_to.sim.result.n()) - 1; k >= 0; --k) {
                    if (floating_point_equal(node_to.sim.result.times[k], t_dep, 1e-10, 1e-10)) {
                        v0 = &node_to.sim.result.values[k];
                        break;
                    }
                }
                assert(v0 && "unexpected error.");
                // calculate_mobility_returns, instant.h:460-471: one explicit Euler step with pop = total
                Vec& mp = mobile_population.values[i];
                {
                    Vec y0 = mp;
                    Vec dydt(mp.size());
                    node_to.sim.model.get_derivatives(*v0, y0, t_dep, dydt);
                    for (size_t j = 0; j < mp.size(); ++j) {
                        mp[j] = y0[j] + dt * dydt[j];
                    }
                }
                Vec& to_last = node_to.sim.result.last();
                Vec remaining(to_last.size());
                for (size_t j = 0; j < remaining.size(); ++j) {
                    remaining[j] = to_last[j] - mp[j];
                }
                const int nc = node_to.sim.model.num_compartments();
                for (int j = 0; j < int(remaining.size()); ++j) {
                    if (remaining[j] < 0) {
                        int group     = j / nc;
                        int max_index = 0;
                        for (int c = 1; c < nc; ++c) {
                            if (remaining[group * nc + c] > remaining[group * nc + max_index]) {
                                max_index = c;
                            }
                        }
                        max_index += group * nc;
                        mp[max_index] -= remaining[j];
                        mp[j] += remaining[j];
                        ++stats.underflows;
                    }
                }
                Vec& from_last = node_from.sim.result.last();
                for (size_t j = 0; j < mp.size(); ++j) {
                    from_last[j] += mp[j];
                    to_last[j] -= mp[j];
                }
                add_mobility_result_time_point(t);
                mobile_population.remove(i);
                return_times.erase(return_times.begin() + i);
            }
        }
        if (!return_mobile_population) {
            Vec c    = coeffs_at(t);
            bool any = false;
            for (double v : c) {
                any = any || (v > 0.0);
            }
            if (any) {
                Vec f = node_from.sim.model.get_mobility_factors(t, node_from.last_state);
                Vec mp(c.size());
                for (size_t j = 0; j < mp.size(); ++j) {
                    mp[j] = node_from.last_state[j] * c[j] * f[j];
                }
                mobile_population.add(t, mp);
                return_times.push_back(t + dt);
                node_from.sim.model.test_commuters(node_from.sim.result.last(), mobile_population.last(), t);
                Vec& to_last   = node_to.sim.result.last();
                Vec& from_last = node_from.sim.result.last();
                const Vec& m   = mobile_population.last();
                for (size_t j = 0; j < m.size(); ++j) {
                    to_last[j] += m[j];
                    from_last[j] -= m[j];
                }
                add_mobility_result_time_point(t);
            }
        }
        return_mobile_population = !return_mobile_population;
    }
};

// cpp/memilio/mobility/graph.h:145-300 (edges kept sorted by (start,end), :233-241) and
// graph_simulation.h:101-120 GraphSimulation::advance
template <class M>
struct GraphSim {
    struct Edge {
        int start, end;
        MobilityEdge property;
    };
    std::vector<SimulationNode<M>> nodes;
    std::vector<Edge> edges;
    double t, dt;

    GraphSim(double t0, double dt_)
        : t(t0)
        , dt(dt_)
    {
    }
    void add_node(const M& m, double t0, double dt_node = 0.1)
    {
        nodes.emplace_back(m, t0, dt_node);
    }
    void add_edge(int s, int e, const MobilityEdge& me)
    {
        Edge ed{s, e, me};
        auto lb = std::lower_bound(edges.begin(), edges.end(), ed, [](const Edge& a, const Edge& b) {
            return a.start == b.start ? a.end < b.end : a.start < b.start;
        });
        if (lb != edges.end() && lb->start == s && lb->end == e) {
            *lb = ed;
        }
        else {
            edges.insert(lb, ed);
        }
    }
    void advance(double t_max)
    {
        double dts = dt;
        while (t < t_max) {
            if (t + dts > t_max) {
                dts = t_max - t;
            }
            for (auto& n : nodes) {
                n.advance(t, dts);
            }
            t += dts;
            for (auto& e : edges) {
                e.property.apply_mobility(t, dts, nodes[e.start], nodes[e.end]);
            }
        }
    }
};

} // namespace orc

// study.cpp -- host side of ParameterStudy::run for the B200 backend: counter based RNG, parameter sampling in run
// order and the damping cache that turns sampled dampings into the flat contact / mobility tables of mio_cuda.h.
//
// Replaces (reference paths): cpp/memilio/compartments/parameter_studies.h:290-367 (run_impl counter logic,
// distribute_runs), cpp/models/ode_secir/parameter_space.h:125-251 (draw_sample*), cpp/memilio/utils/
// random_number_generator.h:195-218,237-263,320-331, utils/parameter_distributions.h (sample()),
// epidemiology/uncertain_matrix.h:165-200, epidemiology/damping_sampling.h:256-307, epidemiology/damping.h:571-660,
// ode_secir/parameters.h:465-615, epidemiology/populations.h:136-142,194-205,264-281.
//
// Sampling has to be bit-identical to the reference, which draws through libstdc++'s <random> distributions from a
// Random123 threefry2x32 engine. The distributions here are the same libstdc++ classes (constructed fresh per draw,
// like DistributionAdapter::operator(), random_number_generator.h:474-483); only the engine is restated.
#include "../../include/mio_cuda.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <random>
#include <string>
#include <tuple>
#include <cstdlib>
#include <thread>
#include <vector>

int mio_set_last_error(int code, const std::string& msg); // capi.cu

namespace
{

// ---- Random123 v1.14.0 threefry2x32, 20 rounds (published algorithm: Salmon et al., SC'11; Skein rotation constants)
inline uint32_t rotl32(uint32_t x, unsigned n)
{
    return (x << n) | (x >> (32u - n));
}

inline void threefry2x32_20(uint32_t c0, uint32_t c1, uint32_t k0, uint32_t k1, uint32_t& o0, uint32_t& o1)
{
    static const unsigned rot[8] = {13, 15, 26, 6, 17, 29, 16, 24};
    const uint32_t ks[3]         = {k0, k1, 0x1BD11BDAu ^ k0 ^ k1};
    uint32_t x0 = c0 + ks[0], x1 = c1 + ks[1];
    for (unsigned r = 0; r < 20; ++r) {
        x0 += x1;
        x1 = rotl32(x1, rot[r % 8]);
        x1 ^= x0;
        if ((r & 3u) == 3u) {
            const unsigned s = (r + 1) / 4;
            x0 += ks[s % 3];
            x1 += ks[(s + 1) % 3] + s;
        }
    }
    o0 = x0;
    o1 = x1;
}

// mio::RandomNumberGenerator: UniformRandomBitGenerator with 64 bit results, key and counter are uint64 split into
// two 32 bit words in memory order (details::to_r123_array / to_uint64, random_number_generator.h:150-167)
struct Engine {
    using result_type = uint64_t;
    uint64_t key = 0, counter = 0;
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
        uint32_t o0, o1;
        threefry2x32_20(uint32_t(counter), uint32_t(counter >> 32), uint32_t(key), uint32_t(key >> 32), o0, o1);
        ++counter;
        return uint64_t(o0) | (uint64_t(o1) << 32);
    }
    // RandomNumberGenerator::seed + seed_rng_key: threefry2x32_key_t::seed(seed_seq) fills the two key words with
    // seed_seq.generate (Random123 array.h seed())
    void seed(const uint32_t* seeds, int n)
    {
        std::seed_seq sq(seeds, seeds + n);
        uint32_t u32[2];
        sq.generate(u32, u32 + 2);
        key = uint64_t(u32[0]) | (uint64_t(u32[1]) << 32);
    }
};

// ---- UncertainValue::draw_sample -------------------------------------------------------------------------------
struct Uncertain {
    mio_uncertain_value d;
    double draw(Engine& rng)
    {
        switch (d.kind) {
        case MIO_DIST_UNIFORM:
            // ParameterDistributionUniform::sample, parameter_distributions.h:480-487
            d.value = std::uniform_real_distribution<double>(d.a, d.b)(rng);
            break;
        case MIO_DIST_NORMAL:
            d.value = draw_normal(rng);
            break;
        case MIO_DIST_LOGNORMAL:
            d.value = std::lognormal_distribution<double>(d.a, d.b)(rng);
            break;
        case MIO_DIST_EXPONENTIAL:
            d.value = std::exponential_distribution<double>(d.a)(rng);
            break;
        case MIO_DIST_CONSTANT:
            d.value = d.a;
            break;
        default:
            break; // no distribution: value unchanged, no draw (uncertain_value.h:181-188)
        }
        return d.value;
    }
    // ParameterDistributionNormal::sample, parameter_distributions.h:320-350 with check_quantiles :226-262
    double draw_normal(Engine& rng)
    {
        if (d.upper == d.lower) {
            return d.lower;
        }
        double& mean = d.a;
        double& sd   = d.b;
        if (mean < d.lower || mean > d.upper) {
            mean = 0.5 * (d.upper - d.lower);
        }
        if (mean + sd * d.quantile > d.upper) {
            sd = (d.upper - mean) / d.quantile;
        }
        if (mean - sd * d.quantile < d.lower) {
            sd = (mean - d.lower) / d.quantile;
        }
        int i         = 0;
        const int retries = 10;
        double x      = std::normal_distribution<double>(mean, sd)(rng);
        while ((x > d.upper || x < d.lower) && i < retries) {
            x = std::normal_distribution<double>(mean, sd)(rng);
            i++;
            if (i == retries) {
                x = x > d.upper ? d.upper : d.lower;
            }
        }
        return x;
    }
};

// ---- Dampings cache (damping.h:545-660) on flat row-major matrices ----------------------------------------------
struct RawDamping {
    double time;
    int level, type;
    std::vector<double> m;
};

bool fp_equal(double a, double b, double abs_tol, double rel_tol)
{
    const double diff = std::fabs(a - b);
    return diff <= abs_tol || diff <= std::max(std::fabs(a), std::fabs(b)) * rel_tol;
}

struct DampingCache {
    size_t size = 0;
    std::vector<RawDamping> list; // sorted by (time, type, level), unique keys (insert_sorted_replace, stl_util.h:82-96)
    std::vector<double> times;            // change times, sentinels omitted
    std::vector<std::vector<double>> acc; // accumulated damping active from times[k] on

    void add(RawDamping d)
    {
        auto key = [](const RawDamping& x) {
            return std::make_tuple(x.time, x.type, x.level);
        };
        auto range = std::equal_range(list.begin(), list.end(), d, [&](const RawDamping& a, const RawDamping& b) {
            return key(a) < key(b);
        });
        if (range.second - range.first == 1) {
            *range.first = std::move(d);
        }
        else {
            list.insert(range.first, std::move(d));
        }
    }

    void update()
    {
        times.clear();
        acc.clear();
        // entry -1 is the zero matrix at lowest(); kept implicit
        struct Active {
            const std::vector<double>* m;
            int level, type;
        };
        struct LevelSum {
            std::vector<double> m;
            int level;
        };
        std::vector<Active> active;
        std::vector<LevelSum> sums;
        double last_time = std::numeric_limits<double>::lowest();
        bool have_entry  = false; // false: the "back" of the cache is the lowest() sentinel
        for (const RawDamping& d : list) {
            auto ia = std::find_if(active.begin(), active.end(), [&](const Active& a) {
                return a.level == d.level && a.type == d.type;
            });
            if (ia != active.end()) {
                auto is = std::find_if(sums.begin(), sums.end(), [&](const LevelSum& s) {
                    return s.level == d.level;
                });
                for (size_t i = 0; i < size; ++i) {
                    is->m[i] += d.m[i] - (*ia->m)[i];
                }
                ia->m = &d.m;
            }
            else {
                active.push_back({&d.m, d.level, d.type});
                auto is = std::find_if(sums.begin(), sums.end(), [&](const LevelSum& s) {
                    return s.level == d.level;
                });
                if (is != sums.end()) {
                    for (size_t i = 0; i < size; ++i) {
                        is->m[i] += d.m[i];
                    }
                }
                else {
                    sums.push_back({d.m, d.level});
                }
            }
            // inclusive_exclusive_sum: sum = sum + b - sum*b over the levels
            std::vector<double> comb = sums.front().m;
            for (size_t l = 1; l < sums.size(); ++l) {
                for (size_t i = 0; i < size; ++i) {
                    const double prod = comb[i] * sums[l].m[i];
                    comb[i]           = comb[i] + sums[l].m[i] - prod;
                }
            }
            if (fp_equal(d.time, last_time, 1e-15, 1e-15)) {
                if (have_entry) {
                    acc.back() = comb;
                }
                else {
                    // replaces the matrix of the lowest() sentinel: cannot happen for finite damping times
                    times.push_back(last_time);
                    acc.push_back(comb);
                    have_entry = true;
                }
            }
            else {
                times.push_back(d.time);
                acc.push_back(comb);
                last_time  = d.time;
                have_entry = true;
            }
        }
    }
};

constexpr int NC_OSECIR = 10;
// InfectionState order of ode_secir/infection_state.h:33-47
enum
{
    cS = 0,
    cE,
    cINS,
    cINSC,
    cISy,
    cISyC,
    cISev,
    cICr,
    cR,
    cD
};

} // namespace

struct mio_study {
    mio_study_desc d;
    std::vector<mio_uncertain_value> header, age, local, pop;
    std::vector<double> baseline, minimum;
    std::vector<mio_damping_sampling> dampings;
    mio_damping_sampling holiday_damping;
    std::vector<int32_t> node_set, holiday_ptr;
    std::vector<double> holiday_periods;
    std::vector<int32_t> node_ctable;
    int num_ctables = 1;
    Engine rng;
};

namespace
{

int64_t ctable_size(int G, int max_loc, int max_pts)
{
    return 1 + max_loc + int64_t(max_loc) * (max_pts + int64_t(2 + max_pts) * G * G);
}

// make_contact_damping_matrix, damping_sampling.h:282-289: d_ij = 1 - sqrt((1 - g_i) * (1 - g_j))
void contact_damping_matrix(const double* g, int G, std::vector<double>& m)
{
    m.resize(size_t(G) * G);
    for (int i = 0; i < G; ++i) {
        for (int j = 0; j < G; ++j) {
            const double prod = (1.0 - g[i]) * (1.0 - g[j]);
            m[size_t(i) * G + j] = 1.0 - std::sqrt(prod);
        }
    }
}

// Parameters::apply_constraints, ode_secir/parameters.h:465-615 on the flat layout
void apply_constraints(double* pset, int G, double* local)
{
    if (pset[MIO_PH_SEASONALITY] < 0.0 || pset[MIO_PH_SEASONALITY] > 0.5) {
        pset[MIO_PH_SEASONALITY] = 0;
    }
    if (local[MIO_NL_ICU_CAPACITY] < 0.0) {
        local[MIO_NL_ICU_CAPACITY] = 0;
    }
    if (local[MIO_NL_TNT_CAPACITY] < 0.0) {
        local[MIO_NL_TNT_CAPACITY] = 0;
    }
    if (pset[MIO_PH_TNT_MAX_RISK] < 0.0) {
        pset[MIO_PH_TNT_MAX_RISK] = 0;
    }
    const double tol_times = 1e-1;
    for (int i = 0; i < G; ++i) {
        double* a = pset + MIO_PSET_HEADER + i * MIO_OSECIR_AGE_STRIDE;
        for (int s : {MIO_PA_TIME_EXPOSED, MIO_PA_TIME_INS, MIO_PA_TIME_ISY, MIO_PA_TIME_ISEV, MIO_PA_TIME_ICR}) {
            if (a[s] < tol_times) {
                a[s] = tol_times;
            }
        }
        if (a[MIO_PA_TRANS_PROB] < 0.0 || a[MIO_PA_TRANS_PROB] > 1.0) {
            a[MIO_PA_TRANS_PROB] = 0.0;
        }
        if (a[MIO_PA_REL_TRANS_NOSYMP] < 0.0) {
            a[MIO_PA_REL_TRANS_NOSYMP] = 0;
        }
        for (int s : {MIO_PA_REC_PER_INS, MIO_PA_RISK_SYMP, MIO_PA_SEV_PER_ISY, MIO_PA_CRIT_PER_SEV,
                      MIO_PA_DEATHS_PER_SEV}) {
            if (a[s] < 0.0 || a[s] > 1.0) {
                a[s] = 0;
            }
        }
        if (a[MIO_PA_CRIT_PER_SEV] + a[MIO_PA_DEATHS_PER_SEV] > 1.0) {
            a[MIO_PA_DEATHS_PER_SEV] = 0;
        }
        if (a[MIO_PA_DEATHS_PER_CRIT] < 0.0 || a[MIO_PA_DEATHS_PER_CRIT] > 1.0) {
            a[MIO_PA_DEATHS_PER_CRIT] = 0;
        }
    }
}

// Parameters::apply_constraints, ode_secirvvs/parameters.h:710-938 on the flat layout (age stride 32)
void apply_constraints_vvs(double* pset, int G, double* local)
{
    if (pset[MIO_PH_SEASONALITY] < 0.0 || pset[MIO_PH_SEASONALITY] > 0.5) {
        pset[MIO_PH_SEASONALITY] = 0;
    }
    if (local[MIO_NL_ICU_CAPACITY] < 0.0) {
        local[MIO_NL_ICU_CAPACITY] = 0;
    }
    if (local[MIO_NL_TNT_CAPACITY] < 0.0) {
        local[MIO_NL_TNT_CAPACITY] = 0;
    }
    if (pset[MIO_PH_TNT_MAX_RISK] < 0.0) { // TestAndTraceCapacityMaxRiskSymptoms
        pset[MIO_PH_TNT_MAX_RISK] = 0;
    }
    if (pset[MIO_PH_TNT_MAX_RISK_NOSYMP] < 0.0) {
        pset[MIO_PH_TNT_MAX_RISK_NOSYMP] = 0;
    }
    const double tol_times = 1e-1;
    for (int i = 0; i < G; ++i) {
        double* a = pset + MIO_PSET_HEADER + i * MIO_OSECIRVVS_AGE_STRIDE;
        for (int s : {MIO_PA_TIME_EXPOSED, MIO_PA_TIME_INS, MIO_PA_TIME_ISY, MIO_PA_TIME_ISEV, MIO_PA_TIME_ICR}) {
            if (a[s] < tol_times) {
                a[s] = tol_times;
            }
        }
        if (a[MIO_PA_TRANS_PROB] < 0.0 || a[MIO_PA_TRANS_PROB] > 1.0) {
            a[MIO_PA_TRANS_PROB] = 0.0;
        }
        if (a[MIO_PA_REL_TRANS_NOSYMP] < 0.0) {
            a[MIO_PA_REL_TRANS_NOSYMP] = 0;
        }
        for (int s : {MIO_PA_REC_PER_INS, MIO_PA_RISK_SYMP, MIO_PA_SEV_PER_ISY, MIO_PA_CRIT_PER_SEV,
                      MIO_PA_DEATHS_PER_SEV}) {
            if (a[s] < 0.0 || a[s] > 1.0) {
                a[s] = 0;
            }
        }
        if (a[MIO_PA_CRIT_PER_SEV] + a[MIO_PA_DEATHS_PER_SEV] > 1.0) {
            a[MIO_PA_DEATHS_PER_SEV] = 0;
        }
        if (a[MIO_PA_DEATHS_PER_CRIT] < 0.0 || a[MIO_PA_DEATHS_PER_CRIT] > 1.0) {
            a[MIO_PA_DEATHS_PER_CRIT] = 0;
        }
        for (int s : {MIO_PV_DAYS_UNTIL_EFFECTIVE_PI, MIO_PV_DAYS_UNTIL_EFFECTIVE_II}) {
            if (a[s] < 0.0) {
                a[s] = 0;
            }
        }
        for (int s : {MIO_PV_REDUC_EXPOSED_PI, MIO_PV_REDUC_EXPOSED_II, MIO_PV_REDUC_INF_SYMP_PI,
                      MIO_PV_REDUC_INF_SYMP_II, MIO_PV_REDUC_SEV_CRIT_DEAD_PI, MIO_PV_REDUC_SEV_CRIT_DEAD_II,
                      MIO_PV_REDUC_TIME_MILD}) {
            if (a[s] <= 0.0 || a[s] > 1.0) {
                a[s] = 1;
            }
        }
        if (a[MIO_PV_INFECTIOUSNESS_NEW_VARIANT] < 0.0) {
            a[MIO_PV_INFECTIOUSNESS_NEW_VARIANT] = 1.0;
        }
        if (a[MIO_PV_VACCINATION_GAP] < 0.0) {
            a[MIO_PV_VACCINATION_GAP] = 0;
        }
    }
}

void write_cache(const DampingCache& c, int max_pts, size_t sz, double* times_out, double* mats_out)
{
    for (size_t k = 0; k < c.times.size(); ++k) {
        times_out[k] = c.times[k];
        std::copy(c.acc[k].begin(), c.acc[k].end(), mats_out + k * sz);
    }
    (void)max_pts;
}

int sample_one_run(const mio_study* s, uint64_t counter, double* pset, double* ctables, double* node_local, double* y0,
                   double* mobtable, uint64_t* draws)
{
    const mio_study_desc& d = s->d;
    const bool vvs = d.model == MIO_MODEL_OSECIRVVS;
    const int NCm  = vvs ? 27 : NC_OSECIR;                                    // compartments per age group
    const int AS   = vvs ? MIO_OSECIRVVS_AGE_STRIDE : MIO_OSECIR_AGE_STRIDE;  // age record stride
    const int G = d.num_groups, N = d.num_nodes, L = d.num_locations, n = G * NCm;
    Engine rng  = s->rng; // private copy: runs are sampled concurrently
    rng.counter = counter;

    // working copies: every run starts from the study's distributions (the example copies the graph per run,
    // cpp/examples/ode_secir_parameter_study_graph.cpp:278-282)
    std::vector<Uncertain> header(MIO_PSET_HEADER), age(size_t(G) * AS);
    for (int k = 0; k < MIO_PSET_HEADER; ++k) {
        header[k].d = s->header[k];
    }
    for (size_t k = 0; k < age.size(); ++k) {
        age[k].d = s->age[k];
    }
    auto A = [&](int i, int slot) -> Uncertain& {
        return age[size_t(i) * AS + slot];
    };

    if (!vvs) {
        // draw_sample_infection, ode_secir/parameter_space.h:155-191
        header[MIO_PH_SEASONALITY].draw(rng);
        A(0, MIO_PA_TIME_EXPOSED).draw(rng);
        A(0, MIO_PA_TIME_INS).draw(rng);
        A(0, MIO_PA_TIME_ISY).draw(rng);
        A(0, MIO_PA_REL_TRANS_NOSYMP).draw(rng);
        A(0, MIO_PA_RISK_SYMP).draw(rng);
        A(0, MIO_PA_MAX_RISK_SYMP).draw(rng);
        for (int i = 0; i < G; ++i) {
            for (int slot : {MIO_PA_TIME_EXPOSED, MIO_PA_TIME_INS, MIO_PA_REL_TRANS_NOSYMP, MIO_PA_RISK_SYMP,
                             MIO_PA_MAX_RISK_SYMP}) {
                A(i, slot) = A(0, slot);
            }
            A(i, MIO_PA_TIME_ISY).draw(rng);
            A(i, MIO_PA_TIME_ISEV).draw(rng);
            A(i, MIO_PA_TIME_ICR).draw(rng);
            A(i, MIO_PA_TRANS_PROB).draw(rng);
            A(i, MIO_PA_REC_PER_INS).draw(rng);
            A(i, MIO_PA_DEATHS_PER_CRIT).draw(rng);
            A(i, MIO_PA_SEV_PER_ISY).draw(rng);
            A(i, MIO_PA_CRIT_PER_SEV).draw(rng);
        }
    }
    else {
        // draw_sample_infection, ode_secirvvs/parameter_space.h:86-141
        static const int shared_slots[] = {MIO_PA_TIME_EXPOSED,           MIO_PA_TIME_INS,
                                           MIO_PA_REL_TRANS_NOSYMP,       MIO_PA_RISK_SYMP,
                                           MIO_PA_MAX_RISK_SYMP,          MIO_PV_REDUC_EXPOSED_PI,
                                           MIO_PV_REDUC_EXPOSED_II,       MIO_PV_REDUC_INF_SYMP_PI,
                                           MIO_PV_REDUC_INF_SYMP_II,      MIO_PV_REDUC_SEV_CRIT_DEAD_PI,
                                           MIO_PV_REDUC_SEV_CRIT_DEAD_II, MIO_PV_REDUC_TIME_MILD};
        header[MIO_PH_SEASONALITY].draw(rng);
        for (int slot : shared_slots) {
            A(0, slot).draw(rng);
        }
        for (int i = 0; i < G; ++i) {
            for (int slot : shared_slots) {
                A(i, slot) = A(0, slot);
            }
            A(i, MIO_PA_TIME_ISY).draw(rng);
            A(i, MIO_PA_TIME_ISEV).draw(rng);
            A(i, MIO_PA_TIME_ICR).draw(rng);
            A(i, MIO_PA_TRANS_PROB).draw(rng);
            A(i, MIO_PA_REC_PER_INS).draw(rng);
            A(i, MIO_PA_DEATHS_PER_CRIT).draw(rng);
            A(i, MIO_PA_SEV_PER_ISY).draw(rng);
            A(i, MIO_PA_CRIT_PER_SEV).draw(rng);
        }
    }
    // shared_contacts.draw_sample_dampings(), uncertain_matrix.h:165-171
    std::vector<double> dvalue(s->dampings.size());
    for (size_t k = 0; k < s->dampings.size(); ++k) {
        Uncertain u{s->dampings[k].value};
        dvalue[k] = u.draw(rng);
    }
    double holiday_value;
    {
        Uncertain u{s->holiday_damping.value};
        holiday_value = u.draw(rng);
    }
    // (DynamicNPIsInfectedSymptoms::draw_sample: no dynamic NPIs in this descriptor)

    for (int k = 0; k < MIO_PSET_HEADER; ++k) {
        pset[k] = header[k].d.value;
    }
    for (int i = 0; i < G; ++i) {
        for (int slot = 0; slot < AS; ++slot) {
            pset[MIO_PSET_HEADER + i * AS + slot] = A(i, slot).d.value;
        }
    }

    // per node: draw_sample_demographics (ode_secir/parameter_space.h:125-148, ode_secirvvs/parameter_space.h:40-78)
    for (int k = 0; k < N; ++k) {
        Uncertain icu{s->local[size_t(k) * 2 + MIO_NL_ICU_CAPACITY]}, tnt{s->local[size_t(k) * 2 + MIO_NL_TNT_CAPACITY]};
        double* loc              = node_local + size_t(k) * 2;
        loc[MIO_NL_ICU_CAPACITY] = icu.draw(rng);
        loc[MIO_NL_TNT_CAPACITY] = tnt.draw(rng);
        double* y                = y0 + size_t(k) * n;
        const mio_uncertain_value* pk = s->pop.data() + size_t(k) * n;
        for (int j = 0; j < n; ++j) {
            y[j] = pk[j].value;
        }
        auto group_total = [&](int i) { // Populations::get_group_total, populations.h:136-142
            double acc = 0.0;
            for (int c = 0; c < NCm; ++c) {
                acc = acc + y[i * NCm + c];
            }
            return acc;
        };
        auto set_difference = [&](int i, int comp, double total) { // populations.h:194-205
            double current = group_total(i);
            current -= y[i * NCm + comp];
            y[i * NCm + comp] = total - current;
        };
        for (int i = 0; i < G; ++i) {
            const double total = group_total(i);
            if (!vvs) {
                for (int c : {cE, cINS, cISy, cISev, cICr, cR}) {
                    Uncertain u{pk[i * NCm + c]};
                    y[i * NCm + c] = u.draw(rng);
                }
                set_difference(i, cS, total);
                set_difference(i, cS, group_total(i));
            }
            else {
                // every compartment but SusceptibleNaive (0) and the three Dead ones (24, 25, 26)
                for (int c = 1; c < 24; ++c) {
                    Uncertain u{pk[i * NCm + c]};
                    y[i * NCm + c] = u.draw(rng);
                }
                y[i * NCm + 0]    = 0;
                const double diff = group_total(i) - total;
                if (diff > 0) {
                    y[i * NCm + 23] -= diff; // SusceptibleImprovedImmunity
                }
                set_difference(i, 0, total);
            }
        }
        // apply_constraints of the node model: parameters (shared set + this node's capacities), populations
        if (vvs) {
            apply_constraints_vvs(pset, G, loc);
        }
        else {
            apply_constraints(pset, G, loc);
        }
        for (int j = 0; j < n; ++j) {
            if (y[j] < 0.0) {
                y[j] = 0.0;
            }
        }
    }

    // make_matrix, uncertain_matrix.h:178-200: sampled dampings (+ school holidays of the table's holiday set)
    const int max_pts  = d.ct_max_pts;
    const int64_t csz  = ctable_size(G, L, max_pts);
    const size_t GG    = size_t(G) * G;
    const int64_t lstr = max_pts + int64_t(2 + max_pts) * G * G;
    std::vector<double> gw(G), mat;
    for (int tbl = 0; tbl < s->num_ctables; ++tbl) {
        double* ct = ctables + tbl * csz;
        std::fill(ct, ct + csz, 0.0);
        ct[0] = L;
        for (int l = 0; l < L; ++l) {
            DampingCache cache;
            cache.size = GG;
            for (size_t k = 0; k < s->dampings.size(); ++k) {
                const mio_damping_sampling& ds = s->dampings[k];
                if (!((ds.matrix_mask >> l) & 1u)) {
                    continue;
                }
                for (int i = 0; i < G; ++i) {
                    gw[i] = dvalue[k] * ds.group_weights[i];
                }
                contact_damping_matrix(gw.data(), G, mat);
                cache.add({ds.time, ds.level, ds.type, mat});
            }
            if (tbl < d.num_holiday_sets) {
                const mio_damping_sampling& hs = s->holiday_damping;
                if ((hs.matrix_mask >> l) & 1u) {
                    for (int h = s->holiday_ptr[tbl]; h < s->holiday_ptr[tbl + 1]; ++h) {
                        for (int i = 0; i < G; ++i) {
                            gw[i] = holiday_value * hs.group_weights[i];
                        }
                        contact_damping_matrix(gw.data(), G, mat);
                        cache.add({s->holiday_periods[size_t(h) * 2], hs.level, hs.type, mat});
                        for (int i = 0; i < G; ++i) {
                            gw[i] = 0.0 * hs.group_weights[i];
                        }
                        contact_damping_matrix(gw.data(), G, mat);
                        cache.add({s->holiday_periods[size_t(h) * 2 + 1], hs.level, hs.type, mat});
                    }
                }
            }
            cache.update();
            if ((int)cache.times.size() > max_pts) {
                return mio_set_last_error(MIO_OUT_OF_RANGE, "sampled contact dampings exceed ct_max_pts");
            }
            ct[1 + l] = double(cache.times.size());
            double* p = ct + 1 + L + l * lstr;
            std::copy(s->baseline.begin() + l * GG, s->baseline.begin() + (l + 1) * GG, p + max_pts);
            std::copy(s->minimum.begin() + l * GG, s->minimum.begin() + (l + 1) * GG, p + max_pts + GG);
            write_cache(cache, max_pts, GG, p, p + max_pts + 2 * GG);
        }
    }

    // edges: apply_dampings(mobility coefficients, shared dampings, make_mobility_damping_vector),
    // parameter_space.h:241-248
    if (d.mob_max_pts > 0 && mobtable != nullptr && vvs) {
        // ode_secirvvs draw_sample copies the edges unchanged (parameter_space.h:205-211): no mobility dampings
        const int64_t msz0 = 1 + d.mob_max_pts + int64_t(d.mob_max_pts) * n;
        std::fill(mobtable, mobtable + msz0, 0.0);
    }
    else if (d.mob_max_pts > 0 && mobtable != nullptr) {
        const int64_t msz = 1 + d.mob_max_pts + int64_t(d.mob_max_pts) * n;
        std::fill(mobtable, mobtable + msz, 0.0);
        DampingCache cache;
        cache.size = size_t(n);
        std::vector<double> vec(n);
        for (size_t k = 0; k < s->dampings.size(); ++k) {
            const mio_damping_sampling& ds = s->dampings[k];
            if (!((ds.matrix_mask >> d.mobility_location) & 1u)) {
                continue;
            }
            for (int j = 0; j < n; ++j) {
                vec[j] = dvalue[k] * ds.group_weights[j / NCm];
            }
            cache.add({ds.time, ds.level, ds.type, vec});
        }
        cache.update();
        if ((int)cache.times.size() > d.mob_max_pts) {
            return mio_set_last_error(MIO_OUT_OF_RANGE, "sampled mobility dampings exceed mob_max_pts");
        }
        mobtable[0] = double(cache.times.size());
        write_cache(cache, d.mob_max_pts, size_t(n), mobtable + 1, mobtable + 1 + d.mob_max_pts);
    }
    if (draws) {
        *draws = rng.counter - counter;
    }
    return MIO_OK;
}

} // namespace

extern "C" {

int mio_study_create(const mio_study_desc* d, mio_study** out)
{
    if (!d || !out) {
        return mio_set_last_error(MIO_INVALID_VALUE, "null argument");
    }
    *out = nullptr;
    if (d->model != MIO_MODEL_OSECIR && d->model != MIO_MODEL_OSECIRVVS) {
        return mio_set_last_error(MIO_INVALID_VALUE, "study sampling: unknown model kind");
    }
    if (d->num_groups < 1 || d->num_groups > MIO_MAX_GROUPS || d->num_nodes < 1 || d->num_locations < 1 ||
        d->num_locations > 64 || d->num_dampings < 0 || d->num_holiday_sets < 0 || d->ct_max_pts < 0 ||
        d->mob_max_pts < 0 || d->num_seeds < 1 || d->num_seeds > 8 || d->mobility_location < 0 ||
        d->mobility_location >= 64) {
        return mio_set_last_error(MIO_INVALID_VALUE, "invalid study descriptor sizes");
    }
    if (!d->header || !d->age_params || !d->node_local || !d->populations || !d->baseline ||
        (d->num_dampings > 0 && !d->dampings) ||
        (d->num_holiday_sets > 0 && (!d->holiday_ptr || !d->holiday_periods || !d->node_holiday_set))) {
        return mio_set_last_error(MIO_INVALID_VALUE, "study descriptor: missing array");
    }
    const bool vvs = d->model == MIO_MODEL_OSECIRVVS;
    const int G = d->num_groups, N = d->num_nodes, L = d->num_locations, n = G * (vvs ? 27 : NC_OSECIR);
    auto* s = new mio_study();
    s->d    = *d;
    s->header.assign(d->header, d->header + MIO_PSET_HEADER);
    s->age.assign(d->age_params,
                  d->age_params + size_t(G) * (vvs ? MIO_OSECIRVVS_AGE_STRIDE : MIO_OSECIR_AGE_STRIDE));
    s->local.assign(d->node_local, d->node_local + size_t(N) * 2);
    s->pop.assign(d->populations, d->populations + size_t(N) * n);
    s->baseline.assign(d->baseline, d->baseline + size_t(L) * G * G);
    if (d->minimum) {
        s->minimum.assign(d->minimum, d->minimum + size_t(L) * G * G);
    }
    else {
        s->minimum.assign(size_t(L) * G * G, 0.0);
    }
    if (d->num_dampings > 0) {
        s->dampings.assign(d->dampings, d->dampings + d->num_dampings);
    }
    if (d->school_holiday_damping) {
        s->holiday_damping = *d->school_holiday_damping;
    }
    else {
        // UncertainContactMatrix ctor, uncertain_matrix.h:53-55: value 0, level 0, type 0, time 0, no matrices
        std::memset(&s->holiday_damping, 0, sizeof(s->holiday_damping));
        s->holiday_damping.value.kind = MIO_DIST_NONE;
    }
    bool any_none = false;
    s->node_ctable.assign(N, 0);
    if (d->num_holiday_sets > 0) {
        s->holiday_ptr.assign(d->holiday_ptr, d->holiday_ptr + d->num_holiday_sets + 1);
        s->holiday_periods.assign(d->holiday_periods, d->holiday_periods + size_t(s->holiday_ptr.back()) * 2);
        s->node_set.assign(d->node_holiday_set, d->node_holiday_set + N);
        for (int k = 0; k < N; ++k) {
            if (s->node_set[k] >= d->num_holiday_sets) {
                delete s;
                return mio_set_last_error(MIO_OUT_OF_RANGE, "node_holiday_set out of range");
            }
            any_none = any_none || s->node_set[k] < 0;
        }
        for (int k = 0; k < N; ++k) {
            s->node_ctable[k] = s->node_set[k] < 0 ? d->num_holiday_sets : s->node_set[k];
        }
        s->num_ctables = d->num_holiday_sets + (any_none ? 1 : 0);
    }
    s->rng.seed(d->seeds, d->num_seeds);
    s->rng.counter         = d->rng_counter;
    s->d.header            = nullptr; // host pointers of the caller are not valid after this call
    s->d.age_params        = nullptr;
    s->d.node_local        = nullptr;
    s->d.populations       = nullptr;
    s->d.baseline          = nullptr;
    s->d.minimum           = nullptr;
    s->d.dampings          = nullptr;
    s->d.school_holiday_damping = nullptr;
    s->d.node_holiday_set  = nullptr;
    s->d.holiday_ptr       = nullptr;
    s->d.holiday_periods   = nullptr;
    *out = s;
    return MIO_OK;
}

void mio_study_destroy(mio_study* s)
{
    delete s;
}

int mio_study_num_ctables(const mio_study* s)
{
    return s ? s->num_ctables : -1;
}

int mio_study_node_ctable(const mio_study* s, int32_t* node_ctable)
{
    if (!s || !node_ctable) {
        return mio_set_last_error(MIO_INVALID_VALUE, "null argument");
    }
    std::copy(s->node_ctable.begin(), s->node_ctable.end(), node_ctable);
    return MIO_OK;
}

int mio_study_sample_runs(mio_study* s, int64_t run0, int32_t count, double* psets, double* ctables,
                          double* node_local, double* y0, double* mobtables, uint64_t* draws)
{
    if (!s || !psets || !ctables || !node_local || !y0) {
        return mio_set_last_error(MIO_INVALID_VALUE, "null argument");
    }
    if (run0 < 0 || count < 0 || uint64_t(run0) + uint64_t(count) > (uint64_t(1) << 32)) {
        return mio_set_last_error(MIO_OUT_OF_RANGE, "run index exceeds the 32 bit subsequence index");
    }
    const mio_study_desc& d = s->d;
    const bool vvs    = d.model == MIO_MODEL_OSECIRVVS;
    const int G = d.num_groups, N = d.num_nodes, n = G * (vvs ? 27 : NC_OSECIR);
    const int64_t psz = MIO_PSET_HEADER + int64_t(vvs ? MIO_OSECIRVVS_AGE_STRIDE : MIO_OSECIR_AGE_STRIDE) * G;
    const int64_t csz = ctable_size(G, d.num_locations, d.ct_max_pts) * s->num_ctables;
    const int64_t msz = d.mob_max_pts > 0 ? 1 + d.mob_max_pts + int64_t(d.mob_max_pts) * n : 0;
    // parameter_studies.h:310-313: counter = m_rng.counter + rng_totalsequence_counter(run_idx, 0). Every run has its
    // own counter subsequence and its own output slices, so the runs are sampled by several host threads; the values
    // are the same bit for bit as in the reference's serial loop.
    auto sample_range = [&](int32_t r0, int32_t r1) -> int {
        for (int32_t r = r0; r < r1; ++r) {
            const uint64_t counter = d.rng_counter + (uint64_t(run0 + r) << 32);
            int rc = sample_one_run(s, counter, psets + r * psz, ctables + r * csz, node_local + size_t(r) * N * 2,
                                    y0 + size_t(r) * N * n, (mobtables && msz) ? mobtables + r * msz : nullptr,
                                    draws ? draws + r : nullptr);
            if (rc != MIO_OK) {
                return rc;
            }
        }
        return MIO_OK;
    };
    int num_threads = 1;
    if (const char* env = std::getenv("MIO_SAMPLE_THREADS")) {
        num_threads = std::max(1, std::atoi(env));
    }
    else {
        num_threads = (int)std::min<unsigned>(16u, std::max(1u, std::thread::hardware_concurrency()));
    }
    num_threads = std::min<int>(num_threads, std::max<int32_t>(1, count / 16));
    if (num_threads <= 1) {
        return sample_range(0, count);
    }
    std::vector<int> rcs(num_threads, MIO_OK);
    std::vector<std::thread> workers;
    for (int t = 0; t < num_threads; ++t) {
        const int32_t r0 = (int32_t)((int64_t)count * t / num_threads);
        const int32_t r1 = (int32_t)((int64_t)count * (t + 1) / num_threads);
        workers.emplace_back([&, t, r0, r1] { rcs[t] = sample_range(r0, r1); });
    }
    for (auto& w : workers) {
        w.join();
    }
    for (int rc : rcs) {
        if (rc != MIO_OK) {
            return mio_set_last_error(rc, "sampling failed in a worker thread (dampings exceed the table sizes?)");
        }
    }
    return MIO_OK;
}

int mio_study_distribute_runs(int64_t num_runs, int32_t num_procs, int32_t rank, int64_t* run0, int64_t* count)
{
    if (num_runs < 0 || num_procs < 1 || rank < 0 || rank >= num_procs || !run0 || !count) {
        return mio_set_last_error(MIO_INVALID_VALUE, "invalid arguments");
    }
    const int64_t local = num_runs / num_procs, rem = num_runs % num_procs;
    *count = local + (rank < rem ? 1 : 0);
    *run0  = int64_t(rank) * local + std::min<int64_t>(rank, rem);
    return MIO_OK;
}

int mio_study_rng_raw(const uint32_t* seeds, int32_t num_seeds, uint64_t counter, int32_t count, uint64_t* out,
                      uint64_t* key_out)
{
    if (!seeds || num_seeds < 1 || count < 0 || (count > 0 && !out)) {
        return mio_set_last_error(MIO_INVALID_VALUE, "invalid arguments");
    }
    Engine e;
    e.seed(seeds, num_seeds);
    e.counter = counter;
    if (key_out) {
        *key_out = e.key;
    }
    for (int i = 0; i < count; ++i) {
        out[i] = e();
    }
    return MIO_OK;
}

int mio_damping_cache_build(int32_t rows, int32_t cols, int32_t K, const double* times, const int32_t* levels,
                            const int32_t* types, const double* mats, int32_t* num_out, double* out_times,
                            double* out_mats)
{
    if (rows < 1 || cols < 1 || K < 0 || !num_out || (K > 0 && (!times || !levels || !types || !mats))) {
        return mio_set_last_error(MIO_INVALID_VALUE, "invalid arguments");
    }
    DampingCache c;
    c.size = size_t(rows) * cols;
    for (int k = 0; k < K; ++k) {
        c.add({times[k], levels[k], types[k],
               std::vector<double>(mats + size_t(k) * c.size, mats + size_t(k + 1) * c.size)});
    }
    c.update();
    *num_out = int32_t(c.times.size());
    if (out_times && out_mats) {
        write_cache(c, K, c.size, out_times, out_mats);
    }
    return MIO_OK;
}

} // extern "C"

// mio_oracle_sampling.hpp -- CPU ORACLE (test infrastructure, NOT product code): ParameterStudy sampling.
//
// Literal restatement of the reference's sampling chain, independent of the product's study.cpp:
//   RandomNumberGenerator        cpp/memilio/utils/random_number_generator.h:195-218,237-263,295-338
//   Random123 v1.14.0 threefry2x32 (20 rounds) -- third party, not under /root/reference; restated from the published
//       Threefish/Threefry definition (rotation constants R_32x2, key-schedule parity 0x1BD11BDA). Pinned by the
//       Random123 known-answer vectors (tests/test_sampling.py).
//   std::seed_seq::generate      ISO C++ [rand.util.seedseq] (restated by hand)
//   std::generate_canonical / uniform_real_distribution / normal_distribution (libstdc++ 13 algorithms, restated by
//       hand: one 64 bit draw per canonical value, Marsaglia polar method without the cached second value because the
//       reference constructs the distribution afresh for every draw, random_number_generator.h:474-483)
//   ParameterDistribution*       cpp/memilio/utils/parameter_distributions.h:159-350,439-490
//   UncertainValue::draw_sample  cpp/memilio/utils/uncertain_value.h:181-188
//   osecir::draw_sample(Graph&)  cpp/models/ode_secir/parameter_space.h:125-251
//   osecirvvs::draw_sample(Graph&) cpp/models/ode_secirvvs/parameter_space.h:40-214, ode_secirvvs/parameters.h:710-938
//   UncertainContactMatrix       cpp/memilio/epidemiology/uncertain_matrix.h:150-200
//   damping_sampling.h:256-307, populations.h:136-205,264-281, ode_secir/parameters.h:465-615
// PARITY UNPINNED by the reference's tests for sampled VALUES (only properties are tested there); this file and the
// product agree bit for bit, and the raw generator is pinned by Random123's KAT.
#pragma once

#include "mio_oracle.hpp"

#include <cstring>
#include <memory>
#include <optional>

namespace orc
{
namespace sampling
{

// ---- threefry2x32, rounds unrolled the way Random123's threefry.h writes them -----------------------------------
struct R123Array2x32 {
    uint32_t v[2];
};

inline uint32_t RotL_32(uint32_t x, unsigned int N)
{
    return (x << (N & 31)) | (x >> ((32 - N) & 31));
}

inline R123Array2x32 threefry2x32(R123Array2x32 in, R123Array2x32 k)
{
    enum
    {
        R_32x2_0_0 = 13,
        R_32x2_1_0 = 15,
        R_32x2_2_0 = 26,
        R_32x2_3_0 = 6,
        R_32x2_4_0 = 17,
        R_32x2_5_0 = 29,
        R_32x2_6_0 = 16,
        R_32x2_7_0 = 24
    };
    uint32_t X0, X1;
    uint32_t ks[3];
    ks[2] = 0x1BD11BDA;
    for (int i = 0; i < 2; i++) {
        ks[i] = k.v[i];
        ks[2] ^= k.v[i];
    }
    X0 = in.v[0] + ks[0];
    X1 = in.v[1] + ks[1];
#define ORC_TF_ROUND(R)                                                                                                \
    X0 += X1;                                                                                                          \
    X1 = RotL_32(X1, R);                                                                                               \
    X1 ^= X0;
    ORC_TF_ROUND(R_32x2_0_0) ORC_TF_ROUND(R_32x2_1_0) ORC_TF_ROUND(R_32x2_2_0) ORC_TF_ROUND(R_32x2_3_0)
    X0 += ks[1];
    X1 += ks[2];
    X1 += 1;
    ORC_TF_ROUND(R_32x2_4_0) ORC_TF_ROUND(R_32x2_5_0) ORC_TF_ROUND(R_32x2_6_0) ORC_TF_ROUND(R_32x2_7_0)
    X0 += ks[2];
    X1 += ks[0];
    X1 += 2;
    ORC_TF_ROUND(R_32x2_0_0) ORC_TF_ROUND(R_32x2_1_0) ORC_TF_ROUND(R_32x2_2_0) ORC_TF_ROUND(R_32x2_3_0)
    X0 += ks[0];
    X1 += ks[1];
    X1 += 3;
    ORC_TF_ROUND(R_32x2_4_0) ORC_TF_ROUND(R_32x2_5_0) ORC_TF_ROUND(R_32x2_6_0) ORC_TF_ROUND(R_32x2_7_0)
    X0 += ks[1];
    X1 += ks[2];
    X1 += 4;
    ORC_TF_ROUND(R_32x2_0_0) ORC_TF_ROUND(R_32x2_1_0) ORC_TF_ROUND(R_32x2_2_0) ORC_TF_ROUND(R_32x2_3_0)
    X0 += ks[2];
    X1 += ks[0];
    X1 += 5;
#undef ORC_TF_ROUND
    return R123Array2x32{{X0, X1}};
}

// ISO C++ [rand.util.seedseq] generate for n = 2 outputs
inline void seed_seq_generate(const std::vector<uint32_t>& v, uint32_t* begin, size_t n)
{
    if (n == 0) {
        return;
    }
    std::fill(begin, begin + n, 0x8b8b8b8bu);
    const size_t s = v.size();
    const size_t t = (n >= 623) ? 11 : (n >= 68) ? 7 : (n >= 39) ? 5 : (n >= 7) ? 3 : (n - 1) / 2;
    const size_t p = (n - t) / 2;
    const size_t q = p + t;
    const size_t m = std::max(s + 1, n);
    auto T         = [](uint32_t x) {
        return x ^ (x >> 27);
    };
    for (size_t k = 0; k < m; ++k) {
        uint32_t r1 = 1664525u * T(begin[k % n] ^ begin[(k + p) % n] ^ begin[(k + n - 1) % n]);
        uint32_t r2 = r1;
        if (k == 0) {
            r2 += uint32_t(s);
        }
        else if (k <= s) {
            r2 += uint32_t(k % n) + v[k - 1];
        }
        else {
            r2 += uint32_t(k % n);
        }
        begin[(k + p) % n] += r1;
        begin[(k + q) % n] += r2;
        begin[k % n] = r2;
    }
    for (size_t k = m; k < m + n; ++k) {
        uint32_t r3 = 1566083941u * T(begin[k % n] + begin[(k + p) % n] + begin[(k + n - 1) % n]);
        uint32_t r4 = r3 - uint32_t(k % n);
        begin[(k + p) % n] ^= r3;
        begin[(k + q) % n] ^= r4;
        begin[k % n] = r4;
    }
}

// mio::RandomNumberGenerator
struct RandomNumberGenerator {
    uint64_t key     = 0;
    uint64_t counter = 0;
    std::vector<uint32_t> seeds;

    void seed(const std::vector<uint32_t>& s)
    {
        seeds = s;
        uint32_t u32[2];
        seed_seq_generate(seeds, u32, 2); // threefry2x32_key_t::seed(seed_seq)
        std::memcpy(&key, u32, sizeof(uint64_t));
    }
    uint64_t operator()()
    {
        R123Array2x32 c, k;
        std::memcpy(c.v, &counter, sizeof(uint64_t));
        std::memcpy(k.v, &key, sizeof(uint64_t));
        R123Array2x32 r = threefry2x32(c, k);
        uint64_t out;
        std::memcpy(&out, r.v, sizeof(uint64_t));
        ++counter;
        return out;
    }
    // libstdc++ std::generate_canonical<double, 53>(urng) for a 64 bit generator: k = 1 draw, R = 2^64
    double canonical()
    {
        const long double r = static_cast<long double>(std::numeric_limits<uint64_t>::max()) - 0.0L + 1.0L;
        double sum          = 0.0;
        double tmp          = 1.0;
        sum += double((*this)() - 0) * tmp;
        tmp *= r;
        double ret = sum / tmp;
        if (ret >= 1.0) {
            ret = std::nextafter(1.0, 0.0);
        }
        return ret;
    }
};

// ---- distributions ------------------------------------------------------------------------------------------------
struct Distribution {
    int kind = 0; // MIO_DIST_* numbering
    double a = 0, b = 0, lower = 0, upper = 0, quantile = 0;

    double sample(RandomNumberGenerator& rng)
    {
        switch (kind) {
        case 1: // std::uniform_real_distribution: (b - a) * canonical + a
            return (b - a) * rng.canonical() + a;
        case 2:
            return sample_normal(rng);
        case 3: // std::lognormal_distribution: exp(s * N(0,1) + m)
            return std::exp(b * std_normal(rng, 0.0, 1.0) + a);
        case 4: // std::exponential_distribution: -log(1 - canonical) / lambda
            return -std::log(1.0 - rng.canonical()) / a;
        default:
            return a;
        }
    }
    // libstdc++ normal_distribution::operator() with no saved value
    static double std_normal(RandomNumberGenerator& rng, double mean, double stddev)
    {
        double x, y, r2;
        do {
            x  = 2.0 * rng.canonical() - 1.0;
            y  = 2.0 * rng.canonical() - 1.0;
            r2 = x * x + y * y;
        } while (r2 > 1.0 || r2 == 0.0);
        const double mult = std::sqrt(-2 * std::log(r2) / r2);
        double ret        = y * mult;
        return ret * stddev + mean;
    }
    // parameter_distributions.h:226-262,320-350
    double sample_normal(RandomNumberGenerator& rng)
    {
        if (upper == lower) {
            return lower;
        }
        double& m_mean         = a;
        double& m_standard_dev = b;
        if (m_mean < lower || m_mean > upper) {
            m_mean = 0.5 * (upper - lower);
        }
        if (m_mean + m_standard_dev * quantile > upper) {
            m_standard_dev = (upper - m_mean) / quantile;
        }
        if (m_mean - m_standard_dev * quantile < lower) {
            m_standard_dev = (m_mean - lower) / quantile;
        }
        int i        = 0;
        int retries  = 10;
        double rnumb = std_normal(rng, m_mean, m_standard_dev);
        while ((rnumb > upper || rnumb < lower) && i < retries) {
            rnumb = std_normal(rng, m_mean, m_standard_dev);
            i++;
            if (i == retries) {
                if (rnumb > upper) {
                    rnumb = upper;
                }
                else {
                    rnumb = lower;
                }
            }
        }
        return rnumb;
    }
};

struct UncertainValue {
    double value = 0.0;
    std::optional<Distribution> dist;
    double draw_sample(RandomNumberGenerator& rng)
    {
        if (dist) {
            value = dist->sample(rng);
        }
        return value;
    }
};

struct DampingSampling {
    UncertainValue value;
    int level = 0, type = 0;
    double time = 0.0;
    std::vector<size_t> matrices;
    Vec groups;
};

// osecir::Parameters restricted to what is sampled, as UncertainValues per age group
struct SecirParamsU {
    int G     = 0;
    int model = 0; // 0: osecir::Parameters, 1: osecirvvs::Parameters
    UncertainValue StartDay, Seasonality, TestAndTraceCapacityMaxRisk, commuter_nondetection, start_commuter_detection,
        end_commuter_detection;
    // osecirvvs: TestAndTraceCapacityMaxRisk above is ...MaxRiskSymptoms
    UncertainValue TestAndTraceCapacityMaxRiskNoSymptoms, StartDayNewVariant;
    UncertainValue ICUCapacity, TestAndTraceCapacity;
    // [tag][age], tag order = MIO_PA_*
    std::vector<std::vector<UncertainValue>> age;
    // UncertainContactMatrix
    std::vector<Mat> baseline, minimum;
    std::vector<DampingSampling> dampings;
    DampingSampling school_holiday_damping;
    std::vector<std::pair<double, double>> school_holidays;
    std::vector<Dampings> cont_freq_dampings; // per location, rebuilt by make_matrix
};

struct NodeModelU {
    SecirParamsU parameters;
    std::vector<UncertainValue> populations; // [G*C]
    int C = 10;                              // compartments per age group: 10 (osecir), 27 (osecirvvs)
};

enum Tag
{
    TimeExposed = 0,
    TimeInfectedNoSymptoms,
    TimeInfectedSymptoms,
    TimeInfectedSevere,
    TimeInfectedCritical,
    TransmissionProbabilityOnContact,
    RelativeTransmissionNoSymptoms,
    RecoveredPerInfectedNoSymptoms,
    RiskOfInfectionFromSymptomatic,
    MaxRiskOfInfectionFromSymptomatic,
    SeverePerInfectedSymptoms,
    CriticalPerSevere,
    DeathsPerSevere,
    DeathsPerCritical,
    NumTags,
    // osecirvvs only (ode_secirvvs/parameters.h:422-625), same order as MIO_PV_*
    ReducExposedPartialImmunity = NumTags,
    ReducExposedImprovedImmunity,
    ReducInfectedSymptomsPartialImmunity,
    ReducInfectedSymptomsImprovedImmunity,
    ReducInfectedSevereCriticalDeadPartialImmunity,
    ReducInfectedSevereCriticalDeadImprovedImmunity,
    ReducTimeInfectedMild,
    InfectiousnessNewVariant,
    VaccinationGap,
    DaysUntilEffectivePartialImmunity,
    DaysUntilEffectiveImprovedImmunity,
    NumTagsVvs
};

// parameter_space.h:155-191
inline void draw_sample_infection(NodeModelU& model, RandomNumberGenerator& rng)
{
    auto& p = model.parameters;
    p.Seasonality.draw_sample(rng);
    p.age[TimeExposed][0].draw_sample(rng);
    p.age[TimeInfectedNoSymptoms][0].draw_sample(rng);
    p.age[TimeInfectedSymptoms][0].draw_sample(rng);
    p.age[RelativeTransmissionNoSymptoms][0].draw_sample(rng);
    p.age[RiskOfInfectionFromSymptomatic][0].draw_sample(rng);
    p.age[MaxRiskOfInfectionFromSymptomatic][0].draw_sample(rng);
    for (int i = 0; i < p.G; i++) {
        p.age[TimeExposed][i]                       = p.age[TimeExposed][0];
        p.age[TimeInfectedNoSymptoms][i]            = p.age[TimeInfectedNoSymptoms][0];
        p.age[RelativeTransmissionNoSymptoms][i]    = p.age[RelativeTransmissionNoSymptoms][0];
        p.age[RiskOfInfectionFromSymptomatic][i]    = p.age[RiskOfInfectionFromSymptomatic][0];
        p.age[MaxRiskOfInfectionFromSymptomatic][i] = p.age[MaxRiskOfInfectionFromSymptomatic][0];
        p.age[TimeInfectedSymptoms][i].draw_sample(rng);
        p.age[TimeInfectedSevere][i].draw_sample(rng);
        p.age[TimeInfectedCritical][i].draw_sample(rng);
        p.age[TransmissionProbabilityOnContact][i].draw_sample(rng);
        p.age[RecoveredPerInfectedNoSymptoms][i].draw_sample(rng);
        p.age[DeathsPerCritical][i].draw_sample(rng);
        p.age[SeverePerInfectedSymptoms][i].draw_sample(rng);
        p.age[CriticalPerSevere][i].draw_sample(rng);
    }
}

inline double get_group_total(const NodeModelU& m, int i)
{
    double acc = 0.0;
    for (int c = 0; c < m.C; ++c) {
        acc = acc + m.populations[size_t(i) * m.C + c].value;
    }
    return acc;
}

inline void set_difference_from_group_total(NodeModelU& m, int i, int comp, double total_group_population)
{
    double current_population = get_group_total(m, i);
    size_t idx                = size_t(i) * m.C + comp;
    current_population -= m.populations[idx].value;
    m.populations[idx].value = total_group_population - current_population;
}

// parameter_space.h:125-148
inline void draw_sample_demographics(NodeModelU& model, RandomNumberGenerator& rng)
{
    model.parameters.ICUCapacity.draw_sample(rng);
    model.parameters.TestAndTraceCapacity.draw_sample(rng);
    for (int i = 0; i < model.parameters.G; i++) {
        double group_total = get_group_total(model, i);
        for (int c : {1, 2, 4, 6, 7, 8}) { // Exposed, InfectedNoSymptoms, InfectedSymptoms, Severe, Critical, Recovered
            model.populations[size_t(i) * 10 + c].draw_sample(rng);
        }
        set_difference_from_group_total(model, i, 0, group_total);
        set_difference_from_group_total(model, i, 0, get_group_total(model, i));
    }
}

// ode_secirvvs/parameter_space.h:86-141
inline void draw_sample_infection_vvs(NodeModelU& model, RandomNumberGenerator& rng)
{
    auto& p = model.parameters;
    p.Seasonality.draw_sample(rng);
    // not age dependent
    p.age[TimeExposed][0].draw_sample(rng);
    p.age[TimeInfectedNoSymptoms][0].draw_sample(rng);
    p.age[RelativeTransmissionNoSymptoms][0].draw_sample(rng);
    p.age[RiskOfInfectionFromSymptomatic][0].draw_sample(rng);
    p.age[MaxRiskOfInfectionFromSymptomatic][0].draw_sample(rng);
    p.age[ReducExposedPartialImmunity][0].draw_sample(rng);
    p.age[ReducExposedImprovedImmunity][0].draw_sample(rng);
    p.age[ReducInfectedSymptomsPartialImmunity][0].draw_sample(rng);
    p.age[ReducInfectedSymptomsImprovedImmunity][0].draw_sample(rng);
    p.age[ReducInfectedSevereCriticalDeadPartialImmunity][0].draw_sample(rng);
    p.age[ReducInfectedSevereCriticalDeadImprovedImmunity][0].draw_sample(rng);
    p.age[ReducTimeInfectedMild][0].draw_sample(rng);
    for (int i = 0; i < p.G; i++) {
        p.age[TimeExposed][i]                                     = p.age[TimeExposed][0];
        p.age[TimeInfectedNoSymptoms][i]                          = p.age[TimeInfectedNoSymptoms][0];
        p.age[RelativeTransmissionNoSymptoms][i]                  = p.age[RelativeTransmissionNoSymptoms][0];
        p.age[RiskOfInfectionFromSymptomatic][i]                  = p.age[RiskOfInfectionFromSymptomatic][0];
        p.age[MaxRiskOfInfectionFromSymptomatic][i]               = p.age[MaxRiskOfInfectionFromSymptomatic][0];
        p.age[ReducExposedPartialImmunity][i]                     = p.age[ReducExposedPartialImmunity][0];
        p.age[ReducExposedImprovedImmunity][i]                    = p.age[ReducExposedImprovedImmunity][0];
        p.age[ReducInfectedSymptomsPartialImmunity][i]            = p.age[ReducInfectedSymptomsPartialImmunity][0];
        p.age[ReducInfectedSymptomsImprovedImmunity][i]           = p.age[ReducInfectedSymptomsImprovedImmunity][0];
        p.age[ReducInfectedSevereCriticalDeadPartialImmunity][i]  = p.age[ReducInfectedSevereCriticalDeadPartialImmunity][0];
        p.age[ReducInfectedSevereCriticalDeadImprovedImmunity][i] = p.age[ReducInfectedSevereCriticalDeadImprovedImmunity][0];
        p.age[ReducTimeInfectedMild][i]                           = p.age[ReducTimeInfectedMild][0];
        // age dependent
        p.age[TimeInfectedSymptoms][i].draw_sample(rng);
        p.age[TimeInfectedSevere][i].draw_sample(rng);
        p.age[TimeInfectedCritical][i].draw_sample(rng);
        p.age[TransmissionProbabilityOnContact][i].draw_sample(rng);
        p.age[RecoveredPerInfectedNoSymptoms][i].draw_sample(rng);
        p.age[DeathsPerCritical][i].draw_sample(rng);
        p.age[SeverePerInfectedSymptoms][i].draw_sample(rng);
        p.age[CriticalPerSevere][i].draw_sample(rng);
    }
}

// ode_secirvvs/parameter_space.h:40-78; InfectionState order ode_secirvvs/infection_state.h:36-67
inline void draw_sample_demographics_vvs(NodeModelU& model, RandomNumberGenerator& rng)
{
    enum
    {
        SusceptibleNaive            = 0,
        SusceptibleImprovedImmunity = 23,
        DeadNaive                   = 24,
        DeadPartialImmunity         = 25,
        DeadImprovedImmunity        = 26,
        Count                       = 27
    };
    model.parameters.ICUCapacity.draw_sample(rng);
    model.parameters.TestAndTraceCapacity.draw_sample(rng);
    for (int i = 0; i < model.parameters.G; i++) {
        double group_total = get_group_total(model, i);
        for (int inf_state = 0; inf_state < Count; ++inf_state) {
            if (inf_state != SusceptibleNaive && inf_state != DeadNaive && inf_state != DeadPartialImmunity &&
                inf_state != DeadImprovedImmunity) {
                model.populations[size_t(i) * Count + inf_state].draw_sample(rng);
            }
        }
        model.populations[size_t(i) * Count + SusceptibleNaive].value = 0;
        double diff = get_group_total(model, i) - group_total;
        if (diff > 0) {
            model.populations[size_t(i) * Count + SusceptibleImprovedImmunity].value -= diff;
        }
        set_difference_from_group_total(model, i, SusceptibleNaive, group_total);
    }
}

// damping_sampling.h:282-289
inline Mat make_contact_damping_matrix(const Vec& groups)
{
    const int G = int(groups.size());
    Mat m(G, G);
    for (int i = 0; i < G; ++i) {
        for (int j = 0; j < G; ++j) {
            m(i, j) = 1.0 - std::sqrt((1.0 - groups[i]) * (1.0 - groups[j]));
        }
    }
    return m;
}

// uncertain_matrix.h:178-200 with damping_sampling.h:256-267
inline void make_matrix(SecirParamsU& p)
{
    const size_t L = p.baseline.size();
    p.cont_freq_dampings.assign(L, Dampings(p.G, p.G)); // clear_dampings
    auto apply = [&](const DampingSampling& d) {
        for (size_t i : d.matrices) {
            Vec g(d.groups.size());
            for (size_t k = 0; k < g.size(); ++k) {
                g[k] = d.value.value * d.groups[k];
            }
            p.cont_freq_dampings[i].add(Damping{d.time, d.level, d.type, make_contact_damping_matrix(g)});
        }
    };
    for (auto& d : p.dampings) {
        apply(d);
    }
    for (auto h : p.school_holidays) {
        auto damping = p.school_holiday_damping;
        damping.time = h.first;
        apply(damping);
        damping.value.value = 0.0;
        damping.time        = h.second;
        apply(damping);
    }
}

// ode_secir/parameters.h:465-615 (values only)
inline void apply_constraints(SecirParamsU& p)
{
    auto clamp0 = [](UncertainValue& v, bool bad) {
        if (bad) {
            v.value = 0;
        }
    };
    clamp0(p.Seasonality, p.Seasonality.value < 0.0 || p.Seasonality.value > 0.5);
    clamp0(p.ICUCapacity, p.ICUCapacity.value < 0.0);
    clamp0(p.TestAndTraceCapacity, p.TestAndTraceCapacity.value < 0.0);
    clamp0(p.TestAndTraceCapacityMaxRisk, p.TestAndTraceCapacityMaxRisk.value < 0.0);
    for (int i = 0; i < p.G; ++i) {
        for (int tag : {TimeExposed, TimeInfectedNoSymptoms, TimeInfectedSymptoms, TimeInfectedSevere,
                        TimeInfectedCritical}) {
            if (p.age[tag][i].value < 1e-1) {
                p.age[tag][i].value = 1e-1;
            }
        }
        auto& tp = p.age[TransmissionProbabilityOnContact][i];
        clamp0(tp, tp.value < 0.0 || tp.value > 1.0);
        auto& rt = p.age[RelativeTransmissionNoSymptoms][i];
        clamp0(rt, rt.value < 0.0);
        for (int tag : {RecoveredPerInfectedNoSymptoms, RiskOfInfectionFromSymptomatic, SeverePerInfectedSymptoms,
                        CriticalPerSevere, DeathsPerSevere}) {
            auto& v = p.age[tag][i];
            clamp0(v, v.value < 0.0 || v.value > 1.0);
        }
        if (p.age[CriticalPerSevere][i].value + p.age[DeathsPerSevere][i].value > 1.0) {
            p.age[DeathsPerSevere][i].value = 0;
        }
        auto& dc = p.age[DeathsPerCritical][i];
        clamp0(dc, dc.value < 0.0 || dc.value > 1.0);
    }
}

// ode_secirvvs/parameters.h:710-938 (values only)
inline void apply_constraints_vvs(SecirParamsU& p)
{
    auto set_if = [](UncertainValue& v, bool bad, double to) {
        if (bad) {
            v.value = to;
        }
    };
    set_if(p.Seasonality, p.Seasonality.value < 0.0 || p.Seasonality.value > 0.5, 0);
    set_if(p.ICUCapacity, p.ICUCapacity.value < 0.0, 0);
    set_if(p.TestAndTraceCapacity, p.TestAndTraceCapacity.value < 0.0, 0);
    set_if(p.TestAndTraceCapacityMaxRisk, p.TestAndTraceCapacityMaxRisk.value < 0.0, 0);
    set_if(p.TestAndTraceCapacityMaxRiskNoSymptoms, p.TestAndTraceCapacityMaxRiskNoSymptoms.value < 0.0, 0);
    const double tol_times = 1e-1;
    for (int i = 0; i < p.G; ++i) {
        auto v = [&](int tag) -> UncertainValue& {
            return p.age[tag][i];
        };
        set_if(v(TimeExposed), v(TimeExposed).value < tol_times, tol_times);
        set_if(v(TimeInfectedNoSymptoms), v(TimeInfectedNoSymptoms).value < tol_times, tol_times);
        set_if(v(TimeInfectedSymptoms), v(TimeInfectedSymptoms).value < tol_times, tol_times);
        set_if(v(TimeInfectedSevere), v(TimeInfectedSevere).value < tol_times, tol_times);
        set_if(v(TimeInfectedCritical), v(TimeInfectedCritical).value < tol_times, tol_times);
        set_if(v(TransmissionProbabilityOnContact),
               v(TransmissionProbabilityOnContact).value < 0.0 || v(TransmissionProbabilityOnContact).value > 1.0, 0.0);
        set_if(v(RelativeTransmissionNoSymptoms), v(RelativeTransmissionNoSymptoms).value < 0.0, 0);
        set_if(v(RecoveredPerInfectedNoSymptoms),
               v(RecoveredPerInfectedNoSymptoms).value < 0.0 || v(RecoveredPerInfectedNoSymptoms).value > 1.0, 0);
        set_if(v(RiskOfInfectionFromSymptomatic),
               v(RiskOfInfectionFromSymptomatic).value < 0.0 || v(RiskOfInfectionFromSymptomatic).value > 1.0, 0);
        set_if(v(SeverePerInfectedSymptoms),
               v(SeverePerInfectedSymptoms).value < 0.0 || v(SeverePerInfectedSymptoms).value > 1.0, 0);
        set_if(v(CriticalPerSevere), v(CriticalPerSevere).value < 0.0 || v(CriticalPerSevere).value > 1.0, 0);
        set_if(v(DeathsPerSevere), v(DeathsPerSevere).value < 0.0 || v(DeathsPerSevere).value > 1.0, 0);
        set_if(v(DeathsPerSevere), v(CriticalPerSevere).value + v(DeathsPerSevere).value > 1.0, 0);
        set_if(v(DeathsPerCritical), v(DeathsPerCritical).value < 0.0 || v(DeathsPerCritical).value > 1.0, 0);
        set_if(v(DaysUntilEffectivePartialImmunity), v(DaysUntilEffectivePartialImmunity).value < 0.0, 0);
        set_if(v(DaysUntilEffectiveImprovedImmunity), v(DaysUntilEffectiveImprovedImmunity).value < 0.0, 0);
        for (int tag : {ReducExposedPartialImmunity, ReducExposedImprovedImmunity, ReducInfectedSymptomsPartialImmunity,
                        ReducInfectedSymptomsImprovedImmunity, ReducInfectedSevereCriticalDeadPartialImmunity,
                        ReducInfectedSevereCriticalDeadImprovedImmunity, ReducTimeInfectedMild}) {
            set_if(v(tag), v(tag).value <= 0.0 || v(tag).value > 1.0, 1);
        }
        set_if(v(InfectiousnessNewVariant), v(InfectiousnessNewVariant).value < 0.0, 1.0);
        set_if(v(VaccinationGap), v(VaccinationGap).value < 0.0, 0);
    }
}

struct EdgeU {
    int start, end;
    Dampings coeff_dampings; // of the MobilityCoefficients member that carries the baseline
};

struct GraphU {
    std::vector<NodeModelU> nodes;
    int num_edges_proxy = 1; // all edges get the same dampings: one representative is enough
    int n               = 0;
    int mobility_location = 0;
};

// parameter_space.h:207-251; returns the sampled graph (nodes) and the mobility coefficient dampings of the edges
inline std::pair<std::vector<NodeModelU>, Dampings> draw_sample(GraphU& graph, RandomNumberGenerator& rng)
{
    std::vector<NodeModelU> sampled;
    auto& shared_params_model = graph.nodes[0];
    const bool vvs            = shared_params_model.parameters.model == 1;
    if (vvs) {
        draw_sample_infection_vvs(shared_params_model, rng);
    }
    else {
        draw_sample_infection(shared_params_model, rng);
    }
    auto& shared_contacts = shared_params_model.parameters;
    for (auto& d : shared_contacts.dampings) { // draw_sample_dampings
        d.value.draw_sample(rng);
    }
    shared_contacts.school_holiday_damping.value.draw_sample(rng);
    // shared_dynamic_npis.draw_sample(): none configured

    for (size_t k = 0; k < graph.nodes.size(); ++k) {
        auto& node_model = graph.nodes[k];
        if (vvs) {
            draw_sample_demographics_vvs(node_model, rng);
        }
        else {
            draw_sample_demographics(node_model, rng);
        }
        auto local_icu_capacity = node_model.parameters.ICUCapacity;
        auto local_tnt_capacity = node_model.parameters.TestAndTraceCapacity;
        auto local_holidays     = node_model.parameters.school_holidays;
        if (k != 0) {
            node_model.parameters = shared_params_model.parameters;
        }
        node_model.parameters.ICUCapacity          = local_icu_capacity;
        node_model.parameters.TestAndTraceCapacity = local_tnt_capacity;
        node_model.parameters.school_holidays      = local_holidays;
        make_matrix(node_model.parameters);
        if (vvs) {
            apply_constraints_vvs(node_model.parameters);
        }
        else {
            apply_constraints(node_model.parameters);
        }
        for (auto& v : node_model.populations) { // Populations::apply_constraints
            if (v.value < 0.0) {
                v.value = 0.0;
            }
        }
        sampled.push_back(node_model);
    }
    // edges: apply_dampings(coefficients, shared dampings, make_mobility_damping_vector); the osecirvvs draw_sample
    // copies the edges unchanged (ode_secirvvs/parameter_space.h:205-211)
    Dampings mob(graph.n, 1);
    for (auto& d : shared_contacts.dampings) {
        if (vvs) {
            break;
        }
        for (size_t i : d.matrices) {
            if (int(i) != graph.mobility_location) {
                continue; // other group members have a zero baseline (set_edges, graph.h:420-440)
            }
            Mat v(graph.n, 1);
            const int num_compartments = graph.n / int(d.groups.size());
            for (int j = 0; j < graph.n; ++j) {
                v.a[j] = (d.value.value * d.groups[size_t(j / num_compartments)]);
            }
            mob.add(Damping{d.time, d.level, d.type, v});
        }
    }
    return {sampled, mob};
}

} // namespace sampling
} // namespace orc

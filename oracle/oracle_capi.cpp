// oracle_capi.cpp -- C entry points of the CPU ORACLE (test infrastructure, NOT product code; see mio_oracle.hpp).
// They take the same flat descriptors as the CUDA C-ABI (include/mio_cuda.h) so that tests feed identical bytes to
// both sides. Loaded with ctypes by tests/, __graft_entry__.smoke() and the cpu_baseline legs of bench.py only.
#include "mio_oracle.hpp"
#include "mio_oracle_sampling.hpp"
#include "mio_oracle_vvs.hpp"
#include "mio_oracle_analyze.hpp"
#include "mio_oracle_seir.hpp"
#include "mio_oracle_npis.hpp"
#include "../include/mio_cuda.h"

#include <cstring>

using namespace orc;

namespace
{

int ncomp_of(int model)
{
    return model == MIO_MODEL_OSECIR ? 10 : 27;
}

int64_t pset_size(int model, int G)
{
    return MIO_PSET_HEADER + int64_t(model == MIO_MODEL_OSECIR ? MIO_OSECIR_AGE_STRIDE : MIO_OSECIRVVS_AGE_STRIDE) * G;
}

int64_t ctable_size(int G, int max_loc, int max_pts)
{
    return 1 + max_loc + int64_t(max_loc) * (max_pts + int64_t(2 + max_pts) * G * G);
}

DampingExprGroup contacts_from_flat(const double* ct, int G, int max_loc, int max_pts)
{
    DampingExprGroup g;
    const int L          = int(ct[0]);
    const int64_t stride = max_pts + int64_t(2 + max_pts) * G * G;
    for (int l = 0; l < L; ++l) {
        const int K      = int(ct[1 + l]);
        const double* p  = ct + 1 + max_loc + l * stride;
        const double* tm = p;
        const double* B  = p + max_pts;
        const double* M  = B + G * G;
        const double* D  = M + G * G;
        Mat b(G, G), m(G, G);
        std::copy(B, B + G * G, b.a.begin());
        std::copy(M, M + G * G, m.a.begin());
        DampingExpr e(b, m);
        std::vector<double> times(tm, tm + K);
        std::vector<Mat> mats;
        for (int k = 0; k < K; ++k) {
            Mat d(G, G);
            std::copy(D + int64_t(k) * G * G, D + int64_t(k + 1) * G * G, d.a.begin());
            mats.push_back(d);
        }
        e.dampings.set_cache(times, mats);
        g.m.push_back(e);
    }
    return g;
}

osecir::Model osecir_from_flat(int G, const double* pset, const double* local, const double* ct, int max_loc,
                               int max_pts, const double* y0)
{
    osecir::Model m(G);
    auto& p                       = m.parameters;
    p.StartDay                    = pset[MIO_PH_START_DAY];
    p.Seasonality                 = pset[MIO_PH_SEASONALITY];
    p.TestAndTraceCapacityMaxRisk = pset[MIO_PH_TNT_MAX_RISK];
    p.commuter_nondetection       = pset[MIO_PH_COMMUTER_NONDETECTION];
    p.start_commuter_detection    = pset[MIO_PH_START_COMMUTER_DETECTION];
    p.end_commuter_detection      = pset[MIO_PH_END_COMMUTER_DETECTION];
    p.ICUCapacity                 = local[MIO_NL_ICU_CAPACITY];
    p.TestAndTraceCapacity        = local[MIO_NL_TNT_CAPACITY];
    for (int i = 0; i < G; ++i) {
        const double* a                        = pset + MIO_PSET_HEADER + i * MIO_OSECIR_AGE_STRIDE;
        p.TimeExposed[i]                       = a[MIO_PA_TIME_EXPOSED];
        p.TimeInfectedNoSymptoms[i]            = a[MIO_PA_TIME_INS];
        p.TimeInfectedSymptoms[i]              = a[MIO_PA_TIME_ISY];
        p.TimeInfectedSevere[i]                = a[MIO_PA_TIME_ISEV];
        p.TimeInfectedCritical[i]              = a[MIO_PA_TIME_ICR];
        p.TransmissionProbabilityOnContact[i]  = a[MIO_PA_TRANS_PROB];
        p.RelativeTransmissionNoSymptoms[i]    = a[MIO_PA_REL_TRANS_NOSYMP];
        p.RecoveredPerInfectedNoSymptoms[i]    = a[MIO_PA_REC_PER_INS];
        p.RiskOfInfectionFromSymptomatic[i]    = a[MIO_PA_RISK_SYMP];
        p.MaxRiskOfInfectionFromSymptomatic[i] = a[MIO_PA_MAX_RISK_SYMP];
        p.SeverePerInfectedSymptoms[i]         = a[MIO_PA_SEV_PER_ISY];
        p.CriticalPerSevere[i]                 = a[MIO_PA_CRIT_PER_SEV];
        p.DeathsPerSevere[i]                   = a[MIO_PA_DEATHS_PER_SEV];
        p.DeathsPerCritical[i]                 = a[MIO_PA_DEATHS_PER_CRIT];
    }
    p.ContactPatterns = contacts_from_flat(ct, G, max_loc, max_pts);
    std::copy(y0, y0 + size_t(G) * osecir::NC, m.populations.begin());
    return m;
}

osecirvvs::Model osecirvvs_from_flat(int G, const double* pset, const double* local, const double* ct, int max_loc,
                                     int max_pts, const double* y0, int vacc_days, const double* vacc_partial,
                                     const double* vacc_full)
{
    osecirvvs::Model m(G);
    auto& p                                 = m.parameters;
    p.StartDay                              = pset[MIO_PH_START_DAY];
    p.Seasonality                           = pset[MIO_PH_SEASONALITY];
    p.TestAndTraceCapacityMaxRiskSymptoms   = pset[MIO_PH_TNT_MAX_RISK];
    p.TestAndTraceCapacityMaxRiskNoSymptoms = pset[MIO_PH_TNT_MAX_RISK_NOSYMP];
    p.StartDayNewVariant                    = pset[MIO_PH_START_DAY_NEW_VARIANT];
    p.commuter_nondetection                 = pset[MIO_PH_COMMUTER_NONDETECTION];
    p.start_commuter_detection              = pset[MIO_PH_START_COMMUTER_DETECTION];
    p.end_commuter_detection                = pset[MIO_PH_END_COMMUTER_DETECTION];
    p.ICUCapacity                           = local[MIO_NL_ICU_CAPACITY];
    p.TestAndTraceCapacity                  = local[MIO_NL_TNT_CAPACITY];
    for (int i = 0; i < G; ++i) {
        const double* a                        = pset + MIO_PSET_HEADER + i * MIO_OSECIRVVS_AGE_STRIDE;
        p.TimeExposed[i]                       = a[MIO_PA_TIME_EXPOSED];
        p.TimeInfectedNoSymptoms[i]            = a[MIO_PA_TIME_INS];
        p.TimeInfectedSymptoms[i]              = a[MIO_PA_TIME_ISY];
        p.TimeInfectedSevere[i]                = a[MIO_PA_TIME_ISEV];
        p.TimeInfectedCritical[i]              = a[MIO_PA_TIME_ICR];
        p.TransmissionProbabilityOnContact[i]  = a[MIO_PA_TRANS_PROB];
        p.RelativeTransmissionNoSymptoms[i]    = a[MIO_PA_REL_TRANS_NOSYMP];
        p.RecoveredPerInfectedNoSymptoms[i]    = a[MIO_PA_REC_PER_INS];
        p.RiskOfInfectionFromSymptomatic[i]    = a[MIO_PA_RISK_SYMP];
        p.MaxRiskOfInfectionFromSymptomatic[i] = a[MIO_PA_MAX_RISK_SYMP];
        p.SeverePerInfectedSymptoms[i]         = a[MIO_PA_SEV_PER_ISY];
        p.CriticalPerSevere[i]                 = a[MIO_PA_CRIT_PER_SEV];
        p.DeathsPerSevere[i]                   = a[MIO_PA_DEATHS_PER_SEV];
        p.DeathsPerCritical[i]                 = a[MIO_PA_DEATHS_PER_CRIT];
        p.ReducExposedPartialImmunity[i]       = a[MIO_PV_REDUC_EXPOSED_PI];
        p.ReducExposedImprovedImmunity[i]      = a[MIO_PV_REDUC_EXPOSED_II];
        p.ReducInfectedSymptomsPartialImmunity[i]  = a[MIO_PV_REDUC_INF_SYMP_PI];
        p.ReducInfectedSymptomsImprovedImmunity[i] = a[MIO_PV_REDUC_INF_SYMP_II];
        p.ReducInfectedSevereCriticalDeadPartialImmunity[i]  = a[MIO_PV_REDUC_SEV_CRIT_DEAD_PI];
        p.ReducInfectedSevereCriticalDeadImprovedImmunity[i] = a[MIO_PV_REDUC_SEV_CRIT_DEAD_II];
        p.ReducTimeInfectedMild[i]             = a[MIO_PV_REDUC_TIME_MILD];
        p.InfectiousnessNewVariant[i]          = a[MIO_PV_INFECTIOUSNESS_NEW_VARIANT];
        p.VaccinationGap[i]                    = a[MIO_PV_VACCINATION_GAP];
        p.DaysUntilEffectivePartialImmunity[i] = a[MIO_PV_DAYS_UNTIL_EFFECTIVE_PI];
        p.DaysUntilEffectiveImprovedImmunity[i] = a[MIO_PV_DAYS_UNTIL_EFFECTIVE_II];
        if (vacc_days > 0 && vacc_partial && vacc_full) {
            p.DailyPartialVaccinations[i].assign(vacc_partial + int64_t(i) * vacc_days,
                                                 vacc_partial + int64_t(i + 1) * vacc_days);
            p.DailyFullVaccinations[i].assign(vacc_full + int64_t(i) * vacc_days,
                                              vacc_full + int64_t(i + 1) * vacc_days);
        }
    }
    p.ContactPatterns = contacts_from_flat(ct, G, max_loc, max_pts);
    std::copy(y0, y0 + size_t(G) * osecirvvs::NC, m.populations.begin());
    return m;
}

std::unique_ptr<IntegratorCore> make_core(const mio_integrator_desc& d)
{
    switch (d.kind) {
    case MIO_INTEGRATOR_CK54:
        return std::make_unique<CashKarp54Core>(d.abs_tol, d.rel_tol, d.dt_min, d.dt_max);
    case MIO_INTEGRATOR_RKF45:
        return std::make_unique<RKF45Core>(d.abs_tol, d.rel_tol, d.dt_min, d.dt_max);
    default:
        return std::make_unique<EulerCore>();
    }
}

} // namespace

extern "C" {

// Dampings::update_cache on K raw dampings of shape rows x cols; returns the number of cache entries (without the
// two sentinels) and writes change times and accumulated matrices.
int orc_dampings_cache(int rows, int cols, int K, const double* times, const int* levels, const int* types,
                       const double* mats, double* out_times, double* out_mats)
{
    Dampings d(rows, cols);
    const int sz = rows * cols;
    for (int k = 0; k < K; ++k) {
        Damping dm{times[k], levels[k], types[k], Mat(rows, cols)};
        std::copy(mats + int64_t(k) * sz, mats + int64_t(k + 1) * sz, dm.m.a.begin());
        d.add(dm);
    }
    int n = int(d.cache.size()) - 2;
    for (int k = 0; k < n; ++k) {
        out_times[k] = d.cache[k + 1].second;
        std::copy(d.cache[k + 1].first.a.begin(), d.cache[k + 1].first.a.end(), out_mats + int64_t(k) * sz);
    }
    return n;
}

// ContactMatrixGroup::get_matrix_at(t) from a flat contact table
void orc_contact_matrix_at(const double* ct, int G, int max_loc, int max_pts, double t, double* out)
{
    auto g = contacts_from_flat(ct, G, max_loc, max_pts);
    for (int i = 0; i < G; ++i) {
        for (int j = 0; j < G; ++j) {
            out[i * G + j] = g.at(t, i, j);
        }
    }
}

// osecir right-hand side (get_derivatives with pop, y distinct)
void orc_osecir_rhs(int G, const double* pset, const double* local, const double* ct, int max_loc, int max_pts,
                    const double* pop, const double* y, double t, double* dydt)
{
    Vec zero(size_t(G) * osecir::NC, 0.0);
    auto m = osecir_from_flat(G, pset, local, ct, max_loc, max_pts, zero.data());
    Vec p(pop, pop + zero.size()), yy(y, y + zero.size()), d(zero.size());
    m.get_derivatives(p, yy, t, d);
    std::copy(d.begin(), d.end(), dydt);
}

// mio::simulate (compartments/simulation.h) for one node. Returns the number of time points (all of them are counted,
// at most max_pts are written); counters = {accepted, rejected, forced}.
int orc_simulate(int model, int G, const double* pset, const double* local, const double* ct, int max_loc,
                 int max_pts_ct, const double* y0, double t0, double tmax, double dt,
                 const mio_integrator_desc* integ, int max_pts, double* out_t, double* out_y, double* out_dt,
                 int64_t* counters)
{
    if (model != MIO_MODEL_OSECIR) {
        return -1;
    }
    auto m = osecir_from_flat(G, pset, local, ct, max_loc, max_pts_ct, y0);
    Simulation<osecir::Model> sim(m, t0, dt);
    sim.set_integrator_core(make_core(*integ));
    sim.advance(tmax);
    const int n  = int(sim.result.n());
    const int ne = G * ncomp_of(model);
    for (int k = 0; k < std::min(n, max_pts); ++k) {
        out_t[k] = sim.result.times[k];
        std::copy(sim.result.values[k].begin(), sim.result.values[k].end(), out_y + int64_t(k) * ne);
    }
    if (out_dt) {
        *out_dt = sim.dt;
    }
    if (counters) {
        counters[0] = sim.integrator.n_steps;
        counters[1] = sim.integrator.core->n_rejected;
        counters[2] = sim.integrator.core->n_forced;
    }
    return n;
}

// mio::FlowSimulation<osecir::Model> for one node, advanced to tmaxs[0], tmaxs[1], ... in turn (one advance call
// each, flow_simulation.h:63-92): out_f[npts][G*15] the cumulative flows at every time point, out_y the derived
// populations (simulate_flows, flow_simulation.h:99-120, is the case of one advance call)
int orc_simulate_flows_stepwise(int G, const double* pset, const double* local, const double* ct, int max_loc,
                                int max_pts_ct, const double* y0, double t0, int num_t, const double* tmaxs, double dt,
                                const mio_integrator_desc* integ, int max_pts, double* out_t, double* out_y,
                                double* out_f, double* out_dt, int64_t* counters)
{
    auto m = osecir_from_flat(G, pset, local, ct, max_loc, max_pts_ct, y0);
    Simulation<osecir::Model> sim(m, t0, dt);
    sim.set_integrator_core(make_core(*integ));
    sim.make_flow_simulation();
    for (int k = 0; k < num_t; ++k) {
        sim.advance(tmaxs[k]);
    }
    const int n = int(sim.result.n());
    for (int k = 0; k < std::min(n, max_pts); ++k) {
        out_t[k] = sim.result.times[k];
        std::copy(sim.result.values[k].begin(), sim.result.values[k].end(), out_y + int64_t(k) * G * osecir::NC);
        std::copy(sim.flow_result.values[k].begin(), sim.flow_result.values[k].end(),
                  out_f + int64_t(k) * G * osecir::NF);
    }
    if (out_dt) {
        *out_dt = sim.dt;
    }
    if (counters) {
        counters[0] = sim.integrator.n_steps;
        counters[1] = sim.integrator.core->n_rejected;
        counters[2] = sim.integrator.core->n_forced;
    }
    return n;
}
int orc_simulate_flows(int G, const double* pset, const double* local, const double* ct, int max_loc, int max_pts_ct,
                       const double* y0, double t0, double tmax, double dt, const mio_integrator_desc* integ,
                       int max_pts, double* out_t, double* out_y, double* out_f, double* out_dt, int64_t* counters)
{
    return orc_simulate_flows_stepwise(G, pset, local, ct, max_loc, max_pts_ct, y0, t0, 1, &tmax, dt, integ, max_pts,
                                       out_t, out_y, out_f, out_dt, counters);
}

// ---- dynamic NPIs (mio_oracle_npis.hpp) ----------------------------------------------------------------------------
// Flat description of DynamicNPIs<FP> (epidemiology/dynamic_npis.h:40-120): per threshold ONE list of DampingSamplings.
struct orc_npi_desc {
    int32_t num_thresholds;
    const double* thresholds;  // [num_thresholds]
    const int32_t* npi_ptr;    // [num_thresholds + 1] into the DampingSampling arrays
    const double* npi_value;   // [total]
    const int32_t* npi_level;  // [total]
    const int32_t* npi_type;   // [total]
    const int32_t* idx_ptr;    // [total + 1] into npi_matrix_idx
    const int32_t* npi_matrix_idx;
    const double* npi_weights; // [total][num_weights]
    int32_t num_weights;
    int32_t reserved;
    double duration, delay, base, directive_begin, directive_end;
};

static DynamicNPIs npis_from_flat(const orc_npi_desc* d)
{
    DynamicNPIs n;
    n.duration        = d->duration;
    n.delay           = d->delay;
    n.base            = d->base;
    n.directive_begin = d->directive_begin;
    n.directive_end   = d->directive_end;
    for (int t = 0; t < d->num_thresholds; ++t) {
        std::vector<DampingSamplingO> list;
        for (int k = d->npi_ptr[t]; k < d->npi_ptr[t + 1]; ++k) {
            DampingSamplingO ds;
            ds.value = d->npi_value[k];
            ds.level = d->npi_level[k];
            ds.type  = d->npi_type[k];
            ds.matrix_indices.assign(d->npi_matrix_idx + d->idx_ptr[k], d->npi_matrix_idx + d->idx_ptr[k + 1]);
            ds.group_weights.assign(d->npi_weights + size_t(k) * d->num_weights,
                                    d->npi_weights + size_t(k + 1) * d->num_weights);
            list.push_back(ds);
        }
        n.set_threshold(d->thresholds[t], list);
    }
    return n;
}

// a group of column vector damping expressions (rows x 1) with constant-valued dampings given as lists
static DampingExprGroup group_from_lists(int num_mats, int rows, const int32_t* counts, const double* times,
                                         const int32_t* levels, const int32_t* types, const double* values)
{
    DampingExprGroup g;
    int k = 0;
    for (int m = 0; m < num_mats; ++m) {
        g.m.emplace_back(Mat(rows, 1, 1.0));
        for (int i = 0; i < counts[m]; ++i, ++k) {
            g.m.back().dampings.add(Damping{times[k], levels[k], types[k], Mat(rows, 1, values[k])});
        }
    }
    return g;
}

// implement_dynamic_npis (dynamic_npis.h:335-395) on a group of rows x 1 expressions, make_matrix = identity (the
// make_mask of cpp/tests/test_dynamic_npis.cpp). Output per matrix: count, then (time, level, type, coefficients[rows]).
int orc_npi_implement(int num_mats, int rows, const int32_t* counts, const double* times, const int32_t* levels,
                      const int32_t* types, const double* values, int num_calls, const orc_npi_desc* calls,
                      const double* begins, const double* ends, int32_t* out_counts, double* out_times,
                      int32_t* out_levels, int32_t* out_types, double* out_coeffs, int capacity)
{
    auto g = group_from_lists(num_mats, rows, counts, times, levels, types, values);
    for (int c = 0; c < num_calls; ++c) {
        auto n = npis_from_flat(&calls[c]);
        implement_dynamic_npis(g, n.thresholds.at(0).second, begins[c], ends[c], [rows](const Vec& w) {
            Mat m(rows, 1);
            for (int i = 0; i < rows; ++i) {
                m(i, 0) = w[size_t(i)];
            }
            return m;
        });
    }
    int k = 0;
    for (int m = 0; m < num_mats; ++m) {
        out_counts[m] = int32_t(g.m[size_t(m)].dampings.list.size());
        for (auto& d : g.m[size_t(m)].dampings.list) {
            if (k >= capacity) {
                return -1;
            }
            out_times[k]  = d.time;
            out_levels[k] = d.level;
            out_types[k]  = d.type;
            std::copy(d.m.a.begin(), d.m.a.end(), out_coeffs + size_t(k) * rows);
            ++k;
        }
    }
    return k;
}

// get_active_damping / get_damping_indices (dynamic_npis.h:270-306) on one 1 x 1 expression
double orc_npi_active_damping(int count, const double* times, const int32_t* levels, const int32_t* types,
                              const double* values, int lvl, int type, double t)
{
    int32_t c = count;
    auto g    = group_from_lists(1, 1, &c, times, levels, types, values);
    return get_active_damping(g.m[0], lvl, type, t)(0, 0);
}
int orc_npi_damping_indices(int count, const double* times, const int32_t* levels, const int32_t* types,
                            const double* values, int lvl, int type, double begin, double end, int32_t* out)
{
    int32_t c = count;
    auto g    = group_from_lists(1, 1, &c, times, levels, types, values);
    auto idx  = get_damping_indices(g.m[0], lvl, type, begin, end);
    for (size_t i = 0; i < idx.size(); ++i) {
        out[i] = int32_t(idx[i]);
    }
    return int(idx.size());
}

// osecir::Simulation with DynamicNPIsInfectedSymptoms (model.h:289-367) for one node. The contact locations are
// given by their baselines in `ct` (no change points there) plus the raw list of predefined dampings
// (st_*: location, time, level, type, matrix [G*G]). frozen: the MockSimulation of the reference's tests.
// Outputs: the trajectory like orc_simulate; the contact matrix at the query times q_t, out_cm[nq][G*G]; the number
// of dampings of every location afterwards, out_ndamp[max_loc]; returns the number of time points.
int orc_npi_simulate(int G, const double* pset, const double* local, const double* ct, int max_loc, int max_pts_ct,
                     const double* y0, double t0, int num_t, const double* tmaxs, double dt,
                     const mio_integrator_desc* integ, const orc_npi_desc* npi, int frozen, int num_static,
                     const int32_t* st_loc, const double* st_time, const int32_t* st_level, const int32_t* st_type,
                     const double* st_mat, int nq, const double* q_t, double* out_cm, int32_t* out_ndamp, int max_pts,
                     double* out_t, double* out_y, int64_t* counters)
{
    auto m = osecir_from_flat(G, pset, local, ct, max_loc, max_pts_ct, y0);
    auto& cp = m.parameters.ContactPatterns;
    for (auto& e : cp.m) {
        e.dampings.clear();
    }
    for (int k = 0; k < num_static; ++k) {
        Mat dm(G, G);
        std::copy(st_mat + size_t(k) * G * G, st_mat + size_t(k + 1) * G * G, dm.a.begin());
        cp.m.at(size_t(st_loc[k])).dampings.add(Damping{st_time[k], st_level[k], st_type[k], dm});
    }
    osecir::NpiSimulation sim(m, npis_from_flat(npi), t0, dt);
    sim.ctl.frozen = frozen != 0;
    sim.sim.set_integrator_core(make_core(*integ));
    for (int k = 0; k < num_t; ++k) {
        sim.advance(tmaxs[k]);
    }
    const auto& cpo = sim.sim.model.parameters.ContactPatterns;
    for (int k = 0; k < nq; ++k) {
        for (int i = 0; i < G; ++i) {
            for (int j = 0; j < G; ++j) {
                out_cm[(size_t(k) * G + i) * G + j] = cpo.at(q_t[k], i, j);
            }
        }
    }
    for (size_t l = 0; l < cpo.m.size(); ++l) {
        out_ndamp[l] = int32_t(cpo.m[l].dampings.list.size());
    }
    const int n = int(sim.sim.result.n());
    for (int k = 0; k < std::min(n, max_pts); ++k) {
        out_t[k] = sim.sim.result.times[k];
        std::copy(sim.sim.result.values[k].begin(), sim.sim.result.values[k].end(), out_y + int64_t(k) * G * osecir::NC);
    }
    if (counters) {
        counters[0] = sim.sim.integrator.n_steps;
        counters[1] = sim.sim.integrator.core->n_rejected;
        counters[2] = sim.sim.integrator.core->n_forced;
        counters[3] = sim.ctl.num_implemented;
    }
    return n;
}

// osecirvvs::Simulation (sim_kind = MIO_SIM_MODEL) or mio::Simulation<osecirvvs::Model> (MIO_SIM_BASE) for one node
int orc_simulate_vvs(int sim_kind, int G, const double* pset, const double* local, const double* ct, int max_loc,
                     int max_pts_ct, const double* y0, int vacc_days, const double* vacc_partial,
                     const double* vacc_full, double t0, double tmax, double dt, const mio_integrator_desc* integ,
                     int max_pts, double* out_t, double* out_y, double* out_dt, int64_t* counters)
{
    auto m = osecirvvs_from_flat(G, pset, local, ct, max_loc, max_pts_ct, y0, vacc_days, vacc_partial, vacc_full);
    Simulation<osecirvvs::Model> sim(m, t0, dt);
    sim.model_sim = sim_kind == MIO_SIM_MODEL;
    sim.set_integrator_core(make_core(*integ));
    sim.advance(tmax);
    const int n  = int(sim.result.n());
    const int ne = G * osecirvvs::NC;
    for (int k = 0; k < std::min(n, max_pts); ++k) {
        out_t[k] = sim.result.times[k];
        std::copy(sim.result.values[k].begin(), sim.result.values[k].end(), out_y + int64_t(k) * ne);
    }
    if (out_dt) {
        *out_dt = sim.dt;
    }
    if (counters) {
        counters[0] = sim.integrator.n_steps;
        counters[1] = sim.integrator.core->n_rejected;
        counters[2] = sim.integrator.core->n_forced;
    }
    return n;
}

// make_mobility_sim + GraphSimulation::advance for ONE run described by the flat descriptors of mio_cuda.h.
//   out_state[N][n]          final get_result().get_last_value() of every node
//   out_dt[N]                final integrator dt of every node
//   out_counters[N][4]       accepted, rejected, forced, underflow fix-ups (as destination)
//   out_halfsteps            optional [num_graph_steps][N][n]: node states after every graph step's edge phase
//   max_halfsteps            capacity of out_halfsteps (in graph steps)
// returns number of graph steps taken, or <0 on error.
} // extern "C"

template <class M, class MakeModel>
static int graph_simulate_impl(const mio_graph_desc* desc, const mio_run_data* run, double t_max, double* out_state,
                               double* out_dt, int64_t* out_counters, double* out_halfsteps, int max_halfsteps,
                               int ncomp, MakeModel&& make_model,
                               const std::function<void(GraphSim<M>&)>& on_done = nullptr,
                               const std::vector<std::vector<size_t>>& save_indices = {},
                               const std::function<void(int, SimulationNode<M>&, double, double)>& node_advance = nullptr)
{
    const int G = desc->num_groups, N = desc->num_nodes, E = desc->num_edges;
    const int n = G * ncomp;
    GraphSim<M> g(desc->t0, desc->dt_graph);
    for (int k = 0; k < N; ++k) {
        g.add_node(make_model(k), desc->t0, run->dt0 ? run->dt0[k] : 0.1);
        g.nodes.back().sim.set_integrator_core(make_core(desc->integrator));
        if constexpr (requires(Simulation<M>& sm) { sm.make_flow_simulation(); }) {
            if (desc->flow_simulation) {
                g.nodes.back().sim.make_flow_simulation();
            }
        }
    }
    for (int e = 0; e < E; ++e) {
        Vec c(desc->edge_coeff + int64_t(e) * n, desc->edge_coeff + int64_t(e + 1) * n);
        MobilityEdge me(c, save_indices);
        if (run->mobtables && desc->mob_max_pts > 0) {
            const double* mt = run->mobtables;
            const int K      = int(mt[0]);
            std::vector<double> times(mt + 1, mt + 1 + K);
            std::vector<Mat> mats;
            for (int k = 0; k < K; ++k) {
                Mat d(n, 1);
                std::copy(mt + 1 + desc->mob_max_pts + int64_t(k) * n,
                          mt + 1 + desc->mob_max_pts + int64_t(k + 1) * n, d.a.begin());
                mats.push_back(d);
            }
            me.coefficients.m[0].dampings.set_cache(times, mats);
        }
        g.add_edge(desc->edge_start[e], desc->edge_end[e], me);
    }
    // GraphSimulation::advance, unrolled here only to snapshot the state after each graph step
    int steps  = 0;
    double dts = g.dt;
    while (g.t < t_max) {
        if (g.t + dts > t_max) {
            dts = t_max - g.t;
        }
        for (int k = 0; k < N; ++k) {
            if (node_advance) { // (nodes whose simulation type overrides advance, e.g. dynamic NPIs)
                node_advance(k, g.nodes[size_t(k)], g.t, dts);
            }
            else {
                g.nodes[size_t(k)].advance(g.t, dts);
            }
        }
        g.t += dts;
        for (auto& e : g.edges) {
            e.property.apply_mobility(g.t, dts, g.nodes[e.start], g.nodes[e.end]);
        }
        if (out_halfsteps && steps < max_halfsteps) {
            for (int k = 0; k < N; ++k) {
                std::copy(g.nodes[k].sim.result.last().begin(), g.nodes[k].sim.result.last().end(),
                          out_halfsteps + (int64_t(steps) * N + k) * n);
            }
        }
        ++steps;
    }
    for (int k = 0; k < N; ++k) {
        std::copy(g.nodes[k].sim.result.last().begin(), g.nodes[k].sim.result.last().end(),
                  out_state + int64_t(k) * n);
        if (out_dt) {
            out_dt[k] = g.nodes[k].sim.dt;
        }
        if (out_counters) {
            out_counters[k * 4 + 0] = g.nodes[k].sim.integrator.n_steps;
            out_counters[k * 4 + 1] = g.nodes[k].sim.integrator.core->n_rejected;
            out_counters[k * 4 + 2] = g.nodes[k].sim.integrator.core->n_forced;
            out_counters[k * 4 + 3] = 0;
        }
    }
    if (out_counters) {
        for (auto& e : g.edges) {
            out_counters[e.end * 4 + 3] += e.property.stats.underflows;
        }
    }
    if (on_done) {
        on_done(g);
    }
    return steps;
}

namespace
{

// orc_graph_simulate for either model; on_done(GraphSim<M>&) sees the finished simulation (node time series, edges)
template <class OnDone>
int graph_simulate_dispatch(const mio_graph_desc* desc, const mio_run_data* run, double t_max, double* out_state,
                            double* out_dt, int64_t* out_counters, double* out_halfsteps, int max_halfsteps,
                            OnDone&& on_done, const std::vector<std::vector<size_t>>& save_indices = {})
{
    const int G       = desc->num_groups;
    const int64_t psz = pset_size(desc->model, G);
    const int64_t csz = ctable_size(G, desc->ct_max_loc, desc->ct_max_pts);
    const int n       = G * ncomp_of(desc->model);
    auto ps_of        = [&](int k) {
        return run->psets + (desc->node_pset ? desc->node_pset[k] : 0) * psz;
    };
    auto ct_of = [&](int k) {
        return run->ctables + (desc->node_ctable ? desc->node_ctable[k] : 0) * csz;
    };
    if (desc->model == MIO_MODEL_OSECIR) {
        return graph_simulate_impl<osecir::Model>(
            desc, run, t_max, out_state, out_dt, out_counters, out_halfsteps, max_halfsteps, osecir::NC,
            [&](int k) {
                return osecir_from_flat(G, ps_of(k), run->node_local + int64_t(k) * MIO_NODE_LOCAL, ct_of(k),
                                        desc->ct_max_loc, desc->ct_max_pts, run->y0 + int64_t(k) * n);
            },
            [&](GraphSim<osecir::Model>& g) {
                on_done(g);
            },
            save_indices);
    }
    if (desc->model == MIO_MODEL_OSECIRVVS) {
        return graph_simulate_impl<osecirvvs::Model>(
            desc, run, t_max, out_state, out_dt, out_counters, out_halfsteps, max_halfsteps, osecirvvs::NC,
            [&](int k) {
                const int64_t vo = int64_t(k) * G * desc->vacc_days;
                auto m = osecirvvs_from_flat(G, ps_of(k), run->node_local + int64_t(k) * MIO_NODE_LOCAL, ct_of(k),
                                             desc->ct_max_loc, desc->ct_max_pts, run->y0 + int64_t(k) * n,
                                             desc->vacc_days, desc->vacc_partial ? desc->vacc_partial + vo : nullptr,
                                             desc->vacc_full ? desc->vacc_full + vo : nullptr);
                m.model_sim = desc->sim_kind == MIO_SIM_MODEL;
                return m;
            },
            [&](GraphSim<osecirvvs::Model>& g) {
                on_done(g);
            },
            save_indices);
    }
    return -1;
}

// interpolate_simulation_result(graph) (analyze_result.h:121-132) of a finished graph simulation into
// out_daily[N][num_days][n]; returns false when a node does not yield exactly num_days integer days
template <class GS>
bool daily_from_graph(GS& g, int num_days, int n, double* out_daily)
{
    const int N = int(g.nodes.size());
    for (int k = 0; k < N; ++k) {
        const TimeSeries ts = interpolate_simulation_result(g.nodes[k].sim.result);
        if (int(ts.n()) != num_days) {
            return false;
        }
        for (int d = 0; d < num_days; ++d) {
            std::copy(ts.values[d].begin(), ts.values[d].end(), out_daily + (int64_t(k) * num_days + d) * n);
        }
    }
    return true;
}

} // namespace

extern "C" {

int orc_graph_simulate(const mio_graph_desc* desc, const mio_run_data* run, double t_max, double* out_state,
                       double* out_dt, int64_t* out_counters, double* out_halfsteps, int max_halfsteps)
{
    return graph_simulate_dispatch(desc, run, t_max, out_state, out_dt, out_counters, out_halfsteps, max_halfsteps,
                                   [](auto&) {});
}

// orc_graph_simulate (ODE-SECIR) with DynamicNPIsInfectedSymptoms on the nodes: node_npis[N] (num_thresholds = 0: none).
// The nodes' contact tables must come without change points (the dampings live in the lists the NPIs edit).
// out_cm[N][nq][G*G]: the nodes' contact matrices at the query times afterwards.
int orc_npi_graph_simulate(const mio_graph_desc* desc, const mio_run_data* run, const orc_npi_desc* node_npis,
                           double t_max, double* out_state, double* out_halfsteps, int max_halfsteps, int nq,
                           const double* q_t, double* out_cm, int32_t* out_implemented)
{
    if (desc->model != MIO_MODEL_OSECIR) {
        return -1;
    }
    const int G = desc->num_groups, N = desc->num_nodes;
    const int64_t psz = pset_size(desc->model, G);
    const int64_t csz = ctable_size(G, desc->ct_max_loc, desc->ct_max_pts);
    const int n       = G * osecir::NC;
    std::vector<osecir::NpiController> ctl(size_t(N), osecir::NpiController{});
    for (int k = 0; k < N; ++k) {
        ctl[size_t(k)].npis = npis_from_flat(&node_npis[k]);
    }
    return graph_simulate_impl<osecir::Model>(
        desc, run, t_max, out_state, nullptr, nullptr, out_halfsteps, max_halfsteps, osecir::NC,
        [&](int k) {
            auto m = osecir_from_flat(G, run->psets + (desc->node_pset ? desc->node_pset[k] : 0) * psz,
                                      run->node_local + int64_t(k) * MIO_NODE_LOCAL,
                                      run->ctables + (desc->node_ctable ? desc->node_ctable[k] : 0) * csz,
                                      desc->ct_max_loc, desc->ct_max_pts, run->y0 + int64_t(k) * n);
            for (auto& e : m.parameters.ContactPatterns.m) {
                e.dampings.clear();
            }
            return m;
        },
        [&](GraphSim<osecir::Model>& g) {
            for (int k = 0; k < N; ++k) {
                const auto& cp = g.nodes[size_t(k)].sim.model.parameters.ContactPatterns;
                for (int q = 0; q < nq; ++q) {
                    for (int i = 0; i < G; ++i) {
                        for (int j = 0; j < G; ++j) {
                            out_cm[((size_t(k) * nq + q) * G + i) * G + j] = cp.at(q_t[q], i, j);
                        }
                    }
                }
                if (out_implemented) {
                    out_implemented[k] = ctl[size_t(k)].num_implemented;
                }
            }
        },
        {},
        [&](int k, SimulationNode<osecir::Model>& nd, double t, double dts) {
            // SimulationNode::advance (instant.h:83-90) with osecir::Simulation::advance as the node's advance
            ctl[size_t(k)].advance(nd.sim, t + dts);
            nd.last_state = nd.sim.result.last();
        });
}

// orc_graph_simulate of a graph of FlowSimulation nodes (desc->flow_simulation = 1, osecir): also the cumulative flows
// of every node at the end, out_flows[N][G*15] (FlowSimulation::get_flows().get_last_value())
int orc_graph_simulate_flows(const mio_graph_desc* desc, const mio_run_data* run, double t_max, double* out_state,
                             double* out_flows, int64_t* out_counters)
{
    return graph_simulate_dispatch(desc, run, t_max, out_state, nullptr, out_counters, nullptr, 0, [&](auto& g) {
        for (size_t k = 0; k < g.nodes.size(); ++k) {
            if constexpr (requires { g.nodes[k].sim.flow_result; }) {
                const auto& f = g.nodes[k].sim.flow_result.last();
                std::copy(f.begin(), f.end(), out_flows + k * f.size());
            }
        }
    });
}

// orc_graph_simulate + interpolate_simulation_result(graph): out_daily[N][num_days][n] (the per-node time series of
// every adaptive step interpolated at the integer days, analyze_result.h:121-132,186-265). Returns the number of graph
// steps, -2 when the number of days differs from num_days.
int orc_graph_simulate_daily(const mio_graph_desc* desc, const mio_run_data* run, double t_max, int num_days,
                             double* out_daily, double* out_state, int64_t* out_counters)
{
    const int n = desc->num_groups * ncomp_of(desc->model);
    bool ok     = true;
    std::vector<double> state(out_state ? 0 : size_t(desc->num_nodes) * n);
    const int steps = graph_simulate_dispatch(desc, run, t_max, out_state ? out_state : state.data(), nullptr,
                                              out_counters, nullptr, 0, [&](auto& g) {
                                                  ok = daily_from_graph(g, num_days, n, out_daily);
                                              });
    return ok ? steps : -2;
}

// MobilityEdge::get_mobility_results (instant.h:421-445) of every edge: the edges are built with the same
// saved_compartment_indices (num_groups index groups, group_ptr[num_groups+1] into group_idx). out_times[E][cap],
// out_values[E][cap][num_groups+1], out_npts[E]. Returns the number of graph steps.
int orc_graph_simulate_mobility(const mio_graph_desc* desc, const mio_run_data* run, double t_max, int num_save_groups,
                                const int* group_ptr, const int* group_idx, int cap, int* out_npts, double* out_times,
                                double* out_values, double* out_state)
{
    std::vector<std::vector<size_t>> save(num_save_groups);
    for (int gI = 0; gI < num_save_groups; ++gI) {
        save[gI].assign(group_idx + group_ptr[gI], group_idx + group_ptr[gI + 1]);
    }
    const int n = desc->num_groups * ncomp_of(desc->model);
    std::vector<double> state(out_state ? 0 : size_t(desc->num_nodes) * n);
    return graph_simulate_dispatch(
        desc, run, t_max, out_state ? out_state : state.data(), nullptr, nullptr, nullptr, 0,
        [&](auto& g) {
            const int K = num_save_groups + 1;
            for (size_t e = 0; e < g.edges.size(); ++e) {
                const TimeSeries& mr = g.edges[e].property.mobility_results;
                out_npts[e]          = int(mr.n());
                for (int i = 0; i < int(mr.n()) && i < cap; ++i) {
                    out_times[e * cap + i] = mr.times[i];
                    std::copy(mr.values[i].begin(), mr.values[i].end(), out_values + (e * cap + i) * K);
                }
            }
        },
        save);
}

// analyze_result.h:186-265 on one time series: values[npts][n] at times[npts] -> integer days (abs_tol < 0: default)
// or the given interpolation times. Returns the number of interpolated points (<= cap are stored).
int orc_interpolate(const double* times, const double* values, int npts, int n, const double* tps, int ntps,
                    double abs_tol, int cap, double* out_times, double* out_values)
{
    TimeSeries ts;
    for (int i = 0; i < npts; ++i) {
        ts.add(times[i], Vec(values + int64_t(i) * n, values + int64_t(i + 1) * n));
    }
    TimeSeries r = tps ? interpolate_simulation_result(ts, std::vector<double>(tps, tps + ntps))
                       : (abs_tol < 0 ? interpolate_simulation_result(ts) : interpolate_simulation_result(ts, abs_tol));
    for (int i = 0; i < int(r.n()) && i < cap; ++i) {
        out_times[i] = r.times[i];
        std::copy(r.values[i].begin(), r.values[i].end(), out_values + int64_t(i) * n);
    }
    return int(r.n());
}

// sum_nodes / ensemble_mean / ensemble_percentile (analyze_result.h:267-343) of data[runs][nodes][T][n]:
// out_sum[runs][T][n] (nullable), out_mean[nodes][T][n] (nullable), out_pct[num_p][nodes][T][n]
void orc_ensemble_reduce(const double* data, int runs, int nodes, int T, int n, int num_p, const double* p,
                         double* out_sum, double* out_mean, double* out_pct)
{
    EnsembleResult ens(runs, std::vector<TimeSeries>(nodes));
    for (int r = 0; r < runs; ++r) {
        for (int k = 0; k < nodes; ++k) {
            for (int t = 0; t < T; ++t) {
                const double* v = data + ((int64_t(r) * nodes + k) * T + t) * n;
                ens[r][k].add(double(t), Vec(v, v + n));
            }
        }
    }
    auto dump = [&](const std::vector<TimeSeries>& res, double* out) {
        for (int k = 0; k < int(res.size()); ++k) {
            for (int t = 0; t < T; ++t) {
                std::copy(res[k].values[t].begin(), res[k].values[t].end(), out + (int64_t(k) * T + t) * n);
            }
        }
    };
    if (out_sum) {
        auto s = sum_nodes(ens);
        for (int r = 0; r < runs; ++r) {
            dump(s[r], out_sum + int64_t(r) * T * n);
        }
    }
    if (out_mean) {
        dump(ensemble_mean(ens), out_mean);
    }
    for (int q = 0; q < num_p; ++q) {
        dump(ensemble_percentile(ens, p[q]), out_pct + int64_t(q) * nodes * T * n);
    }
}

// The call sequence of TestMobility.edgeApplyMobility (cpp/tests/test_mobility.cpp:154-215): two identical G=1 nodes,
// one edge with coefficient vector coeff[n]; apply_mobility at t (departure), advance both nodes by dt, apply_mobility
// (return), advance, apply_mobility (departure again). out[3][2][n] = node1/node2 last values after each apply.
void orc_edge_apply_sequence(int G, const double* pset, const double* local, const double* ct, int max_loc,
                             int max_pts, const double* y0, const double* coeff, double t, double dt, double* out)
{
    auto m      = osecir_from_flat(G, pset, local, ct, max_loc, max_pts, y0);
    const int n = G * osecir::NC;
    SimulationNode<osecir::Model> node1(m, t), node2(m, t);
    MobilityEdge edge(Vec(coeff, coeff + n));
    for (int phase = 0; phase < 3; ++phase) {
        if (phase > 0) {
            node1.advance(t, dt);
            node2.advance(t, dt);
            t += dt;
        }
        edge.apply_mobility(t, dt, node1, node2);
        std::copy(node1.sim.result.last().begin(), node1.sim.result.last().end(), out + (phase * 2 + 0) * n);
        std::copy(node2.sim.result.last().begin(), node2.sim.result.last().end(), out + (phase * 2 + 1) * n);
    }
}

// ---- ParameterStudy sampling (oracle side of mio_study_sample_runs) ---------------------------------------------
static orc::sampling::UncertainValue uv_from(const mio_uncertain_value& u)
{
    orc::sampling::UncertainValue v;
    v.value = u.value;
    if (u.kind != MIO_DIST_NONE) {
        orc::sampling::Distribution d;
        d.kind     = u.kind;
        d.a        = u.a;
        d.b        = u.b;
        d.lower    = u.lower;
        d.upper    = u.upper;
        d.quantile = u.quantile;
        v.dist     = d;
    }
    return v;
}

static orc::sampling::DampingSampling ds_from(const mio_damping_sampling& s, int G)
{
    orc::sampling::DampingSampling d;
    d.value = uv_from(s.value);
    d.level = s.level;
    d.type  = s.type;
    d.time  = s.time;
    for (size_t l = 0; l < 64; ++l) {
        if ((s.matrix_mask >> l) & 1u) {
            d.matrices.push_back(l);
        }
    }
    d.groups.assign(s.group_weights, s.group_weights + G);
    return d;
}

// raw RandomNumberGenerator output and key for `count` consecutive counters
void orc_rng_raw(const uint32_t* seeds, int num_seeds, uint64_t counter, int count, uint64_t* out, uint64_t* key_out)
{
    orc::sampling::RandomNumberGenerator rng;
    rng.seed(std::vector<uint32_t>(seeds, seeds + num_seeds));
    rng.counter = counter;
    if (key_out) {
        *key_out = rng.key;
    }
    for (int i = 0; i < count; ++i) {
        out[i] = rng();
    }
}

// threefry2x32-20 on explicit counter/key words (Random123 known-answer tests)
void orc_threefry2x32(const uint32_t* ctr, const uint32_t* key, uint32_t* out)
{
    auto r = orc::sampling::threefry2x32({{ctr[0], ctr[1]}}, {{key[0], key[1]}});
    out[0] = r.v[0];
    out[1] = r.v[1];
}

// one uniform / normal draw sequence (distribution level checks)
void orc_draw_values(const uint32_t* seeds, int num_seeds, uint64_t counter, const mio_uncertain_value* u, int count,
                     double* out)
{
    orc::sampling::RandomNumberGenerator rng;
    rng.seed(std::vector<uint32_t>(seeds, seeds + num_seeds));
    rng.counter = counter;
    for (int i = 0; i < count; ++i) {
        auto v = uv_from(u[i]);
        out[i] = v.draw_sample(rng);
    }
}

// ParameterStudy::run_impl sampling step for runs [run0, run0+count): copy the parameter graph, set the counter,
// osecir::draw_sample(copy); outputs in the flat layouts of mio_run_data. Returns 0 on success.
int orc_study_sample_runs(const mio_study_desc* d, int64_t run0, int count, double* psets, double* ctables,
                          double* node_local, double* y0, double* mobtables, uint64_t* draws)
{
    namespace sp = orc::sampling;
    const bool vvs   = d->model == MIO_MODEL_OSECIRVVS;
    const int Cm     = vvs ? 27 : 10;
    const int AS     = vvs ? MIO_OSECIRVVS_AGE_STRIDE : MIO_OSECIR_AGE_STRIDE;
    const int ntags  = vvs ? int(sp::NumTagsVvs) : int(sp::NumTags);
    const int G = d->num_groups, N = d->num_nodes, L = d->num_locations, n = G * Cm;
    // the parameter graph: every node carries node 0's parameter distributions (only node 0's are ever used)
    sp::GraphU graph;
    graph.n                 = n;
    graph.mobility_location = d->mobility_location;
    sp::SecirParamsU p0;
    p0.G                            = G;
    p0.model                        = vvs ? 1 : 0;
    p0.TestAndTraceCapacityMaxRiskNoSymptoms = uv_from(d->header[MIO_PH_TNT_MAX_RISK_NOSYMP]);
    p0.StartDayNewVariant                    = uv_from(d->header[MIO_PH_START_DAY_NEW_VARIANT]);
    p0.StartDay                     = uv_from(d->header[MIO_PH_START_DAY]);
    p0.Seasonality                  = uv_from(d->header[MIO_PH_SEASONALITY]);
    p0.TestAndTraceCapacityMaxRisk  = uv_from(d->header[MIO_PH_TNT_MAX_RISK]);
    p0.commuter_nondetection        = uv_from(d->header[MIO_PH_COMMUTER_NONDETECTION]);
    p0.start_commuter_detection     = uv_from(d->header[MIO_PH_START_COMMUTER_DETECTION]);
    p0.end_commuter_detection       = uv_from(d->header[MIO_PH_END_COMMUTER_DETECTION]);
    p0.age.assign(ntags, std::vector<sp::UncertainValue>(G));
    for (int i = 0; i < G; ++i) {
        for (int tag = 0; tag < ntags; ++tag) {
            p0.age[tag][i] = uv_from(d->age_params[i * AS + tag]);
        }
    }
    for (int l = 0; l < L; ++l) {
        Mat b(G, G), m(G, G);
        std::copy(d->baseline + int64_t(l) * G * G, d->baseline + int64_t(l + 1) * G * G, b.a.begin());
        if (d->minimum) {
            std::copy(d->minimum + int64_t(l) * G * G, d->minimum + int64_t(l + 1) * G * G, m.a.begin());
        }
        p0.baseline.push_back(b);
        p0.minimum.push_back(m);
    }
    for (int k = 0; k < d->num_dampings; ++k) {
        p0.dampings.push_back(ds_from(d->dampings[k], G));
    }
    if (d->school_holiday_damping) {
        p0.school_holiday_damping = ds_from(*d->school_holiday_damping, G);
    }
    else {
        p0.school_holiday_damping.groups.assign(G, 0.0);
    }
    int num_tables = 1;
    bool any_none  = false;
    if (d->num_holiday_sets > 0) {
        for (int k = 0; k < N; ++k) {
            any_none = any_none || d->node_holiday_set[k] < 0;
        }
        num_tables = d->num_holiday_sets + (any_none ? 1 : 0);
    }
    for (int k = 0; k < N; ++k) {
        sp::NodeModelU node;
        node.C                               = Cm;
        node.parameters                      = p0;
        node.parameters.ICUCapacity          = uv_from(d->node_local[k * 2 + MIO_NL_ICU_CAPACITY]);
        node.parameters.TestAndTraceCapacity = uv_from(d->node_local[k * 2 + MIO_NL_TNT_CAPACITY]);
        if (d->num_holiday_sets > 0 && d->node_holiday_set[k] >= 0) {
            const int hs = d->node_holiday_set[k];
            for (int h = d->holiday_ptr[hs]; h < d->holiday_ptr[hs + 1]; ++h) {
                node.parameters.school_holidays.emplace_back(d->holiday_periods[h * 2], d->holiday_periods[h * 2 + 1]);
            }
        }
        for (int j = 0; j < n; ++j) {
            node.populations.push_back(uv_from(d->populations[int64_t(k) * n + j]));
        }
        graph.nodes.push_back(node);
    }
    sp::RandomNumberGenerator rng;
    rng.seed(std::vector<uint32_t>(d->seeds, d->seeds + d->num_seeds));
    const int64_t psz = pset_size(d->model, G);
    const int64_t csz = ctable_size(G, L, d->ct_max_pts);
    const int64_t msz = d->mob_max_pts > 0 ? 1 + d->mob_max_pts + int64_t(d->mob_max_pts) * n : 0;
    const int64_t lstr = d->ct_max_pts + int64_t(2 + d->ct_max_pts) * G * G;
    for (int r = 0; r < count; ++r) {
        const uint64_t c0 = d->rng_counter + (uint64_t(run0 + r) << 32);
        rng.counter       = c0;
        auto copy         = graph;
        auto [nodes, mob] = sp::draw_sample(copy, rng);
        if (draws) {
            draws[r] = rng.counter - c0;
        }
        // all nodes share the sampled parameter set: flatten node 0's
        double* ps = psets + r * psz;
        std::fill(ps, ps + psz, 0.0);
        const auto& q                     = nodes[0].parameters;
        ps[MIO_PH_START_DAY]              = q.StartDay.value;
        ps[MIO_PH_SEASONALITY]            = q.Seasonality.value;
        ps[MIO_PH_TNT_MAX_RISK]           = q.TestAndTraceCapacityMaxRisk.value;
        ps[MIO_PH_COMMUTER_NONDETECTION]  = q.commuter_nondetection.value;
        ps[MIO_PH_START_COMMUTER_DETECTION] = q.start_commuter_detection.value;
        ps[MIO_PH_END_COMMUTER_DETECTION] = q.end_commuter_detection.value;
        ps[MIO_PH_RESERVED6]              = vvs ? q.TestAndTraceCapacityMaxRiskNoSymptoms.value
                                                : d->header[MIO_PH_RESERVED6].value;
        ps[MIO_PH_RESERVED7]              = vvs ? q.StartDayNewVariant.value : d->header[MIO_PH_RESERVED7].value;
        for (int i = 0; i < G; ++i) {
            for (int tag = 0; tag < ntags; ++tag) {
                ps[MIO_PSET_HEADER + i * AS + tag] = q.age[tag][i].value;
            }
            for (int tag = ntags; tag < AS; ++tag) {
                ps[MIO_PSET_HEADER + i * AS + tag] = d->age_params[i * AS + tag].value;
            }
        }
        for (int k = 0; k < N; ++k) {
            node_local[(int64_t(r) * N + k) * 2 + 0] = nodes[k].parameters.ICUCapacity.value;
            node_local[(int64_t(r) * N + k) * 2 + 1] = nodes[k].parameters.TestAndTraceCapacity.value;
            for (int j = 0; j < n; ++j) {
                y0[(int64_t(r) * N + k) * n + j] = nodes[k].populations[j].value;
            }
        }
        // contact tables: one per holiday set, taken from the first node that uses it
        for (int tbl = 0; tbl < num_tables; ++tbl) {
            int rep = 0;
            if (d->num_holiday_sets > 0) {
                for (int k = 0; k < N; ++k) {
                    const int hs = d->node_holiday_set[k] < 0 ? d->num_holiday_sets : d->node_holiday_set[k];
                    if (hs == tbl) {
                        rep = k;
                        break;
                    }
                }
            }
            double* ct = ctables + (int64_t(r) * num_tables + tbl) * csz;
            std::fill(ct, ct + csz, 0.0);
            ct[0] = L;
            for (int l = 0; l < L; ++l) {
                const auto& dm = nodes[rep].parameters.cont_freq_dampings[l];
                const int K    = int(dm.cache.size()) - 2;
                if (K > d->ct_max_pts) {
                    return 2;
                }
                ct[1 + l] = K;
                double* p = ct + 1 + L + l * lstr;
                for (int k = 0; k < K; ++k) {
                    p[k] = dm.cache[k + 1].second;
                    std::copy(dm.cache[k + 1].first.a.begin(), dm.cache[k + 1].first.a.end(),
                              p + d->ct_max_pts + int64_t(2 + k) * G * G);
                }
                std::copy(q.baseline[l].a.begin(), q.baseline[l].a.end(), p + d->ct_max_pts);
                std::copy(q.minimum[l].a.begin(), q.minimum[l].a.end(), p + d->ct_max_pts + G * G);
            }
        }
        if (msz > 0 && mobtables) {
            double* mt = mobtables + r * msz;
            std::fill(mt, mt + msz, 0.0);
            const int K = int(mob.cache.size()) - 2;
            if (K > d->mob_max_pts) {
                return 2;
            }
            mt[0] = K;
            for (int k = 0; k < K; ++k) {
                mt[1 + k] = mob.cache[k + 1].second;
                std::copy(mob.cache[k + 1].first.a.begin(), mob.cache[k + 1].first.a.end(),
                          mt + 1 + d->mob_max_pts + int64_t(k) * n);
            }
        }
    }
    return 0;
}

// ParameterStudy::run_impl (parameter_studies.h:290-345) for runs [run0, run0+count) of this rank, literally per run:
// copy + draw_sample of the parameter graph (RNG counter base + (run_idx << 32)), make_sampled_graph_simulation,
// advance(t_max), then the process_result of the study example: interpolate_simulation_result(graph)
// (cpp/examples/ode_secir_parameter_study_graph.cpp:292-301). Outputs (all nullable except checksum):
//   out_daily[count][N][num_days][n], out_state[count][N][n], out_counters[count][N][4];
//   checksum = sum of all interpolated values (keeps the interpolation alive when out_daily is null).
// Returns 0, or a negative/positive error code.
int orc_study_run_block(const mio_study_desc* sd, const mio_graph_desc* gd, int64_t run0, int count, double t_max,
                        int num_days, double* out_daily, double* out_state, int64_t* out_counters, double* checksum)
{
    const int G = gd->num_groups, N = gd->num_nodes;
    const int n = G * ncomp_of(gd->model);
    const int64_t psz = pset_size(gd->model, G);
    const int64_t csz = ctable_size(G, gd->ct_max_loc, gd->ct_max_pts);
    const int64_t msz = gd->mob_max_pts > 0 ? 1 + gd->mob_max_pts + int64_t(gd->mob_max_pts) * n : 0;
    std::vector<double> psets(psz * gd->num_psets), ctables(csz * gd->num_ctables), local(size_t(N) * 2),
        y0(size_t(N) * n), mob(msz), daily(size_t(N) * num_days * n), state(size_t(N) * n);
    double sum = 0.0;
    for (int r = 0; r < count; ++r) {
        int rc = orc_study_sample_runs(sd, run0 + r, 1, psets.data(), ctables.data(), local.data(), y0.data(),
                                       msz > 0 ? mob.data() : nullptr, nullptr);
        if (rc != 0) {
            return rc;
        }
        mio_run_data rd{};
        rd.psets      = psets.data();
        rd.ctables    = ctables.data();
        rd.node_local = local.data();
        rd.y0         = y0.data();
        rd.dt0        = nullptr;
        rd.mobtables  = msz > 0 ? mob.data() : nullptr;
        double* dly   = out_daily ? out_daily + int64_t(r) * N * num_days * n : daily.data();
        rc = orc_graph_simulate_daily(gd, &rd, t_max, num_days, dly, out_state ? out_state + int64_t(r) * N * n : state.data(),
                                      out_counters ? out_counters + int64_t(r) * N * 4 : nullptr);
        if (rc < 0) {
            return rc;
        }
        for (size_t i = 0; i < size_t(N) * num_days * n; ++i) {
            sum += dly[i];
        }
    }
    if (checksum) {
        *checksum = sum;
    }
    return 0;
}

// The setup of TestGraphSimulation.consistencyFlowMobility (cpp/tests/test_graph_simulation.cpp:256-339): three
// identical ODE-SEIR nodes, all six directed edges with coefficient 0.001 on every compartment, t0 = 0, tmax = 1,
// dt = 0.001. Returns the number of time points of node 0's result; out_times[cap], out_values[cap][4].
int orc_seir_graph_golden_ex(int flow_sim, int cap, double* out_times, double* out_values);
int orc_seir_graph_golden(int cap, double* out_times, double* out_values)
{
    return orc_seir_graph_golden_ex(0, cap, out_times, out_values);
}
// flow_sim != 0: the nodes are mio::FlowSimulation (sim_flows of the reference test, the one the golden file was
// written from, test_graph_simulation.cpp:307-311,328-338)
int orc_seir_graph_golden_ex(int flow_sim, int cap, double* out_times, double* out_values)
{
    oseir::Model model(1);
    const double total_population = 10000;
    model.populations[oseir::E]   = 100;
    model.populations[oseir::I]   = 100;
    model.populations[oseir::R]   = 100;
    model.populations[oseir::S] =
        total_population - model.populations[oseir::E] - model.populations[oseir::I] - model.populations[oseir::R];
    model.TimeExposed[0]                      = 5.2;
    model.TimeInfected[0]                     = 6;
    model.TransmissionProbabilityOnContact[0] = 0.04;
    Mat b(1, 1);
    b.a[0] = 10;
    model.ContactPatterns.m.emplace_back(b);
    GraphSim<oseir::Model> g(0.0, 0.001);
    for (int k = 0; k < 3; ++k) {
        g.add_node(model, 0.0);
        if (flow_sim) {
            g.nodes.back().sim.make_flow_simulation();
        }
    }
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) {
            if (i != j) {
                g.add_edge(i, j, MobilityEdge(Vec(4, 0.001)));
            }
        }
    }
    g.advance(1.0);
    const TimeSeries& r = g.nodes[0].sim.result;
    for (int i = 0; i < int(r.n()) && i < cap; ++i) {
        out_times[i] = r.times[i];
        std::copy(r.values[i].begin(), r.values[i].end(), out_values + int64_t(i) * 4);
    }
    return int(r.n());
}

} // extern "C"

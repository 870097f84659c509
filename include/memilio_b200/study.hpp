// study.hpp -- mio::ParameterStudy, osecir::draw_sample(Graph&), make_sampled_graph_simulation and the result
// post-processing helpers with the reference's signatures (cpp/memilio/compartments/parameter_studies.h:47-399,
// cpp/models/ode_secir/parameter_space.h:207-251, cpp/memilio/data/analyze_result.h:100-343).
// ParameterStudy::run keeps `run(create_simulation, process_result)`; the simulations a batch of runs creates are
// advanced together in one device handle (one run slot each) instead of one after the other.
#pragma once

#include "graph.hpp"
#include "osecir.hpp"

#include <numeric>

namespace mio
{

namespace osecir
{
namespace detail
{
// what stays a node's own when the shared draw is laid over it: the capacities drawn with the node's demographics and
// its school holidays (parameter_space.h:224-236)
template <class FP>
struct NodeOwnValues {
    UncertainValue<FP> icu, tnt;
    std::vector<std::pair<SimulationTime<FP>, SimulationTime<FP>>> holidays;
    explicit NodeOwnValues(const Parameters<FP>& p)
        : icu(p.template get<ICUCapacity<FP>>())
        , tnt(p.template get<TestAndTraceCapacity<FP>>())
        , holidays(p.template get<ContactPatterns<FP>>().get_school_holidays())
    {
    }
    void put_back(Parameters<FP>& p) const
    {
        p.template get<ICUCapacity<FP>>()                           = icu;
        p.template get<TestAndTraceCapacity<FP>>()                  = tnt;
        p.template get<ContactPatterns<FP>>().get_school_holidays() = holidays;
    }
};

// the sampled contact dampings act on the commuters as well: every damping of the shared contact patterns becomes a
// damping of the edge's coefficient vectors it names, spread over the compartments by the age group weights
// (parameter_space.h:241-248, damping_sampling.h:256-307)
template <class FP, class Contacts>
MobilityParameters<FP> damp_mobility(MobilityParameters<FP> edge, const Contacts& contacts)
{
    auto& members = edge.get_coefficients();
    for (const auto& ds : contacts.get_dampings()) {
        const auto& weights = ds.get_group_weights();
        for (size_t member : ds.get_matrix_indices()) {
            if (member < members.get_num_matrices()) {
                const Eigen::Index len       = members[member].get_baseline().size();
                const Eigen::Index per_group = len / weights.size();
                Eigen::VectorX<FP> factor(len);
                for (Eigen::Index j = 0; j < len; ++j) {
                    factor[j] = ds.get_value().value() * weights[j / per_group];
                }
                members[member].add_damping(factor, ds.get_level(), ds.get_type(), ds.get_time());
            }
        }
    }
    return edge;
}
} // namespace detail

// draw_sample(Graph<Model, MobilityParameters>&), parameter_space.h:207-251. The draw order is the reference's (it fixes
// the random numbers every parameter gets): the infection parameters and the contact dampings once, on the first
// node's model, then per node its demographics; every node ends up with the shared draw plus its own capacities and
// holidays. Dynamic NPIs are not part of this backend.
template <class FP>
Graph<Model<FP>, MobilityParameters<FP>> draw_sample(Graph<Model<FP>, MobilityParameters<FP>>& graph)
{
    auto& nodes = graph.nodes();
    auto& lead  = nodes.front().property;
    draw_sample_infection(lead);
    lead.parameters.template get<ContactPatterns<FP>>().draw_sample_dampings();

    Graph<Model<FP>, MobilityParameters<FP>> sampled;
    for (size_t k = 0; k < nodes.size(); ++k) {
        auto& model = nodes[k].property;
        draw_sample_demographics(model);
        const detail::NodeOwnValues<FP> own(model.parameters);
        if (k != 0) {
            model.parameters = lead.parameters;
        }
        own.put_back(model.parameters);
        model.parameters.template get<ContactPatterns<FP>>().make_matrix();
        model.apply_constraints();
        sampled.add_node(nodes[k].id, model);
    }
    const auto& contacts = lead.parameters.template get<ContactPatterns<FP>>();
    for (auto& edge : graph.edges()) {
        sampled.add_edge(edge.start_node_idx, edge.end_node_idx, detail::damp_mobility<FP>(edge.property, contacts));
    }
    return sampled;
}
} // namespace osecir

// make_sampled_graph_simulation, parameter_studies.h:383-399
template <class FP, class Sim>
auto make_sampled_graph_simulation(const Graph<typename Sim::Model, MobilityParameters<FP>>& sampled_graph, FP t0,
                                   FP dt_node_sim, FP dt_graph_sim)
{
    using SimGraph = Graph<SimulationNode<FP, Sim>, MobilityEdge<FP>>;
    SimGraph sim_graph;
    for (auto&& node : sampled_graph.nodes()) {
        sim_graph.add_node(node.id, node.property, t0, dt_node_sim);
    }
    for (auto&& edge : sampled_graph.edges()) {
        sim_graph.add_edge(edge.start_node_idx, edge.end_node_idx, edge.property);
    }
    return make_mobility_sim<FP, Sim>(t0, dt_graph_sim, std::move(sim_graph));
}

namespace detail
{
// advance a batch of graph simulations (same graph topology) to t_max in ONE device handle; every simulation gets
// the node values of each integer day (and of t_max) appended to its node results
template <class SimT>
void advance_batch(std::vector<SimT>& sims, double t_max)
{
    if (sims.empty()) {
        return;
    }
    using FP = double;
    std::vector<PackedGraph> packed;
    packed.reserve(sims.size());
    int ct_pts = 1;
    for (auto& s : sims) {
        packed.push_back(pack_graph<FP>(s.get_graph()));
        ct_pts = std::max(ct_pts, packed.back().ct_max_pts);
        if (packed.back().dynamic_npis) {
            // the threshold decisions are taken between graph steps by GraphSimulation::advance (one run); a batch
            // is advanced in one call, so the NPIs would be silently ignored
            throw std::runtime_error("ParameterStudy: nodes with dynamic NPIs are not supported in batched runs by "
                                     "this backend (advance the simulations one by one with GraphSimulation::advance)");
        }
        for (auto& e : s.get_graph().edges()) {
            if (!e.property.get_parameters().get_dynamic_npis_infected().get_thresholds().empty()) {
                throw std::runtime_error("ParameterStudy: dynamic NPIs on mobility edges are not supported by this "
                                         "backend");
            }
        }
    }
    int num_psets = 1, num_ctables = 1, mob_pts = 0;
    for (size_t r = 0; r < sims.size(); ++r) {
        if (packed[r].ct_max_pts != ct_pts) {
            packed[r] = pack_graph<FP>(sims[r].get_graph(), ct_pts);
        }
        if (!packed[r].same_shape(packed[0])) {
            throw std::runtime_error("ParameterStudy: the simulations of one study must share the graph topology, "
                                     "mobility coefficients and integrator");
        }
        num_psets   = std::max(num_psets, int(packed[r].psets.size()));
        num_ctables = std::max(num_ctables, int(packed[r].ctables.size()));
        mob_pts     = std::max(mob_pts, packed[r].mob_max_pts);
    }
    for (auto& p : packed) { // same mobility-table capacity for every run
        if (p.mob_max_pts != mob_pts) {
            const size_t tsz = size_t(1 + mob_pts + mob_pts * p.n);
            std::vector<double> tbl(tsz, 0.0);
            if (p.mob_max_pts > 0) {
                const int K = p.mob_max_pts;
                tbl[0]      = p.mobtable[0];
                std::copy(p.mobtable.begin() + 1, p.mobtable.begin() + 1 + K, tbl.begin() + 1);
                std::copy(p.mobtable.begin() + 1 + K, p.mobtable.end(), tbl.begin() + 1 + mob_pts);
            }
            p.mobtable    = std::move(tbl);
            p.mob_max_pts = mob_pts;
        }
        // node -> set/table indices must agree with run 0 for a shared handle; pad by repeating
        if (p.node_pset != packed[0].node_pset || p.node_ctable != packed[0].node_ctable) {
            throw std::runtime_error("ParameterStudy: runs assign parameter sets / contact tables to nodes "
                                     "differently");
        }
    }
    const double t0  = sims[0].get_t();
    const int day0   = int(std::ceil(t0 - 1e-14));
    const int days   = std::max(1, int(std::floor(t_max + 1e-14)) - day0 + 1);
    const double dtg = sims[0].get_dt();
    DeviceGraph dev(packed[0], dtg, int(sims.size()), num_psets, num_ctables, days);
    for (size_t r = 0; r < sims.size(); ++r) {
        dev.set_run(int(r), packed[r], num_psets, num_ctables);
    }
    // MobilityEdge(coeffs, save_indices): get_mobility_results of every run's edges
    std::vector<std::vector<size_t>> save;
    for (auto& e : sims[0].get_graph().edges()) {
        const auto& si = e.property.get_parameters().get_save_indices();
        if (!si.empty() && save.empty()) {
            save = si;
        }
        if (!save.empty() && si != save) {
            throw std::runtime_error("ParameterStudy: all edges must share the saved_compartment_indices");
        }
    }
    int mob_cap = 0;
    if (!save.empty()) {
        std::vector<int32_t> ptr(1, 0), idx;
        for (auto& grp : save) {
            for (auto i : grp) {
                idx.push_back(int32_t(i));
            }
            ptr.push_back(int32_t(idx.size()));
        }
        mob_cap = int(std::ceil((t_max - t0) / dtg)) + 2;
        cuda_check(mio_cuda_enable_mobility_results(dev.get(), int32_t(save.size()), ptr.data(), idx.data(), mob_cap));
    }
    cuda_check(mio_cuda_advance(dev.get(), t_max));
    if (!save.empty()) {
        const size_t E = sims[0].get_graph().edges().size(), K1 = save.size() + 1, cap = size_t(mob_cap);
        std::vector<int32_t> npts(sims.size() * E);
        std::vector<double> times(cap), values(sims.size() * E * cap * K1);
        cuda_check(mio_cuda_get_mobility_results(dev.get(), 0, int(sims.size()), npts.data(), times.data(),
                                                 values.data()));
        for (size_t r = 0; r < sims.size(); ++r) {
            for (size_t e = 0; e < E; ++e) {
                auto& ts = sims[r].get_graph().edges()[e].property.get_mobility_results();
                for (int i = 0; i < npts[r * E + e]; ++i) {
                    Eigen::VectorX<FP> v{Eigen::Index(K1)};
                    const size_t off = ((r * E + e) * cap + size_t(i)) * K1;
                    std::copy(values.begin() + off, values.begin() + off + K1, v.data());
                    ts.add_time_point(times[size_t(i)], v);
                }
            }
        }
    }
    const size_t N = size_t(dev.N()), n = size_t(dev.n());
    const size_t nsim = sims.size(), ndays = size_t(days);
    std::vector<double> daily(nsim * N * ndays * n), state(nsim * N * n), day_state(N * n);
    int32_t nd = 0;
    cuda_check(mio_cuda_get_daily(dev.get(), 0, int(sims.size()), &nd, daily.data(), int64_t(daily.size())));
    cuda_check(mio_cuda_get_state(dev.get(), 0, int(sims.size()), state.data()));
    for (size_t r = 0; r < sims.size(); ++r) {
        for (int d = 1; d < nd; ++d) { // day 0 is the start value the node results already hold
            for (size_t k = 0; k < N; ++k) {
                std::copy(daily.begin() + ((r * N + k) * size_t(nd) + size_t(d)) * n,
                          daily.begin() + ((r * N + k) * size_t(nd) + size_t(d) + 1) * n, day_state.begin() + k * n);
            }
            sims[r].append_time_point(double(day0 + d), day_state.data());
        }
        if (std::floor(t_max) != t_max) {
            sims[r].append_time_point(t_max, state.data() + r * N * n);
        }
        sims[r].set_t(t_max);
    }
}
} // namespace detail

// ParameterStudy, parameter_studies.h:47-374
template <class ParameterType, class Time = ScalarType, class Step = ScalarType>
class ParameterStudy
{
public:
    ParameterStudy(const ParameterType& parameters, Time t0, Time tmax, Step dt, size_t num_runs)
        : m_parameters(parameters)
        , m_num_runs(num_runs)
        , m_t0{t0}
        , m_tmax{tmax}
        , m_dt(dt)
    {
    }
    template <class CreateSimulationFunction, class ProcessSimulationResultFunction>
    auto run(CreateSimulationFunction&& create_simulation, ProcessSimulationResultFunction&& process_simulation_result)
    {
        return run_impl(0, m_num_runs, create_simulation, process_simulation_result);
    }
    template <class CreateSimulationFunction, class ProcessSimulationResultFunction>
    auto run_serial(CreateSimulationFunction&& create_simulation,
                    ProcessSimulationResultFunction&& process_simulation_result)
    {
        return run_impl(0, m_num_runs, create_simulation, process_simulation_result);
    }
    size_t get_num_runs() const
    {
        return m_num_runs;
    }
    Time get_t0() const
    {
        return m_t0;
    }
    Time get_tmax() const
    {
        return m_tmax;
    }
    Step get_dt() const
    {
        return m_dt;
    }
    const ParameterType& get_parameters() const
    {
        return m_parameters;
    }
    ParameterType& get_parameters()
    {
        return m_parameters;
    }
    RandomNumberGenerator& get_rng()
    {
        return m_rng;
    }
    // runs advanced together on the device (memory: ~ N*(4 + days)*n*8 bytes per run)
    void set_batch_size(size_t b)
    {
        m_batch = std::max<size_t>(1, b);
    }
    // ParameterStudy::distribute_runs, parameter_studies.h:354-367
    static std::vector<size_t> distribute_runs(size_t num_runs, int num_procs)
    {
        size_t num_runs_local = num_runs / size_t(num_procs);
        size_t remainder      = num_runs % size_t(num_procs);
        std::vector<size_t> run_distribution(size_t(num_procs), num_runs_local);
        std::fill(run_distribution.begin(), run_distribution.begin() + remainder, num_runs_local + 1);
        return run_distribution;
    }

private:
    template <class CreateF, class ProcessF>
    auto run_impl(size_t start_run_idx, size_t end_run_idx, CreateF& create_simulation, ProcessF& process)
    {
        using SimT    = std::decay_t<decltype(create_simulation(std::as_const(m_parameters), std::as_const(m_t0),
                                                             std::as_const(m_dt), size_t(0)))>;
        using ResultT = decltype(process(std::declval<SimT&&>(), size_t(0)));
        constexpr bool gather = !std::is_void_v<ResultT>;
        std::vector<std::conditional_t<gather, ResultT, int>> ensemble_result;
        thread_local_rng() = m_rng; // parameter_studies.h:299
        for (size_t b0 = start_run_idx; b0 < end_run_idx; b0 += m_batch) {
            const size_t b1 = std::min(end_run_idx, b0 + m_batch);
            std::vector<SimT> sims;
            sims.reserve(b1 - b0);
            for (size_t run_idx = b0; run_idx < b1; ++run_idx) {
                // parameter_studies.h:310-313: counter = m_rng counter + (run_idx << 32)
                thread_local_rng().set_counter(m_rng.get_counter() + (uint64_t(uint32_t(run_idx)) << 32));
                sims.push_back(create_simulation(std::as_const(m_parameters), std::as_const(m_t0),
                                                 std::as_const(m_dt), run_idx));
            }
            detail::advance_batch(sims, m_tmax);
            for (size_t run_idx = b0; run_idx < b1; ++run_idx) {
                if constexpr (gather) {
                    ensemble_result.emplace_back(process(std::move(sims[run_idx - b0]), run_idx));
                }
                else {
                    process(std::move(sims[run_idx - b0]), run_idx);
                }
            }
        }
        m_rng.set_counter(m_rng.get_counter() + (uint64_t(m_num_runs) << 32)); // parameter_studies.h:339
        if constexpr (gather) {
            return ensemble_result;
        }
    }

    ParameterType m_parameters;
    size_t m_num_runs;
    Time m_t0, m_tmax;
    Step m_dt;
    RandomNumberGenerator m_rng;
    size_t m_batch = 256;
};

// ---- analyze_result.h ---------------------------------------------------------------------------------------------
// interpolate_simulation_result(TimeSeries, interpolation_times), analyze_result.h:213-265 (math/interpolation.h:44)
template <class FP>
TimeSeries<FP> interpolate_simulation_result(const TimeSeries<FP>& simulation_result,
                                             const std::vector<FP>& interpolation_times)
{
    TimeSeries<FP> interpolated(simulation_result.get_num_elements());
    if (interpolation_times.empty()) {
        return interpolated;
    }
    size_t interp_idx = 0;
    if (simulation_result.get_time(0) >= interpolation_times[0]) {
        interpolated.add_time_point(interpolation_times[0], simulation_result[0]);
        ++interp_idx;
    }
    for (Eigen::Index sim_idx = 0;
         sim_idx < simulation_result.get_num_time_points() - 1 && interp_idx < interpolation_times.size();) {
        const FP t1 = simulation_result.get_time(sim_idx), t2 = simulation_result.get_time(sim_idx + 1);
        const FP t = interpolation_times[interp_idx];
        if (t1 < t && t2 >= t) {
            const FP weight = (t - t1) / (t2 - t1);
            const auto &y1 = simulation_result[sim_idx], &y2 = simulation_result[sim_idx + 1];
            Eigen::VectorX<FP> v(y1.size());
            for (Eigen::Index j = 0; j < y1.size(); ++j) {
                v[j] = y1[j] + (y2[j] - y1[j]) * weight;
            }
            interpolated.add_time_point(t, v);
            ++interp_idx;
        }
        else {
            ++sim_idx;
        }
    }
    if (interp_idx < interpolation_times.size() && simulation_result.get_last_time() < interpolation_times[interp_idx]) {
        interpolated.add_time_point(interpolation_times[interp_idx], simulation_result.get_last_value());
    }
    return interpolated;
}

template <class FP>
TimeSeries<FP> interpolate_simulation_result(const TimeSeries<FP>& simulation_result, const FP abs_tol = FP{100.} *
                                                                                      std::numeric_limits<FP>::epsilon())
{
    const auto t0      = simulation_result.get_time(0);
    const auto t_max   = simulation_result.get_last_time();
    const auto day0    = static_cast<int>(std::ceil(t0 - abs_tol));
    const auto day_max = static_cast<int>(std::floor(t_max + abs_tol));
    const size_t ntp = size_t(day_max - day0 + 1);
    std::vector<FP> tps(ntp);
    std::iota(tps.begin(), tps.end(), FP(day0));
    return interpolate_simulation_result<FP>(simulation_result, tps);
}

template <class FP, class Simulation>
std::vector<TimeSeries<FP>>
interpolate_simulation_result(const Graph<SimulationNode<FP, Simulation>, MobilityEdge<FP>>& graph_result)
{
    std::vector<TimeSeries<FP>> interpolated;
    interpolated.reserve(graph_result.nodes().size());
    for (auto& n : graph_result.nodes()) {
        interpolated.push_back(interpolate_simulation_result(n.property.get_result()));
    }
    return interpolated;
}

namespace detail
{
// ensemble [run][node] TimeSeries -> column-major device matrix [column][run] (columns = node, time, element)
template <class FP>
std::vector<TimeSeries<FP>> reduce_ensemble(const std::vector<std::vector<TimeSeries<FP>>>& ens, const FP* p)
{
    const size_t R = ens.size(), N = ens[0].size();
    const size_t T = size_t(ens[0][0].get_num_time_points()), El = size_t(ens[0][0].get_num_elements());
    const size_t ncol = N * T * El;
    std::vector<double> host(ncol * R);
    for (size_t r = 0; r < R; ++r) {
        for (size_t k = 0; k < N; ++k) {
            for (size_t t = 0; t < T; ++t) {
                for (size_t e = 0; e < El; ++e) {
                    host[((k * T + t) * El + e) * R + r] = ens[r][k][Eigen::Index(t)][Eigen::Index(e)];
                }
            }
        }
    }
    std::vector<double> out(ncol);
    cuda_check(mio_cuda_reduce_host_columns(host.data(), int64_t(ncol), int64_t(R), p ? 1 : 0, p, p ? out.data() : nullptr,
                                            p ? nullptr : out.data()));
    std::vector<TimeSeries<FP>> res(N, TimeSeries<FP>(Eigen::Index(El)));
    for (size_t k = 0; k < N; ++k) {
        for (size_t t = 0; t < T; ++t) {
            Eigen::VectorX<FP> v{Eigen::Index(El)};
            for (size_t e = 0; e < El; ++e) {
                v[Eigen::Index(e)] = out[(k * T + t) * El + e];
            }
            res[k].add_time_point(ens[0][k].get_time(Eigen::Index(t)), v);
        }
    }
    return res;
}
} // namespace detail

// ensemble_percentile / ensemble_mean, analyze_result.h:289-343: reduced on the device
template <class FP>
std::vector<TimeSeries<FP>> ensemble_percentile(const std::vector<std::vector<TimeSeries<FP>>>& ensemble_result, FP p)
{
    return detail::reduce_ensemble(ensemble_result, &p);
}
template <class FP>
std::vector<TimeSeries<FP>> ensemble_mean(const std::vector<std::vector<TimeSeries<FP>>>& ensemble_result)
{
    return detail::reduce_ensemble<FP>(ensemble_result, nullptr);
}
// sum_nodes, analyze_result.h:267-287 (a host-side sum; no device work worth a launch)
template <class FP>
std::vector<std::vector<TimeSeries<FP>>> sum_nodes(const std::vector<std::vector<TimeSeries<FP>>>& ensemble_result)
{
    std::vector<std::vector<TimeSeries<FP>>> sum_result;
    for (auto& run : ensemble_result) {
        const auto T = run[0].get_num_time_points(), El = run[0].get_num_elements();
        TimeSeries<FP> s(El);
        for (Eigen::Index t = 0; t < T; ++t) {
            auto& v = s.add_time_point(run[0].get_time(t));
            for (auto& node : run) {
                v += node[t];
            }
        }
        sum_result.push_back({s});
    }
    return sum_result;
}

} // namespace mio

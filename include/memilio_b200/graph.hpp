// graph.hpp -- mio::Graph, SimulationNode, MobilityParameters, MobilityEdge, GraphSimulation, make_mobility_sim with
// the reference's signatures (cpp/memilio/mobility/graph.h:145-300, metapopulation_mobility_instant.h:40-101,115-260,
// 270-420,656-678, graph_simulation.h:41-121) on top of the C-ABI. A GraphSimulation owns one run slot of a device
// handle; ParameterStudy batches many of them into one handle (study.hpp).
#pragma once

#include "mio.hpp"

#include <cstring>

namespace mio
{

template <class NodePropertyT>
struct Node {
    template <class... Args>
    Node(int node_id, Args&&... args)
        : id{node_id}
        , property(std::forward<Args>(args)...)
    {
    }
    int id;
    NodePropertyT property;
};

template <class EdgePropertyT>
struct Edge {
    template <class... Args>
    Edge(size_t start, size_t end, Args&&... args)
        : start_node_idx(start)
        , end_node_idx(end)
        , property(std::forward<Args>(args)...)
    {
    }
    size_t start_node_idx, end_node_idx;
    EdgePropertyT property;
};

template <class NodePropertyT, class EdgePropertyT>
class Graph
{
public:
    using NodeProperty = NodePropertyT;
    using EdgeProperty = EdgePropertyT;
    template <class... Args>
    void add_node(int id, Args&&... args)
    {
        m_nodes.emplace_back(id, std::forward<Args>(args)...);
    }
    // edges stay sorted by (start, end); an existing edge is replaced (graph.h:233-241, stl_util.h:82-96)
    template <class... Args>
    void add_edge(size_t start_node_idx, size_t end_node_idx, Args&&... args)
    {
        assert(m_nodes.size() > start_node_idx && m_nodes.size() > end_node_idx);
        Edge<EdgePropertyT> e(start_node_idx, end_node_idx, std::forward<Args>(args)...);
        auto less = [](const Edge<EdgePropertyT>& a, const Edge<EdgePropertyT>& b) {
            return a.start_node_idx == b.start_node_idx ? a.end_node_idx < b.end_node_idx
                                                        : a.start_node_idx < b.start_node_idx;
        };
        auto range = std::equal_range(m_edges.begin(), m_edges.end(), e, less);
        if (range.second - range.first == 1) {
            *range.first = std::move(e);
        }
        else {
            m_edges.insert(range.first, std::move(e));
        }
    }
    std::vector<Node<NodePropertyT>>& nodes()
    {
        return m_nodes;
    }
    const std::vector<Node<NodePropertyT>>& nodes() const
    {
        return m_nodes;
    }
    std::vector<Edge<EdgePropertyT>>& edges()
    {
        return m_edges;
    }
    const std::vector<Edge<EdgePropertyT>>& edges() const
    {
        return m_edges;
    }

private:
    std::vector<Node<NodePropertyT>> m_nodes;
    std::vector<Edge<EdgePropertyT>> m_edges;
};

// MobilityCoefficients = DampingMatrixExpression<VectorDampings> (metapopulation_mobility_instant.h:103-113)
template <class FP = ScalarType>
class MobilityCoefficients
{
public:
    struct RawDamping {
        Eigen::VectorX<FP> v;
        int level, type;
        FP time;
    };
    explicit MobilityCoefficients(const Eigen::VectorX<FP>& baseline)
        : m_baseline(baseline)
    {
    }
    explicit MobilityCoefficients(size_t n)
        : m_baseline(Eigen::VectorX<FP>::Zero(Eigen::Index(n)))
    {
    }
    Eigen::VectorX<FP>& get_baseline()
    {
        return m_baseline;
    }
    const Eigen::VectorX<FP>& get_baseline() const
    {
        return m_baseline;
    }
    void add_damping(const Eigen::VectorX<FP>& v, DampingLevel l, DampingType ty, SimulationTime<FP> t)
    {
        m_dampings.push_back(RawDamping{v, l.value, ty.value, t.get()});
    }
    void add_damping(FP d, SimulationTime<FP> t)
    {
        add_damping(Eigen::VectorX<FP>::Constant(m_baseline.size(), d), DampingLevel(0), DampingType(0), t);
    }
    const std::vector<RawDamping>& get_dampings() const
    {
        return m_dampings;
    }

private:
    Eigen::VectorX<FP> m_baseline;
    std::vector<RawDamping> m_dampings;
};

template <class FP = ScalarType>
class MobilityCoefficientGroup
{
public:
    MobilityCoefficientGroup(size_t num_members, size_t n)
        : m_members(num_members, MobilityCoefficients<FP>(n))
    {
    }
    MobilityCoefficientGroup(std::initializer_list<MobilityCoefficients<FP>> l)
        : m_members(l)
    {
    }
    MobilityCoefficients<FP>& operator[](size_t i)
    {
        return m_members[i];
    }
    const MobilityCoefficients<FP>& operator[](size_t i) const
    {
        return m_members[i];
    }
    size_t get_num_matrices() const
    {
        return m_members.size();
    }

private:
    std::vector<MobilityCoefficients<FP>> m_members;
};

// MobilityParameters, metapopulation_mobility_instant.h:115-260
template <class FP = ScalarType>
class MobilityParameters
{
public:
    MobilityParameters(const MobilityCoefficientGroup<FP>& coeffs)
        : m_coefficients(coeffs)
    {
    }
    MobilityParameters(const Eigen::VectorX<FP>& coeffs)
        : m_coefficients({MobilityCoefficients<FP>(coeffs)})
    {
    }
    MobilityParameters(const MobilityCoefficientGroup<FP>& coeffs,
                       const std::vector<std::vector<size_t>>& save_indices)
        : m_coefficients(coeffs)
        , m_saved_compartment_indices(save_indices)
    {
    }
    MobilityParameters(const Eigen::VectorX<FP>& coeffs, const std::vector<std::vector<size_t>>& save_indices)
        : m_coefficients({MobilityCoefficients<FP>(coeffs)})
        , m_saved_compartment_indices(save_indices)
    {
    }
    const MobilityCoefficientGroup<FP>& get_coefficients() const
    {
        return m_coefficients;
    }
    MobilityCoefficientGroup<FP>& get_coefficients()
    {
        return m_coefficients;
    }
    const std::vector<std::vector<size_t>>& get_save_indices() const
    {
        return m_saved_compartment_indices;
    }
    // dynamic NPIs on the mobility coefficients (metapopulation_mobility_instant.h:225-245, :549-566). The device keeps
    // one mobility damping table per run, so edges that adapt their coefficients individually are refused by
    // GraphSimulation::advance (loudly) rather than ignored.
    const DynamicNPIs<FP>& get_dynamic_npis_infected() const
    {
        return m_dynamic_npis;
    }
    void set_dynamic_npis_infected(const DynamicNPIs<FP>& v)
    {
        m_dynamic_npis = v;
    }

private:
    MobilityCoefficientGroup<FP> m_coefficients;
    std::vector<std::vector<size_t>> m_saved_compartment_indices;
    DynamicNPIs<FP> m_dynamic_npis;
};

// MobilityEdge, metapopulation_mobility_instant.h:270-420: parameters only; apply_mobility runs in csrc/edges.cu
template <class FP = ScalarType>
class MobilityEdge
{
public:
    MobilityEdge(const MobilityParameters<FP>& params)
        : m_parameters(params)
        , m_mobility_results(Eigen::Index(params.get_save_indices().size() + 1))
    {
    }
    MobilityEdge(const Eigen::VectorX<FP>& coeffs)
        : MobilityEdge(MobilityParameters<FP>(coeffs))
    {
    }
    MobilityEdge(const Eigen::VectorX<FP>& coeffs, const std::vector<std::vector<size_t>>& save_indices)
        : MobilityEdge(MobilityParameters<FP>(coeffs, save_indices))
    {
    }
    const MobilityParameters<FP>& get_parameters() const
    {
        return m_parameters;
    }
    // add_mobility_result_time_point (instant.h:421-445): GraphSimulation::advance appends one time point per graph
    // step when the edge was built with saved_compartment_indices (mio_cuda_enable_mobility_results)
    TimeSeries<FP>& get_mobility_results()
    {
        return m_mobility_results;
    }
    const TimeSeries<FP>& get_mobility_results() const
    {
        return m_mobility_results;
    }

private:
    MobilityParameters<FP> m_parameters;
    TimeSeries<FP> m_mobility_results;
};

// SimulationNode, metapopulation_mobility_instant.h:40-101
template <class FP, class Sim>
class SimulationNode
{
public:
    using Simulation = Sim;
    template <class... Args>
    SimulationNode(Args&&... args)
        : m_simulation(std::forward<Args>(args)...)
        , m_last_state(m_simulation.get_result().get_last_value())
        , m_t0(m_simulation.get_result().get_last_time())
    {
    }
    decltype(auto) get_result() const
    {
        return m_simulation.get_result();
    }
    decltype(auto) get_result()
    {
        return m_simulation.get_result();
    }
    Sim& get_simulation()
    {
        return m_simulation;
    }
    const Sim& get_simulation() const
    {
        return m_simulation;
    }
    const Eigen::VectorX<FP>& get_last_state() const
    {
        return m_last_state;
    }
    FP get_t0() const
    {
        return m_t0;
    }

private:
    Sim m_simulation;
    Eigen::VectorX<FP> m_last_state;
    FP m_t0;
};

namespace detail
{

// Flat image of one graph: descriptors of mio_cuda.h. Parameter sets and contact tables are deduplicated by value
// (after draw_sample all nodes share one parameter set; contact tables differ only by school holidays).
// what the packing needs to know about a node's simulation type: model kind, state size per age group, contact
// patterns, and (osecirvvs) the node type and the cumulative vaccination arrays. Specialised next to the models.
template <class Sim>
struct SimTraits;
template <class FP, class BaseT>
struct SimTraits<osecir::Simulation<FP, BaseT>> {
    static constexpr int model_kind = MIO_MODEL_OSECIR;
    static constexpr int sim_kind   = MIO_SIM_MODEL;
    static constexpr int NC         = osecir::detail::NC;
    static constexpr bool flow_simulation = osecir::Simulation<FP, BaseT>::is_flow_simulation;
    static bool has_dynamic_npis(const osecir::Model<FP>& m)
    {
        return !m.parameters.template get<osecir::DynamicNPIsInfectedSymptoms<FP>>().get_thresholds().empty();
    }
    // the dynamic NPI decision of a node before a graph step that starts at t (see osecir::Simulation)
    static bool check_dynamic_npis(osecir::Simulation<FP, BaseT>& sim, FP t, FP t_step_end)
    {
        const bool changed = sim.check_dynamic_npis(t);
        const FP npi_end   = sim.get_dynamic_npi_end();
        if (t < npi_end && npi_end < t_step_end - 1e-10) {
            throw std::runtime_error("GraphSimulation: a dynamic NPI ends inside a graph step (the node would be "
                                     "integrated in two chunks); align duration and delay with the graph's dt");
        }
        return changed;
    }
    static const ContactMatrixGroup<FP>& contacts(const osecir::Model<FP>& m)
    {
        return static_cast<const ContactMatrixGroup<FP>&>(m.parameters.template get<osecir::ContactPatterns<FP>>());
    }
    static void vaccinations(const osecir::Model<FP>&, std::vector<double>&, std::vector<double>&, int& days)
    {
        days = 0;
    }
};

template <class FP>
struct SimTraits<Simulation<FP, osecir::Model<FP>>> : SimTraits<osecir::Simulation<FP>> {
};
template <class FP>
struct SimTraits<FlowSimulation<FP, osecir::Model<FP>>>
    : SimTraits<osecir::Simulation<FP, FlowSimulation<FP, osecir::Model<FP>>>> {
};

struct PackedGraph {
    bool dynamic_npis = false;  // nodes carry dynamic NPIs: one contact table per node, room for added change points
    int flow_simulation = 0;    // the nodes are FlowSimulations
    std::vector<double> flows0; // [N][G*15] the flows they continue from
    int model_kind = MIO_MODEL_OSECIR, sim_kind = MIO_SIM_MODEL, vacc_days = 0;
    std::vector<double> vacc_partial, vacc_full; // [N][G][vacc_days]
    int G = 0, N = 0, E = 0, n = 0, ct_max_loc = 1, ct_max_pts = 1, mob_max_pts = 0;
    std::vector<int32_t> edge_start, edge_end, node_pset, node_ctable;
    std::vector<double> edge_coeff;
    std::vector<std::vector<double>> psets, ctables;
    std::vector<double> node_local, y0, dt0, mobtable;
    mio_integrator_desc integrator{};
    double t0 = 0.0;

    bool same_shape(const PackedGraph& o) const
    {
        return G == o.G && N == o.N && E == o.E && ct_max_loc == o.ct_max_loc && edge_start == o.edge_start &&
               edge_end == o.edge_end && edge_coeff == o.edge_coeff && t0 == o.t0 &&
               std::memcmp(&integrator, &o.integrator, sizeof integrator) == 0;
    }
};

template <class FP, class GraphT>
PackedGraph pack_graph(const GraphT& g, int ct_max_pts_min = 1)
{
    PackedGraph p;
    if (g.nodes().empty()) {
        throw std::runtime_error("pack_graph: graph without nodes");
    }
    using Sim    = std::decay_t<decltype(g.nodes()[0].property.get_simulation())>;
    using Traits = SimTraits<Sim>;
    const auto& m0 = g.nodes()[0].property.get_simulation().get_model();
    p.model_kind   = Traits::model_kind;
    p.sim_kind     = Traits::sim_kind;
    p.flow_simulation = Traits::flow_simulation ? 1 : 0;
    p.G            = int(m0.parameters.get_num_groups().value);
    p.N            = int(g.nodes().size());
    p.E            = int(g.edges().size());
    p.n            = p.G * Traits::NC;
    p.t0           = g.nodes()[0].property.get_result().get_last_time();
    p.integrator   = g.nodes()[0].property.get_simulation().get_integrator_core().desc();
    int L = 1, K = ct_max_pts_min;
    for (auto& nd : g.nodes()) {
        p.dynamic_npis = p.dynamic_npis || Traits::has_dynamic_npis(nd.property.get_simulation().get_model());
    }
    if (p.dynamic_npis) {
        K = std::max(K, 24);
    }
    for (auto& nd : g.nodes()) {
        const auto& cm = Traits::contacts(nd.property.get_simulation().get_model());
        L = std::max(L, int(cm.get_num_matrices()));
        K = std::max(K, osecir::detail::num_change_points(cm));
    }
    p.ct_max_loc = L;
    p.ct_max_pts = K;
    const size_t psz = size_t(mio_pset_size(p.model_kind, p.G));
    for (auto& nd : g.nodes()) {
        const auto& sim   = nd.property.get_simulation();
        const auto& model = sim.get_model();
        std::vector<double> ps(psz, 0.0);
        double local[MIO_NODE_LOCAL];
        model.parameters.pack_values(ps.data(), local);
        auto ip = std::find(p.psets.begin(), p.psets.end(), ps);
        p.node_pset.push_back(int32_t(ip - p.psets.begin()));
        if (ip == p.psets.end()) {
            p.psets.push_back(std::move(ps));
        }
        auto ct = osecir::detail::contact_table(Traits::contacts(model), p.G, L, K);
        auto ic = p.dynamic_npis ? p.ctables.end() : std::find(p.ctables.begin(), p.ctables.end(), ct);
        p.node_ctable.push_back(int32_t(ic - p.ctables.begin()));
        if (ic == p.ctables.end()) {
            p.ctables.push_back(std::move(ct));
        }
        p.node_local.push_back(local[0]);
        p.node_local.push_back(local[1]);
        const auto& y = nd.property.get_result().get_last_value();
        p.y0.insert(p.y0.end(), y.data(), y.data() + p.n);
        p.dt0.push_back(sim.get_dt());
        if constexpr (Traits::flow_simulation) {
            const auto& f = sim.get_flows().get_last_value();
            p.flows0.insert(p.flows0.end(), f.data(), f.data() + f.size());
        }
        int days = 0;
        Traits::vaccinations(model, p.vacc_partial, p.vacc_full, days);
        if (&nd != &g.nodes()[0] && days != p.vacc_days) {
            throw std::runtime_error("pack_graph: the nodes' vaccination arrays must cover the same days");
        }
        p.vacc_days = days;
    }
    // edges: baseline = sum of the group's members; dampings must be common to all edges (one table per run)
    bool have_mob = false;
    for (auto& e : g.edges()) {
        p.edge_start.push_back(int32_t(e.start_node_idx));
        p.edge_end.push_back(int32_t(e.end_node_idx));
        const auto& grp = e.property.get_parameters().get_coefficients();
        const size_t nn = size_t(p.n);
        std::vector<double> c(nn, 0.0);
        for (size_t mbr = 0; mbr < grp.get_num_matrices(); ++mbr) {
            if (grp[mbr].get_baseline().size() != p.n) {
                throw std::runtime_error("pack_graph: mobility coefficient vector has the wrong size");
            }
            bool nonzero = false;
            for (int j = 0; j < p.n; ++j) {
                c[size_t(j)] += grp[mbr].get_baseline()[j];
                nonzero = nonzero || grp[mbr].get_baseline()[j] != 0.0;
            }
            if (!grp[mbr].get_dampings().empty() && nonzero) {
                const auto& ds = grp[mbr].get_dampings();
                const int Kd   = int(ds.size());
                std::vector<double> times, mats;
                std::vector<int32_t> levels, types;
                for (auto& d : ds) {
                    times.push_back(d.time);
                    levels.push_back(d.level);
                    types.push_back(d.type);
                    mats.insert(mats.end(), d.v.data(), d.v.data() + p.n);
                }
                const size_t Ku = size_t(Kd);
                std::vector<double> ot(Ku), om(Ku * size_t(p.n));
                int32_t nout = 0;
                cuda_check(mio_damping_cache_build(p.n, 1, Kd, times.data(), levels.data(), types.data(), mats.data(),
                                                   &nout, ot.data(), om.data()));
                const size_t tsz = size_t(1 + Kd + Kd * p.n);
                std::vector<double> tbl(tsz, 0.0);
                tbl[0] = nout;
                std::copy(ot.begin(), ot.begin() + nout, tbl.begin() + 1);
                std::copy(om.begin(), om.begin() + size_t(nout) * p.n, tbl.begin() + 1 + Kd);
                if (have_mob && (p.mob_max_pts != Kd || tbl != p.mobtable)) {
                    throw std::runtime_error("pack_graph: edges with different mobility dampings are not supported "
                                             "(one mobility damping table per run)");
                }
                p.mob_max_pts = Kd;
                p.mobtable    = std::move(tbl);
                have_mob      = true;
            }
        }
        p.edge_coeff.insert(p.edge_coeff.end(), c.begin(), c.end());
    }
    return p;
}

// device handle shared by a batch of runs with the same graph
class DeviceGraph
{
public:
    DeviceGraph(const PackedGraph& p, double dt_graph, int max_runs, int num_psets, int num_ctables, int record_days,
                int device = 0)
        : m_n(p.n)
        , m_N(p.N)
        , m_days(record_days)
    {
        mio_graph_desc d{};
        d.model        = p.model_kind;
        d.sim_kind     = p.sim_kind;
        d.vacc_days    = p.vacc_days;
        d.vacc_partial = p.vacc_days > 0 ? p.vacc_partial.data() : nullptr;
        d.vacc_full    = p.vacc_days > 0 ? p.vacc_full.data() : nullptr;
        d.num_groups   = p.G;
        d.num_nodes    = p.N;
        d.num_edges    = p.E;
        d.num_psets    = num_psets;
        d.num_ctables  = num_ctables;
        d.ct_max_loc   = p.ct_max_loc;
        d.ct_max_pts   = p.ct_max_pts;
        d.mob_max_pts  = p.mob_max_pts;
        d.max_runs     = max_runs;
        d.record_daily = record_days > 0;
        d.max_days     = record_days;
        d.edge_start   = p.edge_start.data();
        d.edge_end     = p.edge_end.data();
        d.edge_coeff   = p.edge_coeff.data();
        d.node_pset    = p.node_pset.data();
        d.node_ctable  = p.node_ctable.data();
        d.integrator   = p.integrator;
        d.t0           = p.t0;
        d.dt_graph     = dt_graph;
        d.flow_simulation = p.flow_simulation;
        cuda_check(mio_cuda_graph_create(&d, device, &m_h));
    }
    DeviceGraph(const DeviceGraph&) = delete;
    DeviceGraph& operator=(const DeviceGraph&) = delete;
    ~DeviceGraph()
    {
        mio_cuda_graph_destroy(m_h);
    }
    mio_handle* get() const
    {
        return m_h;
    }
    // upload run `slot` from a packed graph (padding parameter sets / tables up to the handle's counts)
    void set_run(int slot, const PackedGraph& p, int num_psets, int num_ctables)
    {
        std::vector<double> ps, ct;
        for (int i = 0; i < num_psets; ++i) {
            const auto& s = p.psets[size_t(std::min<int>(i, int(p.psets.size()) - 1))];
            ps.insert(ps.end(), s.begin(), s.end());
        }
        for (int i = 0; i < num_ctables; ++i) {
            const auto& s = p.ctables[size_t(std::min<int>(i, int(p.ctables.size()) - 1))];
            ct.insert(ct.end(), s.begin(), s.end());
        }
        mio_run_data rd{ps.data(),       ct.data(),        p.node_local.data(), p.y0.data(),
                        p.dt0.data(),    p.mob_max_pts > 0 ? p.mobtable.data() : nullptr};
        cuda_check(mio_cuda_set_runs(m_h, slot, 1, &rd));
        if (p.flow_simulation) {
            cuda_check(mio_cuda_set_flows(m_h, slot, 1, p.flows0.data()));
        }
    }
    int n() const
    {
        return m_n;
    }
    int N() const
    {
        return m_N;
    }

private:
    mio_handle* m_h = nullptr;
    int m_n, m_N, m_days;
};

} // namespace detail

// GraphSimulation, graph_simulation.h:41-121 (advance :101-120)
template <class FP, class GraphT>
class GraphSimulation
{
public:
    using Graph = GraphT;
    GraphSimulation(FP t0, FP dt, const GraphT& g)
        : m_graph(g)
        , m_t(t0)
        , m_dt(dt)
    {
    }
    GraphSimulation(FP t0, FP dt, GraphT&& g)
        : m_graph(std::move(g))
        , m_t(t0)
        , m_dt(dt)
    {
    }
    // one run: every graph step's node values (after the edge phase) become a time point of the node results
    void advance(FP t_max = 1.0)
    {
        if (!m_dev) {
            auto p = detail::pack_graph<FP>(m_graph);
            m_dev  = std::make_shared<detail::DeviceGraph>(p, m_dt, 1, int(p.psets.size()), int(p.ctables.size()), 0);
            m_dev->set_run(0, p, int(p.psets.size()), int(p.ctables.size()));
            m_ct_shape = {p.ct_max_loc, p.ct_max_pts, int(p.ctables.size())};
            // MobilityEdge(coeffs, save_indices): the backend records one set of index groups for all edges
            for (auto& e : m_graph.edges()) {
                const auto& si = e.property.get_parameters().get_save_indices();
                if (!si.empty()) {
                    if (m_save.empty()) {
                        m_save = si;
                    }
                    else if (m_save != si) {
                        throw std::runtime_error("GraphSimulation: edges with different saved_compartment_indices are "
                                                 "not supported by this backend");
                    }
                }
            }
            for (auto& e : m_graph.edges()) {
                if (!m_save.empty() && e.property.get_parameters().get_save_indices().empty()) {
                    throw std::runtime_error("GraphSimulation: either all edges or none must save mobility results");
                }
            }
        }
        if (!m_save.empty()) {
            // room for the graph steps of this call
            const int steps = int(std::ceil((t_max - m_t) / m_dt)) + 2;
            std::vector<int32_t> ptr(1, 0), idx;
            for (auto& grp : m_save) {
                for (auto i : grp) {
                    idx.push_back(int32_t(i));
                }
                ptr.push_back(int32_t(idx.size()));
            }
            m_mob_cap = m_mob_seen + steps;
            cuda_check(mio_cuda_enable_mobility_results(m_dev->get(), int32_t(m_save.size()), ptr.data(), idx.data(),
                                                        m_mob_cap));
        }
        const size_t total = size_t(m_dev->N()) * size_t(m_dev->n());
        std::vector<double> state(total);
        FP dt = m_dt;
        while (m_t < t_max) {
            if (m_t + dt > t_max) {
                dt = t_max - m_t;
            }
            update_dynamic_npis(m_t, m_t + dt);
            m_t += dt;
            cuda_check(mio_cuda_advance(m_dev->get(), m_t));
            cuda_check(mio_cuda_get_state(m_dev->get(), 0, 1, state.data()));
            append_time_point(m_t, state.data());
            append_flow_time_point(m_t);
        }
        if (!m_save.empty()) {
            const size_t E = m_graph.edges().size(), K1 = m_save.size() + 1, cap = size_t(m_mob_cap);
            std::vector<int32_t> npts(E);
            std::vector<double> times(cap), values(E * cap * K1);
            cuda_check(mio_cuda_get_mobility_results(m_dev->get(), 0, 1, npts.data(), times.data(), values.data()));
            int seen = m_mob_seen;
            for (size_t e = 0; e < E; ++e) {
                auto& ts = m_graph.edges()[e].property.get_mobility_results();
                for (int i = m_mob_seen; i < npts[e]; ++i) {
                    Eigen::VectorX<FP> v{Eigen::Index(K1)};
                    std::copy(values.begin() + (e * cap + size_t(i)) * K1, values.begin() + (e * cap + size_t(i) + 1) * K1,
                              v.data());
                    ts.add_time_point(times[size_t(i)], v);
                }
                seen = std::max(seen, int(npts[e]));
            }
            m_mob_seen = seen;
        }
        const size_t Nn = size_t(m_dev->N());
        std::vector<double> dts(Nn);
        cuda_check(mio_cuda_get_node_dt(m_dev->get(), 0, 1, dts.data()));
        for (size_t k = 0; k < m_graph.nodes().size(); ++k) {
            m_graph.nodes()[k].property.get_simulation().get_dt() = dts[k];
        }
    }
    FP get_t() const
    {
        return m_t;
    }
    GraphT& get_graph() &
    {
        return m_graph;
    }
    const GraphT& get_graph() const&
    {
        return m_graph;
    }
    GraphT&& get_graph() &&
    {
        return std::move(m_graph);
    }
    // used by ParameterStudy: results computed in a batch handle are written back here
    void append_time_point(FP t, const double* state)
    {
        const int n = int(m_graph.nodes()[0].property.get_result().get_num_elements());
        for (size_t k = 0; k < m_graph.nodes().size(); ++k) {
            Eigen::VectorX<FP> v(n);
            std::copy(state + k * n, state + (k + 1) * n, v.data());
            m_graph.nodes()[k].property.get_result().add_time_point(t, v);
        }
    }
    // dynamic NPIs of the nodes (ode_secir/model.h:299-353): every node takes its decision before the graph step; the
    // contact tables of nodes whose contact patterns changed are replaced on the device
    void update_dynamic_npis(FP t, FP t_step_end)
    {
        using Sim    = std::decay_t<decltype(m_graph.nodes()[0].property.get_simulation())>;
        using Traits = detail::SimTraits<Sim>;
        bool any = false, changed = false;
        for (auto& e : m_graph.edges()) {
            if (!e.property.get_parameters().get_dynamic_npis_infected().get_thresholds().empty()) {
                throw std::runtime_error("GraphSimulation: dynamic NPIs on mobility edges are not supported by this "
                                         "backend (one mobility damping table per run)");
            }
        }
        for (auto& nd : m_graph.nodes()) {
            any = any || Traits::has_dynamic_npis(nd.property.get_simulation().get_model());
        }
        if (!any) {
            return;
        }
        if (m_dt > 1.0) {
            throw std::runtime_error("GraphSimulation: dynamic NPIs need a graph step of at most one day");
        }
        for (auto& nd : m_graph.nodes()) {
            changed = Traits::check_dynamic_npis(nd.property.get_simulation(), t, t_step_end) || changed;
        }
        if (changed) {
            auto p = detail::pack_graph<FP>(m_graph);
            if (m_ct_shape != std::array<int, 3>{p.ct_max_loc, p.ct_max_pts, int(p.ctables.size())}) {
                throw std::runtime_error("GraphSimulation: the dynamic NPIs added more contact change points than the "
                                         "device tables hold");
            }
            std::vector<double> ct;
            for (auto& tbl : p.ctables) {
                ct.insert(ct.end(), tbl.begin(), tbl.end());
            }
            cuda_check(mio_cuda_set_contact_tables(m_dev->get(), 0, 1, ct.data()));
        }
    }
    // graphs of FlowSimulation nodes: the flows of every node at the end of the graph step
    void append_flow_time_point(FP t)
    {
        using Sim = std::decay_t<decltype(m_graph.nodes()[0].property.get_simulation())>;
        if constexpr (detail::SimTraits<Sim>::flow_simulation) {
            const size_t nf = size_t(m_graph.nodes()[0].property.get_simulation().get_flows().get_num_elements());
            std::vector<double> flows(m_graph.nodes().size() * nf);
            cuda_check(mio_cuda_get_flows(m_dev->get(), 0, 1, flows.data()));
            for (size_t k = 0; k < m_graph.nodes().size(); ++k) {
                Eigen::VectorX<FP> v{Eigen::Index(nf)};
                std::copy(flows.begin() + k * nf, flows.begin() + (k + 1) * nf, v.data());
                m_graph.nodes()[k].property.get_simulation().get_flows().add_time_point(t, v);
            }
        }
    }
    void set_t(FP t)
    {
        m_t = t;
    }
    FP get_dt() const
    {
        return m_dt;
    }

private:
    GraphT m_graph;
    FP m_t, m_dt;
    std::shared_ptr<detail::DeviceGraph> m_dev;
    std::vector<std::vector<size_t>> m_save; // saved_compartment_indices common to the edges
    int m_mob_cap = 0, m_mob_seen = 0;
    std::array<int, 3> m_ct_shape{0, 0, 0}; // contact table shape of the device handle (locations, points, tables)
};

// make_mobility_sim, metapopulation_mobility_instant.h:656-678
template <class FP, class Sim>
auto make_mobility_sim(FP t0, FP dt, const Graph<SimulationNode<FP, Sim>, MobilityEdge<FP>>& graph)
{
    return GraphSimulation<FP, Graph<SimulationNode<FP, Sim>, MobilityEdge<FP>>>(t0, dt, graph);
}
template <class FP, class Sim>
auto make_mobility_sim(FP t0, FP dt, Graph<SimulationNode<FP, Sim>, MobilityEdge<FP>>&& graph)
{
    return GraphSimulation<FP, Graph<SimulationNode<FP, Sim>, MobilityEdge<FP>>>(t0, dt, std::move(graph));
}

} // namespace mio

// flows -- mio::FlowSimulation through the facade:
//   "single": mio::simulate_flows of the three-age-group scenario of the reference's golden tests (Cash-Karp
//             controller, 50 days), advanced in two calls; prints "t y0..y29 | f0..f44" for every time point;
//   "graph":  two FlowSimulation nodes exchanging commuters (the scenario of graph_two_nodes), 10 days; prints
//             "t node y0..y9 | f0..f14" after every graph step.
#include "memilio_b200/mio.hpp"

#include <cstdio>
#include <cstring>

using namespace mio::osecir;

static void set_rates(Model<double>& m, double transmission)
{
    auto& p = m.parameters;
    p.set<StartDay<double>>(60);
    p.set<Seasonality<double>>(0.2);
    p.get<TimeExposed<double>>()                       = 3.2;
    p.get<TimeInfectedNoSymptoms<double>>()            = 2.0;
    p.get<TimeInfectedSymptoms<double>>()              = 5.8;
    p.get<TimeInfectedSevere<double>>()                = 9.5;
    p.get<TimeInfectedCritical<double>>()              = 7.1;
    p.get<TransmissionProbabilityOnContact<double>>()  = transmission;
    p.get<RelativeTransmissionNoSymptoms<double>>()    = 0.7;
    p.get<RecoveredPerInfectedNoSymptoms<double>>()    = 0.09;
    p.get<RiskOfInfectionFromSymptomatic<double>>()    = 0.25;
    p.get<MaxRiskOfInfectionFromSymptomatic<double>>() = 0.45;
    p.get<TestAndTraceCapacity<double>>()              = 35;
    p.get<SeverePerInfectedSymptoms<double>>()         = 0.2;
    p.get<CriticalPerSevere<double>>()                 = 0.25;
    p.get<DeathsPerCritical<double>>()                 = 0.3;
}

static void print_rows(const mio::TimeSeries<double>& y, const mio::TimeSeries<double>& f, int node)
{
    for (Eigen::Index i = 0; i < y.get_num_time_points(); ++i) {
        std::printf("%.17g", y.get_time(i));
        if (node >= 0) {
            std::printf(" %d", node);
        }
        for (Eigen::Index c = 0; c < y.get_num_elements(); ++c) {
            std::printf(" %.17g", y[i][c]);
        }
        for (Eigen::Index c = 0; c < f.get_num_elements(); ++c) {
            std::printf(" %.17g", f[i][c]);
        }
        std::printf("\n");
    }
}

static int single()
{
    const int G = 3;
    Model<double> model(G);
    set_rates(model, 0.05);
    const double fact = 1.0 / G;
    for (auto i = mio::AgeGroup(0); i < mio::AgeGroup(G); i++) {
        model.populations[{i, InfectionState::Exposed}]            = fact * 100;
        model.populations[{i, InfectionState::InfectedNoSymptoms}] = fact * 50;
        model.populations[{i, InfectionState::InfectedSymptoms}]   = fact * 50;
        model.populations[{i, InfectionState::InfectedSevere}]     = fact * 20;
        model.populations[{i, InfectionState::InfectedCritical}]   = fact * 10;
        model.populations[{i, InfectionState::Recovered}]          = fact * 10;
        model.populations[{i, InfectionState::Dead}]               = 0;
        model.populations.set_difference_from_group_total<mio::AgeGroup>({i, InfectionState::Susceptible},
                                                                          fact * 10000);
    }
    mio::ContactMatrixGroup<double>& cm = model.parameters.get<ContactPatterns<double>>();
    cm[0] = mio::ContactMatrix<double>(Eigen::MatrixX<double>::Constant(G, G, fact * 10));
    cm[0].add_damping(0.7, mio::SimulationTime<double>(30.));
    model.apply_constraints();

    mio::FlowSimulation<double, Model<double>> sim(model, 0.0, 0.1);
    sim.advance(20.0);
    sim.advance(50.0);
    print_rows(sim.get_result(), sim.get_flows(), -1);
    // the same through simulate_flows (one advance call): only its last row, marked by a leading "#"
    auto res = mio::simulate_flows<double>(0.0, 50.0, 0.1, model);
    std::printf("# %td %.17g %.17g\n", res[1].get_num_time_points(), res[0].get_last_value()[0],
                res[1].get_last_value()[model.get_flat_flow_index<InfectionState::Susceptible, InfectionState::Exposed>(
                    mio::AgeGroup(0))]);
    return 0;
}

static int graph()
{
    using Sim = mio::FlowSimulation<double, Model<double>>;
    Model<double> base(1);
    set_rates(base, 0.1);
    base.parameters.set<StartDay<double>>(0);
    mio::ContactMatrixGroup<double>& cm = base.parameters.get<ContactPatterns<double>>();
    cm[0] = mio::ContactMatrix<double>(Eigen::MatrixX<double>::Constant(1, 1, 10));
    base.populations[{mio::AgeGroup(0), InfectionState::Susceptible}] = 10000;
    Model<double> first = base, second = base;
    mio::ContactMatrixGroup<double>& cm1 = first.parameters.get<ContactPatterns<double>>();
    cm1[0].add_damping(0.7, mio::SimulationTime<double>(15.0));
    first.populations[{mio::AgeGroup(0), InfectionState::Susceptible}] = 9990;
    first.populations[{mio::AgeGroup(0), InfectionState::Exposed}]     = 100;

    mio::Graph<mio::SimulationNode<double, Sim>, mio::MobilityEdge<double>> g;
    g.add_node(1, first, 0.0);
    g.add_node(2, second, 0.0);
    const auto coeff = Eigen::VectorX<double>::Constant(Eigen::Index(InfectionState::Count), 0.1);
    g.add_edge(0, 1, coeff);
    g.add_edge(1, 0, coeff);
    auto sim = mio::make_mobility_sim<double>(0.0, 0.5, std::move(g));
    sim.advance(10.0);
    for (size_t k = 0; k < sim.get_graph().nodes().size(); ++k) {
        const auto& node = sim.get_graph().nodes()[k].property;
        print_rows(node.get_result(), node.get_simulation().get_flows(), int(k));
    }
    return 0;
}

int main(int argc, char** argv)
{
    if (argc > 1 && std::strcmp(argv[1], "graph") == 0) {
        return graph();
    }
    return single();
}

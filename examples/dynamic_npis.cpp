// dynamic_npis -- DynamicNPIsInfectedSymptoms through the facade (cpp/models/ode_secir/model.h:289-367):
//   "single": one node, 10 of 100 symptomatic, contact frequency 10; above 1000 per 50 000 symptomatic infections the
//             contacts are reduced by 0.8 for 4.5 days (renewed while the threshold stays exceeded); 12 days.
//             Prints "t y0..y9" for every time point, then "# c(t)" lines: the contact matrix entry at t = 0..13.
//   "predefined": the same with predefined dampings on the contact matrix (same slot as the NPI and another one).
//   "graph":  the two nodes of graph_two_nodes with such NPIs (thresholds crossed during the run); 10 days; prints
//             "t node y0..y9" after every graph step and "# node c(t)" lines.
#include "memilio_b200/mio.hpp"

#include <cstdio>
#include <cstring>

using namespace mio::osecir;

static mio::DynamicNPIs<double> make_npis(double threshold, double value, double duration)
{
    mio::DynamicNPIs<double> npis;
    npis.set_threshold(threshold, {mio::DampingSampling<double>(value, mio::DampingLevel(0), mio::DampingType(0),
                                                                mio::SimulationTime<double>(0), {0},
                                                                Eigen::VectorX<double>::Constant(1, 1.0))});
    npis.set_duration(mio::SimulationTime<double>(duration));
    npis.set_base_value(50000);
    return npis;
}

static int single(bool with_predefined)
{
    Model<double> model(1);
    model.populations[{mio::AgeGroup(0), InfectionState::InfectedSymptoms}] = 10;
    model.populations.set_difference_from_total({mio::AgeGroup(0), InfectionState::Susceptible}, 100);
    mio::ContactMatrixGroup<double>& cm = model.parameters.get<ContactPatterns<double>>();
    cm[0] = mio::ContactMatrix<double>(Eigen::MatrixX<double>::Constant(1, 1, 10.0));
    if (with_predefined) {
        // predefined dampings of the NPI's (level, type) slot inside and after the NPI window, and one of another slot:
        // implement_dynamic_npis has to raise, keep and restore them
        cm[0].add_damping(0.5, mio::DampingLevel(0), mio::DampingType(0), mio::SimulationTime<double>(3.0));
        cm[0].add_damping(0.9, mio::DampingLevel(0), mio::DampingType(0), mio::SimulationTime<double>(8.0));
        cm[0].add_damping(0.3, mio::DampingLevel(1), mio::DampingType(1), mio::SimulationTime<double>(6.0));
    }
    model.parameters.get<DynamicNPIsInfectedSymptoms<double>>() = make_npis(1000.0, 0.8, 4.5);
    Simulation<double> sim(model, 0.0, 0.1);
    sim.advance(12.0);
    const auto& res = sim.get_result();
    for (Eigen::Index i = 0; i < res.get_num_time_points(); ++i) {
        std::printf("%.17g", res.get_time(i));
        for (Eigen::Index c = 0; c < res.get_num_elements(); ++c) {
            std::printf(" %.17g", res[i][c]);
        }
        std::printf("\n");
    }
    const mio::ContactMatrixGroup<double>& after = sim.get_model().parameters.get<ContactPatterns<double>>();
    for (int t = 0; t <= 13; ++t) {
        std::printf("# %.17g\n", after.get_matrix_at(mio::SimulationTime<double>(t))(0, 0));
    }
    return 0;
}

static int graph()
{
    using Sim = Simulation<double>;
    Model<double> base(1);
    auto& p = base.parameters;
    p.set<StartDay<double>>(0);
    p.set<Seasonality<double>>(0.2);
    p.get<TimeExposed<double>>()                       = 3.2;
    p.get<TimeInfectedNoSymptoms<double>>()            = 2.0;
    p.get<TimeInfectedSymptoms<double>>()              = 5.8;
    p.get<TimeInfectedSevere<double>>()                = 9.5;
    p.get<TimeInfectedCritical<double>>()              = 7.1;
    p.get<TransmissionProbabilityOnContact<double>>()  = 0.1;
    p.get<RelativeTransmissionNoSymptoms<double>>()    = 0.7;
    p.get<RecoveredPerInfectedNoSymptoms<double>>()    = 0.09;
    p.get<RiskOfInfectionFromSymptomatic<double>>()    = 0.25;
    p.get<MaxRiskOfInfectionFromSymptomatic<double>>() = 0.45;
    p.get<TestAndTraceCapacity<double>>()              = 35;
    p.get<SeverePerInfectedSymptoms<double>>()         = 0.2;
    p.get<CriticalPerSevere<double>>()                 = 0.25;
    p.get<DeathsPerCritical<double>>()                 = 0.3;
    mio::ContactMatrixGroup<double>& cm = p.get<ContactPatterns<double>>();
    cm[0] = mio::ContactMatrix<double>(Eigen::MatrixX<double>::Constant(1, 1, 10));
    base.populations[{mio::AgeGroup(0), InfectionState::Susceptible}] = 10000;
    Model<double> first = base, second = base;
    first.populations[{mio::AgeGroup(0), InfectionState::Susceptible}]      = 9900;
    first.populations[{mio::AgeGroup(0), InfectionState::Exposed}]          = 100;
    first.populations[{mio::AgeGroup(0), InfectionState::InfectedSymptoms}] = 50;
    // 50 symptomatic of 10 050 are 249 per 50 000: the threshold of 400 is crossed after some days in the first node,
    // the second node (threshold 100) follows once the commuters have carried the infection over
    first.parameters.get<DynamicNPIsInfectedSymptoms<double>>()  = make_npis(400.0, 0.6, 4.0);
    second.parameters.get<DynamicNPIsInfectedSymptoms<double>>() = make_npis(100.0, 0.5, 3.0);

    mio::Graph<mio::SimulationNode<double, Sim>, mio::MobilityEdge<double>> g;
    g.add_node(1, first, 0.0);
    g.add_node(2, second, 0.0);
    const auto coeff = Eigen::VectorX<double>::Constant(Eigen::Index(InfectionState::Count), 0.1);
    g.add_edge(0, 1, coeff);
    g.add_edge(1, 0, coeff);
    auto sim = mio::make_mobility_sim<double>(0.0, 0.5, std::move(g));
    sim.advance(10.0);
    for (size_t k = 0; k < sim.get_graph().nodes().size(); ++k) {
        const auto& res = sim.get_graph().nodes()[k].property.get_result();
        for (Eigen::Index i = 0; i < res.get_num_time_points(); ++i) {
            std::printf("%.17g %zu", res.get_time(i), k);
            for (Eigen::Index c = 0; c < res.get_num_elements(); ++c) {
                std::printf(" %.17g", res[i][c]);
            }
            std::printf("\n");
        }
    }
    for (size_t k = 0; k < sim.get_graph().nodes().size(); ++k) {
        const mio::ContactMatrixGroup<double>& after =
            sim.get_graph().nodes()[k].property.get_simulation().get_model().parameters.get<ContactPatterns<double>>();
        for (int t = 0; t <= 11; ++t) {
            std::printf("# %zu %.17g\n", k, after.get_matrix_at(mio::SimulationTime<double>(t))(0, 0));
        }
    }
    return 0;
}

int main(int argc, char** argv)
{
    if (argc > 1 && std::strcmp(argv[1], "graph") == 0) {
        return graph();
    }
    if (argc > 1 && std::strcmp(argv[1], "predefined") == 0) {
        return single(true);
    }
    return single(false);
}

// graph_two_nodes -- BASELINE config 2: two ODE-SECIR nodes exchanging commuters (coefficient 0.1 both ways), graph
// step 0.5 days, 30 days, a contact damping of 0.7 from day 15 in the first node only (the scenario of the reference's
// cpp/examples/ode_secir_graph.cpp). Prints "t node y0 .. y9" after every graph step.
#include "memilio_b200/mio.hpp"

#include <cstdio>

int main()
{
    using namespace mio::osecir;
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
    sim.advance(30.0);

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
    return 0;
}

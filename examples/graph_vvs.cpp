// graph_vvs -- three ODE-SECIRVVS nodes (the model of cpp/examples/ode_secirvvs.cpp, one age group) exchanging
// commuters, graph step 0.5 days, 10 days; node type osecirvvs::Simulation (day chunks, vaccination, the model's
// mobility hooks) -- the node / edge types of BASELINE configs[4] through the C++ facade. Prints "node y0 .. y26".
#include "memilio_b200/mio.hpp"

#include <cstdio>

int main()
{
    using namespace mio::osecirvvs;
    using IS  = InfectionState;
    using Sim = Simulation<double>;
    const mio::AgeGroup a0(0);
    Model<double> model(1);
    auto& pop = model.populations;
    pop[{a0, IS::ExposedNaive}] = 10, pop[{a0, IS::ExposedImprovedImmunity}] = 11, pop[{a0, IS::ExposedPartialImmunity}] = 12;
    pop[{a0, IS::InfectedNoSymptomsNaive}] = 13, pop[{a0, IS::InfectedNoSymptomsNaiveConfirmed}] = 13;
    pop[{a0, IS::InfectedNoSymptomsPartialImmunity}] = 14, pop[{a0, IS::InfectedNoSymptomsPartialImmunityConfirmed}] = 14;
    pop[{a0, IS::InfectedNoSymptomsImprovedImmunity}] = 15, pop[{a0, IS::InfectedNoSymptomsImprovedImmunityConfirmed}] = 15;
    pop[{a0, IS::InfectedSymptomsNaive}] = 5, pop[{a0, IS::InfectedSymptomsNaiveConfirmed}] = 5;
    pop[{a0, IS::InfectedSymptomsPartialImmunity}] = 6, pop[{a0, IS::InfectedSymptomsPartialImmunityConfirmed}] = 6;
    pop[{a0, IS::InfectedSymptomsImprovedImmunity}] = 7, pop[{a0, IS::InfectedSymptomsImprovedImmunityConfirmed}] = 7;
    pop[{a0, IS::InfectedSevereNaive}] = 8, pop[{a0, IS::InfectedSevereImprovedImmunity}] = 1;
    pop[{a0, IS::InfectedSeverePartialImmunity}] = 2, pop[{a0, IS::InfectedCriticalNaive}] = 3;
    pop[{a0, IS::InfectedCriticalPartialImmunity}] = 4, pop[{a0, IS::InfectedCriticalImprovedImmunity}] = 5;
    pop[{a0, IS::SusceptibleImprovedImmunity}] = 6, pop[{a0, IS::SusceptiblePartialImmunity}] = 7;
    pop.set_difference_from_group_total<mio::AgeGroup>({a0, IS::SusceptibleNaive}, 1000);

    auto& p = model.parameters;
    p.get<ICUCapacity<double>>()          = 100;
    p.get<TestAndTraceCapacity<double>>() = 0.0143;
    p.get<DailyPartialVaccinations<double>>().resize(mio::SimulationDay(size_t(31)));
    p.get<DailyFullVaccinations<double>>().resize(mio::SimulationDay(size_t(31)));
    for (size_t d = 0; d < 31; ++d) {
        p.get<DailyPartialVaccinations<double>>()[{a0, mio::SimulationDay(d)}] = double(d * 10);
        p.get<DailyFullVaccinations<double>>()[{a0, mio::SimulationDay(d)}]    = double(d * 10);
    }
    auto& cm = p.get<ContactPatterns<double>>().get_cont_freq_mat();
    cm[0].get_baseline().setConstant(5.0);
    cm[0].add_damping(0.3, mio::SimulationTime<double>(5.0));
    p.get<TimeExposed<double>>()[a0]                       = 3.33;
    p.get<TimeInfectedNoSymptoms<double>>()[a0]            = 1.87;
    p.get<TimeInfectedSymptoms<double>>()[a0]              = 7;
    p.get<TimeInfectedSevere<double>>()[a0]                = 6;
    p.get<TimeInfectedCritical<double>>()[a0]              = 7;
    p.get<TransmissionProbabilityOnContact<double>>()[a0]  = 0.15;
    p.get<RelativeTransmissionNoSymptoms<double>>()[a0]    = 0.5;
    p.get<RiskOfInfectionFromSymptomatic<double>>()[a0]    = 0.0;
    p.get<MaxRiskOfInfectionFromSymptomatic<double>>()[a0] = 0.4;
    p.get<RecoveredPerInfectedNoSymptoms<double>>()[a0]    = 0.2;
    p.get<SeverePerInfectedSymptoms<double>>()[a0]         = 0.1;
    p.get<CriticalPerSevere<double>>()[a0]                 = 0.1;
    p.get<DeathsPerCritical<double>>()[a0]                 = 0.1;
    p.get<ReducExposedPartialImmunity<double>>()[a0]                     = 0.8;
    p.get<ReducExposedImprovedImmunity<double>>()[a0]                    = 0.331;
    p.get<ReducInfectedSymptomsPartialImmunity<double>>()[a0]            = 0.65;
    p.get<ReducInfectedSymptomsImprovedImmunity<double>>()[a0]           = 0.243;
    p.get<ReducInfectedSevereCriticalDeadPartialImmunity<double>>()[a0]  = 0.1;
    p.get<ReducInfectedSevereCriticalDeadImprovedImmunity<double>>()[a0] = 0.091;
    p.get<ReducTimeInfectedMild<double>>()[a0]                           = 0.9;
    p.get<Seasonality<double>>()                                         = 0.2;
    model.apply_constraints();

    mio::Graph<mio::SimulationNode<double, Sim>, mio::MobilityEdge<double>> g;
    for (int k = 0; k < 3; ++k) {
        Model<double> m = model;
        m.populations[{a0, IS::ExposedNaive}] = 10.0 + 20.0 * k; // the nodes differ in their exposed population
        m.populations.set_difference_from_group_total<mio::AgeGroup>({a0, IS::SusceptibleNaive}, 1000);
        g.add_node(k, m, 0.0);
    }
    const auto coeff = Eigen::VectorX<double>::Constant(Eigen::Index(IS::Count), 0.05);
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) {
            if (i != j) {
                g.add_edge(size_t(i), size_t(j), coeff);
            }
        }
    }
    auto sim = mio::make_mobility_sim<double>(0.0, 0.5, std::move(g));
    sim.advance(10.0);
    for (size_t k = 0; k < sim.get_graph().nodes().size(); ++k) {
        const auto& res = sim.get_graph().nodes()[k].property.get_result();
        std::printf("%zu", k);
        for (Eigen::Index c = 0; c < res.get_num_elements(); ++c) {
            std::printf(" %.17g", res.get_last_value()[c]);
        }
        std::printf("\n");
    }
    return 0;
}

// parameter_study -- a small Graph-ODE-SECIR ParameterStudy through the facade: 3 counties, 4 commuter edges, 6 age
// groups, uniform parameter distributions, seeds {1,..,6}; usage: parameter_study [num_runs] [days].
// Prints, per run, "run node day y0 .. y59" of the interpolated result, then the p50 table ("p50 node day ...").
#include "memilio_b200/mio.hpp"

#include <cstdio>
#include <cstdlib>

int main(int argc, char** argv)
{
    using namespace mio::osecir;
    const size_t num_runs = argc > 1 ? size_t(std::atoi(argv[1])) : 6;
    const double tmax     = argc > 2 ? std::atof(argv[2]) : 10.0;
    const int G           = 6;
    Model<double> model(G);
    auto uni = [](mio::UncertainValue<double>& v, double lo, double hi) {
        v = mio::UncertainValue<double>(0.5 * (lo + hi));
        v.set_distribution(mio::ParameterDistributionUniform(lo, hi));
    };
    auto& p = model.parameters;
    uni(p.get<Seasonality<double>>(), 0.1, 0.3);
    const double tsev_lo[] = {3.925, 3.925, 4.85, 6.4, 7.2, 9.}, tsev_hi[] = {6.075, 6.075, 7., 8.7, 9.8, 13.};
    const double prob_lo[] = {0.02, 0.05, 0.05, 0.05, 0.08, 0.15}, prob_hi[] = {0.04, 0.07, 0.07, 0.07, 0.10, 0.20};
    for (mio::AgeGroup i = 0; i < mio::AgeGroup(G); ++i) {
        uni(p.get<TimeExposed<double>>()[i], 2.67, 4.0);
        uni(p.get<TimeInfectedNoSymptoms<double>>()[i], 1.2, 2.53);
        uni(p.get<TimeInfectedSymptoms<double>>()[i], 5.6, 8.4);
        uni(p.get<TimeInfectedSevere<double>>()[i], tsev_lo[i.value], tsev_hi[i.value]);
        uni(p.get<TimeInfectedCritical<double>>()[i], 4.95, 8.95);
        uni(p.get<TransmissionProbabilityOnContact<double>>()[i], prob_lo[i.value], prob_hi[i.value]);
        uni(p.get<RelativeTransmissionNoSymptoms<double>>()[i], 0.5, 1.0);
        uni(p.get<RiskOfInfectionFromSymptomatic<double>>()[i], 0.1, 0.3);
        uni(p.get<MaxRiskOfInfectionFromSymptomatic<double>>()[i], 0.3, 0.5);
        uni(p.get<RecoveredPerInfectedNoSymptoms<double>>()[i], 0.15, 0.3);
        uni(p.get<SeverePerInfectedSymptoms<double>>()[i], 0.006, 0.2);
        uni(p.get<CriticalPerSevere<double>>()[i], 0.05, 0.35);
        uni(p.get<DeathsPerCritical<double>>()[i], 0.0, 0.5);
    }
    mio::ContactMatrixGroup<double>& cm = p.get<ContactPatterns<double>>();
    cm[0] = mio::ContactMatrix<double>(Eigen::MatrixX<double>::Constant(G, G, 10.0 / G));
    // one sampled contact damping at day 4 that also damps the commuter flows (parameter_space.h:241-248)
    p.get<ContactPatterns<double>>().get_dampings().emplace_back(
        mio::UncertainValue<double>(0.5, mio::ParameterDistributionUniform(0.3, 0.7)), mio::DampingLevel(0),
        mio::DampingType(0), mio::SimulationTime<double>(4.0), std::vector<size_t>{0},
        Eigen::VectorX<double>::Constant(G, 1.0));
    for (mio::AgeGroup i = 0; i < mio::AgeGroup(G); ++i) {
        model.populations[{i, InfectionState::Exposed}] = 20;
        model.populations[{i, InfectionState::Exposed}].set_distribution(mio::ParameterDistributionUniform(10, 30));
        model.populations.set_difference_from_group_total<mio::AgeGroup>({i, InfectionState::Susceptible}, 10000);
    }
    uni(p.get<TestAndTraceCapacity<double>>(), 20, 40);

    mio::Graph<Model<double>, mio::MobilityParameters<double>> params_graph;
    params_graph.add_node(1001, model);
    params_graph.add_node(1002, model);
    params_graph.add_node(1003, model);
    const Eigen::Index n = G * Eigen::Index(InfectionState::Count);
    params_graph.add_edge(0, 1, Eigen::VectorX<double>::Constant(n, 0.05));
    params_graph.add_edge(1, 0, Eigen::VectorX<double>::Constant(n, 0.1));
    params_graph.add_edge(1, 2, Eigen::VectorX<double>::Constant(n, 0.15));
    params_graph.add_edge(2, 1, Eigen::VectorX<double>::Constant(n, 0.2));

    mio::ParameterStudy study(params_graph, 0.0, tmax, 0.5, num_runs);
    study.get_rng().seed({1, 2, 3, 4, 5, 6});
    auto ensemble = study.run(
        [](auto&& graph, double t0, double dt, size_t) {
            auto copy = graph;
            return mio::make_sampled_graph_simulation<double, Simulation<double>>(draw_sample(copy), t0, dt, dt);
        },
        [](auto&& sim, size_t) {
            return mio::interpolate_simulation_result(sim.get_graph());
        });
    for (size_t r = 0; r < ensemble.size(); ++r) {
        for (size_t k = 0; k < ensemble[r].size(); ++k) {
            for (Eigen::Index d = 0; d < ensemble[r][k].get_num_time_points(); ++d) {
                std::printf("%zu %zu %g", r, k, ensemble[r][k].get_time(d));
                for (Eigen::Index e = 0; e < n; ++e) {
                    std::printf(" %.17g", ensemble[r][k][d][e]);
                }
                std::printf("\n");
            }
        }
    }
    auto p50 = mio::ensemble_percentile(ensemble, 0.5);
    for (size_t k = 0; k < p50.size(); ++k) {
        for (Eigen::Index d = 0; d < p50[k].get_num_time_points(); ++d) {
            std::printf("p50 %zu %g", k, p50[k].get_time(d));
            for (Eigen::Index e = 0; e < n; ++e) {
                std::printf(" %.17g", p50[k][d][e]);
            }
            std::printf("\n");
        }
    }
    return 0;
}

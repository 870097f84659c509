// secir_single -- one ODE-SECIR node, 1 age group, default adaptive integrator, 50 days: the scenario behind the
// reference's golden file pycode/memilio-simulation/tests/data/ode-secihurd-compare.csv (setup restated from
// pycode/memilio-simulation/tests/test_osecir.py:38-88). Prints "t y0 .. y9" for every accepted integrator step.
#include "memilio_b200/mio.hpp"

#include <cstdio>

int main()
{
    using namespace mio::osecir;
    const double t0 = 0, tmax = 50, dt = 0.1;
    Model<double> model(1);
    auto& p = model.parameters;
    p.set<StartDay<double>>(60);
    p.set<Seasonality<double>>(0.2);
    p.get<TimeExposed<double>>()                       = 3.2;
    p.get<TimeInfectedNoSymptoms<double>>()            = 2.0;
    p.get<TimeInfectedSymptoms<double>>()              = 5.8;
    p.get<TimeInfectedSevere<double>>()                = 9.5;
    p.get<TimeInfectedCritical<double>>()              = 7.1;
    p.get<TransmissionProbabilityOnContact<double>>()  = 0.05;
    p.get<RelativeTransmissionNoSymptoms<double>>()    = 0.7;
    p.get<RecoveredPerInfectedNoSymptoms<double>>()    = 0.09;
    p.get<RiskOfInfectionFromSymptomatic<double>>()    = 0.25;
    p.get<MaxRiskOfInfectionFromSymptomatic<double>>() = 0.45;
    p.get<TestAndTraceCapacity<double>>()              = 35;
    p.get<SeverePerInfectedSymptoms<double>>()         = 0.2;
    p.get<CriticalPerSevere<double>>()                 = 0.25;
    p.get<DeathsPerCritical<double>>()                 = 0.3;
    mio::ContactMatrixGroup<double>& cm = p.get<ContactPatterns<double>>();
    cm[0] = mio::ContactMatrix<double>(Eigen::MatrixX<double>::Constant(1, 1, 10.0));
    cm[0].add_damping(0.7, mio::SimulationTime<double>(30.));
    const mio::AgeGroup a0(0);
    model.populations[{a0, InfectionState::Exposed}]            = 100;
    model.populations[{a0, InfectionState::InfectedNoSymptoms}] = 50;
    model.populations[{a0, InfectionState::InfectedSymptoms}]   = 50;
    model.populations[{a0, InfectionState::InfectedSevere}]     = 20;
    model.populations[{a0, InfectionState::InfectedCritical}]   = 10;
    model.populations[{a0, InfectionState::Recovered}]          = 10;
    model.populations[{a0, InfectionState::Dead}]               = 0;
    model.populations.set_difference_from_total({a0, InfectionState::Susceptible}, 10000);
    model.apply_constraints();

    mio::TimeSeries<double> secir = simulate<double>(t0, tmax, dt, model);
    for (Eigen::Index i = 0; i < secir.get_num_time_points(); ++i) {
        std::printf("%.17g", secir.get_time(i));
        for (Eigen::Index k = 0; k < secir.get_num_elements(); ++k) {
            std::printf(" %.17g", secir[i][k]);
        }
        std::printf("\n");
    }
    return 0;
}

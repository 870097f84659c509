// mio_oracle_seir.hpp -- CPU ORACLE (test infrastructure, NOT product code).
//
// ODE-SEIR (cpp/models/ode_seir/model.h:39-107) restated only to pin the graph mechanics of the oracle
// (GraphSim, SimulationNode, MobilityEdge, integrator) by the reference's own golden file
// cpp/tests/data/graphsimulation-compare.csv (TestGraphSimulation.consistencyFlowMobility,
// cpp/tests/test_graph_simulation.cpp:256-339): the SEIR model has no ADL mobility hooks, so the edges use unit
// mobility factors and no commuter testing (metapopulation_mobility_instant.h:513-518,537-541).
#pragma once

#include "mio_oracle.hpp"

namespace orc
{
namespace oseir
{
enum : int
{
    S = 0,
    E,
    I,
    R,
    NC
};
constexpr int NF = 3;
static const int FLOW_SRC[NF] = {S, E, I};
static const int FLOW_DST[NF] = {E, I, R};

struct Model {
    int G;
    Vec TimeExposed, TimeInfected, TransmissionProbabilityOnContact;
    DampingExprGroup ContactPatterns;
    Vec populations;
    explicit Model(int g = 1)
        : G(g)
        , TimeExposed(g, 5.2)
        , TimeInfected(g, 6.0)
        , TransmissionProbabilityOnContact(g, 1.0)
        , populations(size_t(g) * NC, 0.0)
    {
    }
    int num_compartments() const
    {
        return NC;
    }
    // ode_seir/model.h:69-99
    void get_flows(const Vec& pop, const Vec& y, double t, Vec& flows) const
    {
        for (int i = 0; i < G; ++i) {
            const int Si = i * NC + S, Ei = i * NC + E, Ii = i * NC + I;
            for (int j = 0; j < G; ++j) {
                const int Sj = j * NC + S, Ej = j * NC + E, Ij = j * NC + I, Rj = j * NC + R;
                const double Nj        = pop[Sj] + pop[Ej] + pop[Ij] + pop[Rj];
                const double divNj     = (Nj < 1e-12) ? 0.0 : 1.0 / Nj;
                const double coeffStoE = ContactPatterns.at(t, i, j) * TransmissionProbabilityOnContact[i] * divNj;
                flows[i * NF + 0] += coeffStoE * y[Si] * pop[Ij];
            }
            flows[i * NF + 1] = (1.0 / TimeExposed[i]) * y[Ei];
            flows[i * NF + 2] = (1.0 / TimeInfected[i]) * y[Ii];
        }
    }
    mutable Vec m_flow_values;
    void get_derivatives(const Vec& pop, const Vec& y, double t, Vec& dydt) const
    {
        m_flow_values.assign(size_t(G) * NF, 0.0);
        get_flows(pop, y, t, m_flow_values);
        std::fill(dydt.begin(), dydt.end(), 0.0);
        for (int i = 0; i < G; ++i) {
            for (int k = 0; k < NF; ++k) {
                dydt[i * NC + FLOW_SRC[k]] -= m_flow_values[i * NF + k];
                dydt[i * NC + FLOW_DST[k]] += m_flow_values[i * NF + k];
            }
        }
    }
    void eval_right_hand_side(const Vec& pop, const Vec& y, double t, Vec& dydt) const
    {
        std::fill(dydt.begin(), dydt.end(), 0.0);
        get_derivatives(pop, y, t, dydt);
    }
    int num_flows() const
    {
        return G * NF;
    }
    void derivatives_from_flows(const Vec& flows, Vec& dydt) const
    {
        std::fill(dydt.begin(), dydt.end(), 0.0);
        for (int i = 0; i < G; ++i) {
            for (int k = 0; k < NF; ++k) {
                dydt[i * NC + FLOW_SRC[k]] -= flows[i * NF + k];
                dydt[i * NC + FLOW_DST[k]] += flows[i * NF + k];
            }
        }
    }
    // no ADL overloads for this model: instant.h:513-518, :537-541
    Vec get_mobility_factors(double, const Vec& y) const
    {
        return Vec(y.size(), 1.0);
    }
    void test_commuters(Vec&, Vec&, double) const
    {
    }
};
} // namespace oseir
} // namespace orc

/*
 * mio_cuda.h -- C-ABI of the B200 backend for MEmilio's Graph-ODE ensemble hot path.
 *
 * The reference (SciCompMod/memilio @ 8851d5be) has no FFI layer; it is a header-template C++ library. The seams
 * this ABI sits behind are (paths relative to the reference tree):
 *   - mio::make_mobility_sim / GraphSimulation::advance      cpp/memilio/mobility/metapopulation_mobility_instant.h:656-678,
 *                                                            cpp/memilio/mobility/graph_simulation.h:101-120
 *   - SimulationNode::advance + Simulation::advance          cpp/memilio/mobility/metapopulation_mobility_instant.h:91-95,
 *                                                            cpp/memilio/compartments/simulation.h:66-73,
 *                                                            cpp/memilio/math/integrator.h:184-249
 *   - MobilityEdge::apply_mobility                           cpp/memilio/mobility/metapopulation_mobility_instant.h:546-623
 *   - ParameterStudy::run (create_simulation / advance /     cpp/memilio/compartments/parameter_studies.h:158-217,290-345
 *     process_result per run)
 *   - ensemble_percentile / ensemble_mean / sum_nodes        cpp/memilio/data/analyze_result.h:267-343
 *
 * The C++ facade in include/memilio_b200/ keeps the reference's names and signatures (mio::Graph, SimulationNode,
 * MobilityEdge, make_mobility_sim, ParameterStudy) and packs into the flat descriptors below.
 *
 * Conventions
 *   - plain C, no exceptions, no torch types; every function returns a mio_status (numbering of mio::StatusCode,
 *     cpp/memilio/io/io.h:49-59) and mio_cuda_last_error_string() holds the message of the last failure.
 *   - the caller owns every host buffer for the duration of the call only; the library copies.
 *   - there is NO CPU fallback: if no sm_100 device is usable every compute entry point fails with MIO_UNKNOWN_ERROR.
 *   - a handle is not thread safe; one handle per GPU.
 *   - all floating point data is FP64. State vectors use the reference's flat layout idx = age*C + compartment
 *     (cpp/memilio/utils/custom_index_array.h:59-95).
 */
#ifndef MIO_CUDA_H
#define MIO_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define MIO_API
#else
#define MIO_API __attribute__((visibility("default")))
#endif

/* mio::StatusCode, cpp/memilio/io/io.h:49-59 */
typedef enum {
    MIO_OK            = 0,
    MIO_UNKNOWN_ERROR = 1,
    MIO_OUT_OF_RANGE  = 2,
    MIO_INVALID_VALUE = 3,
    MIO_INVALID_FILE_FORMAT = 4,
    MIO_KEY_NOT_FOUND = 5,
    MIO_INVALID_TYPE  = 6,
    MIO_FILE_NOT_FOUND = 7
} mio_status;

/* model kinds */
#define MIO_MODEL_OSECIR 0    /* cpp/models/ode_secir/model.h   : 10 compartments, 15 flows per age group */
#define MIO_MODEL_OSECIRVVS 1 /* cpp/models/ode_secirvvs/model.h: 27 compartments, 45 flows per age group */

/* integrator cores */
#define MIO_INTEGRATOR_CK54 0  /* ControlledStepperWrapper<boost cash_karp54>, cpp/memilio/math/stepper_wrapper.h:59-202 (default) */
#define MIO_INTEGRATOR_RKF45 1 /* mio::RKIntegratorCore, cpp/memilio/math/adapt_rk.h:145-310 */
#define MIO_INTEGRATOR_EULER 2 /* mio::EulerIntegratorCore, cpp/memilio/math/euler.h:33-63 */

/* node simulation kinds (only relevant for OSECIRVVS) */
#define MIO_SIM_BASE 0  /* mio::Simulation<FP, Model>: plain advance, unit mobility factors, no commuter testing */
#define MIO_SIM_MODEL 1 /* osecir::Simulation / osecirvvs::Simulation: model-specific advance + ADL mobility hooks */

#define MIO_MAX_GROUPS 8

/* ---- flat parameter-set layout --------------------------------------------------------------------------------
 * A parameter set is MIO_PSET_HEADER doubles followed by num_groups age records of mio_pset_age_stride(model)
 * doubles (age-major, so that the device lane owning age group i loads one contiguous record).
 */
#define MIO_PSET_HEADER 8
/* header slots (both models) */
#define MIO_PH_START_DAY 0            /* StartDay                       ode_secir/parameters.h:47  */
#define MIO_PH_SEASONALITY 1          /* Seasonality                    ode_secir/parameters.h:64  */
#define MIO_PH_TNT_MAX_RISK 2         /* TestAndTraceCapacityMaxRisk    ode_secir/parameters.h:374 */
#define MIO_PH_COMMUTER_NONDETECTION 3 /* m_commuter_nondetection        ode_secir/parameters.h:417 */
#define MIO_PH_START_COMMUTER_DETECTION 4
#define MIO_PH_END_COMMUTER_DETECTION 5
#define MIO_PH_RESERVED6 6
#define MIO_PH_RESERVED7 7

/* osecir age record (16 doubles, 14 used) -- tags of cpp/models/ode_secir/parameters.h:96-318 */
#define MIO_OSECIR_AGE_STRIDE 16
#define MIO_PA_TIME_EXPOSED 0
#define MIO_PA_TIME_INS 1       /* TimeInfectedNoSymptoms */
#define MIO_PA_TIME_ISY 2       /* TimeInfectedSymptoms */
#define MIO_PA_TIME_ISEV 3      /* TimeInfectedSevere */
#define MIO_PA_TIME_ICR 4       /* TimeInfectedCritical */
#define MIO_PA_TRANS_PROB 5     /* TransmissionProbabilityOnContact */
#define MIO_PA_REL_TRANS_NOSYMP 6 /* RelativeTransmissionNoSymptoms */
#define MIO_PA_REC_PER_INS 7    /* RecoveredPerInfectedNoSymptoms */
#define MIO_PA_RISK_SYMP 8      /* RiskOfInfectionFromSymptomatic */
#define MIO_PA_MAX_RISK_SYMP 9  /* MaxRiskOfInfectionFromSymptomatic */
#define MIO_PA_SEV_PER_ISY 10   /* SeverePerInfectedSymptoms */
#define MIO_PA_CRIT_PER_SEV 11  /* CriticalPerSevere */
#define MIO_PA_DEATHS_PER_SEV 12 /* DeathsPerSevere */
#define MIO_PA_DEATHS_PER_CRIT 13 /* DeathsPerCritical */

/* osecirvvs: header slot 2 holds TestAndTraceCapacityMaxRiskSymptoms, slots 6/7 the two extra scalars
 * (cpp/models/ode_secirvvs/parameters.h:62-155) */
#define MIO_PH_TNT_MAX_RISK_NOSYMP 6   /* TestAndTraceCapacityMaxRiskNoSymptoms (osecirvvs) */
#define MIO_PH_START_DAY_NEW_VARIANT 7 /* StartDayNewVariant (osecirvvs) */
/* osecirvvs age record (32 doubles): slots 0..13 as for osecir, then the tags of ode_secirvvs/parameters.h:422-625 */
#define MIO_OSECIRVVS_AGE_STRIDE 32
#define MIO_PV_REDUC_EXPOSED_PI 14      /* ReducExposedPartialImmunity */
#define MIO_PV_REDUC_EXPOSED_II 15      /* ReducExposedImprovedImmunity */
#define MIO_PV_REDUC_INF_SYMP_PI 16     /* ReducInfectedSymptomsPartialImmunity */
#define MIO_PV_REDUC_INF_SYMP_II 17     /* ReducInfectedSymptomsImprovedImmunity */
#define MIO_PV_REDUC_SEV_CRIT_DEAD_PI 18 /* ReducInfectedSevereCriticalDeadPartialImmunity */
#define MIO_PV_REDUC_SEV_CRIT_DEAD_II 19 /* ReducInfectedSevereCriticalDeadImprovedImmunity */
#define MIO_PV_REDUC_TIME_MILD 20       /* ReducTimeInfectedMild */
#define MIO_PV_INFECTIOUSNESS_NEW_VARIANT 21 /* InfectiousnessNewVariant */
#define MIO_PV_VACCINATION_GAP 22
#define MIO_PV_DAYS_UNTIL_EFFECTIVE_PI 23
#define MIO_PV_DAYS_UNTIL_EFFECTIVE_II 24

/* per-node local parameters: [ICUCapacity, TestAndTraceCapacity] (sampled per node, parameter_space.h:125-148) */
#define MIO_NODE_LOCAL 2
#define MIO_NL_ICU_CAPACITY 0
#define MIO_NL_TNT_CAPACITY 1

/* ---- flat contact-table layout ----------------------------------------------------------------------------------
 * One contact table is the evaluated damping cache of a ContactMatrixGroup (cpp/memilio/epidemiology/
 * contact_matrix.h:271-400, damping.h:571-600): for each of L contact locations the baseline B, the minimum M and the
 * K accumulated dampings D_k active from change time t_k on (sentinels lowest()/max() are implied).
 *   doubles[0]                    = L                       (1 <= L <= ct_max_loc)
 *   doubles[1 .. ct_max_loc]      = K_l                     (0 <= K_l <= ct_max_pts)
 *   then per location l, stride ct_max_pts + (2 + ct_max_pts)*G*G:
 *        times[ct_max_pts], B[G*G], M[G*G], D[ct_max_pts][G*G]      (matrices row-major, (i,j) -> i*G + j)
 * mio_ctable_size() returns the total number of doubles.
 * The mobility table of a run (dampings applied to the mobility coefficients of all edges,
 * parameter_space.h:241-248) has the same shape with L = 1 and vectors of length n = G*C in place of matrices, B
 * and M omitted (baseline is the per-edge coefficient vector, minimum is zero):
 *   doubles[0] = K, times[mob_max_pts], D[mob_max_pts][n]
 */

typedef struct {
    int32_t kind;     /* MIO_INTEGRATOR_* */
    int32_t reserved;
    double abs_tol;   /* default 1e-10 */
    double rel_tol;   /* default 1e-5  */
    double dt_min;    /* default DBL_MIN */
    double dt_max;    /* default DBL_MAX */
} mio_integrator_desc;

typedef struct {
    int32_t model;       /* MIO_MODEL_*  */
    int32_t sim_kind;    /* MIO_SIM_*    */
    int32_t num_groups;  /* G, 1..MIO_MAX_GROUPS */
    int32_t num_nodes;   /* N */
    int32_t num_edges;   /* E (may be 0) */
    int32_t num_psets;   /* parameter sets per run  (>=1) */
    int32_t num_ctables; /* contact tables per run  (>=1) */
    int32_t ct_max_loc;  /* capacity: contact locations per table */
    int32_t ct_max_pts;  /* capacity: damping change points per location */
    int32_t mob_max_pts; /* capacity: change points of the per-run mobility damping table; 0 = none */
    int32_t max_runs;    /* number of run slots held on the device */
    int32_t record_daily; /* 1: keep the state of every node at every integer day (after the edge phase) */
    int32_t max_days;    /* capacity of the daily record (days since t0, inclusive of day 0) */
    int32_t traj_capacity; /* >0: record every accepted integrator step of every node, up to this many points */
    const int32_t* edge_start; /* [E] source node, reference edge order = sorted by (start,end), mobility/graph.h:233-241 */
    const int32_t* edge_end;   /* [E] destination node */
    const double* edge_coeff;  /* [E][n] baseline mobility coefficients (MobilityCoefficientGroup summed over its members) */
    const int32_t* node_pset;  /* [N] parameter set used by each node   (NULL = all 0) */
    const int32_t* node_ctable; /* [N] contact table used by each node  (NULL = all 0) */
    mio_integrator_desc integrator;
    double t0;
    double dt_graph;     /* GraphSimulation m_dt; +inf for make_no_mobility_sim */
    /* osecirvvs only: cumulative vaccinations DailyPartialVaccinations / DailyFullVaccinations
     * (ode_secirvvs/parameters.h:470-497), run-invariant, [N][G][vacc_days]; NULL / 0 for osecir */
    int32_t vacc_days;
    int32_t flow_simulation; /* 1: the nodes are mio::FlowSimulation (compartments/flow_simulation.h:37-97): the cumulative
                              * flows are integrated (error control on the flows), the populations derived from them;
                              * osecir only. 0: mio::Simulation */
    const double* vacc_partial;
    const double* vacc_full;
} mio_graph_desc;

typedef struct {
    const double* psets;      /* [count][num_psets][mio_pset_size]       */
    const double* ctables;    /* [count][num_ctables][mio_ctable_size]   */
    const double* node_local; /* [count][N][MIO_NODE_LOCAL]              */
    const double* y0;         /* [count][N][n]                            */
    const double* dt0;        /* [count][N] initial integrator dt per node, NULL = 0.1 (Simulation ctor default) */
    const double* mobtables;  /* [count][mio_mobtable_size], NULL = no mobility dampings */
} mio_run_data;

/* per (run,node) counters returned by mio_cuda_get_counters */
#define MIO_NUM_COUNTERS 4
#define MIO_CNT_ACCEPTED 0      /* accepted integrator steps (= time points appended by SystemIntegrator::advance) */
#define MIO_CNT_REJECTED 1      /* failed try_step attempts */
#define MIO_CNT_FORCED 2        /* steps forced at dt_min (integrator.h:236-239 "Adaptive step sizing failed") */
#define MIO_CNT_UNDERFLOW 3     /* return-phase underflow fix-ups applied with this node as destination (instant.h:579-596) */

typedef struct mio_handle mio_handle;

MIO_API const char* mio_cuda_last_error_string(void);
MIO_API int mio_cuda_version(void);
MIO_API int mio_cuda_device_count(void);

MIO_API int mio_num_compartments(int model);
MIO_API int mio_pset_age_stride(int model);
MIO_API int64_t mio_pset_size(int model, int num_groups);
MIO_API int64_t mio_ctable_size(int num_groups, int ct_max_loc, int ct_max_pts);
MIO_API int64_t mio_mobtable_size(int model, int num_groups, int mob_max_pts);

/* make_mobility_sim: uploads the graph, allocates max_runs run slots on the current CUDA device. */
MIO_API int mio_cuda_graph_create(const mio_graph_desc* desc, int device, mio_handle** out);
MIO_API void mio_cuda_graph_destroy(mio_handle* h);

/* (re)initialise run slots [slot0, slot0+count): parameters, initial state, t = t0 */
MIO_API int mio_cuda_set_runs(mio_handle* h, int32_t slot0, int32_t count, const mio_run_data* data);
/* rewind the simulation clock and the edges' departure/return toggle to t0; follow with set_runs for every slot */
MIO_API int mio_cuda_reset(mio_handle* h);
/* number of slots that take part in advance() (default max_runs) */
MIO_API int mio_cuda_set_active_runs(mio_handle* h, int32_t count);

/* GraphSimulation::advance(t_max) for all active run slots */
MIO_API int mio_cuda_advance(mio_handle* h, double t_max);
MIO_API int mio_cuda_get_time(mio_handle* h, double* t);

/* node state get_result().get_last_value(): out[count][N][n] */
MIO_API int mio_cuda_get_state(mio_handle* h, int32_t slot0, int32_t count, double* out);
/* SimulationBase::get_dt() per node: out[count][N] */
MIO_API int mio_cuda_get_node_dt(mio_handle* h, int32_t slot0, int32_t count, double* out);
/* out[count][N][MIO_NUM_COUNTERS] */
MIO_API int mio_cuda_get_counters(mio_handle* h, int32_t slot0, int32_t count, int64_t* out);
/* daily record (record_daily=1): out[count][N][num_days][n], num_days = days recorded so far (returned) */
MIO_API int mio_cuda_get_daily(mio_handle* h, int32_t slot0, int32_t count, int32_t* num_days, double* out,
                               int64_t out_capacity);
/* recorded trajectories (traj_capacity>0): times[count][N][cap], values[count][N][cap][n], npts[count][N] */
MIO_API int mio_cuda_get_trajectory(mio_handle* h, int32_t slot0, int32_t count, int32_t* npts, double* times,
                                    double* values);

/* mio::FlowSimulation::get_flows() (cpp/memilio/compartments/flow_simulation.h:104-119) of graphs created with
 * flow_simulation = 1: get_flows().get_last_value() of every node, out[count][N][G*15] (age-major, the 15 flows in the
 * order of osecir's Flows type list, cpp/models/ode_secir/model.h:45-59) */
MIO_API int mio_cuda_get_flows(mio_handle* h, int32_t slot0, int32_t count, double* out);
/* the flows a FlowSimulation continues from (mio_cuda_set_runs starts them at zero, FlowModel::get_initial_flows,
 * cpp/memilio/compartments/flow_model.h:146-149): flows[count][N][G*15], to be called after mio_cuda_set_runs */
MIO_API int mio_cuda_set_flows(mio_handle* h, int32_t slot0, int32_t count, const double* flows);
/* Dynamic NPIs (cpp/models/ode_secir/model.h:299-353, cpp/memilio/epidemiology/dynamic_npis.h:335-395) change the
 * contact patterns of a node while the simulation runs: the host evaluates the thresholds between graph steps (the
 * C++ facade does, osecir::Simulation / GraphSimulation) and replaces the contact tables of the run,
 * ctables[count][num_ctables][mio_ctable_size] as in mio_run_data. The simulation continues where it is. */
MIO_API int mio_cuda_set_contact_tables(mio_handle* h, int32_t slot0, int32_t count, const double* ctables);
/* ... and the whole flow series (needs traj_capacity > 0): values[count][N][traj_capacity][G*15]; the time points
 * are those of mio_cuda_get_trajectory, whose values are the populations that follow from the flows
 * (compute_population_results, cpp/memilio/compartments/flow_simulation_base.h:84-97) */
MIO_API int mio_cuda_get_flow_trajectory(mio_handle* h, int32_t slot0, int32_t count, double* values);

/* statistics: kernels launched by this handle so far; (run, step) pairs whose return phase needed the literal
 * in-order replay because of underflow fix-ups */
/* MobilityEdge::get_mobility_results (cpp/memilio/mobility/metapopulation_mobility_instant.h:421-445, called at :599
 * and :620): every edge is given the same saved_compartment_indices -- num_groups index groups, group_ptr[num_groups+1]
 * into group_idx (reference element indices age*C + compartment). From then on every graph step appends one time point
 * per edge that carries commuters: the sums of the mobile population over each index group and, last, its total. The
 * per-edge traveler state this needs is materialised only for handles that ask for it (R*E*n doubles), and their return
 * steps are replayed with the literal sequential edge loop. capacity = time points kept per edge; a second call with a
 * larger capacity keeps what was recorded. */
MIO_API int mio_cuda_enable_mobility_results(mio_handle* h, int32_t num_groups, const int32_t* group_ptr,
                                             const int32_t* group_idx, int32_t capacity);
/* npts[count][E] points recorded per edge, times[capacity] (the graph times of the steps, common to all edges),
 * values[count][E][capacity][num_groups + 1] */
MIO_API int mio_cuda_get_mobility_results(mio_handle* h, int32_t slot0, int32_t count, int32_t* npts, double* times,
                                          double* values);
MIO_API int mio_cuda_get_stats(mio_handle* h, int64_t* launches, int64_t* exact_replays);

/* ---- ParameterStudy sampling (host C++, no device needed) --------------------------------------------------------
 * Replaces, per run: ParameterStudy::run_impl's counter logic (cpp/memilio/compartments/parameter_studies.h:304-339),
 * osecir::draw_sample(Graph&) (cpp/models/ode_secir/parameter_space.h:207-251) with draw_sample_infection :155-191 and
 * draw_sample_demographics :125-148, UncertainValue::draw_sample (cpp/memilio/utils/uncertain_value.h:181-188),
 * ParameterDistribution*::sample (cpp/memilio/utils/parameter_distributions.h), UncertainContactMatrix::
 * draw_sample_dampings / make_matrix (cpp/memilio/epidemiology/uncertain_matrix.h:165-200), apply_dampings /
 * make_contact_damping_matrix / make_mobility_damping_vector (epidemiology/damping_sampling.h:256-307),
 * Dampings::update_cache (epidemiology/damping.h:571-660) and Parameters::apply_constraints
 * (ode_secir/parameters.h:465-615). Random numbers: threefry2x32-20 keyed by std::seed_seq(seeds)
 * (utils/random_number_generator.h:195-218,320-331) under the libstdc++ distributions, exactly as the reference.
 */
#define MIO_DIST_NONE 0        /* UncertainValue without distribution: draw_sample() keeps the value, no RNG draw */
#define MIO_DIST_UNIFORM 1     /* ParameterDistributionUniform(a = lower, b = upper) */
#define MIO_DIST_NORMAL 2      /* ParameterDistributionNormal(lower, upper, mean = a, stddev = b, quantile) */
#define MIO_DIST_LOGNORMAL 3   /* ParameterDistributionLogNormal(log_mean = a, log_stddev = b) */
#define MIO_DIST_EXPONENTIAL 4 /* ParameterDistributionExponential(rate = a) */
#define MIO_DIST_CONSTANT 5    /* ParameterDistributionConstant(a) */

typedef struct {
    int32_t kind; /* MIO_DIST_* */
    int32_t reserved;
    double value; /* UncertainValue::m_value before sampling */
    double a, b;
    double lower, upper, quantile; /* normal only */
} mio_uncertain_value;

/* DampingSampling, cpp/memilio/epidemiology/damping_sampling.h:40-120 */
typedef struct {
    mio_uncertain_value value;
    int32_t level, type;
    double time;
    uint64_t matrix_mask; /* bit l set: the damping applies to contact location l (get_matrix_indices) */
    double group_weights[MIO_MAX_GROUPS];
} mio_damping_sampling;

typedef struct {
    int32_t model;         /* MIO_MODEL_OSECIR or MIO_MODEL_OSECIRVVS (ode_secirvvs/parameter_space.h:40-211) */
    int32_t num_groups;    /* G */
    int32_t num_nodes;     /* N */
    int32_t num_locations; /* L contact matrices of ContactPatterns */
    int32_t num_dampings;  /* entries of UncertainContactMatrix::get_dampings() */
    int32_t num_holiday_sets; /* distinct school holiday period lists (0 = no holidays anywhere) */
    int32_t ct_max_pts;    /* capacity of the produced contact tables (change points per location) */
    int32_t mob_max_pts;   /* capacity of the produced mobility tables; 0 = do not produce mobility tables */
    int32_t mobility_location; /* member of the edges' MobilityCoefficientGroup that carries the baseline (set_edges:
                                  ContactLocation::Work, graph.h:420-440; 0 for single-member groups) */
    int32_t num_seeds;     /* <= 8; RandomNumberGenerator::seed(seeds) */
    uint32_t seeds[8];
    uint64_t rng_counter;  /* m_rng.get_counter() before run() (0 for a fresh study) */
    const mio_uncertain_value* header;      /* [MIO_PSET_HEADER]  of node 0 (the shared parameter model) */
    const mio_uncertain_value* age_params;  /* [G][mio_pset_age_stride] of node 0 */
    const mio_uncertain_value* node_local;  /* [N][MIO_NODE_LOCAL] */
    const mio_uncertain_value* populations; /* [N][n] */
    const double* baseline;                 /* [L][G*G] */
    const double* minimum;                  /* [L][G*G] */
    const mio_damping_sampling* dampings;   /* [num_dampings] */
    const mio_damping_sampling* school_holiday_damping; /* NULL = the default (value 0, no distribution, level/type 0) */
    const int32_t* node_holiday_set;        /* [N] index into the holiday sets, -1 = none; NULL = none */
    const int32_t* holiday_ptr;             /* [num_holiday_sets + 1] */
    const double* holiday_periods;          /* [holiday_ptr[last]][2] = (start, end) */
} mio_study_desc;

typedef struct mio_study mio_study;

MIO_API int mio_study_create(const mio_study_desc* desc, mio_study** out);
MIO_API void mio_study_destroy(mio_study* s);
/* number of contact tables per run the study produces (= max(1, num_holiday_sets + [some node has none])) and the
 * node -> contact table map to pass as mio_graph_desc.node_ctable */
MIO_API int mio_study_num_ctables(const mio_study* s);
MIO_API int mio_study_node_ctable(const mio_study* s, int32_t* node_ctable /*[N]*/);
/* draw_sample(graph) for runs [run0, run0 + count) of the study, each with the RNG counter
 * rng_counter + (run_idx << 32). Outputs use the layouts of mio_run_data (ct_max_loc = num_locations);
 * mobtables may be NULL when mob_max_pts == 0; draws (optional) receives the number of random numbers each run used. */
MIO_API int mio_study_sample_runs(mio_study* s, int64_t run0, int32_t count, double* psets, double* ctables,
                                  double* node_local, double* y0, double* mobtables, uint64_t* draws);
/* ParameterStudy::distribute_runs (parameter_studies.h:354-367): first run and run count of `rank` */
MIO_API int mio_study_distribute_runs(int64_t num_runs, int32_t num_procs, int32_t rank, int64_t* run0, int64_t* count);
/* raw generator output: RandomNumberGenerator::operator() for `count` consecutive counters (known-answer tests) */
MIO_API int mio_study_rng_raw(const uint32_t* seeds, int32_t num_seeds, uint64_t counter, int32_t count, uint64_t* out,
                              uint64_t* key_out);
/* Dampings::update_cache for K raw dampings of shape rows x cols (matrices row-major); returns the accumulated
 * dampings and their change times in the layout the contact / mobility tables use. *num_out = cache entries. */
MIO_API int mio_damping_cache_build(int32_t rows, int32_t cols, int32_t K, const double* times, const int32_t* levels,
                                    const int32_t* types, const double* mats, int32_t* num_out, double* out_times,
                                    double* out_mats);

/* ---- ensemble results -------------------------------------------------------------------------------------------
 * Device-resident record of the interpolated (daily) results of all runs of a study and their reduction.
 * Replaces interpolate_simulation_result / sum_nodes / ensemble_mean / ensemble_percentile
 * (cpp/memilio/data/analyze_result.h:186-343) as used by save_results (cpp/memilio/io/result_io.h:135-202).
 * The record is a column-major matrix [column][run]: per-node columns (node k, day d, element e) ->
 * (k * num_days + d) * El + e with El = n (MIO_ENS_FULL) or C (MIO_ENS_AGE_SUM: age groups summed per compartment),
 * followed by the node-summed full-resolution columns N*num_days*El + d * n + j (sum_nodes).
 */
#define MIO_ENS_FULL 0
#define MIO_ENS_AGE_SUM 1
typedef struct mio_ensemble mio_ensemble;
MIO_API int mio_cuda_ensemble_create(int device, int32_t model, int32_t num_groups, int32_t num_nodes,
                                     int32_t num_days, int32_t mode, int64_t capacity_runs, mio_ensemble** out);
MIO_API void mio_cuda_ensemble_destroy(mio_ensemble* e);
/* from now on set_runs() records day 0 and advance() every further integer day of the active run slots into the
 * runs [run_offset, run_offset + active) of the record (t0 must be an integer day); e = NULL detaches */
MIO_API int mio_cuda_attach_ensemble(mio_handle* h, mio_ensemble* e, int64_t run_offset);
MIO_API int mio_cuda_ensemble_info(mio_ensemble* e, int64_t* num_columns, int64_t* stride, int64_t* num_runs,
                                   void** device_ptr);
/* ---- multi-GPU: ParameterStudy::run's MPI part (cpp/memilio/compartments/parameter_studies.h:163-213) ---------------
 * One process per GPU; rank r simulates the block distribute_runs gives it (:354-367) into its own ensemble record.
 * mio_cuda_ensemble_reduce_sharded re-shards the records from runs to columns with one grouped NCCL send/receive
 * exchange over NVLink, selects the exact order statistics and the mean of the rank's column block over the runs of
 * all ranks, and gathers the rows on rank 0 like the reference gathers on its root (:190-212).
 * The communicator: rank 0 calls mio_cuda_comm_get_unique_id and hands the 128 bytes to the other ranks by whatever
 * means the application has (MPI_Bcast in the reference's setting); every rank then calls mio_cuda_comm_create.
 * NCCL is bound at run time (libnccl.so.2). At most 8 ranks. */
#define MIO_COMM_ID_BYTES 128
typedef struct mio_comm mio_comm;
MIO_API int mio_cuda_comm_get_unique_id(char* id /*[MIO_COMM_ID_BYTES]*/);
MIO_API int mio_cuda_comm_create(int device, int rank, int world, const char* id, mio_comm** out);
MIO_API void mio_cuda_comm_destroy(mio_comm* c);
/* e: this rank's record holding num_runs_local runs (NULL and 0 for a rank without runs; num_columns is then taken
 * from the argument). Outputs on rank 0 only: percentiles[num_p][num_columns], mean[num_columns] (either may be NULL);
 * other ranks may pass NULL. Collective: every rank of the communicator must call it. */
MIO_API int mio_cuda_ensemble_reduce_sharded(mio_ensemble* e, mio_comm* c, int64_t num_runs_local, int64_t num_columns,
                                             int32_t num_p, const double* p, double* percentiles, double* mean);
/* Day window of the record: the days [first_day, first_day + num_days) since t0 are recorded into the slots
 * [0, num_days), other days are skipped (default first_day = 0). A study whose full record does not fit device memory
 * (ODE-SECIRVVS, 12 500 runs x 1.1 M columns = 110 GB) keeps all its runs resident, advances them one window at a
 * time and reduces window by window: the record is sized for one window (interpolate_simulation_result +
 * ensemble_percentile of cpp/memilio/data/analyze_result.h:186-343 are per day, so the result is the same). */
MIO_API int mio_cuda_ensemble_set_window(mio_ensemble* e, int32_t first_day);
/* ensemble_percentile (order statistic size_t(num_runs * p) per column, exact) and ensemble_mean over the recorded
 * runs. Host outputs percentiles[num_p][num_columns], mean[num_columns]; either may be NULL. num_p <= 8. */
MIO_API int mio_cuda_ensemble_reduce(mio_ensemble* e, int32_t num_p, const double* p, double* percentiles,
                                     double* mean);
/* the same reduction for columns whose runs are spread over up to 8 device segments (multi-GPU: after the
 * run-sharded -> column-sharded exchange); segment s holds [num_columns][seg_stride[s]] doubles, seg_count[s] runs */
MIO_API int mio_cuda_select_percentiles(int device, int32_t num_segments, const void* const* seg_ptr,
                                        const int64_t* seg_stride, const int64_t* seg_count, int64_t num_columns,
                                        int32_t num_p, const double* p, double* percentiles, double* mean);

/* convenience for host callers (the C++ facade's ensemble_percentile / ensemble_mean): host matrix
 * columns[num_columns][num_runs] -> percentiles[num_p][num_columns] and/or mean[num_columns], computed on the device */
MIO_API int mio_cuda_reduce_host_columns(const double* columns, int64_t num_columns, int64_t num_runs, int32_t num_p,
                                         const double* p, double* percentiles, double* mean);

/* ---- measurement hooks ------------------------------------------------------------------------------------------ */
#define MIO_NUM_KERNEL_CLASSES 6
#define MIO_KC_INTEGRATE 0    /* node phase: adaptive RK integration of every (run, node) */
#define MIO_KC_DEPART 1       /* edge phase: commuter departure gather */
#define MIO_KC_RETURN 2       /* edge phase: commuter return gather (optimistic) */
#define MIO_KC_RETURN_EXACT 3 /* edge phase: in-order replay of runs with underflow fix-ups */
#define MIO_KC_RECORD 4       /* daily record copy */
#define MIO_KC_ENSEMBLE 5     /* ensemble record (transpose + node sum) */
/* the CUDA stream (cudaStream_t) all kernels of this handle are launched on */
MIO_API int mio_cuda_get_stream(mio_handle* h, void** stream);
/* per-kernel-class device time, measured with CUDA events on that stream while profiling is on; resets the sums */
MIO_API int mio_cuda_set_profiling(mio_handle* h, int on);
MIO_API int mio_cuda_get_profile(mio_handle* h, double* ms /*[MIO_NUM_KERNEL_CLASSES]*/,
                                 int64_t* launches /*[MIO_NUM_KERNEL_CLASSES]*/);
/* device pointer of the live node state [max_runs][N][n] (valid until the next advance) */
MIO_API int mio_cuda_device_state_ptr(mio_handle* h, void** y, int64_t* num_doubles);
/* sustained FP64 FMA throughput of the device (DFMA chains on all SMs), the roofline denominator of the node phase */
MIO_API int mio_cuda_measure_fp64_peak(int device, double seconds, double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* MIO_CUDA_H */

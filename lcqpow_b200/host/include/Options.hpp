// Options.hpp -- LCQPow::Options for the B200 build.
//
// Same public surface as /root/reference/include/Options.hpp:35-185 (defaults src/Options.cpp:296-333,
// validations :85-259).  The reference embeds a qpOASES::Options and an OSQPSettings*; this build contains
// neither solver, so the two pass-through types below carry the fields that map onto the device QP solver
//  and keep user code of the form
//     options.getqpOASESOptions().terminationTolerance = ...;   options.setOSQPOptions(&settings);
// compiling.  toCuda() flattens everything into the POD that crosses the C ABI (include/lcqp_cuda.h).
#ifndef LCQPOW_B200_OPTIONS_HPP
#define LCQPOW_B200_OPTIONS_HPP

#include "Utilities.hpp"
#include "../../../include/lcqp_cuda.h"

#ifndef QPOASES_OPTIONS_HPP
namespace qpOASES {
enum PrintLevel { PL_DEBUG_ITER = -2, PL_TABULAR = -1, PL_NONE = 0, PL_LOW = 1, PL_MEDIUM = 2, PL_HIGH = 3 };
// The subset of qpOASES::Options (external/qpOASES/include/qpOASES/Options.hpp) that has a meaning for the
// device active-set solver.
struct Options {
    PrintLevel printLevel;
    double terminationTolerance;   // qpOASES: 5e6 * EPS      -> qp_dual_tol (multiplier sign test)
    double boundTolerance;         // qpOASES: 1e6 * EPS      -> qp_feas_tol (bound violation test)
    double epsRegularisation;      // qpOASES: 1e3 * EPS      -> qp_delta when > 0 and enableRegularisation
    int numRefinementSteps;        // qpOASES: 1              -> qp_refine_iter (at least 2 are always done)
    bool enableRegularisation;
    Options() { setToDefault(); }
    void setToDefault()
    {
        printLevel = PL_MEDIUM;
        terminationTolerance = 0.0;   // 0: keep the device solver's own default
        boundTolerance = 0.0;
        epsRegularisation = 0.0;
        numRefinementSteps = 0;
        enableRegularisation = false;
    }
};
}  // namespace qpOASES
#endif

#ifndef OSQP_H
// The subset of OSQPSettings (external/osqp/include/types.h:158-195) that maps onto the device solver's
// ADMM phase; 0 means "keep the device default".
struct OSQPSettings {
    double rho;        // ADMM step (0.1)
    double sigma;      // 1e-6
    double alpha;      // relaxation (1.6)
    double delta;      // polish regularisation (1e-6)
    int max_iter;      // 4000
    int check_termination;   // iterations between termination checks (25); exact-vertex solver: between active-set probes (10)
    int polish_refine_iter;  // refinement passes (3)
    int polish;        // 1 (LCQPow switches it on, Options.cpp:331)
    int verbose;
    double eps_prim_inf;
    // read by the OSQP restatement only (Options::setOSQPADMM)
    double eps_abs, eps_rel, eps_dual_inf, adaptive_rho_tolerance;
    int scaling, adaptive_rho, adaptive_rho_interval;
};
// 0 / 0.0 means "keep the device default" (the defaults of constants.h:59-114 live in lcqp_cuda_default_options)
inline void osqp_set_default_settings(OSQPSettings* s)
{
    s->rho = 0.0; s->sigma = 0.0; s->alpha = 0.0; s->delta = 0.0; s->max_iter = 0; s->check_termination = 0;
    s->polish_refine_iter = 0; s->polish = 1; s->verbose = 0; s->eps_prim_inf = 0.0;
    s->eps_abs = 0.0; s->eps_rel = 0.0; s->eps_dual_inf = 0.0; s->adaptive_rho_tolerance = 0.0;
    s->scaling = -1; s->adaptive_rho = -1; s->adaptive_rho_interval = -1;
}
#endif

namespace LCQPow {

class Options {
public:
    Options();
    Options(const Options& rhs);
    ~Options();
    Options& operator=(const Options& rhs);

    void setToDefault();

    double getStationarityTolerance() const;
    ReturnValue setStationarityTolerance(double val);
    double getComplementarityTolerance() const;
    ReturnValue setComplementarityTolerance(double val);
    double getInitialPenaltyParameter() const;
    ReturnValue setInitialPenaltyParameter(double val);
    double getPenaltyUpdateFactor() const;
    ReturnValue setPenaltyUpdateFactor(double val);
    bool getSolveZeroPenaltyFirst() const;
    ReturnValue setSolveZeroPenaltyFirst(bool val);
    bool getPerturbStep() const;
    ReturnValue setPerturbStep(bool val);
    int getMaxIterations() const;
    ReturnValue setMaxIterations(int val);
    double getMaxPenaltyParameter() const;
    ReturnValue setMaxPenaltyParameter(double val);
    int getNDynamicPenalty() const;
    ReturnValue setNDynamicPenalty(int val);
    double getEtaDynamicPenalty() const;
    ReturnValue setEtaDynamicPenalty(double val);
    PrintLevel getPrintLevel() const;
    ReturnValue setPrintLevel(PrintLevel val);
    ReturnValue setPrintLevel(int val);
    bool getStoreSteps() const;
    ReturnValue setStoreSteps(bool val);
    QPSolver getQPSolver() const;
    ReturnValue setQPSolver(QPSolver val);
    ReturnValue setQPSolver(int val);

    ReturnValue setqpOASESOptions(const qpOASES::Options& _options);
    qpOASES::Options& getqpOASESOptions();
    ReturnValue setOSQPOptions(OSQPSettings* _options);   // deep copy (reference: Options.cpp:275-287)
    OSQPSettings* getOSQPOptions();

    // --- additions of the B200 build -------------------------------------------------------------------
    // perturbStep draws come from a counter-based generator keyed by (seed, instance, iterate, coordinate)
    // instead of srand(time)/rand() (reference: LCQProblem.cpp:1016,1353-1362): runs are reproducible.
    unsigned long long getPerturbSeed() const;
    ReturnValue setPerturbSeed(unsigned long long seed);
    // QPSolver::OSQP_SPARSE runs the exact-vertex solver behind the OSQP dual layout unless this is set: then the
    // device runs the restatement of OSQP itself (ADMM + polish, lcqp_osqp.cuh), as the reference does
    // (SubsolverOSQP.cpp:124-200).  Device loop only (no step tracking / iteration table).
    bool getOSQPADMM() const;
    ReturnValue setOSQPADMM(bool on);
    // CUDA device ordinal used by LCQProblem / SubsolverCUDA
    int getDevice() const;
    ReturnValue setDevice(int dev);
    // the POD that crosses the C ABI
    void toCuda(lcqp_cuda_options& out) const;

protected:
    void copy(const Options& rhs);

    double stationarityTolerance;
    double complementarityTolerance;
    double initialPenaltyParameter;
    double penaltyUpdateFactor;
    bool solveZeroPenaltyFirst;
    bool perturbStep;
    int maxIterations;
    double maxPenaltyParameter;
    int nDynamicPenalty;
    double etaDynamicPenalty;
    bool storeSteps;
    QPSolver qpSolver;
    PrintLevel printLevel;
    qpOASES::Options qpOASES_opts;
    OSQPSettings* OSQP_opts = nullptr;
    unsigned long long perturbSeed;
    int device;
    bool osqpADMM = false;
};

}  // namespace LCQPow

#endif

/*
 * oracle/ref_shim.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * A thin extern "C" door onto the UNMODIFIED reference library (LCQPow + qpOASES 3.2 +
 * OSQP 0.6.2), compiled from the sources where they lie under /root/reference by
 * oracle/Makefile into oracle/_ref/liblcqpow_ref.so.  It drives the reference through its
 * own public API only (LCQProblem::loadLCQP / switchToSparseMode / runSolver /
 * getPrimalSolution / getDualSolution / getOutputStatistics, Options setters), exactly
 * like /root/reference/examples/warm_up.cpp:45-85 does.
 *
 * Used by: tests/ (golden-vector generation + live parity when the .so is present) and
 * bench.py's cpu_baseline / --impl reference arm.
 */
#include "LCQProblem.hpp"

#include <chrono>
#include <cstring>
#include <vector>

extern "C" {
#include <osqp.h>
}

extern "C" {

/* Plain-C mirror of the options the reference exposes (Options.hpp:192-213). */
typedef struct {
    double stationarityTolerance;
    double complementarityTolerance;
    double initialPenaltyParameter;
    double penaltyUpdateFactor;
    double maxPenaltyParameter;
    double etaDynamicPenalty;
    int solveZeroPenaltyFirst;
    int perturbStep;
    int maxIterations;
    int nDynamicPenalty;
    int qpSolver;                    /* 0 QPOASES_DENSE, 1 QPOASES_SPARSE, 2 OSQP_SPARSE */
    int osqp_adaptive_rho_interval;  /* <0: leave the reference default (0 = wall-clock driven) */
} lcqp_ref_options;

typedef struct {
    int ret;          /* LCQPow::ReturnValue of runSolver (or of loadLCQP if that failed) */
    int status;       /* LCQPow::AlgorithmStatus */
    int iterTotal;
    int iterOuter;
    int subproblemIter;
    int qpExitFlag;
    int nDuals;
    int pad;
    double rhoOpt;
    double seconds;   /* wall time of loadLCQP + (switchToSparseMode) + runSolver */
} lcqp_ref_result;

/* Debug door for the parity work (tests/ only): qpOASES print level of the next solves (its own
 * per-iteration table, PL_TABULAR = -1, goes to stdout) and whether the outer loop records its iterates
 * (Options::setStoreSteps).  The recorded run is read back with lcqpow_ref_last_trace. */
static int g_qpoases_print_level = 0; /* qpOASES::PL_NONE */
static int g_store_steps = 0;
static std::vector<double> g_trace_x, g_trace_alpha, g_trace_stat;
static std::vector<int> g_trace_sub;
static int g_trace_nV = 0;

/* Subsolver options pass-through (Options::setqpOASESOptions / setOSQPOptions): values <= 0 keep the reference's
 * defaults.  Used by the tests of the pass-through. */
static double g_qpoases_termtol = 0.0, g_qpoases_boundtol = 0.0, g_osqp_eps = 0.0;
static int g_osqp_max_iter = 0;

void lcqpow_ref_set_subsolver_options(double qpoases_terminationTolerance, double qpoases_boundTolerance, double osqp_eps_abs_rel, int osqp_max_iter)
{
    g_qpoases_termtol = qpoases_terminationTolerance;
    g_qpoases_boundtol = qpoases_boundTolerance;
    g_osqp_eps = osqp_eps_abs_rel;
    g_osqp_max_iter = osqp_max_iter;
}

void lcqpow_ref_set_debug(int qpoases_print_level, int store_steps)
{
    g_qpoases_print_level = qpoases_print_level;
    g_store_steps = store_steps;
}

/* Copies the iterates of the last solve: xsteps[step][nV], sub[step] (qp iterations of that pass),
 * alpha[step], stat[step].  Returns the number of recorded steps (may exceed cap: only cap are copied). */
int lcqpow_ref_last_trace(int cap, double* xsteps, int* sub, double* alpha, double* stat)
{
    const int ns = (int)g_trace_sub.size();
    for (int s = 0; s < ns && s < cap; s++) {
        if (xsteps) std::memcpy(xsteps + (size_t)s * g_trace_nV, &g_trace_x[(size_t)s * g_trace_nV], sizeof(double) * g_trace_nV);
        if (sub) sub[s] = g_trace_sub[s];
        if (alpha) alpha[s] = g_trace_alpha[s];
        if (stat) stat[s] = g_trace_stat[s];
    }
    return ns;
}

void lcqpow_ref_default_options(lcqp_ref_options* o)
{
    LCQPow::Options d;
    o->stationarityTolerance = d.getStationarityTolerance();
    o->complementarityTolerance = d.getComplementarityTolerance();
    o->initialPenaltyParameter = d.getInitialPenaltyParameter();
    o->penaltyUpdateFactor = d.getPenaltyUpdateFactor();
    o->maxPenaltyParameter = d.getMaxPenaltyParameter();
    o->etaDynamicPenalty = d.getEtaDynamicPenalty();
    o->solveZeroPenaltyFirst = d.getSolveZeroPenaltyFirst() ? 1 : 0;
    o->perturbStep = d.getPerturbStep() ? 1 : 0;
    o->maxIterations = d.getMaxIterations();
    o->nDynamicPenalty = d.getNDynamicPenalty();
    o->qpSolver = (int)d.getQPSolver();
    o->osqp_adaptive_rho_interval = -1;
}

/* Solve ONE dense LCQP with the reference.  Pointer arguments follow
 * LCQProblem::loadLCQP (LCQProblem.hpp:87-103): any of lbL..y0 may be NULL.
 * x must hold nV doubles, y must hold nV + nC + 2*nComp doubles. */
int lcqpow_ref_solve(int nV, int nC, int nComp,
                     const double* Q, const double* g, const double* L, const double* R,
                     const double* lbL, const double* ubL, const double* lbR, const double* ubR,
                     const double* A, const double* lbA, const double* ubA,
                     const double* lb, const double* ub,
                     const double* x0, const double* y0,
                     const lcqp_ref_options* o, double* x, double* y, lcqp_ref_result* res)
{
    using namespace LCQPow;
    std::memset(res, 0, sizeof(*res));

    auto t0 = std::chrono::steady_clock::now();

    LCQProblem lcqp(nV, nC, nComp);
    Options options;
    options.setPrintLevel(PrintLevel::NONE);
    options.setStationarityTolerance(o->stationarityTolerance);
    options.setComplementarityTolerance(o->complementarityTolerance);
    options.setInitialPenaltyParameter(o->initialPenaltyParameter);
    options.setPenaltyUpdateFactor(o->penaltyUpdateFactor);
    options.setMaxPenaltyParameter(o->maxPenaltyParameter);
    options.setEtaDynamicPenalty(o->etaDynamicPenalty);
    options.setSolveZeroPenaltyFirst(o->solveZeroPenaltyFirst != 0);
    options.setPerturbStep(o->perturbStep != 0);
    options.setMaxIterations(o->maxIterations);
    options.setNDynamicPenalty(o->nDynamicPenalty);
    options.setQPSolver(o->qpSolver);
    if (o->osqp_adaptive_rho_interval >= 0) {
        OSQPSettings* s = options.getOSQPOptions();
        s->adaptive_rho_interval = o->osqp_adaptive_rho_interval;
    }
    if (g_osqp_eps > 0.0 || g_osqp_max_iter > 0) {
        OSQPSettings* s = options.getOSQPOptions();
        if (g_osqp_eps > 0.0) { s->eps_abs = g_osqp_eps; s->eps_rel = g_osqp_eps; }
        if (g_osqp_max_iter > 0) s->max_iter = g_osqp_max_iter;
    }
    if (g_qpoases_termtol > 0.0 || g_qpoases_boundtol > 0.0) {
        qpOASES::Options qo = options.getqpOASESOptions();
        if (g_qpoases_termtol > 0.0) qo.terminationTolerance = g_qpoases_termtol;
        if (g_qpoases_boundtol > 0.0) qo.boundTolerance = g_qpoases_boundtol;
        options.setqpOASESOptions(qo);
    }
    if (g_qpoases_print_level != 0) {
        qpOASES::Options qo = options.getqpOASESOptions();
        qo.printLevel = (qpOASES::PrintLevel)g_qpoases_print_level;
        options.setqpOASESOptions(qo);
    }
    if (g_store_steps) options.setStoreSteps(true);
    lcqp.setOptions(options);

    ReturnValue rv = lcqp.loadLCQP(Q, g, L, R, lbL, ubL, lbR, ubR, A, lbA, ubA, lb, ub, x0, y0);
    if (rv != SUCCESSFUL_RETURN) {
        res->ret = (int)rv;
        return res->ret;
    }
    if (o->qpSolver >= (int)QPOASES_SPARSE) {
        rv = lcqp.switchToSparseMode();
        if (rv != SUCCESSFUL_RETURN) {
            res->ret = (int)rv;
            return res->ret;
        }
    }

    rv = lcqp.runSolver();
    auto t1 = std::chrono::steady_clock::now();

    OutputStatistics stats;
    lcqp.getOutputStatistics(stats);
    res->ret = (int)rv;
    res->status = (int)lcqp.getPrimalSolution(x);
    lcqp.getDualSolution(y);
    res->nDuals = lcqp.getNumberOfDuals();
    res->iterTotal = stats.getIterTotal();
    res->iterOuter = stats.getIterOuter();
    res->subproblemIter = stats.getSubproblemIter();
    res->qpExitFlag = stats.getQPSolverExitFlag();
    res->rhoOpt = stats.getRhoOpt();
    res->seconds = std::chrono::duration<double>(t1 - t0).count();
    if (g_store_steps) {
        g_trace_nV = nV;
        g_trace_x.clear();
        for (const std::vector<double>& xs : stats.getxStepsStdVec()) g_trace_x.insert(g_trace_x.end(), xs.begin(), xs.end());
        g_trace_sub = stats.getSubproblemItersStdVec();
        g_trace_alpha = stats.getStepLengthStdVec();
        g_trace_stat = stats.getStatValsStdVec();
    }
    return res->ret;
}

/* Solve a contiguous batch serially (this is how the CPU baseline is timed: one process per
 * core, each calling this on its slice).  Arrays flagged in shared_mask are read once, the
 * others are strided per instance.  Bit order: 0 Q,1 g,2 L,3 R,4 lbL,5 ubL,6 lbR,7 ubR,
 * 8 A,9 lbA,10 ubA,11 lb,12 ub,13 x0,14 y0. */
int lcqpow_ref_solve_batch(int batch, int nV, int nC, int nComp, unsigned shared_mask,
                           const double* Q, const double* g, const double* L, const double* R,
                           const double* lbL, const double* ubL, const double* lbR, const double* ubR,
                           const double* A, const double* lbA, const double* ubA,
                           const double* lb, const double* ub,
                           const double* x0, const double* y0,
                           const lcqp_ref_options* o, double* x, double* y, lcqp_ref_result* res)
{
    const size_t nD = (size_t)nV + nC + 2 * nComp;
    const size_t len[15] = {(size_t)nV * nV, (size_t)nV, (size_t)nComp * nV, (size_t)nComp * nV,
                            (size_t)nComp, (size_t)nComp, (size_t)nComp, (size_t)nComp,
                            (size_t)nC * nV, (size_t)nC, (size_t)nC, (size_t)nV, (size_t)nV,
                            (size_t)nV, nD};
    const double* base[15] = {Q, g, L, R, lbL, ubL, lbR, ubR, A, lbA, ubA, lb, ub, x0, y0};
    int nfail = 0;
    for (int b = 0; b < batch; b++) {
        const double* p[15];
        for (int k = 0; k < 15; k++) {
            if (!base[k]) p[k] = 0;
            else p[k] = (shared_mask >> k) & 1u ? base[k] : base[k] + (size_t)b * len[k];
        }
        int rv = lcqpow_ref_solve(nV, nC, nComp, p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7],
                                  p[8], p[9], p[10], p[11], p[12], p[13], p[14], o,
                                  x + (size_t)b * nV, y + (size_t)b * nD, res + b);
        if (rv != 0) nfail++;
    }
    return nfail;
}

} /* extern "C" */

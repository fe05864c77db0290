// Options.cpp -- see ../include/Options.hpp.  Validation rules and defaults follow
// /root/reference/src/Options.cpp (:85-259 setters, :296-333 defaults); the code is new.
#include "Options.hpp"

#include <cstdlib>
#include <cstring>

namespace LCQPow {

Options::Options() { setToDefault(); }

Options::Options(const Options& rhs) { copy(rhs); }

Options::~Options()
{
    std::free(OSQP_opts);
    OSQP_opts = nullptr;
}

Options& Options::operator=(const Options& rhs)
{
    if (this != &rhs) copy(rhs);
    return *this;
}

void Options::copy(const Options& rhs)
{
    stationarityTolerance = rhs.stationarityTolerance;
    complementarityTolerance = rhs.complementarityTolerance;
    initialPenaltyParameter = rhs.initialPenaltyParameter;
    penaltyUpdateFactor = rhs.penaltyUpdateFactor;
    solveZeroPenaltyFirst = rhs.solveZeroPenaltyFirst;
    perturbStep = rhs.perturbStep;
    maxIterations = rhs.maxIterations;
    maxPenaltyParameter = rhs.maxPenaltyParameter;
    nDynamicPenalty = rhs.nDynamicPenalty;
    etaDynamicPenalty = rhs.etaDynamicPenalty;
    storeSteps = rhs.storeSteps;
    qpSolver = rhs.qpSolver;
    printLevel = rhs.printLevel;
    qpOASES_opts = rhs.qpOASES_opts;
    perturbSeed = rhs.perturbSeed;
    osqpADMM = rhs.osqpADMM;
    device = rhs.device;
    setOSQPOptions(rhs.OSQP_opts);
}

void Options::setToDefault()
{
    complementarityTolerance = 1.0e3 * Utilities::EPS;
    stationarityTolerance = 1.0e6 * Utilities::EPS;
    initialPenaltyParameter = 0.01;
    penaltyUpdateFactor = 2.0;
    solveZeroPenaltyFirst = true;
    perturbStep = true;
    maxIterations = 1000;
    maxPenaltyParameter = 1e8;
    nDynamicPenalty = 3;
    etaDynamicPenalty = 0.9;
    printLevel = INNER_LOOP_ITERATES;
    storeSteps = false;
    qpSolver = QPOASES_DENSE;
    qpOASES_opts.setToDefault();
    qpOASES_opts.printLevel = qpOASES::PL_NONE;
    OSQPSettings def;
    osqp_set_default_settings(&def);
    def.eps_prim_inf = Utilities::EPS;
    def.verbose = 0;
    def.polish = 1;
    setOSQPOptions(&def);
    perturbSeed = 1;
    osqpADMM = false;
    device = 0;
}

static ReturnValue warn(ReturnValue r) { return MessageHandler::PrintMessage(r, WARNING); }

double Options::getStationarityTolerance() const { return stationarityTolerance; }
ReturnValue Options::setStationarityTolerance(double val)
{
    if (val <= Utilities::EPS) return warn(INVALID_STATIONARITY_TOLERANCE);
    stationarityTolerance = val;
    return SUCCESSFUL_RETURN;
}

double Options::getComplementarityTolerance() const { return complementarityTolerance; }
ReturnValue Options::setComplementarityTolerance(double val)
{
    if (val <= Utilities::EPS) return warn(INVALID_COMPLEMENTARITY_TOLERANCE);
    complementarityTolerance = val;
    return SUCCESSFUL_RETURN;
}

double Options::getInitialPenaltyParameter() const { return initialPenaltyParameter; }
ReturnValue Options::setInitialPenaltyParameter(double val)
{
    if (val <= Utilities::ZERO) return warn(INVALID_INITIAL_PENALTY_VALUE);
    initialPenaltyParameter = val;
    return SUCCESSFUL_RETURN;
}

double Options::getPenaltyUpdateFactor() const { return penaltyUpdateFactor; }
ReturnValue Options::setPenaltyUpdateFactor(double val)
{
    if (val <= 1) return warn(INVALID_PENALTY_UPDATE_VALUE);
    penaltyUpdateFactor = val;
    return SUCCESSFUL_RETURN;
}

bool Options::getSolveZeroPenaltyFirst() const { return solveZeroPenaltyFirst; }
ReturnValue Options::setSolveZeroPenaltyFirst(bool val) { solveZeroPenaltyFirst = val; return SUCCESSFUL_RETURN; }

bool Options::getPerturbStep() const { return perturbStep; }
ReturnValue Options::setPerturbStep(bool val) { perturbStep = val; return SUCCESSFUL_RETURN; }

int Options::getMaxIterations() const { return maxIterations; }
ReturnValue Options::setMaxIterations(int val)
{
    if (val <= 0) return warn(INVALID_MAX_ITERATIONS_VALUE);
    maxIterations = val;
    return SUCCESSFUL_RETURN;
}

double Options::getMaxPenaltyParameter() const { return maxPenaltyParameter; }
ReturnValue Options::setMaxPenaltyParameter(double val)
{
    if (val <= 0) return warn(INVALID_MAX_RHO_VALUE);
    maxPenaltyParameter = val;
    return SUCCESSFUL_RETURN;
}

int Options::getNDynamicPenalty() const { return nDynamicPenalty; }
ReturnValue Options::setNDynamicPenalty(int val) { nDynamicPenalty = val; return SUCCESSFUL_RETURN; }

double Options::getEtaDynamicPenalty() const { return etaDynamicPenalty; }
ReturnValue Options::setEtaDynamicPenalty(double val)
{
    if (val <= Utilities::EPS || val >= 1) return warn(INVALID_ETA_VALUE);
    etaDynamicPenalty = val;
    return SUCCESSFUL_RETURN;
}

PrintLevel Options::getPrintLevel() const { return printLevel; }
ReturnValue Options::setPrintLevel(PrintLevel val) { printLevel = val; return SUCCESSFUL_RETURN; }
ReturnValue Options::setPrintLevel(int val)
{
    if (val < NONE || val > INNER_LOOP_ITERATES) return warn(INVALID_PRINT_LEVEL_VALUE);
    printLevel = (PrintLevel)val;
    return SUCCESSFUL_RETURN;
}

bool Options::getStoreSteps() const { return storeSteps; }
ReturnValue Options::setStoreSteps(bool val) { storeSteps = val; return SUCCESSFUL_RETURN; }

QPSolver Options::getQPSolver() const { return qpSolver; }
ReturnValue Options::setQPSolver(QPSolver val) { qpSolver = val; return SUCCESSFUL_RETURN; }
// the reference accepts QPOASES_DENSE..OSQP_SPARSE (Options.cpp:253-259); the range is widened by CUDA_DENSE
ReturnValue Options::setQPSolver(int val)
{
    if (val < QPOASES_DENSE || val > CUDA_DENSE) return warn(INVALID_QPSOLVER);
    qpSolver = (QPSolver)val;
    return SUCCESSFUL_RETURN;
}

ReturnValue Options::setqpOASESOptions(const qpOASES::Options& _options)
{
    qpOASES_opts = _options;
    return SUCCESSFUL_RETURN;
}

qpOASES::Options& Options::getqpOASESOptions() { return qpOASES_opts; }

ReturnValue Options::setOSQPOptions(OSQPSettings* _options)
{
    OSQPSettings* fresh = nullptr;
    if (_options) {
        fresh = (OSQPSettings*)std::malloc(sizeof(OSQPSettings));
        if (fresh) std::memcpy(fresh, _options, sizeof(OSQPSettings));
    }
    std::free(OSQP_opts);
    OSQP_opts = fresh;
    return SUCCESSFUL_RETURN;
}

OSQPSettings* Options::getOSQPOptions() { return OSQP_opts; }

bool Options::getOSQPADMM() const { return osqpADMM; }
ReturnValue Options::setOSQPADMM(bool on) { osqpADMM = on; return SUCCESSFUL_RETURN; }

unsigned long long Options::getPerturbSeed() const { return perturbSeed; }
ReturnValue Options::setPerturbSeed(unsigned long long seed) { perturbSeed = seed; return SUCCESSFUL_RETURN; }

int Options::getDevice() const { return device; }
ReturnValue Options::setDevice(int dev)
{
    if (dev < 0) return warn(INVALID_ARGUMENT);
    device = dev;
    return SUCCESSFUL_RETURN;
}

void Options::toCuda(lcqp_cuda_options& o) const
{
    lcqp_cuda_default_options(&o);
    o.stationarityTolerance = stationarityTolerance;
    o.complementarityTolerance = complementarityTolerance;
    o.initialPenaltyParameter = initialPenaltyParameter;
    o.penaltyUpdateFactor = penaltyUpdateFactor;
    o.maxPenaltyParameter = maxPenaltyParameter;
    o.etaDynamicPenalty = etaDynamicPenalty;
    o.solveZeroPenaltyFirst = solveZeroPenaltyFirst ? 1 : 0;
    o.perturbStep = perturbStep ? 1 : 0;
    o.maxIterations = maxIterations;
    o.nDynamicPenalty = nDynamicPenalty;
    // dual layout: OSQP_SPARSE keeps the OSQP-style layout (no box duals); everything else is qpOASES-style
    o.qpSolver = (qpSolver == OSQP_SPARSE) ? 2 : 0;
    o.perturb_seed = perturbSeed;
    // subsolver pass-through: only values the user set (non-zero) override the defaults
    if (qpSolver == OSQP_SPARSE && OSQP_opts) {
        const OSQPSettings& s = *OSQP_opts;
        if (s.rho > 0) o.qp_rho = s.rho;
        if (s.sigma > 0) o.qp_sigma = s.sigma;
        if (s.alpha > 0 && s.alpha < 2) o.qp_alpha = s.alpha;
        if (s.delta > 0) o.qp_delta = s.delta;
        if (s.max_iter > 0) o.qp_max_iter = s.max_iter;
        if (s.check_termination > 0) o.qp_check_interval = s.check_termination;
        if (s.polish_refine_iter > 0) o.qp_refine_iter = s.polish_refine_iter < 2 ? 2 : s.polish_refine_iter;
        // the OSQP restatement reads the settings under their own names
        if (s.rho > 0) o.osqp_rho = s.rho;
        if (s.sigma > 0) o.osqp_sigma = s.sigma;
        if (s.alpha > 0 && s.alpha < 2) o.osqp_alpha = s.alpha;
        if (s.delta > 0) o.osqp_delta = s.delta;
        if (s.max_iter > 0) o.osqp_max_iter = s.max_iter;
        if (s.check_termination > 0) o.osqp_check_termination = s.check_termination;
        if (s.polish_refine_iter > 0) o.osqp_polish_refine_iter = s.polish_refine_iter;
        o.osqp_polish = s.polish ? 1 : 0;
        if (s.eps_prim_inf > 0) o.osqp_eps_prim_inf = s.eps_prim_inf;
        if (s.eps_abs > 0) o.osqp_eps_abs = s.eps_abs;
        if (s.eps_rel > 0) o.osqp_eps_rel = s.eps_rel;
        if (s.eps_dual_inf > 0) o.osqp_eps_dual_inf = s.eps_dual_inf;
        if (s.adaptive_rho_tolerance >= 1) o.osqp_adaptive_rho_tolerance = s.adaptive_rho_tolerance;
        if (s.scaling >= 0) o.osqp_scaling = s.scaling;
        if (s.adaptive_rho >= 0) o.osqp_adaptive_rho = s.adaptive_rho;
        if (s.adaptive_rho_interval >= 0) o.osqp_adaptive_rho_interval = s.adaptive_rho_interval;
    } else if (qpSolver != OSQP_SPARSE) {
        const qpOASES::Options& q = qpOASES_opts;
        if (q.terminationTolerance > 0) o.qp_dual_tol = q.terminationTolerance;
        if (q.boundTolerance > 0) o.qp_feas_tol = q.boundTolerance;
        if (q.enableRegularisation && q.epsRegularisation > 0) o.qp_delta = q.epsRegularisation;
        if (q.numRefinementSteps > 0) o.qp_refine_iter = q.numRefinementSteps < 2 ? 2 : q.numRefinementSteps;
        // the parametric active-set solver (positive definite reduced Hessians) reads qpOASES' own two tolerances
        if (q.terminationTolerance > 0) o.qpoases_terminationTolerance = q.terminationTolerance;
        if (q.boundTolerance > 0) o.qpoases_boundTolerance = q.boundTolerance;
    }
    o.osqp_admm = (qpSolver == OSQP_SPARSE && osqpADMM) ? 1 : 0;
}

}  // namespace LCQPow

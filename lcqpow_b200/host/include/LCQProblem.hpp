// LCQProblem.hpp -- LCQPow::LCQProblem for the B200 build: the reference's public API
// (/root/reference/include/LCQProblem.hpp:47-242) in front of the CUDA hot path.
//
//   min 1/2 x'Qx + g'x   s.t.  lbA <= Ax <= ubA,  lb <= x <= ub,
//                              lbL <= Lx,  lbR <= Rx,  (Lx - lbL)'(Rx - lbR) = 0
//
// runSolver() has two routes, both on the GPU, no CPU solver exists in this build:
//   * device loop (printLevel NONE and storeSteps off): the whole penalty homotopy of
//     LCQProblem::runSolver (reference src/LCQProblem.cpp:444-560) runs inside one persistent kernel --
//     the batched front door of the C ABI with a batch of one;
//   * host loop (printing or step tracking requested): the reference's loop structure on the host, every
//     convex QP solved by the SubsolverCUDA plugin through SubsolverBase::solve -- this is the drop-in the
//     reference's own runSolver would use if SubsolverCUDA were added to its Subsolver switch.
// LCQProblemBatch (LCQProblemBatch.hpp) is the same front door for many instances.
#ifndef LCQPOW_B200_LCQPROBLEM_HPP
#define LCQPOW_B200_LCQPROBLEM_HPP

#include <string>
#include <vector>

#include "Options.hpp"
#include "OutputStatistics.hpp"
#include "Subsolver.hpp"
#include "Utilities.hpp"

namespace LCQPow {

class LCQProblem {
public:
    LCQProblem();
    // nV variables, nC rows of A, nComp complementarity pairs (reference: LCQProblem.cpp:43-79)
    LCQProblem(int _nV, int _nC, int _nComp);
    virtual ~LCQProblem();

    // dense row-major data; NULL = absent (reference: LCQProblem.hpp:87-103, LCQProblem.cpp:87-144)
    ReturnValue loadLCQP(const double* const _Q, const double* const _g, const double* const _L, const double* const _R,
                         const double* const lbL = 0, const double* const ubL = 0, const double* const lbR = 0,
                         const double* const ubR = 0, const double* const _A = 0, const double* const _lbA = 0,
                         const double* const _ubA = 0, const double* const _lb = 0, const double* const _ub = 0,
                         const double* const _x0 = 0, const double* const _y0 = 0);
    // text files, one value per line (reference: LCQProblem.hpp:127-143, LCQProblem.cpp:147-306)
    ReturnValue loadLCQP(const char* const Q_file, const char* const g_file, const char* const L_file, const char* const R_file,
                         const char* const lbL_file = 0, const char* const ubL_file = 0, const char* const lbR_file = 0,
                         const char* const ubR_file = 0, const char* const A_file = 0, const char* const lbA_file = 0,
                         const char* const ubA_file = 0, const char* const lb_file = 0, const char* const ub_file = 0,
                         const char* const x0_file = 0, const char* const y0_file = 0);
    // csc data (reference: LCQProblem.hpp:166-182, LCQProblem.cpp:309-387)
    ReturnValue loadLCQP(const csc* const _Q, const double* const _g, const csc* const _L, const csc* const _R,
                         const double* const lbL = 0, const double* const ubL = 0, const double* const lbR = 0,
                         const double* const ubR = 0, const csc* const _A = 0, const double* const _lbA = 0,
                         const double* const _ubA = 0, const double* const _lb = 0, const double* const _ub = 0,
                         const double* const _x0 = 0, const double* const _y0 = 0);

    ReturnValue switchToSparseMode();   // reference :1037-1068
    ReturnValue switchToDenseMode();    // reference :1071-1102

    ReturnValue runSolver();

    virtual AlgorithmStatus getPrimalSolution(double* const xOpt) const;
    virtual AlgorithmStatus getDualSolution(double* const yOpt) const;
    int getNumberOfPrimals() const;
    virtual int getNumberOfDuals() const;
    virtual void getOutputStatistics(OutputStatistics& stats) const;
    void setOptions(const Options& _options);

    // the csc views that exist in sparse mode (A_full = [A; L; R]); NULL in dense mode
    const csc* getSparseQ() const { return Q_sparse; }
    const csc* getSparseA() const { return A_sparse; }
    const csc* getSparseC() const { return C_sparse; }
    bool isSparseMode() const { return sparseSolver; }
    // text of the last C-ABI error of the device-loop route
    const char* getLastDeviceError() const { return deviceError.c_str(); }

protected:
    void clear();
    ReturnValue setConstraints(const double* L_new, const double* R_new, const double* A_new, const double* lbA_new,
                               const double* ubA_new);
    ReturnValue setComplementarityBounds(const double* lbL_new, const double* ubL_new, const double* lbR_new,
                                         const double* ubR_new);
    ReturnValue runDeviceLoop();
    ReturnValue runHostLoop();

    // host loop helpers (reference: LCQProblem.cpp:1105-1482)
    void applyQk(const double* v, const double* add, double* out);   // Q v + rho C v + add
    double getObj();
    double getPhi();
    void printIteration(int outerIter, int innerIter, double statInf, double phi, double rho, double pNorm, double alphak, int qpIterk);
    void printHeader();
    void printLine();

    int nV = 0, nC = 0, nComp = 0, nDuals = 0, boxDualOffset = 0;
    bool sparseSolver = false;
    bool loaded = false;

    // deep copies of the problem data, dense row-major (always kept: the device consumes dense fp64)
    std::vector<double> Q, g, L, R, A, Afull;           // Afull = [A; L; R], (nC + 2 nComp) x nV
    std::vector<double> lbA, ubA;                        // nC + 2 nComp, as setConstraints/setComplementarityBounds build them
    std::vector<double> lbL, ubL, lbR, ubR, lbAuser, ubAuser, lb, ub, x0, y0;
    bool has_lbL = false, has_ubL = false, has_lbR = false, has_ubR = false, has_lbA = false, has_ubA = false, has_lb = false,
         has_ub = false, has_x0 = false, has_y0 = false;
    std::vector<double> C;                               // L'R + R'L (nV x nV)
    csc *Q_sparse = nullptr, *A_sparse = nullptr, *L_sparse = nullptr, *R_sparse = nullptr, *C_sparse = nullptr;

    // results
    std::vector<double> xk, yk;
    AlgorithmStatus algoStat = PROBLEM_NOT_SOLVED;
    OutputStatistics stats;
    Options options;
    Subsolver subsolver;
    std::string deviceError;

    // host-loop state
    double rho = 0.0;
    std::vector<double> gphi;
    double phi_const = 0.0;
};

}  // namespace LCQPow

#endif

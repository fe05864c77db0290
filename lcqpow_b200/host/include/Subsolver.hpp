// Subsolver.hpp -- the by-value subsolver holder LCQProblem owns (reference: include/Subsolver.hpp:33-114,
// src/Subsolver.cpp:36-136).  In the reference it is a tagged union over SubsolverQPOASES / SubsolverOSQP that
// switches on the QPSolver enum; here every enum value is served by SubsolverCUDA -- the value only selects
// the dual layout LCQProblem uses around it (qpOASES-style or OSQP-style).
#ifndef LCQPOW_B200_SUBSOLVER_HPP
#define LCQPOW_B200_SUBSOLVER_HPP

#include "SubsolverCUDA.hpp"

namespace LCQPow {

class Subsolver {
public:
    Subsolver();
    // dense data (reference: Subsolver(int nV, int nC, double* Q, double* A), Subsolver.cpp:36-42)
    Subsolver(int nV, int nC, const double* Q, const double* A, QPSolver qpSolver = QPOASES_DENSE, int device = 0);
    // sparse data (reference: Subsolver(int, int, csc*, csc*, QPSolver) :44-64 and Subsolver(const csc*, const csc*))
    Subsolver(int nV, int nC, const csc* Q, const csc* A, QPSolver qpSolver, int device = 0);
    Subsolver(const csc* Q, const csc* A, int device = 0);
    Subsolver(const Subsolver& rhs);
    virtual ~Subsolver();
    virtual Subsolver& operator=(const Subsolver& rhs);

    void getSolution(double* x, double* y);
    ReturnValue solve(bool initialSolve, int& iterations, int& exit_flag, const double* const g, const double* const lbA,
                      const double* const ubA, const double* const x0 = 0, const double* const y0 = 0,
                      const double* const lb = 0, const double* const ub = 0);

    void setOptions(const Options& options);
    void setOptions(qpOASES::Options& options);   // pass-through of the subsolver knobs (Subsolver.hpp:92-96)
    void setOptions(OSQPSettings* settings);

    QPSolver getQPSolver() const { return qpSolver; }
    bool isValid() const { return solverCUDA.isValid(); }

protected:
    void copy(const Subsolver& rhs);

private:
    QPSolver qpSolver = CUDA_DENSE;
    SubsolverCUDA solverCUDA;
    Options opts;
};

}  // namespace LCQPow

#endif

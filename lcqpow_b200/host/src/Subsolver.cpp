// Subsolver.cpp -- see ../include/Subsolver.hpp.
#include "Subsolver.hpp"

namespace LCQPow {

Subsolver::Subsolver() {}

Subsolver::Subsolver(int nV, int nC, const double* Q, const double* A, QPSolver _qpSolver, int device)
    : qpSolver(_qpSolver), solverCUDA(nV, nC, Q, A, device)
{
}

Subsolver::Subsolver(int nV, int nC, const csc* Q, const csc* A, QPSolver _qpSolver, int device)
    : qpSolver(_qpSolver), solverCUDA(Q, A, device)
{
    (void)nV;
    (void)nC;
}

Subsolver::Subsolver(const csc* Q, const csc* A, int device) : qpSolver(OSQP_SPARSE), solverCUDA(Q, A, device) {}

Subsolver::Subsolver(const Subsolver& rhs) { copy(rhs); }

Subsolver::~Subsolver() {}

Subsolver& Subsolver::operator=(const Subsolver& rhs)
{
    if (this != &rhs) copy(rhs);
    return *this;
}

void Subsolver::copy(const Subsolver& rhs)
{
    qpSolver = rhs.qpSolver;
    opts = rhs.opts;
    solverCUDA = rhs.solverCUDA;   // a fresh device solver on the same data (reference: Subsolver.cpp:125-136)
}

void Subsolver::getSolution(double* x, double* y) { solverCUDA.getSolution(x, y); }

ReturnValue Subsolver::solve(bool initialSolve, int& iterations, int& exit_flag, const double* const g,
                             const double* const lbA, const double* const ubA, const double* const x0,
                             const double* const y0, const double* const lb, const double* const ub)
{
    return solverCUDA.solve(initialSolve, iterations, exit_flag, g, lbA, ubA, x0, y0, lb, ub);
}

void Subsolver::setOptions(const Options& options)
{
    opts = options;
    solverCUDA.setOptions(opts);
}

void Subsolver::setOptions(qpOASES::Options& options)
{
    opts.setqpOASESOptions(options);
    solverCUDA.setOptions(opts);
}

void Subsolver::setOptions(OSQPSettings* settings)
{
    opts.setOSQPOptions(settings);
    solverCUDA.setOptions(opts);
}

}  // namespace LCQPow

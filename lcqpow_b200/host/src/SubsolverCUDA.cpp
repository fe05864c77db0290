// SubsolverCUDA.cpp -- see ../include/SubsolverCUDA.hpp.
#include "SubsolverCUDA.hpp"

#include <cstdlib>
#include <cstring>

namespace LCQPow {

SubsolverCUDA::SubsolverCUDA() {}

SubsolverCUDA::SubsolverCUDA(int _nV, int _nC, const double* _Q, const double* _A, int _device)
{
    nV = _nV;
    nC = _nC;
    device = _device;
    Q = new double[(size_t)nV * nV];
    std::memcpy(Q, _Q, sizeof(double) * (size_t)nV * nV);
    if (nC > 0 && _A) {
        A = new double[(size_t)nC * nV];
        std::memcpy(A, _A, sizeof(double) * (size_t)nC * nV);
    }
    create();
}

SubsolverCUDA::SubsolverCUDA(const csc* _Q, const csc* _A, int _device)
{
    nV = _Q->n;
    nC = _A ? _A->m : 0;
    device = _device;
    Q = Utilities::csc_to_dns(_Q);
    // OSQP keeps the upper triangle of a symmetric Hessian only (SubsolverOSQP.cpp:44-45); a full csc is
    // accepted as it is, an upper-triangular one is mirrored
    bool lowerEmpty = true;
    for (int r = 1; r < nV && lowerEmpty; ++r)
        for (int c = 0; c < r; ++c)
            if (Q[(size_t)r * nV + c] != 0.0) { lowerEmpty = false; break; }
    if (lowerEmpty)
        for (int r = 1; r < nV; ++r)
            for (int c = 0; c < r; ++c) Q[(size_t)r * nV + c] = Q[(size_t)c * nV + r];
    if (nC > 0) A = Utilities::csc_to_dns(_A);
    create();
}

SubsolverCUDA::SubsolverCUDA(const SubsolverCUDA& rhs) { copy(rhs); }

SubsolverCUDA::~SubsolverCUDA() { clear(); }

SubsolverCUDA& SubsolverCUDA::operator=(const SubsolverCUDA& rhs)
{
    if (this != &rhs) {
        clear();
        copy(rhs);
    }
    return *this;
}

void SubsolverCUDA::clear()
{
    if (handle) lcqp_cuda_qp_destroy(handle);
    handle = nullptr;
    delete[] Q;
    delete[] A;
    Q = nullptr;
    A = nullptr;
}

// Like the reference's subsolvers, a copy is a fresh solver on the same data (Subsolver.cpp:125-136): the
// factorisation and the hot-start state are NOT shared.
void SubsolverCUDA::copy(const SubsolverCUDA& rhs)
{
    nV = rhs.nV;
    nC = rhs.nC;
    device = rhs.device;
    opts = rhs.opts;
    haveOpts = rhs.haveOpts;
    if (rhs.Q) {
        Q = new double[(size_t)nV * nV];
        std::memcpy(Q, rhs.Q, sizeof(double) * (size_t)nV * nV);
    }
    if (rhs.A) {
        A = new double[(size_t)nC * nV];
        std::memcpy(A, rhs.A, sizeof(double) * (size_t)nC * nV);
    }
    if (rhs.Q) create();
}

void SubsolverCUDA::create()
{
    lastCode = lcqp_cuda_qp_create(nV, nC, Q, A, device, &handle);
    if (lastCode != LCQP_CUDA_OK) {
        handle = nullptr;
        return;
    }
    if (haveOpts) lcqp_cuda_qp_set_options(handle, &opts);
}

void SubsolverCUDA::setOptions(const lcqp_cuda_options& options)
{
    opts = options;
    haveOpts = true;
    if (handle) lastCode = lcqp_cuda_qp_set_options(handle, &opts);
}

void SubsolverCUDA::setOptions(const Options& options)
{
    lcqp_cuda_options o;
    options.toCuda(o);
    setOptions(o);
}

ReturnValue SubsolverCUDA::solve(bool initialSolve, int& iterations, int& exit_flag, const double* const g,
                                 const double* const lbA, const double* const ubA, const double* const x0,
                                 const double* const y0, const double* const lb, const double* const ub)
{
    iterations = 0;
    if (!handle) {
        // no usable device: there is no CPU path behind this plugin
        exit_flag = lastCode ? lastCode : LCQP_CUDA_NO_DEVICE;
        return SUBPROBLEM_SOLVER_ERROR;
    }
    lastCode = lcqp_cuda_qp_solve(handle, initialSolve ? 1 : 0, &iterations, &exit_flag, g, lbA, ubA, x0, y0, lb, ub);
    if (lastCode == LCQP_CUDA_OK) return SUCCESSFUL_RETURN;
    if (lastCode >= LCQP_CUDA_NO_DEVICE) exit_flag = lastCode;
    return SUBPROBLEM_SOLVER_ERROR;
}

void SubsolverCUDA::getSolution(double* x, double* y)
{
    if (!handle) return;
    lastCode = lcqp_cuda_qp_get_solution(handle, x, y);
}

}  // namespace LCQPow

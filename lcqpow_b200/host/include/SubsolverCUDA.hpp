// SubsolverCUDA.hpp -- the B200 QP subsolver plugin (north_star: "a new SubsolverCUDA plugin").
//
// Sits where SubsolverQPOASES / SubsolverOSQP sit in the reference (include/SubsolverQPOASES.hpp:33-146,
// include/SubsolverOSQP.hpp:34-117): constructed from (nV, nC, Q, A) dense row-major -- nC counts ALL rows
// of A_full = [A; L; R] -- or from csc matrices; solve()/getSolution() as SubsolverBase defines them.
// All arithmetic happens in liblcqp_cuda.so through the plugin door of the C ABI
// (lcqp_cuda_qp_create / _solve / _get_solution / _destroy, include/lcqp_cuda.h); no CPU solver is linked.
#ifndef LCQPOW_B200_SUBSOLVERCUDA_HPP
#define LCQPOW_B200_SUBSOLVERCUDA_HPP

#include "Options.hpp"
#include "SubsolverBase.hpp"

namespace LCQPow {

class SubsolverCUDA : public SubsolverBase {
public:
    SubsolverCUDA();
    // like SubsolverQPOASES(int nV, int nC, double* Q, double* A) (SubsolverQPOASES.hpp:44-47)
    SubsolverCUDA(int nV, int nC, const double* Q, const double* A, int device = 0);
    // like SubsolverOSQP(const csc* Q, const csc* A) (SubsolverOSQP.hpp:46-48): Q nV x nV, A nC x nV
    SubsolverCUDA(const csc* Q, const csc* A, int device = 0);
    SubsolverCUDA(const SubsolverCUDA& rhs);
    virtual ~SubsolverCUDA();
    virtual SubsolverCUDA& operator=(const SubsolverCUDA& rhs);

    // the knobs of the device solver (qp_* fields) and the tolerances
    void setOptions(const lcqp_cuda_options& options);
    void setOptions(const Options& options);

    ReturnValue solve(bool initialSolve, int& iterations, int& exit_flag, const double* const g, const double* const lbA,
                      const double* const ubA, const double* const x0 = 0, const double* const y0 = 0,
                      const double* const lb = 0, const double* const ub = 0) override;
    void getSolution(double* x, double* y) override;

    bool isValid() const { return handle != nullptr; }
    int lastCudaCode() const { return lastCode; }   // return code of the last C-ABI call (0 or >= 500)

protected:
    void copy(const SubsolverCUDA& rhs);
    void clear();
    void create();

private:
    int nV = 0;
    int nC = 0;
    int device = 0;
    double* Q = nullptr;   // deep copies, as every layer of the reference keeps its own (SubsolverQPOASES.cpp:41-45)
    double* A = nullptr;
    lcqp_cuda_options opts;
    bool haveOpts = false;
    lcqp_cuda_qp handle = nullptr;
    int lastCode = 0;
};

}  // namespace LCQPow

#endif

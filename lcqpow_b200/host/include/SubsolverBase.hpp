// SubsolverBase.hpp -- the QP subsolver plugin interface of LCQPow, unchanged in shape:
// /root/reference/include/SubsolverBase.hpp:28-58.  A plugin solves the convex QP
//     min 1/2 x'Qx + g'x   s.t.  lbA <= A x <= ubA,  lb <= x <= ub
// for a fixed (Q, A) and is called repeatedly with new g (and formally new bounds): the first call has
// initialSolve = true and may use (x0, y0); later calls hot-start from the plugin's own previous solution.
#ifndef LCQPOW_B200_SUBSOLVERBASE_HPP
#define LCQPOW_B200_SUBSOLVERBASE_HPP

#include "Utilities.hpp"

namespace LCQPow {

class SubsolverBase {
public:
    virtual ~SubsolverBase() {}

    // x[nV]; y[nV + nC] with the qpOASES sign convention Qx + g = A'y_A + y_box (box duals first)
    virtual void getSolution(double* x, double* y) = 0;

    // iterations: inner iterations of this call; exit_flag: backend status (0 = solved)
    virtual ReturnValue solve(bool initialSolve, int& iterations, int& exit_flag, const double* const _g,
                              const double* const _lbA, const double* const _ubA, const double* const x0 = 0,
                              const double* const y0 = 0, const double* const _lb = 0, const double* const _ub = 0) = 0;
};

}  // namespace LCQPow

#endif

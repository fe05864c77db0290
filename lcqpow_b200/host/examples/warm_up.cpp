// warm_up -- the reference's first example (/root/reference/examples/warm_up.cpp) against the B200 build:
// same API calls, the QPs are solved by the SubsolverCUDA plugin (host loop, iteration table printed).
#include <cstdio>

#include "LCQProblem.hpp"
#include "problems.hpp"

using namespace LCQPow;

int main()
{
    const examples::Problem p = examples::warmUp();
    LCQProblem lcqp(p.nV, p.nC, p.nComp);
    Options options;
    options.setPrintLevel(PrintLevel::INNER_LOOP_ITERATES);
    options.setQPSolver(QPSolver::QPOASES_DENSE);
    lcqp.setOptions(options);

    ReturnValue ret = lcqp.loadLCQP(p.Q.data(), p.g.data(), p.L.data(), p.R.data(), 0, 0, 0, 0, 0, 0, 0, 0, 0, p.x0.data(),
                                    p.y0.data());
    if (ret != SUCCESSFUL_RETURN) { std::printf("Failed to load LCQP.\n"); return 1; }
    ret = lcqp.runSolver();
    if (ret != SUCCESSFUL_RETURN) { std::printf("Failed to solve LCQP (%d).\n", (int)ret); return 1; }

    double x[2], y[4];
    OutputStatistics stats;
    lcqp.getPrimalSolution(x);
    lcqp.getDualSolution(y);
    lcqp.getOutputStatistics(stats);
    std::printf("\nxOpt = [ %g, %g ];  yOpt = [ %g, %g, %g, %g ]; i = %d; k = %d; rho = %g; WSR = %d \n\n", x[0], x[1], y[0], y[1],
                y[2], y[3], stats.getIterTotal(), stats.getIterOuter(), stats.getRhoOpt(), stats.getSubproblemIter());
    return 0;
}

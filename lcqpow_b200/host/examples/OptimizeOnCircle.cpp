// OptimizeOnCircle -- the reference's second example (/root/reference/examples/OptimizeOnCircle.cpp):
// N = 100 facets, stationarityTolerance 1e-2, sparse mode, OSQP-style dual layout, iteration table.
#include <cstdio>
#include <vector>

#include "LCQProblem.hpp"
#include "problems.hpp"

using namespace LCQPow;

int main()
{
    std::printf("Preparing unit circle optimization problem...\n");
    const examples::Problem p = examples::circle(100, 0.5, -0.6);
    LCQProblem lcqp(p.nV, p.nC, p.nComp);
    Options options;
    options.setPrintLevel(PrintLevel::INNER_LOOP_ITERATES);
    options.setQPSolver(QPSolver::OSQP_SPARSE);
    options.setStationarityTolerance(10e-3);
    lcqp.setOptions(options);

    ReturnValue ret = lcqp.loadLCQP(p.Q.data(), p.g.data(), p.L.data(), p.R.data(), 0, 0, 0, 0, p.A.data(), p.lbA.data(),
                                    p.ubA.data(), 0, 0, p.x0.data());
    if (ret != SUCCESSFUL_RETURN) { std::printf("Failed to load LCQP.\n"); return 1; }
    if (options.getQPSolver() >= QPOASES_SPARSE && lcqp.switchToSparseMode() != SUCCESSFUL_RETURN) {
        std::printf("Failed to switch to sparse mode LCQP.\n");
        return 1;
    }
    ret = lcqp.runSolver();
    if (ret != SUCCESSFUL_RETURN) { std::printf("Failed to solve LCQP (%d).\n", (int)ret); return 1; }

    std::vector<double> x(p.nV), y(p.nV + p.nC + 2 * p.nComp);
    OutputStatistics stats;
    lcqp.getPrimalSolution(x.data());
    lcqp.getDualSolution(y.data());
    lcqp.getOutputStatistics(stats);
    std::printf("\nxOpt = [ %g, %g ];  i = %d; k = %d; rho = %g; WSR = %d \n\n", x[0], x[1], stats.getIterTotal(),
                stats.getIterOuter(), stats.getRhoOpt(), stats.getSubproblemIter());
    std::printf("For reference: Global solution is at:  [ %g, %g ]\n", 0.1811, -0.9835);
    std::printf("               Another local solution: [ %g, %g ]\n", 0.9764, -0.2183);
    return 0;
}

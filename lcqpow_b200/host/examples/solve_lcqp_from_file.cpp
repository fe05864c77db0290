// solve_lcqp_from_file -- the reference's third example (/root/reference/examples/solve_lcqp_from_file.cpp):
// reads Q, g, L, R, A, bounds, x0 from text files (one value per line) in the directory given as argv[1]
// (default: example_data), dimensions from the file lengths, solves on the device loop.
#include <cstdio>
#include <fstream>
#include <string>
#include <vector>

#include "LCQProblem.hpp"

using namespace LCQPow;

static int countLines(const std::string& f)
{
    std::ifstream in(f);
    if (!in) return -1;
    int k = 0;
    std::string tok;
    while (in >> tok) ++k;
    return k;
}

static bool exists(const std::string& f) { return std::ifstream(f).good(); }

int main(int argc, char** argv)
{
    const std::string dir = (argc > 1 ? argv[1] : "example_data");
    auto path = [&](const char* name) { return dir + "/" + name + ".txt"; };
    const int nV = countLines(path("g"));
    if (nV <= 0) { std::printf("Could not read %s\n", path("g").c_str()); return 1; }
    const int nComp = countLines(path("L")) / nV;
    const int nC = exists(path("A")) ? countLines(path("A")) / nV : 0;
    std::printf("nV = %d, nC = %d, nComp = %d\n", nV, nC, nComp);

    std::string f[15];
    const char* names[15] = {"Q", "g", "L", "R", "lbL", "ubL", "lbR", "ubR", "A", "lbA", "ubA", "lb", "ub", "x0", "y0"};
    const char* fp[15];
    for (int k = 0; k < 15; ++k) {
        f[k] = path(names[k]);
        fp[k] = exists(f[k]) ? f[k].c_str() : nullptr;
    }
    LCQProblem lcqp(nV, nC, nComp);
    Options options;
    options.setPrintLevel(argc > 2 ? PrintLevel::INNER_LOOP_ITERATES : PrintLevel::NONE);
    options.setQPSolver(QPSolver::QPOASES_SPARSE);
    lcqp.setOptions(options);
    ReturnValue ret = lcqp.loadLCQP(fp[0], fp[1], fp[2], fp[3], fp[4], fp[5], fp[6], fp[7], fp[8], fp[9], fp[10], fp[11], fp[12],
                                    fp[13], fp[14]);
    if (ret != SUCCESSFUL_RETURN) { std::printf("Failed to load LCQP.\n"); return 1; }
    if (lcqp.switchToSparseMode() != SUCCESSFUL_RETURN) { std::printf("Failed to switch to sparse mode.\n"); return 1; }
    ret = lcqp.runSolver();
    if (ret != SUCCESSFUL_RETURN) { std::printf("Failed to solve LCQP (%d).\n", (int)ret); return 1; }
    std::vector<double> x(nV);
    OutputStatistics stats;
    const AlgorithmStatus st = lcqp.getPrimalSolution(x.data());
    lcqp.getOutputStatistics(stats);
    std::printf("status = %d; x[0..2] = [ %.10g, %.10g, %.10g ]; i = %d; k = %d; rho = %g; WSR = %d\n", (int)st, x[0],
                nV > 1 ? x[1] : 0.0, nV > 2 ? x[2] : 0.0, stats.getIterTotal(), stats.getIterOuter(), stats.getRhoOpt(),
                stats.getSubproblemIter());
    return 0;
}

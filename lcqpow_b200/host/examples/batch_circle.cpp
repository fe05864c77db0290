// batch_circle -- the batched front door: B OptimizeOnCircle-shaped LCQPs that share Q/A/L/R/lbA/ubA and
// differ in g and x0 (config C2 of SURVEY.md 8d), sharded over the GPUs given on the command line.
//   batch_circle [batch=4096] [gpu ...]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "LCQProblemBatch.hpp"
#include "problems.hpp"

using namespace LCQPow;

int main(int argc, char** argv)
{
    const int B = argc > 1 ? std::atoi(argv[1]) : 4096;
    std::vector<int> devs;
    for (int k = 2; k < argc; ++k) devs.push_back(std::atoi(argv[k]));
    const examples::Problem p = examples::circle(100, 0.5, -0.6);
    const int n = p.nV;
    std::vector<double> g((size_t)B * n, 0.0), x0((size_t)B * n, 1.0);
    std::mt19937_64 rng(20000);
    std::uniform_real_distribution<double> U(-1.0, 1.0);
    for (int b = 0; b < B; ++b) {
        double xr = 0.5, yr = -0.6;
        if (b > 0) do { xr = U(rng); yr = U(rng); } while (xr * xr + yr * yr > 0.95 * 0.95);
        g[(size_t)b * n + 0] = -(17.0 * xr - 15.0 * yr);
        g[(size_t)b * n + 1] = -(-15.0 * xr + 17.0 * yr);
        x0[(size_t)b * n + 0] = xr;
        x0[(size_t)b * n + 1] = yr;
    }
    LCQProblemBatch batch(p.nV, p.nC, p.nComp, B, devs);
    if (!batch.isValid()) { std::printf("no usable device: %s\n", batch.getLastError().c_str()); return 1; }
    Options options;
    options.setStationarityTolerance(10e-3);
    options.setPrintLevel(PrintLevel::NONE);
    batch.setOptions(options);
    const unsigned shared = (1u << LCQP_Q) | (1u << LCQP_L) | (1u << LCQP_R) | (1u << LCQP_A) | (1u << LCQP_LBA) | (1u << LCQP_UBA);
    ReturnValue ret = batch.loadLCQP(shared, p.Q.data(), g.data(), p.L.data(), p.R.data(), 0, 0, 0, 0, p.A.data(), p.lbA.data(),
                                     p.ubA.data(), 0, 0, x0.data());
    if (ret != SUCCESSFUL_RETURN) { std::printf("load failed: %s\n", batch.getLastError().c_str()); return 1; }
    const auto t0 = std::chrono::steady_clock::now();
    ret = batch.runSolver();
    const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (ret != SUCCESSFUL_RETURN) { std::printf("run failed: %s\n", batch.getLastError().c_str()); return 1; }
    std::vector<double> x((size_t)B * n);
    batch.getPrimalSolution(x.data());
    const std::vector<int> rv = batch.getReturnValues();
    int solved = 0;
    for (int r : rv) solved += (r == SUCCESSFUL_RETURN);
    std::printf("%d of %d LCQPs solved on %d GPU(s) in %.3f s (%.0f LCQP/s); instance 0: x = [ %.9f, %.9f ]\n", solved, B,
                batch.getNumberOfShards(), sec, solved / sec, x[0], x[1]);
    return 0;
}

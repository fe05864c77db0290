// host_tests -- unit and integration tests of the C++ host side, modelled on the reference's own suite
// (/root/reference/test/RunUnitTests.cpp and test/examples/*.cpp): same fixtures, same expected values.
//   host_tests --cpu   everything that needs no GPU (Utilities known answers, Options, OutputStatistics,
//                      csc round trips, load errors, "no device" is loud)
//   host_tests --gpu   the solver tests (RunWarmUp, CheckQPReturnFlag, DenseToSparse, test_max_penalty,
//                      OptimizeOnCircle, host loop vs device loop, batched door + sharding, plugin door)
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include "LCQProblem.hpp"
#include "LCQProblemBatch.hpp"
#include "../examples/problems.hpp"

using namespace LCQPow;

static int g_fail = 0, g_checks = 0;
static const char* g_test = "";
#define CHECK(cond)                                                                                  \
    do {                                                                                             \
        ++g_checks;                                                                                  \
        if (!(cond)) { ++g_fail; std::printf("FAILED %s:%d [%s]: %s\n", __FILE__, __LINE__, g_test, #cond); } \
    } while (0)
#define CHECK_EQ(a, b) CHECK((a) == (b))
#define CHECK_NEAR(a, b, tol) CHECK(std::fabs((a) - (b)) <= (tol))
#define TEST(name) g_test = name

// ---- Utilities known answers (RunUnitTests.cpp:33-246) ------------------------------------------------------
static void testUtilities()
{
    TEST("Utilities.MatrixMultiplication");   // :33-57
    {
        const double A[6] = {1, 0, 2, 3, 1, 1}, B[12] = {2, 0, 0, 2, 1, 0, 0, 1, 0, -1, -1, 0};
        double C[8];
        Utilities::MatrixMultiplication(A, B, C, 2, 3, 4);
        const double want[8] = {2, -2, -2, 2, 7, -1, -1, 7};
        for (int i = 0; i < 8; ++i) CHECK_EQ(C[i], want[i]);
    }
    TEST("Utilities.TransposedMatrixMultiplication");   // :60-78
    {
        const double A[6] = {1, 0, 2, 3, 1, 1}, b[2] = {98, -10};
        double c[3];
        Utilities::TransponsedMatrixMultiplication(A, b, c, 2, 3, 1);
        CHECK_EQ(c[0], 68.0); CHECK_EQ(c[1], -10.0); CHECK_EQ(c[2], 186.0);
    }
    TEST("Utilities.MatrixSymmetrization");   // :81-104
    {
        const double A[6] = {1, 0, 2, 3, 1, 1}, B[6] = {2, 0, 1, 0, 0, -1};
        double C[9];
        Utilities::MatrixSymmetrizationProduct(A, B, C, 2, 3);
        const double want[9] = {4, 0, 2, 0, 0, -1, 2, -1, 2};
        for (int i = 0; i < 9; ++i) CHECK_EQ(C[i], want[i]);
    }
    TEST("Utilities.AffineTransformation");   // :107-129
    {
        const double A[6] = {1, 0, 2, 3, 1, 1}, b[3] = {1, 1, 1}, c[2] = {-1, 1};
        double d[2];
        Utilities::AffineLinearTransformation(2, A, b, c, d, 2, 3);
        CHECK_EQ(d[0], 5.0); CHECK_EQ(d[1], 11.0);
    }
    TEST("Utilities.VectorAdd/DotProduct/QuadraticForm/MaxAbs");   // :162-246
    {
        const double a[3] = {1, 2, 3}, b[3] = {3, 1, 2};
        double c[3];
        Utilities::WeightedVectorAdd(2, a, -1, b, c, 3);
        CHECK_EQ(c[0], -1.0); CHECK_EQ(c[1], 3.0); CHECK_EQ(c[2], 4.0);
        CHECK_EQ(Utilities::DotProduct(a, b, 3), 11.0);
        const double Q[4] = {2, 1, 1, 2}, p[2] = {2, 2};
        CHECK_EQ(Utilities::QuadraticFormProduct(Q, p, 2), 24.0);
        const double v[4] = {1, -7, 3, 5};
        CHECK_EQ(Utilities::MaxAbs(v, 4), 7.0);
    }
    TEST("Utilities.csc round trip");   // :265-375
    {
        std::vector<double> M(7 * 5, 0.0);
        unsigned s = 12345u;
        for (double& v : M) {
            s = s * 1664525u + 1013904223u;
            if ((s >> 28) < 5) v = (double)((int)((s >> 8) % 19) - 9);
        }
        csc* S = Utilities::dns_to_csc(M.data(), 7, 5);
        CHECK(S != nullptr);
        double* D = Utilities::csc_to_dns(S);
        for (size_t i = 0; i < M.size(); ++i) CHECK_EQ(D[i], M[i]);
        csc* S2 = Utilities::copyCSC(S);
        CHECK_EQ(S2->p[5], S->p[5]);
        delete[] D;
        Utilities::ClearSparseMat(&S);
        Utilities::ClearSparseMat(&S2);
        CHECK(S == nullptr);
    }
}

static void testOptionsAndStatistics()
{
    TEST("Options.defaults");   // Options.cpp:296-333
    Options o;
    CHECK_NEAR(o.getComplementarityTolerance(), 1e3 * Utilities::EPS, 1e-30);
    CHECK_NEAR(o.getStationarityTolerance(), 1e6 * Utilities::EPS, 1e-30);
    CHECK_EQ(o.getInitialPenaltyParameter(), 0.01);
    CHECK_EQ(o.getPenaltyUpdateFactor(), 2.0);
    CHECK(o.getSolveZeroPenaltyFirst() && o.getPerturbStep() && !o.getStoreSteps());
    CHECK_EQ(o.getMaxIterations(), 1000);
    CHECK_EQ(o.getMaxPenaltyParameter(), 1e8);
    CHECK_EQ(o.getNDynamicPenalty(), 3);
    CHECK_EQ(o.getEtaDynamicPenalty(), 0.9);
    CHECK_EQ(o.getPrintLevel(), INNER_LOOP_ITERATES);
    CHECK_EQ(o.getQPSolver(), QPOASES_DENSE);
    TEST("Options.validation");   // Options.cpp:85-259
    std::printf("(the WARNING lines below are expected)\n");
    CHECK_EQ(o.setStationarityTolerance(1e-17), INVALID_STATIONARITY_TOLERANCE);
    CHECK_EQ(o.setComplementarityTolerance(0.0), INVALID_COMPLEMENTARITY_TOLERANCE);
    CHECK_EQ(o.setInitialPenaltyParameter(0.0), INVALID_INITIAL_PENALTY_VALUE);
    CHECK_EQ(o.setPenaltyUpdateFactor(1.0), INVALID_PENALTY_UPDATE_VALUE);
    CHECK_EQ(o.setMaxIterations(0), INVALID_MAX_ITERATIONS_VALUE);
    CHECK_EQ(o.setMaxPenaltyParameter(0.0), INVALID_MAX_RHO_VALUE);
    CHECK_EQ(o.setEtaDynamicPenalty(1.0), INVALID_ETA_VALUE);
    CHECK_EQ(o.setPrintLevel(3), INVALID_PRINT_LEVEL_VALUE);
    CHECK_EQ(o.setQPSolver(4), INVALID_QPSOLVER);
    CHECK_EQ(o.setQPSolver(3), SUCCESSFUL_RETURN);
    CHECK_NEAR(o.getStationarityTolerance(), 1e6 * Utilities::EPS, 1e-30);   // rejected values leave the option alone
    TEST("Options.copy");   // RunUnitTests.cpp:249-262
    Options a;
    a.setQPSolver(OSQP_SPARSE);
    a.getOSQPOptions()->max_iter = 777;
    Options b(a);
    CHECK_EQ(b.getQPSolver(), OSQP_SPARSE);
    CHECK_EQ(b.getOSQPOptions()->max_iter, 777);
    CHECK(b.getOSQPOptions() != a.getOSQPOptions());   // deep copy (Options.cpp:275-287)
    b.setQPSolver(QPOASES_DENSE);
    CHECK_EQ(a.getQPSolver(), OSQP_SPARSE);
    lcqp_cuda_options co;
    a.toCuda(co);
    CHECK_EQ(co.qpSolver, 2);
    CHECK_EQ(co.qp_max_iter, 777);

    TEST("OutputStatistics");   // OutputStatistics.cpp:81-164
    OutputStatistics st;
    CHECK_EQ(st.updateIterTotal(-1), INVALID_TOTAL_ITER_COUNT);
    CHECK_EQ(st.updateIterOuter(-1), INVALID_TOTAL_OUTER_ITER);
    CHECK_EQ(st.updateSubproblemIter(-1), IVALID_SUBPROBLEM_ITER);
    CHECK_EQ(st.updateRhoOpt(0.0), INVALID_RHO_OPT);
    st.updateIterTotal(3); st.updateIterOuter(2); st.updateSubproblemIter(11); st.updateRhoOpt(0.5);
    double xs[2] = {1, 2};
    st.updateTrackingVectors(xs, 0, 4, 1.0, 0.1, 0.2, 0.3, 0.4, 0.5, 2);
    st.updateTrackingVectors(xs, 1, 5, 1.0, 0.1, 0.2, 0.3, 0.4, 0.5, 2);
    CHECK_EQ(st.getIterTotal(), 3); CHECK_EQ(st.getIterOuter(), 2); CHECK_EQ(st.getSubproblemIter(), 11);
    CHECK_EQ(st.getAccuSubproblemItersStdVec().back(), 9);
    CHECK_EQ(st.getxStepsStdVec().size(), (size_t)2);
    OutputStatistics cp;
    cp = st;
    st.reset();
    CHECK_EQ(st.getIterTotal(), 0); CHECK_EQ(cp.getIterTotal(), 3); CHECK_EQ(cp.getInnerItersStdVec().size(), (size_t)2);
}

static void testLoadErrors(const std::string& tmpdir)
{
    TEST("LCQProblem.load errors");
    std::printf("(the ERROR lines below are expected)\n");
    const examples::Problem p = examples::warmUp();
    LCQProblem bad(0, 0, 1);   // LCQProblem.cpp:46-56
    CHECK_EQ(bad.loadLCQP(p.Q.data(), p.g.data(), p.L.data(), p.R.data()), LCQPOBJECT_NOT_SETUP);
    LCQProblem q(2, 0, 1);
    CHECK_EQ(q.loadLCQP(p.Q.data(), (const double*)nullptr, p.L.data(), p.R.data()), INVALID_OBJECTIVE_LINEAR_TERM);
    CHECK_EQ(q.loadLCQP(p.Q.data(), p.g.data(), (const double*)nullptr, p.R.data()), INVALID_COMPLEMENTARITY_MATRIX);
    const double ninf[1] = {-std::numeric_limits<double>::infinity()};
    CHECK_EQ(q.loadLCQP(p.Q.data(), p.g.data(), p.L.data(), p.R.data(), ninf), INVALID_LOWER_COMPLEMENTARITY_BOUND);   // :747
    LCQProblem withA(2, 1, 1);
    CHECK_EQ(withA.loadLCQP(p.Q.data(), p.g.data(), p.L.data(), p.R.data()), INVALID_CONSTRAINT_MATRIX);
    CHECK_EQ(q.runSolver(), LCQPOBJECT_NOT_SETUP);
    CHECK_EQ(q.loadLCQP(p.Q.data(), p.g.data(), p.L.data(), p.R.data(), 0, 0, 0, 0, 0, 0, 0, 0, 0, p.x0.data(), p.y0.data()),
             SUCCESSFUL_RETURN);
    TEST("LCQProblem.sparse mode");   // :1037-1102
    CHECK_EQ(q.switchToSparseMode(), SUCCESSFUL_RETURN);
    CHECK(q.isSparseMode() && q.getSparseQ() && q.getSparseA() && q.getSparseC());
    CHECK_EQ(q.getSparseA()->m, 2);
    CHECK_EQ(q.getSparseC()->p[2], 2);   // C = L'R + R'L = [[0,1],[1,0]]
    CHECK_EQ(q.switchToDenseMode(), SUCCESSFUL_RETURN);
    CHECK(!q.isSparseMode() && q.getSparseQ() == nullptr);

    TEST("LCQProblem.file loader");   // Utilities.cpp:341-395, LCQProblem.cpp:147-306
    auto wr = [&](const char* name, const std::vector<double>& v) {
        const std::string f = tmpdir + "/" + name + ".txt";
        CHECK_EQ(Utilities::writeToFile(v.data(), (int)v.size(), f.c_str()), SUCCESSFUL_RETURN);
        return f;
    };
    const std::string fQ = wr("Q", p.Q), fg = wr("g", p.g), fL = wr("L", p.L), fR = wr("R", p.R), fx = wr("x0", p.x0);
    LCQProblem ff(2, 0, 1);
    CHECK_EQ(ff.loadLCQP(fQ.c_str(), fg.c_str(), fL.c_str(), fR.c_str(), 0, 0, 0, 0, 0, 0, 0, 0, 0, fx.c_str()), SUCCESSFUL_RETURN);
    CHECK_EQ(ff.loadLCQP((tmpdir + "/missing.txt").c_str(), fg.c_str(), fL.c_str(), fR.c_str()), UNABLE_TO_READ_FILE);
}

static void testNoDeviceIsLoud()
{
    TEST("no device is loud");
    const examples::Problem p = examples::warmUp();
    LCQProblem q(2, 0, 1);
    Options o;
    o.setPrintLevel(NONE);
    q.setOptions(o);
    q.loadLCQP(p.Q.data(), p.g.data(), p.L.data(), p.R.data(), 0, 0, 0, 0, 0, 0, 0, 0, 0, p.x0.data());
    std::printf("(the ERROR line below is expected)\n");
    CHECK_EQ(q.runSolver(), SUBPROBLEM_SOLVER_ERROR);   // no CPU fallback
    OutputStatistics st;
    q.getOutputStatistics(st);
    CHECK_EQ(st.getQPSolverExitFlag(), LCQP_CUDA_NO_DEVICE);
    SubsolverCUDA sub(2, 2, p.Q.data(), p.Q.data());
    CHECK(!sub.isValid());
    int it = 0, fl = 0;
    const double lo[2] = {0, 0}, up[2] = {1, 1};
    CHECK_EQ(sub.solve(true, it, fl, p.g.data(), lo, up), SUBPROBLEM_SOLVER_ERROR);
    CHECK_EQ(fl, LCQP_CUDA_NO_DEVICE);
    LCQProblemBatch batch(2, 0, 1, 4);
    CHECK(!batch.isValid());
}

// ---- GPU tests ---------------------------------------------------------------------------------------------
static ReturnValue solveWarmUp(LCQProblem& lcqp, const Options& o, bool withGuess = true, double g1 = -2.0)
{
    examples::Problem p = examples::warmUp();
    p.g[1] = g1;
    lcqp.setOptions(o);
    ReturnValue r = lcqp.loadLCQP(p.Q.data(), p.g.data(), p.L.data(), p.R.data(), 0, 0, 0, 0, 0, 0, 0, 0, 0,
                                  withGuess ? p.x0.data() : nullptr, withGuess ? p.y0.data() : nullptr);
    if (r != SUCCESSFUL_RETURN) return r;
    return lcqp.runSolver();
}

static void testRunWarmUp()
{
    TEST("SolverTest.RunWarmUp");   // RunUnitTests.cpp:505-551: 100 reloads, x in {(1,0),(0,1)}, dual stationarity
    for (int rep = 0; rep < 100; ++rep) {
        LCQProblem lcqp(2, 0, 1);
        Options o;
        o.setPrintLevel(NONE);
        o.setPerturbSeed((unsigned long long)rep + 1);   // the reference reseeds from time(); here: one seed per repetition
        CHECK_EQ(solveWarmUp(lcqp, o), SUCCESSFUL_RETURN);
        double x[2], y[4];
        CHECK_EQ(lcqp.getPrimalSolution(x), S_STATIONARY_SOLUTION);
        lcqp.getDualSolution(y);
        const double tol = o.getStationarityTolerance();
        const bool sol1 = std::fabs(x[0] - 1) <= tol && std::fabs(x[1]) <= tol;
        const bool sol2 = std::fabs(x[1] - 1) <= tol && std::fabs(x[0]) <= tol;
        CHECK(sol1 || sol2);
        for (int i = 0; i < 2; ++i) CHECK(std::fabs(2 * x[i] - 2 - y[i] - y[2 + i]) <= tol);
        OutputStatistics st;
        lcqp.getOutputStatistics(st);
        CHECK(st.getIterOuter() == 15 || st.getIterOuter() == 16);   // k = 15 in every reference run
        CHECK_EQ(lcqp.getNumberOfDuals(), 4);
    }
}

static void testHostLoopMatchesDeviceLoop()
{
    TEST("host loop (plugin door) == device loop");
    Options dev, host;
    dev.setPrintLevel(NONE);
    dev.setPerturbStep(false);
    host = dev;
    host.setStoreSteps(true);   // forces the host loop over SubsolverCUDA
    // warm_up with g = (-2, -3): the shipped g = (-2, -2) is exactly symmetric in (x1, x2) and, without the
    // perturbation, follows a saddle path that any difference in round-off leaves (tests/conftest.py,
    // SYMMETRIC_SADDLE) -- the two routes need not leave it in the same pass.  The asymmetric problem has one
    // trajectory: the unmodified reference (perturbStep off) ends S-stationary at (0, 1.5) with k = 8, i = 39.
    LCQProblem a(2, 0, 1), b(2, 0, 1);
    CHECK_EQ(solveWarmUp(a, dev, true, -3.0), SUCCESSFUL_RETURN);
    CHECK_EQ(solveWarmUp(b, host, true, -3.0), SUCCESSFUL_RETURN);
    OutputStatistics sa, sb;
    a.getOutputStatistics(sa);
    b.getOutputStatistics(sb);
    CHECK_EQ(sa.getIterOuter(), 8);
    CHECK_EQ(sa.getIterTotal(), 39);
    CHECK_EQ(sb.getIterOuter(), sa.getIterOuter());
    CHECK_EQ(sb.getIterTotal(), sa.getIterTotal());
    CHECK_EQ(sb.getRhoOpt(), sa.getRhoOpt());
    CHECK_EQ((int)sb.getSolutionStatus(), (int)sa.getSolutionStatus());
    CHECK_EQ(sb.getxStepsStdVec().size(), (size_t)sb.getIterTotal());
    double xa[2], xb[2];
    a.getPrimalSolution(xa);
    b.getPrimalSolution(xb);
    CHECK_NEAR(xa[0], xb[0], 1e-9);
    CHECK_NEAR(xa[1], xb[1], 1e-9);

    // the circle example through both routes, OSQP-style dual layout, sparse mode
    const examples::Problem p = examples::circle(100, 0.5, -0.6);
    for (int route = 0; route < 2; ++route) {
        LCQProblem c(p.nV, p.nC, p.nComp);
        Options o;
        o.setPrintLevel(NONE);
        o.setStoreSteps(route == 1);
        o.setQPSolver(OSQP_SPARSE);
        o.setStationarityTolerance(10e-3);
        o.setPerturbStep(false);
        c.setOptions(o);
        CHECK_EQ(c.loadLCQP(p.Q.data(), p.g.data(), p.L.data(), p.R.data(), 0, 0, 0, 0, p.A.data(), p.lbA.data(), p.ubA.data(), 0,
                            0, p.x0.data()),
                 SUCCESSFUL_RETURN);
        CHECK_EQ(c.switchToSparseMode(), SUCCESSFUL_RETURN);
        CHECK_EQ(c.runSolver(), SUCCESSFUL_RETURN);
        std::vector<double> x(p.nV);
        CHECK_EQ(c.getPrimalSolution(x.data()), S_STATIONARY_SOLUTION);
        CHECK_NEAR(x[0], 0.181110968, 1e-6);   // the known global solution (test/examples/OptimizeOnCircle.cpp:143)
        CHECK_NEAR(x[1], -0.983483383, 1e-6);
        OutputStatistics st;
        c.getOutputStatistics(st);
        CHECK_EQ(st.getIterOuter(), 8);   // the reference's exact-QP (qpOASES) run: k = 8, i = 19, rho = 2.56
        CHECK_EQ(st.getIterTotal(), 19);
        CHECK_NEAR(st.getRhoOpt(), 2.56, 1e-12);
        CHECK_EQ(c.getNumberOfDuals(), p.nC + 2 * p.nComp);
    }
}

static void testReturnFlags()
{
    TEST("OutputStatisticsTest.CheckQPReturnFlag");   // RunUnitTests.cpp:463-502: lbA = 0 > ubA = -1
    {
        const double Q[4] = {2, 0, 0, 2}, g[2] = {-2, -2}, L[2] = {1, 0}, R[2] = {0, 1}, A[2] = {1, 1}, lbA[1] = {0}, ubA[1] = {-1};
        for (int route = 0; route < 2; ++route) {
            LCQProblem lcqp(2, 1, 1);
            Options o;
            o.setPrintLevel(NONE);
            o.setStoreSteps(route == 1);
            lcqp.setOptions(o);
            CHECK_EQ(lcqp.loadLCQP(Q, g, L, R, 0, 0, 0, 0, A, lbA, ubA), SUCCESSFUL_RETURN);
            std::printf("(the ERROR line below is expected)\n");
            CHECK_EQ(lcqp.runSolver(), SUBPROBLEM_SOLVER_ERROR);
            OutputStatistics st;
            lcqp.getOutputStatistics(st);
            CHECK(st.getQPSolverExitFlag() != 0);
        }
    }
    TEST("test_max_penalty");   // test/examples/test_max_penalty.cpp:75-79
    {
        LCQProblem lcqp(2, 0, 1);
        Options o;
        o.setPrintLevel(NONE);
        o.setMaxPenaltyParameter(1.0);
        std::printf("(the ERROR line below is expected)\n");
        CHECK_EQ(solveWarmUp(lcqp, o), MAX_PENALTY_REACHED);
    }
    TEST("LoadDataTest.DenseToSparse");   // RunUnitTests.cpp:413-460
    {
        LCQProblem lcqp(2, 0, 1);
        Options o;
        o.setPrintLevel(NONE);
        o.setQPSolver(QPOASES_SPARSE);
        const examples::Problem p = examples::warmUp();
        lcqp.setOptions(o);
        CHECK_EQ(lcqp.loadLCQP(p.Q.data(), p.g.data(), p.L.data(), p.R.data(), 0, 0, 0, 0, 0, 0, 0, 0, 0, p.x0.data(), p.y0.data()),
                 SUCCESSFUL_RETURN);
        CHECK_EQ(lcqp.switchToSparseMode(), SUCCESSFUL_RETURN);
        CHECK_EQ(lcqp.runSolver(), SUCCESSFUL_RETURN);
        o.setQPSolver(QPOASES_DENSE);
        lcqp.setOptions(o);
        CHECK_EQ(lcqp.switchToDenseMode(), SUCCESSFUL_RETURN);
        CHECK_EQ(lcqp.runSolver(), SUCCESSFUL_RETURN);
    }
    TEST("OSQP layout rejects box bounds");   // LCQProblem.cpp:930-932, :955-957
    {
        const examples::Problem p = examples::warmUp();
        const double lb[2] = {0, 0}, ub[2] = {2, 2};
        LCQProblem lcqp(2, 0, 1);
        Options o;
        o.setPrintLevel(NONE);
        o.setQPSolver(OSQP_SPARSE);
        lcqp.setOptions(o);
        CHECK_EQ(lcqp.loadLCQP(p.Q.data(), p.g.data(), p.L.data(), p.R.data(), 0, 0, 0, 0, 0, 0, 0, lb, ub, p.x0.data()),
                 SUCCESSFUL_RETURN);
        std::printf("(the ERROR line below is expected)\n");
        CHECK_EQ(lcqp.runSolver(), INVALID_OSQP_BOX_CONSTRAINTS);
    }
}

static void testCscLoader()
{
    TEST("loadLCQP(csc)");   // LCQProblem.cpp:309-387, test/examples/warm_up_sparse.cpp
    const examples::Problem p = examples::warmUp();
    csc* Q = Utilities::dns_to_csc(p.Q.data(), 2, 2);
    csc* L = Utilities::dns_to_csc(p.L.data(), 1, 2);
    csc* R = Utilities::dns_to_csc(p.R.data(), 1, 2);
    LCQProblem lcqp(2, 0, 1);
    Options o;
    o.setPrintLevel(NONE);
    o.setQPSolver(OSQP_SPARSE);
    lcqp.setOptions(o);
    CHECK_EQ(lcqp.loadLCQP(Q, p.g.data(), L, R, 0, 0, 0, 0, 0, 0, 0, 0, 0, p.x0.data()), SUCCESSFUL_RETURN);
    CHECK(lcqp.isSparseMode());
    CHECK_EQ(lcqp.runSolver(), SUCCESSFUL_RETURN);
    double x[2];
    CHECK_EQ(lcqp.getPrimalSolution(x), S_STATIONARY_SOLUTION);
    CHECK((std::fabs(x[0] - 1) < 1e-6 && std::fabs(x[1]) < 1e-6) || (std::fabs(x[1] - 1) < 1e-6 && std::fabs(x[0]) < 1e-6));
    CHECK_EQ(lcqp.getNumberOfDuals(), 2);
    Utilities::ClearSparseMat(&Q);
    Utilities::ClearSparseMat(&L);
    Utilities::ClearSparseMat(&R);
}

static void testBatchAndSharding(int ngpus)
{
    TEST("LCQProblemBatch + sharding");
    const int B = 37;
    const examples::Problem p = examples::circle(20, 0.5, -0.6);   // N = 20 facets: nV = 42
    const int n = p.nV;
    std::vector<double> g((size_t)B * n, 0.0), x0((size_t)B * n, 1.0);
    for (int b = 0; b < B; ++b) {
        const double ang = 0.37 * b, rad = 0.2 + 0.7 * (b % 7) / 7.0;
        const double xr = b == 0 ? 0.5 : rad * std::cos(ang), yr = b == 0 ? -0.6 : rad * std::sin(ang);
        g[(size_t)b * n] = -(17 * xr - 15 * yr);
        g[(size_t)b * n + 1] = -(-15 * xr + 17 * yr);
        x0[(size_t)b * n] = xr;
        x0[(size_t)b * n + 1] = yr;
    }
    const unsigned shared = (1u << LCQP_Q) | (1u << LCQP_L) | (1u << LCQP_R) | (1u << LCQP_A) | (1u << LCQP_LBA) | (1u << LCQP_UBA);
    Options o;
    o.setPrintLevel(NONE);
    o.setStationarityTolerance(10e-3);
    std::vector<std::vector<double>> xs;
    std::vector<std::vector<int>> iters;
    // one shard, then two shards (on two GPUs when there are two, else twice the same GPU): identical results --
    // the perturbStep draws are keyed by the GLOBAL instance index
    for (int shards = 1; shards <= 2; ++shards) {
        std::vector<int> devs;
        for (int s = 0; s < shards; ++s) devs.push_back(ngpus > 1 ? s : 0);
        LCQProblemBatch batch(p.nV, p.nC, p.nComp, B, devs);
        CHECK(batch.isValid());
        CHECK_EQ(batch.getNumberOfShards(), shards);
        CHECK_EQ(batch.setOptions(o), SUCCESSFUL_RETURN);
        CHECK_EQ(batch.loadLCQP(shared, p.Q.data(), g.data(), p.L.data(), p.R.data(), 0, 0, 0, 0, p.A.data(), p.lbA.data(),
                                p.ubA.data(), 0, 0, x0.data()),
                 SUCCESSFUL_RETURN);
        CHECK_EQ(batch.runSolver(), SUCCESSFUL_RETURN);
        std::vector<double> x((size_t)B * n);
        CHECK_EQ(batch.getPrimalSolution(x.data()), SUCCESSFUL_RETURN);
        std::vector<OutputStatistics> st;
        CHECK_EQ(batch.getOutputStatistics(st), SUCCESSFUL_RETURN);
        std::vector<int> it;
        int solved = 0;
        const std::vector<int> rv = batch.getReturnValues();
        for (int b = 0; b < B; ++b) {
            it.push_back(st[(size_t)b].getIterTotal());
            solved += rv[(size_t)b] == SUCCESSFUL_RETURN;
            if (rv[(size_t)b] != SUCCESSFUL_RETURN) continue;
            // complementarity and feasibility of every solved instance
            double phi = 0;
            for (int i = 0; i < p.nComp; ++i) phi += x[(size_t)b * n + 2 + 2 * i] * x[(size_t)b * n + 3 + 2 * i];
            CHECK(phi < 1e-9);
        }
        CHECK(solved >= B - 2);
        CHECK(batch.getLaunchCount() >= 2 * shards);
        xs.push_back(x);
        iters.push_back(it);
    }
    CHECK(xs[0] == xs[1]);        // bit-identical: sharding does not change any instance's arithmetic
    CHECK(iters[0] == iters[1]);
}

static void testPluginDoor()
{
    TEST("SubsolverCUDA plugin door");   // SubsolverBase::solve semantics: initial solve, then hot starts with new g
    const int n = 3, m = 2;
    const double Q[9] = {2, 0.5, 0, 0.5, 1, 0, 0, 0, 3}, A[6] = {1, 1, 1, 1, -1, 0};
    const double lbA[2] = {1, -0.5}, ubA[2] = {1, 0.5};
    Subsolver sub(n, m, Q, A);
    CHECK(sub.isValid());
    Subsolver copy = sub;   // copies are fresh solvers on the same data (Subsolver.cpp:125-136)
    CHECK(copy.isValid());
    const double gs[2][3] = {{-1, -1, -1}, {3, -2, 0.5}};
    for (int k = 0; k < 2; ++k) {
        int it = -1, fl = -1;
        const double x0[3] = {0, 0, 0};
        CHECK_EQ(copy.solve(k == 0, it, fl, gs[k], lbA, ubA, x0), SUCCESSFUL_RETURN);
        CHECK_EQ(fl, 0);
        double x[3], y[5];
        copy.getSolution(x, y);
        // KKT: Qx + g = A'y_A (qpOASES sign), feasibility, multiplier signs
        for (int j = 0; j < n; ++j) {
            double r = gs[k][j];
            for (int c = 0; c < n; ++c) r += Q[j * n + c] * x[c];
            for (int i = 0; i < m; ++i) r -= A[i * n + j] * y[n + i];
            CHECK(std::fabs(r) < 1e-9);
        }
        const double a0 = x[0] + x[1] + x[2], a1 = x[0] - x[1];
        CHECK(std::fabs(a0 - 1) < 1e-9);
        CHECK(a1 > -0.5 - 1e-9 && a1 < 0.5 + 1e-9);
        if (a1 > -0.5 + 1e-7) CHECK(y[n + 1] <= 1e-9);
        if (a1 < 0.5 - 1e-7) CHECK(y[n + 1] >= -1e-9);
    }
}

int main(int argc, char** argv)
{
    const std::string mode = argc > 1 ? argv[1] : "--cpu";
    const std::string tmpdir = argc > 2 ? argv[2] : "/tmp";
    if (mode == "--cpu") {
        testUtilities();
        testOptionsAndStatistics();
        testLoadErrors(tmpdir);
        if (argc > 3 && std::string(argv[3]) == "--expect-no-device") testNoDeviceIsLoud();
    } else if (mode == "--gpu") {
        const int ngpus = argc > 3 ? std::atoi(argv[3]) : 1;
        testRunWarmUp();
        testHostLoopMatchesDeviceLoop();
        testReturnFlags();
        testCscLoader();
        testBatchAndSharding(ngpus);
        testPluginDoor();
    } else {
        std::printf("usage: host_tests --cpu|--gpu [tmpdir] [--expect-no-device | ngpus]\n");
        return 2;
    }
    std::printf("%s: %d checks, %d failed\n", mode.c_str(), g_checks, g_fail);
    return g_fail ? 1 : 0;
}

// LCQProblem.cpp -- see ../include/LCQProblem.hpp.  Behaviour follows /root/reference/src/LCQProblem.cpp
// (line numbers cited per function); the code is new.
#include "LCQProblem.hpp"

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>

namespace LCQPow {

namespace {

ReturnValue err(ReturnValue r) { return MessageHandler::PrintMessage(r, ERROR); }

void assign(std::vector<double>& dst, bool& has, const double* src, size_t n)
{
    has = (src != nullptr);
    if (has) dst.assign(src, src + n);
    else dst.clear();
}

const double* ptr(const std::vector<double>& v, bool has) { return has ? v.data() : nullptr; }

// the perturbStep generator of the device loop (lcqp_device.cuh perturb_draw): splitmix64 finaliser keyed by
// (seed, instance, iterate, coordinate) -> {-1, 0, +1}.  The reference draws rand() % 3 - 1 after
// srand(time(NULL)) (LCQProblem.cpp:1016, 1353-1362).
int perturbDraw(unsigned long long seed, unsigned long long instance, unsigned iter, unsigned i)
{
    uint64_t z = seed * 0x9E3779B97F4A7C15ull + instance * 0xBF58476D1CE4E5B9ull + (((uint64_t)iter << 32) | i);
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    return (int)(z % 3ull) - 1;
}

}  // namespace

LCQProblem::LCQProblem() {}

// reference: LCQProblem.cpp:43-79 (invalid sizes print a message and leave the object unusable)
LCQProblem::LCQProblem(int _nV, int _nC, int _nComp)
{
    if (_nV <= 0 || _nComp <= 0) {
        err(INVALID_NUMBER_OF_OPTIM_VARS);
        return;
    }
    if (_nC < 0) {
        err(INVALID_NUMBER_OF_CONSTRAINT_VARS);
        return;
    }
    nV = _nV;
    nC = _nC;
    nComp = _nComp;
}

LCQProblem::~LCQProblem() { clear(); }

void LCQProblem::clear()
{
    Utilities::ClearSparseMat(&Q_sparse);
    Utilities::ClearSparseMat(&A_sparse);
    Utilities::ClearSparseMat(&L_sparse);
    Utilities::ClearSparseMat(&R_sparse);
    Utilities::ClearSparseMat(&C_sparse);
}

void LCQProblem::setOptions(const Options& _options) { options = _options; }

// A_full = [A; L; R] and its bounds, C = L'R + R'L (reference: setConstraints :563-626)
ReturnValue LCQProblem::setConstraints(const double* L_new, const double* R_new, const double* A_new, const double* lbA_new,
                                       const double* ubA_new)
{
    if (nV == 0 || nComp == 0) return LCQPOBJECT_NOT_SETUP;
    if (!A_new && nC > 0) return INVALID_CONSTRAINT_MATRIX;
    if (!L_new || !R_new) return INVALID_COMPLEMENTARITY_MATRIX;
    const int m = nC + 2 * nComp;
    const size_t n = (size_t)nV;
    L.assign(L_new, L_new + (size_t)nComp * n);
    R.assign(R_new, R_new + (size_t)nComp * n);
    if (nC > 0) A.assign(A_new, A_new + (size_t)nC * n);
    else A.clear();
    Afull.resize((size_t)m * n);
    if (nC > 0) std::memcpy(Afull.data(), A.data(), sizeof(double) * (size_t)nC * n);
    std::memcpy(Afull.data() + (size_t)nC * n, L.data(), sizeof(double) * (size_t)nComp * n);
    std::memcpy(Afull.data() + (size_t)(nC + nComp) * n, R.data(), sizeof(double) * (size_t)nComp * n);
    assign(lbAuser, has_lbA, lbA_new, (size_t)nC);
    assign(ubAuser, has_ubA, ubA_new, (size_t)nC);
    const double inf = std::numeric_limits<double>::infinity();
    lbA.assign((size_t)m, -inf);
    ubA.assign((size_t)m, inf);
    for (int i = 0; i < nC; ++i) {
        if (has_lbA) lbA[i] = lbAuser[i];
        if (has_ubA) ubA[i] = ubAuser[i];
    }
    C.assign(n * n, 0.0);
    Utilities::MatrixSymmetrizationProduct(L.data(), R.data(), C.data(), nComp, nV);
    return SUCCESSFUL_RETURN;
}

// rows nC .. nC + 2 nComp of lbA / ubA (reference: setComplementarityBounds :726-796)
ReturnValue LCQProblem::setComplementarityBounds(const double* lbL_new, const double* ubL_new, const double* lbR_new,
                                                 const double* ubR_new)
{
    const double inf = std::numeric_limits<double>::infinity();
    assign(lbL, has_lbL, lbL_new, (size_t)nComp);
    assign(ubL, has_ubL, ubL_new, (size_t)nComp);
    assign(lbR, has_lbR, lbR_new, (size_t)nComp);
    assign(ubR, has_ubR, ubR_new, (size_t)nComp);
    for (int i = 0; i < nComp; ++i) {
        if (has_lbL && lbL[i] <= -inf) return INVALID_LOWER_COMPLEMENTARITY_BOUND;
        if (has_lbR && lbR[i] <= -inf) return INVALID_LOWER_COMPLEMENTARITY_BOUND;
        lbA[nC + i] = has_lbL ? lbL[i] : 0.0;
        ubA[nC + i] = has_ubL ? ubL[i] : inf;
        lbA[nC + nComp + i] = has_lbR ? lbR[i] : 0.0;
        ubA[nC + nComp + i] = has_ubR ? ubR[i] : inf;
    }
    return SUCCESSFUL_RETURN;
}

ReturnValue LCQProblem::loadLCQP(const double* const _Q, const double* const _g, const double* const _L, const double* const _R,
                                 const double* const _lbL, const double* const _ubL, const double* const _lbR,
                                 const double* const _ubR, const double* const _A, const double* const _lbA,
                                 const double* const _ubA, const double* const _lb, const double* const _ub,
                                 const double* const _x0, const double* const _y0)
{
    if (nV <= 0 || nComp <= 0) return err(LCQPOBJECT_NOT_SETUP);
    if (!_Q) return err(INVALID_ARGUMENT);
    if (!_g) return err(INVALID_OBJECTIVE_LINEAR_TERM);
    loaded = false;
    clear();
    sparseSolver = false;
    const size_t n = (size_t)nV;
    Q.assign(_Q, _Q + n * n);
    g.assign(_g, _g + n);
    assign(lb, has_lb, _lb, n);
    assign(ub, has_ub, _ub, n);
    ReturnValue ret = setConstraints(_L, _R, _A, _lbA, _ubA);
    if (ret != SUCCESSFUL_RETURN) return err(ret);
    ret = setComplementarityBounds(_lbL, _ubL, _lbR, _ubR);
    if (ret != SUCCESSFUL_RETURN) return err(ret);
    assign(x0, has_x0, _x0, n);                                        // zero vector if absent (LCQProblem.ipp:133-142)
    assign(y0, has_y0, _y0, n + (size_t)nC + 2 * (size_t)nComp);    // box(nV) + A(nC) + L + R (LCQProblem.ipp:144-151)
    xk.assign(n, 0.0);
    if (has_x0) xk = x0;
    yk.clear();
    algoStat = PROBLEM_NOT_SOLVED;
    stats.reset();
    loaded = true;
    return SUCCESSFUL_RETURN;
}

ReturnValue LCQProblem::loadLCQP(const char* const Q_file, const char* const g_file, const char* const L_file,
                                 const char* const R_file, const char* const lbL_file, const char* const ubL_file,
                                 const char* const lbR_file, const char* const ubR_file, const char* const A_file,
                                 const char* const lbA_file, const char* const ubA_file, const char* const lb_file,
                                 const char* const ub_file, const char* const x0_file, const char* const y0_file)
{
    if (nV <= 0 || nComp <= 0) return err(LCQPOBJECT_NOT_SETUP);
    const size_t n = (size_t)nV;
    struct Item { const char* file; size_t len; bool required; std::vector<double> data; bool present; };
    Item items[15] = {
        {Q_file, n * n, true, {}, false},          {g_file, n, true, {}, false},
        {L_file, (size_t)nComp * n, true, {}, false}, {R_file, (size_t)nComp * n, true, {}, false},
        {lbL_file, (size_t)nComp, false, {}, false}, {ubL_file, (size_t)nComp, false, {}, false},
        {lbR_file, (size_t)nComp, false, {}, false}, {ubR_file, (size_t)nComp, false, {}, false},
        {A_file, (size_t)nC * n, false, {}, false},  {lbA_file, (size_t)nC, false, {}, false},
        {ubA_file, (size_t)nC, false, {}, false},    {lb_file, n, false, {}, false},
        {ub_file, n, false, {}, false},              {x0_file, n, false, {}, false},
        {y0_file, n + (size_t)nC + 2 * (size_t)nComp, false, {}, false}};
    for (Item& it : items) {
        if (!it.file) {
            if (it.required) return err(UNABLE_TO_READ_FILE);
            continue;
        }
        it.data.resize(it.len);
        const ReturnValue ret = Utilities::readFromFile(it.data.data(), (int)it.len, it.file);
        if (ret != SUCCESSFUL_RETURN) return err(ret);
        it.present = true;
    }
    auto p = [&](int k) -> const double* { return items[k].present ? items[k].data.data() : nullptr; };
    return loadLCQP(p(0), p(1), p(2), p(3), p(4), p(5), p(6), p(7), p(8), p(9), p(10), p(11), p(12), p(13), p(14));
}

ReturnValue LCQProblem::loadLCQP(const csc* const _Q, const double* const _g, const csc* const _L, const csc* const _R,
                                 const double* const _lbL, const double* const _ubL, const double* const _lbR,
                                 const double* const _ubR, const csc* const _A, const double* const _lbA,
                                 const double* const _ubA, const double* const _lb, const double* const _ub,
                                 const double* const _x0, const double* const _y0)
{
    if (!_Q) return err(INVALID_ARGUMENT);
    if (!_L || !_R) return err(INVALID_COMPLEMENTARITY_MATRIX);
    if (!_A && nC > 0) return err(INVALID_CONSTRAINT_MATRIX);
    if (_Q->m != nV || _Q->n != nV || _L->m != nComp || _L->n != nV || _R->m != nComp || _R->n != nV ||
        (_A && (_A->m != nC || _A->n != nV)))
        return err(DENSE_SPARSE_MISSMATCH);
    double* Qd = Utilities::csc_to_dns(_Q);
    double* Ld = Utilities::csc_to_dns(_L);
    double* Rd = Utilities::csc_to_dns(_R);
    double* Ad = (_A && nC > 0) ? Utilities::csc_to_dns(_A) : nullptr;
    ReturnValue ret = FAILED_SWITCH_TO_DENSE;
    if (Qd && Ld && Rd && (Ad || nC == 0)) {
        ret = loadLCQP(Qd, _g, Ld, Rd, _lbL, _ubL, _lbR, _ubR, Ad, _lbA, _ubA, _lb, _ub, _x0, _y0);
        if (ret == SUCCESSFUL_RETURN) ret = switchToSparseMode();
    }
    delete[] Qd;
    delete[] Ld;
    delete[] Rd;
    delete[] Ad;
    return ret;
}

// The device consumes dense fp64 operands and builds its own CSR copies of whatever is sparse, so the dense
// master copies stay; sparse mode additionally holds the csc views the reference exposes internally.
ReturnValue LCQProblem::switchToSparseMode()
{
    if (sparseSolver) return SUCCESSFUL_RETURN;
    if (Q.empty()) return FAILED_SWITCH_TO_SPARSE;
    clear();
    Q_sparse = Utilities::dns_to_csc(Q.data(), nV, nV);
    A_sparse = Utilities::dns_to_csc(Afull.data(), nC + 2 * nComp, nV);
    L_sparse = Utilities::dns_to_csc(L.data(), nComp, nV);
    R_sparse = Utilities::dns_to_csc(R.data(), nComp, nV);
    C_sparse = Utilities::dns_to_csc(C.data(), nV, nV);
    if (!Q_sparse || !A_sparse || !L_sparse || !R_sparse || !C_sparse) {
        clear();
        return FAILED_SWITCH_TO_SPARSE;
    }
    sparseSolver = true;
    return SUCCESSFUL_RETURN;
}

ReturnValue LCQProblem::switchToDenseMode()
{
    if (!sparseSolver) return SUCCESSFUL_RETURN;
    if (Q.empty()) return FAILED_SWITCH_TO_DENSE;
    clear();
    sparseSolver = false;
    return SUCCESSFUL_RETURN;
}

ReturnValue LCQProblem::runSolver()
{
    if (!loaded) return err(LCQPOBJECT_NOT_SETUP);
    stats.reset();
    algoStat = PROBLEM_NOT_SOLVED;
    const bool osqpLayout = (options.getQPSolver() == OSQP_SPARSE);
    const int m = nC + 2 * nComp;
    // initializeSolver (reference :885-960): dual layout per subsolver flavour; the OSQP layout has no box rows
    if (osqpLayout && (has_lb || has_ub)) return err(INVALID_OSQP_BOX_CONSTRAINTS);
    nDuals = osqpLayout ? m : nV + m;
    boxDualOffset = osqpLayout ? 0 : nV;
    xk.assign((size_t)nV, 0.0);
    if (has_x0) xk = x0;
    yk.assign((size_t)nDuals, 0.0);
    // (the OSQP restatement exists on the device only: no plug-in door, hence no host loop)
    const bool hostLoop = ((options.getPrintLevel() != NONE) || options.getStoreSteps()) && !(osqpLayout && options.getOSQPADMM());
    const ReturnValue ret = hostLoop ? runHostLoop() : runDeviceLoop();
    if (ret != SUCCESSFUL_RETURN) return err(ret);
    return SUCCESSFUL_RETURN;
}

// ---- route 1: the whole loop on the device -------------------------------------------------------------
ReturnValue LCQProblem::runDeviceLoop()
{
    lcqp_cuda_handle h = nullptr;
    int rc = lcqp_cuda_create(nV, nC, nComp, 1, options.getDevice(), &h);
    if (rc != LCQP_CUDA_OK) {
        deviceError = "lcqp_cuda_create failed";
        stats.updateQPSolverExitFlag(rc);
        return rc >= LCQP_CUDA_NO_DEVICE ? SUBPROBLEM_SOLVER_ERROR : (ReturnValue)rc;
    }
    lcqp_cuda_options o;
    options.toCuda(o);
    auto fin = [&](int code) -> ReturnValue {
        if (code != LCQP_CUDA_OK) deviceError = lcqp_cuda_last_error(h);
        lcqp_cuda_destroy(h);
        if (code >= LCQP_CUDA_NO_DEVICE) {
            stats.updateQPSolverExitFlag(code);
            return SUBPROBLEM_SOLVER_ERROR;
        }
        return (ReturnValue)code;
    };
    if ((rc = lcqp_cuda_set_options(h, &o)) != LCQP_CUDA_OK) return fin(rc);
    const unsigned allShared = (1u << LCQP_NUM_ARRAYS) - 1u;
    rc = lcqp_cuda_load(h, 1, allShared, Q.data(), g.data(), L.data(), R.data(), ptr(lbL, has_lbL), ptr(ubL, has_ubL),
                        ptr(lbR, has_lbR), ptr(ubR, has_ubR), nC > 0 ? A.data() : nullptr, ptr(lbAuser, has_lbA),
                        ptr(ubAuser, has_ubA), ptr(lb, has_lb), ptr(ub, has_ub), ptr(x0, has_x0), ptr(y0, has_y0));
    if (rc != LCQP_CUDA_OK) return fin(rc);
    if ((rc = lcqp_cuda_run(h, nullptr)) != LCQP_CUDA_OK) return fin(rc);
    std::vector<double> yfull((size_t)nV + nC + 2 * (size_t)nComp, 0.0);
    lcqp_cuda_stats rec;
    if ((rc = lcqp_cuda_get_primal(h, xk.data())) != LCQP_CUDA_OK) return fin(rc);
    if ((rc = lcqp_cuda_get_dual(h, yfull.data())) != LCQP_CUDA_OK) return fin(rc);
    if ((rc = lcqp_cuda_get_stats(h, &rec)) != LCQP_CUDA_OK) return fin(rc);
    std::memcpy(yk.data(), yfull.data(), sizeof(double) * (size_t)nDuals);
    stats.fromCuda(rec);
    algoStat = (AlgorithmStatus)rec.status;
    return fin(rec.ret);
}

// ---- route 2: the reference's loop on the host over the SubsolverCUDA plugin -----------------------------
// out = Q v + rho * C v + add  (Qk v + add, reference setQk :799-882 keeps Qk explicitly)
void LCQProblem::applyQk(const double* v, const double* add, double* out)
{
    const size_t n = (size_t)nV;
    for (size_t i = 0; i < n; ++i) {
        double s = 0.0, c = 0.0;
        const double* q = Q.data() + i * n;
        const double* cc = C.data() + i * n;
        for (size_t j = 0; j < n; ++j) {
            s += q[j] * v[j];
            c += cc[j] * v[j];
        }
        out[i] = s + rho * c + (add ? add[i] : 0.0);
    }
}

double LCQProblem::getObj()   // reference :1163-1170
{
    return Utilities::DotProduct(g.data(), xk.data(), nV) + 0.5 * Utilities::QuadraticFormProduct(Q.data(), xk.data(), nV);
}

// phi = phi_const + g_phi'x + (Lx)'(Rx)  -- x'Cx/2 written as the device loop writes it (reference :1172-1185
// evaluates the quadratic form with C; the two agree to rounding)
double LCQProblem::getPhi()
{
    double phi = phi_const;
    const size_t n = (size_t)nV;
    for (int i = 0; i < nComp; ++i) {
        double a = 0.0, b = 0.0;
        for (size_t j = 0; j < n; ++j) {
            a += L[(size_t)i * n + j] * xk[j];
            b += R[(size_t)i * n + j] * xk[j];
        }
        phi += a * b;
    }
    if (!gphi.empty()) phi += Utilities::DotProduct(gphi.data(), xk.data(), nV);
    return phi;
}

ReturnValue LCQProblem::runHostLoop()
{
    const int n = nV, m = nC + 2 * nComp;
    const bool osqpLayout = (options.getQPSolver() == OSQP_SPARSE);
    const bool hasBox = !osqpLayout && (has_lb || has_ub);

    // build the plugin exactly where the reference builds its subsolver (:906-907, :926, :959)
    Subsolver tmp(n, m, Q.data(), Afull.data(), options.getQPSolver(), options.getDevice());
    subsolver = tmp;
    if (!subsolver.isValid()) {
        stats.updateQPSolverExitFlag(LCQP_CUDA_NO_DEVICE);
        return SUBPROBLEM_SOLVER_ERROR;
    }
    subsolver.setOptions(options);

    // g_phi = -(R' lbL + L' lbR), phi_const = lbL' lbR (:966-996)
    gphi.clear();
    phi_const = 0.0;
    if (has_lbL || has_lbR) {
        gphi.assign((size_t)n, 0.0);
        for (int i = 0; i < nComp; ++i) {
            const double a = has_lbL ? lbL[i] : 0.0, b = has_lbR ? lbR[i] : 0.0;
            phi_const += a * b;
            for (int j = 0; j < n; ++j) gphi[j] -= R[(size_t)i * n + j] * a + L[(size_t)i * n + j] * b;
        }
    }
    std::vector<double> gk(n), gtilde(g), pk(n, 0.0), statk(n, 0.0), xnew(n), ysol((size_t)n + m, 0.0), ykA(m, 0.0), tmpv(n);
    std::vector<double> hist;
    double alphak = 1.0;
    rho = options.getInitialPenaltyParameter();
    int outerIter = 0, innerIter = 0, totalIter = 0, qpIterk = 0, exitFlag = 0;

    auto linearize = [&]() {   // gk = rho C xk + g_tilde (:1105-1112)
        Utilities::AffineLinearTransformation(rho, C.data(), xk.data(), gtilde.data(), gk.data(), n, n);
    };
    auto updatePenalty = [&]() {   // :1199-1214
        hist.clear();
        rho *= options.getPenaltyUpdateFactor();
        stats.updateRhoOpt(rho);
        for (int j = 0; j < n; ++j) gtilde[j] = g[j] + (gphi.empty() ? 0.0 : rho * gphi[j]);
    };
    auto solveQP = [&](bool initial) -> ReturnValue {   // :1115-1148
        const double* y0p = (initial && has_y0) ? y0.data() : nullptr;
        std::vector<double> y0plugin;
        if (y0p && osqpLayout) {   // the plugin's y0 has box duals first
            y0plugin.assign((size_t)n + m, 0.0);
            std::memcpy(y0plugin.data() + n, y0.data() + n, sizeof(double) * (size_t)m);
            y0p = y0plugin.data();
        }
        const ReturnValue r = subsolver.solve(initial, qpIterk, exitFlag, gk.data(), lbA.data(), ubA.data(), xk.data(), y0p,
                                              hasBox && has_lb ? lb.data() : nullptr, hasBox && has_ub ? ub.data() : nullptr);
        stats.updateQPSolverExitFlag(exitFlag);
        if (r != SUCCESSFUL_RETURN) return r;
        stats.updateSubproblemIter(qpIterk);
        subsolver.getSolution(xnew.data(), ysol.data());
        std::memcpy(ykA.data(), ysol.data() + n, sizeof(double) * (size_t)m);
        for (int j = 0; j < n; ++j) pk[j] = xnew[j] - xk[j];
        return SUCCESSFUL_RETURN;
    };
    auto writeDuals = [&](bool transform) {
        // yk = [box duals ; yk_A]; on success the penalty part is taken out of the complementarity rows
        // (transformDuals :1381-1409)
        if (!osqpLayout)
            for (int j = 0; j < n; ++j) yk[j] = hasBox ? ysol[j] : 0.0;
        std::vector<double> Lx(nComp, 0.0), Rx(nComp, 0.0);
        if (transform) {
            Utilities::MatrixMultiplication(L.data(), xk.data(), Lx.data(), nComp, n, 1);
            Utilities::MatrixMultiplication(R.data(), xk.data(), Rx.data(), nComp, n, 1);
        }
        for (int i = 0; i < m; ++i) {
            double v = ykA[i];
            if (transform && i >= nC && i < nC + nComp) v -= rho * Rx[i - nC];
            else if (transform && i >= nC + nComp) v -= rho * Lx[i - nC - nComp];
            yk[boxDualOffset + i] = v;
        }
    };

    // first QP (:452-467)
    if (options.getSolveZeroPenaltyFirst()) gk = g;
    else linearize();
    ReturnValue ret = solveQP(true);
    if (ret != SUCCESSFUL_RETURN) return ret;
    stats.updateRhoOpt(rho);

    for (;;) {
        for (int j = 0; j < n; ++j) xk[j] += alphak * pk[j];   // updateStep :1240
        // updateStationarity :1246-1272
        applyQk(xk.data(), gtilde.data(), statk.data());
        for (int i = 0; i < m; ++i) {
            const double yi = ykA[i];
            if (yi == 0.0) continue;
            const double* row = Afull.data() + (size_t)i * n;
            for (int j = 0; j < n; ++j) statk[j] -= row[j] * yi;
        }
        if (hasBox)
            for (int j = 0; j < n; ++j) statk[j] -= ysol[j];
        const double statInf = Utilities::MaxAbs(statk.data(), n);
        const double pNorm = Utilities::MaxAbs(pk.data(), n);
        double phi = getPhi();
        printIteration(outerIter, innerIter, statInf, phi, rho, pNorm, alphak, qpIterk);
        if (options.getStoreSteps()) {   // storeSteps :1365-1378
            const double merit = Utilities::DotProduct(g.data(), xk.data(), n) +
                                 0.5 * (Utilities::QuadraticFormProduct(Q.data(), xk.data(), n) +
                                        rho * Utilities::QuadraticFormProduct(C.data(), xk.data(), n));
            stats.updateTrackingVectors(xk.data(), innerIter, qpIterk, alphak, pNorm, statInf, getObj(), phi, merit, n);
        }
        totalIter++;
        innerIter++;
        stats.updateIterTotal(1);

        // leyfferCheckPositive :1275-1313
        const int nd = options.getNDynamicPenalty();
        if (nd > 0) {
            bool fire = false;
            if ((int)hist.size() < nd) hist.push_back(phi);
            else {
                if (!(phi < options.getComplementarityTolerance())) {
                    fire = true;
                    for (int i = 0; i < nd; ++i)
                        if (phi < options.getEtaDynamicPenalty() * hist[i]) { fire = false; break; }
                }
                hist.erase(hist.begin());
                hist.push_back(phi);
            }
            if (fire) {
                updatePenalty();
                outerIter++;
                innerIter = 0;
                stats.updateIterOuter(1);
            }
        }
        if (statInf < options.getStationarityTolerance()) {   // :511
            if (phi < options.getComplementarityTolerance()) {
                // determineStationarityType :1412-1453 on yk_A, weak set :1456-1482
                std::vector<double> Lx(nComp), Rx(nComp);
                Utilities::MatrixMultiplication(L.data(), xk.data(), Lx.data(), nComp, n, 1);
                Utilities::MatrixMultiplication(R.data(), xk.data(), Rx.data(), nComp, n, 1);
                const double tc = options.getComplementarityTolerance();
                bool sOk = true, mOk = true, weakOnly = false;
                for (int i = 0; i < nComp && !weakOnly; ++i) {
                    if (!(Lx[i] <= tc && Rx[i] <= tc)) continue;
                    const double yl = ykA[nC + i], yr = ykA[nC + nComp + i];
                    const double prod = yl * yr, mn = yl < yr ? yl : yr;
                    if (mn < 0) sOk = false;
                    if (std::fabs(prod) >= tc && mn <= 0) {
                        if (prod <= tc) weakOnly = true;
                        else mOk = false;
                    }
                }
                algoStat = weakOnly ? W_STATIONARY_SOLUTION
                                    : (sOk ? S_STATIONARY_SOLUTION : (mOk ? M_STATIONARY_SOLUTION : C_STATIONARY_SOLUTION));
                stats.updateSolutionStatus(algoStat);
                writeDuals(true);
                if (options.getPrintLevel() != NONE) MessageHandler::PrintSolution(algoStat);
                return SUCCESSFUL_RETURN;
            }
            updatePenalty();
            outerIter++;
            innerIter = 0;
            stats.updateIterOuter(1);
        }
        if (totalIter > options.getMaxIterations()) { writeDuals(false); return MAX_ITERATIONS_REACHED; }   // :537
        if (rho > options.getMaxPenaltyParameter()) { writeDuals(false); return MAX_PENALTY_REACHED; }     // :541

        linearize();                                        // :545
        ret = solveQP(false);                               // :548
        if (ret != SUCCESSFUL_RETURN) { writeDuals(false); return ret; }

        if (options.getPerturbStep())                       // :553-555, :1353-1362
            for (int j = 0; j < n; ++j)
                xk[j] += perturbDraw(options.getPerturbSeed(), 0ull, (unsigned)totalIter, (unsigned)j) * Utilities::EPS;

        // getOptimalStepLength :1217-1237
        applyQk(pk.data(), nullptr, tmpv.data());
        const double qk = Utilities::DotProduct(tmpv.data(), pk.data(), n);
        applyQk(xk.data(), gtilde.data(), tmpv.data());
        const double lk = Utilities::DotProduct(tmpv.data(), pk.data(), n);
        alphak = 1.0;
        if (qk > 0 && lk < 0) alphak = (-lk / qk < 1.0) ? -lk / qk : 1.0;
    }
}

AlgorithmStatus LCQProblem::getPrimalSolution(double* const xOpt) const
{
    if (xOpt && !xk.empty()) std::memcpy(xOpt, xk.data(), sizeof(double) * (size_t)nV);
    return algoStat;
}

AlgorithmStatus LCQProblem::getDualSolution(double* const yOpt) const
{
    if (yOpt && !yk.empty()) std::memcpy(yOpt, yk.data(), sizeof(double) * (size_t)nDuals);
    return algoStat;
}

int LCQProblem::getNumberOfPrimals() const { return nV; }
int LCQProblem::getNumberOfDuals() const { return nDuals; }
void LCQProblem::getOutputStatistics(OutputStatistics& _stats) const { _stats = stats; }

// the iteration table of the reference (:1528-1637): same columns and widths
void LCQProblem::printIteration(int outerIter, int innerIter, double statInf, double phi, double rhoNow, double pNorm,
                                double alphak, int qpIterk)
{
    const PrintLevel pl = options.getPrintLevel();
    if (pl == NONE) return;
    if (pl == OUTER_LOOP_ITERATES && innerIter > 0) return;
    const bool inner = (pl >= INNER_LOOP_ITERATES);
    if ((inner && innerIter % 10 == 0) || (!inner && outerIter % 10 == 0)) printHeader();
    std::printf("%6d", outerIter);
    if (inner) std::printf(" | %6d", innerIter);
    std::printf(" | %10.3g | %10.3g | %10.3g | %10.3g", statInf, phi, rhoNow, pNorm);
    if (inner) std::printf(" | %10.3g | %6d", alphak, qpIterk);
    std::printf(" \n");
}

void LCQProblem::printHeader()
{
    const bool inner = (options.getPrintLevel() >= INNER_LOOP_ITERATES);
    printLine();
    std::printf(" outer");
    if (inner) std::printf(" |  inner");
    std::printf(" |   station  |   complem  |     rho    |   norm p  ");
    if (inner) std::printf(" |    alpha   | sub it");
    std::printf(" \n");
    printLine();
}

void LCQProblem::printLine()
{
    const bool inner = (options.getPrintLevel() >= INNER_LOOP_ITERATES);
    std::printf("------");
    if (inner) std::printf("-+-------");
    for (int k = 0; k < 4; ++k) std::printf("-+-----------");
    if (inner) std::printf("-+------------+-------");
    std::printf("-\n");
}

}  // namespace LCQPow

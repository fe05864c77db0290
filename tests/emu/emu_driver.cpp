// tests/emu/emu_driver.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Compiles the DEVICE code of the product (lcqpow_b200/csrc/lcqp_device.cuh) as plain single-threaded C++
// (-DLCQP_HOST_EMU: one "thread" per CTA, barriers are no-ops) and drives it exactly the way the kernels in
// lcqp_cabi.cu do: batch-level preparation (prepare_shared_kernel), then run_instance per instance
// (lcqp_solve_kernel).  The CPU test-suite (-m "not gpu") uses it to check the kernel's logic against the
// oracle and the golden vectors of the real reference.  It is never built, imported or called by the product.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../lcqpow_b200/csrc/lcqp_device.cuh"

using namespace lcqp;

static size_t field_len(int k, int n, int c, int p)
{
    switch (k) {
        case LCQP_Q: return (size_t)n * n;
        case LCQP_G: return n;
        case LCQP_L: case LCQP_R: return (size_t)p * n;
        case LCQP_LBL: case LCQP_UBL: case LCQP_LBR: case LCQP_UBR: return p;
        case LCQP_A: return (size_t)c * n;
        case LCQP_LBA: case LCQP_UBA: return c;
        case LCQP_LB: case LCQP_UB: case LCQP_X0: return n;
        case LCQP_Y0: return (size_t)n + c + 2 * (size_t)p;
    }
    return 0;
}

// global index of instance 0 (lcqp_cuda_set_instance_offset of the product): keys the perturbStep draws
static unsigned long long g_instance_offset = 0;
extern "C" void lcqp_emu_set_instance_offset(unsigned long long off) { g_instance_offset = off; }

extern "C" int lcqp_emu_solve_batch(int batch, int nV, int nC, int nComp, unsigned shared_mask_in,
                                    const double* Q, const double* g, const double* L, const double* R,
                                    const double* lbL, const double* ubL, const double* lbR, const double* ubR,
                                    const double* A, const double* lbA, const double* ubA,
                                    const double* lb, const double* ub, const double* x0, const double* y0,
                                    const lcqp_cuda_options* o, double* x, double* y, lcqp_cuda_stats* res)
{
    const double* base[LCQP_NUM_ARRAYS] = {Q, g, L, R, lbL, ubL, lbR, ubR, A, lbA, ubA, lb, ub, x0, y0};
    Dims d = make_dims(nV, nC, nComp, (o->qpSolver != 2) && (lb || ub));
    const unsigned all_bits = (1u << LCQP_NUM_ARRAYS) - 1u;
    const unsigned shared_mask = (batch == 1) ? all_bits : shared_mask_in;
    const unsigned mat_bits = (1u << LCQP_Q) | (1u << LCQP_L) | (1u << LCQP_R) | (nC > 0 ? (1u << LCQP_A) : 0u);
    const bool mats_shared = (shared_mask & mat_bits) == mat_bits;
    auto inst = [&](int b) {
        Inst in;
        const double** p = reinterpret_cast<const double**>(&in);
        for (int k = 0; k < LCQP_NUM_ARRAYS; k++)
            p[k] = base[k] ? base[k] + (((shared_mask >> k) & 1u) ? 0 : field_len(k, nV, nC, nComp) * (size_t)b) : nullptr;
        return in;
    };
    Scalars sc;
    // batch-level preparation (prepare_shared_kernel)
    const size_t pool_cap = (size_t)d.n * d.n + (size_t)d.m * d.n + (size_t)d.nComp * d.n + (size_t)d.nC * d.n / 4 + 32ull * (d.m + d.n) + 1024;
    std::vector<int> pool_i(pool_cap);
    std::vector<double> pool_d(pool_cap);
    int pool_used[2] = {0, 0};
    CsrPool pool = {pool_i.data(), pool_d.data(), (int)pool_cap, (int)pool_cap, pool_used};
    std::vector<double> store(mats_doubles(d));
    Mats shared_mt;
    RawOps shared_ro;
    {
        const Inst in0 = inst(0);
        raw_build_ops(d, in0, shared_ro, shared_mask, pool, &sc);
        if (mats_shared) {
            std::vector<double> v1(d.n), v2(d.n), e1(d.m), e2(d.m), e3(d.m), lo(d.m), up(d.m);
            std::vector<signed char> ctype(d.m);
            carve_mats(shared_mt, store.data(), d);
            prepare_scale(d, in0, shared_mt, v1.data(), e1.data());
            mats_build_ops_pre(d, shared_mt, pool, &sc);
            set_bounds(d, in0, shared_mt.E, lo.data(), up.data(), ctype.data(), &sc);
            int rc = prepare_factor(d, shared_mt, ctype.data(), *o, v1.data(), v2.data(), e1.data(), e2.data(), e3.data(), &sc);
            if (rc == 0) mats_dense_ops_post(d, shared_mt);
            if (rc == 0) mats_build_ops_post(d, shared_mt, pool, &sc);
            shared_mt.status = rc;
            if (rc == 0) shrink_dims(d, shared_mt.mE);
        }
    }
    // the persistent solver CTA (lcqp_solve_kernel); LCQP_EMU_PLAN selects the memory plan like the host code does:
    //   A (default) everything in "shared memory" + operator cache, B working-set inverse in "global memory"
    //   (full storage) + operator cache, C no operator cache
    const char* pl = getenv("LCQP_EMU_PLAN");
    const char plan_id = pl ? pl[0] : 'A';
    const SmemPlan plan = plan_id == 'B' ? make_plan(d, 0, true) : make_plan(d, 227 * 1024);
    size_t cache_bytes = 0;
    int cache_what = 0;
    const size_t cache_offset = (plan.bytes + 15) / 16 * 16;
    if (mats_shared && shared_mt.status == 0 && plan_id != 'C') {
        cache_requirements(shared_mt, shared_ro);
        cache_bytes = shared_mt.cache_bytes_hot; cache_what = 2;
        if (plan_id == 'A') { cache_bytes += shared_mt.cache_bytes_se + shared_mt.cache_bytes_raw; cache_what = 7; }
    }
    std::vector<double> smem_d((cache_offset + cache_bytes + 64) / 8 + 2);
    unsigned char* smem = reinterpret_cast<unsigned char*>(smem_d.data());
    std::vector<double> gl(plan.gl_doubles + 1);
    std::vector<double> ws(mats_shared ? 1 : mats_doubles(d));
    QP s;
    Work wk;
    s.d = &d;
    s.o = o;
    s.w = &wk;
    carve(wk, d, plan, smem, gl.data());
    Mats mt;
    if (mats_shared) mt = shared_mt; else carve_mats(mt, ws.data(), d);
    RawOps ro = shared_ro;
    if (cache_bytes) cache_shared_operators(d, mt, ro, smem + cache_offset, cache_bytes, cache_what);
    const int nD = nV + nC + 2 * nComp;
    int nfail = 0;
    for (int b = 0; b < batch; b++) {
        const Inst in = inst(b);
        raw_dense_ops(d, in, ro, shared_mask);
        LoopOut out;
        run_instance(s, mt, mats_shared, in, ro, g_instance_offset + (unsigned long long)b, x + (size_t)b * nV, y + (size_t)b * nD, out);
        lcqp_cuda_stats st;
        st.ret = out.ret; st.status = out.status; st.iterTotal = out.iterTotal; st.iterOuter = out.iterOuter;
        st.subproblemIter = out.subIter; st.qpExitFlag = out.exitFlag;
        st.nDuals = (o->qpSolver == 2) ? d.mA : nD;
        st.kktSolves = (int)s.n_pass;
        st.rhoOpt = out.rhoOpt; st.admmIters = (double)s.n_admm;
        res[b] = st;
        nfail += (out.ret != 0);
    }
    return nfail;
}

extern "C" void lcqp_emu_default_options(lcqp_cuda_options* o)
{
    memset(o, 0, sizeof(*o));
    o->complementarityTolerance = 1.0e3 * kEPS;
    o->stationarityTolerance = 1.0e6 * kEPS;
    o->initialPenaltyParameter = 0.01;
    o->penaltyUpdateFactor = 2.0;
    o->solveZeroPenaltyFirst = 1;
    o->perturbStep = 1;
    o->maxIterations = 1000;
    o->maxPenaltyParameter = 1e8;
    o->nDynamicPenalty = 3;
    o->etaDynamicPenalty = 0.9;
    o->qpSolver = 0;
    o->qp_rho = 0.1;
    o->qp_sigma = 1e-6;
    o->qp_alpha = 1.6;
    o->qp_delta = 1e-6;
    o->qp_feas_tol = 1e-12;
    o->qp_dual_tol = 1e-14;
    o->qp_max_iter = 4000;
    o->qp_check_interval = 10;
    o->qp_refine_iter = 10;
    o->qp_adaptive_rho = 0;
    o->perturb_seed = 1;
}

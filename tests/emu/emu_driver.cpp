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

#include "../../lcqpow_b200/csrc/lcqp_pas.cuh"
#include "../../lcqpow_b200/csrc/lcqp_osqp.cuh"
#include "../../lcqpow_b200/csrc/lcqp_sparse_host.hpp"

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

static int legacy_solve_batch(int batch, int nV, int nC, int nComp, unsigned shared_mask_in,
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

// ---- the OSQP flavour (lcqp_osqp.cuh): pattern analysis on the host, then one scalar run per instance ------------
static long long g_osqp_nnzL = 0, g_osqp_factor_flops = 0;
extern "C" void lcqp_emu_osqp_info(long long* nnzL, long long* factor_flops) { *nnzL = g_osqp_nnzL; *factor_flops = g_osqp_factor_flops; }
static int g_osqp_stream[3] = {0, 0, 0};
static long long g_osqp_sstat[8] = {0};   // per direction: levels, rows, padded entries, max E
extern "C" void lcqp_emu_osqp_stream_stats(long long* out) { for (int k = 0; k < 8; k++) out[k] = g_osqp_sstat[k]; }
static std::vector<int> g_osqp_streamI[2];
static int g_osqp_N = 0;
// the raw index stream of the last analysis (dir 0 forward, 1 backward): 16-bit words, kStreamIdx per chunk; returns the
// number of words (copies at most cap), *N = order of the KKT system
extern "C" long long lcqp_emu_osqp_stream_words(int dir, unsigned short* out, long long cap, int* N, int* words_per_chunk)
{
    const std::vector<int>& v = g_osqp_streamI[dir ? 1 : 0];
    const long long n = (long long)v.size() * 2;
    if (out) std::memcpy(out, v.data(), (size_t)std::min(n, cap) * sizeof(unsigned short));
    if (N) *N = g_osqp_N;
    if (words_per_chunk) *words_per_chunk = lcqp::osq::kStreamIdx;
    return n;
}
static void stream_stats(const std::vector<int>& Ipack, int nchunks, long long* o)
{
    const unsigned short* I = reinterpret_cast<const unsigned short*>(Ipack.data());
    o[0] = o[1] = o[2] = o[3] = 0;
    for (int c = 0; c < nchunks; c++) {
        const unsigned short* ic = I + (size_t)c * lcqp::osq::kStreamIdx;
        int p = 1;
        for (int lv = 0; lv < ic[0]; lv++) { const int r = ic[p], E = ic[p + 1]; o[0]++; o[1] += r; o[2] += r * E; if (E > o[3]) o[3] = E; p += 2 + r + r * E; }
    }
}
extern "C" void lcqp_emu_osqp_stream(int* on, int* fwd_chunks, int* bwd_chunks) { *on = g_osqp_stream[0]; *fwd_chunks = g_osqp_stream[1]; *bwd_chunks = g_osqp_stream[2]; }

static void sym_to_dev(const osq::Symbolic& S, int nC, int nComp, osq::SymDev& D)
{
    D.n = S.n; D.m = S.m; D.N = S.N; D.nC = nC; D.nComp = nComp;
    D.nnzP = (int)S.Pi.size(); D.nnzA = (int)S.Ai.size(); D.nnzQ = (int)S.Qi.size(); D.nnzK = (int)S.Ki.size(); D.nnzL = (int)S.Li.size();
    D.nflev = (int)S.flP.size() - 1; D.nblev = (int)S.blP.size() - 1;
    D.Pp = S.Pp.data(); D.Pi = S.Pi.data(); D.Psrc = S.Psrc.data(); D.Ap = S.Ap.data(); D.Ai = S.Ai.data(); D.Asrc = S.Asrc.data();
    D.Qp = S.Qp.data(); D.Qi = S.Qi.data(); D.Qsrc = S.Qsrc.data(); D.perm = S.perm.data(); D.Kp = S.Kp.data(); D.Ki = S.Ki.data(); D.Ksrc = S.Ksrc.data();
    D.Lp = S.Lp.data(); D.Li = S.Li.data(); D.rp = S.rp.data(); D.rcol = S.rcol.data(); D.rpos = S.rpos.data();
    D.Pcol = S.Pcol.data(); D.Acol = S.Acol.data(); D.Qcol = S.Qcol.data();
    D.ArP = S.ArP.data(); D.ArE = S.ArE.data(); D.PrP = S.PrP.data(); D.PrE = S.PrE.data(); D.QrP = S.QrP.data(); D.QrE = S.QrE.data();
    D.LrP = S.LrP.data(); D.LrC = S.LrC.data(); D.rposr = S.rposr.data(); D.flP = S.flP.data(); D.flR = S.flR.data(); D.blP = S.blP.data(); D.blC = S.blC.data();
    // streamed sweeps (the warp build reads the streams in place; LCQP_EMU_NOSTREAM: the level-by-level sweeps)
    D.stream = getenv("LCQP_EMU_NOSTREAM") ? 0 : S.stream; D.fsChunks = S.fsChunks; D.bsChunks = S.bsChunks;
    D.fsI = S.fsI.data(); D.bsI = S.bsI.data(); D.fsSrc = S.fsSrc.data(); D.bsSrc = S.bsSrc.data();
    D.sOff = S.sOff.data(); D.fpIdx = S.fpIdx.data(); D.fpLi = S.fpLi.data(); D.fpStep = S.fpStep.data(); D.rowPair = S.rowPair.data();
}

// osqp_admm = 1: the one-thread-per-instance build of the solver; osqp_admm = 2 (test only): the one-warp-per-instance
// build, whose CPU variant walks every loop that is declared parallel BACKWARDS (lcqp_osqp_impl.inc)
static int osqp_solve_batch(int batch, int nV, int nC, int nComp, unsigned shared_mask_in, const double* const* base,
                            const lcqp_cuda_options* o, double* x, double* y, lcqp_cuda_stats* res)
{
    const unsigned all_bits = (1u << LCQP_NUM_ARRAYS) - 1u;
    const unsigned shared_mask = (batch == 1) ? all_bits : shared_mask_in;
    const int mA = nC + 2 * nComp, nD = nV + mA;
    auto ptr = [&](int k, int b) { return base[k] ? base[k] + (((shared_mask >> k) & 1u) ? 0 : field_len(k, nV, nC, nComp) * (size_t)b) : nullptr; };
    if (base[LCQP_LB] || base[LCQP_UB]) {   // LCQProblem.cpp:930-957
        for (int b = 0; b < batch; b++) {
            memset(&res[b], 0, sizeof(res[b]));
            res[b].ret = RET_INVALID_OSQP_BOX; res[b].nDuals = mA;
            for (int j = 0; j < nV; j++) x[(size_t)b * nV + j] = base[LCQP_X0] ? ptr(LCQP_X0, b)[j] : 0.0;
            for (int j = 0; j < nD; j++) y[(size_t)b * nD + j] = 0.0;
        }
        return batch;
    }
    // union of the non-zeros over the batch
    auto mask_of = [&](int k, size_t len) {
        std::vector<unsigned char> mk(len, 0);
        if (!base[k]) return mk;
        const int reps = ((shared_mask >> k) & 1u) ? 1 : batch;
        for (int b = 0; b < reps; b++) { const double* a = base[k] + len * (size_t)b; for (size_t e = 0; e < len; e++) if (a[e] != 0.0) mk[e] = 1; }
        return mk;
    };
    const std::vector<unsigned char> Qm = mask_of(LCQP_Q, (size_t)nV * nV), Am = mask_of(LCQP_A, (size_t)nC * nV),
                                     Lm = mask_of(LCQP_L, (size_t)nComp * nV), Rm = mask_of(LCQP_R, (size_t)nComp * nV);
    std::vector<osq::Trip> Qpat, Apat;
    osq::dense_patterns(nV, nC, nComp, Qm.data(), Am.data(), Lm.data(), Rm.data(), Qpat, Apat);
    osq::Symbolic S;
    osq::analyse(nV, mA, Qpat, Apat, S);
    g_osqp_nnzL = (long long)S.Li.size(); g_osqp_factor_flops = S.factor_flops;
    g_osqp_stream[0] = S.stream; g_osqp_stream[1] = S.fsChunks; g_osqp_stream[2] = S.bsChunks;
    if (S.stream) { stream_stats(S.fsI, S.fsChunks, g_osqp_sstat); stream_stats(S.bsI, S.bsChunks, g_osqp_sstat + 4); }
    g_osqp_streamI[0] = S.fsI; g_osqp_streamI[1] = S.bsI; g_osqp_N = S.N;
    osq::SymDev D;
    sym_to_dev(S, nC, nComp, D);
    std::vector<double> ws(osq::ws_doubles(D) + 8);
    const bool warp_build = (o->osqp_admm == 2);
    osqt::Work wt;
    osqw::Work ww;
    osqt::carve(wt, D, ws.data());
    osqw::carve(ww, D, ws.data());
    int nfail = 0;
    for (int b = 0; b < batch; b++) {
        osq::View v;
        v.Q = ptr(LCQP_Q, b); v.A = ptr(LCQP_A, b); v.L = ptr(LCQP_L, b); v.R = ptr(LCQP_R, b); v.g = ptr(LCQP_G, b);
        v.lbL = ptr(LCQP_LBL, b); v.ubL = ptr(LCQP_UBL, b); v.lbR = ptr(LCQP_LBR, b); v.ubR = ptr(LCQP_UBR, b);
        v.lbA = ptr(LCQP_LBA, b); v.ubA = ptr(LCQP_UBA, b); v.x0 = ptr(LCQP_X0, b); v.y0 = ptr(LCQP_Y0, b);
        LoopOut out;
        osq::State st;
        memset(&st, 0, sizeof(st));
        if (warp_build) osqw::lcqp_loop(D, v, *o, ww, g_instance_offset + (unsigned long long)b, x + (size_t)b * nV, y + (size_t)b * nD, out, st);
        else osqt::lcqp_loop(D, v, *o, wt, g_instance_offset + (unsigned long long)b, x + (size_t)b * nV, y + (size_t)b * nD, out, st);
        for (int j = mA; j < nD; j++) y[(size_t)b * nD + j] = 0.0;
        lcqp_cuda_stats r;
        r.ret = out.ret; r.status = out.status; r.iterTotal = out.iterTotal; r.iterOuter = out.iterOuter; r.subproblemIter = out.subIter;
        r.qpExitFlag = out.exitFlag; r.nDuals = mA; r.kktSolves = (int)st.factor_count; r.rhoOpt = out.rhoOpt; r.admmIters = (double)st.admm_total;
        res[b] = r;
        nfail += (out.ret != 0);
    }
    return nfail;
}


// ---- the parametric active-set path (lcqp_pas.cuh), driven the way lcqp_cabi.cu drives it: the equality mask over the
// batch, batch-level preparation when Q/A/L/R are shared, then pas_run_instance per instance.  Instances whose Hessian is
// not positive definite (status 1) go to the regularised solver above, as in the product.
static long long g_last_solves = 0, g_last_changes = 0, g_last_polish = 0, g_last_mac = 0, g_last_byte = 0;
extern "C" void lcqp_emu_last_work(long long* macs, long long* bytes) { *macs = g_last_mac; *bytes = g_last_byte; }
extern "C" void lcqp_emu_last_counts(long long* solves, long long* changes, long long* polish)
{
    *solves = g_last_solves; *changes = g_last_changes; *polish = g_last_polish;
}

extern "C" int lcqp_emu_solve_batch(int batch, int nV, int nC, int nComp, unsigned shared_mask_in,
                                    const double* Q, const double* g, const double* L, const double* R,
                                    const double* lbL, const double* ubL, const double* lbR, const double* ubR,
                                    const double* A, const double* lbA, const double* ubA,
                                    const double* lb, const double* ub, const double* x0, const double* y0,
                                    const lcqp_cuda_options* o, double* x, double* y, lcqp_cuda_stats* res)
{
    using namespace lcqp::pas;
    if (getenv("LCQP_EMU_LEGACY"))
        return legacy_solve_batch(batch, nV, nC, nComp, shared_mask_in, Q, g, L, R, lbL, ubL, lbR, ubR, A, lbA, ubA, lb, ub, x0, y0, o, x, y, res);
    const double* base[LCQP_NUM_ARRAYS] = {Q, g, L, R, lbL, ubL, lbR, ubR, A, lbA, ubA, lb, ub, x0, y0};
    if (o->qpSolver == 2 && o->osqp_admm) return osqp_solve_batch(batch, nV, nC, nComp, shared_mask_in, base, o, x, y, res);
    const PDims d = make_pdims(nV, nC, nComp, (o->qpSolver != 2) && (lb || ub));
    Dims dold = make_dims(nV, nC, nComp, d.has_box);
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
    const size_t pool_cap = 4 * ((size_t)d.n * d.n + 3 * (size_t)d.m * d.n) / 4 + 64ull * (d.m + d.n) + 4096;
    std::vector<int> pool_i(pool_cap);
    std::vector<double> pool_d(pool_cap);
    int pool_used[2] = {0, 0};
    CsrPool pool = {pool_i.data(), pool_d.data(), (int)pool_cap, (int)pool_cap, pool_used};
    std::vector<double> store(pmats_doubles(d));
    std::vector<double> v1(d.n + d.m + 2), v2(d.n + d.m + 2);
    std::vector<signed char> eq(d.m + 16);
    PMats mt;
    carve_pmats(mt, store.data(), d);
    RawOps ro;
    {
        const Inst in0 = inst(0);
        raw_build_ops(dold, in0, ro, shared_mask, pool, &sc);
        if (mats_shared) {
            // rows that are equalities (l == u, finite) in every instance
            for (int r = 0; r < d.m; r++) {
                int all = 1;
                for (int b = 0; b < batch && all; b++) { const Inst in = inst(b); double lo, up; row_bounds(d, in, r, lo, up); all = (lo == up && lo > -qINFTY && lo < qINFTY); }
                eq[r] = (signed char)all;
            }
            pas_prepare(d, in0, mt, eq.data(), &sc);
            if (mt.status == 0) { pmats_dense_ops(d, mt); pmats_build_ops(d, mt, pool, &sc); }
        }
    }
    if (mats_shared && mt.status == 1)
        return legacy_solve_batch(batch, nV, nC, nComp, shared_mask_in, Q, g, L, R, lbL, ubL, lbR, ubR, A, lbA, ubA, lb, ub, x0, y0, o, x, y, res);
    const int mEc = mats_shared ? mt.mE : (d.m < d.n ? d.m : d.n);
    const int mIc = mats_shared ? mt.mI : d.m;
    const int capc = mats_shared ? pas_cap(d, mt.mE, mt.mI) : ((d.n < d.m) ? d.n : d.m);
    std::vector<double> smem_d(pas_smem_bytes(d, mEc, mIc, capc) / 8 + 8);
    std::vector<double> gl(pas_gl_doubles(d, mEc, capc) + 8);
    PWork wk;
    pas_carve(wk, d, mEc, mIc, capc, reinterpret_cast<unsigned char*>(smem_d.data()), gl.data());
    PQP s;
    s.d = &d; s.o = o; s.w = &wk; s.mt = &mt;
    const int nD = nV + nC + 2 * nComp;
    int nfail = 0;
    g_last_solves = g_last_changes = g_last_polish = g_last_mac = g_last_byte = 0;
    for (int b = 0; b < batch; b++) {
        const Inst in = inst(b);
        s.in = &in;
        raw_dense_ops(dold, in, ro, shared_mask);
        LoopOut out;
        const bool ok = pas_run_instance(s, mt, mats_shared, ro, g_instance_offset + (unsigned long long)b, x + (size_t)b * nV, y + (size_t)b * nD, out, eq.data());
        if (!ok) {   // semidefinite Hessian of this instance: the regularised solver
            const unsigned one_all = all_bits;
            legacy_solve_batch(1, nV, nC, nComp, one_all, in.Q, in.g, in.L, in.R, in.lbL, in.ubL, in.lbR, in.ubR, in.A, in.lbA, in.ubA, in.lb, in.ub, in.x0, in.y0, o,
                               x + (size_t)b * nV, y + (size_t)b * nD, res + b);
            nfail += (res[b].ret != 0);
            continue;
        }
        lcqp_cuda_stats st;
        st.ret = out.ret; st.status = out.status; st.iterTotal = out.iterTotal; st.iterOuter = out.iterOuter;
        st.subproblemIter = out.subIter; st.qpExitFlag = out.exitFlag;
        st.nDuals = (o->qpSolver == 2) ? d.mA : nD;
        st.kktSolves = (int)s.n_solve;
        st.rhoOpt = out.rhoOpt; st.admmIters = 0.0;
        res[b] = st;
        nfail += (out.ret != 0);
        g_last_solves += s.n_solve; g_last_changes += s.n_change; g_last_polish += s.n_polish; g_last_mac += s.n_mac; g_last_byte += s.n_byte;
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
    o->osqp_admm = 0;
    o->osqp_rho = 0.1; o->osqp_sigma = 1e-6; o->osqp_alpha = 1.6; o->osqp_delta = 1e-6;
    o->osqp_eps_abs = 1e-3; o->osqp_eps_rel = 1e-3; o->osqp_eps_prim_inf = kEPS; o->osqp_eps_dual_inf = 1e-4;
    o->osqp_adaptive_rho_tolerance = 5.0;
    o->osqp_max_iter = 4000; o->osqp_check_termination = 25; o->osqp_scaling = 10;
    o->osqp_adaptive_rho = 1; o->osqp_adaptive_rho_interval = 0; o->osqp_polish = 1; o->osqp_polish_refine_iter = 3;
    o->osqp_reserved = 0;
    o->qpoases_terminationTolerance = 5.0e6 * 2.221e-16;   // qpOASES Options.cpp:115
    o->qpoases_boundTolerance = 1.0e6 * 2.221e-16;         // :116
}

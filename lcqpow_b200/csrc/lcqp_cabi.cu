// lcqp_cabi.cu -- kernels + the extern "C" layer declared in include/lcqp_cuda.h.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC (see build.py)
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "lcqp_pas.cuh"
#include "lcqp_osqp.cuh"
#include "lcqp_sparse_host.hpp"

namespace lcqp {

#ifdef LCQP_PROFILE
__device__ unsigned long long g_prof[16];
#endif

constexpr int kThreads = 256;           // plugin-door CTA (one group)
constexpr int kMaxCtaThreads = 512;     // solver CTA: groups x threads per group (128 registers per thread)
constexpr int kMaxGroups = 4;           // groups (instances in flight) per solver CTA
constexpr int kPrepThreads = 1024;      // batch-level preparation CTA
constexpr size_t kSmemMax = 227 * 1024; // opt-in dynamic shared memory per CTA on sm_100

struct KernelArgs {
    Dims d;
    lcqp_cuda_options o;
    const double* arr[LCQP_NUM_ARRAYS];
    unsigned long long stride[LCQP_NUM_ARRAYS];  // 0 when shared
    int batch;
    unsigned shared_mask;        // loadLCQP arguments shared by the batch
    int mats_shared;             // Q, L, R, A all shared: one preparation for the batch
    SmemPlan plan;
    unsigned long long group_bytes;      // dynamic shared memory of one group (its instance state)
    unsigned long long bounds_offset;    // byte offset of the bounds shared by the groups (plan.bounds_shared), before the cache
    unsigned long long cache_offset;     // byte offset of the operator cache in dynamic shared memory (after the groups)
    unsigned long long cache_bytes;      // its size (0: operators stay in L2)
    int cache_what;                      // bit0 packed SEinv, bit1 inner-pass operators, bit2 outer-loop operators
    Mats* shared_mats;           // prepared operands of the batch (device struct, written by prepare_shared_kernel)
    RawOps* shared_raw;          // CSR / dense operators on the unscaled shared matrices
    double* shared_mats_store;   // backing store of shared_mats
    CsrPool pool;
    double* workspace;                   // per-CTA scratch: [Mats block when not shared][what did not fit in shared memory]
    unsigned long long ws_stride;        // doubles per CTA
    unsigned long long ws_mats_doubles;  // doubles of the per-CTA Mats block (0 when shared)
    double* xout;
    double* yout;
    lcqp_cuda_stats* stats;
    unsigned int* counter;
    unsigned long long instance_offset;  // global index of instance 0 (perturbStep RNG key under sharding)
};

__device__ __forceinline__ Inst make_inst(const KernelArgs& a, int b)
{
    Inst in;
    const double** p = reinterpret_cast<const double**>(&in);
    for (int k = 0; k < LCQP_NUM_ARRAYS; k++) p[k] = a.arr[k] ? a.arr[k] + a.stride[k] * (unsigned long long)b : nullptr;
    return in;
}

// Operands shared by the whole batch are prepared once by a single CTA: CSR copies of the shared unscaled
// matrices (outer loop) and, when Q, L, R and A are all shared, scaling + factorisations + their operators.
__global__ void __launch_bounds__(kPrepThreads) prepare_shared_kernel(const __grid_constant__ KernelArgs a)
{
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ Mats mt;
    __shared__ RawOps ro;
    __shared__ Scalars sc;
    const Dims& d = a.d;
    double* v1 = reinterpret_cast<double*>(smem);
    double* v2 = v1 + d.n;
    double* e1 = v2 + d.n;
    double* e2 = e1 + d.m;
    double* e3 = e2 + d.m;
    double* lo = e3 + d.m;
    double* up = lo + d.m;
    signed char* ctype = reinterpret_cast<signed char*>(up + d.m);
    CsrPool pool = a.pool;
    if (threadIdx.x == 0) { pool.used[0] = 0; pool.used[1] = 0; }
    __syncthreads();
    const Inst in = make_inst(a, 0);
    raw_build_ops(d, in, ro, a.shared_mask, pool, &sc);
    if (threadIdx.x == 0) *a.shared_raw = ro;
    if (a.mats_shared) {
        if (threadIdx.x == 0) { carve_mats(mt, a.shared_mats_store, d); mt.status = 1; }
        __syncthreads();
        prepare_scale(d, in, mt, v1, e1);
        mats_build_ops_pre(d, mt, pool, &sc);
        const int bflags = set_bounds(d, in, mt.E, lo, up, ctype, &sc);
        int rc = 1;
        if (!(bflags & 1)) rc = prepare_factor(d, mt, ctype, a.o, v1, v2, e1, e2, e3, &sc);
        else {
            // instance 0 has inconsistent bounds: prepare with its row types anyway (other instances may be fine)
            rc = prepare_factor(d, mt, ctype, a.o, v1, v2, e1, e2, e3, &sc);
        }
        if (rc == 0) mats_build_ops_post(d, mt, pool, &sc);
        __syncthreads();
        if (threadIdx.x == 0) {
            mt.status = rc;
            if (rc == 0) cache_requirements(mt, ro);
            *a.shared_mats = mt;
        }
    }
}

// The solver: persistent CTAs of blockDim.y groups x blockDim.x threads; every group works on one LCQP
// instance at a time (pulled from a global counter) with its own control flow and its own named barrier.
// The groups of a CTA share the operator cache (CSR copies of the batch-shared operators, packed SEinv).
__global__ void __launch_bounds__(kMaxCtaThreads, 1) lcqp_solve_kernel(const __grid_constant__ KernelArgs a)
{
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ Mats mt_s[kMaxGroups];
    __shared__ RawOps ro_s[kMaxGroups];
    __shared__ Work wk_s[kMaxGroups];
    __shared__ Dims dm;
    __shared__ lcqp_cuda_options opt;
    const int g = threadIdx.y, G = blockDim.y;
    __shared__ QP qp_s[kMaxGroups];
    Mats& mt = mt_s[g];
    RawOps& ro = ro_s[g];
    Work& wk = wk_s[g];
    QP& s = qp_s[g];
    double* ws = a.workspace + a.ws_stride * ((unsigned long long)blockIdx.x * G + g);
    if (threadIdx.x == 0) {
        s.d = &dm; s.o = &opt; s.w = &wk; s.mt = &mt;
        s.nw = 0; s.have_W = 0; s.tinv_valid = 0; s.n_admm = 0; s.n_pass = 0; s.n_changes = 0;
        if (g == 0) { dm = a.d; opt = a.o; }
        carve(wk, a.d, a.plan, smem + (size_t)g * a.group_bytes, ws + a.ws_mats_doubles,
              a.plan.bounds_shared ? reinterpret_cast<double*>(smem + a.bounds_offset) : nullptr);
        if (a.mats_shared) mt = *a.shared_mats;
        else carve_mats(mt, ws, a.d);
        ro = *a.shared_raw;
    }
    __syncthreads();
    if (a.mats_shared && a.cache_bytes > 0 && mt_s[0].status == 0) {
        // group 0 fills the cache; the others take over its operator descriptors
        if (g == 0) cache_shared_operators(a.d, mt, ro, smem + a.cache_offset, (size_t)a.cache_bytes, a.cache_what);
        __syncthreads();
        if (g > 0 && threadIdx.x == 0) { mt = mt_s[0]; ro = ro_s[0]; }
        __syncthreads();
    }
    // bounds shared by the batch: the one copy that the groups alias is written once, here
    const bool bounds_ready = a.plan.bounds_shared != 0 && mt_s[0].status == 0;
    if (bounds_ready) {
        if (g == 0) { const Inst in0 = make_inst(a, 0); set_bounds(a.d, in0, mt.E, wk.l, wk.ub, wk.ctype, wk.sc, true); }
        __syncthreads();
    }
    const int nD = a.d.n + a.d.mA;
    const unsigned mat_bits = (1u << LCQP_Q) | (1u << LCQP_L) | (1u << LCQP_R) | (1u << LCQP_A);
    const bool raw_all_shared = (a.shared_mask & mat_bits) == mat_bits || (a.d.nC == 0 && (a.shared_mask & mat_bits) == (mat_bits & ~(1u << LCQP_A)));

    for (;;) {
        LCQ_SYNC();
        if (threadIdx.x == 0) wk.sc->bidx = (int)atomicAdd(a.counter, 1u);
        LCQ_SYNC();
        const int b = wk.sc->bidx;
        if (b >= a.batch) break;
        const Inst in = make_inst(a, b);
        if (!raw_all_shared) {
            if (threadIdx.x == 0) raw_dense_ops(a.d, in, ro, a.shared_mask);
            LCQ_SYNC();
        }
#ifdef LCQP_PROFILE
        if (threadIdx.x == 0) { for (int k = 0; k < 16; k++) wk.sc->prof[k] = 0; wk.sc->prof_last = clock64(); wk.sc->prof_cur = 13; }
#endif
        LoopOut out;
        double* xo = a.xout + (size_t)b * a.d.n;
        double* yo = a.yout + (size_t)b * nD;
        run_instance(s, mt, a.mats_shared != 0, in, ro, a.instance_offset + (unsigned long long)b, xo, yo, out, bounds_ready);
#ifdef LCQP_PROFILE
        LCQ_PROF(wk.sc, 13);
        if (threadIdx.x == 0) for (int k = 0; k < 16; k++) atomicAdd(&g_prof[k], (unsigned long long)wk.sc->prof[k]);
#endif
        if (threadIdx.x == 0) {
            lcqp_cuda_stats st;
            st.ret = out.ret; st.status = out.status; st.iterTotal = out.iterTotal; st.iterOuter = out.iterOuter;
            st.subproblemIter = out.subIter; st.qpExitFlag = out.exitFlag;
            st.nDuals = (a.o.qpSolver == 2) ? a.d.mA : nD;
            st.kktSolves = (int)s.n_pass;
            st.rhoOpt = out.rhoOpt; st.admmIters = (double)s.n_admm;
            a.stats[b] = st;
        }
    }
}

// =================================================================================================
// The parametric active-set path (lcqp_pas.cuh)
// =================================================================================================
constexpr int kPasGroupsMax = 8;
#ifndef LCQP_PAS_CTA_THREADS
#define LCQP_PAS_CTA_THREADS 512
#endif
constexpr int kPasCtaThreads = LCQP_PAS_CTA_THREADS;   // threads of a solver CTA (the register budget per thread follows)

struct PasArgs {
    pas::PDims d;
    lcqp_cuda_options o;
    const double* arr[LCQP_NUM_ARRAYS];
    unsigned long long stride[LCQP_NUM_ARRAYS];  // 0 when shared
    int batch;
    unsigned shared_mask;
    int mats_shared;
    int bounds_vary;             // some bound array has one copy per instance
    pas::PMats* shared_mats;     // prepared operands of the batch (device struct)
    RawOps* shared_raw;
    double* shared_store;
    signed char* eqmask;         // m: row is an equality in every instance
    CsrPool pool;
    int mEc, mIc, capc;          // sizes the per-group buffers were carved for
    unsigned long long group_smem;    // dynamic shared memory per group
    double* workspace;
    unsigned long long ws_stride;     // doubles per group
    unsigned long long ws_mats;       // doubles of the per-group PMats block (0 when shared)
    double* xout;
    double* yout;
    lcqp_cuda_stats* stats;
    unsigned int* counter;
    int* fallback_flag;          // set when an instance's reduced Hessian is not positive definite
    unsigned long long* work;    // [2]: fp64 multiply-adds and bytes of the run's dense products (pas::PQP::n_mac, n_byte)
    unsigned long long instance_offset;
};

__device__ __forceinline__ Inst pas_make_inst(const PasArgs& a, int b)
{
    Inst in;
    const double** p = reinterpret_cast<const double**>(&in);
    for (int k = 0; k < LCQP_NUM_ARRAYS; k++) p[k] = a.arr[k] ? a.arr[k] + a.stride[k] * (unsigned long long)b : nullptr;
    return in;
}

// eqmask[r] stays 1 only if row r has l = u (finite) in every instance
__global__ void pas_eqmask_kernel(const __grid_constant__ PasArgs a)
{
    const int nb = a.bounds_vary ? a.batch : 1;
    for (int b = blockIdx.x; b < nb; b += gridDim.x) {
        const Inst in = pas_make_inst(a, b);
        for (int r = threadIdx.x; r < a.d.m; r += blockDim.x) {
            double lo, up;
            pas::row_bounds(a.d, in, r, lo, up);
            if (!(lo == up && lo > -pas::qINFTY && lo < pas::qINFTY)) a.eqmask[r] = 0;
        }
    }
}

// Batch-level preparation by one CTA: CSR copies of the shared unscaled matrices (outer loop) and, when Q, L, R
// and A are all shared, the prepared operands of the subsolver.
__global__ void __launch_bounds__(kPrepThreads) pas_prepare_kernel(const __grid_constant__ PasArgs a)
{
    __shared__ pas::PMats mt;
    __shared__ RawOps ro;
    __shared__ Scalars sc;
    CsrPool pool = a.pool;
    if (threadIdx.x == 0) { pool.used[0] = 0; pool.used[1] = 0; }
    __syncthreads();
    const Inst in = pas_make_inst(a, 0);
    const Dims dold = make_dims(a.d.n, a.d.nC, a.d.nComp, a.d.has_box);
    raw_build_ops(dold, in, ro, a.shared_mask, pool, &sc);
    if (threadIdx.x == 0) *a.shared_raw = ro;
    if (a.mats_shared) {
        if (threadIdx.x == 0) pas::carve_pmats(mt, a.shared_store, a.d);
        __syncthreads();
        pas::pas_prepare(a.d, in, mt, a.eqmask, &sc);
        if (mt.status == 0) {
            if (threadIdx.x == 0) pas::pmats_dense_ops(a.d, mt);
            __syncthreads();
            pas::pmats_build_ops(a.d, mt, pool, &sc);
        }
        __syncthreads();
        if (threadIdx.x == 0) *a.shared_mats = mt;
    }
}

// The solver: persistent CTAs of blockDim.y groups x blockDim.x threads; every group works on one LCQP instance at a
// time (pulled from a global counter) with its own control flow and its own named barrier.
__global__ void __launch_bounds__(kPasCtaThreads, 1) lcqp_pas_kernel(const __grid_constant__ PasArgs a)
{
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ pas::PMats mt_s[kPasGroupsMax];
    __shared__ RawOps ro_s[kPasGroupsMax];
    __shared__ pas::PWork wk_s[kPasGroupsMax];
    __shared__ pas::PQP qp_s[kPasGroupsMax];
    __shared__ Inst in_s[kPasGroupsMax];
    __shared__ pas::PDims dm;
    __shared__ lcqp_cuda_options opt;
    const int g = threadIdx.y, G = blockDim.y;
    pas::PMats& mt = mt_s[g];
    RawOps& ro = ro_s[g];
    pas::PWork& wk = wk_s[g];
    pas::PQP& s = qp_s[g];
    double* ws = a.workspace + a.ws_stride * ((unsigned long long)blockIdx.x * G + g);
    if (threadIdx.x == 0) {
        if (g == 0) { dm = a.d; opt = a.o; }
        s.d = &dm; s.o = &opt; s.w = &wk; s.mt = &mt; s.in = &in_s[g];
        s.nw = 0; s.n_solve = 0; s.n_change = 0; s.n_polish = 0; s.nwsr = 0;
        pas::pas_carve(wk, a.d, a.mEc, a.mIc, a.capc, smem + (size_t)g * a.group_smem, ws + a.ws_mats);
        if (a.mats_shared) mt = *a.shared_mats;
        else pas::carve_pmats(mt, ws, a.d);
        ro = *a.shared_raw;
    }
    __syncthreads();
    const int nD = a.d.n + a.d.mA;
    const unsigned mat_bits = (1u << LCQP_Q) | (1u << LCQP_L) | (1u << LCQP_R) | (1u << LCQP_A);
    const bool raw_all_shared = (a.shared_mask & mat_bits) == mat_bits || (a.d.nC == 0 && (a.shared_mask & mat_bits) == (mat_bits & ~(1u << LCQP_A)));
    const Dims dold = make_dims(a.d.n, a.d.nC, a.d.nComp, a.d.has_box);
    signed char* eq_scratch = reinterpret_cast<signed char*>(wk.tm2);   // free until the first QP of an instance
    for (;;) {
        LCQ_SYNC();
        if (threadIdx.x == 0) wk.sc->bidx = (int)atomicAdd(a.counter, 1u);
        LCQ_SYNC();
        const int b = wk.sc->bidx;
        if (b >= a.batch) break;
        if (threadIdx.x == 0) {
            in_s[g] = pas_make_inst(a, b);
            if (!raw_all_shared) raw_dense_ops(dold, in_s[g], ro, a.shared_mask);
        }
        LCQ_SYNC();
#ifdef LCQP_PROFILE
        if (threadIdx.x == 0) { for (int k = 0; k < 16; k++) wk.sc->prof[k] = 0; wk.sc->prof_last = clock64(); wk.sc->prof_cur = 11; }
#endif
        LoopOut out;
        double* xo = a.xout + (size_t)b * a.d.n;
        double* yo = a.yout + (size_t)b * nD;
        const bool ok = pas::pas_run_instance(s, mt, a.mats_shared != 0, ro, a.instance_offset + (unsigned long long)b, xo, yo, out, eq_scratch);
#ifdef LCQP_PROFILE
        LCQ_PROF(wk.sc, 13);
        if (threadIdx.x == 0) for (int k = 0; k < 16; k++) atomicAdd(&g_prof[k], (unsigned long long)wk.sc->prof[k]);
#endif
        if (threadIdx.x == 0) {
            lcqp_cuda_stats st;
            st.ret = ok ? out.ret : -1; st.status = out.status; st.iterTotal = out.iterTotal; st.iterOuter = out.iterOuter;
            st.subproblemIter = out.subIter; st.qpExitFlag = out.exitFlag;
            st.nDuals = (a.o.qpSolver == 2) ? a.d.mA : nD;
            st.kktSolves = (int)s.n_solve;
            st.rhoOpt = out.rhoOpt; st.admmIters = 0.0;
            a.stats[b] = st;
#ifdef LCQP_COUNT_WORK
            atomicAdd(a.work, (unsigned long long)s.n_mac);
            atomicAdd(a.work + 1, (unsigned long long)s.n_byte);
#endif
            if (!ok) atomicExch(a.fallback_flag, 1);
        }
    }
}

// fp64 FMA rate of the device (roofline denominator for the SIMT fp64 work of the solver; bench.py reports it)
__global__ void fp64_fma_probe_kernel(double* out, int iters)
{
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; i++) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

// ---- plugin door: one QP with persistent state ----------------------------------------------------
struct QPState {
    int nw, have_W, tinv_valid, prepared;
    int infeasible, iterations, flag, pad;
};

struct QPKernelArgs {
    Dims d;
    lcqp_cuda_options o;
    Inst in;            // Q, A (nC = nCtot rows), lbA, ubA, lb, ub, x0, y0, g  (device staging)
    SmemPlan plan;
    double* mats_store;
    Mats* mats;         // persistent device copy of the Mats struct
    double* gl;         // global scratch for what does not fit in shared memory
    unsigned char* saved_smem;
    unsigned long long smem_bytes;
    QPState* state;
    double* xout;  // n
    double* yout;  // n + mA
    int initial;
};

__global__ void __launch_bounds__(kThreads) qp_plugin_kernel(const __grid_constant__ QPKernelArgs a)
{
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ Mats mt;
    __shared__ Work wk;
    __shared__ Dims dm;
    __shared__ lcqp_cuda_options opt;
    __shared__ QP s;
    if (threadIdx.x == 0) {
        dm = a.d; opt = a.o; carve(wk, a.d, a.plan, smem, a.gl);
        s.d = &dm; s.o = &opt; s.w = &wk; s.mt = &mt;
        s.n_admm = 0; s.n_pass = 0; s.n_changes = 0;
    }
    __syncthreads();
    if (!a.initial) {
        for (size_t k = threadIdx.x; k < a.smem_bytes / 8; k += blockDim.x)
            reinterpret_cast<double*>(smem)[k] = reinterpret_cast<const double*>(a.saved_smem)[k];
        if (threadIdx.x == 0) { s.nw = a.state->nw; s.have_W = a.state->have_W; s.tinv_valid = a.state->tinv_valid; mt = *a.mats; }
    } else {
        if (threadIdx.x == 0) { s.nw = 0; s.have_W = 0; s.tinv_valid = 0; carve_mats(mt, a.mats_store, a.d); mats_dense_ops(a.d, mt); }
    }
    if (threadIdx.x < 8) wk.sc->wph[threadIdx.x] = 0;
    __syncthreads();
    int infeasible = a.initial ? 0 : a.state->infeasible;
    int flag = 0, iters = 0;
    if (a.initial) {
        prepare_scale(a.d, a.in, mt, wk.u, wk.zx);
        const int bflags = set_bounds(a.d, a.in, mt.E, wk.l, wk.ub, wk.ctype, wk.sc);
        infeasible = bflags & 1;
        if (!infeasible) {
            if (prepare_factor(a.d, mt, wk.ctype, a.o, wk.u, wk.t, wk.zx, wk.zp, wk.w, wk.sc)) flag = 38;
            else {
                if (threadIdx.x == 0) mats_dense_ops_post(a.d, mt);
                for (int i = LCQ_TID; i < a.d.m; i += LCQ_NT) wk.ctype[i] = mt.ctype[i];
                __syncthreads();
            }
        }
    }
    if (flag == 0) {
        const double* y0A = a.in.y0 ? a.in.y0 + a.d.n : nullptr;
        const double* y0box = (a.in.y0 && a.d.has_box) ? a.in.y0 : nullptr;
        flag = qp_solve(s, a.initial != 0, a.in.g, a.in.x0, y0A, y0box, &iters, infeasible != 0);
    }
    if (flag == 0) {
        for (int j = LCQ_TID; j < a.d.n; j += LCQ_NT) {
            a.xout[j] = mt.D[j] * wk.x[j];
            a.yout[j] = a.d.has_box ? wk.ys[a.d.mA + j] : 0.0;
        }
        for (int i = LCQ_TID; i < a.d.mA; i += LCQ_NT) a.yout[a.d.n + i] = wk.ys[i];
    }
    __syncthreads();
    for (size_t k = threadIdx.x; k < a.smem_bytes / 8; k += blockDim.x)
        reinterpret_cast<double*>(a.saved_smem)[k] = reinterpret_cast<const double*>(smem)[k];
    if (threadIdx.x == 0) {
        *a.mats = mt;
        a.state->nw = s.nw; a.state->have_W = s.have_W; a.state->tinv_valid = s.tinv_valid;
        a.state->infeasible = infeasible; a.state->iterations = iters; a.state->flag = flag;
    }
}

// ---- plugin door on the parametric active-set solver: one QP, state kept in device memory between the calls ----
struct PasQPState {
    int nw, cap, e_moving, stamp_next, rampOffset, prepared;
    int iterations, flag;
    double phi, len0;
};

struct PasQPArgs {
    pas::PDims d;
    lcqp_cuda_options o;
    Inst in;                 // Q, A (nC = nCtot rows), g, lbA, ubA, lb, ub, x0, y0 (device staging)
    int mEc, mIc, capc;
    double* mats_store;
    pas::PMats* mats;        // persistent header of the prepared operands
    double* gl;              // persistent global scratch (working-set inverse, gradients, duals)
    unsigned char* saved_smem;
    unsigned long long smem_bytes;
    PasQPState* state;
    double* xout;            // n
    double* yout;            // n + mA
    int initial;
};

__global__ void __launch_bounds__(kThreads) pas_plugin_kernel(const __grid_constant__ PasQPArgs a)
{
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ pas::PMats mt;
    __shared__ pas::PWork wk;
    __shared__ pas::PQP s;
    __shared__ pas::PDims dm;
    __shared__ lcqp_cuda_options opt;
    __shared__ Inst in;
    __shared__ RawOps ro;
    if (threadIdx.x == 0) {
        dm = a.d; opt = a.o; in = a.in;
        pas::pas_carve(wk, a.d, a.mEc, a.mIc, a.capc, smem, a.gl);
        s.d = &dm; s.o = &opt; s.w = &wk; s.mt = &mt; s.in = &in;
        s.n_solve = 0; s.n_change = 0; s.n_polish = 0; s.nwsr = 0;
        ro.Q = dense_op(a.in.Q, a.d.n, a.d.n, a.d.n, 0);
    }
    __syncthreads();
    if (threadIdx.x < 8) wk.sc->wph[threadIdx.x] = 0;
    int flag = 0;
    if (a.initial) {
        if (threadIdx.x == 0) pas::carve_pmats(mt, a.mats_store, a.d);
        __syncthreads();
        signed char* eq = reinterpret_cast<signed char*>(wk.tm2);
        for (int r = threadIdx.x; r < a.d.m; r += blockDim.x) { double lo, up; pas::row_bounds(a.d, in, r, lo, up); eq[r] = (lo == up && lo > -pas::qINFTY && lo < pas::qINFTY) ? 1 : 0; }
        __syncthreads();
        pas::pas_prepare(a.d, in, mt, eq, wk.sc);
        if (threadIdx.x == 0 && mt.status == 0) pas::pmats_dense_ops(a.d, mt);
        __syncthreads();
        if (mt.status != 0) flag = pas::QP_SETUP;
        if (threadIdx.x == 0) { s.cap = pas::pas_cap(a.d, mt.mE, mt.mI); s.nw = 0; }
        __syncthreads();
    } else {
        for (size_t k = threadIdx.x; k < a.smem_bytes / 8; k += blockDim.x)
            reinterpret_cast<double*>(smem)[k] = reinterpret_cast<const double*>(a.saved_smem)[k];
        if (threadIdx.x == 0) {
            mt = *a.mats;
            s.nw = a.state->nw; s.cap = a.state->cap; s.e_moving = a.state->e_moving; s.stamp_next = a.state->stamp_next;
            s.rampOffset = a.state->rampOffset; s.phi = a.state->phi; s.len0 = a.state->len0;
        }
        __syncthreads();
    }
    if (flag == 0) {
        for (int j = threadIdx.x; j < a.d.n; j += blockDim.x) wk.gk[j] = in.g[j];
        __syncthreads();
        if (a.initial) {
            for (int j = threadIdx.x; j < a.d.n; j += blockDim.x) wk.xk[j] = in.x0 ? in.x0[j] : 0.0;
            __syncthreads();
            op_mv(ro.Q, wk.xk, nullptr, -1.0, wk.gq);
            __syncthreads();
            const double* y0A = in.y0 ? in.y0 + a.d.n : nullptr;
            const double* y0box = (in.y0 && a.d.has_box) ? in.y0 : nullptr;
            flag = pas::pas_init(s, ro, wk.xk, y0A, y0box);
        } else {
            flag = pas::pas_hotstart(s, ro);
        }
    }
    if (flag == 0) {
        for (int j = threadIdx.x; j < a.d.n; j += blockDim.x) {
            a.xout[j] = wk.xq[j];
            a.yout[j] = a.d.has_box ? wk.ys[a.d.mA + j] : 0.0;
        }
        for (int i = threadIdx.x; i < a.d.mA; i += blockDim.x) a.yout[a.d.n + i] = wk.ys[i];
    }
    __syncthreads();
    for (size_t k = threadIdx.x; k < a.smem_bytes / 8; k += blockDim.x)
        reinterpret_cast<double*>(a.saved_smem)[k] = reinterpret_cast<const double*>(smem)[k];
    if (threadIdx.x == 0) {
        *a.mats = mt;
        a.state->nw = s.nw; a.state->cap = s.cap; a.state->e_moving = s.e_moving; a.state->stamp_next = s.stamp_next;
        a.state->rampOffset = s.rampOffset; a.state->phi = s.phi; a.state->len0 = s.len0;
        a.state->iterations = s.nwsr; a.state->flag = flag; a.state->prepared = 1 + mt.status;
    }
}


// ---- the OSQP flavour (lcqp_osqp.cuh): one warp per CTA; mode T = 32 instances per warp, mode W = one instance per warp ----
struct OsqpArgs {
    osq::SymDev S;
    lcqp_cuda_options o;
    const double* arr[LCQP_NUM_ARRAYS];          // Q, A, L, R hold VALUE arrays indexed through S.*src (dense or csc layout)
    unsigned long long stride[LCQP_NUM_ARRAYS];  // doubles between consecutive instances (0: shared)
    int batch;
    double* workspace;                           // per resident warp: ws_doubles x (32 | 1) doubles
    unsigned long long ws_doubles;
    double* xout;
    double* yout;
    lcqp_cuda_stats* stats;
    unsigned int* counter;
    unsigned long long instance_offset;
    int box_given;                               // lb / ub were loaded: INVALID_OSQP_BOX_CONSTRAINTS (LCQProblem.cpp:930-957)
    unsigned smem_bytes;                         // solve / factor scratch per warp in shared memory when that fits, else 0 (global, L2)
    unsigned ring_offset;                        // mode W, streamed sweeps: byte offset of the bulk-copy ring in the dynamic shared memory (0: none)
};

LCQ_DEV osq::View osqp_view(const OsqpArgs& a, int b)
{
    osq::View v;
    auto at = [&](int k) -> const double* { return a.arr[k] ? a.arr[k] + a.stride[k] * (unsigned long long)b : nullptr; };
    v.Q = at(LCQP_Q); v.A = at(LCQP_A); v.L = at(LCQP_L); v.R = at(LCQP_R); v.g = at(LCQP_G);
    v.lbL = at(LCQP_LBL); v.ubL = at(LCQP_UBL); v.lbR = at(LCQP_LBR); v.ubR = at(LCQP_UBR);
    v.lbA = at(LCQP_LBA); v.ubA = at(LCQP_UBA); v.x0 = at(LCQP_X0); v.y0 = at(LCQP_Y0);
    return v;
}
LCQ_DEV void osqp_state_init(osq::State& st)
{
    st.rho = 0; st.c = 1; st.cinv = 1; st.pri_res = 0; st.dua_res = 0; st.status_val = 0; st.iter = 0; st.interval = 0;
    st.factor_bad = 0; st.admm_total = 0; st.factor_count = 0;
}
LCQ_DEV lcqp_cuda_stats osqp_stats(const LoopOut& out, const osq::State& st, int mA)
{
    lcqp_cuda_stats r;
    r.ret = out.ret; r.status = out.status; r.iterTotal = out.iterTotal; r.iterOuter = out.iterOuter;
    r.subproblemIter = out.subIter; r.qpExitFlag = out.exitFlag; r.nDuals = mA; r.kktSolves = (int)st.factor_count;
    r.rhoOpt = out.rhoOpt; r.admmIters = (double)st.admm_total;
    return r;
}

// mode T: thread = instance; the last tile is padded with copies of the last instance (every lane takes part in the
// warp's votes; the padded lanes recompute instance batch - 1 bit for bit, so their stores carry the same values)
__global__ void __launch_bounds__(32) lcqp_osqp_kernel(const __grid_constant__ OsqpArgs a)
{
    const int lane = threadIdx.x;
    extern __shared__ __align__(16) unsigned char osqp_smem[];
    osqt::Work w;
    osqt::carve(w, a.S, a.workspace + (size_t)blockIdx.x * a.ws_doubles * 32 + lane, a.smem_bytes ? reinterpret_cast<double*>(osqp_smem) + lane : nullptr);
    const int nV = a.S.n, mA = a.S.m, nD = nV + mA;
    for (;;) {
        unsigned tile = 0;
        if (lane == 0) tile = atomicAdd(a.counter, 1u);
        tile = __shfl_sync(0xffffffffu, tile, 0);
        if ((unsigned long long)tile * 32ull >= (unsigned long long)a.batch) break;
        const int b_lane = (int)tile * 32 + lane;
        const bool valid = b_lane < a.batch;
        const int b = valid ? b_lane : a.batch - 1;
        const osq::View v = osqp_view(a, b);
        LoopOut out;
        osq::State st;
        osqp_state_init(st);
        double* xo = a.xout + (size_t)b * nV;
        double* yo = a.yout + (size_t)b * nD;
        if (a.box_given) {
            out.ret = RET_INVALID_OSQP_BOX; out.status = 0; out.iterTotal = 0; out.iterOuter = 0; out.subIter = 0; out.exitFlag = 0; out.rhoOpt = 0;
            for (int j = 0; j < nV; j++) xo[j] = v.x0 ? v.x0[j] : 0.0;
            for (int j = 0; j < mA; j++) yo[j] = 0.0;
        } else
            osqt::lcqp_loop(a.S, v, a.o, w, a.instance_offset + (unsigned long long)b, xo, yo, out, st);
        for (int j = mA; j < nD; j++) yo[j] = 0.0;
        if (valid) a.stats[b] = osqp_stats(out, st, mA);
        __syncwarp();
    }
}

// mode W: warp = instance
__global__ void __launch_bounds__(32) lcqp_osqpw_kernel(const __grid_constant__ OsqpArgs a)
{
    const int lane = threadIdx.x;
    extern __shared__ __align__(16) unsigned char osqp_smem[];
    osqw::Work w;
    osqw::carve(w, a.S, a.workspace + (size_t)blockIdx.x * a.ws_doubles, a.smem_bytes ? reinterpret_cast<double*>(osqp_smem) : nullptr);
    if (a.ring_offset) {
        w.ring = osqp_smem + a.ring_offset;
        w.ring_uses = reinterpret_cast<unsigned*>(w.ring + osq::kStreamStages * osq::stream_stage_bytes() + 8 * osq::kStreamStages);
        osqw::ring_init(w);
    }
#ifdef LCQP_PROFILE
    __shared__ long long oprof[16];
    if (lane == 0) { for (int k = 0; k < 16; k++) oprof[k] = 0; oprof[14] = clock64(); }
    w.prof = oprof;
    __syncwarp();
#endif
    const int nV = a.S.n, mA = a.S.m, nD = nV + mA;
    for (;;) {
        unsigned b = 0;
        if (lane == 0) b = atomicAdd(a.counter, 1u);
        b = __shfl_sync(0xffffffffu, b, 0);
        if (b >= (unsigned)a.batch) break;
        const osq::View v = osqp_view(a, (int)b);
        LoopOut out;
        osq::State st;
        osqp_state_init(st);
        double* xo = a.xout + (size_t)b * nV;
        double* yo = a.yout + (size_t)b * nD;
        if (a.box_given) {
            out.ret = RET_INVALID_OSQP_BOX; out.status = 0; out.iterTotal = 0; out.iterOuter = 0; out.subIter = 0; out.exitFlag = 0; out.rhoOpt = 0;
            for (int j = lane; j < nV; j += 32) xo[j] = v.x0 ? v.x0[j] : 0.0;
            for (int j = lane; j < mA; j += 32) yo[j] = 0.0;
        } else
            osqw::lcqp_loop(a.S, v, a.o, w, a.instance_offset + (unsigned long long)b, xo, yo, out, st);
        for (int j = mA + lane; j < nD; j += 32) yo[j] = 0.0;
        if (lane == 0) a.stats[b] = osqp_stats(out, st, mA);
        __syncwarp();
    }
#ifdef LCQP_PROFILE
    OSQ_PROF(w, 0);
    if (lane == 0) for (int k = 0; k < 14; k++) atomicAdd(&g_prof[k], (unsigned long long)oprof[k]);
#endif
}

}  // namespace lcqp

// =================================================================================================
// host side
// =================================================================================================
using namespace lcqp;

// Launch-plan overrides for development A/B runs (tools/gpu_ab.sh): compiled in only with -DLCQP_TUNING, so the shipped
// library takes no decision from the environment (LCQP_CUDA_VERBOSE only prints the plan).
static inline const char* tune_env(const char* name)
{
#ifdef LCQP_TUNING
    return getenv(name);
#else
    (void)name;
    return nullptr;
#endif
}

struct lcqp_cuda_handle_s {
    int nV, nC, nComp, capacity, device;
    lcqp_cuda_options opts;
    int batch = 0;
    unsigned shared_mask = 0;
    bool loaded = false, ran = false;
    const double* dev_in[LCQP_NUM_ARRAYS] = {};
    double* own_in[LCQP_NUM_ARRAYS] = {};
    size_t own_in_cap[LCQP_NUM_ARRAYS] = {};
    double* xout = nullptr;
    double* yout = nullptr;
    lcqp_cuda_stats* stats = nullptr;
    unsigned int* counter = nullptr;
    double* shared_store = nullptr;
    size_t shared_store_cap = 0;
    Mats* shared_mats = nullptr;
    RawOps* shared_raw = nullptr;
    int* pool_i = nullptr;
    double* pool_d = nullptr;
    size_t pool_cap = 0;
    int* pool_used = nullptr;
    Mats* host_mats = nullptr;  // pinned read-back of the prepared header (mE, status)
    double* workspace = nullptr;
    size_t workspace_cap = 0;
    int num_sms = 0;
    long long launches = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr;
    cudaStream_t last_stream = nullptr;
    unsigned long long instance_offset = 0;
    int last_grid = 0, last_smem = 0, last_mE = 0;
    // parametric active-set path
    cudaStream_t load_stream = nullptr;   // copies + preparation of a load (lcqp_cuda_load returns when they are done)
    pas::PMats* pas_mats = nullptr;       // device header of the prepared operands
    pas::PMats* pas_host = nullptr;       // pinned read-back (mE, mI, status)
    double* pas_store = nullptr;
    size_t pas_store_cap = 0;
    signed char* eqmask = nullptr;
    size_t eqmask_cap = 0;
    int* fallback_flag = nullptr;
    unsigned long long* work = nullptr;   // device, [2] (see PasArgs::work)
    int* fallback_host = nullptr;         // pinned
    bool pas_ready = false;               // prepared for the current load
    bool use_legacy = false;              // reduced Hessian not positive definite: the regularised solver runs
    int has_box = 0;
    // OSQP flavour (qpSolver == 2 with osqp_admm): symbolic analysis of the batch's pattern + value arrays
    bool osqp_ready = false;
    osq::Symbolic* sym = nullptr;
    osq::SymDev symdev;
    int* sym_ints = nullptr;
    size_t sym_ints_cap = 0;
    double* csc_vals[4] = {nullptr, nullptr, nullptr, nullptr};   // Q, A, L, R values of the sparse door
    size_t csc_vals_cap[4] = {0, 0, 0, 0};
    const double* osqp_arr[LCQP_NUM_ARRAYS] = {};
    unsigned long long osqp_stride[LCQP_NUM_ARRAYS] = {};
    long long osqp_nnzL = 0;
    int osqp_warp_mode = 0;
    std::string err;
};

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

static int fail(lcqp_cuda_handle h, int code, const char* what, cudaError_t e = cudaSuccess)
{
    if (h) {
        h->err = what;
        if (e != cudaSuccess) { h->err += ": "; h->err += cudaGetErrorString(e); }
    }
    return code;
}

#define CK(call, code) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(h, code, #call, e_); } while (0)

static size_t prep_smem_bytes(const Dims& d) { return (2ull * d.n + 5ull * d.m) * sizeof(double) + d.m + 64; }

extern "C" {

int lcqp_cuda_abi_version(void) { return LCQP_CUDA_ABI_VERSION; }

void lcqp_cuda_default_options(lcqp_cuda_options* o)
{
    memset(o, 0, sizeof(*o));
    o->complementarityTolerance = 1.0e3 * kEPS;  // Options.cpp:297-298
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
    o->osqp_rho = 0.1; o->osqp_sigma = 1e-6; o->osqp_alpha = 1.6; o->osqp_delta = 1e-6;   // constants.h:59-78
    o->osqp_eps_abs = 1e-3; o->osqp_eps_rel = 1e-3; o->osqp_eps_prim_inf = kEPS; o->osqp_eps_dual_inf = 1e-4;   // Options.cpp:329
    o->osqp_adaptive_rho_tolerance = 5.0;
    o->osqp_max_iter = 4000; o->osqp_check_termination = 25; o->osqp_scaling = 10;
    o->osqp_adaptive_rho = 1; o->osqp_adaptive_rho_interval = 0; o->osqp_polish = 1; o->osqp_polish_refine_iter = 3;   // Options.cpp:331
    o->osqp_reserved = 0;
    o->qpoases_terminationTolerance = 5.0e6 * 2.221e-16;   // qpOASES Options.cpp:115
    o->qpoases_boundTolerance = 1.0e6 * 2.221e-16;         // :116
}

int lcqp_cuda_create(int nV, int nC, int nComp, int batch_capacity, int device, lcqp_cuda_handle* out)
{
    if (!out) return LCQP_CUDA_BAD_ARGUMENT;
    *out = nullptr;
    if (nV <= 0 || nComp <= 0) return 106;  // INVALID_NUMBER_OF_OPTIM_VARS (LCQProblem.cpp:46-56)
    if (nC < 0) return 108;                 // INVALID_NUMBER_OF_CONSTRAINT_VARS (:58-63)
    if (batch_capacity <= 0) return LCQP_CUDA_BAD_ARGUMENT;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) return LCQP_CUDA_NO_DEVICE;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return LCQP_CUDA_NO_DEVICE;
    if (prop.major != 10) return LCQP_CUDA_NO_DEVICE;  // sm_100a cubin only
    lcqp_cuda_handle h = new (std::nothrow) lcqp_cuda_handle_s();
    if (!h) return LCQP_CUDA_OUT_OF_MEMORY;
    h->nV = nV; h->nC = nC; h->nComp = nComp; h->capacity = batch_capacity; h->device = device;
    h->num_sms = prop.multiProcessorCount;
    lcqp_cuda_default_options(&h->opts);
    if (cudaSetDevice(device) != cudaSuccess) { delete h; return LCQP_CUDA_NO_DEVICE; }
    const size_t nD = (size_t)nV + nC + 2 * (size_t)nComp;
    bool ok = cudaMalloc(&h->xout, sizeof(double) * nV * (size_t)batch_capacity) == cudaSuccess &&
              cudaMalloc(&h->yout, sizeof(double) * nD * (size_t)batch_capacity) == cudaSuccess &&
              cudaMalloc(&h->stats, sizeof(lcqp_cuda_stats) * (size_t)batch_capacity) == cudaSuccess &&
              cudaMalloc(&h->counter, sizeof(unsigned int)) == cudaSuccess &&
              cudaMalloc(&h->shared_mats, sizeof(Mats)) == cudaSuccess &&
              cudaMalloc(&h->shared_raw, sizeof(RawOps)) == cudaSuccess &&
              cudaMalloc(&h->pool_used, 2 * sizeof(int)) == cudaSuccess &&
              cudaMallocHost(&h->host_mats, sizeof(Mats)) == cudaSuccess &&
              cudaMalloc(&h->pas_mats, sizeof(pas::PMats)) == cudaSuccess &&
              cudaMallocHost(&h->pas_host, sizeof(pas::PMats)) == cudaSuccess &&
              cudaMalloc(&h->fallback_flag, sizeof(int)) == cudaSuccess &&
              cudaMalloc(&h->work, 2 * sizeof(unsigned long long)) == cudaSuccess &&
              cudaMallocHost(&h->fallback_host, sizeof(int)) == cudaSuccess &&
              cudaStreamCreateWithFlags(&h->load_stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaEventCreate(&h->ev0) == cudaSuccess && cudaEventCreate(&h->ev1) == cudaSuccess &&
              cudaEventCreate(&h->ev2) == cudaSuccess;
    if (!ok) { cudaGetLastError(); lcqp_cuda_destroy(h); return LCQP_CUDA_OUT_OF_MEMORY; }
    *out = h;
    return LCQP_CUDA_OK;
}

int lcqp_cuda_destroy(lcqp_cuda_handle h)
{
    if (!h) return LCQP_CUDA_BAD_HANDLE;
    cudaSetDevice(h->device);
    for (int k = 0; k < LCQP_NUM_ARRAYS; k++) if (h->own_in[k]) cudaFree(h->own_in[k]);
    cudaFree(h->xout); cudaFree(h->yout); cudaFree(h->stats); cudaFree(h->counter);
    cudaFree(h->shared_store); cudaFree(h->shared_mats); cudaFree(h->shared_raw);
    cudaFree(h->pool_i); cudaFree(h->pool_d); cudaFree(h->pool_used); cudaFree(h->workspace);
    if (h->host_mats) cudaFreeHost(h->host_mats);
    cudaFree(h->pas_mats); cudaFree(h->pas_store); cudaFree(h->eqmask); cudaFree(h->fallback_flag); cudaFree(h->work);
    cudaFree(h->sym_ints);
    for (int k = 0; k < 4; k++) cudaFree(h->csc_vals[k]);
    delete h->sym;
    if (h->pas_host) cudaFreeHost(h->pas_host);
    if (h->fallback_host) cudaFreeHost(h->fallback_host);
    if (h->load_stream) cudaStreamDestroy(h->load_stream);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->ev2) cudaEventDestroy(h->ev2);
    delete h;
    return LCQP_CUDA_OK;
}

// knobs of the regularised solver (semidefinite Hessians): a non-positive interval would spin its ADMM loop forever
static bool qp_knobs_valid(const lcqp_cuda_options* o)
{
    return o->qp_check_interval >= 1 && o->qp_max_iter >= 1 && o->qp_rho > 0 && o->qp_sigma > 0 && o->qp_delta > 0 && o->qp_alpha > 0 &&
           o->qp_alpha < 2 && o->qp_refine_iter >= 0;
}

int lcqp_cuda_set_options(lcqp_cuda_handle h, const lcqp_cuda_options* o)
{
    if (!h) return LCQP_CUDA_BAD_HANDLE;
    if (!o) return LCQP_CUDA_BAD_ARGUMENT;
    // the validations of Options' setters (/root/reference/src/Options.cpp:85-259)
    if (o->stationarityTolerance <= kEPS) return 105;
    if (o->complementarityTolerance <= kEPS) return 102;
    if (o->initialPenaltyParameter <= 1.0e-25) return 103;
    if (o->penaltyUpdateFactor <= 1) return 101;
    if (o->maxIterations <= 0) return 104;
    if (o->maxPenaltyParameter <= 1.0e-25) return 121;
    if (o->etaDynamicPenalty <= 0 || o->etaDynamicPenalty >= 1) return 119;
    if (o->qpSolver < 0 || o->qpSolver > 2) return 109;
    if (o->nDynamicPenalty > kMaxLeyffer) return LCQP_CUDA_BAD_ARGUMENT;
    if (!qp_knobs_valid(o)) return LCQP_CUDA_BAD_ARGUMENT;
    if (!(o->qpoases_terminationTolerance > 0) || !(o->qpoases_boundTolerance > 0)) return LCQP_CUDA_BAD_ARGUMENT;
    if (o->qpSolver == 2 && o->osqp_admm &&
        (!(o->osqp_rho > 0) || !(o->osqp_sigma > 0) || !(o->osqp_alpha > 0 && o->osqp_alpha < 2) || !(o->osqp_delta > 0) || o->osqp_max_iter < 1 ||
         o->osqp_check_termination < 0 || o->osqp_scaling < 0 || o->osqp_polish_refine_iter < 0 || !(o->osqp_adaptive_rho_tolerance >= 1) ||
         !(o->osqp_eps_abs >= 0) || !(o->osqp_eps_rel >= 0) || !(o->osqp_eps_prim_inf > 0) || !(o->osqp_eps_dual_inf > 0)))
        return LCQP_CUDA_BAD_ARGUMENT;   // validate_settings, external/osqp/src/auxil.c:880-1000
    h->opts = *o;
    return LCQP_CUDA_OK;
}

static int pas_prepare_load(lcqp_cuda_handle h);
static int osqp_prepare_dense(lcqp_cuda_handle h, int batch, unsigned shared_mask, const double* const* ptr);
static int osqp_upload_symbolic(lcqp_cuda_handle h);

static int load_common(lcqp_cuda_handle h, int batch, unsigned shared_mask, const double* const* ptr, bool device_ptrs)
{
    if (!h) return LCQP_CUDA_BAD_HANDLE;
    if (batch <= 0 || batch > h->capacity) return fail(h, LCQP_CUDA_BAD_ARGUMENT, "batch out of range");
    // a previous run may still be reading the staging buffers
    if (h->ran && h->last_stream != nullptr) cudaStreamSynchronize(h->last_stream);
    else if (h->ran) cudaDeviceSynchronize();
    // loadLCQP argument checks (LCQProblem.cpp:87-144, :563-626, LCQProblem.ipp:39-50)
    if (!ptr[LCQP_Q]) return fail(h, LCQP_CUDA_BAD_ARGUMENT, "Q is NULL");
    if (!ptr[LCQP_G]) return 116;                                    // INVALID_OBJECTIVE_LINEAR_TERM
    if (!ptr[LCQP_A] && h->nC > 0) return 117;                       // INVALID_CONSTRAINT_MATRIX
    if (!ptr[LCQP_L] || !ptr[LCQP_R]) return 118;                    // INVALID_COMPLEMENTARITY_MATRIX
    CK(cudaSetDevice(h->device), LCQP_CUDA_NO_DEVICE);
    for (int k = 0; k < LCQP_NUM_ARRAYS; k++) {
        const size_t len = field_len(k, h->nV, h->nC, h->nComp);
        if (!ptr[k] || len == 0) { h->dev_in[k] = nullptr; continue; }
        const size_t count = ((shared_mask >> k) & 1u) ? len : len * (size_t)batch;
        if (device_ptrs) { h->dev_in[k] = ptr[k]; continue; }
        if (h->own_in_cap[k] < count) {
            if (h->own_in[k]) cudaFree(h->own_in[k]);
            h->own_in[k] = nullptr; h->own_in_cap[k] = 0;
            const size_t want = ((shared_mask >> k) & 1u) ? count : len * (size_t)h->capacity;
            if (cudaMalloc(&h->own_in[k], want * sizeof(double)) != cudaSuccess) { cudaGetLastError(); return fail(h, LCQP_CUDA_OUT_OF_MEMORY, "cudaMalloc(input)"); }
            h->own_in_cap[k] = want;
        }
        CK(cudaMemcpyAsync(h->own_in[k], ptr[k], count * sizeof(double), cudaMemcpyHostToDevice, h->load_stream), LCQP_CUDA_LAUNCH_FAILED);
        h->dev_in[k] = h->own_in[k];
    }
    h->batch = batch;
    h->shared_mask = shared_mask;
    h->loaded = true;
    h->ran = false;
    h->pas_ready = false;
    h->osqp_ready = false;
    if (h->opts.qpSolver == 2 && h->opts.osqp_admm) {
        // the OSQP flavour: pattern analysis on the host (needs the caller's HOST arrays), no operand preparation
        if (device_ptrs) return fail(h, LCQP_CUDA_BAD_ARGUMENT, "the OSQP flavour (osqp_admm) loads through lcqp_cuda_load or lcqp_cuda_load_csc");
        const int rc = osqp_prepare_dense(h, batch, shared_mask, ptr);
        if (rc != LCQP_CUDA_OK) return rc;
        CK(cudaStreamSynchronize(h->load_stream), LCQP_CUDA_LAUNCH_FAILED);
        h->osqp_ready = true;
        return LCQP_CUDA_OK;
    }
    // a load is complete when it returns: the copies are done (the caller may reuse its buffers, and any stream may
    // run the batch) and the batch-level operands are prepared, so that lcqp_cuda_run is one asynchronous launch
    const int rc = pas_prepare_load(h);
    if (rc != LCQP_CUDA_OK) return rc;
    CK(cudaStreamSynchronize(h->load_stream), LCQP_CUDA_LAUNCH_FAILED);
    h->use_legacy = (h->pas_host->status == 1);
    if (h->pas_host->status > 1) return fail(h, LCQP_CUDA_LAUNCH_FAILED, "preparation of the shared operands failed");
    h->pas_ready = true;
    return LCQP_CUDA_OK;
}

int lcqp_cuda_load(lcqp_cuda_handle h, int batch, unsigned shared_mask,
                   const double* Q, const double* g, const double* L, const double* R,
                   const double* lbL, const double* ubL, const double* lbR, const double* ubR,
                   const double* A, const double* lbA, const double* ubA,
                   const double* lb, const double* ub, const double* x0, const double* y0)
{
    const double* p[LCQP_NUM_ARRAYS] = {Q, g, L, R, lbL, ubL, lbR, ubR, A, lbA, ubA, lb, ub, x0, y0};
    return load_common(h, batch, shared_mask, p, false);
}

int lcqp_cuda_load_device(lcqp_cuda_handle h, int batch, unsigned shared_mask,
                          const double* Q, const double* g, const double* L, const double* R,
                          const double* lbL, const double* ubL, const double* lbR, const double* ubR,
                          const double* A, const double* lbA, const double* ubA,
                          const double* lb, const double* ub, const double* x0, const double* y0)
{
    const double* p[LCQP_NUM_ARRAYS] = {Q, g, L, R, lbL, ubL, lbR, ubR, A, lbA, ubA, lb, ub, x0, y0};
    return load_common(h, batch, shared_mask, p, true);
}

int lcqp_cuda_load_csc(lcqp_cuda_handle h, int batch, unsigned shared_mask,
                       const int* Q_p, const int* Q_i, const double* Q_x, const double* g,
                       const int* L_p, const int* L_i, const double* L_x,
                       const int* R_p, const int* R_i, const double* R_x,
                       const double* lbL, const double* ubL, const double* lbR, const double* ubR,
                       const int* A_p, const int* A_i, const double* A_x, const double* lbA, const double* ubA,
                       const double* x0, const double* y0)
{
    if (!h) return LCQP_CUDA_BAD_HANDLE;
    if (batch <= 0 || batch > h->capacity) return fail(h, LCQP_CUDA_BAD_ARGUMENT, "batch out of range");
    if (!(h->opts.qpSolver == 2 && h->opts.osqp_admm)) return fail(h, LCQP_CUDA_BAD_ARGUMENT, "the sparse door serves the OSQP flavour (qpSolver = 2, osqp_admm = 1)");
    if (!Q_p || !Q_i || !Q_x) return fail(h, LCQP_CUDA_BAD_ARGUMENT, "Q is NULL");
    if (!g) return 116;
    if (h->nC > 0 && (!A_p || !A_i || !A_x)) return 117;
    if (!L_p || !L_i || !L_x || !R_p || !R_i || !R_x) return 118;
    if (h->ran && h->last_stream != nullptr) cudaStreamSynchronize(h->last_stream);
    else if (h->ran) cudaDeviceSynchronize();
    CK(cudaSetDevice(h->device), LCQP_CUDA_NO_DEVICE);
    const int n = h->nV, nC = h->nC, nComp = h->nComp;
    // index arrays: every column pointer monotone, every row index in range (Utilities.hpp INVALID_INDEX_POINTER / _ARRAY)
    auto check = [&](const int* p, const int* i, int rows) -> int {
        if (p[0] != 0) return 400;
        for (int c = 0; c < n; c++) if (p[c + 1] < p[c]) return 400;
        for (int e = 0; e < p[n]; e++) if (i[e] < 0 || i[e] >= rows) return 401;
        return 0;
    };
    int rc = check(Q_p, Q_i, n);
    if (!rc) rc = check(L_p, L_i, nComp);
    if (!rc) rc = check(R_p, R_i, nComp);
    if (!rc && nC > 0) rc = check(A_p, A_i, nC);
    if (rc) return fail(h, rc, "invalid csc index arrays");
    // vectors through the staging buffers of the dense door; matrices as value arrays
    const double* vec[LCQP_NUM_ARRAYS] = {nullptr, g, nullptr, nullptr, lbL, ubL, lbR, ubR, nullptr, lbA, ubA, nullptr, nullptr, x0, y0};
    for (int k = 0; k < LCQP_NUM_ARRAYS; k++) {
        h->dev_in[k] = nullptr; h->osqp_arr[k] = nullptr; h->osqp_stride[k] = 0;
        const size_t len = field_len(k, n, nC, nComp);
        if (!vec[k] || len == 0) continue;
        const bool sh = (shared_mask >> k) & 1u;
        const size_t count = sh ? len : len * (size_t)batch;
        if (h->own_in_cap[k] < count) {
            if (h->own_in[k]) cudaFree(h->own_in[k]);
            h->own_in[k] = nullptr; h->own_in_cap[k] = 0;
            const size_t want = sh ? count : len * (size_t)h->capacity;
            if (cudaMalloc(&h->own_in[k], want * sizeof(double)) != cudaSuccess) { cudaGetLastError(); return fail(h, LCQP_CUDA_OUT_OF_MEMORY, "cudaMalloc(input)"); }
            h->own_in_cap[k] = want;
        }
        CK(cudaMemcpyAsync(h->own_in[k], vec[k], count * sizeof(double), cudaMemcpyHostToDevice, h->load_stream), LCQP_CUDA_LAUNCH_FAILED);
        h->osqp_arr[k] = h->own_in[k];
        h->osqp_stride[k] = sh ? 0ull : (unsigned long long)len;
    }
    const int which[4] = {LCQP_Q, LCQP_A, LCQP_L, LCQP_R};
    const double* vals[4] = {Q_x, A_x, L_x, R_x};
    const size_t nnz[4] = {(size_t)Q_p[n], (nC > 0) ? (size_t)A_p[n] : 0, (size_t)L_p[n], (size_t)R_p[n]};
    for (int k = 0; k < 4; k++) {
        if (!vals[k] || nnz[k] == 0) continue;
        const bool sh = (shared_mask >> which[k]) & 1u;
        const size_t count = sh ? nnz[k] : nnz[k] * (size_t)batch;
        if (h->csc_vals_cap[k] < count) {
            if (h->csc_vals[k]) cudaFree(h->csc_vals[k]);
            h->csc_vals[k] = nullptr; h->csc_vals_cap[k] = 0;
            if (cudaMalloc(&h->csc_vals[k], count * sizeof(double)) != cudaSuccess) { cudaGetLastError(); return fail(h, LCQP_CUDA_OUT_OF_MEMORY, "cudaMalloc(csc values)"); }
            h->csc_vals_cap[k] = count;
        }
        CK(cudaMemcpyAsync(h->csc_vals[k], vals[k], count * sizeof(double), cudaMemcpyHostToDevice, h->load_stream), LCQP_CUDA_LAUNCH_FAILED);
        h->osqp_arr[which[k]] = h->csc_vals[k];
        h->osqp_stride[which[k]] = sh ? 0ull : (unsigned long long)nnz[k];
    }
    std::vector<osq::Trip> Qpat, Apat;
    osq::csc_patterns(n, nC, nComp, Q_p, Q_i, nC > 0 ? A_p : nullptr, A_i, L_p, L_i, R_p, R_i, Qpat, Apat);
    if (!h->sym) h->sym = new (std::nothrow) osq::Symbolic();
    if (!h->sym) return fail(h, LCQP_CUDA_OUT_OF_MEMORY, "symbolic");
    osq::analyse(n, nC + 2 * nComp, Qpat, Apat, *h->sym);
    rc = osqp_upload_symbolic(h);
    if (rc != LCQP_CUDA_OK) return rc;
    CK(cudaStreamSynchronize(h->load_stream), LCQP_CUDA_LAUNCH_FAILED);
    h->batch = batch;
    h->shared_mask = shared_mask;
    h->loaded = true;
    h->ran = false;
    h->pas_ready = false;
    h->osqp_ready = true;
    return LCQP_CUDA_OK;
}

int lcqp_cuda_set_instance_offset(lcqp_cuda_handle h, unsigned long long off)
{
    if (!h) return LCQP_CUDA_BAD_HANDLE;
    h->instance_offset = off;
    return LCQP_CUDA_OK;
}

static int run_legacy(lcqp_cuda_handle h, cudaStream_t stream)
{
    KernelArgs a;
    memset(&a, 0, sizeof(a));
    Dims& d = a.d;
    d = make_dims(h->nV, h->nC, h->nComp, (h->opts.qpSolver != 2) && (h->dev_in[LCQP_LB] || h->dev_in[LCQP_UB]));
    a.o = h->opts;
    for (int k = 0; k < LCQP_NUM_ARRAYS; k++) {
        a.arr[k] = h->dev_in[k];
        a.stride[k] = ((h->shared_mask >> k) & 1u) ? 0ull : (unsigned long long)field_len(k, h->nV, h->nC, h->nComp);
    }
    a.batch = h->batch;
    const unsigned all_bits = (1u << LCQP_NUM_ARRAYS) - 1u;
    a.shared_mask = (h->batch == 1) ? all_bits : h->shared_mask;
    const unsigned mat_bits = (1u << LCQP_Q) | (1u << LCQP_L) | (1u << LCQP_R) | (h->nC > 0 ? (1u << LCQP_A) : 0u);
    a.mats_shared = ((a.shared_mask & mat_bits) == mat_bits);

    // batch-level store: Mats backing store + CSR pool
    const size_t md = mats_doubles(d);
    if (md > h->shared_store_cap) {
        if (h->shared_store) cudaFree(h->shared_store);
        h->shared_store = nullptr; h->shared_store_cap = 0;
        if (cudaMalloc(&h->shared_store, md * sizeof(double)) != cudaSuccess) { cudaGetLastError(); return fail(h, LCQP_CUDA_OUT_OF_MEMORY, "cudaMalloc(shared store)"); }
        h->shared_store_cap = md;
    }
    const size_t pool_cap = (size_t)d.n * d.n + (size_t)d.m * d.n + (size_t)d.nComp * d.n + (size_t)d.nC * d.n / 4 + 32ull * (d.m + d.n) + 1024;
    if (pool_cap > h->pool_cap) {
        if (h->pool_i) cudaFree(h->pool_i);
        if (h->pool_d) cudaFree(h->pool_d);
        h->pool_i = nullptr; h->pool_d = nullptr; h->pool_cap = 0;
        if (cudaMalloc(&h->pool_i, pool_cap * sizeof(int)) != cudaSuccess || cudaMalloc(&h->pool_d, pool_cap * sizeof(double)) != cudaSuccess) {
            cudaGetLastError();
            return fail(h, LCQP_CUDA_OUT_OF_MEMORY, "cudaMalloc(csr pool)");
        }
        h->pool_cap = pool_cap;
    }
    a.shared_mats = h->shared_mats;
    a.shared_raw = h->shared_raw;
    a.shared_mats_store = h->shared_store;
    a.pool.ibuf = h->pool_i; a.pool.dbuf = h->pool_d; a.pool.icap = (int)h->pool_cap; a.pool.dcap = (int)h->pool_cap; a.pool.used = h->pool_used;
    a.xout = h->xout; a.yout = h->yout; a.stats = h->stats; a.counter = h->counter;
    a.instance_offset = h->instance_offset;

    CK(cudaMemsetAsync(h->counter, 0, sizeof(unsigned int), stream), LCQP_CUDA_LAUNCH_FAILED);
    CK(cudaEventRecord(h->ev0, stream), LCQP_CUDA_LAUNCH_FAILED);
    {
        const size_t psm = prep_smem_bytes(d);
        if (psm > kSmemMax) return fail(h, LCQP_CUDA_TOO_LARGE, "instance does not fit the shared-memory budget (prepare)");
        CK(cudaFuncSetAttribute(prepare_shared_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psm), LCQP_CUDA_LAUNCH_FAILED);
        prepare_shared_kernel<<<1, kPrepThreads, psm, stream>>>(a);
        h->launches++;
        CK(cudaGetLastError(), LCQP_CUDA_LAUNCH_FAILED);
    }
    int mE = -1;
    if (a.mats_shared) {
        // the order of the static equality block decides the shared-memory plan: read the header back
        CK(cudaMemcpyAsync(h->host_mats, h->shared_mats, sizeof(Mats), cudaMemcpyDeviceToHost, stream), LCQP_CUDA_LAUNCH_FAILED);
        CK(cudaStreamSynchronize(stream), LCQP_CUDA_LAUNCH_FAILED);
        mE = h->host_mats->mE;
        if (h->host_mats->status == 0) shrink_dims(d, mE);
    }
    h->last_mE = mE;
    CK(cudaEventRecord(h->ev1, stream), LCQP_CUDA_LAUNCH_FAILED);

    // Shared-memory plan.  A CTA is G groups of T threads (G*T <= 512: the solver needs 128 registers per
    // thread); each group holds the state of one instance in shared memory, the operator cache behind the
    // groups is shared by all of them.  G is maximised first (independent instances hide each other's
    // latencies), then, as far as shared memory goes: the working-set inverse of each group (else it lives in
    // L2, full storage), the inner-pass operators, the packed inverse of the static equality block, the
    // outer-loop vectors, the outer-loop operators.
    // Tuning aids: LCQP_CUDA_THREADS (T), LCQP_CUDA_GROUPS (upper bound on G), LCQP_CUDA_TINV=l2|smem,
    // LCQP_CUDA_CACHE (bit mask: 1 SEinv, 2 inner-pass operators, 4 outer-loop operators).
    const bool can_cache = a.mats_shared && h->host_mats->status == 0;
    const size_t c_se = can_cache ? (size_t)h->host_mats->cache_bytes_se : 0;
    const size_t c_hot = can_cache ? (size_t)h->host_mats->cache_bytes_hot : 0;
    const size_t c_raw = can_cache ? (size_t)h->host_mats->cache_bytes_raw : 0;
    int threads = 128;
    if (const char* t = tune_env("LCQP_CUDA_THREADS")) { const int v = atoi(t); if (v >= 32 && v <= 256 && v % 32 == 0) threads = v; }
    int gmax = kMaxCtaThreads / threads;
    if (gmax > kMaxGroups) gmax = kMaxGroups;
    if (const char* t = tune_env("LCQP_CUDA_GROUPS")) { const int v = atoi(t); if (v >= 1 && v < gmax) gmax = v; }
    if (gmax > h->batch) gmax = h->batch;
    int tinv_pref = -1;   // -1: in shared memory when it fits
    if (const char* t = tune_env("LCQP_CUDA_TINV")) tinv_pref = (t[0] == 's') ? 1 : (t[0] == 'l' ? 0 : -1);
    int cache_allow = 7;
    if (const char* t = tune_env("LCQP_CUDA_CACHE")) cache_allow = atoi(t) & 7;
    cudaFuncAttributes fattr;
    CK(cudaFuncGetAttributes(&fattr, lcqp_solve_kernel), LCQP_CUDA_LAUNCH_FAILED);
    const size_t budget = kSmemMax - fattr.sharedSizeBytes;   // static shared memory: the descriptors of the groups
    SmemPlan plan;
    int groups = 0;
    a.cache_bytes = 0; a.cache_what = 0; a.cache_offset = 0;
    const SmemPlan pmin = make_plan(d, 0, true);   // the instance's QP vectors only
    // Bounds that the whole batch shares (every bound array shared or absent, and shared scaling) scale to the
    // same l / ub for every instance: the groups alias ONE copy (written once per CTA before the instance loop).
    const unsigned bound_bits = (1u << LCQP_LBL) | (1u << LCQP_UBL) | (1u << LCQP_LBR) | (1u << LCQP_UBR) | (1u << LCQP_LBA) | (1u << LCQP_UBA) | (1u << LCQP_LB) | (1u << LCQP_UB);
    bool bounds_shared = can_cache;
    for (int k = 0; k < LCQP_NUM_ARRAYS; k++)
        if (((bound_bits >> k) & 1u) && h->dev_in[k] && !((a.shared_mask >> k) & 1u)) bounds_shared = false;
    if (tune_env("LCQP_CUDA_NO_SLIM")) bounds_shared = false;   // tuning aid
    const size_t bnd_bytes = bounds_shared ? 2 * ev((size_t)d.m) * sizeof(double) : 0;
    const size_t ys_bytes = ev((size_t)d.m) * sizeof(double);
    size_t bounds_bytes_used = 0;
    for (int G = gmax; G >= 1 && !groups; G--) {
        const size_t base = (pmin.bytes - bnd_bytes + 15) / 16 * 16;
        if ((size_t)G * base + bnd_bytes > budget) continue;
        // per-group extras, in order of preference
        size_t per = base;
        size_t cache = bnd_bytes;
        int what = 0;
        bool tinv_smem = false, outer_smem = false, ys_global = false;
        const size_t tb = tinv_doubles(d) * sizeof(double), ob = outer_doubles(d) * sizeof(double);
        auto fits = [&](size_t per_new, size_t cache_new) { return (size_t)G * ((per_new + 15) / 16 * 16) + cache_new <= budget; };
        if (tinv_pref != 0 && fits(per + tb, cache)) { per += tb; tinv_smem = true; }
        if (tinv_pref == 1 && !tinv_smem) continue;
        if ((cache_allow & 2) && c_hot && fits(per, cache + c_hot)) { cache += c_hot; what |= 2; }
        if ((cache_allow & 1) && c_se) {
            if (fits(per, cache + c_se)) { cache += c_se; what |= 1; }
            else if (!tune_env("LCQP_CUDA_NO_SLIM") && fits(per - ys_bytes, cache + c_se)) {
                // the accepted duals move to the global scratch (with the outer-loop vectors) to make room
                per -= ys_bytes; ys_global = true; cache += c_se; what |= 1;
            }
        }
        if (!ys_global && fits(per + ob, cache)) { per += ob; outer_smem = true; }
        if ((cache_allow & 4) && c_raw && (what & 2) && fits(per, cache + c_raw)) { cache += c_raw; what |= 4; }
        plan = pmin;
        plan.tinv_in_smem = tinv_smem;
        plan.outer_in_smem = outer_smem;
        plan.ys_global = ys_global;
        plan.bounds_shared = bounds_shared;
        plan.gl_doubles = (tinv_smem ? 0 : tinv_gl_doubles(d)) + (outer_smem ? 0 : outer_doubles(d)) + (ys_global ? ev((size_t)d.m) : 0);
        plan.bytes = (per + 15) / 16 * 16;
        groups = G;
        a.cache_bytes = cache - bnd_bytes; a.cache_what = what;
        bounds_bytes_used = bnd_bytes;
    }
    if (!groups) return fail(h, LCQP_CUDA_TOO_LARGE, "instance does not fit the shared-memory budget");
    a.plan = plan;
    a.group_bytes = plan.bytes;
    a.bounds_offset = (unsigned long long)groups * plan.bytes;
    a.cache_offset = a.bounds_offset + bounds_bytes_used;
    const size_t smem = (size_t)a.cache_offset + a.cache_bytes;
    CK(cudaFuncSetAttribute(lcqp_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), LCQP_CUDA_LAUNCH_FAILED);
    int per_sm = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lcqp_solve_kernel, threads * groups, smem), LCQP_CUDA_LAUNCH_FAILED);
    if (const char* t = tune_env("LCQP_CUDA_CTAS_PER_SM")) { const int v = atoi(t); if (v >= 1 && v < per_sm) per_sm = v; }  // tuning aid
    if (per_sm < 1) return fail(h, LCQP_CUDA_TOO_LARGE, "kernel cannot be resident");
    int grid = per_sm * h->num_sms;
    if ((long long)grid * groups > h->batch) grid = (h->batch + groups - 1) / groups;

    // per-CTA global scratch
    a.ws_mats_doubles = a.mats_shared ? 0 : md;
    a.ws_stride = a.ws_mats_doubles + plan.gl_doubles;
    const size_t ws_total = a.ws_stride * (size_t)grid * groups;
    if (ws_total > h->workspace_cap) {
        if (h->workspace) cudaFree(h->workspace);
        h->workspace = nullptr; h->workspace_cap = 0;
        if (cudaMalloc(&h->workspace, (ws_total ? ws_total : 1) * sizeof(double)) != cudaSuccess) { cudaGetLastError(); return fail(h, LCQP_CUDA_OUT_OF_MEMORY, "cudaMalloc(workspace)"); }
        h->workspace_cap = ws_total;
    }
    a.workspace = h->workspace;

    if (getenv("LCQP_CUDA_VERBOSE"))
        fprintf(stderr, "lcqp_cuda: grid %d x (%d threads x %d groups), %d CTA/SM, smem %zu B = %d x %zu (tinv %s, outer %s, ys %s, bounds %s) + cache %llu (what %d; se %zu hot %zu raw %zu), mE %d cap %d\n",
                grid, threads, groups, per_sm, smem, groups, (size_t)plan.bytes, plan.tinv_in_smem ? "smem" : "L2", plan.outer_in_smem ? "smem" : "L2",
                plan.ys_global ? "L2" : "smem", plan.bounds_shared ? "shared" : "own",
                a.cache_bytes, a.cache_what, c_se, c_hot, c_raw, mE, d.cap);
    lcqp_solve_kernel<<<grid, dim3(threads, groups), smem, stream>>>(a);
    h->launches++;
    CK(cudaGetLastError(), LCQP_CUDA_LAUNCH_FAILED);
    CK(cudaEventRecord(h->ev2, stream), LCQP_CUDA_LAUNCH_FAILED);
#ifdef LCQP_PROFILE
    if (getenv("LCQP_CUDA_VERBOSE")) {
        cudaStreamSynchronize(stream);
        unsigned long long pr[16];
        cudaMemcpyFromSymbol(pr, g_prof, sizeof(pr));
        unsigned long long z[16] = {};
        cudaMemcpyToSymbol(g_prof, z, sizeof(z));
        double tot = 0;
        for (int k = 0; k < 16; k++) tot += (double)pr[k];
        static const char* nm[16] = {"kkt_residual", "kkt1 K^-1", "kkt2 rows+Tinv", "kkt3 At+K^-1", "linesearch", "ratio", "tinv_append", "as misc/update", "kkt_check", "tinv_remove", "admm", "outer", "qp misc", "setup/out", "res: oP+oA", "res: oAt"};
        fprintf(stderr, "lcqp_cuda profile (group cycles per LCQP, %% of total):\n");
        for (int k = 0; k < 16; k++) fprintf(stderr, "   %-16s %12.0f  %5.1f%%\n", nm[k], (double)pr[k] / h->batch, 100.0 * pr[k] / (tot > 0 ? tot : 1));
    }
#endif
    h->last_stream = stream;
    h->last_grid = grid * groups;
    h->last_smem = (int)smem;
    h->ran = true;
    return LCQP_CUDA_OK;
}

// ---- the parametric active-set path -----------------------------------------------------------------------
static void pas_fill_args(lcqp_cuda_handle h, PasArgs& a)
{
    memset(&a, 0, sizeof(a));
    a.d = pas::make_pdims(h->nV, h->nC, h->nComp, h->has_box);
    a.o = h->opts;
    unsigned bound_bits = (1u << LCQP_LBL) | (1u << LCQP_UBL) | (1u << LCQP_LBR) | (1u << LCQP_UBR) | (1u << LCQP_LBA) | (1u << LCQP_UBA) | (1u << LCQP_LB) | (1u << LCQP_UB);
    const unsigned all_bits = (1u << LCQP_NUM_ARRAYS) - 1u;
    a.shared_mask = (h->batch == 1) ? all_bits : h->shared_mask;
    for (int k = 0; k < LCQP_NUM_ARRAYS; k++) {
        a.arr[k] = h->dev_in[k];
        a.stride[k] = ((a.shared_mask >> k) & 1u) ? 0ull : (unsigned long long)field_len(k, h->nV, h->nC, h->nComp);
        if (((bound_bits >> k) & 1u) && h->dev_in[k] && a.stride[k]) a.bounds_vary = 1;
    }
    a.batch = h->batch;
    const unsigned mat_bits = (1u << LCQP_Q) | (1u << LCQP_L) | (1u << LCQP_R) | (h->nC > 0 ? (1u << LCQP_A) : 0u);
    a.mats_shared = ((a.shared_mask & mat_bits) == mat_bits);
    a.shared_mats = h->pas_mats;
    a.shared_raw = h->shared_raw;
    a.shared_store = h->pas_store;
    a.eqmask = h->eqmask;
    a.pool.ibuf = h->pool_i; a.pool.dbuf = h->pool_d; a.pool.icap = (int)h->pool_cap; a.pool.dcap = (int)h->pool_cap; a.pool.used = h->pool_used;
    a.xout = h->xout; a.yout = h->yout; a.stats = h->stats; a.counter = h->counter;
    a.fallback_flag = h->fallback_flag;
    a.work = h->work;
    a.instance_offset = h->instance_offset;
}

// Batch-level preparation of a load (on the load stream): equality mask over the batch, CSR copies of the shared
// raw matrices, prepared operands when Q, L, R, A are shared.  The header (mE, mI, status) is read back.
static int pas_prepare_load(lcqp_cuda_handle h)
{
    h->has_box = (h->dev_in[LCQP_LB] || h->dev_in[LCQP_UB]) ? 1 : 0;
    const pas::PDims d = pas::make_pdims(h->nV, h->nC, h->nComp, h->has_box);
    const size_t md = pas::pmats_doubles(d);
    if (md > h->pas_store_cap) {
        if (h->pas_store) cudaFree(h->pas_store);
        h->pas_store = nullptr; h->pas_store_cap = 0;
        if (cudaMalloc(&h->pas_store, md * sizeof(double)) != cudaSuccess) { cudaGetLastError(); return fail(h, LCQP_CUDA_OUT_OF_MEMORY, "cudaMalloc(prepared operands)"); }
        h->pas_store_cap = md;
    }
    if ((size_t)d.m + 16 > h->eqmask_cap) {
        if (h->eqmask) cudaFree(h->eqmask);
        h->eqmask = nullptr; h->eqmask_cap = 0;
        if (cudaMalloc(&h->eqmask, (size_t)d.m + 16) != cudaSuccess) { cudaGetLastError(); return fail(h, LCQP_CUDA_OUT_OF_MEMORY, "cudaMalloc(eqmask)"); }
        h->eqmask_cap = (size_t)d.m + 16;
    }
    const size_t pool_cap = 2 * (size_t)d.n * d.n + 3 * (size_t)d.m * d.n + 64ull * (d.m + d.n) + 4096;
    if (pool_cap > h->pool_cap) {
        if (h->pool_i) cudaFree(h->pool_i);
        if (h->pool_d) cudaFree(h->pool_d);
        h->pool_i = nullptr; h->pool_d = nullptr; h->pool_cap = 0;
        if (cudaMalloc(&h->pool_i, pool_cap * sizeof(int)) != cudaSuccess || cudaMalloc(&h->pool_d, pool_cap * sizeof(double)) != cudaSuccess) {
            cudaGetLastError();
            return fail(h, LCQP_CUDA_OUT_OF_MEMORY, "cudaMalloc(csr pool)");
        }
        h->pool_cap = pool_cap;
    }
    PasArgs a;
    pas_fill_args(h, a);
    cudaStream_t st = h->load_stream;
    CK(cudaMemsetAsync(h->eqmask, 1, (size_t)d.m, st), LCQP_CUDA_LAUNCH_FAILED);
    {
        int grid = a.bounds_vary ? (h->batch < 4 * h->num_sms ? h->batch : 4 * h->num_sms) : 1;
        pas_eqmask_kernel<<<grid, 128, 0, st>>>(a);
        h->launches++;
        CK(cudaGetLastError(), LCQP_CUDA_LAUNCH_FAILED);
    }
    pas_prepare_kernel<<<1, kPrepThreads, 0, st>>>(a);
    h->launches++;
    CK(cudaGetLastError(), LCQP_CUDA_LAUNCH_FAILED);
    memset(h->pas_host, 0, sizeof(pas::PMats));
    if (a.mats_shared) CK(cudaMemcpyAsync(h->pas_host, h->pas_mats, sizeof(pas::PMats), cudaMemcpyDeviceToHost, st), LCQP_CUDA_LAUNCH_FAILED);
    return LCQP_CUDA_OK;
}

// ---- the OSQP flavour: symbolic analysis at load time, one launch per run -----------------------------------------
static int osqp_upload_symbolic(lcqp_cuda_handle h)
{
    const osq::Symbolic& S = *h->sym;
    constexpr int NV = osq::kSymArrays;
    const std::vector<int>* vs[NV] = {&S.Pp, &S.Pi, &S.Psrc, &S.Ap, &S.Ai, &S.Asrc, &S.Qp, &S.Qi, &S.Qsrc, &S.perm, &S.Kp, &S.Ki, &S.Ksrc,
                                      &S.Lp, &S.Li, &S.rp, &S.rcol, &S.rpos, &S.Pcol, &S.Acol, &S.Qcol,
                                      &S.ArP, &S.ArE, &S.PrP, &S.PrE, &S.QrP, &S.QrE, &S.LrP, &S.LrC, &S.rposr, &S.flP, &S.flR, &S.blP, &S.blC,
                                      &S.fsI, &S.bsI, &S.fsSrc, &S.bsSrc, &S.sOff, &S.fpIdx, &S.fpLi, &S.fpStep, &S.rowPair};
    size_t total = 0;
    size_t off[NV];
    for (int k = 0; k < NV; k++) { off[k] = total; total += (vs[k]->size() + 3) & ~(size_t)3; }
    std::vector<int> pack(total ? total : 4, 0);
    for (int k = 0; k < NV; k++) if (!vs[k]->empty()) memcpy(pack.data() + off[k], vs[k]->data(), vs[k]->size() * sizeof(int));
    if (total > h->sym_ints_cap) {
        if (h->sym_ints) cudaFree(h->sym_ints);
        h->sym_ints = nullptr; h->sym_ints_cap = 0;
        if (cudaMalloc(&h->sym_ints, (total ? total : 4) * sizeof(int)) != cudaSuccess) { cudaGetLastError(); return fail(h, LCQP_CUDA_OUT_OF_MEMORY, "cudaMalloc(symbolic)"); }
        h->sym_ints_cap = total;
    }
    CK(cudaMemcpyAsync(h->sym_ints, pack.data(), pack.size() * sizeof(int), cudaMemcpyHostToDevice, h->load_stream), LCQP_CUDA_LAUNCH_FAILED);
    CK(cudaStreamSynchronize(h->load_stream), LCQP_CUDA_LAUNCH_FAILED);   // `pack` goes out of scope
    osq::SymDev& D = h->symdev;
    D.n = S.n; D.m = S.m; D.N = S.N; D.nC = h->nC; D.nComp = h->nComp;
    D.nnzP = (int)S.Pi.size(); D.nnzA = (int)S.Ai.size(); D.nnzQ = (int)S.Qi.size(); D.nnzK = (int)S.Ki.size(); D.nnzL = (int)S.Li.size();
    D.nflev = (int)S.flP.size() - 1; D.nblev = (int)S.blP.size() - 1;
    D.fsChunks = S.fsChunks; D.bsChunks = S.bsChunks; D.stream = S.stream;
    const int** dst[NV] = {&D.Pp, &D.Pi, &D.Psrc, &D.Ap, &D.Ai, &D.Asrc, &D.Qp, &D.Qi, &D.Qsrc, &D.perm, &D.Kp, &D.Ki, &D.Ksrc, &D.Lp, &D.Li, &D.rp, &D.rcol, &D.rpos,
                           &D.Pcol, &D.Acol, &D.Qcol, &D.ArP, &D.ArE, &D.PrP, &D.PrE, &D.QrP, &D.QrE, &D.LrP, &D.LrC, &D.rposr, &D.flP, &D.flR, &D.blP, &D.blC,
                           &D.fsI, &D.bsI, &D.fsSrc, &D.bsSrc, &D.sOff, &D.fpIdx, &D.fpLi, &D.fpStep, &D.rowPair};
    for (int k = 0; k < NV; k++) *dst[k] = h->sym_ints + off[k];
    h->osqp_nnzL = (long long)S.Li.size();
    return LCQP_CUDA_OK;
}

// dense door: the batch's union of non-zeros decides the pattern (host scan of the caller's arrays)
static int osqp_prepare_dense(lcqp_cuda_handle h, int batch, unsigned shared_mask, const double* const* ptr)
{
    const int n = h->nV, nC = h->nC, nComp = h->nComp;
    auto mask_of = [&](int k, size_t len) {
        std::vector<unsigned char> mk(len ? len : 1, 0);
        if (!ptr[k]) return mk;
        const int reps = ((shared_mask >> k) & 1u) ? 1 : batch;
        for (int b = 0; b < reps; b++) { const double* a = ptr[k] + len * (size_t)b; for (size_t e = 0; e < len; e++) if (a[e] != 0.0) mk[e] = 1; }
        return mk;
    };
    const std::vector<unsigned char> Qm = mask_of(LCQP_Q, (size_t)n * n), Am = mask_of(LCQP_A, (size_t)nC * n),
                                     Lm = mask_of(LCQP_L, (size_t)nComp * n), Rm = mask_of(LCQP_R, (size_t)nComp * n);
    std::vector<osq::Trip> Qpat, Apat;
    osq::dense_patterns(n, nC, nComp, Qm.data(), Am.data(), Lm.data(), Rm.data(), Qpat, Apat);
    if (!h->sym) h->sym = new (std::nothrow) osq::Symbolic();
    if (!h->sym) return fail(h, LCQP_CUDA_OUT_OF_MEMORY, "symbolic");
    osq::analyse(n, nC + 2 * nComp, Qpat, Apat, *h->sym);
    for (int k = 0; k < LCQP_NUM_ARRAYS; k++) {
        h->osqp_arr[k] = h->dev_in[k];
        h->osqp_stride[k] = ((shared_mask >> k) & 1u) ? 0ull : (unsigned long long)field_len(k, n, nC, nComp);
    }
    return osqp_upload_symbolic(h);
}

static int run_osqp(lcqp_cuda_handle h, cudaStream_t stream)
{
    if (!h->osqp_ready) return fail(h, LCQP_CUDA_NOT_LOADED, "osqp_admm was set after the load: load again (the pattern analysis happens at load time)");
    OsqpArgs a;
    memset(&a, 0, sizeof(a));
    a.S = h->symdev;
    a.o = h->opts;
    for (int k = 0; k < LCQP_NUM_ARRAYS; k++) { a.arr[k] = h->osqp_arr[k]; a.stride[k] = (h->batch == 1) ? 0ull : h->osqp_stride[k]; }
    a.batch = h->batch;
    a.box_given = (h->osqp_arr[LCQP_LB] || h->osqp_arr[LCQP_UB]) ? 1 : 0;
    a.ws_doubles = (osq::ws_doubles(a.S) + 1) & ~(size_t)1;
    // One warp per instance when the factor is sparse (the level sets give the lanes something to share) or the problem
    // is large, or when the batch would not fill the GPU with one thread per instance; else one thread per instance.
    const long long tiles32 = ((long long)h->batch + 31) / 32;
    bool warp_mode = (a.S.nnzL <= 8 * a.S.N) || a.S.N > 512 || tiles32 < 4ll * h->num_sms;
    if (const char* t = tune_env("LCQP_CUDA_OSQP_MODE")) warp_mode = (t[0] == 'w');
    const int lanes = warp_mode ? 1 : 32;
    // the scratch vector of the triangular solves and of the factorisation (chains of dependent read-modify-writes) sits
    // in shared memory when it fits the CTA's budget
    const size_t sm_need = osq::sm_len(a.S) * lanes * sizeof(double);
    a.smem_bytes = (sm_need <= (size_t)kSmemMax - 1024) ? (unsigned)sm_need : 0u;
    if (tune_env("LCQP_CUDA_OSQP_NOSMEM")) a.smem_bytes = 0;
    // streamed triangular solves (one warp per instance): the ring that cp.async.bulk fills sits behind the scratch vector
    size_t dyn_smem = a.smem_bytes;
    a.ring_offset = 0;
    if (!warp_mode) a.S.stream = 0;   // (one thread per instance reads L in place; its workspace holds no streams)
    // (they pay when a sweep is a long chain of small levels -- C4: 334 levels of ~6 rows; with a handful of wide levels
    //  -- C2: 4 levels -- reading L in place is faster: 1.66k against 1.13k LCQP/s, r2p)
    if (warp_mode && a.S.stream && (a.S.nflev + a.S.nblev >= 32 || tune_env("LCQP_CUDA_OSQP_FORCESTREAM")) && !tune_env("LCQP_CUDA_OSQP_NOSTREAM")) {
        const size_t off = a.smem_bytes ? ((size_t)a.smem_bytes + 127) & ~(size_t)127 : 128;
        if (off + osq::stream_ring_bytes() <= (size_t)kSmemMax - 1024) { a.ring_offset = (unsigned)off; dyn_smem = off + osq::stream_ring_bytes(); }
    }
    if (!a.ring_offset) a.S.stream = 0;
    a.ws_doubles = (osq::ws_doubles(a.S) + 1) & ~(size_t)1;
    auto kern = warp_mode ? lcqp_osqpw_kernel : lcqp_osqp_kernel;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_smem), LCQP_CUDA_LAUNCH_FAILED);
    int per_sm = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32, dyn_smem), LCQP_CUDA_LAUNCH_FAILED);
    if (per_sm < 1) return fail(h, LCQP_CUDA_TOO_LARGE, "kernel cannot be resident");
    if (per_sm > 16) per_sm = 16;
    // one warp per CTA; as many resident warps as there are work items, up to per_sm per SM and a quarter of the device memory
    const long long items = warp_mode ? (long long)h->batch : tiles32;
    long long warps = (long long)h->num_sms * per_sm;
    if (warps > items) warps = items;
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    const size_t per_warp = a.ws_doubles * lanes * sizeof(double);
    const size_t budget = total_b / 4 > h->workspace_cap * sizeof(double) ? total_b / 4 : h->workspace_cap * sizeof(double);
    while (warps > 1 && (size_t)warps * per_warp > budget) warps = (warps * 3) / 4;
    const size_t ws_total = (size_t)warps * a.ws_doubles * lanes;
    if (ws_total > h->workspace_cap) {
        if (h->workspace) { cudaDeviceSynchronize(); cudaFree(h->workspace); }
        h->workspace = nullptr; h->workspace_cap = 0;
        if (cudaMalloc(&h->workspace, ws_total * sizeof(double)) != cudaSuccess) { cudaGetLastError(); return fail(h, LCQP_CUDA_OUT_OF_MEMORY, "cudaMalloc(workspace)"); }
        h->workspace_cap = ws_total;
    }
    a.workspace = h->workspace;
    a.xout = h->xout; a.yout = h->yout; a.stats = h->stats; a.counter = h->counter;
    a.instance_offset = h->instance_offset;
    CK(cudaMemsetAsync(h->counter, 0, sizeof(unsigned int), stream), LCQP_CUDA_LAUNCH_FAILED);
    CK(cudaEventRecord(h->ev0, stream), LCQP_CUDA_LAUNCH_FAILED);
    CK(cudaEventRecord(h->ev1, stream), LCQP_CUDA_LAUNCH_FAILED);
    if (getenv("LCQP_CUDA_VERBOSE"))
        fprintf(stderr, "lcqp_cuda (OSQP flavour, %s): %lld warps (%d per SM), N %d, nnz(L) %d, nnz(K) %d, levels %d + %d, workspace %.2f MB/warp, smem %zu B/warp, factor flops %lld, streamed sweeps %s (%d + %d chunks)\n",
                warp_mode ? "one warp per instance" : "one thread per instance", warps, per_sm, a.S.N, a.S.nnzL, a.S.nnzK, a.S.nflev, a.S.nblev, per_warp / 1.0e6,
                dyn_smem, h->sym ? h->sym->factor_flops : 0ll, a.S.stream ? "on" : "off", a.S.fsChunks, a.S.bsChunks);
    h->osqp_warp_mode = warp_mode ? 1 : 0;
    if (warp_mode) lcqp_osqpw_kernel<<<(unsigned)warps, 32, dyn_smem, stream>>>(a);
    else lcqp_osqp_kernel<<<(unsigned)warps, 32, dyn_smem, stream>>>(a);
    h->launches++;
    CK(cudaGetLastError(), LCQP_CUDA_LAUNCH_FAILED);
    CK(cudaEventRecord(h->ev2, stream), LCQP_CUDA_LAUNCH_FAILED);
#ifdef LCQP_PROFILE
    if (getenv("LCQP_CUDA_VERBOSE") && warp_mode) {
        cudaStreamSynchronize(stream);
        unsigned long long pr[16];
        cudaMemcpyFromSymbol(pr, g_prof, sizeof(pr));
        unsigned long long z[16] = {};
        cudaMemcpyToSymbol(g_prof, z, sizeof(z));
        double tot = 0;
        for (int k = 0; k < 14; k++) tot += (double)pr[k];
        static const char* nm[14] = {"outer loop / other", "factor: elimination", "factor: pack streams", "solve: permute/scale", "solve: forward sweep", "solve: backward sweep",
                                     "admm: vector updates", "admm: residuals/term.", "polish (its own part)", "sweeps: waiting for a chunk", "-", "-", "-", "-"};
        fprintf(stderr, "lcqp_cuda OSQP-flavour profile (warp cycles per LCQP, %% of total):\n");
        for (int k = 0; k < 10; k++) fprintf(stderr, "   %-28s %14.0f  %5.1f%%\n", nm[k], (double)pr[k] / h->batch, 100.0 * pr[k] / (tot > 0 ? tot : 1));
        if (h->sym && h->sym->stream) {
            long long lf = 0, lb = 0;
            for (int c = 0; c < h->sym->fsChunks; c++) lf += (unsigned short)(h->sym->fsI[(size_t)c * osq::kStreamIdx / 2] & 0xffff);
            for (int c = 0; c < h->sym->bsChunks; c++) lb += (unsigned short)(h->sym->bsI[(size_t)c * osq::kStreamIdx / 2] & 0xffff);
            fprintf(stderr, "   stream levels: forward %lld, backward %lld\n", lf, lb);
        }
    }
#endif
    h->last_stream = stream;
    h->last_grid = (int)warps;
    h->last_smem = (int)dyn_smem;
    h->last_mE = -1;
    h->ran = true;
    return LCQP_CUDA_OK;
}

static int run_pas(lcqp_cuda_handle h, cudaStream_t stream)
{
    PasArgs a;
    pas_fill_args(h, a);
    const pas::PDims& d = a.d;
    if (a.mats_shared) { a.mEc = h->pas_host->mE; a.mIc = h->pas_host->mI; a.capc = pas::pas_cap(d, a.mEc, a.mIc); }
    else { a.mEc = pas::pas_mEmax(d); a.mIc = d.m; a.capc = d.n < d.m ? d.n : d.m; }
    h->last_mE = a.mats_shared ? a.mEc : -1;
    // a CTA is G groups of T threads; every group keeps the row vectors of its instance in shared memory
    int threads = 128;
    if (const char* t = tune_env("LCQP_CUDA_THREADS")) { const int v = atoi(t); if (v >= 32 && v <= 256 && v % 32 == 0) threads = v; }
    int gmax = kPasCtaThreads / threads;
    if (gmax > kPasGroupsMax) gmax = kPasGroupsMax;
    if (const char* t = tune_env("LCQP_CUDA_GROUPS")) { const int v = atoi(t); if (v >= 1 && v < gmax) gmax = v; }
    if (gmax > h->batch) gmax = h->batch;
    cudaFuncAttributes fattr;
    CK(cudaFuncGetAttributes(&fattr, lcqp_pas_kernel), LCQP_CUDA_LAUNCH_FAILED);
    const size_t budget = kSmemMax - fattr.sharedSizeBytes;
    a.group_smem = (pas::pas_smem_bytes(d, a.mEc, a.mIc, a.capc) + 15) / 16 * 16;
    int groups = gmax;
    while (groups > 1 && (size_t)groups * a.group_smem > budget) groups--;
    if ((size_t)groups * a.group_smem > budget) return fail(h, LCQP_CUDA_TOO_LARGE, "instance does not fit the shared-memory budget");
    const size_t smem = (size_t)groups * a.group_smem;
    CK(cudaFuncSetAttribute(lcqp_pas_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), LCQP_CUDA_LAUNCH_FAILED);
    int per_sm = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lcqp_pas_kernel, threads * groups, smem), LCQP_CUDA_LAUNCH_FAILED);
    if (per_sm < 1) return fail(h, LCQP_CUDA_TOO_LARGE, "kernel cannot be resident");
    int grid = per_sm * h->num_sms;
    if ((long long)grid * groups > h->batch) grid = (h->batch + groups - 1) / groups;
    a.ws_mats = a.mats_shared ? 0 : pas::pmats_doubles(d);
    a.ws_stride = a.ws_mats + pas::pas_gl_doubles(d, a.mEc, a.capc) + 16;
    a.ws_stride = (a.ws_stride + 1) & ~1ull;
    const size_t ws_total = a.ws_stride * (size_t)grid * groups;
    if (ws_total > h->workspace_cap) {
        // (the previous run on another stream may still use the old block: wait for it)
        if (h->workspace) { cudaDeviceSynchronize(); cudaFree(h->workspace); }
        h->workspace = nullptr; h->workspace_cap = 0;
        if (cudaMalloc(&h->workspace, (ws_total ? ws_total : 1) * sizeof(double)) != cudaSuccess) { cudaGetLastError(); return fail(h, LCQP_CUDA_OUT_OF_MEMORY, "cudaMalloc(workspace)"); }
        h->workspace_cap = ws_total;
    }
    a.workspace = h->workspace;
    CK(cudaMemsetAsync(h->counter, 0, sizeof(unsigned int), stream), LCQP_CUDA_LAUNCH_FAILED);
    CK(cudaMemsetAsync(h->fallback_flag, 0, sizeof(int), stream), LCQP_CUDA_LAUNCH_FAILED);
    CK(cudaMemsetAsync(h->work, 0, 2 * sizeof(unsigned long long), stream), LCQP_CUDA_LAUNCH_FAILED);
    CK(cudaEventRecord(h->ev0, stream), LCQP_CUDA_LAUNCH_FAILED);
    CK(cudaEventRecord(h->ev1, stream), LCQP_CUDA_LAUNCH_FAILED);
    if (getenv("LCQP_CUDA_VERBOSE"))
        fprintf(stderr, "lcqp_cuda (parametric active set): grid %d x (%d threads x %d groups), %d CTA/SM, smem %zu B = %d x %llu, mE %d mI %d cap %d, scratch %llu doubles/group\n",
                grid, threads, groups, per_sm, smem, groups, a.group_smem, a.mEc, a.mIc, a.capc, a.ws_stride);
    lcqp_pas_kernel<<<grid, dim3(threads, groups), smem, stream>>>(a);
    h->launches++;
    CK(cudaGetLastError(), LCQP_CUDA_LAUNCH_FAILED);
    CK(cudaEventRecord(h->ev2, stream), LCQP_CUDA_LAUNCH_FAILED);
#ifdef LCQP_PROFILE
    if (getenv("LCQP_CUDA_VERBOSE")) {
        cudaStreamSynchronize(stream);
        unsigned long long pr[16];
        cudaMemcpyFromSymbol(pr, g_prof, sizeof(pr));
        unsigned long long z[16] = {};
        cudaMemcpyToSymbol(g_prof, z, sizeof(z));
        double tot = 0;
        for (int k = 0; k < 16; k++) tot += (double)pr[k];
        static const char* nm[16] = {"dy = Sinv rhs", "dz = Tt[:,W] dy", "ratio tests", "step + hom.len", "ws remove", "ws add", "drift", "ramping", "dc / target", "finish: rebase", "finish: polish", "outer loop", "init (aux QP)", "store stats", "hotstart misc", "-"};
        fprintf(stderr, "lcqp_cuda profile (group cycles per LCQP, %% of total):\n");
        for (int k = 0; k < 16; k++) fprintf(stderr, "   %-16s %12.0f  %5.1f%%\n", nm[k], (double)pr[k] / h->batch, 100.0 * pr[k] / (tot > 0 ? tot : 1));
    }
#endif
    h->last_stream = stream;
    h->last_grid = grid * groups;
    h->last_smem = (int)smem;
    h->ran = true;
    if (!a.mats_shared) {
        // per-instance matrices: an instance with a semidefinite reduced Hessian sends the batch to the regularised solver
        CK(cudaMemcpyAsync(h->fallback_host, h->fallback_flag, sizeof(int), cudaMemcpyDeviceToHost, stream), LCQP_CUDA_LAUNCH_FAILED);
        CK(cudaStreamSynchronize(stream), LCQP_CUDA_LAUNCH_FAILED);
        if (*h->fallback_host) { h->use_legacy = true; return run_legacy(h, stream); }
    }
    return LCQP_CUDA_OK;
}

int lcqp_cuda_run(lcqp_cuda_handle h, void* stream_v)
{
    if (!h) return LCQP_CUDA_BAD_HANDLE;
    if (!h->loaded) return fail(h, LCQP_CUDA_NOT_LOADED, "run before load");
    cudaStream_t stream = (cudaStream_t)stream_v;
    CK(cudaSetDevice(h->device), LCQP_CUDA_NO_DEVICE);
    if (h->opts.qpSolver == 2 && h->opts.osqp_admm) return run_osqp(h, stream);
    if (h->osqp_ready && !h->pas_ready) return fail(h, LCQP_CUDA_NOT_LOADED, "osqp_admm was cleared after the load: load again");
    if (h->use_legacy || !h->pas_ready || tune_env("LCQP_CUDA_LEGACY")) return run_legacy(h, stream);
    return run_pas(h, stream);
}

int lcqp_cuda_synchronize(lcqp_cuda_handle h)
{
    if (!h) return LCQP_CUDA_BAD_HANDLE;
    CK(cudaSetDevice(h->device), LCQP_CUDA_NO_DEVICE);
    CK(cudaStreamSynchronize(h->last_stream), LCQP_CUDA_LAUNCH_FAILED);
    return LCQP_CUDA_OK;
}

static int get_common(lcqp_cuda_handle h, void* dst, const void* src, size_t bytes)
{
    if (!h) return LCQP_CUDA_BAD_HANDLE;
    if (!h->ran) return fail(h, LCQP_CUDA_NOT_RUN, "get before run");
    if (!dst) return LCQP_CUDA_BAD_ARGUMENT;
    CK(cudaSetDevice(h->device), LCQP_CUDA_NO_DEVICE);
    CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, h->last_stream), LCQP_CUDA_LAUNCH_FAILED);
    CK(cudaStreamSynchronize(h->last_stream), LCQP_CUDA_LAUNCH_FAILED);
    return LCQP_CUDA_OK;
}

int lcqp_cuda_get_primal(lcqp_cuda_handle h, double* x)
{
    return h ? get_common(h, x, h->xout, sizeof(double) * h->nV * (size_t)h->batch) : LCQP_CUDA_BAD_HANDLE;
}

int lcqp_cuda_get_dual(lcqp_cuda_handle h, double* y)
{
    return h ? get_common(h, y, h->yout, sizeof(double) * ((size_t)h->nV + h->nC + 2 * (size_t)h->nComp) * (size_t)h->batch) : LCQP_CUDA_BAD_HANDLE;
}

int lcqp_cuda_get_stats(lcqp_cuda_handle h, lcqp_cuda_stats* st)
{
    return h ? get_common(h, st, h->stats, sizeof(lcqp_cuda_stats) * (size_t)h->batch) : LCQP_CUDA_BAD_HANDLE;
}

int lcqp_cuda_get_device_results(lcqp_cuda_handle h, const double** x, const double** y, const lcqp_cuda_stats** st)
{
    if (!h) return LCQP_CUDA_BAD_HANDLE;
    if (!h->ran) return fail(h, LCQP_CUDA_NOT_RUN, "get before run");
    if (x) *x = h->xout;
    if (y) *y = h->yout;
    if (st) *st = h->stats;
    return LCQP_CUDA_OK;
}

int lcqp_cuda_num_duals(lcqp_cuda_handle h)
{
    if (!h) return -1;
    const int mA = h->nC + 2 * h->nComp;
    return h->opts.qpSolver == 2 ? mA : h->nV + mA;  // LCQProblem.cpp:889, :934
}

long long lcqp_cuda_launch_count(lcqp_cuda_handle h) { return h ? h->launches : -1; }

int lcqp_cuda_last_run_ms(lcqp_cuda_handle h, float* solve_ms, float* total_ms)
{
    if (!h) return LCQP_CUDA_BAD_HANDLE;
    if (!h->ran) return fail(h, LCQP_CUDA_NOT_RUN, "timing before run");
    CK(cudaSetDevice(h->device), LCQP_CUDA_NO_DEVICE);
    CK(cudaEventSynchronize(h->ev2), LCQP_CUDA_LAUNCH_FAILED);
    if (solve_ms) CK(cudaEventElapsedTime(solve_ms, h->ev1, h->ev2), LCQP_CUDA_LAUNCH_FAILED);
    if (total_ms) CK(cudaEventElapsedTime(total_ms, h->ev0, h->ev2), LCQP_CUDA_LAUNCH_FAILED);
    return LCQP_CUDA_OK;
}

int lcqp_cuda_last_launch_info(lcqp_cuda_handle h, int* grid, int* smem_bytes, int* equality_rows)
{
    if (!h) return LCQP_CUDA_BAD_HANDLE;
    if (!h->ran) return fail(h, LCQP_CUDA_NOT_RUN, "launch info before run");
    if (grid) *grid = h->last_grid;
    if (smem_bytes) *smem_bytes = h->last_smem;
    if (equality_rows) *equality_rows = h->last_mE;
    return LCQP_CUDA_OK;
}

const char* lcqp_cuda_last_error(lcqp_cuda_handle h) { return h ? h->err.c_str() : "bad handle"; }

int lcqp_cuda_osqp_info(lcqp_cuda_handle h, int* N, int* nnzL, int* levels, int* mode)
{
    if (!h) return LCQP_CUDA_BAD_HANDLE;
    if (!h->osqp_ready) return fail(h, LCQP_CUDA_NOT_LOADED, "no OSQP-flavour load");
    if (N) *N = h->symdev.N;
    if (nnzL) *nnzL = h->symdev.nnzL;
    if (levels) *levels = h->symdev.nflev + h->symdev.nblev;
    if (mode) *mode = h->osqp_warp_mode;
    return LCQP_CUDA_OK;
}

int lcqp_cuda_last_work(lcqp_cuda_handle h, double* fp64_macs, double* bytes)
{
    if (!h) return LCQP_CUDA_BAD_HANDLE;
#ifndef LCQP_COUNT_WORK
    return fail(h, LCQP_CUDA_BAD_ARGUMENT, "this build does not count (the counting build is liblcqp_cuda_work.so, -DLCQP_COUNT_WORK)");
#endif
    if (!h->ran || h->use_legacy || !h->pas_ready || (h->opts.qpSolver == 2 && h->opts.osqp_admm))
        return fail(h, LCQP_CUDA_NOT_LOADED, "the work counters belong to a run of the parametric active-set kernel");
    CK(cudaSetDevice(h->device), LCQP_CUDA_NO_DEVICE);
    CK(cudaStreamSynchronize(h->last_stream), LCQP_CUDA_LAUNCH_FAILED);
    unsigned long long v[2] = {0, 0};
    CK(cudaMemcpy(v, h->work, sizeof(v), cudaMemcpyDeviceToHost), LCQP_CUDA_LAUNCH_FAILED);
    if (fp64_macs) *fp64_macs = (double)v[0];
    if (bytes) *bytes = (double)v[1];
    return LCQP_CUDA_OK;
}

// every thread sums 16-byte loads of a buffer that fits the L2 (volatile: the loads are issued, eight in flight)
__global__ void l2_read_probe_kernel(const double2* buf, size_t n16, int rounds, double* out)
{
    double acc = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int r = 0; r < rounds; r++) {
        size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
        for (; i + 7 * stride < n16; i += 8 * stride) {
            double2 v[8];
#pragma unroll
            for (int k = 0; k < 8; k++) asm volatile("ld.volatile.global.v2.f64 {%0, %1}, [%2];" : "=d"(v[k].x), "=d"(v[k].y) : "l"(buf + i + k * stride));
#pragma unroll
            for (int k = 0; k < 8; k++) acc += v[k].x + v[k].y;
        }
    }
    if (acc == 1.2345e-300) out[0] = acc;
}

int lcqp_cuda_measure_l2_gbs(int device, double* gbs)
{
    if (!gbs) return LCQP_CUDA_BAD_ARGUMENT;
    cudaDeviceProp prop;
    if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&prop, device) != cudaSuccess) return LCQP_CUDA_NO_DEVICE;
    const size_t bytes = 48ull << 20;   // well inside the 126 MB L2 (each half of it holds the buffer)
    const size_t n16 = bytes / 16;
    double2* buf = nullptr;
    double* out = nullptr;
    if (cudaMalloc(&buf, bytes) != cudaSuccess || cudaMalloc(&out, 8) != cudaSuccess) { cudaGetLastError(); cudaFree(buf); return LCQP_CUDA_OUT_OF_MEMORY; }
    cudaMemset(buf, 0, bytes);
    const int blocks = prop.multiProcessorCount * 4, threads = 512, rounds = 256;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    l2_read_probe_kernel<<<blocks, threads>>>(buf, n16, 2, out);   // brings the buffer into the L2
    float best = 1e30f;
    for (int r = 0; r < 3; r++) {
        cudaEventRecord(e0);
        l2_read_probe_kernel<<<blocks, threads>>>(buf, n16, rounds, out);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(buf); cudaFree(out);
    if (cudaGetLastError() != cudaSuccess) return LCQP_CUDA_LAUNCH_FAILED;
    const size_t per_round = (n16 / ((size_t)blocks * threads * 8)) * ((size_t)blocks * threads * 8) * 16;
    *gbs = (double)per_round * rounds / (best * 1e-3) / 1e9;
    return LCQP_CUDA_OK;
}

int lcqp_cuda_measure_fp64_tflops(int device, double* tflops)
{
    if (!tflops) return LCQP_CUDA_BAD_ARGUMENT;
    cudaDeviceProp prop;
    if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&prop, device) != cudaSuccess) return LCQP_CUDA_NO_DEVICE;
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 14;
    double* buf = nullptr;
    if (cudaMalloc(&buf, sizeof(double) * blocks * threads) != cudaSuccess) { cudaGetLastError(); return LCQP_CUDA_OUT_OF_MEMORY; }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    fp64_fma_probe_kernel<<<blocks, threads>>>(buf, iters);
    float best = 1e30f;
    for (int r = 0; r < 3; r++) {
        cudaEventRecord(e0);
        fp64_fma_probe_kernel<<<blocks, threads>>>(buf, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(buf);
    if (cudaGetLastError() != cudaSuccess) return LCQP_CUDA_LAUNCH_FAILED;
    *tflops = 2.0 * 8.0 * (double)iters * blocks * threads / (best * 1e-3) / 1e12;
    return LCQP_CUDA_OK;
}

// ---- plugin door ---------------------------------------------------------------------------------
struct lcqp_cuda_qp_s {
    int nV, nCtot, device;
    lcqp_cuda_options opts;
    double *Q = nullptr, *A = nullptr;
    double* stage = nullptr;  // g n | lbA mA | ubA mA | lb n | ub n | x0 n | y0 n+mA
    double* mats_store = nullptr;
    Mats* mats = nullptr;
    double* gl = nullptr;
    unsigned char* saved = nullptr;
    QPState* state = nullptr;
    double* xout = nullptr;
    double* yout = nullptr;
    int has_box = 0;
    bool initialised = false;
    long long launches = 0;
    // parametric active-set solver (the default); the regularised solver above takes over for semidefinite Hessians
    double* pas_store = nullptr;
    pas::PMats* pas_mats = nullptr;
    double* pas_gl = nullptr;
    unsigned char* pas_saved = nullptr;
    PasQPState* pas_state = nullptr;
    bool use_legacy = false;
    bool pas_has[4] = {false, false, false, false};   // lbA, ubA, lb, ub were given at the initial solve
    std::vector<char> eq_rows;   // rows that were equalities (l == u) at the initial solve: eliminated for good
};

static pas::PDims pas_qp_dims(int nV, int nCtot, int has_box)
{
    pas::PDims d = pas::make_pdims(nV, 0, 0, has_box);
    d.nC = nCtot; d.mA = nCtot;
    d.m = d.mA + (has_box ? nV : 0);
    return d;
}

static Dims qp_dims(int nV, int nCtot, int has_box)
{
    Dims d = make_dims(nV, 0, 0, has_box);
    d.nC = nCtot; d.mA = nCtot;
    d.m = d.mA + (has_box ? d.n : 0);
    d.ldE = ((d.m < d.n ? d.m : d.n) + 1) & ~1;
    d.capE = d.ldE > 0 ? d.ldE : 1;
    d.cap = d.m < d.n + 8 ? d.m : d.n + 8;
    if (d.cap < 1) d.cap = 1;
    return d;
}

int lcqp_cuda_qp_create(int nV, int nCtot, const double* Q, const double* A, int device, lcqp_cuda_qp* out)
{
    if (!out) return LCQP_CUDA_BAD_ARGUMENT;
    *out = nullptr;
    if (nV <= 0 || nCtot < 0 || !Q || (nCtot > 0 && !A)) return LCQP_CUDA_BAD_ARGUMENT;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return LCQP_CUDA_NO_DEVICE;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major != 10) return LCQP_CUDA_NO_DEVICE;
    lcqp_cuda_qp q = new (std::nothrow) lcqp_cuda_qp_s();
    if (!q) return LCQP_CUDA_OUT_OF_MEMORY;
    q->nV = nV; q->nCtot = nCtot; q->device = device;
    lcqp_cuda_default_options(&q->opts);
    cudaSetDevice(device);
    const size_t n = nV, mA = nCtot;
    const Dims d = qp_dims(nV, nCtot, 1);  // the larger of the two layouts
    const SmemPlan plan = make_plan(d, 0);  // everything optional goes to global scratch
    bool ok = cudaMalloc(&q->Q, n * n * 8) == cudaSuccess && cudaMalloc(&q->A, (mA * n + 1) * 8) == cudaSuccess &&
              cudaMalloc(&q->stage, (5 * n + 3 * mA + 8) * 8) == cudaSuccess &&
              cudaMalloc(&q->mats_store, mats_doubles(d) * 8) == cudaSuccess &&
              cudaMalloc(&q->mats, sizeof(Mats)) == cudaSuccess &&
              cudaMalloc(&q->gl, (plan.gl_doubles + 1) * 8) == cudaSuccess &&
              cudaMalloc(&q->saved, plan.bytes + 64) == cudaSuccess &&
              cudaMalloc(&q->state, sizeof(QPState)) == cudaSuccess &&
              cudaMalloc(&q->xout, n * 8) == cudaSuccess && cudaMalloc(&q->yout, (n + mA) * 8) == cudaSuccess;
    {
        const pas::PDims pd = pas_qp_dims(nV, nCtot, 1);
        const int mEc = pas::pas_mEmax(pd), capc = pd.n < pd.m ? pd.n : pd.m;
        ok = ok && cudaMalloc(&q->pas_store, pas::pmats_doubles(pd) * 8) == cudaSuccess &&
             cudaMalloc(&q->pas_mats, sizeof(pas::PMats)) == cudaSuccess &&
             cudaMalloc(&q->pas_gl, (pas::pas_gl_doubles(pd, mEc, capc) + 16) * 8) == cudaSuccess &&
             cudaMalloc(&q->pas_saved, pas::pas_smem_bytes(pd, mEc, pd.m, capc) + 64) == cudaSuccess &&
             cudaMalloc(&q->pas_state, sizeof(PasQPState)) == cudaSuccess;
        if (ok) ok = cudaMemset(q->pas_state, 0, sizeof(PasQPState)) == cudaSuccess;
    }
    if (ok) ok = cudaMemcpy(q->Q, Q, n * n * 8, cudaMemcpyHostToDevice) == cudaSuccess;
    if (ok && mA) ok = cudaMemcpy(q->A, A, mA * n * 8, cudaMemcpyHostToDevice) == cudaSuccess;
    if (ok) ok = cudaMemset(q->state, 0, sizeof(QPState)) == cudaSuccess;
    if (!ok) { cudaGetLastError(); lcqp_cuda_qp_destroy(q); return LCQP_CUDA_OUT_OF_MEMORY; }
    *out = q;
    return LCQP_CUDA_OK;
}

int lcqp_cuda_qp_destroy(lcqp_cuda_qp q)
{
    if (!q) return LCQP_CUDA_BAD_HANDLE;
    cudaSetDevice(q->device);
    cudaFree(q->Q); cudaFree(q->A); cudaFree(q->stage); cudaFree(q->mats_store); cudaFree(q->mats); cudaFree(q->gl);
    cudaFree(q->saved); cudaFree(q->state); cudaFree(q->xout); cudaFree(q->yout);
    cudaFree(q->pas_store); cudaFree(q->pas_mats); cudaFree(q->pas_gl); cudaFree(q->pas_saved); cudaFree(q->pas_state);
    delete q;
    return LCQP_CUDA_OK;
}

int lcqp_cuda_qp_set_options(lcqp_cuda_qp q, const lcqp_cuda_options* o)
{
    if (!q) return LCQP_CUDA_BAD_HANDLE;
    if (!o) return LCQP_CUDA_BAD_ARGUMENT;
    if (!qp_knobs_valid(o)) return LCQP_CUDA_BAD_ARGUMENT;
    q->opts = *o;
    return LCQP_CUDA_OK;
}

int lcqp_cuda_qp_solve(lcqp_cuda_qp q, int initialSolve, int* iterations, int* exit_flag,
                       const double* g, const double* lbA, const double* ubA,
                       const double* x0, const double* y0, const double* lb, const double* ub)
{
    if (!q) return LCQP_CUDA_BAD_HANDLE;
    if (!g || !iterations || !exit_flag) return LCQP_CUDA_BAD_ARGUMENT;
    if (!initialSolve && !q->initialised) return LCQP_CUDA_BAD_ARGUMENT;
    if (cudaSetDevice(q->device) != cudaSuccess) return LCQP_CUDA_NO_DEVICE;
    const size_t n = q->nV, mA = q->nCtot;
    if (initialSolve) q->has_box = (lb || ub) ? 1 : 0;
    QPKernelArgs a;
    memset(&a, 0, sizeof(a));
    a.d = qp_dims(q->nV, q->nCtot, q->has_box);
    a.o = q->opts;
    double* st = q->stage;
    double* d_g = st; st += n;
    double* d_lbA = st; st += mA;
    double* d_ubA = st; st += mA;
    double* d_lb = st; st += n;
    double* d_ub = st; st += n;
    double* d_x0 = st; st += n;
    double* d_y0 = st;
    bool ok = cudaMemcpy(d_g, g, n * 8, cudaMemcpyHostToDevice) == cudaSuccess;
    auto up = [&](double* dst, const double* src, size_t cnt) { if (src && cnt) ok = ok && cudaMemcpy(dst, src, cnt * 8, cudaMemcpyHostToDevice) == cudaSuccess; };
    if (initialSolve) {
        up(d_lbA, lbA, mA); up(d_ubA, ubA, mA); up(d_lb, lb, n); up(d_ub, ub, n); up(d_x0, x0, n); up(d_y0, y0, n + mA);
        q->use_legacy = false;
        q->eq_rows.assign(mA + n, 0);
        for (size_t i = 0; i < mA; i++) q->eq_rows[i] = (lbA && ubA && lbA[i] == ubA[i] && lbA[i] > -1e20 && lbA[i] < 1e20);
        for (size_t j = 0; j < n; j++) q->eq_rows[mA + j] = (lb && ub && lb[j] == ub[j] && lb[j] > -1e20 && lb[j] < 1e20);
    } else if (!q->use_legacy) {
        // SubsolverQPOASES forwards the bounds of every call to qp.hotstart (SubsolverQPOASES.cpp:156-158): a hot start
        // follows the homotopy to the new bounds as well.  The rows eliminated at the initial solve (l == u) must stay
        // equalities; a NULL pointer keeps the bounds of the previous call.
        for (size_t i = 0; i < mA && lbA && ubA; i++) if (q->eq_rows[i] && lbA[i] != ubA[i]) return LCQP_CUDA_BAD_ARGUMENT;
        for (size_t j = 0; j < n && lb && ub && q->has_box; j++) if (q->eq_rows[mA + j] && lb[j] != ub[j]) return LCQP_CUDA_BAD_ARGUMENT;
        up(d_lbA, lbA, mA); up(d_ubA, ubA, mA);
        if (q->has_box) { up(d_lb, lb, n); up(d_ub, ub, n); }
    }
    if (!ok) { cudaGetLastError(); return LCQP_CUDA_LAUNCH_FAILED; }
    if (!q->use_legacy) {
        PasQPArgs pa;
        memset(&pa, 0, sizeof(pa));
        pa.d = pas_qp_dims(q->nV, q->nCtot, q->has_box);
        pa.o = q->opts;
        pa.in.Q = q->Q; pa.in.A = q->A; pa.in.g = d_g;
        // the staging copies of the bounds persist between the calls (a pointer that was NULL at the initial solve stays absent)
        static_assert(sizeof(bool) == 1, "");
        if (initialSolve) { q->initialised = false; }
        pa.in.lbA = (initialSolve ? lbA != nullptr : q->pas_has[0]) ? d_lbA : nullptr;
        pa.in.ubA = (initialSolve ? ubA != nullptr : q->pas_has[1]) ? d_ubA : nullptr;
        pa.in.lb = (initialSolve ? lb != nullptr : q->pas_has[2]) ? d_lb : nullptr;
        pa.in.ub = (initialSolve ? ub != nullptr : q->pas_has[3]) ? d_ub : nullptr;
        if (initialSolve) { q->pas_has[0] = lbA != nullptr; q->pas_has[1] = ubA != nullptr; q->pas_has[2] = lb != nullptr; q->pas_has[3] = ub != nullptr; }
        pa.in.x0 = (initialSolve && x0) ? d_x0 : nullptr;
        pa.in.y0 = (initialSolve && y0) ? d_y0 : nullptr;
        pa.mEc = pas::pas_mEmax(pa.d); pa.mIc = pa.d.m; pa.capc = pa.d.n < pa.d.m ? pa.d.n : pa.d.m;
        pa.mats_store = q->pas_store; pa.mats = q->pas_mats; pa.gl = q->pas_gl; pa.saved_smem = q->pas_saved;
        pa.smem_bytes = (pas::pas_smem_bytes(pa.d, pa.mEc, pa.mIc, pa.capc) + 7) / 8 * 8;
        pa.state = q->pas_state; pa.xout = q->xout; pa.yout = q->yout; pa.initial = initialSolve ? 1 : 0;
        if (pa.smem_bytes > kSmemMax) return LCQP_CUDA_TOO_LARGE;
        if (cudaFuncSetAttribute(pas_plugin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pa.smem_bytes) != cudaSuccess) { cudaGetLastError(); return LCQP_CUDA_LAUNCH_FAILED; }
        pas_plugin_kernel<<<1, kThreads, pa.smem_bytes>>>(pa);
        q->launches++;
        if (cudaGetLastError() != cudaSuccess) return LCQP_CUDA_LAUNCH_FAILED;
        PasQPState hs;
        if (cudaMemcpy(&hs, q->pas_state, sizeof(hs), cudaMemcpyDeviceToHost) != cudaSuccess) { cudaGetLastError(); return LCQP_CUDA_LAUNCH_FAILED; }
        if (!(initialSolve && hs.prepared == 2)) {   // prepared == 1 + status; status 1: semidefinite Hessian -> the regularised solver
            q->initialised = true;
            *iterations = hs.iterations;
            *exit_flag = hs.flag;
            return hs.flag == 0 ? 0 : 203;  // SUBPROBLEM_SOLVER_ERROR (SubsolverQPOASES.cpp:165-166)
        }
        q->use_legacy = true;
    }
    memset(&a.in, 0, sizeof(a.in));
    a.in.Q = q->Q; a.in.A = q->A; a.in.g = d_g;
    // the QP's bounds are those given at the initial solve (LCQPow never changes them between calls)
    if (initialSolve) { a.in.lbA = lbA ? d_lbA : nullptr; a.in.ubA = ubA ? d_ubA : nullptr; a.in.lb = lb ? d_lb : nullptr; a.in.ub = ub ? d_ub : nullptr;
                        a.in.x0 = x0 ? d_x0 : nullptr; a.in.y0 = y0 ? d_y0 : nullptr; }
    a.plan = make_plan(a.d, 0);
    if (a.plan.bytes > kSmemMax) return LCQP_CUDA_TOO_LARGE;
    a.mats_store = q->mats_store;
    a.mats = q->mats;
    a.gl = q->gl;
    a.saved_smem = q->saved;
    a.smem_bytes = (a.plan.bytes + 7) / 8 * 8;
    a.state = q->state;
    a.xout = q->xout; a.yout = q->yout;
    a.initial = initialSolve ? 1 : 0;
    if (cudaFuncSetAttribute(qp_plugin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)a.smem_bytes) != cudaSuccess) { cudaGetLastError(); return LCQP_CUDA_LAUNCH_FAILED; }
    qp_plugin_kernel<<<1, kThreads, a.smem_bytes>>>(a);
    q->launches++;
    if (cudaGetLastError() != cudaSuccess) return LCQP_CUDA_LAUNCH_FAILED;
    QPState hs;
    if (cudaMemcpy(&hs, q->state, sizeof(hs), cudaMemcpyDeviceToHost) != cudaSuccess) { cudaGetLastError(); return LCQP_CUDA_LAUNCH_FAILED; }
    q->initialised = true;
    *iterations = hs.iterations;
    *exit_flag = hs.flag;
    return hs.flag == 0 ? 0 : 203;  // SUBPROBLEM_SOLVER_ERROR (SubsolverQPOASES.cpp:165-166)
}

int lcqp_cuda_qp_get_solution(lcqp_cuda_qp q, double* x, double* y)
{
    if (!q) return LCQP_CUDA_BAD_HANDLE;
    if (!q->initialised) return LCQP_CUDA_NOT_RUN;
    if (cudaSetDevice(q->device) != cudaSuccess) return LCQP_CUDA_NO_DEVICE;
    bool ok = true;
    if (x) ok = ok && cudaMemcpy(x, q->xout, (size_t)q->nV * 8, cudaMemcpyDeviceToHost) == cudaSuccess;
    if (y) ok = ok && cudaMemcpy(y, q->yout, ((size_t)q->nV + q->nCtot) * 8, cudaMemcpyDeviceToHost) == cudaSuccess;
    return ok ? LCQP_CUDA_OK : LCQP_CUDA_LAUNCH_FAILED;
}

}  // extern "C"

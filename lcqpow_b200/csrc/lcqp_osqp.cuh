// lcqp_osqp.cuh -- the OSQP flavour of the B200-native LCQP path (sm_100a device code; compiles as plain C++ under
// LCQP_HOST_EMU for the CPU test-suite).
//
// What it restates.  With QPSolver::OSQP_SPARSE the reference hands every inner QP to OSQP 0.6.2
// (/root/reference/src/SubsolverOSQP.cpp:124-200): Ruiz equilibration with cost scaling (external/osqp/src/scaling.c:44-156),
// the ADMM iteration on the quasi-definite KKT system (osqp.c:354-519, auxil.c:161-225, proj.c:4-14), a sparse LDL'
// factorisation of K = [P + sigma I, A'; A, -diag(1/rho)] (kkt.c:6-177, qdldl.c:72-281), residuals, termination and
// infeasibility certificates (auxil.c:240-520, :681-786), rho adaptation with refactorisation (auxil.c:13-74,
// osqp.c:1268-1318), polish with iterative refinement and its accept rule (polish.c:19-350), warm start from the
// previous QP (osqp.c:752-832, :954-994), and the adapter's conventions (exit flag, negated duals).  Each function
// below cites the lines it follows; the arithmetic is fp64 and in the reference's order of operations.
//
// How it is built for the GPU.  ONE THREAD OWNS ONE INSTANCE.  A batch shares its sparsity pattern, so the symbolic
// analysis (ordering, elimination tree, pattern of L, the row programs of the factorisation: lcqp_sparse_host.hpp)
// is done once on the host and every thread executes the SAME index sequence on its own values.  All per-instance
// arrays of the 32 instances of a warp are interleaved (element i of lane l at [i * 32 + l]): every load and store
// of the hot loops -- L in the triangular solves, the iterates, the matrix values -- is one fully coalesced 256-byte
// transaction per warp, the index arrays are warp-uniform broadcast loads, and there is no shared memory, no
// barrier and no reduction tree anywhere (norms are per-thread running maxima).  Instances that need more ADMM
// iterations or more penalty updates than their warp mates simply run longer; warps pull tiles of 32 instances
// from a global counter.  The path is bound by HBM/L2 streaming of L (16 nnz(L) + 8 N + 8 (3n + 5m)
// bytes per instance-iteration).
//
// Code size matters more than inlining here: a thread runs long dependent chains, the warps of an SM are few, and an
// ADMM iteration whose instructions do not stay in the instruction cache is bound by instruction fetch (the first
// version, everything inlined, was 90 k instructions and ran at ~600 cycles per matrix entry).  Every routine below is
// therefore ONE non-inlined function; only the entry loops are unrolled (eight loads ahead of the dependent chain).
//
// Deviations from the reference, stated: (1) the fill-reducing ordering is minimum degree, not AMD -- L differs in
// pattern, the computed iterates differ by round-off only; (2) `adaptive_rho_interval = 0` means "every
// 4 x check_termination iterations" (OSQP's own rule when it is built without its wall-clock profiler) -- the
// reference derives the interval from timings, which makes its own iteration counts machine dependent
// (osqp.c:459-485); (3) a user-supplied y0 warm-starts the duals of [A; L; R] (the reference copies nDuals BYTES of
// them, /root/reference/src/LCQProblem.cpp:947).
#pragma once

#include "lcqp_device.cuh"

namespace lcqp {
namespace osq {

#ifdef LCQP_HOST_EMU
#define OSQ_STRIDE 1
#define OSQ_ANY(mask, pred) (pred)
#define OSQ_BALLOT(pred) ((pred) ? 1u : 0u)
#define OSQ_SYNCWARP(mask) ((void)0)
#else
#define OSQ_STRIDE 32
// The 32 instances of a warp must stay in LOCKSTEP: the interleaved layout is only coalesced, and the warp only issues
// one instruction stream, while the lanes execute the same instruction.  Data-dependent loops (ADMM iterations of a QP,
// passes of the penalty loop) therefore run a WARP-UNIFORM number of times -- until no lane needs another one -- with the
// lanes that are done masked out, and the lanes re-converge explicitly after every trip.  (With plain `break`s the
// lanes drifted apart after the first few QPs and ran one after the other: 20 x slower, measured.)
#define OSQ_ANY(mask, pred) __any_sync(mask, pred)
#define OSQ_BALLOT(pred) __ballot_sync(0xffffffffu, pred)
#define OSQ_SYNCWARP(mask) __syncwarp(mask)
#endif

// constants.h:59-114
constexpr double kRhoMin = 1e-6, kRhoMax = 1e6, kRhoEqOverIneq = 1e3, kRhoTol = 1e-4;
constexpr double kMinScaling = 1e-4, kMaxScaling = 1e4;
constexpr double kOsqpInfty = 1e30;
constexpr double kDivTol = 1.0 / 1e30;
enum { OSQP_SOLVED = 1, OSQP_SOLVED_INACCURATE = 2, OSQP_PRIMAL_INFEASIBLE_INACCURATE = 3, OSQP_DUAL_INFEASIBLE_INACCURATE = 4,
       OSQP_MAX_ITER_REACHED = -2, OSQP_PRIMAL_INFEASIBLE = -3, OSQP_DUAL_INFEASIBLE = -4, OSQP_NON_CVX = -7, OSQP_UNSOLVED = -10 };

struct Vec {
    double* p;
    LCQ_HD double& operator[](int i) const { return p[(size_t)i * OSQ_STRIDE]; }
};

// device view of osq::Symbolic (lcqp_sparse_host.hpp)
struct SymDev {
    int n, m, N, nC, nComp, nnzP, nnzA, nnzQ, nnzK, nnzL;
    const int *Pp, *Pi, *Psrc, *Ap, *Ai, *Asrc, *Qp, *Qi, *Qsrc, *perm, *Kp, *Ki, *Ksrc, *Lp, *Li, *rp, *rcol, *rpos, *Lcol, *Lrev, *Pcol, *Acol, *Qcol;
};

// one instance's inputs (value arrays as the caller laid them out; NULL = absent)
struct View {
    const double *Q, *A, *L, *R, *g, *lbL, *ubL, *lbR, *ubR, *lbA, *ubA, *x0, *y0;
};

struct Work {
    Vec Px, Ax, Lx, Dinv, Lxp, Dinvp;                          // scaled P (upper), scaled A, the ADMM factor, the polish factor
    Vec sD, sDi, q, x, xp, dx, Pxv, Aty, px, xk, pk, gk, gt, gphi, stat, tn1, tn2;   // n
    Vec sE, sEi, l, u, z, zp, y, dy, Axv, rv, riv, pz, py, act, yk, tm1, tm2;       // m
    Vec xt, bp, w3;                                            // N
    Vec sm;                                                    // N: scratch of the factorisation and of the triangular solves -- shared memory when it fits
};

LCQ_HD inline size_t sm_len(const SymDev& S) { return (size_t)(S.N > 2 * S.n ? S.N : 2 * S.n); }   // solve / factor scratch, two accumulators of P v

LCQ_HD inline size_t ws_doubles(const SymDev& S)
{
    return (size_t)S.nnzP + S.nnzA + 2 * (size_t)S.nnzL + 2 * (size_t)S.N + 17 * (size_t)S.n + 17 * (size_t)S.m + 2 * (size_t)S.N + sm_len(S);
}

// `base` points at element 0 of this lane (tile base + lane); consecutive vectors follow each other
LCQ_HD inline void carve(Work& w, const SymDev& S, double* base, double* smem_lane = nullptr)
{
    size_t off = 0;
    auto take = [&](size_t k) { Vec v; v.p = base + off * OSQ_STRIDE; off += k; return v; };
    w.Px = take(S.nnzP); w.Ax = take(S.nnzA); w.Lx = take(S.nnzL); w.Dinv = take(S.N); w.Lxp = take(S.nnzL); w.Dinvp = take(S.N);
    Vec* nv[17] = {&w.sD, &w.sDi, &w.q, &w.x, &w.xp, &w.dx, &w.Pxv, &w.Aty, &w.px, &w.xk, &w.pk, &w.gk, &w.gt, &w.gphi, &w.stat, &w.tn1, &w.tn2};
    for (int k = 0; k < 17; k++) *nv[k] = take(S.n);
    Vec* mv[17] = {&w.sE, &w.sEi, &w.l, &w.u, &w.z, &w.zp, &w.y, &w.dy, &w.Axv, &w.rv, &w.riv, &w.pz, &w.py, &w.act, &w.yk, &w.tm1, &w.tm2};
    for (int k = 0; k < 17; k++) *mv[k] = take(S.m);
    w.xt = take(S.N); w.bp = take(sm_len(S)); w.w3 = take(S.N);
    w.sm = w.bp;
    if (smem_lane) w.sm.p = smem_lane;
}

struct State {
    double rho, c, cinv;
    double pri_res, dua_res;
    int status_val, iter, interval;
    int factor_bad;          // a zero pivot, or not exactly n positive pivots (qdldl_interface.c:93-99)
    long long admm_total, factor_count;
};

LCQ_DEV double a_val(const SymDev& S, const View& v, int p)
{
    const int src = S.Asrc[p], mat = src >> 28, idx = src & 0x0fffffff;
    return (mat == 0) ? v.A[idx] : (mat == 1 ? v.L[idx] : v.R[idx]);
}

LCQ_DEV double limit_scaling(double d) { d = d < kMinScaling ? 1.0 : d; return d > kMaxScaling ? kMaxScaling : d; }   // scaling.c:7-14

// ---- sparse products (lin_alg.c mat_vec / mat_tpose_vec) -----------------------------------------------------
// acc[tgt[p]] += val(p) * v[src[p]] over the flat entry list, in storage order -- the reference's order of
// accumulation.  The accumulator is the scratch vector (shared memory when it fits): a chain of read-modify-writes in
// global memory would cost an L2 round trip per entry.  The matrix values and vector entries of KB entries are loaded
// before the first accumulation (the compiler must assume that the vectors alias and would serialise otherwise).
template <class F>
LCQ_DEV void flat_acc(int nnz, const int* tgt, const int* src, F val, const Vec& v, const Vec& acc, bool skip_diag = false)
{
    constexpr int KB = 8;
    for (int p0 = 0; p0 < nnz; p0 += KB) {
        double a[KB], x[KB];
#pragma unroll
        for (int k = 0; k < KB; k++) { const int p = p0 + k < nnz ? p0 + k : nnz - 1; a[k] = val(p); x[k] = v[src[p]]; }
#pragma unroll
        for (int k = 0; k < KB; k++)
            if (p0 + k < nnz && !(skip_diag && tgt[p0 + k] == src[p0 + k])) acc[tgt[p0 + k]] += a[k] * x[k];
    }
}
LCQ_DEV void vec_zero(const Vec& a, int len) { for (int i = 0; i < len; i++) a[i] = 0.0; }
// out[i] = a[i] (+ b[i]) for i < len, eight loads ahead of the stores
LCQ_DEV void vec_out(const Vec& out, const Vec& a, const Vec* b, int len)
{
    for (int i0 = 0; i0 < len; i0 += 8) {
        double t[8];
#pragma unroll
        for (int k = 0; k < 8; k++) { const int i = i0 + k < len ? i0 + k : len - 1; t[k] = b ? a[i] + (*b)[i] : a[i]; }
#pragma unroll
        for (int k = 0; k < 8; k++) if (i0 + k < len) out[i0 + k] = t[k];
    }
}
// out = A v                         (out: m, v: n)
LCQ_DEVN void A_mul(const SymDev& S, const Work& w, const Vec& v, const Vec& out)
{
    vec_zero(w.sm, S.m);
    flat_acc(S.nnzA, S.Ai, S.Acol, [&](int p) { return w.Ax[p]; }, v, w.sm);
    vec_out(out, w.sm, nullptr, S.m);
}
// out = A' v                        (out: n, v: m)
LCQ_DEVN void At_mul(const SymDev& S, const Work& w, const Vec& v, const Vec& out)
{
    vec_zero(w.sm, S.n);
    flat_acc(S.nnzA, S.Acol, S.Ai, [&](int p) { return w.Ax[p]; }, v, w.sm);
    vec_out(out, w.sm, nullptr, S.n);
}
// out = P v with P stored as its upper triangle (auxil.c:294-300: upper part, then the strict lower part)
LCQ_DEVN void P_mul(const SymDev& S, const Work& w, const Vec& v, const Vec& out)
{
    Vec lo;
    lo.p = &w.sm[S.n];
    vec_zero(w.sm, 2 * S.n);
    flat_acc(S.nnzP, S.Pi, S.Pcol, [&](int p) { return w.Px[p]; }, v, w.sm);
    flat_acc(S.nnzP, S.Pcol, S.Pi, [&](int p) { return w.Px[p]; }, v, lo, true);
    vec_out(out, w.sm, &lo, S.n);
}

// ---- scale_data (scaling.c:44-156) ---------------------------------------------------------------------------
LCQ_DEVN void scale_data(const SymDev& S, const View& v, const lcqp_cuda_options& o, const Work& w, State& st)
{
    const int n = S.n, m = S.m;
    for (int p = 0; p < S.nnzP; p++) w.Px[p] = v.Q[S.Psrc[p]];
    for (int p = 0; p < S.nnzA; p++) w.Ax[p] = a_val(S, v, p);
    for (int j = 0; j < n; j++) { w.sD[j] = 1.0; w.q[j] = w.gk[j]; }
    for (int i = 0; i < m; i++) w.sE[i] = 1.0;
    double c = 1.0;
    for (int it = 0; it < o.osqp_scaling; it++) {
        // column norms of [P; A] into tn1, row norms of A into tm1
        for (int j = 0; j < n; j++) w.tn1[j] = 0.0;
        for (int i = 0; i < m; i++) w.tm1[i] = 0.0;
        for (int j = 0; j < n; j++)
            for (int p = S.Pp[j]; p < S.Pp[j + 1]; p++) {
                const int i = S.Pi[p];
                const double a = fabs(w.Px[p]);
                w.tn1[j] = fmax(a, w.tn1[j]);
                if (i != j) w.tn1[i] = fmax(a, w.tn1[i]);
            }
        for (int j = 0; j < n; j++) {
            double cn = 0.0;
            for (int p = S.Ap[j]; p < S.Ap[j + 1]; p++) { const double a = fabs(w.Ax[p]); cn = fmax(cn, a); const int i = S.Ai[p]; w.tm1[i] = fmax(w.tm1[i], a); }
            w.tn1[j] = fmax(w.tn1[j], cn);
        }
        for (int j = 0; j < n; j++) w.tn1[j] = 1.0 / sqrt(limit_scaling(w.tn1[j]));
        for (int i = 0; i < m; i++) w.tm1[i] = 1.0 / sqrt(limit_scaling(w.tm1[i]));
        // P <- D P D, A <- E A D, q <- D q
        for (int j = 0; j < n; j++) {
            const double dj = w.tn1[j];
            for (int p = S.Pp[j]; p < S.Pp[j + 1]; p++) { double t = w.Px[p]; t *= w.tn1[S.Pi[p]]; t *= dj; w.Px[p] = t; }
            for (int p = S.Ap[j]; p < S.Ap[j + 1]; p++) { double t = w.Ax[p]; t *= w.tm1[S.Ai[p]]; t *= dj; w.Ax[p] = t; }
            w.q[j] = dj * w.q[j];
            w.sD[j] = w.sD[j] * dj;
        }
        for (int i = 0; i < m; i++) w.sE[i] = w.sE[i] * w.tm1[i];
        // cost normalisation
        for (int j = 0; j < n; j++) w.tn1[j] = 0.0;
        for (int j = 0; j < n; j++)
            for (int p = S.Pp[j]; p < S.Pp[j + 1]; p++) {
                const int i = S.Pi[p];
                const double a = fabs(w.Px[p]);
                w.tn1[j] = fmax(a, w.tn1[j]);
                if (i != j) w.tn1[i] = fmax(a, w.tn1[i]);
            }
        double mean = 0.0, nq = 0.0;
        for (int j = 0; j < n; j++) { mean += w.tn1[j]; nq = fmax(nq, fabs(w.q[j])); }
        mean /= (double)n;
        nq = limit_scaling(nq);
        double ct = limit_scaling(fmax(mean, nq));
        ct = 1.0 / ct;
        for (int p = 0; p < S.nnzP; p++) w.Px[p] *= ct;
        for (int j = 0; j < n; j++) w.q[j] *= ct;
        c *= ct;
    }
    st.c = c;
    st.cinv = 1.0 / c;
    for (int j = 0; j < n; j++) w.sDi[j] = 1.0 / w.sD[j];
    for (int i = 0; i < m; i++) w.sEi[i] = 1.0 / w.sE[i];
    for (int i = 0; i < m; i++) { w.l[i] = w.sE[i] * w.l[i]; w.u[i] = w.sE[i] * w.u[i]; }
}

// ---- set_rho_vec (auxil.c:76-100); rv = rho_vec, riv = rho_inv_vec, act doubles as constr_type here is NOT used:
// the type is recomputed from the bounds where osqp_update_rho needs it (the bounds never change)
LCQ_DEV int constr_type(const Work& w, int i)
{
    if (w.l[i] < -kOsqpInfty * kMinScaling && w.u[i] > kOsqpInfty * kMinScaling) return -1;
    if (w.u[i] - w.l[i] < kRhoTol) return 1;
    return 0;
}
LCQ_DEVN void set_rho_vec(const SymDev& S, const Work& w, State& st)
{
    st.rho = fmin(fmax(st.rho, kRhoMin), kRhoMax);
    for (int i = 0; i < S.m; i++) {
        const int t = constr_type(w, i);
        const double r = (t < 0) ? kRhoMin : (t > 0 ? kRhoEqOverIneq * st.rho : st.rho);
        w.rv[i] = r;
        w.riv[i] = 1.0 / r;
    }
}
// osqp_update_rho (osqp.c:1268-1318): loose rows keep RHO_MIN
LCQ_DEVN void update_rho_vec(const SymDev& S, const Work& w, State& st, double rho_new)
{
    st.rho = fmin(fmax(rho_new, kRhoMin), kRhoMax);
    for (int i = 0; i < S.m; i++) {
        const int t = constr_type(w, i);
        if (t == 0) { w.rv[i] = st.rho; w.riv[i] = 1.0 / st.rho; }
        else if (t == 1) { w.rv[i] = kRhoEqOverIneq * st.rho; w.riv[i] = 1.0 / w.rv[i]; }
    }
}

// ---- numeric LDL' of the permuted KKT matrix (kkt.c:6-177 for the values, qdldl.c:72-233 for the elimination) ----
// polish = 0:  [P + sigma I, A'; A, -diag(1/rho)]     polish = 1:  [P + delta I, Aact'; Aact, -delta I]
LCQ_DEV double kkt_value(const SymDev& S, const Work& w, int e, bool diag, int polish, double sigma, double delta)
{
    const int src = S.Ksrc[e], type = src >> 28, idx = src & 0x0fffffff;
    if (type == 0) return w.Px[idx] + (diag ? (polish ? delta : sigma) : 0.0);
    if (type == 1) return (polish && w.act[S.Ai[idx]] == 0.0) ? 0.0 : w.Ax[idx];
    if (type == 2) return polish ? -delta : -w.riv[idx];
    return polish ? delta : sigma;
}

LCQ_DEVN void kkt_factor(const SymDev& S, const Work& w, State& st, int polish, double sigma, double delta)
{
    const int N = S.N;
    const Vec Lx = polish ? w.Lxp : w.Lx, Dinv = polish ? w.Dinvp : w.Dinv;
    const Vec yv = w.sm;
    int positive = 0, bad = 0;
    for (int k = 0; k < N; k++) yv[k] = 0.0;
    for (int k = 0; k < N; k++) {
        double dk = 0.0;
        for (int e0 = S.Kp[k]; e0 < S.Kp[k + 1]; e0 += 8) {
            double kv[8];
            const int e1 = S.Kp[k + 1];
#pragma unroll
            for (int q = 0; q < 8; q++) { const int e = e0 + q < e1 ? e0 + q : e1 - 1; kv[q] = kkt_value(S, w, e, S.Ki[e] == k, polish, sigma, delta); }
#pragma unroll
            for (int q = 0; q < 8; q++)
                if (e0 + q < e1) { const int i = S.Ki[e0 + q]; if (i == k) dk = kv[q]; else yv[i] = kv[q]; }
        }
        for (int t = S.rp[k]; t < S.rp[k + 1]; t++) {
            const int c = S.rcol[t], pos = S.rpos[t];
            const double yc = yv[c];
            yv[c] = 0.0;
            // (the rows of one column are distinct: four independent updates in flight)
            for (int p0 = S.Lp[c]; p0 < pos; p0 += 8) {
                double l[8];
#pragma unroll
                for (int q = 0; q < 8; q++) l[q] = (p0 + q < pos) ? Lx[p0 + q] : 0.0;
#pragma unroll
                for (int q = 0; q < 8; q++) if (p0 + q < pos) yv[S.Li[p0 + q]] -= l[q] * yc;
            }
            const double lkc = yc * Dinv[c];
            Lx[pos] = lkc;
            dk -= yc * lkc;
        }
        if (dk == 0.0) bad = 1;
        if (dk > 0.0) positive++;
        Dinv[k] = 1.0 / dk;
    }
    // quasi-definite: exactly n positive pivots (the polish system keeps n positive ones as well)
    st.factor_bad = (bad || positive != S.n) ? 1 : 0;
    st.factor_count++;
}

// b (N, natural order) <- solution.  polish = 0 follows qdldl_interface.c:341-376: the x part is the solution, the
// z part becomes b_z + rho_inv * nu.  polish = 1: plain solve.
LCQ_DEVN void kkt_solve(const SymDev& S, const Work& w, const Vec& b, int polish)
{
    // The sweeps run over the entries of L in storage order with the matrix values of the next KB entries loaded ahead
    // (independent, streaming loads) of the chain of dependent read-modify-writes on the scratch vector; the
    // arithmetic is that of qdldl.c:236-281 operation for operation (bp[col] is final when its column is reached).
    constexpr int KB = 8;
    const int N = S.N, n = S.n, nnzL = S.nnzL;
    const Vec bp = w.sm;
    const Vec Lx = polish ? w.Lxp : w.Lx, Dinv = polish ? w.Dinvp : w.Dinv;
    for (int k0 = 0; k0 < N; k0 += 8) {
        double t[8];
#pragma unroll
        for (int k = 0; k < 8; k++) t[k] = (k0 + k < N) ? b[S.perm[k0 + k]] : 0.0;
#pragma unroll
        for (int k = 0; k < 8; k++) if (k0 + k < N) bp[k0 + k] = t[k];
    }
    for (int p0 = 0; p0 < nnzL; p0 += KB) {
        double l[KB];
#pragma unroll
        for (int k = 0; k < KB; k++) l[k] = (p0 + k < nnzL) ? Lx[p0 + k] : 0.0;
#pragma unroll
        for (int k = 0; k < KB; k++)
            if (p0 + k < nnzL) { const int r = S.Li[p0 + k], c = S.Lcol[p0 + k]; bp[r] -= l[k] * bp[c]; }
    }
    for (int k0 = 0; k0 < N; k0 += 8) {
        double t[8];
#pragma unroll
        for (int k = 0; k < 8; k++) t[k] = (k0 + k < N) ? Dinv[k0 + k] : 0.0;
#pragma unroll
        for (int k = 0; k < 8; k++) if (k0 + k < N) bp[k0 + k] *= t[k];
    }
    for (int q0 = 0; q0 < nnzL; q0 += KB) {
        double l[KB];
#pragma unroll
        for (int k = 0; k < KB; k++) l[k] = (q0 + k < nnzL) ? Lx[S.Lrev[q0 + k]] : 0.0;
#pragma unroll
        for (int k = 0; k < KB; k++)
            if (q0 + k < nnzL) { const int p = S.Lrev[q0 + k]; const int r = S.Li[p], c = S.Lcol[p]; bp[c] -= l[k] * bp[r]; }
    }
    for (int k0 = 0; k0 < N; k0 += 8) {
        double t[8], u[8];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int old = (k0 + k < N) ? S.perm[k0 + k] : 0;
            t[k] = (k0 + k < N) ? bp[k0 + k] : 0.0;
            u[k] = (k0 + k < N && !polish && old >= n) ? b[old] + w.riv[old - n] * t[k] : t[k];
        }
#pragma unroll
        for (int k = 0; k < 8; k++) if (k0 + k < N) b[S.perm[k0 + k]] = u[k];
    }
}

// ---- update_info (auxil.c:564-629) for the iterate (x, z, y): residuals, leaving Ax, Px, A'y behind and -- as the
// reference does -- the (scaled) residual vectors in zp and xp, which compute_rho_estimate reads (auxil.c:26-27)
LCQ_DEVN void residuals(const SymDev& S, const Work& w, const State& st, const Vec& x, const Vec& z, const Vec& y, double& pri, double& dua)
{
    const int n = S.n, m = S.m;
    if (m == 0) pri = 0.0;
    else {
        A_mul(S, w, x, w.Axv);
        double r = 0.0;
        for (int i = 0; i < m; i++) { const double d = w.Axv[i] - z[i]; w.zp[i] = d; r = fmax(r, fabs(w.sEi[i] * d)); }
        pri = r;
    }
    P_mul(S, w, x, w.Pxv);
    if (m > 0) At_mul(S, w, y, w.Aty);
    double r = 0.0;
    for (int j = 0; j < n; j++) {
        double d = w.q[j];
        d = d + w.Pxv[j];
        if (m > 0) d = d + w.Aty[j];
        w.xp[j] = d;
        r = fmax(r, fabs(w.sDi[j] * d));
    }
    dua = st.cinv * r;
}

// compute_rho_estimate (auxil.c:13-52)
LCQ_DEVN double rho_estimate(const SymDev& S, const Work& w, const State& st)
{
    const int n = S.n, m = S.m;
    double pri = 0, dua = 0, a = 0, b = 0, c1 = 0, c2 = 0, c3 = 0;
    for (int i = 0; i < m; i++) { pri = fmax(pri, fabs(w.zp[i])); a = fmax(a, fabs(w.z[i])); b = fmax(b, fabs(w.Axv[i])); }
    for (int j = 0; j < n; j++) { dua = fmax(dua, fabs(w.xp[j])); c1 = fmax(c1, fabs(w.q[j])); c2 = fmax(c2, fabs(w.Aty[j])); c3 = fmax(c3, fabs(w.Pxv[j])); }
    pri /= (fmax(a, b) + kDivTol);
    dua /= (fmax(fmax(c1, c2), c3) + kDivTol);
    const double e = st.rho * sqrt(pri / dua);
    return fmin(fmax(e, kRhoMin), kRhoMax);
}

// is_primal_infeasible (auxil.c:362-424); modifies dy like the reference
LCQ_DEVN bool primal_infeasible(const SymDev& S, const Work& w, double eps)
{
    const int n = S.n, m = S.m;
    for (int i = 0; i < m; i++) {
        if (w.u[i] > kOsqpInfty * kMinScaling) {
            if (w.l[i] < -kOsqpInfty * kMinScaling) w.dy[i] = 0.0;
            else w.dy[i] = fmin(w.dy[i], 0.0);
        } else if (w.l[i] < -kOsqpInfty * kMinScaling) w.dy[i] = fmax(w.dy[i], 0.0);
    }
    double nrm = 0.0;
    for (int i = 0; i < m; i++) nrm = fmax(nrm, fabs(w.sE[i] * w.dy[i]));
    if (nrm > kDivTol) {
        double lhs = 0.0;
        for (int i = 0; i < m; i++) lhs += w.u[i] * fmax(w.dy[i], 0.0) + w.l[i] * fmin(w.dy[i], 0.0);
        if (lhs < eps * nrm) {
            At_mul(S, w, w.dy, w.tn1);
            double r = 0.0;
            for (int j = 0; j < n; j++) r = fmax(r, fabs(w.sDi[j] * w.tn1[j]));
            return r < eps * nrm;
        }
    }
    return false;
}

// is_dual_infeasible (auxil.c:426-520)
LCQ_DEVN bool dual_infeasible(const SymDev& S, const Work& w, const State& st, double eps)
{
    const int n = S.n, m = S.m;
    double nrm = 0.0;
    for (int j = 0; j < n; j++) nrm = fmax(nrm, fabs(w.sD[j] * w.dx[j]));
    if (!(nrm > kDivTol)) return false;
    double qd = 0.0;
    for (int j = 0; j < n; j++) qd += w.q[j] * w.dx[j];
    if (!(qd < st.c * eps * nrm)) return false;
    P_mul(S, w, w.dx, w.tn2);
    double r = 0.0;
    for (int j = 0; j < n; j++) r = fmax(r, fabs(w.sDi[j] * w.tn2[j]));
    if (!(r < st.c * eps * nrm)) return false;
    A_mul(S, w, w.dx, w.tm1);
    for (int i = 0; i < m; i++) {
        const double a = w.sEi[i] * w.tm1[i];
        if ((w.u[i] < kOsqpInfty * kMinScaling && a > eps * nrm) || (w.l[i] > -kOsqpInfty * kMinScaling && a < -eps * nrm)) return false;
    }
    return true;
}

// check_termination (auxil.c:681-786); returns 1 when the loop ends
LCQ_DEVN int check_termination(const SymDev& S, const lcqp_cuda_options& o, const Work& w, State& st, int approximate)
{
    const int n = S.n, m = S.m;
    double eps_abs = o.osqp_eps_abs, eps_rel = o.osqp_eps_rel, eps_pi = o.osqp_eps_prim_inf, eps_di = o.osqp_eps_dual_inf;
    if (st.pri_res > kOsqpInfty || st.dua_res > kOsqpInfty) { st.status_val = OSQP_NON_CVX; return 1; }
    if (approximate) { eps_abs *= 10; eps_rel *= 10; eps_pi *= 10; eps_di *= 10; }
    bool prim_ok = false, dual_ok = false, prim_inf = false, dual_inf = false;
    if (m == 0) prim_ok = true;
    else {
        double a = 0, b = 0;
        for (int i = 0; i < m; i++) { a = fmax(a, fabs(w.sEi[i] * w.z[i])); b = fmax(b, fabs(w.sEi[i] * w.Axv[i])); }
        const double eps_prim = eps_abs + eps_rel * fmax(a, b);
        if (st.pri_res < eps_prim) prim_ok = true;
        else prim_inf = primal_infeasible(S, w, eps_pi);
    }
    {
        double a = 0, b = 0, c = 0;
        for (int j = 0; j < n; j++) { a = fmax(a, fabs(w.sDi[j] * w.q[j])); b = fmax(b, fabs(w.sDi[j] * w.Aty[j])); c = fmax(c, fabs(w.sDi[j] * w.Pxv[j])); }
        double mx = fmax(fmax(a, b), c);
        mx *= st.cinv;
        const double eps_dual = eps_abs + eps_rel * mx;
        if (st.dua_res < eps_dual) dual_ok = true;
        else dual_inf = dual_infeasible(S, w, st, eps_di);
    }
    if (prim_ok && dual_ok) { st.status_val = approximate ? OSQP_SOLVED_INACCURATE : OSQP_SOLVED; return 1; }
    if (prim_inf) { st.status_val = approximate ? OSQP_PRIMAL_INFEASIBLE_INACCURATE : OSQP_PRIMAL_INFEASIBLE; return 1; }
    if (dual_inf) { st.status_val = approximate ? OSQP_DUAL_INFEASIBLE_INACCURATE : OSQP_DUAL_INFEASIBLE; return 1; }
    return 0;
}

// ---- polish (polish.c:237-350) --------------------------------------------------------------------------------
LCQ_DEVN void polish(const SymDev& S, const lcqp_cuda_options& o, const Work& w, State& st)
{
    const int n = S.n, m = S.m, N = S.N;
    // form_Ared (:19-104): act = -1 lower-active, +1 upper-active, 0 inactive
    for (int i = 0; i < m; i++) {
        double a = 0.0;
        if (w.z[i] - w.l[i] < -w.y[i]) a = -1.0;
        else if (w.u[i] - w.z[i] < w.y[i]) a = 1.0;
        w.act[i] = a;
    }
    const double sig = o.osqp_sigma, del = o.osqp_delta;
    kkt_factor(S, w, st, 1, sig, del);
    if (st.factor_bad) { st.factor_bad = 0; return; }   // polishing failed (:262-270); the ADMM factor is untouched
    // rhs_red = [-q; l_low; u_upp] (:112-128) in xt; solve, then iterative refinement against the unregularised K (:141-197)
    const Vec sol = w.xt;
    for (int j = 0; j < n; j++) sol[j] = -w.q[j];
    for (int i = 0; i < m; i++) sol[n + i] = (w.act[i] < 0.0) ? w.l[i] : (w.act[i] > 0.0 ? w.u[i] : 0.0);
    kkt_solve(S, w, sol, 1);
    for (int it = 0; it < o.osqp_polish_refine_iter; it++) {
        // r = b - K z  with  K = [P, Aact'; Aact, 0]
        const Vec r = w.w3;
        Vec ys;
        ys.p = &sol[n];
        P_mul(S, w, sol, w.tn1);                       // P x            (sol[0..n) is x)
        vec_zero(w.sm, n);                             // Aact' y_red
        flat_acc(S.nnzA, S.Acol, S.Ai, [&](int p) { return w.act[S.Ai[p]] != 0.0 ? w.Ax[p] : 0.0; }, ys, w.sm);
        for (int j = 0; j < n; j++) r[j] = ((-w.q[j]) - w.tn1[j]) - w.sm[j];
        A_mul(S, w, sol, w.tm1);                       // A x
        for (int i = 0; i < m; i++) r[n + i] = (w.act[i] < 0.0) ? (w.l[i] - w.tm1[i]) : (w.act[i] > 0.0 ? (w.u[i] - w.tm1[i]) : 0.0);
        // (the factorisation and the solves use their own scratch, w.sm)
        kkt_solve(S, w, r, 1);
        for (int k = 0; k < N; k++) sol[k] += r[k];
    }
    // polished (x, z, y) (:302-308): px, pz = A px, py from the reduced duals, then the normal-cone projection (proj.c:17-31)
    for (int j = 0; j < n; j++) w.px[j] = sol[j];
    A_mul(S, w, w.px, w.pz);
    for (int i = 0; i < m; i++) {
        const double yi = (w.act[i] != 0.0) ? sol[n + i] : 0.0;
        const double t = w.pz[i] + yi;
        const double zi = fmin(fmax(t, w.l[i]), w.u[i]);
        w.pz[i] = zi;
        w.py[i] = t - zi;
    }
    // residuals of the polished point (update_info(work, 0, 1, 1)); they overwrite Ax, Px, A'y and zp/xp like the reference
    double ppri, pdua;
    residuals(S, w, st, w.px, w.pz, w.py, ppri, pdua);
    const bool ok = (ppri < st.pri_res && pdua < st.dua_res) || (ppri < st.pri_res && st.dua_res < 1e-10) || (pdua < st.dua_res && st.pri_res < 1e-10);
    if (ok) {
        st.pri_res = ppri; st.dua_res = pdua;
        for (int j = 0; j < n; j++) w.x[j] = w.px[j];
        for (int i = 0; i < m; i++) { w.z[i] = w.pz[i]; w.y[i] = w.py[i]; }
    }
}

// ---- osqp_solve (osqp.c:288-641).  Returns the exit flag the adapter reads (status_val). ----------------------
// `mask`: the lanes of the warp that solve a QP in this pass (all of them call this function together)
LCQ_DEVN int osqp_solve(const SymDev& S, const lcqp_cuda_options& o, Work& w, State& st, unsigned mask)
{
    const int n = S.n, m = S.m;
    const double sigma = o.osqp_sigma, alpha = o.osqp_alpha;
    const int check = o.osqp_check_termination;
    st.status_val = OSQP_UNSOLVED;
    int iter = 1, it_end = 0;
    bool can_check = false, running = true;
    for (iter = 1; OSQ_ANY(mask, running); iter++) {
      if (running && iter > o.osqp_max_iter) { running = false; it_end = iter; }
      if (running) {
        { const Vec t = w.x; w.x = w.xp; w.xp = t; }
        { const Vec t = w.z; w.z = w.zp; w.zp = t; }
        // update_xz_tilde (auxil.c:161-183).  (Elementwise loops: the inputs of four elements are loaded before the
        // first store -- the compiler must assume that the vectors alias and would serialise load -> store -> load.)
        for (int j0 = 0; j0 < n; j0 += 4) {
            double a[4], c[4];
#pragma unroll
            for (int k = 0; k < 4; k++) { const int j = j0 + k < n ? j0 + k : n - 1; a[k] = w.xp[j]; c[k] = w.q[j]; }
#pragma unroll
            for (int k = 0; k < 4; k++) if (j0 + k < n) w.xt[j0 + k] = sigma * a[k] - c[k];
        }
        for (int i0 = 0; i0 < m; i0 += 4) {
            double a[4], c[4], d[4];
#pragma unroll
            for (int k = 0; k < 4; k++) { const int i = i0 + k < m ? i0 + k : m - 1; a[k] = w.zp[i]; c[k] = w.riv[i]; d[k] = w.y[i]; }
#pragma unroll
            for (int k = 0; k < 4; k++) if (i0 + k < m) w.xt[n + i0 + k] = a[k] - c[k] * d[k];
        }
        kkt_solve(S, w, w.xt, 0);
        // update_x, update_z, update_y (auxil.c:185-225)
        for (int j0 = 0; j0 < n; j0 += 4) {
            double a[4], c[4];
#pragma unroll
            for (int k = 0; k < 4; k++) { const int j = j0 + k < n ? j0 + k : n - 1; a[k] = w.xt[j]; c[k] = w.xp[j]; }
#pragma unroll
            for (int k = 0; k < 4; k++)
                if (j0 + k < n) { const double xn = alpha * a[k] + (1.0 - alpha) * c[k]; w.x[j0 + k] = xn; w.dx[j0 + k] = xn - c[k]; }
        }
        for (int i0 = 0; i0 < m; i0 += 4) {
            double zt[4], zo[4], ri[4], yy[4], lo[4], up[4], rr[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int i = i0 + k < m ? i0 + k : m - 1;
                zt[k] = w.xt[n + i]; zo[k] = w.zp[i]; ri[k] = w.riv[i]; yy[k] = w.y[i]; lo[k] = w.l[i]; up[k] = w.u[i]; rr[k] = w.rv[i];
            }
#pragma unroll
            for (int k = 0; k < 4; k++)
                if (i0 + k < m) {
                    double zn = alpha * zt[k] + (1.0 - alpha) * zo[k] + ri[k] * yy[k];
                    zn = fmin(fmax(zn, lo[k]), up[k]);
                    w.z[i0 + k] = zn;
                    const double d = rr[k] * (alpha * zt[k] + (1.0 - alpha) * zo[k] - zn);
                    w.dy[i0 + k] = d;
                    w.y[i0 + k] = yy[k] + d;
                }
        }
        st.admm_total++;
        can_check = check && (iter % check == 0);
        if (can_check) {
            st.iter = iter;
            residuals(S, w, st, w.x, w.z, w.y, st.pri_res, st.dua_res);
            if (check_termination(S, o, w, st, 0)) { running = false; it_end = iter; }
        }
        if (running && o.osqp_adaptive_rho && st.interval && (iter % st.interval == 0)) {
            if (!can_check) { st.iter = iter; residuals(S, w, st, w.x, w.z, w.y, st.pri_res, st.dua_res); }
            // adapt_rho (auxil.c:54-74)
            const double rn = rho_estimate(S, w, st);
            if (rn > st.rho * o.osqp_adaptive_rho_tolerance || rn < st.rho / o.osqp_adaptive_rho_tolerance) {
                update_rho_vec(S, w, st, rn);
                kkt_factor(S, w, st, 0, sigma, o.osqp_delta);
            }
        }
      }
      OSQ_SYNCWARP(mask);
    }
    iter = it_end;
    if (!can_check) {
        st.iter = iter - 1;
        residuals(S, w, st, w.x, w.z, w.y, st.pri_res, st.dua_res);
        check_termination(S, o, w, st, 0);
    }
    if (st.status_val == OSQP_UNSOLVED) {
        if (!check_termination(S, o, w, st, 1)) st.status_val = OSQP_MAX_ITER_REACHED;
    }
    if (o.osqp_polish && st.status_val == OSQP_SOLVED) polish(S, o, w, st);
    return st.status_val;
}

LCQ_DEV bool has_solution(int s)
{
    return s != OSQP_PRIMAL_INFEASIBLE && s != OSQP_PRIMAL_INFEASIBLE_INACCURATE && s != OSQP_DUAL_INFEASIBLE && s != OSQP_DUAL_INFEASIBLE_INACCURATE && s != OSQP_NON_CVX;
}

// ------------------------------------------------------------------------------------------------
// The penalty loop of LCQProblem::runSolver (/root/reference/src/LCQProblem.cpp:444-560, helpers :1105-1482) for
// one instance over the OSQP flavour: nDuals = nC + 2 nComp, boxDualOffset = 0 (:934-935).
// ------------------------------------------------------------------------------------------------
// unscaled products with the caller's values
LCQ_DEVN void Q_mul_raw(const SymDev& S, const View& v, const Work& w, const Vec& x, const Vec& out)   // out = Q x (full symmetric pattern)
{
    vec_zero(w.sm, S.n);
    flat_acc(S.nnzQ, S.Qi, S.Qcol, [&](int p) { return v.Q[S.Qsrc[p]]; }, x, w.sm);
    vec_out(out, w.sm, nullptr, S.n);
}
LCQ_DEVN void A_mul_raw(const SymDev& S, const View& v, const Work& w, const Vec& x, const Vec& out)   // out = [A; L; R] x
{
    vec_zero(w.sm, S.m);
    flat_acc(S.nnzA, S.Ai, S.Acol, [&](int p) { return a_val(S, v, p); }, x, w.sm);
    vec_out(out, w.sm, nullptr, S.m);
}
LCQ_DEVN void At_mul_raw(const SymDev& S, const View& v, const Work& w, const Vec& y, const Vec& out)  // out = [A; L; R]' y
{
    vec_zero(w.sm, S.n);
    flat_acc(S.nnzA, S.Acol, S.Ai, [&](int p) { return a_val(S, v, p); }, y, w.sm);
    vec_out(out, w.sm, nullptr, S.n);
}

// out = Qk x + add = Q x + rho (L'(R x) + R'(L x)) + add ; leaves [A; L; R] x in tm1.  `add` may alias nothing (null p = none)
LCQ_DEVN void Qk_mul(const SymDev& S, const View& v, const Work& w, double rho, const Vec& x, const Vec* add, const Vec& out)
{
    const int nC = S.nC, nComp = S.nComp;
    Q_mul_raw(S, v, w, x, w.tn1);
    A_mul_raw(S, v, w, x, w.tm1);
    for (int i = 0; i < nC; i++) w.tm2[i] = 0.0;
    for (int i = 0; i < nComp; i++) { w.tm2[nC + i] = w.tm1[nC + nComp + i]; w.tm2[nC + nComp + i] = w.tm1[nC + i]; }
    At_mul_raw(S, v, w, w.tm2, w.tn2);
    for (int j = 0; j < S.n; j++) out[j] = (w.tn1[j] + (add ? (*add)[j] : 0.0)) + rho * w.tn2[j];
}

LCQ_DEV void lcqp_loop(const SymDev& S, const View& v, const lcqp_cuda_options& o, Work& w, unsigned long long instance,
                       double* xout, double* yout, LoopOut& out, State& st)
{
    const int n = S.n, m = S.m, nC = S.nC, nComp = S.nComp;
    double hist[kMaxLeyffer];
    int nh = 0;
    double alphak = 1.0, rho = o.initialPenaltyParameter, phi_const = 0.0;
    int outerIter = 0, totalIter = 0, subIter = 0, exitFlag = 0, status = 0, ret = RET_OK;
    out.rhoOpt = 0.0;
    const bool have_gphi = (v.lbL != nullptr) || (v.lbR != nullptr);
    st.admm_total = 0; st.factor_count = 0; st.factor_bad = 0;

    for (int j = 0; j < n; j++) { w.xk[j] = v.x0 ? v.x0[j] : 0.0; w.gt[j] = v.g[j]; w.pk[j] = 0.0; w.gphi[j] = 0.0; }
    for (int i = 0; i < m; i++) w.yk[i] = 0.0;
    // bounds of [A; L; R] (LCQProblem.cpp:584-608)
    for (int i = 0; i < m; i++) {
        double lo, up;
        if (i < nC) { lo = v.lbA ? v.lbA[i] : -INFINITY; up = v.ubA ? v.ubA[i] : INFINITY; }
        else if (i < nC + nComp) { lo = v.lbL ? v.lbL[i - nC] : 0.0; up = v.ubL ? v.ubL[i - nC] : INFINITY; }
        else { lo = v.lbR ? v.lbR[i - nC - nComp] : 0.0; up = v.ubR ? v.ubR[i - nC - nComp] : INFINITY; }
        w.l[i] = lo; w.u[i] = up;
    }
    if (have_gphi) {  // :970-996
        double pc = 0.0;
        for (int i = 0; i < nComp; i++) pc += (v.lbL ? v.lbL[i] : 0.0) * (v.lbR ? v.lbR[i] : 0.0);
        phi_const = pc;
        for (int i = 0; i < nC; i++) w.tm2[i] = 0.0;
        for (int i = 0; i < nComp; i++) { w.tm2[nC + i] = v.lbR ? v.lbR[i] : 0.0; w.tm2[nC + nComp + i] = v.lbL ? v.lbL[i] : 0.0; }
        At_mul_raw(S, v, w, w.tm2, w.gphi);   // L' lbR + R' lbL
        for (int j = 0; j < n; j++) w.gphi[j] = -w.gphi[j];
    }

    auto phi = [&]() -> double {  // getPhi :1172-1185 ; leaves [A; L; R] xk in tm1
        A_mul_raw(S, v, w, w.xk, w.tm1);
        double p = 0.0;
        for (int i = 0; i < nComp; i++) p += w.tm1[nC + i] * w.tm1[nC + nComp + i];
        if (have_gphi) for (int j = 0; j < n; j++) p += w.gphi[j] * w.xk[j];
        return phi_const + p;
    };
    auto update_penalty = [&]() {  // :1199-1214
        nh = 0;
        rho *= o.penaltyUpdateFactor;
        out.rhoOpt = rho;
        if (have_gphi) for (int j = 0; j < n; j++) w.gt[j] = v.g[j] + rho * w.gphi[j];
    };
    auto linearize = [&]() {  // :1105-1112 : gk = rho C xk + g_tilde
        A_mul_raw(S, v, w, w.xk, w.tm1);
        for (int i = 0; i < nC; i++) w.tm2[i] = 0.0;
        for (int i = 0; i < nComp; i++) { w.tm2[nC + i] = w.tm1[nC + nComp + i]; w.tm2[nC + nComp + i] = w.tm1[nC + i]; }
        At_mul_raw(S, v, w, w.tm2, w.tn2);
        for (int j = 0; j < n; j++) w.gk[j] = rho * w.tn2[j] + w.gt[j];
    };
    // solveQPSubproblem :1115-1148 over SubsolverOSQP::solve, in three steps so that the lanes of the warp enter the ADMM
    // loop together: prepare (setup or cost update; false: the QP cannot start), osqp_solve, finish (read the solution).
    auto qp_prepare = [&](bool initial) -> bool {
        if (initial) {
            // osqp_setup (osqp.c:96-283): validate_data rejects l > u -> the workspace stays NULL and the warm start fails
            for (int i = 0; i < m; i++) if (w.l[i] > w.u[i]) { ret = RET_OSQP_GUESS; return false; }
            st.rho = o.osqp_rho;
            scale_data(S, v, o, w, st);
            set_rho_vec(S, w, st);
            kkt_factor(S, w, st, 0, o.osqp_sigma, o.osqp_delta);
            if (st.factor_bad) { ret = RET_OSQP_GUESS; return false; }
            st.interval = o.osqp_adaptive_rho_interval > 0 ? o.osqp_adaptive_rho_interval : 4 * (o.osqp_check_termination > 0 ? o.osqp_check_termination : 25);
            // cold start, then osqp_warm_start_x (osqp.c:954-974): x = Dinv x0, z = A x
            for (int j = 0; j < n; j++) { w.x[j] = w.sDi[j] * w.xk[j]; w.xp[j] = 0.0; }
            A_mul(S, w, w.x, w.z);
            for (int i = 0; i < m; i++) { w.y[i] = 0.0; w.zp[i] = 0.0; }
            if (v.y0) for (int i = 0; i < m; i++) w.y[i] = st.c * (w.sEi[i] * v.y0[n + i]);   // osqp_warm_start_y (:976-994)
        } else {
            for (int j = 0; j < n; j++) w.q[j] = st.c * (w.sD[j] * w.gk[j]);   // osqp_update_lin_cost (:752-782)
        }
        return true;
    };
    auto qp_finish = [&](int flag) -> bool {
        subIter += st.iter;
        exitFlag = flag;
        if (flag <= 0) { ret = RET_SUBPROBLEM; return false; }   // SubsolverOSQP.cpp:176-181
        // store_solution (auxil.c:533-562) + getSolution (SubsolverOSQP.cpp:187-200): xnew = D x, yk = -(cinv E y)
        if (has_solution(flag)) {
            for (int j = 0; j < n; j++) w.pk[j] = w.sD[j] * w.x[j] - w.xk[j];
            for (int i = 0; i < m; i++) w.yk[i] = -(st.cinv * (w.sE[i] * w.y[i]));
        } else {
            for (int j = 0; j < n; j++) w.pk[j] = NAN;
            for (int i = 0; i < m; i++) w.yk[i] = NAN;
            for (int j = 0; j < n; j++) w.x[j] = 0.0;
            for (int i = 0; i < m; i++) { w.z[i] = 0.0; w.y[i] = 0.0; }
        }
        return true;
    };

    bool failed = false, success = false;
    // (every lane of the warp runs an instance -- the kernel pads the last tile -- so the votes below take all 32 lanes)
    if (o.solveZeroPenaltyFirst) { for (int j = 0; j < n; j++) w.gk[j] = v.g[j]; }
    else linearize();
    {
        const bool go = qp_prepare(true);
        const unsigned qp_mask = OSQ_BALLOT(go);
        if (go) { const int flag = osqp_solve(S, o, w, st, qp_mask); if (!qp_finish(flag)) failed = true; }
        else failed = true;
    }
    out.rhoOpt = failed ? 0.0 : rho;

    // one pass of the loop; returns false when the instance is finished
    auto pass = [&]() -> bool {
        for (int j = 0; j < n; j++) w.xk[j] = w.xk[j] + alphak * w.pk[j];   // updateStep :1240
        // updateStationarity :1246-1272 (no box part on this path)
        Qk_mul(S, v, w, rho, w.xk, &w.gt, w.stat);
        At_mul_raw(S, v, w, w.yk, w.tn1);
        for (int j = 0; j < n; j++) w.stat[j] -= w.tn1[j];
        totalIter++;
        {   // leyfferCheckPositive :1275-1313
            const int nd = o.nDynamicPenalty < kMaxLeyffer ? o.nDynamicPenalty : kMaxLeyffer;
            bool fire = false;
            if (nd > 0) {
                const double cur = phi();
                if (nh < nd) hist[nh++] = cur;
                else {
                    if (!(cur < o.complementarityTolerance)) {
                        fire = true;
                        for (int i = 0; i < nd; i++) if (cur < o.etaDynamicPenalty * hist[i]) { fire = false; break; }
                    }
                    for (int i = 0; i + 1 < nd; i++) hist[i] = hist[i + 1];
                    hist[nd - 1] = cur;
                }
            }
            if (fire) { update_penalty(); outerIter++; }
        }
        double sm = 0.0;
        for (int j = 0; j < n; j++) sm = fmax(sm, fabs(w.stat[j]));
        if (sm < o.stationarityTolerance) {  // :511
            if (phi() < o.complementarityTolerance) {
                const double tc = o.complementarityTolerance;   // determineStationarityType :1412-1453 (tm1 = [A;L;R] xk)
                int fl = 0;
                for (int i = 0; i < nComp; i++) {
                    const double Lx = w.tm1[nC + i], Rx = w.tm1[nC + nComp + i];
                    if (!(Lx <= tc && Rx <= tc)) continue;
                    const double yl = w.yk[nC + i], yr = w.yk[nC + nComp + i];
                    const double prod = yl * yr, mn = fmin(yl, yr);
                    if (mn < 0) fl |= 1;
                    if (fabs(prod) >= tc && mn <= 0) { if (prod <= tc) fl |= 4; else fl |= 2; }
                }
                status = (fl & 4) ? 1 : (!(fl & 1) ? 4 : (!(fl & 2) ? 3 : 2));
                success = true;
                return false;
            } else { update_penalty(); outerIter++; }
        }
        if (totalIter > o.maxIterations) { ret = RET_MAX_ITER; return false; }
        if (rho > o.maxPenaltyParameter) { ret = RET_MAX_PEN; return false; }
        return true;
    };
    // the QP of the pass and the step length (second half of the reference's loop body): every lane that is still
    // running enters its QP in the same trip
    auto pass2 = [&](int flag) -> bool {
        if (!qp_finish(flag)) { failed = true; return false; }
        if (o.perturbStep)
            for (int j = 0; j < n; j++) w.xk[j] += perturb_draw(o.perturb_seed, instance, (unsigned)totalIter, (unsigned)j) * kEPS;
        {   // getOptimalStepLength :1217-1237
            Qk_mul(S, v, w, rho, w.pk, nullptr, w.stat);
            double qk = 0.0;
            for (int j = 0; j < n; j++) qk += w.stat[j] * w.pk[j];
            Qk_mul(S, v, w, rho, w.xk, &w.gt, w.stat);
            double lk = 0.0;
            for (int j = 0; j < n; j++) lk += w.stat[j] * w.pk[j];
            alphak = 1.0;
            if (qk > 0 && lk < 0) alphak = fmin(-lk / qk, 1.0);
        }
        return true;
    };
    bool run = !failed;
    while (OSQ_ANY(0xffffffffu, run)) {
        if (run) run = pass();
        if (run) { linearize(); qp_prepare(false); }   // :545-548
        const unsigned qp_mask = OSQ_BALLOT(run);
        if (run) { const int flag = osqp_solve(S, o, w, st, qp_mask); run = pass2(flag); }
        OSQ_SYNCWARP(0xffffffffu);
    }
    for (int j = 0; j < n; j++) xout[j] = w.xk[j];
    if (success) {   // transformDuals :1381-1409 ; tm1 holds [A; L; R] xk from the last phi()
        for (int i = 0; i < m; i++) {
            double y = w.yk[i];
            if (i >= nC && i < nC + nComp) y -= rho * w.tm1[nComp + i];
            else if (i >= nC + nComp) y -= rho * w.tm1[i - nComp];
            yout[i] = y;
        }
    } else for (int i = 0; i < m; i++) yout[i] = w.yk[i];
    out.ret = ret; out.status = status; out.iterTotal = totalIter; out.iterOuter = outerIter; out.subIter = subIter; out.exitFlag = exitFlag;
}

}  // namespace osq
}  // namespace lcqp

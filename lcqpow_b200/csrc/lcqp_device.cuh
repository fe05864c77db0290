// lcqp_device.cuh -- device code of the B200-native batched LCQP solver (sm_100a).
//
// One CTA owns one LCQP instance from loadLCQP to the stationarity classification: the whole
// penalty-homotopy loop of LCQProblem::runSolver (/root/reference/src/LCQProblem.cpp:444-560) and the
// convex QP under it run inside one persistent kernel; instances are pulled from a global work counter.
// Per-instance iterates, bounds, working set and (when it fits) the inverse Schur complement live in
// shared memory; matrices are streamed from global memory, where the operands shared by a batch
// (Q, A, L, R and everything prepared from them) stay L2-resident.
//
// The convex QP  min 1/2 x'Px + q'x  s.t. l <= Ahat x <= u  is solved EXACTLY (the contract of the
// reference's qpOASES subsolver):
//   phase 1 (first QP of an instance only): OSQP-style ADMM in condensed form
//       (P + sigma I + A'RA) xt = sigma x - q + A'(R z - y),  zt = A xt        (osqp auxil.c:161-225)
//     with the matrix inverted once per instance (or once per batch when Q/A/L/R are shared), probing
//     the active set every qp_check_interval iterations                        (osqp polish.c:33-49);
//   phase 2: primal active-set iteration on the regularised KKT system
//       [P + dI, Aw'; Aw, -dI]                                                 (osqp polish.c:232-300)
//     solved by block elimination through Hinv = (P+dI)^-1 and S = G[W,W] + dI, G = A Hinv A'; both
//     Hinv and G are working-set independent, so a working-set change is a gather from G plus an
//     O(|W|^2) bordering update of the explicit S^-1; every solve is iteratively refined against the
//     unregularised system                                                     (osqp polish.c:134-181);
//   later QPs of the instance hot-start phase 2 from the previous optimum and working set (the
//   analogue of qpOASES' hotstart, /root/reference/src/SubsolverQPOASES.cpp:154-160).
// A QP solution is only accepted when it satisfies the KKT conditions of the full QP.
//
// The file also compiles as plain single-threaded C++ (LCQP_HOST_EMU) -- a debugging aid used from
// scratch builds in the GPU-less dev container; the product never builds or calls that variant.
#pragma once

#include <math.h>
#include <stdint.h>
#include <string.h>
#ifdef LCQP_HOST_EMU
#include <stdio.h>
#include <stdlib.h>
#endif

#ifdef LCQP_HOST_EMU
#define LCQ_DEV
#define LCQ_DEVN
#define LCQ_TID 0
#define LCQ_NT 1
#define LCQ_LANE 0
#define LCQ_WARP 0
#define LCQ_NWARP 1
#define LCQ_LANES 1
#define LCQ_SYNC() ((void)0)
#else
#define LCQ_DEV __device__ __forceinline__
#define LCQ_DEVN __device__ __noinline__
#define LCQ_TID ((int)threadIdx.x)
#define LCQ_NT ((int)blockDim.x)
#define LCQ_LANE ((int)(threadIdx.x & 31))
#define LCQ_WARP ((int)(threadIdx.x >> 5))
#define LCQ_NWARP ((int)(blockDim.x >> 5))
#define LCQ_LANES 32
#define LCQ_SYNC() __syncthreads()
#endif

#include "../../include/lcqp_cuda.h"

namespace lcqp {

constexpr double kEPS = 2.221e-16;       // LCQPow Utilities::EPS (Utilities.hpp:350)
constexpr double kQPInf = 1e20;          // Utilities::INFTY (Utilities.hpp:362)
constexpr double kRhoMin = 1e-6;         // osqp constants.h:48
constexpr double kRhoTol = 1e-4;         // constants.h:50
constexpr double kRhoEqOverIneq = 1e3;   // constants.h:51
constexpr double kMinScaling = 1e-4;     // constants.h:94
constexpr double kMaxScaling = 1e4;
constexpr int kScalingIters = 10;        // constants.h:56
constexpr int kMaxLeyffer = 16;
constexpr double kResTol = 1e-9;         // residual of an EQP solve that still counts as solved

enum { RET_OK = 0, RET_INVALID_OSQP_BOX = 110, RET_INVALID_LOWER_COMP = 120, RET_MAX_ITER = 200, RET_MAX_PEN = 201,
       RET_SUBPROBLEM = 203, RET_OSQP_GUESS = 208 };

// Prepared (scaled) operands of the QP: depend on Q, A_full and the row types only.
struct Prep {
    double* P;     // n*n   D Q D
    double* A;     // m*n   E Ahat D
    double* D;     // n
    double* E;     // m
    double* Hinv;  // n*n   (P + delta I)^-1
    double* G;     // m*m   A Hinv A'
    double* Minv;  // n*n   (P + sigma I + A' R A)^-1
    double* T;     // m*n   scratch (A Hinv)
};

// Unscaled instance data (global memory), loadLCQP argument order.
struct Inst {
    const double *Q, *g, *L, *R, *lbL, *ubL, *lbR, *ubR, *A, *lbA, *ubA, *lb, *ub, *x0, *y0;
};

struct Dims {
    int n, nC, nComp, mA, m, has_box, cap;
};

// block-shared scalars
struct Scalars {
    double red[40];
    int ired[40];
    double bval;
    int bidx;
    int flag;
    int nw;
};

// Shared-memory working set of one instance.
struct Work {
    // QP (scaled space)
    double *q, *x, *xe, *xa, *px, *r1, *t1, *t2, *dx;            // n
    double *z, *y, *l, *u, *rhov, *lam, *r2, *dl, *dl2, *zt, *zt2, *w; // m
    signed char *W, *Wtry, *Wfail, *ctype, *pin;                // m
    int* idx;                                                   // m
    double* Sinv;                                               // cap*cap (shared or global)
    // accepted QP solution, unscaled
    double* xs;  // n
    double* ys;  // m   qpOASES sign
    // outer loop (unscaled)
    double *xk, *pk, *gk, *gt, *gphi, *stat, *tn;               // n
    double *Lx, *Rx;                                            // nComp
    Scalars* sc;
};

// ------------------------------------------------------------------------------------------------
// block-wide primitives
// ------------------------------------------------------------------------------------------------
LCQ_DEV double warp_sum(double v)
{
#ifndef LCQP_HOST_EMU
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
#endif
    return v;
}

LCQ_DEV double warp_max(double v)
{
#ifndef LCQP_HOST_EMU
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
#endif
    return v;
}

// all threads receive the block-wide sum / max of v
LCQ_DEV double block_sum(double v, Scalars* sc)
{
    v = warp_sum(v);
    LCQ_SYNC();
    if (LCQ_LANE == 0) sc->red[LCQ_WARP] = v;
    LCQ_SYNC();
    double s = 0;
    for (int k = 0; k < LCQ_NWARP; k++) s += sc->red[k];
    return s;
}

LCQ_DEV double block_max(double v, Scalars* sc)
{
    v = warp_max(v);
    LCQ_SYNC();
    if (LCQ_LANE == 0) sc->red[LCQ_WARP] = v;
    LCQ_SYNC();
    double s = sc->red[0];
    for (int k = 1; k < LCQ_NWARP; k++) s = fmax(s, sc->red[k]);
    return s;
}

// block-wide argmax with deterministic tie-break (smallest index); v must be >= threshold to count.
// Returns index (or -1) to all threads and the value through *vout.
LCQ_DEV int block_argmax(double v, int i, double* vout, Scalars* sc)
{
#ifndef LCQP_HOST_EMU
    for (int o = 16; o > 0; o >>= 1) {
        double ov = __shfl_xor_sync(0xffffffffu, v, o);
        int oi = __shfl_xor_sync(0xffffffffu, i, o);
        if (oi >= 0 && (i < 0 || ov > v || (ov == v && oi < i))) { v = ov; i = oi; }
    }
#endif
    LCQ_SYNC();
    if (LCQ_LANE == 0) { sc->red[LCQ_WARP] = v; sc->ired[LCQ_WARP] = i; }
    LCQ_SYNC();
    double bv = sc->red[0];
    int bi = sc->ired[0];
    for (int k = 1; k < LCQ_NWARP; k++) {
        double ov = sc->red[k];
        int oi = sc->ired[k];
        if (oi >= 0 && (bi < 0 || ov > bv || (ov == bv && oi < bi))) { bv = ov; bi = oi; }
    }
    *vout = bv;
    return bi;
}

// out[r] = sum_c M[r*ld + c] v[c]   (one warp per row, lanes over columns: coalesced row reads)
LCQ_DEV void mv_rows(const double* __restrict__ M, int rows, int cols, int ld, const double* v, double* out)
{
    for (int r = LCQ_WARP; r < rows; r += LCQ_NWARP) {
        const double* row = M + (size_t)r * ld;
        double s = 0;
        for (int c = LCQ_LANE; c < cols; c += LCQ_LANES) s += row[c] * v[c];
        s = warp_sum(s);
        if (LCQ_LANE == 0) out[r] = s;
    }
}

// out[c] = init[c] (or 0) + scale * sum_r M[r*ld + c] w[r]   (one thread per column: coalesced across threads)
LCQ_DEV void mv_cols(const double* __restrict__ M, int rows, int cols, int ld, const double* w, const double* init, double scale, double* out)
{
    for (int c = LCQ_TID; c < cols; c += LCQ_NT) {
        double s = 0;
        for (int r = 0; r < rows; r++) s += M[(size_t)r * ld + c] * w[r];
        out[c] = (init ? init[c] : 0.0) + scale * s;
    }
}

// out[a] = sum_c M[idx[a]*ld + c] v[c]  for a < na  (rows selected by idx)
LCQ_DEV void mv_rows_idx(const double* __restrict__ M, const int* idx, int na, int cols, int ld, const double* v, double* out)
{
    for (int a = LCQ_WARP; a < na; a += LCQ_NWARP) {
        const double* row = M + (size_t)idx[a] * ld;
        double s = 0;
        for (int c = LCQ_LANE; c < cols; c += LCQ_LANES) s += row[c] * v[c];
        s = warp_sum(s);
        if (LCQ_LANE == 0) out[a] = s;
    }
}

// out[c] = init[c] + scale * sum_a M[idx[a]*ld + c] w[a]
LCQ_DEV void mv_cols_idx(const double* __restrict__ M, const int* idx, int na, int cols, int ld, const double* w, const double* init, double scale, double* out)
{
    for (int c = LCQ_TID; c < cols; c += LCQ_NT) {
        double s = 0;
        for (int a = 0; a < na; a++) s += M[(size_t)idx[a] * ld + c] * w[a];
        out[c] = (init ? init[c] : 0.0) + scale * s;
    }
}

LCQ_DEV double limit_scaling(double v)
{
    if (v < kMinScaling) return 1.0;  // osqp scaling.c:22-30
    if (v > kMaxScaling) return kMaxScaling;
    return v;
}

// In-place inversion of the SPD n x n matrix M (row-major, global memory) by Gauss-Jordan without
// pivoting; colbuf/rowbuf are n-vectors in shared memory.  Returns 0, or 1 on a non-positive pivot.
LCQ_DEVN int spd_invert_inplace(double* M, int n, double* colbuf, double* rowbuf, Scalars* sc)
{
    for (int k = 0; k < n; k++) {
        LCQ_SYNC();
        const double p = M[(size_t)k * n + k];
        if (!(p > 0.0)) return 1;  // uniform: every thread reads the same value
        for (int j = LCQ_TID; j < n; j += LCQ_NT) {
            colbuf[j] = M[(size_t)j * n + k];
            rowbuf[j] = (j == k ? 1.0 : M[(size_t)k * n + j]) / p;
        }
        LCQ_SYNC();
        for (int e = LCQ_TID; e < n * n; e += LCQ_NT) {
            const int i = e / n, j = e - i * n;
            double v;
            if (i == k) v = rowbuf[j];
            else v = (j == k ? 0.0 : M[e]) - colbuf[i] * rowbuf[j];
            M[e] = v;
        }
    }
    LCQ_SYNC();
    (void)sc;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// prepare(): Ruiz equilibration (osqp scaling.c:44-156 without the q-dependent cost scale, c = 1),
// Hinv, Minv, G.  Row types (for rho_vec, auxil.c:76-98) come from `ctype`.
// Uses Work vectors t1,t2 (n) and zt (m) as scratch.
// ------------------------------------------------------------------------------------------------
LCQ_DEVN void prepare_scale(const Dims& d, const Inst& in, const Prep& pr, Work& w)
{
    const int n = d.n, m = d.m, mA = d.mA, nC = d.nC, nComp = d.nComp;
    // P = Q, A = [A; L; R; I]
    for (int e = LCQ_TID; e < n * n; e += LCQ_NT) pr.P[e] = in.Q[e];
    for (int e = LCQ_TID; e < m * n; e += LCQ_NT) {
        const int i = e / n, j = e - i * n;
        double v;
        if (i < nC) v = in.A[e];
        else if (i < nC + nComp) v = in.L[(size_t)(i - nC) * n + j];
        else if (i < mA) v = in.R[(size_t)(i - nC - nComp) * n + j];
        else v = (i - mA == j) ? 1.0 : 0.0;
        pr.A[e] = v;
    }
    for (int j = LCQ_TID; j < n; j += LCQ_NT) pr.D[j] = 1.0;
    for (int i = LCQ_TID; i < m; i += LCQ_NT) pr.E[i] = 1.0;
    double* Dt = w.t1;
    double* Et = w.zt;
    for (int it = 0; it < kScalingIters; it++) {
        LCQ_SYNC();
        for (int j = LCQ_TID; j < n; j += LCQ_NT) {
            double v = 0;
            for (int i = 0; i < n; i++) v = fmax(v, fabs(pr.P[(size_t)i * n + j]));
            for (int i = 0; i < m; i++) v = fmax(v, fabs(pr.A[(size_t)i * n + j]));
            Dt[j] = 1.0 / sqrt(limit_scaling(v));
        }
        for (int i = LCQ_WARP; i < m; i += LCQ_NWARP) {
            double v = 0;
            for (int j = LCQ_LANE; j < n; j += LCQ_LANES) v = fmax(v, fabs(pr.A[(size_t)i * n + j]));
            v = warp_max(v);
            if (LCQ_LANE == 0) Et[i] = 1.0 / sqrt(limit_scaling(v));
        }
        LCQ_SYNC();
        for (int e = LCQ_TID; e < n * n; e += LCQ_NT) {
            const int i = e / n, j = e - i * n;
            pr.P[e] *= Dt[i] * Dt[j];
        }
        for (int e = LCQ_TID; e < m * n; e += LCQ_NT) {
            const int i = e / n, j = e - i * n;
            pr.A[e] *= Et[i] * Dt[j];
        }
        for (int j = LCQ_TID; j < n; j += LCQ_NT) pr.D[j] *= Dt[j];
        for (int i = LCQ_TID; i < m; i += LCQ_NT) pr.E[i] *= Et[i];
    }
    LCQ_SYNC();
}

LCQ_DEVN int prepare_factor(const Dims& d, const Prep& pr, const signed char* ctype, const lcqp_cuda_options& o, Work& w)
{
    const int n = d.n, m = d.m;
    // Hinv
    for (int e = LCQ_TID; e < n * n; e += LCQ_NT) {
        const int i = e / n, j = e - i * n;
        pr.Hinv[e] = pr.P[e] + (i == j ? o.qp_delta : 0.0);
    }
    if (spd_invert_inplace(pr.Hinv, n, w.t1, w.t2, w.sc)) return 1;
    // Minv
    for (int i = LCQ_TID; i < m; i += LCQ_NT)
        w.zt[i] = ctype[i] < 0 ? kRhoMin : (ctype[i] == 1 ? kRhoEqOverIneq * o.qp_rho : o.qp_rho);
    LCQ_SYNC();
    for (int e = LCQ_TID; e < n * n; e += LCQ_NT) {
        const int i = e / n, j = e - i * n;
        double s = pr.P[e] + (i == j ? o.qp_sigma : 0.0);
        for (int r = 0; r < m; r++) s += w.zt[r] * pr.A[(size_t)r * n + i] * pr.A[(size_t)r * n + j];
        pr.Minv[e] = s;
    }
    if (spd_invert_inplace(pr.Minv, n, w.t1, w.t2, w.sc)) return 1;
    // T = A Hinv, G = T A'
    for (int e = LCQ_TID; e < m * n; e += LCQ_NT) {
        const int i = e / n, j = e - i * n;
        double s = 0;
        for (int k = 0; k < n; k++) s += pr.A[(size_t)i * n + k] * pr.Hinv[(size_t)k * n + j];
        pr.T[e] = s;
    }
    LCQ_SYNC();
    for (int e = LCQ_TID; e < m * m; e += LCQ_NT) {
        const int i = e / m, r = e - i * m;
        if (r > i) continue;
        double s = 0;
        for (int k = 0; k < n; k++) s += pr.T[(size_t)i * n + k] * pr.A[(size_t)r * n + k];
        pr.G[(size_t)i * m + r] = s;
        pr.G[(size_t)r * m + i] = s;
    }
    LCQ_SYNC();
    return 0;
}

// ------------------------------------------------------------------------------------------------
// QP solver state machine
// ------------------------------------------------------------------------------------------------
struct QP {
    Dims d;
    Prep pr;
    Work w;
    const lcqp_cuda_options* o;
    int nw;           // rows in the working set = order of Sinv (idx[0..nw))
    int have_W;
    int sinv_valid;   // Sinv matches idx/W
    double eqp_res;
    long long n_admm, n_eqp, n_changes;
};

// Bounds of A_full = [A; L; R] (+ box rows) as setConstraints / setComplementarityBounds build them
// (/root/reference/src/LCQProblem.cpp:584-608, 745-782), scaled by E, with OSQP's row types
// (auxil.c:76-98).  Returns bit0: some l > u, bit1: a complementarity lower bound is -inf (:747,:767).
LCQ_DEVN int qp_set_bounds(QP& s, const Inst& in)
{
    const Dims& d = s.d;
    Work& w = s.w;
    int bad = 0;
    for (int i = LCQ_TID; i < d.m; i += LCQ_NT) {
        double lo, up;
        if (i < d.nC) { lo = in.lbA ? in.lbA[i] : -INFINITY; up = in.ubA ? in.ubA[i] : INFINITY; }
        else if (i < d.nC + d.nComp) {
            const int k = i - d.nC;
            lo = in.lbL ? in.lbL[k] : 0.0; up = in.ubL ? in.ubL[k] : INFINITY;
            if (in.lbL && in.lbL[k] <= -INFINITY) bad |= 2;
        } else if (i < d.mA) {
            const int k = i - d.nC - d.nComp;
            lo = in.lbR ? in.lbR[k] : 0.0; up = in.ubR ? in.ubR[k] : INFINITY;
            if (in.lbR && in.lbR[k] <= -INFINITY) bad |= 2;
        } else { lo = in.lb ? in.lb[i - d.mA] : -INFINITY; up = in.ub ? in.ub[i - d.mA] : INFINITY; }
        if (lo > up) bad |= 1;
        const bool linf = !(lo > -kQPInf), uinf = !(up < kQPInf);
        const double E = s.pr.E[i];
        w.l[i] = linf ? -INFINITY : lo * E;
        w.u[i] = uinf ? INFINITY : up * E;
        signed char t = 0;
        if (linf && uinf) t = -1;
        else if (!linf && !uinf && w.u[i] - w.l[i] < kRhoTol) t = 1;
        w.ctype[i] = t;
        w.rhov[i] = t < 0 ? kRhoMin : (t == 1 ? kRhoEqOverIneq * s.o->qp_rho : s.o->qp_rho);
    }
    const int b0 = block_max((double)(bad & 1), w.sc) > 0.5;
    const int b1 = block_max((double)((bad >> 1) & 1), w.sc) > 0.5;
    return b0 | (b1 << 1);
}

// One ADMM iteration (osqp auxil.c:161-225, condensed KKT solve).
LCQ_DEVN void admm_iter(QP& s)
{
    const int n = s.d.n, m = s.d.m;
    Work& w = s.w;
    const double alpha = s.o->qp_alpha, sigma = s.o->qp_sigma;
    for (int i = LCQ_TID; i < m; i += LCQ_NT) w.w[i] = w.rhov[i] * w.z[i] - w.y[i];
    for (int j = LCQ_TID; j < n; j += LCQ_NT) w.t1[j] = sigma * w.x[j] - w.q[j];
    LCQ_SYNC();
    mv_cols(s.pr.A, m, n, n, w.w, w.t1, 1.0, w.t2);  // rhs = sigma x - q + A'w
    LCQ_SYNC();
    mv_rows(s.pr.Minv, n, n, n, w.t2, w.xe);         // xt
    LCQ_SYNC();
    mv_rows(s.pr.A, m, n, n, w.xe, w.zt);            // zt = A xt
    for (int j = LCQ_TID; j < n; j += LCQ_NT) w.x[j] = alpha * w.xe[j] + (1.0 - alpha) * w.x[j];
    LCQ_SYNC();
    for (int i = LCQ_TID; i < m; i += LCQ_NT) {
        const double v = alpha * w.zt[i] + (1.0 - alpha) * w.z[i];
        double zn = v + w.y[i] / w.rhov[i];
        zn = fmin(fmax(zn, w.l[i]), w.u[i]);
        w.y[i] += w.rhov[i] * (v - zn);
        w.z[i] = zn;
    }
    LCQ_SYNC();
}

// Active-set guess from the ADMM iterate (osqp polish.c:33-49; equality rows always active).
LCQ_DEVN void guess_working_set(QP& s, signed char* W)
{
    Work& w = s.w;
    for (int i = LCQ_TID; i < s.d.m; i += LCQ_NT) {
        signed char v = 0;
        if (w.ctype[i] == 1) v = 1;
        else if (w.ctype[i] < 0) v = 0;
        else if (w.z[i] - w.l[i] < -w.y[i]) v = 1;
        else if (w.u[i] - w.z[i] < w.y[i]) v = 2;
        W[i] = v;
    }
    LCQ_SYNC();
}

// ---- explicit inverse of S = G[W,W] + delta I, maintained by bordering -------------------------
// Append row j of the constraint matrix to the working set (position nw).  A row that is numerically a
// combination of the rows already in the set -- Schur pivot kappa at the level of the regularisation,
// kappa <= 10 delta (1 + u'u) -- is NOT appended (the working set is kept linearly independent, as the
// active-set theory requires); returns 1 in that case, 0 otherwise.
LCQ_DEVN int sinv_append(QP& s, int j)
{
    Work& w = s.w;
    const int nw = s.nw, cap = s.d.cap, m = s.d.m;
    double* Si = w.Sinv;
    const double* Gj = s.pr.G + (size_t)j * m;
    // sv = G[j, idx[a]]  ->  dl2 ;  u = Sinv sv -> dl
    for (int a = LCQ_TID; a < nw; a += LCQ_NT) w.dl2[a] = Gj[w.idx[a]];
    LCQ_SYNC();
    mv_rows(Si, nw, nw, cap, w.dl2, w.dl);
    LCQ_SYNC();
    double part = 0, part2 = 0;
    for (int a = LCQ_TID; a < nw; a += LCQ_NT) { part += w.dl2[a] * w.dl[a]; part2 += w.dl[a] * w.dl[a]; }
    const double su = block_sum(part, w.sc);
    const double uu = block_sum(part2, w.sc);
    const double kappa = Gj[j] + s.o->qp_delta - su;
    if (!(kappa > 10.0 * s.o->qp_delta * (1.0 + uu))) return 1;
    const double ik = 1.0 / kappa;
    for (int e = LCQ_TID; e < nw * nw; e += LCQ_NT) {
        const int a = e / nw, b = e - a * nw;
        Si[(size_t)a * cap + b] += w.dl[a] * w.dl[b] * ik;
    }
    for (int a = LCQ_TID; a < nw; a += LCQ_NT) {
        Si[(size_t)a * cap + nw] = -w.dl[a] * ik;
        Si[(size_t)nw * cap + a] = -w.dl[a] * ik;
    }
    if (LCQ_TID == 0) { Si[(size_t)nw * cap + nw] = ik; w.idx[nw] = j; }
    s.nw = nw + 1;
    LCQ_SYNC();
    return 0;
}

// remove position a from the working set (last position is moved into a)
LCQ_DEVN void sinv_remove(QP& s, int a)
{
    Work& w = s.w;
    const int nw = s.nw, cap = s.d.cap, last = nw - 1;
    double* Si = w.Sinv;
    const double c = Si[(size_t)a * cap + a];
    for (int b = LCQ_TID; b < nw; b += LCQ_NT) w.dl2[b] = Si[(size_t)b * cap + a];
    LCQ_SYNC();
    const double ic = 1.0 / c;
    for (int e = LCQ_TID; e < nw * nw; e += LCQ_NT) {
        const int r = e / nw, b = e - r * nw;
        Si[(size_t)r * cap + b] -= w.dl2[r] * w.dl2[b] * ic;
    }
    LCQ_SYNC();
    if (a != last) {
        // move row/col `last` into position a
        for (int b = LCQ_TID; b < nw; b += LCQ_NT) w.dl2[b] = Si[(size_t)last * cap + b];
        LCQ_SYNC();
        for (int b = LCQ_TID; b < last; b += LCQ_NT) {
            const double v = (b == a) ? w.dl2[last] : w.dl2[b];
            Si[(size_t)a * cap + b] = v;
            Si[(size_t)b * cap + a] = v;
        }
        if (LCQ_TID == 0) w.idx[a] = w.idx[last];
    }
    s.nw = last;
    LCQ_SYNC();
}

// Rebuild idx/Sinv from W by successive bordering; rows found linearly dependent on the rows before
// them are taken out of W.  Returns 1 if W has more rows than Sinv can hold.
LCQ_DEVN int sinv_build(QP& s, signed char* W)
{
    const int m = s.d.m;
    s.nw = 0;
    s.sinv_valid = 0;
    int cnt = 0;
    for (int i = LCQ_TID; i < m; i += LCQ_NT) cnt += (W[i] != 0);
    const int total = (int)(block_sum((double)cnt, s.w.sc) + 0.5);
    (void)total;
    // equality rows first, then the rest in row order (W is in shared memory: uniform branches)
    for (int pass = 0; pass < 2; pass++)
        for (int i = 0; i < m; i++) {
            if (!W[i] || ((s.w.ctype[i] == 1) != (pass == 0))) continue;
            if (s.nw >= s.d.cap || sinv_append(s, i)) {
                LCQ_SYNC();
                if (LCQ_TID == 0) W[i] = 0;
                LCQ_SYNC();
            }
        }
    s.sinv_valid = 1;
    return 0;
}

// EQP on the working set (idx[0..nw), sides in W): regularised KKT + refinement (see file header).
// Result: w.xe (x), w.lam (compact multipliers, OSQP sign).  s.eqp_res = final residual.
LCQ_DEVN void eqp_solve(QP& s, const signed char* W)
{
    const int n = s.d.n, nw = s.nw, cap = s.d.cap;
    Work& w = s.w;
    const Prep& pr = s.pr;
    s.n_eqp++;
    for (int j = LCQ_TID; j < n; j += LCQ_NT) { w.xe[j] = 0.0; w.dx[j] = 0.0; }
    for (int a = LCQ_TID; a < nw; a += LCQ_NT) { w.lam[a] = 0.0; w.dl[a] = 0.0; }
    LCQ_SYNC();
    double best = INFINITY;
    for (int pass = 0;; pass++) {
        // r1 = -q - P x - Aw' lam ; r2 = b - Aw x
        mv_rows(pr.P, n, n, n, w.xe, w.px);
        mv_rows_idx(pr.A, w.idx, nw, n, n, w.xe, w.r2);
        LCQ_SYNC();
        for (int j = LCQ_TID; j < n; j += LCQ_NT) w.t1[j] = -w.q[j] - w.px[j];
        double rn = 0;
        for (int a = LCQ_TID; a < nw; a += LCQ_NT) {
            const int i = w.idx[a];
            const double b = (W[i] == 1) ? w.l[i] : w.u[i];
            w.r2[a] = b - w.r2[a];
            rn = fmax(rn, fabs(w.r2[a]));
        }
        LCQ_SYNC();
        mv_cols_idx(pr.A, w.idx, nw, n, n, w.lam, w.t1, -1.0, w.r1);
        LCQ_SYNC();
        for (int j = LCQ_TID; j < n; j += LCQ_NT) rn = fmax(rn, fabs(w.r1[j]));
        rn = block_max(rn, w.sc);
        if (pass > 0 && !(rn < best)) {  // last correction did not help: undo it and stop
            for (int j = LCQ_TID; j < n; j += LCQ_NT) w.xe[j] -= w.dx[j];
            for (int a = LCQ_TID; a < nw; a += LCQ_NT) w.lam[a] -= w.dl[a];
            break;
        }
        best = rn;
        if (rn < 1e-15 || pass == s.o->qp_refine_iter) break;
        // dlam = Sinv (Aw Hinv r1 - r2) ; dx = Hinv (r1 - Aw' dlam)
        mv_rows(pr.Hinv, n, n, n, w.r1, w.t1);
        LCQ_SYNC();
        mv_rows_idx(pr.A, w.idx, nw, n, n, w.t1, w.dl2);
        LCQ_SYNC();
        for (int a = LCQ_TID; a < nw; a += LCQ_NT) w.dl2[a] -= w.r2[a];
        LCQ_SYNC();
        mv_rows(w.Sinv, nw, nw, cap, w.dl2, w.dl);
        LCQ_SYNC();
        mv_cols_idx(pr.A, w.idx, nw, n, n, w.dl, w.r1, -1.0, w.t2);
        LCQ_SYNC();
        mv_rows(pr.Hinv, n, n, n, w.t2, w.dx);
        LCQ_SYNC();
        for (int j = LCQ_TID; j < n; j += LCQ_NT) w.xe[j] += w.dx[j];
        for (int a = LCQ_TID; a < nw; a += LCQ_NT) w.lam[a] += w.dl[a];
        LCQ_SYNC();
    }
    LCQ_SYNC();
    s.eqp_res = best;
}

// KKT conditions of the full QP at (xe, lam).  0 = satisfied; 2 stationarity, 3 active row off its
// bound, 4 inactive row violated, 5/6 wrong multiplier sign (*worst = position in idx to drop).
LCQ_DEVN int kkt_check(QP& s, const signed char* W, int* worst)
{
    const int n = s.d.n, m = s.d.m, nw = s.nw;
    Work& w = s.w;
    const Prep& pr = s.pr;
    const double ftol = s.o->qp_feas_tol, dtol = s.o->qp_dual_tol;
    mv_rows(pr.P, n, n, n, w.xe, w.px);
    mv_rows(pr.A, m, n, n, w.xe, w.zt);
    LCQ_SYNC();
    for (int j = LCQ_TID; j < n; j += LCQ_NT) w.t1[j] = -w.q[j] - w.px[j];
    double ln = 0;
    for (int a = LCQ_TID; a < nw; a += LCQ_NT) ln = fmax(ln, fabs(w.lam[a]));
    LCQ_SYNC();
    mv_cols_idx(pr.A, w.idx, nw, n, n, w.lam, w.t1, -1.0, w.r1);
    ln = block_max(ln, w.sc);
    double rs = 0;
    for (int j = LCQ_TID; j < n; j += LCQ_NT) rs = fmax(rs, fabs(w.r1[j]));
    rs = block_max(rs, w.sc);
    if (!(rs <= kResTol * (1.0 + ln))) return 2;
    int bad3 = 0, bad4 = 0;
    for (int a = LCQ_TID; a < nw; a += LCQ_NT) {
        const int i = w.idx[a];
        const double b = (W[i] == 1) ? w.l[i] : w.u[i];
        if (!(fabs(b - w.zt[i]) <= kResTol * (1.0 + fabs(b)))) bad3 = 1;
    }
    for (int i = LCQ_TID; i < m; i += LCQ_NT) {
        const double tol = ftol * (1.0 + fabs(w.zt[i]));
        if (w.zt[i] < w.l[i] - tol || w.zt[i] > w.u[i] + tol) bad4 = 1;
    }
    const double bad = block_max((double)(bad3 * 2 + bad4), w.sc);
    if (bad >= 2.0) return 3;
    if (bad >= 1.0) return 4;
    const double thr = dtol * (1.0 + ln);
    double bv = -1.0;
    int bi = -1;
    for (int a = LCQ_TID; a < nw; a += LCQ_NT) {
        const int i = w.idx[a];
        if (w.ctype[i] == 1 || w.pin[i]) continue;
        const double v = (W[i] == 1) ? w.lam[a] : -w.lam[a];  // OSQP sign: lower-active needs lam <= 0
        if (v > thr && (bi < 0 || v > bv)) { bv = v; bi = a; }
    }
    double vout;
    const int pos = block_argmax(bv, bi, &vout, w.sc);
    if (pos < 0) return 0;
    if (worst) *worst = pos;
    return (W[w.idx[pos]] == 1) ? 5 : 6;
}

LCQ_DEVN void accept_solution(QP& s, const signed char* W)
{
    const int n = s.d.n, m = s.d.m, nw = s.nw;
    Work& w = s.w;
    const Prep& pr = s.pr;
    if (W != w.W)
        for (int i = LCQ_TID; i < m; i += LCQ_NT) w.W[i] = W[i];
    for (int i = LCQ_TID; i < m; i += LCQ_NT) w.y[i] = 0.0;
    LCQ_SYNC();
    for (int a = LCQ_TID; a < nw; a += LCQ_NT) w.y[w.idx[a]] = w.lam[a];
    for (int j = LCQ_TID; j < n; j += LCQ_NT) {
        w.xs[j] = pr.D[j] * w.xe[j];  // un-scale (osqp auxil.c:524-562), c = 1
        w.x[j] = w.xe[j];
    }
    LCQ_SYNC();
    for (int i = LCQ_TID; i < m; i += LCQ_NT) w.ys[i] = -(pr.E[i] * w.y[i]);  // qpOASES sign
    mv_rows(pr.A, m, n, n, w.x, w.z);
    s.have_W = 1;
    LCQ_SYNC();
}

// H-metric projection of xin onto the rows of the working set + feasibility test of every row.
LCQ_DEVN int project_feasible(QP& s, const signed char* W, const double* xin, double* xout)
{
    const int n = s.d.n, m = s.d.m, nw = s.nw, cap = s.d.cap;
    Work& w = s.w;
    const Prep& pr = s.pr;
    for (int j = LCQ_TID; j < n; j += LCQ_NT) xout[j] = xin[j];
    LCQ_SYNC();
    for (int pass = 0; pass < 4; pass++) {
        mv_rows_idx(pr.A, w.idx, nw, n, n, xout, w.dl2);
        LCQ_SYNC();
        double rn = 0;
        for (int a = LCQ_TID; a < nw; a += LCQ_NT) {
            const int i = w.idx[a];
            w.dl2[a] = ((W[i] == 1) ? w.l[i] : w.u[i]) - w.dl2[a];
            rn = fmax(rn, fabs(w.dl2[a]));
        }
        rn = block_max(rn, w.sc);
        if (rn < 1e-15) break;
        mv_rows(w.Sinv, nw, nw, cap, w.dl2, w.dl);
        LCQ_SYNC();
        mv_cols_idx(pr.A, w.idx, nw, n, n, w.dl, nullptr, 1.0, w.t2);
        LCQ_SYNC();
        mv_rows(pr.Hinv, n, n, n, w.t2, w.dx);
        LCQ_SYNC();
        for (int j = LCQ_TID; j < n; j += LCQ_NT) xout[j] += w.dx[j];
        LCQ_SYNC();
    }
    mv_rows(pr.A, m, n, n, xout, w.zt);
    LCQ_SYNC();
    const double ftol = s.o->qp_feas_tol;
    int bad = 0;
    for (int i = LCQ_TID; i < m; i += LCQ_NT) {
        const double tol = ftol * (1.0 + fabs(w.zt[i]));
        if (w.zt[i] < w.l[i] - tol || w.zt[i] > w.u[i] + tol) bad = 1;
        if (W[i]) {
            const double b = (W[i] == 1) ? w.l[i] : w.u[i];
            if (fabs(w.zt[i] - b) > tol) bad = 1;
        }
    }
    return block_max((double)bad, w.sc) < 0.5;
}

// Primal active-set iteration from the feasible point x whose active rows contain the working set
// (idx/Sinv valid for W).  Returns 0 with the solution accepted, 1 if it gave up.
LCQ_DEVN int active_set(QP& s, double* x, signed char* W)
{
    const int n = s.d.n, m = s.d.m;
    Work& w = s.w;
    const Prep& pr = s.pr;
    const int cap_it = 20 * (n + m) + 100;
    int last_dropped = -1;
    for (int it = 0; it < cap_it; it++) {
        eqp_solve(s, W);
        // direction to the EQP minimiser, ratio test (Nocedal & Wright alg. 16.3)
        for (int j = LCQ_TID; j < n; j += LCQ_NT) w.dx[j] = w.xe[j] - x[j];
        LCQ_SYNC();
        mv_rows(pr.A, m, n, n, x, w.w);       // A x
        mv_rows(pr.A, m, n, n, w.dx, w.zt2);  // A p
        LCQ_SYNC();
        double apn = 0;
        for (int i = LCQ_TID; i < m; i += LCQ_NT) apn = fmax(apn, fabs(w.zt2[i]));
        apn = block_max(apn, w.sc);
        const double seps = 1e-13 * (1.0 + apn);
        // per-thread best: smallest alpha, tie -> largest |s|, then smallest row
        double ba = 2.0, bs = 0.0;
        int bi = -1;
        for (int i = LCQ_TID; i < m; i += LCQ_NT) {
            if (W[i] || w.ctype[i] < 0) continue;
            const double sv = w.zt2[i];
            double a = 2.0;
            if (sv < -seps && w.l[i] > -INFINITY) a = fmax(w.w[i] - w.l[i], 0.0) / (-sv);
            else if (sv > seps && w.u[i] < INFINITY) a = fmax(w.u[i] - w.w[i], 0.0) / sv;
            else continue;
            if (a < 1.0 && (bi < 0 || a < ba || (a == ba && fabs(sv) > bs))) { ba = a; bs = fabs(sv); bi = i; }
        }
        // two-stage reduction: first the minimal alpha, then among the rows attaining it the largest |s|
        const double amin = -block_max(bi >= 0 ? -ba : -2.0, w.sc);
        int block = -1;
        if (amin < 1.0) {
            double vout;
            block = block_argmax((bi >= 0 && ba == amin) ? bs : -1.0, (bi >= 0 && ba == amin) ? bi : -1, &vout, w.sc);
        }
        if (block >= 0) {
            const signed char side = (w.zt2[block] < 0) ? 1 : 2;
            for (int j = LCQ_TID; j < n; j += LCQ_NT) x[j] += amin * w.dx[j];
            LCQ_SYNC();
            if (s.nw >= s.d.cap) return 1;
            if (sinv_append(s, block)) return 1;  // a blocking row cannot be dependent on W (A_W p = 0)
            // a row that comes straight back after a zero-length step was dropped on multiplier noise:
            // it is weakly active; exempt it from the sign test for the rest of this QP (anti-cycling)
            if (LCQ_TID == 0) { W[block] = side; if (block == last_dropped && amin * (1.0 + apn) <= 1e-12) w.pin[block] = 1; }
            last_dropped = -1;
            LCQ_SYNC();
            s.n_changes++;
#ifdef LCQP_HOST_EMU
            if (getenv("LCQP_EMU_DEBUG") && atoi(getenv("LCQP_EMU_DEBUG")) > 1) fprintf(stderr, "    emu as it=%d nw=%d add row %d side %d alpha=%.3e res=%.1e\n", it, s.nw, block, (int)side, amin, s.eqp_res);
#endif
            continue;
        }
        // full step: x is the EQP minimiser; check its multipliers
        for (int j = LCQ_TID; j < n; j += LCQ_NT) x[j] = w.xe[j];
        LCQ_SYNC();
        int worst = -1;
        const int reason = kkt_check(s, W, &worst);
#ifdef LCQP_HOST_EMU
        if (getenv("LCQP_EMU_DEBUG") && atoi(getenv("LCQP_EMU_DEBUG")) > 1) fprintf(stderr, "    emu as it=%d nw=%d full step reason=%d worst=%d res=%.1e\n", it, s.nw, reason, worst, s.eqp_res);
#endif
        if (reason == 0) { accept_solution(s, W); return 0; }
        if ((reason == 5 || reason == 6) && worst >= 0) {
            const int row = w.idx[worst];
            LCQ_SYNC();
            if (LCQ_TID == 0) W[row] = 0;
            sinv_remove(s, worst);
            last_dropped = row;
            s.n_changes++;
            continue;
        }
#ifdef LCQP_HOST_EMU
        if (getenv("LCQP_EMU_DEBUG")) fprintf(stderr, "    emu as it=%d nw=%d gives up: reason %d res=%.1e\n", it, s.nw, reason, s.eqp_res);
#endif
        return 1;  // EQP not solvable to tolerance on this set
    }
    return 1;
}

// SubsolverBase::solve contract (/root/reference/include/SubsolverBase.hpp:37-56).  g unscaled (shared
// or global memory), x0 / y0 (m entries, qpOASES sign) may be null.  Returns 0 or a non-zero flag;
// *iterations = ADMM iterations + working-set changes.
LCQ_DEVN int qp_solve(QP& s, bool initial, const double* g, const double* x0, const double* y0A, const double* y0box, int* iterations, bool infeasible)
{
    const int n = s.d.n, m = s.d.m, mA = s.d.mA;
    Work& w = s.w;
    const Prep& pr = s.pr;
    const lcqp_cuda_options& o = *s.o;
    *iterations = 0;
    if (infeasible) return 37;
    const long long ch0 = s.n_changes;
    for (int j = LCQ_TID; j < n; j += LCQ_NT) w.q[j] = pr.D[j] * g[j];  // osqp.c:752-779, c = 1
    for (int i = LCQ_TID; i < m; i += LCQ_NT) w.pin[i] = 0;
    LCQ_SYNC();
    if (initial) {
        // osqp_warm_start_x/_y (osqp.c:700-745)
        for (int j = LCQ_TID; j < n; j += LCQ_NT) w.x[j] = x0 ? x0[j] / pr.D[j] : 0.0;
        for (int i = LCQ_TID; i < m; i += LCQ_NT) {
            double yv = 0.0;
            if (i < mA) { if (y0A) yv = y0A[i]; }
            else if (y0box) yv = y0box[i - mA];
            w.y[i] = -yv / pr.E[i];
        }
        LCQ_SYNC();
        mv_rows(pr.A, m, n, n, w.x, w.z);
        s.have_W = 0;
        s.sinv_valid = 0;
        LCQ_SYNC();
    } else if (s.have_W && s.sinv_valid) {
        for (int j = LCQ_TID; j < n; j += LCQ_NT) w.xa[j] = w.x[j];
        for (int i = LCQ_TID; i < m; i += LCQ_NT) w.Wtry[i] = w.W[i];
        LCQ_SYNC();
        if (active_set(s, w.xa, w.Wtry) == 0) { *iterations = (int)(s.n_changes - ch0); return 0; }
        // fall through to ADMM from the previous solution
        s.sinv_valid = 0;
        mv_rows(pr.A, m, n, n, w.x, w.z);
        LCQ_SYNC();
    }
    int have_fail = 0;
    int it = 0;
    while (it < o.qp_max_iter) {
        for (int k = 0; k < o.qp_check_interval && it < o.qp_max_iter; k++, it++) admm_iter(s);
        s.n_admm += o.qp_check_interval;
        guess_working_set(s, w.Wtry);
        int same = 0;
        if (have_fail) {
            int diff = 0;
            for (int i = LCQ_TID; i < m; i += LCQ_NT) diff |= (w.Wtry[i] != w.Wfail[i]);
            same = block_max((double)diff, w.sc) < 0.5;
        }
        if (same) continue;
        for (int i = LCQ_TID; i < m; i += LCQ_NT) w.Wfail[i] = w.Wtry[i];
        have_fail = 1;
        LCQ_SYNC();
        if (sinv_build(s, w.Wtry)) continue;  // guess larger than the Schur complement capacity
        eqp_solve(s, w.Wtry);
        const int reason = kkt_check(s, w.Wtry, nullptr);
#ifdef LCQP_HOST_EMU
        if (getenv("LCQP_EMU_DEBUG")) fprintf(stderr, "  emu admm it=%d nw=%d probe reason=%d res=%.1e\n", it, s.nw, reason, s.eqp_res);
#endif
        if (reason == 0) { accept_solution(s, w.Wtry); *iterations = it + (int)(s.n_changes - ch0); return 0; }
        if (reason == 5 || reason == 6) {
            for (int j = LCQ_TID; j < n; j += LCQ_NT) w.xa[j] = w.xe[j];
            LCQ_SYNC();
            if (active_set(s, w.xa, w.Wtry) == 0) { *iterations = it + (int)(s.n_changes - ch0); return 0; }
            s.sinv_valid = 0;
        } else if (project_feasible(s, w.Wtry, w.x, w.xa)) {
            if (active_set(s, w.xa, w.Wtry) == 0) { *iterations = it + (int)(s.n_changes - ch0); return 0; }
            s.sinv_valid = 0;
        }
    }
    *iterations = it + (int)(s.n_changes - ch0);
    return -2;  // OSQP_MAX_ITER_REACHED class
}

// perturbStep RNG (shared with the oracle): splitmix64 finaliser keyed by (seed, instance, iterate, coordinate)
LCQ_DEV int perturb_draw(unsigned long long seed, unsigned long long instance, unsigned iter, unsigned i)
{
    uint64_t z = seed * 0x9E3779B97F4A7C15ull + instance * 0xBF58476D1CE4E5B9ull + (((uint64_t)iter << 32) | i);
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    return (int)(z % 3ull) - 1;
}

// ------------------------------------------------------------------------------------------------
// The penalty-homotopy loop for one instance (LCQProblem.cpp:444-560 and helpers :1105-1482)
// ------------------------------------------------------------------------------------------------
struct LoopOut {
    int ret, status, iterTotal, iterOuter, subIter, exitFlag;
    double rhoOpt;
};

// t = Q v ; Lx = L v ; Rx = R v ; out = t + rho * (L' Rx + R' Lx) + add   (i.e. Qk v + add)
LCQ_DEVN void qk_apply(const Dims& d, const Inst& in, double rho, const double* v, const double* add, double* out, Work& w)
{
    const int n = d.n, nComp = d.nComp;
    mv_rows(in.Q, n, n, n, v, w.tn);
    mv_rows(in.L, nComp, n, n, v, w.Lx);
    mv_rows(in.R, nComp, n, n, v, w.Rx);
    LCQ_SYNC();
    for (int c = LCQ_TID; c < n; c += LCQ_NT) {
        double sL = 0, sR = 0;
        for (int r = 0; r < nComp; r++) {
            sL += in.L[(size_t)r * n + c] * w.Rx[r];
            sR += in.R[(size_t)r * n + c] * w.Lx[r];
        }
        out[c] = w.tn[c] + rho * (sL + sR) + (add ? add[c] : 0.0);
    }
    LCQ_SYNC();
}

LCQ_DEVN void lcqp_loop(QP& s, const Inst& in, unsigned long long instance,
                       bool infeasible, double* xout, double* yout, LoopOut& out)
{
    const Dims& d = s.d;
    const int n = d.n, nC = d.nC, nComp = d.nComp, mA = d.mA;
    Work& w = s.w;
    const lcqp_cuda_options& o = *s.o;
    const bool osqp_flavour = (o.qpSolver == 2);
    const int boxOff = osqp_flavour ? 0 : n;

    double hist[kMaxLeyffer];
    int nh = 0;
    double alphak = 1.0, rho = o.initialPenaltyParameter, phi_const = 0.0;
    int outerIter = 0, totalIter = 0, subIter = 0, qpIter = 0, exitFlag = 0, status = 0, ret = RET_OK;
    out.rhoOpt = 0.0;
    const bool have_gphi = (in.lbL != nullptr) || (in.lbR != nullptr);

    for (int j = LCQ_TID; j < n; j += LCQ_NT) {
        w.xk[j] = in.x0 ? in.x0[j] : 0.0;   // LCQProblem.ipp:138-142
        w.gt[j] = in.g[j];                  // g_tilde = g (LCQProblem.cpp:966-967)
        w.pk[j] = 0.0;
    }
    if (have_gphi) {  // LCQProblem.cpp:970-996
        double part = 0;
        for (int i = LCQ_TID; i < nComp; i += LCQ_NT) part += (in.lbL ? in.lbL[i] : 0.0) * (in.lbR ? in.lbR[i] : 0.0);
        phi_const = block_sum(part, w.sc);
        for (int c = LCQ_TID; c < n; c += LCQ_NT) {
            double sv = 0;
            if (in.lbL) for (int r = 0; r < nComp; r++) sv += in.R[(size_t)r * n + c] * in.lbL[r];
            if (in.lbR) for (int r = 0; r < nComp; r++) sv += in.L[(size_t)r * n + c] * in.lbR[r];
            w.gphi[c] = -sv;
        }
    }
    LCQ_SYNC();

    auto phi = [&]() -> double {  // getPhi :1172-1185 ; x'Cx/2 = (Lx)'(Rx)
        mv_rows(in.L, nComp, n, n, w.xk, w.Lx);
        mv_rows(in.R, nComp, n, n, w.xk, w.Rx);
        LCQ_SYNC();
        double part = 0;
        for (int i = LCQ_TID; i < nComp; i += LCQ_NT) part += w.Lx[i] * w.Rx[i];
        if (have_gphi) for (int j = LCQ_TID; j < n; j += LCQ_NT) part += w.gphi[j] * w.xk[j];
        return phi_const + block_sum(part, w.sc);
    };
    auto update_penalty = [&]() {  // :1199-1214
        nh = 0;
        rho *= o.penaltyUpdateFactor;
        out.rhoOpt = rho;
        if (have_gphi) {
            for (int j = LCQ_TID; j < n; j += LCQ_NT) w.gt[j] = in.g[j] + rho * w.gphi[j];
            LCQ_SYNC();
        }
    };
    auto linearize = [&]() {  // updateLinearization :1105-1112 : gk = rho C xk + g_tilde
        mv_rows(in.L, nComp, n, n, w.xk, w.Lx);
        mv_rows(in.R, nComp, n, n, w.xk, w.Rx);
        LCQ_SYNC();
        for (int c = LCQ_TID; c < n; c += LCQ_NT) {
            double sL = 0, sR = 0;
            for (int r = 0; r < nComp; r++) {
                sL += in.L[(size_t)r * n + c] * w.Rx[r];
                sR += in.R[(size_t)r * n + c] * w.Lx[r];
            }
            w.gk[c] = rho * (sL + sR) + w.gt[c];
        }
        LCQ_SYNC();
    };
    auto solve_qp = [&](bool initial) -> bool {  // solveQPSubproblem :1115-1148
        const double* y0A = nullptr;
        const double* y0box = nullptr;
        if (initial && in.y0) { y0A = in.y0 + n; y0box = d.has_box ? in.y0 : nullptr; }
        const int fl = qp_solve(s, initial, w.gk, w.xk, y0A, y0box, &qpIter, infeasible);
        subIter += qpIter;
        exitFlag = osqp_flavour ? (fl == 0 ? 1 : fl) : fl;
        if (fl != 0) { ret = (osqp_flavour && infeasible) ? RET_OSQP_GUESS : RET_SUBPROBLEM; return false; }
        for (int j = LCQ_TID; j < n; j += LCQ_NT) w.pk[j] = w.xs[j] - w.xk[j];
        LCQ_SYNC();
        return true;
    };

    bool failed = false, success = false;
    // first QP (:452-467)
    if (o.solveZeroPenaltyFirst) {
        for (int j = LCQ_TID; j < n; j += LCQ_NT) w.gk[j] = in.g[j];
        LCQ_SYNC();
    } else {
        linearize();
    }
    if (!solve_qp(true)) failed = true;
    out.rhoOpt = failed ? 0.0 : rho;  // :473

    while (!failed) {
        // updateStep :1240-1243
        for (int j = LCQ_TID; j < n; j += LCQ_NT) w.xk[j] = w.xk[j] + alphak * w.pk[j];
        LCQ_SYNC();
        // updateStationarity :1246-1272 : stat = Qk xk + g_tilde - A_full' yk_A - yk_box
        qk_apply(d, in, rho, w.xk, w.gt, w.stat, w);
        for (int c = LCQ_TID; c < n; c += LCQ_NT) {
            double sv = 0;
            for (int r = 0; r < nC; r++) sv += in.A[(size_t)r * n + c] * w.ys[r];
            for (int r = 0; r < nComp; r++) sv += in.L[(size_t)r * n + c] * w.ys[nC + r];
            for (int r = 0; r < nComp; r++) sv += in.R[(size_t)r * n + c] * w.ys[nC + nComp + r];
            double st = w.stat[c] - sv;
            if (d.has_box) st -= w.ys[mA + c];
            w.stat[c] = st;
        }
        LCQ_SYNC();
        totalIter++;  // :493-496

        // leyfferCheckPositive :1275-1313
        {
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
        linearize();  // :508

        double sm = 0;
        for (int j = LCQ_TID; j < n; j += LCQ_NT) sm = fmax(sm, fabs(w.stat[j]));
        sm = block_max(sm, w.sc);
        if (sm < o.stationarityTolerance) {  // :511
            if (phi() < o.complementarityTolerance) {
                // determineStationarityType :1412-1453 on yk_A (the PENALISED duals, :1420),
                // weak set :1456-1482 (phi() left Lx, Rx in shared memory)
                const double tc = o.complementarityTolerance;
                int fl = 0;  // bit0: s fails, bit1: m fails, bit2: weakly stationary only
                for (int i = LCQ_TID; i < nComp; i += LCQ_NT) {
                    if (!(w.Lx[i] <= tc && w.Rx[i] <= tc)) continue;
                    const double yl = w.ys[nC + i], yr = w.ys[nC + nComp + i];
                    const double prod = yl * yr, mn = fmin(yl, yr);
                    if (mn < 0) fl |= 1;
                    if (fabs(prod) >= tc && mn <= 0) { if (prod <= tc) fl |= 4; else fl |= 2; }
                }
                // the reference returns W at the FIRST weak index whose test fails c-stationarity; any
                // such index gives W, so an OR over the weak set is the same decision
                int any = 0;
                for (int b = 0; b < 3; b++) if (block_max((double)((fl >> b) & 1), w.sc) > 0.5) any |= (1 << b);
                status = (any & 4) ? 1 : (!(any & 1) ? 4 : (!(any & 2) ? 3 : 2));
                success = true;
                break;
            } else {
                update_penalty();
                outerIter++;
            }
        }
        if (totalIter > o.maxIterations) { ret = RET_MAX_ITER; break; }  // :537
        if (rho > o.maxPenaltyParameter) { ret = RET_MAX_PEN; break; }   // :541

        linearize();                                // :545
        if (!solve_qp(false)) { failed = true; break; }  // :548

        if (o.perturbStep) {  // :553-555, :1353-1362
            for (int j = LCQ_TID; j < n; j += LCQ_NT)
                w.xk[j] += perturb_draw(o.perturb_seed, instance, (unsigned)totalIter, (unsigned)j) * kEPS;
            LCQ_SYNC();
        }
        // getOptimalStepLength :1217-1237
        {
            qk_apply(d, in, rho, w.pk, nullptr, w.stat, w);  // Qk pk
            double part = 0;
            for (int j = LCQ_TID; j < n; j += LCQ_NT) part += w.stat[j] * w.pk[j];
            const double qk = block_sum(part, w.sc);
            qk_apply(d, in, rho, w.xk, w.gt, w.stat, w);     // Qk xk + g_tilde
            part = 0;
            for (int j = LCQ_TID; j < n; j += LCQ_NT) part += w.stat[j] * w.pk[j];
            const double lk = block_sum(part, w.sc);
            alphak = 1.0;
            if (qk > 0 && lk < 0) alphak = fmin(-lk / qk, 1.0);
        }
    }

    // outputs: x = xk ; y = [box duals ; yk_A] (transformDuals :1381-1409 applied on success)
    for (int j = LCQ_TID; j < n; j += LCQ_NT) xout[j] = w.xk[j];
    if (!osqp_flavour)
        for (int j = LCQ_TID; j < n; j += LCQ_NT) yout[j] = (d.has_box && !infeasible) ? w.ys[mA + j] : 0.0;
    if (success) {
        // Lx, Rx from the last phi() call hold L xk, R xk
        for (int i = LCQ_TID; i < mA; i += LCQ_NT) {
            double v = w.ys[i];
            if (i >= nC && i < nC + nComp) v -= rho * w.Rx[i - nC];
            else if (i >= nC + nComp) v -= rho * w.Lx[i - nC - nComp];
            yout[boxOff + i] = v;
        }
    } else {
        for (int i = LCQ_TID; i < mA; i += LCQ_NT) yout[boxOff + i] = infeasible ? 0.0 : w.ys[i];
    }
    LCQ_SYNC();
    out.ret = ret;
    out.status = status;
    out.iterTotal = totalIter;
    out.iterOuter = outerIter;
    out.subIter = subIter;
    out.exitFlag = exitFlag;
}

// ------------------------------------------------------------------------------------------------
// shared-memory carving
// ------------------------------------------------------------------------------------------------
struct SmemPlan {
    size_t bytes;       // total dynamic shared memory
    int sinv_in_smem;
};

inline
#ifndef LCQP_HOST_EMU
__host__ __device__
#endif
size_t work_bytes(const Dims& d, bool sinv_in_smem)
{
    size_t nd = 0;
    nd += 9ull * d.n;       // q x xe xa px r1 t1 t2 dx
    nd += 12ull * d.m;      // z y l u rhov lam r2 dl dl2 zt zt2 w
    nd += 1ull * d.n + d.m; // xs ys
    nd += 7ull * d.n;       // xk pk gk gt gphi stat tn
    nd += 2ull * d.nComp;   // Lx Rx
    if (sinv_in_smem) nd += (size_t)d.cap * d.cap;
    size_t b = nd * sizeof(double);
    b += (size_t)d.m * sizeof(int);          // idx
    b += 5ull * ((d.m + 15) / 16) * 16;      // W Wtry Wfail ctype pin
    b += sizeof(Scalars) + 64;
    return b;
}

LCQ_DEV void carve(Work& w, const Dims& d, unsigned char* base, double* sinv_global)
{
    double* p = reinterpret_cast<double*>(base);
    auto takeN = [&](int k) { double* r = p; p += k; return r; };
    const int n = d.n, m = d.m;
    w.q = takeN(n); w.x = takeN(n); w.xe = takeN(n); w.xa = takeN(n); w.px = takeN(n); w.r1 = takeN(n);
    w.t1 = takeN(n); w.t2 = takeN(n); w.dx = takeN(n);
    w.z = takeN(m); w.y = takeN(m); w.l = takeN(m); w.u = takeN(m); w.rhov = takeN(m); w.lam = takeN(m);
    w.r2 = takeN(m); w.dl = takeN(m); w.dl2 = takeN(m); w.zt = takeN(m); w.zt2 = takeN(m); w.w = takeN(m);
    w.xs = takeN(n); w.ys = takeN(m);
    w.xk = takeN(n); w.pk = takeN(n); w.gk = takeN(n); w.gt = takeN(n); w.gphi = takeN(n); w.stat = takeN(n); w.tn = takeN(n);
    w.Lx = takeN(d.nComp); w.Rx = takeN(d.nComp);
    if (sinv_global) w.Sinv = sinv_global;
    else w.Sinv = takeN(d.cap * d.cap);
    w.idx = reinterpret_cast<int*>(p);
    signed char* c = reinterpret_cast<signed char*>(w.idx + m);
    const int mpad = ((m + 15) / 16) * 16;
    w.W = c; w.Wtry = c + mpad; w.Wfail = c + 2 * mpad; w.ctype = c + 3 * mpad; w.pin = c + 4 * mpad;
    uintptr_t sp = reinterpret_cast<uintptr_t>(c + 5 * mpad);
    sp = (sp + 15) & ~(uintptr_t)15;
    w.sc = reinterpret_cast<Scalars*>(sp);
}

}  // namespace lcqp

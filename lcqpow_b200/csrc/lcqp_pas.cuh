// lcqp_pas.cuh -- the convex-QP subsolver of the B200-native LCQP path: a PARAMETRIC ACTIVE-SET method in range
// space (sm_100a device code; compiles as single-threaded C++ under LCQP_HOST_EMU for the CPU test-suite).
//
// What it restates.  The reference hands every inner QP to qpOASES' online active-set strategy
// (/root/reference/src/SubsolverQPOASES.cpp:134-169): QProblem::init builds an auxiliary QP that has the initial
// guess as its solution and follows the homotopy to the true data; QProblem::hotstart follows the homotopy from
// the previous QP to the next gradient (/root/reference/external/qpOASES/src/QProblem.cpp:1301-1471, :446-640,
// :1477-1747).  The penalty loop above it never sees anything but the END POINT of that homotopy -- and where the
// QP has a (numerically) non-unique solution the end point is decided by the path: ratio tests with their
// tolerances and tie-breaking order, the termination test on the remaining homotopy length, the resolution of
// linearly dependent additions, ramping after zero steps, far bounds.  Those DECISIONS are restated here, each
// citing the lines it follows:
//     ratio tests        QProblem::performStep        QProblem.cpp:4981-5278,  QProblemB::performRatioTest QProblemB.cpp:2065
//     termination        getRelativeHomotopyLength    QProblem.cpp:5372-5410,  QProblemB.cpp:2113, solveQP :1640-1658
//     dependent addition addConstraint_ensureLI       QProblem.cpp:3117-3300
//     ramping            performRamping               QProblem.cpp:5416-5492 (after a zero step, :1707-1710, and in init :1437)
//     drift correction   performDriftCorrection       QProblem.cpp:5559-5652
//     far bounds         hotstart / updateFarBounds   QProblem.cpp:499-629, :5498-5552
//     initial homotopy   solveInitialQP, obtainAuxiliaryWorkingSet, setupAuxiliaryQPbounds   QProblem.cpp:1301-1471, :2199-2344, :2668-2813
// The LINEAR ALGEBRA is not qpOASES' (dense TQ factorisation + projected Cholesky per instance): it is a range-
// space method built for batches that share Q and A_full = [A; L; R; I_box]:
//     rows E with l = u in every instance are eliminated for good (always active, free multipliers):
//         Z  = orthonormal basis of null(A_E) (Householder),   P = Z (Z'QZ)^-1 Z'   (the inverse Hessian on null(A_E))
//         N0 = A_E'(A_E A_E')^-1,   N = (I - P Q) N0            (x = P(A_I'y - g) + N b_E solves the equality-constrained QP)
//         Gt = A_I P,   Tt = Gt A_I',   K = A_I N                (once per batch / per instance)
//     (the null-space route keeps Tt accurate when Q has directions of tiny curvature: the textbook
//      Hinv - Hinv A_E'(A_E Hinv A_E')^-1 A_E Hinv cancels 1/curvature-sized terms and loses them)
//     a homotopy step works on vectors over the remaining rows only:
//         dy_W = (Tt_WW)^-1 (db_W + dc_W),   dz = Tt[:,W] dy_W - dc,   c = Gt g - K b_E
//     with the inverse of Tt_WW kept explicitly per instance (bordering updates, O(|W|^2)).
// No n-dimensional work happens inside the homotopy loop: the gradient is tracked by the remaining fraction phi of
// the step g_new - g, the primal by x += P (A_I'dy - dg) + N db_E once per QP, polished in range space until the
// active rows hold to round-off.  Requires Z'QZ positive definite (Gauss-Jordan pivots > 1e-14 max diag); batches
// with a semidefinite reduced Hessian take the regularised primal active-set kernel of lcqp_device.cuh instead.
#pragma once

#include "lcqp_device.cuh"

namespace lcqp {
namespace pas {

// qpOASES constants / default options (Constants.hpp:50-61, Options.cpp:91-146)
constexpr double qEPS = 2.221e-16;
constexpr double qINFTY = 1.0e20;
// terminationTolerance (5e6 EPS) and boundTolerance (1e6 EPS) come from the options (qpoases_*): Options::setqpOASESOptions
constexpr double kBoundRelax = 1.0e4;            // boundRelaxation
constexpr double kEpsNum = -1.0e3 * qEPS;        // epsNum
constexpr double kEpsDen = 1.0e3 * qEPS;         // epsDen
constexpr double kMaxDualJump = 1.0e8;           // maxDualJump
constexpr double kRamp0 = 0.5, kRamp1 = 1.0;     // initialRamping, finalRamping
constexpr double kFar0 = 1.0e6, kFarGrow = 1.0e3;  // initialFarBounds, growFarBounds
constexpr double kLITol = 1.0e-10;               // bordering pivot / diagonal below which a row counts as dependent
constexpr double kPDTol = 1.0e-14;               // Gauss-Jordan pivot / max diagonal below which Q is not "positive definite"
constexpr double kPolishTol = 1.0e-15;
constexpr int kPolishMax = 8;

enum { ST_INACTIVE = 0, ST_LOWER = 1, ST_UPPER = -1 };
enum { QP_OK = 0, QP_INFEASIBLE_BOUNDS = 31, QP_INFEASIBLE = 37, QP_UNBOUNDED = 38, QP_MAXITER = 64, QP_SETUP = 33 };

// ------------------------------------------------------------------------------------------------
// Prepared operands.  rows: m = nC + 2 nComp (+ n box rows); E = eliminated equality rows, I = the others.
// ------------------------------------------------------------------------------------------------
struct PMats {
    // work arrays of the preparation (row-major, natural leading dimensions)
    double* P;       // n*n      Z (Z'QZ)^-1 Z'
    double* N;       // n*mE     (I - P Q) N0
    double* N0;      // n*mE     A_E'(A_E A_E')^-1
    double* Gt;      // mI*n     A_I P
    double* K;       // mI*mE    A_I N
    double* Af;      // m*n      A_full (dense copy; rows of box constraints are unit rows)
    double* scr;     // preparation scratch (pmats_scratch_doubles)
    // what the solver reads: every operator stored so that consecutive threads read consecutive addresses
    // (out[r] = sum_c X[c*ld + r] v[c], ld even), see op_mv_t<true> in lcqp_device.cuh
    double* Tt;      // mI*ldI   Gt A_I'   (exactly symmetric)
    double* Pp;      // n*ldn    P (symmetric)
    double* GtT;     // n*ldI    Gt'
    double* KT;      // mE*ldI   K'
    double* NT;      // mE*ldn   N'
    double* N0p;     // n*ldE    N0
    double* Afp;     // m*ldn    A_full
    double* AfT;     // n*ldm    A_full'
    int ldn, ldm, ldE;
    int* Eidx;       // mE  full row index of eliminated row e
    int* Iidx;       // mI  full row index of remaining row i
    int* pos;        // m   position of full row r in its list (E or I)
    signed char* isE;  // m
    int mE, mI, ldI;
    int status;      // 0 ok, 1 reduced Hessian not positive definite (take the regularised solver), 2 other failure
    Op oP, oA, oAt, oGt, oK, oN, oN0t;   // n x n, m x n, n x m, mI x n, mI x mE, n x mE, mE x n
};

struct PDims {
    int n, nC, nComp, mA, m, has_box;
};

inline LCQ_HD PDims make_pdims(int nV, int nC, int nComp, int has_box)
{
    PDims d;
    d.n = nV; d.nC = nC; d.nComp = nComp; d.mA = nC + 2 * nComp; d.has_box = has_box;
    d.m = d.mA + (has_box ? nV : 0);
    return d;
}

inline LCQ_HD size_t pev(size_t k) { return (k + 1) & ~(size_t)1; }
inline LCQ_HD int pas_mEmax(const PDims& d) { return d.m < d.n ? d.m : d.n; }
inline LCQ_HD size_t pmats_scratch_doubles(const PDims& d)
{
    const size_t n = d.n, e = pas_mEmax(d);
    return 5 * pev(n * n) + pev(e * e) + pev(n * e) + 5 * pev(n) + 8;
}
inline LCQ_HD size_t pmats_doubles(const PDims& d)
{
    const size_t n = d.n, m = d.m, e = pas_mEmax(d);
    const size_t work = pev(n * n) + 2 * pev(n * e) + 2 * pev(m * n) + pev(m * e);
    const size_t fin = pev(m * (m + 1)) + pev(n * (n + 1)) + pev(n * (m + 1)) + pev(e * (m + 1)) + pev(e * (n + 1)) + pev(n * (e + 1)) + pev(m * (n + 1)) + pev(n * (m + 1));
    return work + fin + 2 * pev((m + 1) / 2) + pev(m) + pev((m + 7) / 8) + pmats_scratch_doubles(d) + 8;
}

LCQ_DEV void carve_pmats(PMats& mt, double* base, const PDims& d)
{
    const size_t n = d.n, m = d.m, e = pas_mEmax(d);
    mt.P = base; base += pev(n * n);
    mt.N = base; base += pev(n * e);
    mt.N0 = base; base += pev(n * e);
    mt.Gt = base; base += pev(m * n);
    mt.Af = base; base += pev(m * n);
    mt.K = base; base += pev(m * e);
    mt.Tt = base; base += pev(m * (m + 1));
    mt.Pp = base; base += pev(n * (n + 1));
    mt.GtT = base; base += pev(n * (m + 1));
    mt.KT = base; base += pev(e * (m + 1));
    mt.NT = base; base += pev(e * (n + 1));
    mt.N0p = base; base += pev(n * (e + 1));
    mt.Afp = base; base += pev(m * (n + 1));
    mt.AfT = base; base += pev(n * (m + 1));
    mt.ldn = (int)pev(n); mt.ldm = (int)pev(m); mt.ldE = 2;
    mt.Eidx = reinterpret_cast<int*>(base); base += pev((m + 1) / 2);
    mt.Iidx = reinterpret_cast<int*>(base); base += pev((m + 1) / 2);
    mt.pos = reinterpret_cast<int*>(base); base += pev(m);   // (m ints fit in m doubles)
    mt.isE = reinterpret_cast<signed char*>(base); base += pev((m + 7) / 8);
    mt.scr = base;
    mt.mE = 0; mt.mI = (int)m; mt.ldI = (int)pev(m); mt.status = 0;
}

// A_full[r][j] from the instance arrays
LCQ_DEV double afull_at(const PDims& d, const Inst& in, int r, int j)
{
    if (r < d.nC) return in.A[(size_t)r * d.n + j];
    if (r < d.nC + d.nComp) return in.L[(size_t)(r - d.nC) * d.n + j];
    if (r < d.mA) return in.R[(size_t)(r - d.nC - d.nComp) * d.n + j];
    return (r - d.mA == j) ? 1.0 : 0.0;
}

// bounds of full row r (LCQProblem.cpp:584-608, :745-782)
LCQ_DEV void row_bounds(const PDims& d, const Inst& in, int r, double& lo, double& up)
{
    if (r < d.nC) { lo = in.lbA ? in.lbA[r] : -INFINITY; up = in.ubA ? in.ubA[r] : INFINITY; }
    else if (r < d.nC + d.nComp) { const int k = r - d.nC; lo = in.lbL ? in.lbL[k] : 0.0; up = in.ubL ? in.ubL[k] : INFINITY; }
    else if (r < d.mA) { const int k = r - d.nC - d.nComp; lo = in.lbR ? in.lbR[k] : 0.0; up = in.ubR ? in.ubR[k] : INFINITY; }
    else { lo = in.lb ? in.lb[r - d.mA] : -INFINITY; up = in.ub ? in.ub[r - d.mA] : INFINITY; }
}

// In-place inversion of an SPD matrix by Gauss-Jordan; fails (returns 1) when a pivot is not above tol * max diagonal.
LCQ_DEVN int spd_invert_checked(double* M, int n, int ld, double tol, double* colbuf, double* rowbuf, Scalars* sc)
{
    double mx = 0;
    LCQ_LOOP for (int j = LCQ_TID; j < n; j += LCQ_NT) mx = fmax(mx, fabs(M[(size_t)j * ld + j]));
    mx = block_max(mx, sc);
    LCQ_LOOP for (int k = 0; k < n; k++) {
        LCQ_SYNC();
        const double p = M[(size_t)k * ld + k];
        if (!(p > tol * mx)) return 1;
        LCQ_LOOP for (int j = LCQ_TID; j < n; j += LCQ_NT) {
            colbuf[j] = M[(size_t)j * ld + k];
            rowbuf[j] = (j == k ? 1.0 : M[(size_t)k * ld + j]) / p;
        }
        LCQ_SYNC();
        LCQ_LOOP for (int i = LCQ_WARP; i < n; i += LCQ_NWARP) {
            double* row = M + (size_t)i * ld;
            if (i == k) {
                LCQ_LOOP for (int j = LCQ_LANE; j < n; j += LCQ_LANES) row[j] = rowbuf[j];
            } else {
                const double ci = colbuf[i];
                if (ci == 0.0) continue;
                LCQ_LOOP for (int j = LCQ_LANE; j < n; j += LCQ_LANES) row[j] = (j == k ? 0.0 : row[j]) - ci * rowbuf[j];
            }
        }
    }
    LCQ_SYNC();
    return 0;
}

// C (ra x cb, ldc) = A (ra x ca, lda) * B (ca x cb, ldb), skipping zeros of A; one warp per row of C
LCQ_DEVN void mm_rows(const double* A, int lda, const double* B, int ldb, double* C, int ldc, int ra, int ca, int cb)
{
    LCQ_LOOP for (int r = LCQ_WARP; r < ra; r += LCQ_NWARP) {
        double* c = C + (size_t)r * ldc;
        const double* a = A + (size_t)r * lda;
        LCQ_LOOP for (int j = LCQ_LANE; j < cb; j += LCQ_LANES) c[j] = 0.0;
        LCQ_LOOP for (int k = 0; k < ca; k++) {
            const double ak = a[k];
            if (ak == 0.0) continue;
            const double* b = B + (size_t)k * ldb;
            LCQ_LOOP for (int j = LCQ_LANE; j < cb; j += LCQ_LANES) c[j] += ak * b[j];
        }
    }
    LCQ_SYNC();
}

// C (ra x rb, ldc) = A (ra x k, lda) * B' (B is rb x k, ldb), skipping zeros of B; one thread per entry
LCQ_DEVN void mm_abt(const double* A, int lda, const double* B, int ldb, double* C, int ldc, int ra, int rb, int k)
{
    LCQ_LOOP for (int e = LCQ_TID; e < ra * rb; e += LCQ_NT) {
        const int r = e / rb, s = e - r * rb;
        const double* a = A + (size_t)r * lda;
        const double* b = B + (size_t)s * ldb;
        double acc = 0;
        LCQ_LOOP for (int j = 0; j < k; j++) { const double bj = b[j]; if (bj != 0.0) acc += a[j] * bj; }
        C[(size_t)r * ldc + s] = acc;
    }
    LCQ_SYNC();
}

// Prepare P, N, N0, Gt, Tt, K.  eqmask[r] != 0: row r has l = u (finite) in EVERY instance that will use these
// operands (rows found linearly dependent on the equality rows before them stay ordinary rows).  Block-cooperative.
LCQ_DEVN void pas_prepare(const PDims& d, const Inst& in, PMats& mt, const signed char* eqmask, Scalars* sc)
{
    const int n = d.n, m = d.m, emax = pas_mEmax(d);
    const size_t nn = pev((size_t)n * n);
    double* Hs = mt.scr;             // n*n  symmetrised Q
    double* Qh = Hs + nn;            // n*n  Householder product, Z = its last n - mE columns
    double* HZ = Qh + nn;            // n*nz
    double* Mi = HZ + nn;            // nz*nz  (Z'QZ)^-1
    double* W1 = Mi + nn;            // n*nz   Z Mi ; later Q N0 (n*mE)
    double* S2 = W1 + nn;            // mE*mE  (A_E A_E')^-1
    double* V = S2 + pev((size_t)emax * emax);   // n*emax reflectors (column k in V[:,k], leading dimension emax)
    double* beta = V + pev((size_t)n * emax);    // emax <= n
    double* va = beta + pev(n);      // n
    double* vb = va + pev(n);        // n
    double* vc = vb + pev(n);        // n
    double* ds = vc + pev(n);        // n  variable scaling 1/sqrt(Q_jj)
    // Everything up to P is computed in the scaled variables x = D xs, D = diag(1/sqrt(Q_jj)): a Hessian whose
    // curvatures differ by many orders of magnitude (the reference's examples regularise with 5e-12) is badly scaled,
    // not ill-conditioned -- in the scaled variables Z'QZ is inverted to full relative accuracy in every block.
    // (curvatures below kPDTol of the largest count as zero: such a Hessian is semidefinite for this solver -- qpOASES
    //  treats them by its zero-curvature exchanges, QProblem.cpp:4278-4511, which the regularised solver stands in for)
    int badd = 0;
    double qmax = 0;
    LCQ_LOOP for (int j = LCQ_TID; j < n; j += LCQ_NT) qmax = fmax(qmax, in.Q[(size_t)j * n + j]);
    qmax = block_max(qmax, sc);
    LCQ_LOOP for (int j = LCQ_TID; j < n; j += LCQ_NT) { const double q = in.Q[(size_t)j * n + j]; if (!(q > kPDTol * qmax)) badd = 1; ds[j] = (q > 0.0) ? 1.0 / sqrt(q) : 1.0; }
    if (block_or(badd, sc)) { if (LCQ_TID == 0) mt.status = 1; LCQ_SYNC(); return; }
    LCQ_LOOP for (int e = LCQ_TID; e < n * n; e += LCQ_NT) { const int i = e / n, j = e - i * n; Hs[e] = 0.5 * (in.Q[e] + in.Q[(size_t)j * n + i]) * ds[i] * ds[j]; }
    LCQ_LOOP for (int e = LCQ_TID; e < m * n; e += LCQ_NT) { const int r = e / n; mt.Af[e] = afull_at(d, in, r, e - r * n); }
    LCQ_LOOP for (int r = LCQ_TID; r < m; r += LCQ_NT) mt.isE[r] = 0;
    LCQ_SYNC();
    // ---- Householder QR of A_E', one equality row at a time
    int mE = 0;
    LCQ_LOOP for (int r = 0; r < m && eqmask; r++) {
        if (!eqmask[r] || mE >= emax) continue;
        LCQ_LOOP for (int j = LCQ_TID; j < n; j += LCQ_NT) va[j] = mt.Af[(size_t)r * n + j] * ds[j];
        LCQ_SYNC();
        double part = 0;
        LCQ_LOOP for (int j = LCQ_TID; j < n; j += LCQ_NT) part += va[j] * va[j];
        const double nrm2 = block_sum(part, sc);
        LCQ_LOOP for (int k = 0; k < mE; k++) {
            part = 0;
            LCQ_LOOP for (int j = k + LCQ_TID; j < n; j += LCQ_NT) part += V[(size_t)j * emax + k] * va[j];
            const double t = beta[k] * block_sum(part, sc);
            LCQ_LOOP for (int j = k + LCQ_TID; j < n; j += LCQ_NT) va[j] -= t * V[(size_t)j * emax + k];
            LCQ_SYNC();
        }
        part = 0;
        LCQ_LOOP for (int j = mE + LCQ_TID; j < n; j += LCQ_NT) part += va[j] * va[j];
        const double rest2 = block_sum(part, sc);
        if (!(rest2 > 1e-20 * nrm2) || !(nrm2 > 0.0)) continue;   // dependent on the equality rows before it
        const double a0 = va[mE];
        const double alpha = (a0 >= 0.0) ? -sqrt(rest2) : sqrt(rest2);
        LCQ_SYNC();
        LCQ_LOOP for (int j = LCQ_TID; j < n; j += LCQ_NT) V[(size_t)j * emax + mE] = (j < mE) ? 0.0 : (j == mE ? a0 - alpha : va[j]);
        if (LCQ_TID == 0) { beta[mE] = 2.0 / (rest2 - a0 * a0 + (a0 - alpha) * (a0 - alpha)); mt.Eidx[mE] = r; mt.isE[r] = 1; }
        mE++;
        LCQ_SYNC();
    }
    LCQ_SYNC();
    if (LCQ_TID == 0) {
        int mI = 0, e = 0;
        for (int r = 0; r < m; r++) {
            if (mt.isE[r]) mt.pos[r] = e++;
            else { mt.Iidx[mI] = r; mt.pos[r] = mI++; }
        }
        mt.mE = mE; mt.mI = mI; mt.ldI = (int)pev(mI);
    }
    LCQ_SYNC();
    const int mI = mt.mI, ldI = mt.ldI, nz = n - mE;
    if (mE == 0) {
        // P = Q^-1 = D (D Q D)^-1 D
        LCQ_LOOP for (int e = LCQ_TID; e < n * n; e += LCQ_NT) mt.P[e] = Hs[e];
        LCQ_SYNC();
        if (spd_invert_checked(mt.P, n, n, kPDTol, va, vb, sc)) { if (LCQ_TID == 0) mt.status = 1; LCQ_SYNC(); return; }
        LCQ_LOOP for (int e = LCQ_TID; e < n * n; e += LCQ_NT) { const int i = e / n, j = e - i * n; mt.P[e] *= ds[i] * ds[j]; }
        LCQ_SYNC();
    } else {
        // Qh = H_0 H_1 ... H_(mE-1) applied to the identity (column j of Qh = Q e_j)
        LCQ_LOOP for (int e = LCQ_TID; e < n * n; e += LCQ_NT) { const int i = e / n, j = e - i * n; Qh[e] = (i == j) ? 1.0 : 0.0; }
        LCQ_SYNC();
        LCQ_LOOP for (int k = mE - 1; k >= 0; k--) {
            // t_j = beta_k v_k' Qh[:, j] ; Qh[:, j] -= t_j v_k        (one thread per column j)
            LCQ_LOOP for (int j = LCQ_TID; j < n; j += LCQ_NT) {
                double acc = 0;
                LCQ_LOOP for (int i = k; i < n; i++) acc += V[(size_t)i * emax + k] * Qh[(size_t)i * n + j];
                vc[j] = beta[k] * acc;
            }
            LCQ_SYNC();
            LCQ_LOOP for (int e = LCQ_TID; e < (n - k) * n; e += LCQ_NT) { const int i = k + e / n, j = e - (i - k) * n; Qh[(size_t)i * n + j] -= V[(size_t)i * emax + k] * vc[j]; }
            LCQ_SYNC();
        }
        // Z = Qh[:, mE:]  (n x nz, leading dimension n, column offset mE)
        if (nz > 0) {
            const double* Z = Qh + mE;
            mm_rows(Hs, n, Z, n, HZ, nz, n, n, nz);                    // HZ = Q Z       (n x nz, ld nz)
            // Mi = Z' HZ  (nz x nz)
            LCQ_LOOP for (int e = LCQ_TID; e < nz * nz; e += LCQ_NT) {
                const int a = e / nz, b = e - a * nz;
                if (b < a) continue;
                double acc = 0;
                LCQ_LOOP for (int i = 0; i < n; i++) acc += Z[(size_t)i * n + a] * HZ[(size_t)i * nz + b];
                Mi[(size_t)a * nz + b] = acc;
            }
            LCQ_SYNC();
            LCQ_LOOP for (int e = LCQ_TID; e < nz * nz; e += LCQ_NT) { const int a = e / nz, b = e - a * nz; if (b < a) Mi[e] = Mi[(size_t)b * nz + a]; }
            LCQ_SYNC();
            if (spd_invert_checked(Mi, nz, nz, kPDTol, va, vb, sc)) { if (LCQ_TID == 0) mt.status = 1; LCQ_SYNC(); return; }
            mm_rows(Z, n, Mi, nz, W1, nz, n, nz, nz);                  // W1 = Z Mi      (n x nz)
            mm_abt(W1, nz, Z, n, mt.P, n, n, n, nz);                   // P = W1 Z'
            // exact symmetry, back to the unscaled variables
            LCQ_LOOP for (int e = LCQ_TID; e < n * n; e += LCQ_NT) { const int i = e / n, j = e - i * n; if (j >= i) { const double v = 0.5 * (mt.P[e] + mt.P[(size_t)j * n + i]) * ds[i] * ds[j]; mt.P[e] = v; mt.P[(size_t)j * n + i] = v; } }
            LCQ_SYNC();
        } else {
            LCQ_LOOP for (int e = LCQ_TID; e < n * n; e += LCQ_NT) mt.P[e] = 0.0;
            LCQ_SYNC();
        }
        // S2 = (A_E A_E')^-1 ; N0 = A_E' S2 ; N = N0 - P (Q N0)
        LCQ_LOOP for (int e = LCQ_TID; e < mE * mE; e += LCQ_NT) {
            const int a = e / mE, b = e - a * mE;
            const double* ra = mt.Af + (size_t)mt.Eidx[a] * n;
            const double* rb = mt.Af + (size_t)mt.Eidx[b] * n;
            double acc = 0;
            LCQ_LOOP for (int j = 0; j < n; j++) { const double x = ra[j]; if (x != 0.0) acc += x * rb[j]; }
            S2[e] = acc;
        }
        LCQ_SYNC();
        if (spd_invert_checked(S2, mE, mE, 1e-13, va, vb, sc)) { if (LCQ_TID == 0) mt.status = 2; LCQ_SYNC(); return; }
        LCQ_LOOP for (int e = LCQ_TID; e < n * mE; e += LCQ_NT) {
            const int j = e / mE, a = e - j * mE;
            double acc = 0;
            LCQ_LOOP for (int b = 0; b < mE; b++) { const double x = mt.Af[(size_t)mt.Eidx[b] * n + j]; if (x != 0.0) acc += x * S2[(size_t)b * mE + a]; }
            mt.N0[e] = acc;
        }
        LCQ_SYNC();
        // W1 = Q N0 (n x mE) with the unscaled Q = D^-1 Hs D^-1: scale the rows of N0, multiply, scale the rows back
        LCQ_LOOP for (int e = LCQ_TID; e < n * mE; e += LCQ_NT) HZ[e] = mt.N0[e] / ds[e / mE];
        LCQ_SYNC();
        mm_rows(Hs, n, HZ, mE, W1, mE, n, n, mE);
        LCQ_LOOP for (int e = LCQ_TID; e < n * mE; e += LCQ_NT) W1[e] /= ds[e / mE];
        LCQ_SYNC();
        mm_rows(mt.P, n, W1, mE, mt.N, mE, n, n, mE);                  // N = P W1
        LCQ_LOOP for (int e = LCQ_TID; e < n * mE; e += LCQ_NT) mt.N[e] = mt.N0[e] - mt.N[e];
        LCQ_SYNC();
    }
    // Gt = A_I P ; Tt = Gt A_I' ; K = A_I N
    LCQ_LOOP for (int i = LCQ_WARP; i < mI; i += LCQ_NWARP) {
        double* g = mt.Gt + (size_t)i * n;
        const double* a = mt.Af + (size_t)mt.Iidx[i] * n;
        LCQ_LOOP for (int j = LCQ_LANE; j < n; j += LCQ_LANES) g[j] = 0.0;
        LCQ_LOOP for (int k = 0; k < n; k++) {
            const double ak = a[k];
            if (ak == 0.0) continue;
            const double* h = mt.P + (size_t)k * n;
            LCQ_LOOP for (int j = LCQ_LANE; j < n; j += LCQ_LANES) g[j] += ak * h[j];
        }
    }
    LCQ_SYNC();
    LCQ_LOOP for (int e = LCQ_TID; e < mI * mI; e += LCQ_NT) {
        const int i = e / mI, j = e - i * mI;
        if (j < i) continue;
        const double* g = mt.Gt + (size_t)i * n;
        const double* a = mt.Af + (size_t)mt.Iidx[j] * n;
        double acc = 0;
        LCQ_LOOP for (int k = 0; k < n; k++) { const double ak = a[k]; if (ak != 0.0) acc += g[k] * ak; }
        mt.Tt[(size_t)i * ldI + j] = acc;
    }
    LCQ_SYNC();
    LCQ_LOOP for (int e = LCQ_TID; e < mI * mI; e += LCQ_NT) { const int i = e / mI, j = e - i * mI; if (j < i) mt.Tt[(size_t)i * ldI + j] = mt.Tt[(size_t)j * ldI + i]; }
    LCQ_LOOP for (int e = LCQ_TID; e < mI * mE; e += LCQ_NT) {
        const int i = e / mE, a = e - i * mE;
        const double* r = mt.Af + (size_t)mt.Iidx[i] * n;
        double acc = 0;
        LCQ_LOOP for (int j = 0; j < n; j++) { const double x = r[j]; if (x != 0.0) acc += x * mt.N[(size_t)j * mE + a]; }
        mt.K[e] = acc;
    }
    LCQ_SYNC();
    // the solver's copies
    const int ldn = mt.ldn, ldm = mt.ldm;
    const int ldE = (int)pev(mE > 0 ? mE : 1);
    if (LCQ_TID == 0) mt.ldE = ldE;
    LCQ_LOOP for (int e = LCQ_TID; e < n * n; e += LCQ_NT) { const int i = e / n, j = e - i * n; mt.Pp[(size_t)i * ldn + j] = mt.P[e]; }
    LCQ_LOOP for (int e = LCQ_TID; e < mI * n; e += LCQ_NT) { const int i = e / n, j = e - i * n; mt.GtT[(size_t)j * ldI + i] = mt.Gt[e]; }
    LCQ_LOOP for (int e = LCQ_TID; e < mI * mE; e += LCQ_NT) { const int i = e / mE, a = e - i * mE; mt.KT[(size_t)a * ldI + i] = mt.K[e]; }
    LCQ_LOOP for (int e = LCQ_TID; e < n * mE; e += LCQ_NT) { const int j = e / mE, a = e - j * mE; mt.NT[(size_t)a * ldn + j] = mt.N[e]; mt.N0p[(size_t)j * ldE + a] = mt.N0[e]; }
    LCQ_LOOP for (int e = LCQ_TID; e < m * n; e += LCQ_NT) { const int r = e / n, j = e - r * n; mt.Afp[(size_t)r * ldn + j] = mt.Af[e]; mt.AfT[(size_t)j * ldm + r] = mt.Af[e]; }
    LCQ_SYNC();
}

// dense operator descriptors (one thread): all in the transposed ("column") form
LCQ_DEV void pmats_dense_ops(const PDims& d, PMats& mt)
{
    const int e1 = mt.mE > 0 ? mt.mE : 1;
    mt.oP = dense_op(mt.Pp, d.n, d.n, mt.ldn, 1);
    mt.oA = dense_op(mt.AfT, d.m, d.n, mt.ldm, 1);
    mt.oAt = dense_op(mt.Afp, d.n, d.m, mt.ldn, 1);
    mt.oGt = dense_op(mt.GtT, mt.mI, d.n, mt.ldI, 1);
    mt.oK = dense_op(mt.KT, mt.mI, e1, mt.ldI, 1);
    mt.oN = dense_op(mt.NT, d.n, e1, mt.ldn, 1);
    mt.oN0t = dense_op(mt.N0p, e1, d.n, mt.ldE, 1);
}

// CSR copies where sparse (batch-shared operands only)
LCQ_DEVN void pmats_build_ops(const PDims& d, PMats& mt, CsrPool& pool, Scalars* sc)
{
    const Op a = build_op(mt.Pp, d.n, d.n, mt.ldn, 1, pool, sc);
    const Op b = build_op(mt.AfT, d.m, d.n, mt.ldm, 1, pool, sc);
    const Op c = build_op(mt.Afp, d.n, d.m, mt.ldn, 1, pool, sc);
    const Op e = build_op(mt.GtT, mt.mI, d.n, mt.ldI, 1, pool, sc);
    if (LCQ_TID == 0) { mt.oP = a; mt.oA = b; mt.oAt = c; mt.oGt = e; }
    LCQ_SYNC();
}

// ------------------------------------------------------------------------------------------------
// Per-instance state
// ------------------------------------------------------------------------------------------------
struct PWork {
    // shared memory: vectors over the remaining rows (mI) and the working set (cap)
    double *y, *z, *l, *u, *c, *dz, *dc, *lf, *uf, *yb;    // mI
    double *va, *vb;                                       // cap
    double *xq, *tn, *sv;                                  // n   (xq: the QP's primal iterate; tn, sv: operator inputs)
    double *tm1, *tm2;                                     // m   (operator inputs / outputs in full row order)
    double *tE1, *tE2, *bEn;                               // mE  (bEn: the true bounds of the eliminated rows)
    int* widx;                                             // cap: rows (I numbering) of the working set
    int* stamp;                                            // mI: position in its index list (insertion order)
    signed char* st;                                       // mI: ST_*
    Scalars* sc;
    // global scratch
    double* Sinv; int ld;                                  // cap x ld, full storage, exactly symmetric
    double *gq, *gk, *dg0, *xk, *pk, *gt, *gphi, *stat, *tq;   // n
    double *Lx, *Rx;                                       // nComp
    double *ys;                                            // m   duals of the last QP, full row order
    double *bE, *bEb, *yE;                                 // mE  (bE: current equality bounds, bEb: at the last rebase, yE: duals)
};

struct PQP {
    const PDims* d;
    const PMats* mt;
    const lcqp_cuda_options* o;
    const Inst* in;
    PWork* w;
    int nw, cap;
    int e_moving;        // the bounds of the eliminated rows are still on their way (initial homotopy only)
    int stamp_next, rampOffset;
    double phi, len0;
    int nwsr;            // working-set changes of the current QP
    long long n_solve;   // explicit-inverse solves
    long long n_change;
    long long n_polish;
    // work done on the per-instance inverse and on Tt (the two streams that dominate a homotopy step): fp64
    // multiply-adds and the bytes they read / write (L2-resident scratch and operands), for the bench's model
    long long n_mac, n_byte;
};
// (thread 0 of the group keeps the books -- only in the counting build, -DLCQP_COUNT_WORK: liblcqp_cuda_work.so; the
//  bookkeeping costs the shipped kernel 3-5 % on the B200, so the product is built without it)
#ifdef LCQP_COUNT_WORK
LCQ_DEV void pas_count(PQP& s, long long mac, long long bytes)
{
    if (LCQ_TID == 0) { s.n_mac += mac; s.n_byte += bytes; }
}
#else
LCQ_DEV void pas_count(PQP&, long long, long long) {}
#endif

inline LCQ_HD int pas_cap(const PDims& d, int mE, int mI)
{
    int c = d.n - mE;   // a linearly independent working set has at most n rows, mE of them are the eliminated ones
    if (c > mI) c = mI;
    if (c < 1) c = 1;
    return c;
}
inline LCQ_HD int pas_ld(int cap) { return (cap + 1) & ~1; }
inline LCQ_HD size_t pas_smem_doubles(const PDims& d, int mE, int mI, int cap)
{
    return 10 * pev(mI) + 2 * pev(cap) + 3 * pev(d.n) + 2 * pev(d.m) + 3 * pev(mE > 0 ? mE : 1);
}
inline LCQ_HD size_t pas_smem_bytes(const PDims& d, int mE, int mI, int cap)
{
    return pas_smem_doubles(d, mE, mI, cap) * sizeof(double) + (size_t)pev(cap) * sizeof(int) + (size_t)pev(mI) * sizeof(int) + ((size_t)(mI + 15) / 16) * 16 + sizeof(Scalars) + 64;
}
inline LCQ_HD size_t pas_gl_doubles(const PDims& d, int mE, int cap)
{
    return (size_t)cap * pas_ld(cap) + 9 * pev(d.n) + 2 * pev(d.nComp) + pev(d.m) + 3 * pev(mE > 0 ? mE : 1);
}

LCQ_DEV void pas_carve(PWork& w, const PDims& d, int mE, int mI, int cap, unsigned char* smem, double* gl)
{
    double* q = reinterpret_cast<double*>(smem);
    auto take = [&](size_t k) { double* r = q; q += pev(k); return r; };
    w.y = take(mI); w.z = take(mI); w.l = take(mI); w.u = take(mI); w.c = take(mI); w.dz = take(mI); w.dc = take(mI);
    w.lf = take(mI); w.uf = take(mI); w.yb = take(mI);
    w.va = take(cap); w.vb = take(cap);
    w.xq = take(d.n); w.tn = take(d.n); w.sv = take(d.n);
    w.tm1 = take(d.m); w.tm2 = take(d.m);
    {
        const int e1 = mE > 0 ? mE : 1;
        w.tE1 = take(e1); w.tE2 = take(e1); w.bEn = take(e1);
    }
    w.widx = reinterpret_cast<int*>(q);
    w.stamp = w.widx + pev(cap);
    w.st = reinterpret_cast<signed char*>(w.stamp + pev(mI));
    uintptr_t sp = reinterpret_cast<uintptr_t>(w.st + ((mI + 15) / 16) * 16);
    sp = (sp + 15) & ~(uintptr_t)15;
    w.sc = reinterpret_cast<Scalars*>(sp);
    auto tg = [&](size_t k) { double* r = gl; gl += pev(k); return r; };
    w.ld = pas_ld(cap);
    w.Sinv = tg((size_t)cap * w.ld);
    w.gq = tg(d.n); w.gk = tg(d.n); w.dg0 = tg(d.n); w.xk = tg(d.n); w.pk = tg(d.n); w.gt = tg(d.n);
    w.gphi = tg(d.n); w.stat = tg(d.n); w.tq = tg(d.n);
    w.Lx = tg(d.nComp); w.Rx = tg(d.nComp);
    w.ys = tg(d.m);
    const int e = mE > 0 ? mE : 1;
    w.bE = tg(e); w.bEb = tg(e); w.yE = tg(e);
}

// ---- small block-wide helpers ----------------------------------------------------------------------
// two block-wide maxima with one pair of barriers
LCQ_DEV void block_max2(double& a, double& b, Scalars* sc)
{
    a = warp_max(a);
    b = warp_max(b);
    LCQ_SYNC();
    if (LCQ_LANE == 0) { sc->red[LCQ_WARP] = a; sc->red[32 + LCQ_WARP] = b; }
    LCQ_SYNC();
    double x = sc->red[0], y = sc->red[32];
    LCQ_LOOP for (int k = 1; k < LCQ_NWARP; k++) { x = fmax(x, sc->red[k]); y = fmax(y, sc->red[32 + k]); }
    a = x; b = y;
}

// lexicographic argmin over (value, rank): smallest value, ties -> smallest rank; idx < 0: no candidate
LCQ_DEV bool lex2_better(double a, int ra, int ia, double b, int rb, int ib)
{
    if (ib < 0) return false;
    if (ia < 0) return true;
    if (b != a) return b < a;
    return rb < ra;
}
LCQ_DEV int block_argmin2(double a, int rank, int idx, double* aout, int* rout, Scalars* sc)
{
#ifndef LCQP_HOST_EMU
    for (int o = 16; o > 0; o >>= 1) {
        const double a2 = __shfl_xor_sync(0xffffffffu, a, o);
        const int r2 = __shfl_xor_sync(0xffffffffu, rank, o);
        const int i2 = __shfl_xor_sync(0xffffffffu, idx, o);
        if (lex2_better(a, rank, idx, a2, r2, i2)) { a = a2; rank = r2; idx = i2; }
    }
#endif
    LCQ_SYNC();
    if (LCQ_LANE == 0) { sc->red[LCQ_WARP] = a; sc->ired[LCQ_WARP] = idx; sc->ior[LCQ_WARP] = rank; }
    LCQ_SYNC();
    double ra = sc->red[0];
    int ri = sc->ired[0], rr = sc->ior[0];
    LCQ_LOOP for (int k = 1; k < LCQ_NWARP; k++)
        if (lex2_better(ra, rr, ri, sc->red[k], sc->ior[k], sc->ired[k])) { ra = sc->red[k]; rr = sc->ior[k]; ri = sc->ired[k]; }
    *aout = ra;
    if (rout) *rout = rr;
    return ri;
}

// is full row r (of remaining row i) a bound row?  local index for the ramp (QProblem.cpp:5425-5483)
LCQ_DEV bool row_is_bound(const PQP& s, int i, int* li)
{
    const int r = s.mt->Iidx[i];
    if (r >= s.d->mA) { *li = r - s.d->mA; return true; }
    *li = r;
    return false;
}

LCQ_DEV void ramp_vals(const PQP& s, int i, double* rP, double* rD)
{
    const int nV = s.d->n, nC = s.d->mA, nRamp = nV + nC + nC + nV, off = s.rampOffset;
    int li;
    const bool isb = row_is_bound(s, i, &li);
    const int kP = isb ? li : nV + li, kD = isb ? nV + nC + nC + li : nV + nC + li;
    const double tP = (double)((kP + off) % nRamp) / (double)(nRamp - 1);
    const double tD = (double)((kD + off) % nRamp) / (double)(nRamp - 1);
    *rP = (1.0 - tP) * kRamp0 + tP * kRamp1;
    *rD = (1.0 - tD) * kRamp0 + tD * kRamp1;
}

// far bounds of the remaining rows (QProblem.cpp:5498-5552 with enableRamping)
LCQ_DEVN void far_bounds(PQP& s, double far)
{
    const PWork& w = *s.w;
    const int nV = s.d->n, nC = s.d->mA, nRamp = nV + nC;
    LCQ_LOOP for (int i = LCQ_TID; i < s.mt->mI; i += LCQ_NT) {
        int li;
        const bool isb = row_is_bound(s, i, &li);
        const double t = (double)(((isb ? li : nV + li) + s.rampOffset) % nRamp) / (double)(nRamp - 1);
        const double rv = far * (1.0 + (1.0 - t) * kRamp0 + t * kRamp1);
        double lo, up;
        row_bounds(*s.d, *s.in, s.mt->Iidx[i], lo, up);
        w.lf[i] = fmax(-rv, lo);
        w.uf[i] = fmin(rv, up);
    }
    LCQ_SYNC();
}

// ---- the explicit inverse of Tt[W,W] ------------------------------------------------------------------
// va = Tt[k, W], vb = Sinv va; returns the bordering pivot Tt[k][k] - va'vb
LCQ_DEVN double pas_pivot(PQP& s, int k)
{
    const PWork& w = *s.w;
    const int nw = s.nw;
    const double* row = s.mt->Tt + (size_t)k * s.mt->ldI;
    LCQ_LOOP for (int a = LCQ_TID; a < nw; a += LCQ_NT) w.va[a] = row[w.widx[a]];
    LCQ_SYNC();
    if (nw > 0) { sym_apply(w.Sinv, w.ld, nw, w.va, 1.0, w.vb, nullptr, nullptr, nullptr); LCQ_SYNC(); }
    pas_count(s, (long long)nw * nw + nw, 8LL * nw * w.ld + 32LL * nw);
    double part = 0;
    LCQ_LOOP for (int a = LCQ_TID; a < nw; a += LCQ_NT) part += w.va[a] * w.vb[a];
    return row[k] - block_sum(part, w.sc);
}

// append row k (pivot p and vb from pas_pivot)
LCQ_DEVN void pas_append(PQP& s, int k, int status, double p)
{
    const PWork& w = *s.w;
    const int nw = s.nw, ld = w.ld;
    const double ip = 1.0 / p;
    if (nw > 0) rank1_update_full(w.Sinv, ld, nw, w.vb, ip);
    pas_count(s, (long long)nw * nw, 16LL * nw * ld + 16LL * nw);
    double* row = w.Sinv + (size_t)nw * ld;
    LCQ_LOOP for (int a = LCQ_TID; a < nw; a += LCQ_NT) { const double v = -w.vb[a] * ip; row[a] = v; w.Sinv[(size_t)a * ld + nw] = v; }
    LCQ_SYNC();   // every thread has read s.nw
    if (LCQ_TID == 0) {
        row[nw] = ip;
        w.widx[nw] = k;
        w.st[k] = (signed char)status;
        w.stamp[k] = s.stamp_next;
        s.stamp_next++;
        s.nw = nw + 1;
        s.n_change++;
    }
    LCQ_SYNC();
}

// remove the row at position p of the working set (the last position moves into p)
LCQ_DEVN void pas_remove(PQP& s, int p)
{
    const PWork& w = *s.w;
    const int nw = s.nw, last = nw - 1, ld = w.ld;
    double* Si = w.Sinv;
    LCQ_LOOP for (int b = LCQ_TID; b < nw; b += LCQ_NT) w.va[b] = Si[(size_t)b * ld + p];
    LCQ_SYNC();
    const double ic = 1.0 / w.va[p];
    rank1_update_full(Si, ld, nw, w.va, -ic);
    LCQ_SYNC();
    pas_count(s, (long long)nw * nw, 16LL * nw * ld + 56LL * nw);
    const int k = w.widx[p];
    if (p != last) {
        LCQ_LOOP for (int b = LCQ_TID; b < nw; b += LCQ_NT) w.vb[b] = Si[(size_t)last * ld + b];
        LCQ_SYNC();
        LCQ_LOOP for (int b = LCQ_TID; b < last; b += LCQ_NT) {
            const double v = (b == p) ? w.vb[last] : w.vb[b];
            Si[(size_t)p * ld + b] = v;
            Si[(size_t)b * ld + p] = v;
        }
    }
    LCQ_SYNC();
    if (LCQ_TID == 0) {
        if (p != last) w.widx[p] = w.widx[last];
        w.st[k] = ST_INACTIVE;
        w.y[k] = 0.0;
        w.stamp[k] = s.stamp_next;
        s.stamp_next++;
        s.nw = last;
        s.n_change++;
    }
    LCQ_SYNC();
}

// out (mI) = Tt[:, W] v - sub   (v over the working set; reads rows W of the symmetric Tt).  v, out, sub in shared
// memory.  A pair of adjacent lanes owns two adjacent columns (one 16-byte load per row) and splits the rows of the
// working set in two halves; twelve loads are in flight per thread; one shuffle joins the halves.
LCQ_DEVN void tt_cols_apply(const PQP& s, const double* v, double* out, const double* sub, bool skip_active = false)
{
    const PWork& w = *s.w;
    const int mI = s.mt->mI, ldI = s.mt->ldI, nw = s.nw;
    const double* Tt = s.mt->Tt;
#ifndef LCQP_HOST_EMU
    const int npair = (mI + 1) >> 1;
    const int half = (nw + 1) >> 1;
    LCQ_LOOP for (int t0 = 0; t0 < 2 * npair; t0 += LCQ_NT) {
        const int t = t0 + LCQ_TID;
        bool act = t < 2 * npair;
        const int c = t >> 1, h = t & 1;
        // (both lanes of a pair take the same decision: the shuffle below stays convergent)
        if (act && skip_active && w.st[2 * c] != ST_INACTIVE && (2 * c + 1 >= mI || w.st[2 * c + 1] != ST_INACTIVE)) act = false;
        int b = act ? (h ? half : 0) : 0;
        const int b1 = act ? (h ? nw : half) : 0;
        double x0 = 0, x1 = 0, y0 = 0, y1 = 0;
        LCQ_LOOP for (; b + 12 <= b1; b += 12) {
            double m0[12], m1[12];
#pragma unroll
            for (int k = 0; k < 12; k++) ldg128v(Tt + ((unsigned)w.widx[b + k] * (unsigned)ldI + 2u * (unsigned)c), m0[k], m1[k]);
#pragma unroll
            for (int k = 0; k < 12; k += 2) {
                const double v0 = v[b + k], v1 = v[b + k + 1];
                x0 += m0[k] * v0; y0 += m1[k] * v0;
                x1 += m0[k + 1] * v1; y1 += m1[k + 1] * v1;
            }
        }
        LCQ_LOOP for (; b + 4 <= b1; b += 4) {
            double m0[4], m1[4];
#pragma unroll
            for (int k = 0; k < 4; k++) ldg128v(Tt + ((unsigned)w.widx[b + k] * (unsigned)ldI + 2u * (unsigned)c), m0[k], m1[k]);
#pragma unroll
            for (int k = 0; k < 4; k += 2) {
                const double v0 = v[b + k], v1 = v[b + k + 1];
                x0 += m0[k] * v0; y0 += m1[k] * v0;
                x1 += m0[k + 1] * v1; y1 += m1[k + 1] * v1;
            }
        }
        LCQ_LOOP for (; b < b1; b++) {
            double m0, m1;
            ldg128(Tt + ((unsigned)w.widx[b] * (unsigned)ldI + 2u * (unsigned)c), m0, m1);
            const double v0 = v[b];
            x0 += m0 * v0; y0 += m1 * v0;
        }
        double sa = x0 + x1, sb = y0 + y1;
        sa += __shfl_xor_sync(0xffffffffu, sa, 1);
        sb += __shfl_xor_sync(0xffffffffu, sb, 1);
        const int i = 2 * c + h;
        if (act && i < mI) out[i] = (h ? sb : sa) - (sub ? sub[i] : 0.0);
    }
#else
    LCQ_LOOP for (int i = LCQ_TID; i < mI; i += LCQ_NT) {
        // (same association as the device loop: two halves of the working set, each split over two accumulators)
        const int half = (nw + 1) >> 1;
        double hsum[2];
        for (int h = 0; h < 2; h++) {
            int b = h ? half : 0;
            const int b1 = h ? nw : half;
            double x0 = 0, x1 = 0;
            for (; b + 2 <= b1; b += 2) {
                x0 += Tt[(size_t)w.widx[b] * ldI + i] * v[b];
                x1 += Tt[(size_t)w.widx[b + 1] * ldI + i] * v[b + 1];
            }
            if (b < b1) x0 += Tt[(size_t)w.widx[b] * ldI + i] * v[b];
            hsum[h] = x0 + x1;
        }
        out[i] = (hsum[0] + hsum[1]) - (sub ? sub[i] : 0.0);
    }
#endif
    LCQ_SYNC();
}

// one application of a prepared operator: dense rows x cols, or the non-zeros of its CSR form (no L2 bytes when the
// CSR arrays sit in the CTA's shared-memory operator cache)
#ifdef LCQP_COUNT_WORK
LCQ_DEV void pas_count_op(PQP& s, const Op& op)
{
    if (LCQ_TID != 0) return;
    if (op.rp) { const long long nz = op.rp[op.rows]; s.n_mac += nz; s.n_byte += op.smem ? 0 : 10 * nz; }
    else { const long long e = (long long)op.rows * op.cols; s.n_mac += e; s.n_byte += 8 * e; }
}
#else
LCQ_DEV void pas_count_op(PQP&, const Op&) {}
#endif

// c-space image of a gradient change v (n) and an equality-bound change dbE (mE, may be null):  out = Gt v - K dbE
// (v, dbE, out in shared memory)
LCQ_DEVN void c_image(PQP& s, const double* v, const double* dbE, double* out)
{
    const PMats& mt = *s.mt;
    op_mv_s(mt.oGt, v, nullptr, 1.0, out);
    LCQ_SYNC();
    pas_count_op(s, mt.oGt);
    if (dbE && mt.mE > 0) { op_mv_s(mt.oK, dbE, out, -1.0, out); LCQ_SYNC(); pas_count_op(s, mt.oK); }
}

// xq += P (A_I' dyI - dgrad) + N dbE;  dyI: full-order vector in tm2 (zero on the eliminated rows), dgrad (n) and
// dbE (mE) may be null.   Scratch: tn, stat is NOT touched.
LCQ_DEVN void x_update(PQP& s, const double* dgrad, const double* dbE)
{
    const PWork& w = *s.w;
    const int n = s.d->n;
    op_mv_s(s.mt->oAt, w.tm2, nullptr, 1.0, w.tn);
    LCQ_SYNC();
    if (dgrad) { LCQ_LOOP for (int j = LCQ_TID; j < n; j += LCQ_NT) w.tn[j] -= dgrad[j]; LCQ_SYNC(); }
    op_mv_s(s.mt->oP, w.tn, w.xq, 1.0, w.xq);
    LCQ_SYNC();
    pas_count_op(s, s.mt->oAt);
    pas_count_op(s, s.mt->oP);
    if (dbE && s.mt->mE > 0) { op_mv_s(s.mt->oN, dbE, w.xq, 1.0, w.xq); LCQ_SYNC(); pas_count_op(s, s.mt->oN); }
}

// Bring xq, gq and the base duals / equality bounds to the current point of the homotopy: g = g_new - phi dg0.
LCQ_DEVN void rebase(PQP& s)
{
    const PWork& w = *s.w;
    const PMats& mt = *s.mt;
    const int n = s.d->n, m = s.d->m, mI = mt.mI, mE = mt.mE;
    const double f = 1.0 - s.phi;   // fraction of dg0 travelled since the last rebase
    LCQ_LOOP for (int j = LCQ_TID; j < n; j += LCQ_NT) w.stat[j] = f * w.dg0[j];
    LCQ_LOOP for (int r = LCQ_TID; r < m; r += LCQ_NT) { const int p = mt.pos[r]; w.tm2[r] = mt.isE[r] ? 0.0 : (w.y[p] - w.yb[p]); }
    LCQ_LOOP for (int e = LCQ_TID; e < mE; e += LCQ_NT) w.tE1[e] = w.bE[e] - w.bEb[e];
    LCQ_SYNC();
    x_update(s, w.stat, s.e_moving ? w.tE1 : nullptr);
    LCQ_LOOP for (int j = LCQ_TID; j < n; j += LCQ_NT) { w.gq[j] += w.stat[j]; w.dg0[j] *= s.phi; }
    LCQ_LOOP for (int e = LCQ_TID; e < mE; e += LCQ_NT) w.bEb[e] = w.bE[e];
    LCQ_LOOP for (int i = LCQ_TID; i < mI; i += LCQ_NT) w.yb[i] = w.y[i];
    LCQ_SYNC();
    if (LCQ_TID == 0) { s.len0 *= s.phi; s.phi = 1.0; }
    LCQ_SYNC();
}

// dg0 = gk - gq, phi = 1, len0 = max |dg0| / max(1, |gk|)  (QProblemB.cpp:2113-2127)
LCQ_DEVN void set_target(PQP& s)
{
    const PWork& w = *s.w;
    double mx = 0;
    LCQ_LOOP for (int j = LCQ_TID; j < s.d->n; j += LCQ_NT) {
        const double dg = w.gk[j] - w.gq[j];
        w.dg0[j] = dg;
        mx = fmax(mx, fabs(dg) / fmax(fabs(w.gk[j]), 1.0));
    }
    mx = block_max(mx, w.sc);
    if (LCQ_TID == 0) { s.len0 = mx; s.phi = 1.0; }
    LCQ_SYNC();
}

// c = Tt[:,W] y_W - z
LCQ_DEVN void c_from_state(PQP& s)
{
    const PWork& w = *s.w;
    LCQ_LOOP for (int a = LCQ_TID; a < s.nw; a += LCQ_NT) w.va[a] = w.y[w.widx[a]];
    LCQ_SYNC();
    tt_cols_apply(s, w.va, w.c, w.z);
    pas_count(s, (long long)s.nw * s.mt->mI, 8LL * s.nw * s.mt->mI);
}

// dc = c-image of the remaining gradient step (phi dg0) and of the remaining equality-bound step
LCQ_DEVN void dc_from_target(PQP& s)
{
    const PWork& w = *s.w;
    const PMats& mt = *s.mt;
    LCQ_LOOP for (int j = LCQ_TID; j < s.d->n; j += LCQ_NT) w.sv[j] = s.phi * w.dg0[j];
    LCQ_LOOP for (int e = LCQ_TID; e < mt.mE; e += LCQ_NT) w.tE2[e] = w.bEn[e] - w.bE[e];
    LCQ_SYNC();
    c_image(s, w.sv, s.e_moving ? w.tE2 : nullptr, w.dc);
}

// performRamping (QProblem.cpp:5416-5492)
LCQ_DEVN void ramping(PQP& s)
{
    const PWork& w = *s.w;
    const PMats& mt = *s.mt;
    const int n = s.d->n, m = s.d->m, mI = mt.mI;
    rebase(s);
    LCQ_LOOP for (int i = LCQ_TID; i < mI; i += LCQ_NT) {
        double rP, rD;
        ramp_vals(s, i, &rP, &rD);
        const double zi = w.z[i], sca = fmax(fabs(zi), 1.0);
        const int st = w.st[i];
        if (st != ST_LOWER) w.l[i] = zi - sca * rP;
        if (st != ST_UPPER) w.u[i] = zi + sca * rP;
        if (st == ST_LOWER) { w.l[i] = zi; w.y[i] = rD; }
        if (st == ST_UPPER) { w.u[i] = zi; w.y[i] = -rD; }
        if (st == ST_INACTIVE) w.y[i] = 0.0;
    }
    LCQ_SYNC();
    // gq = -Q xq + A_full' y   (setupAuxiliaryQPgradient, QProblem.cpp:2602-2641): the change of the gradient is
    // A_I'(y - yb) (x and the equality multipliers stay)
    LCQ_LOOP for (int r = LCQ_TID; r < m; r += LCQ_NT) { const int p = mt.pos[r]; w.tm2[r] = mt.isE[r] ? 0.0 : (w.y[p] - w.yb[p]); }
    LCQ_SYNC();
    op_mv_s(mt.oAt, w.tm2, w.gq, 1.0, w.gq);
    LCQ_SYNC();
    LCQ_LOOP for (int i = LCQ_TID; i < mI; i += LCQ_NT) w.yb[i] = w.y[i];
    LCQ_SYNC();
    (void)n;
    set_target(s);
    c_from_state(s);
    if (LCQ_TID == 0) s.rampOffset++;
    LCQ_SYNC();
}

// performDriftCorrection (QProblem.cpp:5559-5652) on the remaining rows; a clipped multiplier moves c (x stays)
LCQ_DEVN void drift(PQP& s)
{
    const PWork& w = *s.w;
    const int mI = s.mt->mI, ldI = s.mt->ldI;
    int any = 0;
    LCQ_LOOP for (int i = LCQ_TID; i < mI; i += LCQ_NT) {
        const int st = w.st[i];
        const double zi = w.z[i], yo = w.y[i];
        double yn;
        if (st == ST_LOWER) { w.l[i] = zi; w.u[i] = fmax(w.u[i], zi); yn = fmax(yo, 0.0); }
        else if (st == ST_UPPER) { w.u[i] = zi; w.l[i] = fmin(w.l[i], zi); yn = fmin(yo, 0.0); }
        else { w.l[i] = fmin(w.l[i], zi); w.u[i] = fmax(w.u[i], zi); yn = 0.0; }
        w.dz[i] = yn - yo;
        if (yn != yo) { any = 1; w.y[i] = yn; }
    }
    any = block_or(any, w.sc);
    if (any) {
        LCQ_LOOP for (int k = 0; k < mI; k++) {
            const double dk = w.dz[k];
            if (dk == 0.0) continue;
            const double* row = s.mt->Tt + (size_t)k * ldI;
            LCQ_LOOP for (int i = LCQ_TID; i < mI; i += LCQ_NT) w.c[i] += row[i] * dk;
        }
        LCQ_SYNC();
    }
}

// addConstraint with ensureLI (QProblem.cpp:2819-2975, :3117-3300).  Returns 0, or QP_INFEASIBLE.
LCQ_DEVN int add_with_li(PQP& s, int k, int status)
{
    const PWork& w = *s.w;
    double p = pas_pivot(s, k);
    const double tkk = s.mt->Tt[(size_t)k * s.mt->ldI + k];
    if (p > kLITol * tkk && s.nw < s.cap) { pas_append(s, k, status, p); return 0; }
    // linearly dependent: a_k (signed) = sum_a xi_a a_(W_a), xi = sgn vb; ratio test on the duals
    const double sgn = (status == ST_LOWER) ? 1.0 : -1.0;
    double best = INFINITY;
    int brank = 0, bpos = -1;
    LCQ_LOOP for (int a = LCQ_TID; a < s.nw; a += LCQ_NT) {
        const int i = w.widx[a];
        double num = w.y[i], den = sgn * w.vb[a];
        if (w.st[i] == ST_UPPER) { num = -num; den = -den; }
        if (den >= kEpsDen && num >= kEpsNum && num < kMaxDualJump * den) {
            int li;
            const int rank = (row_is_bound(s, i, &li) ? (1 << 30) : 0) + w.stamp[i];
            const double r = num / den;
            if (lex2_better(best, brank, bpos, r, rank, a)) { best = r; brank = rank; bpos = a; }
        }
    }
    double ymin;
    const int jmin = block_argmin2(best, brank, bpos, &ymin, nullptr, w.sc);
    if (jmin < 0) return QP_INFEASIBLE;
    LCQ_LOOP for (int a = LCQ_TID; a < s.nw; a += LCQ_NT) w.y[w.widx[a]] -= ymin * sgn * w.vb[a];
    LCQ_SYNC();
    pas_remove(s, jmin);
    if (LCQ_TID == 0) w.y[k] = (status == ST_LOWER) ? ymin : -ymin;
    LCQ_SYNC();
    p = pas_pivot(s, k);
    if (!(p > 0.0)) return QP_INFEASIBLE;
    pas_append(s, k, status, p);
    return 0;
}

// ---- the homotopy loop (QProblem::solveQP, QProblem.cpp:1477-1747) towards (gk, lf/uf, the true equality bounds) ----
LCQ_DEVN int pas_homotopy(PQP& s)
{
    const PWork& w = *s.w;
    const PMats& mt = *s.mt;
    const int mI = mt.mI, mE = mt.mE;
    const int max_iter = 20 * (s.d->n + s.d->m) + 1000;
    bool dc_dirty = true;
    double tau = 0.0;
    LCQ_LOOP for (int it = 0; it < max_iter; it++) {
        LCQ_PROF(w.sc, 8);
        if (dc_dirty) { dc_from_target(s); dc_dirty = false; }
        else { const double f = 1.0 - tau; LCQ_LOOP for (int i = LCQ_TID; i < mI; i += LCQ_NT) w.dc[i] *= f; LCQ_SYNC(); }
        const int nw = s.nw;
        LCQ_PROF(w.sc, 0);
        // step direction: dy_W = Sinv (db_W + dc_W), dz = Tt[:,W] dy_W - dc
        LCQ_LOOP for (int a = LCQ_TID; a < nw; a += LCQ_NT) {
            const int i = w.widx[a];
            const double db = (w.st[i] == ST_LOWER) ? (w.lf[i] - w.l[i]) : (w.uf[i] - w.u[i]);
            w.va[a] = db + w.dc[i];
        }
        LCQ_SYNC();
        if (nw > 0) {
            sym_apply(w.Sinv, w.ld, nw, w.va, 1.0, w.vb, nullptr, nullptr, nullptr);
            LCQ_SYNC();
            if (LCQ_TID == 0) s.n_solve++;
            // (dz below: only the inactive rows are needed -- the count is that algorithmic minimum)
            pas_count(s, (long long)nw * nw + (long long)nw * (mI - nw), 8LL * nw * w.ld + 8LL * nw * (mI - nw));
        }
        LCQ_PROF(w.sc, 1);
        tt_cols_apply(s, w.vb, w.dz, w.dc, true);
        LCQ_PROF(w.sc, 2);
        LCQ_LOOP for (int a = LCQ_TID; a < nw; a += LCQ_NT) {
            const int i = w.widx[a];
            w.dz[i] = (w.st[i] == ST_LOWER) ? (w.lf[i] - w.l[i]) : (w.uf[i] - w.u[i]);
            w.va[a] = w.vb[a];   // keep dy_W (vb is scratch of the working-set updates)
        }
        LCQ_SYNC();
        // ratio tests (QProblem.cpp:5034-5192): groups in the reference's order, list order inside a group
        double best = INFINITY;
        int brank = 0, bcode = -1;   // code = 3 * row + {0 remove, 1 add lower, 2 add upper}
        LCQ_LOOP for (int a = LCQ_TID; a < nw; a += LCQ_NT) {
            const int i = w.widx[a];
            double num = w.y[i], den = -w.va[a];
            if (w.st[i] == ST_UPPER) { num = -num; den = -den; }
            if (den >= kEpsDen && num >= kEpsNum && num < den) {
                int li;
                const int rank = (row_is_bound(s, i, &li) ? (1 << 28) : 0) + w.stamp[i];
                const double r = num / den;
                if (lex2_better(best, brank, bcode, r, rank, 3 * i)) { best = r; brank = rank; bcode = 3 * i; }
            }
        }
        LCQ_LOOP for (int i = LCQ_TID; i < mI; i += LCQ_NT) {
            if (w.st[i] != ST_INACTIVE) continue;
            int li;
            const int grp = row_is_bound(s, i, &li) ? 4 : 2;
            const double dzi = w.dz[i];
            {
                const double num = fmax(w.z[i] - w.l[i], 0.0), den = (w.lf[i] - w.l[i]) - dzi;
                if (den >= kEpsDen && num >= kEpsNum && num < den) {
                    const int rank = (grp << 28) + w.stamp[i];
                    const double r = num / den;
                    if (lex2_better(best, brank, bcode, r, rank, 3 * i + 1)) { best = r; brank = rank; bcode = 3 * i + 1; }
                }
            }
            {
                const double num = fmax(w.u[i] - w.z[i], 0.0), den = dzi - (w.uf[i] - w.u[i]);
                if (den >= kEpsDen && num >= kEpsNum && num < den) {
                    const int rank = ((grp + 1) << 28) + w.stamp[i];
                    const double r = num / den;
                    if (lex2_better(best, brank, bcode, r, rank, 3 * i + 2)) { best = r; brank = rank; bcode = 3 * i + 2; }
                }
            }
        }
        double tmin;
        const int code = block_argmin2(best, brank, bcode, &tmin, nullptr, w.sc);
        tau = (code >= 0) ? tmin : 1.0;
        if (!(tau > 1e-25)) tau = 0.0;   // ZERO (QProblem.cpp:5212)
        // step
        LCQ_PROF(w.sc, 3);
        double hl = 0;
        // step, and the remaining relative homotopy length (QProblem.cpp:5372-5410) in the same sweep
        if (tau > 0.0) LCQ_LOOP for (int a = LCQ_TID; a < nw; a += LCQ_NT) w.y[w.widx[a]] += tau * w.va[a];
        LCQ_LOOP for (int i = LCQ_TID; i < mI; i += LCQ_NT) {
            const double lfi = w.lf[i], ufi = w.uf[i];
            double li = w.l[i], ui = w.u[i];
            if (tau > 0.0) {
                w.z[i] += tau * w.dz[i];
                w.c[i] += tau * w.dc[i];
                li += tau * (lfi - li); ui += tau * (ufi - ui);
                w.l[i] = li; w.u[i] = ui;
            }
            hl = fmax(hl, fmax(fabs(lfi - li) / fmax(fabs(lfi), 1.0), fabs(ufi - ui) / fmax(fabs(ufi), 1.0)));
        }
        if (s.e_moving) LCQ_LOOP for (int e = LCQ_TID; e < mE; e += LCQ_NT) {
            double be = w.bE[e];
            if (tau > 0.0) { be += tau * (w.bEn[e] - be); w.bE[e] = be; }
            hl = fmax(hl, fabs(w.bEn[e] - be) / fmax(fabs(w.bEn[e]), 1.0));
        }
        if (LCQ_TID == 0 && tau > 0.0) s.phi *= (1.0 - tau);
        hl = block_max(hl, w.sc);
        hl = fmax(hl, s.phi * s.len0);
        if (hl <= s.o->qpoases_terminationTolerance) { LCQ_PROF(w.sc, 14); return QP_OK; }
        if (LCQ_TID == 0) s.nwsr++;
        LCQ_SYNC();
        // change the working set (QProblem.cpp:5284-5365)
        bool ramp = false;
        if (code >= 0) {
            const int row = code / 3, kind = code - 3 * row;
            if (kind == 0) {
                LCQ_PROF(w.sc, 4);
                int p = -1;
                LCQ_LOOP for (int a = LCQ_TID; a < nw; a += LCQ_NT) if (w.widx[a] == row) p = a;
                double dummy;
                p = block_argmin2(p >= 0 ? 0.0 : INFINITY, 0, p, &dummy, nullptr, w.sc);
                pas_remove(s, p);
            } else {
                LCQ_PROF(w.sc, 5);
                const int rc = add_with_li(s, row, kind == 1 ? ST_LOWER : ST_UPPER);
                if (rc) return rc;
            }
            ramp = (tau <= qEPS);
        }
        if (ramp) { LCQ_PROF(w.sc, 7); ramping(s); dc_dirty = true; }
        else { LCQ_PROF(w.sc, 6); drift(s); }
    }
    return QP_MAXITER;
}

// End of a QP: bring xq to the end point, polish the active rows in range space, recompute the row activities and
// the multipliers of the eliminated rows.  `ro`: raw operators (Q) for the equality multipliers.
LCQ_DEVN void pas_finish(PQP& s, const RawOps& ro)
{
    const PWork& w = *s.w;
    const PMats& mt = *s.mt;
    const int n = s.d->n, m = s.d->m, mI = mt.mI, mE = mt.mE;
    LCQ_PROF(w.sc, 9);
    // the eliminated rows finish their way here (the homotopy stops within the termination tolerance of the target)
    if (s.e_moving) { LCQ_LOOP for (int e = LCQ_TID; e < mE; e += LCQ_NT) w.bE[e] = w.bEn[e]; LCQ_SYNC(); }
    rebase(s);
    LCQ_PROF(w.sc, 10);
    LCQ_LOOP for (int pass = 0; pass <= kPolishMax; pass++) {
        op_mv_s(mt.oA, w.xq, nullptr, 1.0, w.tm1);   // A_full xq
        LCQ_SYNC();
        pas_count_op(s, mt.oA);
        const int nw = s.nw;
        double rn = 0;
        LCQ_LOOP for (int e = LCQ_TID; e < mE; e += LCQ_NT) { const double r = w.bE[e] - w.tm1[mt.Eidx[e]]; w.tE2[e] = r; rn = fmax(rn, fabs(r)); }
        LCQ_LOOP for (int a = LCQ_TID; a < nw; a += LCQ_NT) {
            const int i = w.widx[a];
            const double r = ((w.st[i] == ST_LOWER) ? w.l[i] : w.u[i]) - w.tm1[mt.Iidx[i]];
            w.va[a] = r;
            rn = fmax(rn, fabs(r));
        }
        double rnE = 0;
        LCQ_LOOP for (int e = LCQ_TID; e < mE; e += LCQ_NT) rnE = fmax(rnE, fabs(w.tE2[e]));
        block_max2(rn, rnE, w.sc);
        if (rn <= kPolishTol || pass == kPolishMax) break;
        if (LCQ_TID == 0) s.n_polish++;
        const bool eq_part = (mE > 0) && (rnE > 0.01 * kPolishTol);
        // dy_W = Sinv (r2W - K[W] r2E) ; dx = P A_W' dy_W + N r2E        (dz is free here: K r2E)
        if (eq_part) {
            op_mv_s(mt.oK, w.tE2, nullptr, 1.0, w.dz);
            LCQ_SYNC();
            LCQ_LOOP for (int a = LCQ_TID; a < nw; a += LCQ_NT) w.va[a] -= w.dz[w.widx[a]];
            LCQ_SYNC();
        }
        if (nw > 0) { sym_apply(w.Sinv, w.ld, nw, w.va, 1.0, w.vb, nullptr, nullptr, nullptr); LCQ_SYNC(); }
        pas_count(s, (long long)nw * nw, 8LL * nw * w.ld);
        LCQ_LOOP for (int r = LCQ_TID; r < m; r += LCQ_NT) w.tm2[r] = 0.0;
        LCQ_SYNC();
        LCQ_LOOP for (int a = LCQ_TID; a < nw; a += LCQ_NT) { const int i = w.widx[a]; w.tm2[mt.Iidx[i]] = w.vb[a]; w.y[i] += w.vb[a]; }
        LCQ_SYNC();
        x_update(s, nullptr, eq_part ? w.tE2 : nullptr);
    }
    // z = (A_full xq)_I (tm1 holds A_full xq of the accepted point), c consistent with it; duals in full row order.
    // The polish moves multipliers by round-off: their signs are re-established as the drift correction of every
    // homotopy step does (QProblem.cpp:5611-5631) -- the stationarity classification reads these signs.
    LCQ_LOOP for (int i = LCQ_TID; i < mI; i += LCQ_NT) {
        const int st = w.st[i];
        double yi = w.y[i];
        yi = (st == ST_LOWER) ? fmax(yi, 0.0) : ((st == ST_UPPER) ? fmin(yi, 0.0) : 0.0);
        w.y[i] = yi; w.yb[i] = yi;
        w.z[i] = w.tm1[mt.Iidx[i]];
    }
    LCQ_SYNC();
    // multipliers of the eliminated rows: yE = N0' (Q xq + gq - A_I' y_I)
    if (mE > 0) {
        LCQ_LOOP for (int r = LCQ_TID; r < m; r += LCQ_NT) w.tm2[r] = mt.isE[r] ? 0.0 : w.y[mt.pos[r]];
        LCQ_SYNC();
        op_mv(ro.Q, w.xq, w.gq, 1.0, w.tn);
        LCQ_SYNC();
        op_mv_s(mt.oAt, w.tm2, w.tn, -1.0, w.tn);
        LCQ_SYNC();
        op_mv_s(mt.oN0t, w.tn, nullptr, 1.0, w.yE);
        LCQ_SYNC();
    }
    LCQ_LOOP for (int r = LCQ_TID; r < m; r += LCQ_NT) w.ys[r] = mt.isE[r] ? w.yE[mt.pos[r]] : w.y[mt.pos[r]];
    LCQ_SYNC();
    (void)n;
    c_from_state(s);
    if (LCQ_TID == 0) s.e_moving = 0;   // the eliminated rows sit on their true bounds from here on
    LCQ_SYNC();
    LCQ_PROF(w.sc, 11);
}

// QProblem::hotstart (QProblem.cpp:446-640): far bounds around the homotopy loop.  gk holds the new gradient.
LCQ_DEVN int pas_hotstart(PQP& s, const RawOps& ro)
{
    const PWork& w = *s.w;
    const PMats& mt = *s.mt;
    const int m = s.d->m, mI = mt.mI;
    LCQ_PROF(w.sc, 14);
    // areBoundsConsistent (:2647-2662) and the largest finite bound (:528-540)
    int bad = 0;
    double far = kFar0;
    LCQ_LOOP for (int r = LCQ_TID; r < m; r += LCQ_NT) {
        double lo, up;
        row_bounds(*s.d, *s.in, r, lo, up);
        if (lo > up + qEPS) bad = 1;
        if (up < qINFTY && up > far) far = up;
        if (lo > -qINFTY && lo < -far) far = -lo;
    }
    LCQ_LOOP for (int e = LCQ_TID; e < mt.mE; e += LCQ_NT) { double lo, up; row_bounds(*s.d, *s.in, mt.Eidx[e], lo, up); w.bEn[e] = lo; }
    bad = block_or(bad, w.sc);
    if (bad) return QP_INFEASIBLE_BOUNDS;
    far = block_max(far, w.sc);
    far_bounds(s, far);
    if (LCQ_TID == 0) s.nwsr = 0;
    set_target(s);
    int rc = QP_OK;
    LCQ_LOOP for (;;) {
        rc = pas_homotopy(s);
        far *= kFarGrow;
        if (rc == QP_INFEASIBLE) {
            if (far >= qINFTY) break;
            far_bounds(s, far);
        } else if (rc == QP_OK) {
            const double tol = far / kFarGrow * s.o->qpoases_boundTolerance;
            int nact = 0;
            LCQ_LOOP for (int i = LCQ_TID; i < mI; i += LCQ_NT) {
                double lo, up;
                row_bounds(*s.d, *s.in, mt.Iidx[i], lo, up);
                if (w.lf[i] > lo && fabs(w.lf[i] - w.z[i]) < tol) nact = 1;
                if (w.uf[i] < up && fabs(w.uf[i] - w.z[i]) < tol) nact = 1;
            }
            nact = block_or(nact, w.sc);
            if (!nact) break;
            if (far >= qINFTY) { rc = QP_UNBOUNDED; break; }
            far_bounds(s, far);
        } else break;
        if (LCQ_TID == 0) s.rampOffset++;
        LCQ_SYNC();
    }
    if (rc == QP_OK) pas_finish(s, ro);
    return rc;
}

// QProblem::init -> solveInitialQP (QProblem.cpp:1301-1471): auxiliary QP at (x0, y0), then the homotopy to (gk, bounds).
// y0A: duals of [A;L;R] (mA) or null, y0box: duals of the box rows or null.
LCQ_DEVN int pas_init(PQP& s, const RawOps& ro, const double* x0, const double* y0A, const double* y0box)
{
    const PWork& w = *s.w;
    const PMats& mt = *s.mt;
    const PDims& d = *s.d;
    const int n = d.n, m = d.m, mI = mt.mI, mE = mt.mE;
    const bool have_y = (y0A != nullptr);
    LCQ_PROF(w.sc, 12);
    if (LCQ_TID == 0) { s.nw = 0; s.stamp_next = mI; s.rampOffset = 0; s.phi = 1.0; s.len0 = 0.0; s.nwsr = 0; s.e_moving = 1; }
    LCQ_LOOP for (int j = LCQ_TID; j < n; j += LCQ_NT) w.xq[j] = x0 ? x0[j] : 0.0;
    LCQ_SYNC();
    op_mv_s(mt.oA, w.xq, nullptr, 1.0, w.tm1);   // A_full x0
    LCQ_SYNC();
    // full-order duals of the guess in tm2
    LCQ_LOOP for (int r = LCQ_TID; r < m; r += LCQ_NT) w.tm2[r] = !have_y ? 0.0 : (r < d.mA ? y0A[r] : (y0box ? y0box[r - d.mA] : 0.0));
    LCQ_SYNC();
    LCQ_LOOP for (int e = LCQ_TID; e < mE; e += LCQ_NT) { w.bE[e] = w.tm1[mt.Eidx[e]]; w.bEb[e] = w.bE[e]; w.yE[e] = w.tm2[mt.Eidx[e]]; }
    // obtainAuxiliaryWorkingSet (:2199-2344, QProblemB.cpp:1479-1617): dz holds the wanted status
    LCQ_LOOP for (int i = LCQ_TID; i < mI; i += LCQ_NT) {
        const int r = mt.Iidx[i];
        const double zi = w.tm1[r], yi = w.tm2[r];
        double lo, up;
        row_bounds(d, *s.in, r, lo, up);
        int aux = ST_INACTIVE;
        if (have_y) aux = (yi > qEPS) ? ST_LOWER : ((yi < -qEPS) ? ST_UPPER : ST_INACTIVE);
        else if (x0) {
            if (zi - lo <= s.o->qpoases_boundTolerance) aux = ST_LOWER;
            else if (up - zi <= s.o->qpoases_boundTolerance) aux = ST_UPPER;
        } else aux = (r >= d.mA) ? ST_LOWER : ST_INACTIVE;   // initialStatusBounds = ST_LOWER
        w.z[i] = zi; w.y[i] = yi; w.st[i] = ST_INACTIVE; w.stamp[i] = i;
        w.dz[i] = (double)aux;
    }
    LCQ_SYNC();
    // setupAuxiliaryWorkingSet (:2351-2539): bounds first, then constraints, each only if linearly independent
    LCQ_LOOP for (int pass = 0; pass < 2; pass++) {
        LCQ_LOOP for (int i = 0; i < mI; i++) {
            const bool isb = mt.Iidx[i] >= d.mA;
            if (isb != (pass == 0)) continue;
            const int aux = (int)w.dz[i];
            if (aux == ST_INACTIVE || s.nw >= s.cap) continue;
            const double p = pas_pivot(s, i);
            if (p > kLITol * mt.Tt[(size_t)i * mt.ldI + i]) pas_append(s, i, aux, p);
        }
    }
    // setupAuxiliaryQPbounds (:2668-2813, useRelaxation)
    LCQ_LOOP for (int i = LCQ_TID; i < mI; i += LCQ_NT) {
        const int st = w.st[i], aux = (int)w.dz[i];
        const double zi = w.z[i];
        if (st == ST_INACTIVE) { w.l[i] = (aux == ST_LOWER) ? zi : zi - kBoundRelax; w.u[i] = (aux == ST_UPPER) ? zi : zi + kBoundRelax; }
        else if (st == ST_LOWER) { w.l[i] = zi; w.u[i] = zi + kBoundRelax; }
        else { w.u[i] = zi; w.l[i] = zi - kBoundRelax; }
        w.yb[i] = w.y[i];
    }
    LCQ_SYNC();
    // setupAuxiliaryQPgradient (:2602-2641): gq = -H x0 + A_full' y  -- H x0 through the raw Q operator is the caller's
    // (gq arrives holding -Q x0); add A_full' y
    op_mv_s(mt.oAt, w.tm2, w.gq, 1.0, w.gq);
    LCQ_SYNC();
    set_target(s);
    ramping(s);
    return pas_hotstart(s, ro);
}

// ------------------------------------------------------------------------------------------------
// The penalty-homotopy loop for one instance (LCQProblem::runSolver, /root/reference/src/LCQProblem.cpp:444-560 and
// helpers :885-1034, :1105-1482) over the parametric active-set subsolver.
// ------------------------------------------------------------------------------------------------
// out = Q v + rho (L'(R v) + R'(L v)) + add ; leaves Lx = L v, Rx = R v
LCQ_DEVN void pqk_apply(const RawOps& ro, double rho, const double* v, const double* add, double* out, const PWork& w)
{
    op_mv(ro.Q, v, add, 1.0, w.tq);
    op_mv(ro.L, v, nullptr, 1.0, w.Lx);
    op_mv(ro.R, v, nullptr, 1.0, w.Rx);
    LCQ_SYNC();
    op_mv(ro.Lt, w.Rx, w.tq, rho, out);
    LCQ_SYNC();
    op_mv(ro.Rt, w.Lx, out, rho, out);
    LCQ_SYNC();
}

LCQ_DEVN void pas_lcqp_loop(PQP& s, const RawOps& ro, unsigned long long instance, double* xout, double* yout, LoopOut& out)
{
    const PDims& d = *s.d;
    const Inst& in = *s.in;
    const int n = d.n, nC = d.nC, nComp = d.nComp, mA = d.mA;
    const PWork& w = *s.w;
    const lcqp_cuda_options& o = *s.o;
    const bool osqp_flavour = (o.qpSolver == 2);
    const int boxOff = osqp_flavour ? 0 : n;

    double hist[kMaxLeyffer];
    int nh = 0;
    double alphak = 1.0, rho = o.initialPenaltyParameter, phi_const = 0.0;
    int outerIter = 0, totalIter = 0, subIter = 0, exitFlag = 0, status = 0, ret = RET_OK;
    out.rhoOpt = 0.0;
    const bool have_gphi = (in.lbL != nullptr) || (in.lbR != nullptr);

    LCQ_LOOP for (int j = LCQ_TID; j < n; j += LCQ_NT) {
        w.xk[j] = in.x0 ? in.x0[j] : 0.0;   // LCQProblem.ipp:138-142
        w.gt[j] = in.g[j];                  // g_tilde = g (LCQProblem.cpp:966-967)
        w.pk[j] = 0.0;
        w.gphi[j] = 0.0;
    }
    LCQ_LOOP for (int r = LCQ_TID; r < d.m; r += LCQ_NT) w.ys[r] = 0.0;
    LCQ_SYNC();
    if (have_gphi) {  // LCQProblem.cpp:970-996
        double part = 0;
        LCQ_LOOP for (int i = LCQ_TID; i < nComp; i += LCQ_NT) part += (in.lbL ? in.lbL[i] : 0.0) * (in.lbR ? in.lbR[i] : 0.0);
        phi_const = block_sum(part, w.sc);
        if (in.lbL) { op_mv(ro.Rt, in.lbL, w.gphi, -1.0, w.gphi); LCQ_SYNC(); }
        if (in.lbR) { op_mv(ro.Lt, in.lbR, w.gphi, -1.0, w.gphi); LCQ_SYNC(); }
    }

    auto phi = [&]() -> double {  // getPhi :1172-1185 ; x'Cx/2 = (Lx)'(Rx)
        op_mv(ro.L, w.xk, nullptr, 1.0, w.Lx);
        op_mv(ro.R, w.xk, nullptr, 1.0, w.Rx);
        LCQ_SYNC();
        double part = 0;
        LCQ_LOOP for (int i = LCQ_TID; i < nComp; i += LCQ_NT) part += w.Lx[i] * w.Rx[i];
        if (have_gphi) LCQ_LOOP for (int j = LCQ_TID; j < n; j += LCQ_NT) part += w.gphi[j] * w.xk[j];
        return phi_const + block_sum(part, w.sc);
    };
    auto update_penalty = [&]() {  // :1199-1214
        nh = 0;
        rho *= o.penaltyUpdateFactor;
        out.rhoOpt = rho;
        if (have_gphi) {
            LCQ_LOOP for (int j = LCQ_TID; j < n; j += LCQ_NT) w.gt[j] = in.g[j] + rho * w.gphi[j];
            LCQ_SYNC();
        }
    };
    auto linearize = [&]() {  // updateLinearization :1105-1112 : gk = rho C xk + g_tilde
        op_mv(ro.L, w.xk, nullptr, 1.0, w.Lx);
        op_mv(ro.R, w.xk, nullptr, 1.0, w.Rx);
        LCQ_SYNC();
        op_mv(ro.Lt, w.Rx, w.gt, rho, w.gk);
        LCQ_SYNC();
        op_mv(ro.Rt, w.Lx, w.gk, rho, w.gk);
        LCQ_SYNC();
    };
    auto solve_qp = [&](bool initial) -> bool {  // solveQPSubproblem :1115-1148
        int rc;
        if (initial) {
            const double* y0A = nullptr;
            const double* y0box = nullptr;
            if (in.y0) { y0A = osqp_flavour ? in.y0 : in.y0 + n; y0box = (d.has_box && !osqp_flavour) ? in.y0 : nullptr; }
            op_mv(ro.Q, w.xk, nullptr, -1.0, w.gq);   // -Q x0 (the auxiliary gradient is completed in pas_init)
            LCQ_SYNC();
            rc = pas_init(s, ro, w.xk, y0A, y0box);
        } else {
            rc = pas_hotstart(s, ro);
        }
        subIter += s.nwsr;
        exitFlag = rc;
        if (rc != QP_OK) { ret = RET_SUBPROBLEM; return false; }
        LCQ_LOOP for (int j = LCQ_TID; j < n; j += LCQ_NT) w.pk[j] = w.xq[j] - w.xk[j];
        LCQ_SYNC();
        return true;
    };

    bool failed = false, success = false;
    // first QP (:452-467)
    if (o.solveZeroPenaltyFirst) {
        LCQ_LOOP for (int j = LCQ_TID; j < n; j += LCQ_NT) w.gk[j] = in.g[j];
        LCQ_SYNC();
    } else {
        linearize();
    }
    if (!solve_qp(true)) failed = true;
    out.rhoOpt = failed ? 0.0 : rho;  // :473

    while (!failed) {
        // updateStep :1240-1243
        LCQ_LOOP for (int j = LCQ_TID; j < n; j += LCQ_NT) w.xk[j] = w.xk[j] + alphak * w.pk[j];
        LCQ_SYNC();
        // updateStationarity :1246-1272 : stat = Qk xk + g_tilde - A_full' yk_A - yk_box
        pqk_apply(ro, rho, w.xk, w.gt, w.stat, w);
        if (nC > 0) { op_mv(ro.At, w.ys, w.stat, -1.0, w.stat); LCQ_SYNC(); }
        op_mv(ro.Lt, w.ys + nC, w.stat, -1.0, w.stat);
        LCQ_SYNC();
        op_mv(ro.Rt, w.ys + nC + nComp, w.stat, -1.0, w.stat);
        LCQ_SYNC();
        if (d.has_box) {
            LCQ_LOOP for (int c = LCQ_TID; c < n; c += LCQ_NT) w.stat[c] -= w.ys[mA + c];
            LCQ_SYNC();
        }
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
                        LCQ_LOOP for (int i = 0; i < nd; i++) if (cur < o.etaDynamicPenalty * hist[i]) { fire = false; break; }
                    }
                    LCQ_LOOP for (int i = 0; i + 1 < nd; i++) hist[i] = hist[i + 1];
                    hist[nd - 1] = cur;
                }
            }
            if (fire) { update_penalty(); outerIter++; }
        }

        double sm = 0;
        LCQ_LOOP for (int j = LCQ_TID; j < n; j += LCQ_NT) sm = fmax(sm, fabs(w.stat[j]));
        sm = block_max(sm, w.sc);
        if (sm < o.stationarityTolerance) {  // :511
            if (phi() < o.complementarityTolerance) {
                // determineStationarityType :1412-1453 on yk_A (the PENALISED duals, :1420), weak set :1456-1482
                const double tc = o.complementarityTolerance;
                int fl = 0;  // bit0: s fails, bit1: m fails, bit2: weakly stationary only
                LCQ_LOOP for (int i = LCQ_TID; i < nComp; i += LCQ_NT) {
                    if (!(w.Lx[i] <= tc && w.Rx[i] <= tc)) continue;
                    const double yl = w.ys[nC + i], yr = w.ys[nC + nComp + i];
                    const double prod = yl * yr, mn = fmin(yl, yr);
                    if (mn < 0) fl |= 1;
                    if (fabs(prod) >= tc && mn <= 0) { if (prod <= tc) fl |= 4; else fl |= 2; }
                }
                const int any = block_or(fl, w.sc);
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
            LCQ_LOOP for (int j = LCQ_TID; j < n; j += LCQ_NT)
                w.xk[j] += perturb_draw(o.perturb_seed, instance, (unsigned)totalIter, (unsigned)j) * kEPS;
            LCQ_SYNC();
        }
        // getOptimalStepLength :1217-1237
        {
            pqk_apply(ro, rho, w.pk, nullptr, w.stat, w);  // Qk pk
            double p1 = 0;
            LCQ_LOOP for (int j = LCQ_TID; j < n; j += LCQ_NT) p1 += w.stat[j] * w.pk[j];
            const double qk = block_sum(p1, w.sc);
            pqk_apply(ro, rho, w.xk, w.gt, w.stat, w);     // Qk xk + g_tilde
            p1 = 0;
            LCQ_LOOP for (int j = LCQ_TID; j < n; j += LCQ_NT) p1 += w.stat[j] * w.pk[j];
            const double lk = block_sum(p1, w.sc);
            alphak = 1.0;
            if (qk > 0 && lk < 0) alphak = fmin(-lk / qk, 1.0);
        }
    }

    // outputs: x = xk ; y = [box duals ; yk_A] (transformDuals :1381-1409 applied on success)
    LCQ_LOOP for (int j = LCQ_TID; j < n; j += LCQ_NT) xout[j] = w.xk[j];
    if (!osqp_flavour)
        LCQ_LOOP for (int j = LCQ_TID; j < n; j += LCQ_NT) yout[j] = d.has_box ? w.ys[mA + j] : 0.0;
    if (success) {
        // Lx, Rx from the last phi() call hold L xk, R xk
        LCQ_LOOP for (int i = LCQ_TID; i < mA; i += LCQ_NT) {
            double v = w.ys[i];
            if (i >= nC && i < nC + nComp) v -= rho * w.Rx[i - nC];
            else if (i >= nC + nComp) v -= rho * w.Lx[i - nC - nComp];
            yout[boxOff + i] = v;
        }
    } else {
        LCQ_LOOP for (int i = LCQ_TID; i < mA; i += LCQ_NT) yout[boxOff + i] = w.ys[i];
    }
    LCQ_SYNC();
    out.ret = ret;
    out.status = status;
    out.iterTotal = totalIter;
    out.iterOuter = outerIter;
    out.subIter = subIter;
    out.exitFlag = exitFlag;
}

// One instance, start to finish.  `mt` is prepared (batch-shared) or is prepared here (per-instance matrices).
// Returns false when Q is not positive definite (the caller takes the regularised solver of lcqp_device.cuh).
LCQ_DEVN bool pas_run_instance(PQP& s, PMats& mt, bool mats_shared, const RawOps& ro, unsigned long long instance,
                               double* xo, double* yo, LoopOut& out, signed char* eq_scratch)
{
    const PDims& d = *s.d;
    const Inst& in = *s.in;
    const PWork& w = *s.w;
    const lcqp_cuda_options& o = *s.o;
    const int nD = d.n + d.mA;
    LCQ_SYNC();
    if (LCQ_TID == 0) { s.mt = &mt; s.nw = 0; s.n_solve = 0; s.n_change = 0; s.n_polish = 0; s.n_mac = 0; s.n_byte = 0; s.nwsr = 0; }
    LCQ_LOOP for (int k = LCQ_TID; k < 8; k += LCQ_NT) w.sc->wph[k] = 0;
    LCQ_SYNC();
    out.ret = 0; out.status = 0; out.iterTotal = 0; out.iterOuter = 0; out.subIter = 0; out.exitFlag = 0; out.rhoOpt = 0;
    bool skip = false;
    if (o.qpSolver == 2 && (in.lb || in.ub)) { out.ret = RET_INVALID_OSQP_BOX; skip = true; }   // LCQProblem.cpp:930-957
    if (!skip) {
        int bad = 0;
        LCQ_LOOP for (int i = LCQ_TID; i < d.nComp; i += LCQ_NT) {
            if (in.lbL && in.lbL[i] <= -INFINITY) bad = 1;   // loadLCQP fails (:747,:767)
            if (in.lbR && in.lbR[i] <= -INFINITY) bad = 1;
        }
        if (block_or(bad, w.sc)) { out.ret = RET_INVALID_LOWER_COMP; skip = true; }
    }
    if (!skip && !mats_shared) {
        LCQ_LOOP for (int r = LCQ_TID; r < d.m; r += LCQ_NT) { double lo, up; row_bounds(d, in, r, lo, up); eq_scratch[r] = (lo == up && lo > -qINFTY && lo < qINFTY) ? 1 : 0; }
        LCQ_SYNC();
        pas_prepare(d, in, mt, eq_scratch, w.sc);
        if (LCQ_TID == 0 && mt.status == 0) pmats_dense_ops(d, mt);
        LCQ_SYNC();
    }
    if (!skip && mt.status == 1) return false;
    if (!skip && mt.status != 0) { out.ret = RET_SUBPROBLEM; out.exitFlag = QP_SETUP; skip = true; }
    if (!skip) {
        if (LCQ_TID == 0) s.cap = pas_cap(d, mt.mE, mt.mI);
        LCQ_SYNC();
        pas_lcqp_loop(s, ro, instance, xo, yo, out);
    }
    if (skip) {
        LCQ_LOOP for (int j = LCQ_TID; j < d.n; j += LCQ_NT) xo[j] = in.x0 ? in.x0[j] : 0.0;
        LCQ_LOOP for (int j = LCQ_TID; j < nD; j += LCQ_NT) yo[j] = 0.0;
    }
    LCQ_SYNC();
    return true;
}

}  // namespace pas
}  // namespace lcqp

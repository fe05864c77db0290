// lcqp_device.cuh -- device code of the B200-native batched LCQP solver (sm_100a).
//
// One GROUP of threads (128; up to four groups per CTA, each with its own named barrier) owns one LCQP instance
// from loadLCQP to the stationarity classification: the whole penalty-homotopy loop of LCQProblem::runSolver
// (/root/reference/src/LCQProblem.cpp:444-560) and the convex QP under it run inside one persistent kernel;
// groups pull instances from a global work counter.  Per-instance iterates, bounds and working set live in
// shared memory; the inverse Schur complement of the working set lives there too when it fits, else in a
// per-group global scratch that stays in L2.  Operators that a batch shares (Q, A, L, R and all that is prepared
// from them) are applied in CSR form from a shared-memory cache that the groups of a CTA share; per-instance
// (dense) operators are applied from L2 in column form.
//
// The hot loops use explicit address spaces (inline PTX: ld.shared / ld.global with 32-bit shared addresses) and
// issue their L2 loads as batches (ld.volatile.global: ptxas would otherwise sink each load to its use); the
// solver state (struct QP) is in shared memory, not in the threads' local memory.  DESIGN.md section 5 has the
// measurements behind each of these choices.
//
// The convex QP  min 1/2 x'Px + q'x  s.t. l <= Ahat x <= u  is solved EXACTLY (the contract of the
// reference's qpOASES subsolver):
//   phase 1 (first QP of an instance only): OSQP-style ADMM in condensed form
//       (P + sigma I + A'RA) xt = sigma x - q + A'(R z - y),  zt = A xt        (osqp auxil.c:161-225)
//     with the matrix inverted once per instance (or once per batch when Q/A/L/R are shared), probing
//     the active set every qp_check_interval iterations                        (osqp polish.c:33-49);
//   phase 2: primal active-set iteration.  Every pass solves the regularised KKT system
//       [P + dI, Aw'; Aw, -dI] [dx; dlam] = residual of the unregularised system (osqp polish.c:134-181, :232-300)
//     from the CURRENT point, tests the step against the inactive rows (ratio test) and either adds the
//     blocking row or takes the step; at a converged point the multiplier signs decide (drop or accept).
//     The KKT solve is a block elimination with three levels, all but the last working-set independent:
//       Hinv = (P+dI)^-1                      (n x n, prepared once)
//       SEinv = (A_E Hinv A_E' + dI)^-1       (rows E that are equalities l = u: always active, prepared once)
//       Tinv  = (T[W,W] + dI)^-1,  T = A_I Hk A_I',  Hk = Hinv - Hinv A_E' SEinv A_E Hinv
//     Only Tinv (order = number of active INEQUALITY rows) is per instance; it is kept as an explicit packed
//     symmetric inverse in shared memory and updated by bordering in O(|W|^2) per working-set change.
//   later QPs of the instance hot-start phase 2 from the previous optimum, multipliers and working set
//   (the analogue of qpOASES' hotstart, /root/reference/src/SubsolverQPOASES.cpp:154-160).
// A QP solution is only accepted when it satisfies the KKT conditions of the full QP.
//
// The file also compiles as plain single-threaded C++ (LCQP_HOST_EMU): tests/emu builds that variant so
// that the CPU test-suite covers the kernel's logic; the product never builds or calls it.
#pragma once

#include <math.h>
#include <stdint.h>
#include <string.h>
#ifdef LCQP_HOST_EMU
#include <stdio.h>
#include <stdlib.h>
#endif

#ifdef LCQP_HOST_EMU
#define LCQ_DEV static inline
#define LCQ_DEVN static
#define LCQ_HD
#define LCQ_TID 0
#define LCQ_NT 1
#define LCQ_LANE 0
#define LCQ_WARP 0
#define LCQ_NWARP 1
#define LCQ_LANES 1
#define LCQ_SYNC() ((void)0)
#define LCQ_LOOP
#else
#define LCQ_DEV __device__ __forceinline__
#define LCQ_DEVN __device__ __noinline__
#define LCQ_HD __host__ __device__
#define LCQ_TID ((int)threadIdx.x)
#define LCQ_NT ((int)blockDim.x)
#define LCQ_LANE ((int)(threadIdx.x & 31))
#define LCQ_WARP ((int)(threadIdx.x >> 5))
#define LCQ_NWARP ((int)(blockDim.x >> 5))
#define LCQ_LANES 32
// A CTA holds blockDim.y independent GROUPS of blockDim.x threads; a group owns one instance at a time and
// synchronises on its own named barrier (id 1 + group), so that groups run their own control flow while
// sharing the CTA's operator cache.
#define LCQ_GROUP ((int)threadIdx.y)
#define LCQ_SYNC() asm volatile("bar.sync %0, %1;" ::"r"(1 + (int)threadIdx.y), "r"((int)blockDim.x) : "memory")
// The solver is bound by instruction fetch, not by issue slots: its hot path must stay inside the 32 KB
// L1.5 instruction cache, so no loop is unrolled.
#define LCQ_LOOP _Pragma("unroll 1")
#endif

#include "../../include/lcqp_cuda.h"

// inlining of the two dense kernels (code size vs. register allocation)
#ifdef LCQP_SYM_NOINLINE
#define LCQ_SYM_INL LCQ_DEVN
#else
#define LCQ_SYM_INL LCQ_DEV
#endif
#ifdef LCQP_R1_INLINE
#define LCQ_R1_INL LCQ_DEV
#else
#define LCQ_R1_INL LCQ_DEVN
#endif

namespace lcqp {

// Pointer into the CTA's shared memory.  The accessors tell the compiler the address space
// (__builtin_assume(__isShared)), so that element accesses compile to LDS/STS with 32-bit addresses instead of
// generic loads; it converts implicitly to a plain (generic) pointer where a callee takes one.
template <class T> struct SPtr {
    T* p;
    LCQ_HD SPtr() : p(nullptr) {}
    LCQ_HD SPtr(T* q) : p(q) {}
#if defined(LCQP_HOST_EMU) || defined(LCQP_NO_ASSUME)
    LCQ_HD T& operator[](int i) const { return p[i]; }
    LCQ_HD T* operator->() const { return p; }
    LCQ_HD T& operator*() const { return *p; }
#else
    __device__ __forceinline__ T& operator[](int i) const { __builtin_assume(__isShared(p)); return p[i]; }
    __device__ __forceinline__ T* operator->() const { __builtin_assume(__isShared(p)); return p; }
    __device__ __forceinline__ T& operator*() const { __builtin_assume(__isShared(p)); return *p; }
#endif
    LCQ_HD operator T*() const { return p; }
};
typedef SPtr<double> SVec;

#if defined(LCQP_HOST_EMU) || defined(LCQP_NO_ASSUME)
#define LCQ_ASSUME_SHARED(ptr) ((void)0)
#define LCQ_ASSUME_GLOBAL(ptr) ((void)0)
#else
#define LCQ_ASSUME_SHARED(ptr) __builtin_assume(__isShared(ptr))
#define LCQ_ASSUME_GLOBAL(ptr) __builtin_assume(__isGlobal(ptr))
#endif

#ifndef LCQP_HOST_EMU
// Explicit address-space accessors (inline PTX).  The hot loops address shared memory by 32-bit shared-window
// addresses (LDS/STS, immediate offsets) and the L2-resident matrices by ld.global/st.global, instead of generic
// 64-bit loads.  The asm statements are volatile (kept in program order among themselves and against the named
// barriers, which clobber memory); the stores clobber memory.  A function that uses them reaches an array ONLY
// through them between two barriers.
LCQ_DEV unsigned saddr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
LCQ_DEV double lds64(unsigned a) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a)); return v; }
LCQ_DEV void lds128(unsigned a, double& x, double& y) { asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x), "=d"(y) : "r"(a)); }
LCQ_DEV int lds32(unsigned a) { int v; asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
LCQ_DEV unsigned ldsu16(unsigned a) { unsigned short v; asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a)); return v; }
LCQ_DEV void sts64(unsigned a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }
LCQ_DEV double ldg64(const double* p) { double v; asm volatile("ld.global.f64 %0, [%1];" : "=d"(v) : "l"(p)); return v; }
LCQ_DEV void ldg128(const double* p, double& x, double& y) { asm volatile("ld.global.v2.f64 {%0, %1}, [%2];" : "=d"(x), "=d"(y) : "l"(p)); }
// volatile: ptxas keeps volatile loads in program order, i.e. a batch of them is ISSUED before the first use (it
// otherwise sinks each load to its use when the enclosing function is large) -- that is what hides the L2 latency
LCQ_DEV void ldg128v(const double* p, double& x, double& y) { asm volatile("ld.volatile.global.v2.f64 {%0, %1}, [%2];" : "=d"(x), "=d"(y) : "l"(p)); }
LCQ_DEV double ldg64v(const double* p) { double v; asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(p)); return v; }
LCQ_DEV void lds128v(unsigned a, double& x, double& y) { asm volatile("ld.volatile.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x), "=d"(y) : "r"(a)); }
LCQ_DEV void stg64(double* p, double v) { asm volatile("st.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory"); }
LCQ_DEV void stg128(double* p, double x, double y) { asm volatile("st.global.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(x), "d"(y) : "memory"); }
#endif

constexpr double kEPS = 2.221e-16;       // LCQPow Utilities::EPS (Utilities.hpp:350)
constexpr double kQPInf = 1e20;          // Utilities::INFTY (Utilities.hpp:362)
constexpr double kRhoMin = 1e-6;         // osqp constants.h:48
constexpr double kRhoTol = 1e-4;         // constants.h:50
constexpr double kRhoEqOverIneq = 1e3;   // constants.h:51
constexpr double kMinScaling = 1e-4;     // constants.h:94
constexpr double kMaxScaling = 1e4;
constexpr int kScalingIters = 10;        // constants.h:56
constexpr int kMaxLeyffer = 16;
constexpr double kFloorTol = 1e-11;       // KKT residual below which refinement is at its round-off floor
constexpr double kResTol = 1e-9;         // residual of an EQP solve that still counts as solved
constexpr int kLongRow = 32;             // CSR rows longer than this are reduced by a warp

enum { RET_OK = 0, RET_INVALID_OSQP_BOX = 110, RET_INVALID_LOWER_COMP = 120, RET_MAX_ITER = 200, RET_MAX_PEN = 201,
       RET_SUBPROBLEM = 203, RET_OSQP_GUESS = 208 };

// ------------------------------------------------------------------------------------------------
// A linear operator y = M v, M logically rows x cols.
//   rp != null : CSR (rp[rows+1], 16-bit column indices ci, values va); rows with more than kLongRow entries
//                are also listed in lrows
//   else trans == 0 : dense row-major, M[r][c] = dense[r*ld + c]
//   else            : the transpose of a dense row-major matrix, M[r][c] = dense[c*ld + r]
// ------------------------------------------------------------------------------------------------
struct Op {
    const double* dense;
    const int* rp;
    const unsigned short* ci;
    const double* va;
    const int* lrows;
    int rows, cols, ld, trans, nlong;
    int smem;   // the CSR arrays were copied into shared memory (operator cache)
};

LCQ_DEV Op dense_op(const double* M, int rows, int cols, int ld, int trans)
{
    Op o;
    o.dense = M; o.rp = nullptr; o.ci = nullptr; o.va = nullptr; o.lrows = nullptr;
    o.rows = rows; o.cols = cols; o.ld = ld; o.trans = trans; o.nlong = 0; o.smem = 0;
    return o;
}

// Prepared (scaled) operands of the QP: depend on Q, A_full, the row types and the options only.
struct Mats {
    double* P;      // n*n   D Q D
    double* A;      // m*n   E Ahat D
    double* At;     // n*m   its transpose (dense operators read the output index contiguously)
    double* D;      // n
    double* E;      // m
    double* Hinv;   // n*n   (P + delta I)^-1
    double* AH;     // m*n   A Hinv
    double* T;      // m*m   A Hinv A', then the equality rows eliminated (see prepare_factor)
    double* Minv;   // n*n   (P + sigma I + A' R A)^-1
    double* SEinv;  // ldE*ldE (leading dimension d.ldE), order mE
    double* AHE;    // ldE*n  the rows of A Hinv that belong to the static equality block (compact)
    int* eidx;      // equality rows that are always in the working set (mE of them)
    signed char* ctype;  // m: -1 free row, 0 inequality, 1 equality, 2 equality found dependent (never active)
    const double* SEinvP;  // packed (lower triangle) copy of SEinv in shared memory, or null
    int mE;
    int status;     // 0 ok, 1 factorisation failed
    int cache_bytes_se;    // shared memory that the packed SEinv would take,
    int cache_bytes_hot;   // ... the operators applied in every pass,
    int cache_bytes_raw;   // ... and the operators of the outer loop
    Op oP, oA, oAt, oHinv, oAHE, oAHtE, oMinv;   // oAHE: mE x n, oAHtE: n x mE
};

// Operators on the UNSCALED matrices, used by the outer loop.
struct RawOps {
    Op Q, L, R, Lt, Rt, At;   // At: nV x nC
};

// Bump allocator over global memory for the CSR copies built on the device.
struct CsrPool {
    int* ibuf;
    double* dbuf;
    int icap, dcap;
    int* used;   // [0] ints used, [1] doubles used
};

// Unscaled instance data (global memory), loadLCQP argument order.
struct Inst {
    const double *Q, *g, *L, *R, *lbL, *ubL, *lbR, *ubR, *A, *lbA, *ubA, *lb, *ub, *x0, *y0;
};

struct Dims {
    int n, nC, nComp, mA, m, has_box;
    int cap;    // capacity of the inequality working set (order of Tinv)
    int ldE;    // capacity / leading dimension of the static equality block SEinv
    int capE;   // entries of the shared-memory vectors over the equality block (mE when known, else ldE)
};

// Dimensions of a problem with nV variables, nC general rows, nComp complementarity pairs (and box rows).
inline LCQ_HD Dims make_dims(int nV, int nC, int nComp, int has_box)
{
    Dims d;
    d.n = nV; d.nC = nC; d.nComp = nComp; d.mA = nC + 2 * nComp; d.has_box = has_box;
    d.m = d.mA + (has_box ? nV : 0);
    d.ldE = ((d.m < d.n ? d.m : d.n) + 1) & ~1;   // even: the full-storage SEinv is read by 16-byte loads
    d.capE = d.ldE > 0 ? d.ldE : 1;
    d.cap = d.m < d.n + 8 ? d.m : d.n + 8;   // a linearly independent working set has at most n rows
    if (d.cap < 1) d.cap = 1;
    return d;
}

// Once the order mE of the static equality block is known the per-instance buffers shrink.
inline LCQ_HD void shrink_dims(Dims& d, int mE)
{
    if (mE < 0 || mE > d.ldE) return;
    d.capE = mE > 0 ? mE : 1;
    const int capI = d.n - mE + 8, mI = d.m - mE;
    d.cap = capI < mI ? capI : mI;
    if (d.cap < 1) d.cap = 1;
}

// block-shared scalars
struct Scalars {
    double red[64];
    int ired[32];
    double fred[2][3][8];   // double-buffered partials of the single-barrier reductions (solver CTAs: <= 8 warps)
    int fired[2][8];
    int ior[32];
    int bidx;
    int flag;
    int pad[2];
    int wph[8];   // per-warp phase of the single-barrier reductions (lane 0 toggles it after the barrier)
#ifdef LCQP_PROFILE
    long long prof[16];   // cycles per section (thread 0 of the group), see LCQ_PROF
    long long prof_last;
    int prof_cur;
#endif
};

// Section timing (development aid, -DLCQP_PROFILE): thread 0 of a group charges the cycles since the last
// switch to the section that was current, then makes `id` current.
#if defined(LCQP_PROFILE) && !defined(LCQP_HOST_EMU)
#define LCQ_PROF(sc_, id)                                                                   \
    do {                                                                                    \
        if (LCQ_TID == 0) {                                                                 \
            const long long t_ = clock64();                                                 \
            (sc_)->prof[(sc_)->prof_cur] += t_ - (sc_)->prof_last;                          \
            (sc_)->prof_last = t_;                                                          \
            (sc_)->prof_cur = (id);                                                         \
        }                                                                                   \
    } while (0)
#else
#define LCQ_PROF(sc_, id) ((void)0)
#endif

// Working set of one instance (shared memory, except the outer-loop vectors and Tinv when they do not fit).
struct Work {
    // QP (scaled space): always in shared memory
    SVec q, x, xa, px, r1, u, t, dx;                               // n
    SVec z, y, l, ub, lam, dlam, r2, zx, zp, w, yf;                // m
    SVec dI, lI;                                                   // cap
    SVec cE, vE;                                                   // mE
    SPtr<signed char> W, Wtry, Wfail, ctype, pin;                  // m
    SPtr<int> idx;                                                 // cap: rows of the inequality working set
    double* Tinv;                                                  // shared memory: packed lower triangle, cap*(cap+1)/2;
    int tld;                                                       // global memory: full storage, leading dimension tld (0 = packed)
    SVec ys;                                                       // m  accepted multipliers, unscaled, qpOASES sign
    // outer loop (unscaled): shared memory when it fits, else global scratch
    double *xk, *pk, *gk, *gt, *gphi, *stat, *tn;                  // n
    double *Lx, *Rx;                                               // nComp
    SPtr<Scalars> sc;
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
    LCQ_LOOP for (int k = 0; k < LCQ_NWARP; k++) s += sc->red[k];
    return s;
}

LCQ_DEV double block_max(double v, Scalars* sc)
{
    v = warp_max(v);
    LCQ_SYNC();
    if (LCQ_LANE == 0) sc->red[LCQ_WARP] = v;
    LCQ_SYNC();
    double s = sc->red[0];
    LCQ_LOOP for (int k = 1; k < LCQ_NWARP; k++) s = fmax(s, sc->red[k]);
    return s;
}

// two sums with one pair of barriers
LCQ_DEV void block_sum2(double& a, double& b, Scalars* sc)
{
    a = warp_sum(a);
    b = warp_sum(b);
    LCQ_SYNC();
    if (LCQ_LANE == 0) { sc->red[LCQ_WARP] = a; sc->red[32 + LCQ_WARP] = b; }
    LCQ_SYNC();
    double s = 0, t = 0;
    LCQ_LOOP for (int k = 0; k < LCQ_NWARP; k++) { s += sc->red[k]; t += sc->red[32 + k]; }
    a = s; b = t;
}

// block-wide OR of small non-negative ints
LCQ_DEV int block_or(int v, Scalars* sc)
{
#ifndef LCQP_HOST_EMU
    v = (int)__reduce_or_sync(0xffffffffu, (unsigned)v);
    LCQ_SYNC();
    if (LCQ_LANE == 0) sc->ior[LCQ_WARP] = v;
    LCQ_SYNC();
    v = 0;
    LCQ_LOOP for (int k = 0; k < LCQ_NWARP; k++) v |= sc->ior[k];
#else
    (void)sc;
#endif
    return v;
}

// block-wide argmax with deterministic tie-break (smallest index); entries with i < 0 do not count.
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
    LCQ_LOOP for (int k = 1; k < LCQ_NWARP; k++) {
        double ov = sc->red[k];
        int oi = sc->ired[k];
        if (oi >= 0 && (bi < 0 || ov > bv || (ov == bv && oi < bi))) { bv = ov; bi = oi; }
    }
    *vout = bv;
    return bi;
}

// ---- single-barrier reductions for the inner passes ---------------------------------------------------
// The partials go to one of two buffers, alternating from call to call (`ph` is a per-thread copy of the
// same counter): the barrier of call k+1 separates the reads of call k from the writes of call k+2.
// Only for CTAs of at most 8 warps.
LCQ_DEV double fast_max(double v, Scalars* sc)
{
    v = warp_max(v);
    const int ph = sc->wph[LCQ_WARP];
    double* buf = sc->fred[ph][0];
    if (LCQ_LANE == 0) buf[LCQ_WARP] = v;
    LCQ_SYNC();
    if (LCQ_LANE == 0) sc->wph[LCQ_WARP] = ph ^ 1;
    double r = buf[0];
    LCQ_LOOP for (int k = 1; k < LCQ_NWARP; k++) r = fmax(r, buf[k]);
    return r;
}

LCQ_DEV void fast_sum2(double& a, double& b, Scalars* sc)
{
    a = warp_sum(a);
    b = warp_sum(b);
    const int ph = sc->wph[LCQ_WARP];
    double* b0 = sc->fred[ph][0];
    double* b1 = sc->fred[ph][1];
    if (LCQ_LANE == 0) { b0[LCQ_WARP] = a; b1[LCQ_WARP] = b; }
    LCQ_SYNC();
    if (LCQ_LANE == 0) sc->wph[LCQ_WARP] = ph ^ 1;
    double x = 0, y = 0;
    LCQ_LOOP for (int k = 0; k < LCQ_NWARP; k++) { x += b0[k]; y += b1[k]; }
    a = x; b = y;
}

// lexicographic selection over (alpha ascending, weight descending, index ascending); i < 0: no candidate.
LCQ_DEV bool lex_better(double a, double wgt, int i, double a2, double w2, int i2)
{
    if (i2 < 0) return false;
    if (i < 0) return true;
    if (a2 != a) return a2 < a;
    if (w2 != wgt) return w2 > wgt;
    return i2 < i;
}

LCQ_DEV int fast_argmin_lex(double a, double wgt, int i, double* aout, Scalars* sc)
{
#ifndef LCQP_HOST_EMU
    for (int o = 16; o > 0; o >>= 1) {
        const double a2 = __shfl_xor_sync(0xffffffffu, a, o);
        const double w2 = __shfl_xor_sync(0xffffffffu, wgt, o);
        const int i2 = __shfl_xor_sync(0xffffffffu, i, o);
        if (lex_better(a, wgt, i, a2, w2, i2)) { a = a2; wgt = w2; i = i2; }
    }
#endif
    const int ph = sc->wph[LCQ_WARP];
    double* ba = sc->fred[ph][0];
    double* bw = sc->fred[ph][1];
    int* bi = sc->fired[ph];
    if (LCQ_LANE == 0) { ba[LCQ_WARP] = a; bw[LCQ_WARP] = wgt; bi[LCQ_WARP] = i; }
    LCQ_SYNC();
    if (LCQ_LANE == 0) sc->wph[LCQ_WARP] = ph ^ 1;
    double ra = ba[0], rw = bw[0];
    int ri = bi[0];
    LCQ_LOOP for (int k = 1; k < LCQ_NWARP; k++)
        if (lex_better(ra, rw, ri, ba[k], bw[k], bi[k])) { ra = ba[k]; rw = bw[k]; ri = bi[k]; }
    *aout = ra;
    return ri;
}

// ------------------------------------------------------------------------------------------------
// operator application
// ------------------------------------------------------------------------------------------------
// out[r] = (init ? init[iidx ? iidx[r] : r] : 0) + scale * sum_c M[r][c] v[c]      (no barrier inside)
// VS: v, init and out live in shared memory (the inner passes of the QP solver); OS: so do the CSR arrays
// (operator cache).  The address spaces are compile-time facts of each instantiation, so that the loops
// compile to LDS/STS (or LDG) instead of generic accesses.
template <bool VS, bool OS>
LCQ_DEV void csr_mv(const int* __restrict__ rp, const unsigned short* __restrict__ ci, const double* __restrict__ va,
                    const int* __restrict__ lrows, int rows, int nlong, const double* v, const double* init, bool has_init,
                    double scale, double* out, const int* iidx)
{
#ifndef LCQP_HOST_EMU
    if (VS && OS) {
        // everything but iidx lives in shared memory: 32-bit shared addresses, LDS/STS (same operation order
        // as the generic loops below, so the results are bit-identical)
        const unsigned rps = saddr(rp), cis = saddr(ci), vas = saddr(va), vs = saddr(v), is = saddr(init), os = saddr(out), lrs = saddr(lrows);
        // two rows of a thread (r and r + NT) advance in lockstep: two independent load -> multiply -> add chains
        LCQ_LOOP for (int r = LCQ_TID; r < rows; r += 2 * LCQ_NT) {
            const int rb = r + LCQ_NT;
            const bool hb = rb < rows;
            int ka = lds32(rps + 4u * (unsigned)r), ea = lds32(rps + 4u * (unsigned)r + 4u);
            int kb = hb ? lds32(rps + 4u * (unsigned)rb) : 0, eb = hb ? lds32(rps + 4u * (unsigned)rb + 4u) : 0;
            const bool sa_ok = ea - ka <= kLongRow, sb_ok = hb && eb - kb <= kLongRow;
            if (!sa_ok) ea = ka;
            if (!sb_ok) eb = kb;
            const double ia = (has_init && sa_ok) ? lds64(is + 8u * (unsigned)(iidx ? iidx[r] : r)) : 0.0;
            const double ib = (has_init && sb_ok) ? lds64(is + 8u * (unsigned)(iidx ? iidx[rb] : rb)) : 0.0;
            double sa = 0, sb = 0;
            LCQ_LOOP while (ka < ea || kb < eb) {
                if (ka < ea) { sa += lds64(vas + 8u * (unsigned)ka) * lds64(vs + 8u * ldsu16(cis + 2u * (unsigned)ka)); ka++; }
                if (kb < eb) { sb += lds64(vas + 8u * (unsigned)kb) * lds64(vs + 8u * ldsu16(cis + 2u * (unsigned)kb)); kb++; }
            }
            if (sa_ok) sts64(os + 8u * (unsigned)r, ia + scale * sa);
            if (sb_ok) sts64(os + 8u * (unsigned)rb, ib + scale * sb);
        }
        LCQ_LOOP for (int a = LCQ_NWARP - 1 - LCQ_WARP; a < nlong; a += LCQ_NWARP) {
            const int r = lds32(lrs + 4u * (unsigned)a);
            const int k1 = lds32(rps + 4u * (unsigned)r + 4u);
            double s0 = 0, s1 = 0;
            int k = lds32(rps + 4u * (unsigned)r) + LCQ_LANE;
            LCQ_LOOP for (; k + LCQ_LANES < k1; k += 2 * LCQ_LANES) {
                s0 += lds64(vas + 8u * (unsigned)k) * lds64(vs + 8u * ldsu16(cis + 2u * (unsigned)k));
                s1 += lds64(vas + 8u * (unsigned)(k + LCQ_LANES)) * lds64(vs + 8u * ldsu16(cis + 2u * (unsigned)(k + LCQ_LANES)));
            }
            if (k < k1) s0 += lds64(vas + 8u * (unsigned)k) * lds64(vs + 8u * ldsu16(cis + 2u * (unsigned)k));
            const double s = warp_sum(s0 + s1);
            if (LCQ_LANE == 0) {
                const double i0 = has_init ? lds64(is + 8u * (unsigned)(iidx ? iidx[r] : r)) : 0.0;
                sts64(os + 8u * (unsigned)r, i0 + scale * s);
            }
        }
        return;
    }
#endif
#define LCQ_INIT(r) (has_init ? init[iidx ? iidx[r] : (r)] : 0.0)
    LCQ_LOOP for (int r = LCQ_TID; r < rows; r += LCQ_NT) {
        const int k0 = rp[r], k1 = rp[r + 1];
        if (k1 - k0 > kLongRow && LCQ_LANES > 1) continue;
        double s = 0;
        LCQ_LOOP for (int k = k0; k < k1; k++) s += va[k] * v[ci[k]];
        out[r] = LCQ_INIT(r) + scale * s;
    }
    if (LCQ_LANES > 1)
        LCQ_LOOP for (int a = LCQ_NWARP - 1 - LCQ_WARP; a < nlong; a += LCQ_NWARP) {
            const int r = lrows[a];
            const int k1 = rp[r + 1];
            double s0 = 0, s1 = 0;
            int k = rp[r] + LCQ_LANE;
            LCQ_LOOP for (; k + LCQ_LANES < k1; k += 2 * LCQ_LANES) { s0 += va[k] * v[ci[k]]; s1 += va[k + LCQ_LANES] * v[ci[k + LCQ_LANES]]; }
            if (k < k1) s0 += va[k] * v[ci[k]];
            const double s = warp_sum(s0 + s1);
            if (LCQ_LANE == 0) out[r] = LCQ_INIT(r) + scale * s;
        }
#undef LCQ_INIT
}

template <bool VS>
LCQ_DEVN void op_mv_t(const Op& opr, const double* v, const double* init, double scale, double* out, const int* iidx = nullptr)
{
    const Op op = opr;
    const bool has_init = init != nullptr;
    if (!has_init) init = v;   // a valid address for the address-space assumption; never read
    if (op.rp) {
        if (op.smem == 1) csr_mv<VS, true>(op.rp, op.ci, op.va, op.lrows, op.rows, op.nlong, v, init, has_init, scale, out, iidx);
        else csr_mv<VS, false>(op.rp, op.ci, op.va, op.lrows, op.rows, op.nlong, v, init, has_init, scale, out, iidx);
        return;
    }
#ifndef LCQP_HOST_EMU
    if (VS) { LCQ_ASSUME_SHARED(v); LCQ_ASSUME_SHARED(init); LCQ_ASSUME_SHARED(out); }
#endif
#define LCQ_INIT(r) (has_init ? init[iidx ? iidx[r] : (r)] : 0.0)
#ifndef LCQP_HOST_EMU
    // Dense operators live in global memory (L2-resident).  Eight loads per thread are issued as a batch
    // (volatile: see ldg128v); the sums run in the same order as the plain loops of the host build.
    const double* __restrict__ M = op.dense;
    if (!op.trans) {
        const int rows = op.rows, cols = op.cols, ld = op.ld;
        LCQ_LOOP for (int r0 = LCQ_WARP; r0 < rows; r0 += 8 * LCQ_NWARP) {
            double sm[8];
            unsigned off[8];
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const int rk = r0 + k * LCQ_NWARP;
                off[k] = (unsigned)(rk < rows ? rk : r0) * (unsigned)ld;
                sm[k] = 0;
            }
            LCQ_LOOP for (int c = LCQ_LANE; c < cols; c += LCQ_LANES) {
                double mk[8];
#pragma unroll
                for (int k = 0; k < 8; k++) mk[k] = ldg64v(M + (off[k] + (unsigned)c));
                const double vc = v[c];
#pragma unroll
                for (int k = 0; k < 8; k++) sm[k] += mk[k] * vc;
            }
#pragma unroll
            for (int k = 0; k < 8; k++) sm[k] = warp_sum(sm[k]);
            if (LCQ_LANE == 0) {
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const int rk = r0 + k * LCQ_NWARP;
                    if (rk < rows) out[rk] = LCQ_INIT(rk) + scale * sm[k];
                }
            }
        }
    } else if (VS && !(op.ld & 1) && !(reinterpret_cast<uintptr_t>(M) & 15) && !(saddr(v) & 15u)) {
        // out[r] = sum_c M[c*ld + r] v[c], v in shared memory: the scheme of sym_apply's L2 branch -- a pair of
        // adjacent lanes owns outputs (2p, 2p+1) by 16-byte loads and splits the c range in two halves (the first
        // of even length), twelve loads in flight per thread, one shuffle joins the halves
        const int rows = op.rows, cols = op.cols;
        const unsigned ldu = (unsigned)op.ld, vs = saddr(v);
        const int npair = (rows + 1) >> 1;
        int half = (((cols + 1) >> 1) + 1) & ~1;
        if (half > cols) half = cols;
        LCQ_LOOP for (int t0 = 0; t0 < 2 * npair; t0 += LCQ_NT) {
            const int t = t0 + LCQ_TID;
            const bool act = t < 2 * npair;
            const int pr = t >> 1, h = t & 1;
            int c = act ? (h ? half : 0) : 0;
            const int c1 = act ? (h ? cols : half) : 0;
            const double* p = M + (size_t)((unsigned)c * ldu + 2u * (unsigned)pr);
            unsigned va = vs + 8u * (unsigned)c;
            double x0 = 0, x1 = 0, y0 = 0, y1 = 0;
            LCQ_LOOP for (; c + 12 <= c1; c += 12) {
                double m0[12], m1[12];
#pragma unroll
                for (int k = 0; k < 12; k++) { ldg128v(p, m0[k], m1[k]); p += ldu; }
#pragma unroll
                for (int k = 0; k < 12; k += 2) {
                    double v0, v1;
                    lds128v(va + 8u * k, v0, v1);
                    x0 += m0[k] * v0; y0 += m1[k] * v0;
                    x1 += m0[k + 1] * v1; y1 += m1[k + 1] * v1;
                }
                va += 96u;
            }
            LCQ_LOOP for (; c + 2 <= c1; c += 2) {
                double m00, m10, m01, m11, v0, v1;
                ldg128v(p, m00, m10); p += ldu;
                ldg128v(p, m01, m11); p += ldu;
                lds128(va, v0, v1);
                va += 16u;
                x0 += m00 * v0; y0 += m10 * v0;
                x1 += m01 * v1; y1 += m11 * v1;
            }
            if (c < c1) {
                double m0, m1;
                ldg128v(p, m0, m1);
                const double v0 = lds64(va);
                x0 += m0 * v0; y0 += m1 * v0;
            }
            double sa = x0 + x1, sb = y0 + y1;
            sa += __shfl_xor_sync(0xffffffffu, sa, 1);
            sb += __shfl_xor_sync(0xffffffffu, sb, 1);
            const int r = 2 * pr + h;
            if (act && r < rows) out[r] = LCQ_INIT(r) + scale * (h ? sb : sa);
        }
    } else {
        const int rows = op.rows, cols = op.cols, ld = op.ld;
        LCQ_LOOP for (int r = LCQ_TID; r < rows; r += LCQ_NT) {
            double sacc = 0;
            int c = 0;
            LCQ_LOOP for (; c + 8 <= cols; c += 8) {
                double mk[8];
#pragma unroll
                for (int k = 0; k < 8; k++) mk[k] = ldg64v(M + ((unsigned)(c + k) * (unsigned)ld + (unsigned)r));
#pragma unroll
                for (int k = 0; k < 8; k++) sacc += mk[k] * v[c + k];
            }
            LCQ_LOOP for (; c < cols; c++) sacc += ldg64v(M + ((unsigned)c * (unsigned)ld + (unsigned)r)) * v[c];
            out[r] = LCQ_INIT(r) + scale * sacc;
        }
    }
#else
    if (!op.trans) {
        LCQ_LOOP for (int r = 0; r < op.rows; r++) {
            double sacc = 0;
            LCQ_LOOP for (int c = 0; c < op.cols; c++) sacc += op.dense[(size_t)r * op.ld + c] * v[c];
            out[r] = LCQ_INIT(r) + scale * sacc;
        }
    } else {
        LCQ_LOOP for (int r = 0; r < op.rows; r++) {
            double sacc = 0;
            LCQ_LOOP for (int c = 0; c < op.cols; c++) sacc += op.dense[(size_t)c * op.ld + r] * v[c];
            out[r] = LCQ_INIT(r) + scale * sacc;
        }
    }
#endif
#undef LCQ_INIT
}

// generic address spaces (outer loop, preparation) / everything in shared memory (inner passes)
LCQ_DEV void op_mv(const Op& op, const double* v, const double* init, double scale, double* out, const int* iidx = nullptr)
{
    op_mv_t<false>(op, v, init, scale, out, iidx);
}
LCQ_DEV void op_mv_s(const Op& op, const double* v, const double* init, double scale, double* out, const int* iidx = nullptr)
{
    op_mv_t<true>(op, v, init, scale, out, iidx);
}

// out[a] = sum_c M[idx[a]][c] v[c] - (sub ? sub[idx[a]] : 0)   for a < na   (rows selected by idx; op not transposed)
// idx, v, sub, out live in shared memory (inner passes only).
template <bool OS>
LCQ_DEV void csr_mv_rows(const int* __restrict__ rp, const unsigned short* __restrict__ ci, const double* __restrict__ va,
                         const int* idx, int na, const double* v, const double* sub, bool has_sub, double* out)
{
#ifndef LCQP_HOST_EMU
    if (OS) {
        const unsigned rps = saddr(rp), cis = saddr(ci), vas = saddr(va), vs = saddr(v), ss = saddr(sub), os = saddr(out), xs = saddr(idx);
        LCQ_LOOP for (int a = LCQ_TID; a < na; a += LCQ_NT) {
            const int r = lds32(xs + 4u * (unsigned)a);
            const int k0 = lds32(rps + 4u * (unsigned)r), k1 = lds32(rps + 4u * (unsigned)r + 4u);
            double s = 0;
            unsigned ca = cis + 2u * (unsigned)k0, xa = vas + 8u * (unsigned)k0;
            LCQ_LOOP for (int k = k0; k < k1; k++) { s += lds64(xa) * lds64(vs + 8u * ldsu16(ca)); ca += 2u; xa += 8u; }
            sts64(os + 8u * (unsigned)a, s - (has_sub ? lds64(ss + 8u * (unsigned)r) : 0.0));
        }
        return;
    }
#endif
    LCQ_LOOP for (int a = LCQ_TID; a < na; a += LCQ_NT) {
        const int r = idx[a];
        double s = 0;
        const int k1 = rp[r + 1];
        LCQ_LOOP for (int k = rp[r]; k < k1; k++) s += va[k] * v[ci[k]];
        out[a] = s - (has_sub ? sub[r] : 0.0);
    }
}

LCQ_DEVN void op_mv_rows(const Op& opr, const int* idx, int na, const double* v, const double* sub, double* out)
{
    const Op op = opr;
    const bool has_sub = sub != nullptr;
    if (!has_sub) sub = v;
    if (op.rp) {
        if (op.smem == 1) csr_mv_rows<true>(op.rp, op.ci, op.va, idx, na, v, sub, has_sub, out);
        else csr_mv_rows<false>(op.rp, op.ci, op.va, idx, na, v, sub, has_sub, out);
    } else {
#ifndef LCQP_HOST_EMU
        LCQ_ASSUME_SHARED(idx); LCQ_ASSUME_SHARED(v); LCQ_ASSUME_SHARED(sub); LCQ_ASSUME_SHARED(out);
#endif
        if (op.trans) {
            // M[r][c] = dense[c*ld + r]: one thread per selected row, eight loads in flight
            LCQ_LOOP for (int a = LCQ_TID; a < na; a += LCQ_NT) {
                const int r = idx[a];
                const double* col = op.dense + r;
                const int cols = op.cols, ld = op.ld;
                double s = 0;
                int c = 0;
#ifndef LCQP_HOST_EMU
                LCQ_LOOP for (; c + 8 <= cols; c += 8) {
                    double mk[8];
#pragma unroll
                    for (int k = 0; k < 8; k++) mk[k] = ldg64v(col + (unsigned)(c + k) * (unsigned)ld);
#pragma unroll
                    for (int k = 0; k < 8; k++) s += mk[k] * v[c + k];
                }
#endif
                LCQ_LOOP for (; c < cols; c++) s += col[(size_t)c * ld] * v[c];
                out[a] = s - (has_sub ? sub[r] : 0.0);
            }
        } else
        LCQ_LOOP for (int a = LCQ_WARP; a < na; a += LCQ_NWARP) {
            const int r = idx[a];
            const double* row = op.dense + (size_t)r * op.ld;
            double s = 0;
            LCQ_LOOP for (int c = LCQ_LANE; c < op.cols; c += LCQ_LANES) s += row[c] * v[c];
            s = warp_sum(s);
            if (LCQ_LANE == 0) out[a] = s - (has_sub ? sub[r] : 0.0);
        }
    }
}

// plain dense mat-vec with a leading dimension: one warp per row, four rows of a warp in flight (the matrix
// usually sits in L2: the loads of the four rows overlap)
LCQ_DEV void mv_dense(const double* __restrict__ M, int rows, int cols, int ld, const double* v, double* out)
{
    LCQ_LOOP for (int r0 = LCQ_WARP; r0 < rows; r0 += 4 * LCQ_NWARP) {
        const int r1 = r0 + LCQ_NWARP, r2 = r0 + 2 * LCQ_NWARP, r3 = r0 + 3 * LCQ_NWARP;
        const double* a0 = M + (size_t)r0 * ld;
        const double* a1 = M + (size_t)(r1 < rows ? r1 : r0) * ld;
        const double* a2 = M + (size_t)(r2 < rows ? r2 : r0) * ld;
        const double* a3 = M + (size_t)(r3 < rows ? r3 : r0) * ld;
        double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
        LCQ_LOOP for (int c = LCQ_LANE; c < cols; c += LCQ_LANES) {
            const double vc = v[c];
            const double m0 = a0[c], m1 = a1[c], m2 = a2[c], m3 = a3[c];
            s0 += m0 * vc; s1 += m1 * vc; s2 += m2 * vc; s3 += m3 * vc;
        }
        s0 = warp_sum(s0); s1 = warp_sum(s1); s2 = warp_sum(s2); s3 = warp_sum(s3);
        if (LCQ_LANE == 0) {
            out[r0] = s0;
            if (r1 < rows) out[r1] = s1;
            if (r2 < rows) out[r2] = s2;
            if (r3 < rows) out[r3] = s3;
        }
    }
}

// ---- packed symmetric matrix (lower triangle, row-major): S(a,b), b <= a, at a(a+1)/2 + b ----------------
// Every row starts at an even offset (16-byte aligned, for 16-byte shared-memory loads): row a begins at
// prow(a) = a(a+1)/2 + (a+1)/2, the rows of odd length carry one pad entry.
inline LCQ_HD size_t prow(int a) { return (size_t)a * (a + 1) / 2 + (size_t)((a + 1) / 2); }
inline LCQ_HD size_t packed_doubles(int order) { return prow(order); }
LCQ_DEV size_t pidx(int a, int b) { return a >= b ? prow(a) + b : prow(b) + a; }

// y = sgn * S v for a symmetric matrix S of order nw, scattered to up to three places:
//     out[a] = y_a;   o2[sidx[a]] = y_a;   o3[sidx[a]] = y_a          (null pointers are skipped)
// ld == 0: S is a packed lower triangle (shared memory);  ld > 0: full storage with leading dimension ld
// (global memory / L2), exactly symmetric.  One non-inlined copy serves every call site.
LCQ_SYM_INL void sym_apply(const double* __restrict__ S, int ld, int nw, const double* v, double sgn,
                        double* out, const int* sidx, double* o2, double* o3)
{
#define LCQ_EMIT(a, val)                                  \
    do {                                                  \
        const double y_ = sgn * (val);                    \
        if (out) out[a] = y_;                             \
        if (sidx) {                                       \
            const int i_ = sidx[a];                       \
            if (o2) o2[i_] = y_;                          \
            if (o3) o3[i_] = y_;                          \
        }                                                 \
    } while (0)
#ifdef LCQP_HOST_EMU
    for (int a = 0; a < nw; a++) {
        double s = 0;
        for (int b = 0; b < nw; b++) s += (ld ? S[(size_t)a * ld + b] : S[pidx(a, b)]) * v[b];
        LCQ_EMIT(a, s);
    }
#else
    // v, out, o2, o3 live in shared memory; sidx may be global (generic access).
    const unsigned vs = saddr(v);
    if (ld) {
        // Full storage in global memory (L2-resident), exactly symmetric, ld even and S 16-byte aligned: row a is
        // read as column a.  A pair of adjacent lanes owns two adjacent columns (one 16-byte load per row) and
        // splits the rows in two halves; sixteen loads (32 values) are in flight per thread.  The two half sums
        // are exchanged by one shuffle; lane h of the pair emits column 2c + h.
        // (the first half has an even number of rows, so that both halves of v are 16-byte aligned)
        const int npair = (nw + 1) >> 1;
        int half = (((nw + 1) >> 1) + 1) & ~1;
        if (half > nw) half = nw;
        LCQ_LOOP for (int t0 = 0; t0 < 2 * npair; t0 += LCQ_NT) {
            const int t = t0 + LCQ_TID;
            const bool act = t < 2 * npair;
            const int c = t >> 1, h = t & 1;
            int b = act ? (h ? half : 0) : 0;
            const int b1 = act ? (h ? nw : half) : 0;
            const double* p = S + (size_t)((unsigned)b * (unsigned)ld + 2u * (unsigned)c);
            unsigned va = vs + 8u * (unsigned)b;
            double x0 = 0, x1 = 0, y0 = 0, y1 = 0;
            const unsigned ldu = (unsigned)ld;
            LCQ_LOOP for (; b + 12 <= b1; b += 12) {
                double m0[12], m1[12];
#pragma unroll
                for (int k = 0; k < 12; k++) { ldg128v(p, m0[k], m1[k]); p += ldu; }
#pragma unroll
                for (int k = 0; k < 12; k += 2) {
                    double v0, v1;
                    lds128v(va + 8u * k, v0, v1);
                    x0 += m0[k] * v0; y0 += m1[k] * v0;
                    x1 += m0[k + 1] * v1; y1 += m1[k + 1] * v1;
                }
                va += 96u;
            }
            LCQ_LOOP for (; b + 4 <= b1; b += 4) {
                double m0[4], m1[4];
#pragma unroll
                for (int k = 0; k < 4; k++) { ldg128v(p, m0[k], m1[k]); p += ldu; }
#pragma unroll
                for (int k = 0; k < 4; k += 2) {
                    double v0, v1;
                    lds128(va + 8u * k, v0, v1);
                    x0 += m0[k] * v0; y0 += m1[k] * v0;
                    x1 += m0[k + 1] * v1; y1 += m1[k + 1] * v1;
                }
                va += 32u;
            }
            LCQ_LOOP for (; b + 2 <= b1; b += 2) {
                double m00, m10, m01, m11, v0, v1;
                ldg128v(p, m00, m10); p += ldu;
                ldg128v(p, m01, m11); p += ldu;
                lds128(va, v0, v1);
                va += 16u;
                x0 += m00 * v0; y0 += m10 * v0;
                x1 += m01 * v1; y1 += m11 * v1;
            }
            LCQ_LOOP for (; b < b1; b++) {
                double m0, m1;
                ldg128(p, m0, m1);
                p += ldu;
                const double v0 = lds64(va);
                va += 8u;
                x0 += m0 * v0; y0 += m1 * v0;
            }
            double sa = x0 + x1, sb = y0 + y1;
            sa += __shfl_xor_sync(0xffffffffu, sa, 1);
            sb += __shfl_xor_sync(0xffffffffu, sb, 1);
            const int a = 2 * c + h;
            if (act && a < nw) LCQ_EMIT(a, h ? sb : sa);
        }
    } else {
        // Packed lower triangle in shared memory, one thread per row: b <= a walks the contiguous row a (the
        // triangular offsets of 16 consecutive rows fall into 16 distinct 8-byte banks), b > a reads S(b,a) at
        // T(b)+a (consecutive across threads).
        const unsigned Ss = saddr(S);
        LCQ_LOOP for (int a = LCQ_TID; a < nw; a += LCQ_NT) {
            double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
            unsigned row = Ss + 8u * (unsigned)prow(a);   // 16-byte aligned
            unsigned va = vs;
            int b = 0;
            LCQ_LOOP for (; b + 4 <= a + 1; b += 4) {
                double r0, r1, r2, r3, v0, v1, v2, v3;
                lds128(row, r0, r1); lds128(row + 16u, r2, r3);
                lds128(va, v0, v1); lds128(va + 16u, v2, v3);
                s0 += r0 * v0; s1 += r1 * v1; s2 += r2 * v2; s3 += r3 * v3;
                row += 32u; va += 32u;
            }
            LCQ_LOOP for (; b <= a; b++) { s0 += lds64(row) * lds64(va); row += 8u; va += 8u; }
            // b = a + 1: the column part, S(b, a) at prow(b) + a; v is read in aligned pairs
            if ((b & 1) && b < nw) { s3 += lds64(Ss + 8u * (unsigned)(prow(b) + a)) * lds64(va); b++; va += 8u; }
            unsigned q = Ss + 8u * (unsigned)(prow(b) + a);
            LCQ_LOOP for (; b + 2 <= nw; b += 2) {     // b even: prow(b+1) - prow(b) = b + 2, prow(b+2) - prow(b) = 2b + 4
                double v0, v1;
                lds128(va, v0, v1);
                s1 += lds64(q) * v0;
                s2 += lds64(q + 8u * (unsigned)(b + 2)) * v1;
                q += 8u * (unsigned)(2 * b + 4); va += 16u;
            }
            if (b < nw) s3 += lds64(q) * lds64(va);
            LCQ_EMIT(a, (s0 + s1) + (s2 + s3));
        }
    }
#endif
#undef LCQ_EMIT
}

LCQ_DEV double limit_scaling(double v)
{
    if (v < kMinScaling) return 1.0;  // osqp scaling.c:22-30
    if (v > kMaxScaling) return kMaxScaling;
    return v;
}

// In-place inversion of the SPD n x n matrix M (row-major, global memory) by Gauss-Jordan without
// pivoting; colbuf/rowbuf are n-vectors.  Returns 0, or 1 on a non-positive pivot.  Rows whose entry in
// the pivot column is zero are left alone, so a block-diagonal matrix keeps its exact zeros.
LCQ_DEVN int spd_invert_inplace(double* M, int n, double* colbuf, double* rowbuf)
{
    LCQ_LOOP for (int k = 0; k < n; k++) {
        LCQ_SYNC();
        const double p = M[(size_t)k * n + k];
        if (!(p > 0.0)) return 1;  // uniform: every thread reads the same value
        LCQ_LOOP for (int j = LCQ_TID; j < n; j += LCQ_NT) {
            colbuf[j] = M[(size_t)j * n + k];
            rowbuf[j] = (j == k ? 1.0 : M[(size_t)k * n + j]) / p;
        }
        LCQ_SYNC();
        LCQ_LOOP for (int i = LCQ_WARP; i < n; i += LCQ_NWARP) {
            double* row = M + (size_t)i * n;
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

// ------------------------------------------------------------------------------------------------
// CSR construction on the device (block-cooperative).  M is rows x cols; trans: M[r][c] = src[c*ld + r].
// Falls back to the dense operator when the matrix is not sparse enough or the pool is exhausted.
// ------------------------------------------------------------------------------------------------
LCQ_DEVN Op build_op(const double* src, int rows, int cols, int ld, int trans, CsrPool& pool, Scalars* sc)
{
    Op op = dense_op(src, rows, cols, ld, trans);
    LCQ_SYNC();
    if (LCQ_TID == 0) {
        int off = pool.used[0];
        if (pool.ibuf && off + rows + 1 <= pool.icap) { sc->ired[0] = off; pool.used[0] = off + rows + 1; }
        else sc->ired[0] = -1;
    }
    LCQ_SYNC();
    const int rpoff = sc->ired[0];
    if (rpoff < 0) return op;
    int* rp = pool.ibuf + rpoff;
    LCQ_LOOP for (int r = LCQ_TID; r < rows; r += LCQ_NT) {
        int c = 0;
        LCQ_LOOP for (int j = 0; j < cols; j++) c += ((trans ? src[(size_t)j * ld + r] : src[(size_t)r * ld + j]) != 0.0);
        rp[r + 1] = c;
    }
    LCQ_SYNC();
    if (LCQ_TID == 0) {
        int tot = 0, nl = 0;
        rp[0] = 0;
        LCQ_LOOP for (int r = 0; r < rows; r++) { const int c = rp[r + 1]; nl += (c > kLongRow); tot += c; rp[r + 1] = tot; }
        const int io = pool.used[0], dof = pool.used[1];
        const int ci_ints = (tot + 1) / 2;   // 16-bit column indices
        const bool sparse = cols < 65536 && (long long)tot * 4 <= (long long)rows * cols && io + ci_ints + nl <= pool.icap && dof + tot <= pool.dcap;
        if (sparse) {
            pool.used[0] = io + ci_ints + nl;
            pool.used[1] = dof + tot;
            sc->ired[1] = io; sc->ired[2] = dof; sc->ired[3] = nl;
            int* lr = pool.ibuf + io + ci_ints;
            int k = 0;
            LCQ_LOOP for (int r = 0; r < rows; r++) if (rp[r + 1] - rp[r] > kLongRow) lr[k++] = r;
        } else {
            sc->ired[1] = -1;
            pool.used[0] = rpoff;  // give the row pointers back
        }
    }
    LCQ_SYNC();
    const int io = sc->ired[1], dof = sc->ired[2], nl = sc->ired[3];
    if (io < 0) { LCQ_SYNC(); return op; }
    unsigned short* ci = reinterpret_cast<unsigned short*>(pool.ibuf + io);
    double* va = pool.dbuf + dof;
    LCQ_LOOP for (int r = LCQ_TID; r < rows; r += LCQ_NT) {
        int k = rp[r];
        LCQ_LOOP for (int j = 0; j < cols; j++) {
            const double v = trans ? src[(size_t)j * ld + r] : src[(size_t)r * ld + j];
            if (v != 0.0) { ci[k] = (unsigned short)j; va[k] = v; k++; }
        }
    }
    LCQ_SYNC();
    op.rp = rp; op.ci = ci; op.va = va; op.lrows = pool.ibuf + io + (rp[rows] + 1) / 2; op.nlong = nl;
    return op;
}

// ------------------------------------------------------------------------------------------------
// prepare: Ruiz equilibration (osqp scaling.c:44-156 without the q-dependent cost scale, c = 1), row types,
// Hinv, Minv, AH, T, the static equality block.  Scratch: vectors n (v1, v2), m (e1, e2).
// ------------------------------------------------------------------------------------------------
LCQ_DEVN void prepare_scale(const Dims& d, const Inst& in, Mats& mt, double* v1, double* e1)
{
    const int n = d.n, m = d.m, mA = d.mA, nC = d.nC, nComp = d.nComp;
    // P = Q, A = [A; L; R; I]
    LCQ_LOOP for (int e = LCQ_TID; e < n * n; e += LCQ_NT) mt.P[e] = in.Q[e];
    LCQ_LOOP for (int e = LCQ_TID; e < m * n; e += LCQ_NT) {
        const int i = e / n, j = e - i * n;
        double v;
        if (i < nC) v = in.A[e];
        else if (i < nC + nComp) v = in.L[(size_t)(i - nC) * n + j];
        else if (i < mA) v = in.R[(size_t)(i - nC - nComp) * n + j];
        else v = (i - mA == j) ? 1.0 : 0.0;
        mt.A[e] = v;
    }
    LCQ_LOOP for (int j = LCQ_TID; j < n; j += LCQ_NT) mt.D[j] = 1.0;
    LCQ_LOOP for (int i = LCQ_TID; i < m; i += LCQ_NT) mt.E[i] = 1.0;
    double* Dt = v1;
    double* Et = e1;
    LCQ_LOOP for (int it = 0; it < kScalingIters; it++) {
        LCQ_SYNC();
        LCQ_LOOP for (int j = LCQ_TID; j < n; j += LCQ_NT) {
            double v = 0;
            LCQ_LOOP for (int i = 0; i < n; i++) v = fmax(v, fabs(mt.P[(size_t)i * n + j]));
            LCQ_LOOP for (int i = 0; i < m; i++) v = fmax(v, fabs(mt.A[(size_t)i * n + j]));
            Dt[j] = 1.0 / sqrt(limit_scaling(v));
        }
        LCQ_LOOP for (int i = LCQ_WARP; i < m; i += LCQ_NWARP) {
            double v = 0;
            LCQ_LOOP for (int j = LCQ_LANE; j < n; j += LCQ_LANES) v = fmax(v, fabs(mt.A[(size_t)i * n + j]));
            v = warp_max(v);
            if (LCQ_LANE == 0) Et[i] = 1.0 / sqrt(limit_scaling(v));
        }
        LCQ_SYNC();
        LCQ_LOOP for (int e = LCQ_TID; e < n * n; e += LCQ_NT) {
            const int i = e / n, j = e - i * n;
            mt.P[e] *= Dt[i] * Dt[j];
        }
        LCQ_LOOP for (int e = LCQ_TID; e < m * n; e += LCQ_NT) {
            const int i = e / n, j = e - i * n;
            mt.A[e] *= Et[i] * Dt[j];
        }
        LCQ_LOOP for (int j = LCQ_TID; j < n; j += LCQ_NT) mt.D[j] *= Dt[j];
        LCQ_LOOP for (int i = LCQ_TID; i < m; i += LCQ_NT) mt.E[i] *= Et[i];
    }
    LCQ_SYNC();
    LCQ_LOOP for (int e = LCQ_TID; e < m * n; e += LCQ_NT) {
        const int j = e / m, i = e - j * m;
        mt.At[e] = mt.A[(size_t)i * n + j];
    }
    LCQ_SYNC();
}

// Bounds of A_full = [A; L; R] (+ box rows) as setConstraints / setComplementarityBounds build them
// (/root/reference/src/LCQProblem.cpp:584-608, 745-782), scaled by E, with OSQP's row types
// (auxil.c:76-98).  Returns bit0: some l > u, bit1: a complementarity lower bound is -inf (:747,:767).
// store_lu = false: l / u already hold these values (one copy shared by the groups of the CTA, written before the
// instance loop); only the row types and the flags are produced.
LCQ_DEVN int set_bounds(const Dims& d, const Inst& in, const double* E, double* l, double* u, signed char* ctype, Scalars* sc,
                        bool store_lu = true)
{
    int bad = 0;
    LCQ_LOOP for (int i = LCQ_TID; i < d.m; i += LCQ_NT) {
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
        const double Ei = E[i];
        const double li = linf ? -INFINITY : lo * Ei, ui = uinf ? INFINITY : up * Ei;
        if (store_lu) { l[i] = li; u[i] = ui; }
        signed char t = 0;
        if (linf && uinf) t = -1;
        else if (!linf && !uinf && ui - li < kRhoTol) t = 1;
        ctype[i] = t;
    }
    (void)sc;
    return block_or(bad, sc);
}

LCQ_DEV double rho_of(signed char t, double rho)
{
    return t < 0 ? kRhoMin : (t >= 1 ? kRhoEqOverIneq * rho : rho);  // auxil.c:76-98
}

// Everything that follows the scaling.  ctype (m, row types of this instance / of instance 0 of a sharing
// batch) is copied into mt.ctype, where equality rows found linearly dependent are re-typed 2.
// Scratch: v1, v2 (n), e1, e2, e3 (m).
LCQ_DEVN int prepare_factor(const Dims& d, Mats& mt, const signed char* ctype, const lcqp_cuda_options& o,
                            double* v1, double* v2, double* e1, double* e2, double* e3, Scalars* sc)
{
    const int n = d.n, m = d.m;
    const double delta = o.qp_delta;
    // Hinv
    LCQ_LOOP for (int e = LCQ_TID; e < n * n; e += LCQ_NT) {
        const int i = e / n, j = e - i * n;
        mt.Hinv[e] = mt.P[e] + (i == j ? delta : 0.0);
    }
    if (spd_invert_inplace(mt.Hinv, n, v1, v2)) return 1;
    // Minv
    LCQ_LOOP for (int i = LCQ_TID; i < m; i += LCQ_NT) { e1[i] = rho_of(ctype[i], o.qp_rho); mt.ctype[i] = ctype[i]; }
    LCQ_SYNC();
    LCQ_LOOP for (int e = LCQ_TID; e < n * n; e += LCQ_NT) {
        const int i = e / n, j = e - i * n;
        mt.Minv[e] = mt.P[e] + (i == j ? o.qp_sigma : 0.0);
    }
    LCQ_SYNC();
    if (mt.oA.rp) {
        // A' R A accumulated row by row of A (rows are short); one thread per row pair product would race, so
        // rows are processed one after the other and the (a, b) pairs of a row in parallel
        const Op& oa = mt.oA;
        LCQ_LOOP for (int r = 0; r < m; r++) {
            const int k0 = oa.rp[r], len = oa.rp[r + 1] - k0;
            LCQ_LOOP for (int e = LCQ_TID; e < len * len; e += LCQ_NT) {
                const int a = e / len, b = e - a * len;
                mt.Minv[(size_t)oa.ci[k0 + a] * n + oa.ci[k0 + b]] += e1[r] * oa.va[k0 + a] * oa.va[k0 + b];
            }
            LCQ_SYNC();
        }
    } else {
        LCQ_LOOP for (int e = LCQ_TID; e < n * n; e += LCQ_NT) {
            const int i = e / n, j = e - i * n;
            if (j > i) continue;
            double s = 0;
            LCQ_LOOP for (int r = 0; r < m; r++) s += e1[r] * mt.A[(size_t)r * n + i] * mt.A[(size_t)r * n + j];
            mt.Minv[e] += s;
            if (j != i) mt.Minv[(size_t)j * n + i] += s;
        }
    }
    if (spd_invert_inplace(mt.Minv, n, v1, v2)) return 1;
    // AH = A Hinv, G = AH A' (into T); through the CSR rows of A when they exist (mt.oA is set by the caller)
    const Op& oA = mt.oA;
    LCQ_LOOP for (int e = LCQ_TID; e < m * n; e += LCQ_NT) {
        const int i = e / n, j = e - i * n;
        double s = 0;
        if (oA.rp) {
            LCQ_LOOP for (int k = oA.rp[i]; k < oA.rp[i + 1]; k++) s += oA.va[k] * mt.Hinv[(size_t)oA.ci[k] * n + j];
        } else {
            const double* ar = mt.A + (size_t)i * n;
            LCQ_LOOP for (int k = 0; k < n; k++) s += ar[k] * mt.Hinv[(size_t)k * n + j];
        }
        mt.AH[e] = s;
    }
    LCQ_SYNC();
    if (oA.rp) {
        LCQ_LOOP for (int e = LCQ_TID; e < m * m; e += LCQ_NT) {
            const int i = e / m, r = e - i * m;
            if (r > i) continue;
            const double* ah = mt.AH + (size_t)i * n;
            double s = 0;
            LCQ_LOOP for (int k = oA.rp[r]; k < oA.rp[r + 1]; k++) s += ah[oA.ci[k]] * oA.va[k];
            mt.T[(size_t)i * m + r] = s;
            mt.T[(size_t)r * m + i] = s;
        }
    } else {
        LCQ_LOOP for (int e = LCQ_WARP; e < m * m; e += LCQ_NWARP) {
            const int i = e / m, r = e - i * m;
            if (r > i) continue;
            const double* ah = mt.AH + (size_t)i * n;
            const double* ar = mt.A + (size_t)r * n;
            double s = 0;
            LCQ_LOOP for (int k = LCQ_LANE; k < n; k += LCQ_LANES) s += ah[k] * ar[k];
            s = warp_sum(s);
            if (LCQ_LANE == 0) { mt.T[(size_t)i * m + r] = s; mt.T[(size_t)r * m + i] = s; }
        }
    }
    LCQ_SYNC();
    // static equality block: greedy, in row order; SEinv by bordering (full storage, ld = mEcap).  A row whose
    // Schur pivot is at the level of the regularisation is a combination of the rows before it: it is never
    // put into a working set (type 2); its bounds are still checked by the KKT test.
    const int mEcap = d.ldE;
    int mE = 0;
    double* Si = mt.SEinv;
    double* sv = e1;   // G[j, E]
    double* uu = e2;   // SEinv sv
    LCQ_LOOP for (int j = 0; j < m; j++) {
        if (ctype[j] != 1) continue;
        const double* Gj = mt.T + (size_t)j * m;
        LCQ_LOOP for (int a = LCQ_TID; a < mE; a += LCQ_NT) sv[a] = Gj[mt.eidx[a]];
        LCQ_SYNC();
        mv_dense(Si, mE, mE, mEcap, sv, uu);
        LCQ_SYNC();
        double p1 = 0, p2 = 0;
        LCQ_LOOP for (int a = LCQ_TID; a < mE; a += LCQ_NT) { p1 += sv[a] * uu[a]; p2 += uu[a] * uu[a]; }
        block_sum2(p1, p2, sc);
        const double kappa = Gj[j] + delta - p1;
        if (mE >= mEcap || !(kappa > 10.0 * delta * (1.0 + p2))) {
            if (LCQ_TID == 0) mt.ctype[j] = 2;
            LCQ_SYNC();
            continue;
        }
        const double ik = 1.0 / kappa;
        LCQ_LOOP for (int e = LCQ_TID; e < mE * mE; e += LCQ_NT) {
            const int a = e / mE, b = e - a * mE;
            Si[(size_t)a * mEcap + b] += uu[a] * uu[b] * ik;
        }
        LCQ_LOOP for (int a = LCQ_TID; a < mE; a += LCQ_NT) {
            Si[(size_t)a * mEcap + mE] = -uu[a] * ik;
            Si[(size_t)mE * mEcap + a] = -uu[a] * ik;
        }
        if (LCQ_TID == 0) { Si[(size_t)mE * mEcap + mE] = ik; mt.eidx[mE] = j; }
        mE++;
        LCQ_SYNC();
    }
    // T <- G - G[:,E] SEinv G[E,:]   (only entries with both rows outside E are used afterwards)
    if (mE > 0) {
        // X = G[:,E] SEinv, row by row, kept in e3; then T[i,:] -= X[i,:] G[E,:]
        LCQ_LOOP for (int i = 0; i < m; i++) {
            if (mt.ctype[i] == 1) continue;
            double* Ti = mt.T + (size_t)i * m;
            LCQ_LOOP for (int a = LCQ_TID; a < mE; a += LCQ_NT) sv[a] = Ti[mt.eidx[a]];
            LCQ_SYNC();
            int any = 0;
            LCQ_LOOP for (int a = LCQ_TID; a < mE; a += LCQ_NT) any |= (sv[a] != 0.0);
            any = block_or(any, sc);
            if (!any) continue;  // row decoupled from the equality block
            mv_dense(Si, mE, mE, mEcap, sv, e3);
            LCQ_SYNC();
            LCQ_LOOP for (int r = LCQ_TID; r < m; r += LCQ_NT) {
                if (mt.ctype[r] == 1) continue;
                double s = 0;
                LCQ_LOOP for (int a = 0; a < mE; a++) s += e3[a] * mt.T[(size_t)mt.eidx[a] * m + r];
                Ti[r] -= s;
            }
            LCQ_SYNC();
        }
    }
    LCQ_LOOP for (int e = LCQ_TID; e < mE * n; e += LCQ_NT) {
        const int a = e / n, j = e - a * n;
        mt.AHE[e] = mt.AH[(size_t)mt.eidx[a] * n + j];
    }
    if (LCQ_TID == 0) { mt.mE = mE; mt.status = 0; }
    LCQ_SYNC();
    (void)uu;
    return 0;
}

// dense operators over the prepared matrices (per-instance preparation; one thread writes) ...
LCQ_DEV void mats_dense_ops(const Dims& d, Mats& mt)
{
    const int n = d.n, m = d.m;
    // every dense operator is applied in the form out[r] = sum_c M[c*ld + r] v[c] (coalesced, no cross-lane
    // reduction): the symmetric ones as they are (P exactly symmetric, the inverses to rounding), A through At
    mt.oP = dense_op(mt.P, n, n, n, 1);
    mt.oA = dense_op(mt.At, m, n, m, 1);
    mt.oAt = dense_op(mt.A, n, m, n, 1);
    mt.oHinv = dense_op(mt.Hinv, n, n, n, 1);
    mt.oMinv = dense_op(mt.Minv, n, n, n, 1);
}

// the operators of the static equality block need its order (known after prepare_factor)
LCQ_DEV void mats_dense_ops_post(const Dims& d, Mats& mt)
{
    mt.oAHE = dense_op(mt.AHE, mt.mE, d.n, d.n, 0);
    mt.oAHtE = dense_op(mt.AHE, d.n, mt.mE, d.n, 1);
}

// ... or CSR copies where the matrix is sparse (shared preparation, done once per batch): the scaled
// problem matrices before the factorisation (which uses oA), the prepared ones after it.  `mt` must be
// block-shared or thread-local-identical: every thread receives the same Op values.
LCQ_DEVN void mats_build_ops_pre(const Dims& d, Mats& mt, CsrPool& pool, Scalars* sc)
{
    const int n = d.n, m = d.m;
    Op a = build_op(mt.P, n, n, n, 0, pool, sc);
    Op b = build_op(mt.A, m, n, n, 0, pool, sc);
    const Op c = build_op(mt.A, n, m, n, 1, pool, sc);
    if (!a.rp) a = dense_op(mt.P, n, n, n, 1);     // dense fall-back: see mats_dense_ops
    if (!b.rp) b = dense_op(mt.At, m, n, m, 1);
    if (LCQ_TID == 0) { mt.oP = a; mt.oA = b; mt.oAt = c; }
    LCQ_SYNC();
}

LCQ_DEVN void mats_build_ops_post(const Dims& d, Mats& mt, CsrPool& pool, Scalars* sc)
{
    const int n = d.n, m = d.m;
    Op a = build_op(mt.Hinv, n, n, n, 0, pool, sc);
    const Op b = build_op(mt.AHE, mt.mE, n, n, 0, pool, sc);
    const Op c = build_op(mt.AHE, n, mt.mE, n, 1, pool, sc);
    Op e = build_op(mt.Minv, n, n, n, 0, pool, sc);
    if (!a.rp) a = dense_op(mt.Hinv, n, n, n, 1);
    if (!e.rp) e = dense_op(mt.Minv, n, n, n, 1);
    (void)m;
    if (LCQ_TID == 0) { mt.oHinv = a; mt.oAHE = b; mt.oAHtE = c; mt.oMinv = e; }
    LCQ_SYNC();
}

LCQ_DEV void raw_dense_ops(const Dims& d, const Inst& in, RawOps& ro, unsigned keep_mask)
{
    const int n = d.n;
    if (!(keep_mask & (1u << LCQP_Q))) ro.Q = dense_op(in.Q, n, n, n, 0);
    if (!(keep_mask & (1u << LCQP_L))) { ro.L = dense_op(in.L, d.nComp, n, n, 0); ro.Lt = dense_op(in.L, n, d.nComp, n, 1); }
    if (!(keep_mask & (1u << LCQP_R))) { ro.R = dense_op(in.R, d.nComp, n, n, 0); ro.Rt = dense_op(in.R, n, d.nComp, n, 1); }
    if (!(keep_mask & (1u << LCQP_A))) ro.At = dense_op(in.A, n, d.nC, n, 1);
}

// one thread writes `ro` (block-shared)
LCQ_DEVN void raw_build_ops(const Dims& d, const Inst& in, RawOps& ro, unsigned shared_mask, CsrPool& pool, Scalars* sc)
{
    const int n = d.n;
    if (LCQ_TID == 0) raw_dense_ops(d, in, ro, 0u);
    LCQ_SYNC();
    if (shared_mask & (1u << LCQP_Q)) { const Op a = build_op(in.Q, n, n, n, 0, pool, sc); if (LCQ_TID == 0) ro.Q = a; }
    if (shared_mask & (1u << LCQP_L)) {
        const Op a = build_op(in.L, d.nComp, n, n, 0, pool, sc);
        const Op b = build_op(in.L, n, d.nComp, n, 1, pool, sc);
        if (LCQ_TID == 0) { ro.L = a; ro.Lt = b; }
    }
    if (shared_mask & (1u << LCQP_R)) {
        const Op a = build_op(in.R, d.nComp, n, n, 0, pool, sc);
        const Op b = build_op(in.R, n, d.nComp, n, 1, pool, sc);
        if (LCQ_TID == 0) { ro.R = a; ro.Rt = b; }
    }
    if ((shared_mask & (1u << LCQP_A)) && d.nC > 0) { const Op a = build_op(in.A, n, d.nC, n, 1, pool, sc); if (LCQ_TID == 0) ro.At = a; }
    LCQ_SYNC();
}

// ------------------------------------------------------------------------------------------------
// Shared-memory cache of the batch-shared operators.  A persistent CTA copies the CSR arrays of the
// operators it applies in every pass (and the packed inverse of the static equality block) from L2 into its
// own shared memory once, at kernel start; every later application is then free of L2 round trips.
// ------------------------------------------------------------------------------------------------
inline LCQ_HD size_t op_cache_bytes(int rows, int nnz, int nlong)
{
    size_t b = ((size_t)nnz * sizeof(double) + 15) / 16 * 16;
    b += (((size_t)rows + 1 + nlong) * sizeof(int) + 15) / 16 * 16;
    b += ((size_t)nnz * sizeof(unsigned short) + 15) / 16 * 16;
    return b;
}

LCQ_DEV size_t op_cache_bytes(const Op& op) { return op.rp ? op_cache_bytes(op.rows, op.rp[op.rows], op.nlong) : 0; }

// Copy a CSR operator into [*cur, end) if it fits (block-cooperative; every thread gets the same result).
LCQ_DEVN void cache_op(Op& op, unsigned char*& cur, unsigned char* end)
{
    if (!op.rp) return;
    const int rows = op.rows, nnz = op.rp[rows], nl = op.nlong;
    const size_t need = op_cache_bytes(rows, nnz, nl);
    if (cur + need > end) return;
    double* va = reinterpret_cast<double*>(cur);
    int* rp = reinterpret_cast<int*>(cur + ((size_t)nnz * sizeof(double) + 15) / 16 * 16);
    int* lr = rp + rows + 1;
    unsigned short* ci = reinterpret_cast<unsigned short*>(reinterpret_cast<unsigned char*>(rp) + (((size_t)rows + 1 + nl) * sizeof(int) + 15) / 16 * 16);
    LCQ_LOOP for (int k = LCQ_TID; k < nnz; k += LCQ_NT) { va[k] = op.va[k]; ci[k] = op.ci[k]; }
    LCQ_LOOP for (int r = LCQ_TID; r <= rows; r += LCQ_NT) rp[r] = op.rp[r];
    LCQ_LOOP for (int a = LCQ_TID; a < nl; a += LCQ_NT) lr[a] = op.lrows[a];
    op.va = va; op.rp = rp; op.ci = ci; op.lrows = lr; op.smem = 1;
    cur += need;
}

// `mt` / `ro` are this CTA's block-shared copies; one thread writes them.  what: bit0 packed SEinv,
// bit1 the operators of the inner passes, bit2 the operators of the outer loop.
LCQ_DEVN void cache_shared_operators(const Dims& d, Mats& mt, RawOps& ro, unsigned char* base, size_t bytes, int what)
{
    unsigned char* cur = base;
    unsigned char* end = base + bytes;
    const int mE = mt.mE;
    LCQ_SYNC();
    const double* sep = nullptr;
    if (mE > 0 && (what & 1)) {
        const size_t need = (packed_doubles(mE) * sizeof(double) + 15) / 16 * 16;
        if (cur + need <= end) {
            double* P = reinterpret_cast<double*>(cur);
            LCQ_LOOP for (int e = LCQ_TID; e < mE * mE; e += LCQ_NT) {
                const int a = e / mE, b = e - a * mE;
                if (b <= a) P[prow(a) + b] = mt.SEinv[(size_t)a * d.ldE + b];
            }
            sep = P;
            cur += need;
        }
    }
    Op oA = mt.oA, oAt = mt.oAt, oAHE = mt.oAHE, oAHtE = mt.oAHtE, oHinv = mt.oHinv, oP = mt.oP;
    int* eidx_s = nullptr;
    if ((what & 2) && mE > 0) {
        // the index list of the static equality block is read in every pass as well
        const size_t need = ((size_t)mE * sizeof(int) + 15) / 16 * 16;
        if (cur + need <= end) {
            eidx_s = reinterpret_cast<int*>(cur);
            LCQ_LOOP for (int a = LCQ_TID; a < mE; a += LCQ_NT) eidx_s[a] = mt.eidx[a];
            cur += need;
        }
    }
    if (what & 2) {
        cache_op(oA, cur, end); cache_op(oAt, cur, end); cache_op(oAHE, cur, end); cache_op(oAHtE, cur, end);
        cache_op(oHinv, cur, end); cache_op(oP, cur, end);
    }
    Op rL = ro.L, rR = ro.R, rLt = ro.Lt, rRt = ro.Rt, rQ = ro.Q, rAt = ro.At;
    if (what & 4) {
        cache_op(rL, cur, end); cache_op(rR, cur, end); cache_op(rLt, cur, end); cache_op(rRt, cur, end);
        cache_op(rQ, cur, end); cache_op(rAt, cur, end);
    }
    LCQ_SYNC();
    if (LCQ_TID == 0) {
        mt.SEinvP = sep;
        if (eidx_s) mt.eidx = eidx_s;
        mt.oA = oA; mt.oAt = oAt; mt.oAHE = oAHE; mt.oAHtE = oAHtE; mt.oHinv = oHinv; mt.oP = oP;
        ro.L = rL; ro.R = rR; ro.Lt = rLt; ro.Rt = rRt; ro.Q = rQ; ro.At = rAt;
    }
    LCQ_SYNC();
}

// bytes the cache would take (thread 0 of the prepare step calls this)
LCQ_DEV void cache_requirements(Mats& mt, const RawOps& ro)
{
    mt.cache_bytes_se = (int)((packed_doubles(mt.mE) * sizeof(double) + 15) / 16 * 16);
    mt.cache_bytes_hot = (int)(((size_t)mt.mE * sizeof(int) + 15) / 16 * 16 + op_cache_bytes(mt.oA) + op_cache_bytes(mt.oAt) + op_cache_bytes(mt.oAHE) + op_cache_bytes(mt.oAHtE) +
                               op_cache_bytes(mt.oHinv) + op_cache_bytes(mt.oP));
    mt.cache_bytes_raw = (int)(op_cache_bytes(ro.L) + op_cache_bytes(ro.R) + op_cache_bytes(ro.Lt) + op_cache_bytes(ro.Rt) +
                               op_cache_bytes(ro.Q) + op_cache_bytes(ro.At));
}

// ------------------------------------------------------------------------------------------------
// QP solver state machine
// ------------------------------------------------------------------------------------------------
// State of the QP solver of one group.  It lives in SHARED memory (one per group): the non-inlined device
// functions receive a reference to it, and a per-thread copy would sit in local memory, i.e. behind an L2 round
// trip at every function entry (the L1 left beside 220 KB of shared memory does not hold the stacks of 512 threads).
// All threads read it; thread 0 alone writes it, always with a group barrier between a write and the reads on
// either side of it.
struct QP {
    SPtr<const Dims> d;
    SPtr<const Mats> mt;
    SPtr<const Work> w;
    SPtr<const lcqp_cuda_options> o;
    int nw;           // rows in the inequality working set = order of Tinv (idx[0..nw))
    int have_W;
    int tinv_valid;   // Tinv matches idx/W
    int pad;
    long long n_admm, n_pass, n_changes;
};

// One ADMM iteration (osqp auxil.c:161-225, condensed KKT solve).
LCQ_DEVN void admm_iter(QP& s)
{
    const int n = s.d->n, m = s.d->m;
    const Work& w = *s.w;
    const Mats& mt = *s.mt;
    const double alpha = s.o->qp_alpha, sigma = s.o->qp_sigma, rho = s.o->qp_rho;
    LCQ_LOOP for (int i = LCQ_TID; i < m; i += LCQ_NT) w.w[i] = rho_of(w.ctype[i], rho) * w.z[i] - w.y[i];
    LCQ_LOOP for (int j = LCQ_TID; j < n; j += LCQ_NT) w.u[j] = sigma * w.x[j] - w.q[j];
    LCQ_SYNC();
    op_mv_s(mt.oAt, w.w, w.u, 1.0, w.t);        // rhs = sigma x - q + A'w
    LCQ_SYNC();
    op_mv_s(mt.oMinv, w.t, nullptr, 1.0, w.dx); // xt
    LCQ_SYNC();
    op_mv_s(mt.oA, w.dx, nullptr, 1.0, w.zp);   // zt = A xt
    LCQ_LOOP for (int j = LCQ_TID; j < n; j += LCQ_NT) w.x[j] = alpha * w.dx[j] + (1.0 - alpha) * w.x[j];
    LCQ_SYNC();
    LCQ_LOOP for (int i = LCQ_TID; i < m; i += LCQ_NT) {
        const double r = rho_of(w.ctype[i], rho);
        const double v = alpha * w.zp[i] + (1.0 - alpha) * w.z[i];
        double zn = v + w.y[i] / r;
        zn = fmin(fmax(zn, w.l[i]), w.ub[i]);
        w.y[i] += r * (v - zn);
        w.z[i] = zn;
    }
    LCQ_SYNC();
}

// Active-set guess from the ADMM iterate (osqp polish.c:33-49; equality rows always active).
LCQ_DEVN void guess_working_set(QP& s, signed char* W)
{
    const Work& w = *s.w;
    LCQ_LOOP for (int i = LCQ_TID; i < s.d->m; i += LCQ_NT) {
        signed char v = 0;
        const signed char t = w.ctype[i];
        if (t == 1) v = 1;
        else if (t != 0) v = 0;
        else if (w.z[i] - w.l[i] < -w.y[i]) v = 1;
        else if (w.ub[i] - w.z[i] < w.y[i]) v = 2;
        W[i] = v;
    }
    LCQ_SYNC();
}

// S += c * u u' on the leading nw x nw block of a full-storage matrix in global memory (leading dimension ld).
// (u_a u_b) c is commutative in a, b: S stays exactly symmetric.  Eight independent load -> update -> store
// chains per thread, so that eight L2 round trips overlap instead of one.
LCQ_R1_INL void rank1_update_full(double* __restrict__ S, int ld, int nw, const double* u, double c)
{
#ifndef LCQP_HOST_EMU
    // ld even, S 16-byte aligned, u in shared memory (16-byte aligned).  The nw x ceil(nw/2) pairs of adjacent
    // columns are dealt out round-robin (element e -> row e / npair, pair e % npair: every thread gets the same
    // share); BATCH 16-byte loads are issued (volatile: kept together by ptxas) before the first store.
    constexpr int BATCH = 8;
    const unsigned us = saddr(u);
    const int npair = (nw + 1) >> 1, total = nw * npair;
    const int dq = LCQ_NT / npair, dr = LCQ_NT - dq * npair;   // (row, pair) advance per LCQ_NT elements
    int row = LCQ_TID / npair, cp = LCQ_TID - row * npair;
    LCQ_LOOP for (int e0 = LCQ_TID; e0 < total; e0 += BATCH * LCQ_NT) {
        double m0[BATCH], m1[BATCH];
        int key[BATCH];
#pragma unroll
        for (int k = 0; k < BATCH; k++) {
            key[k] = (e0 + k * LCQ_NT < total) ? ((row << 16) | cp) : -1;
            ldg128v(S + (key[k] >= 0 ? (unsigned)(row * ld + 2 * cp) : 0u), m0[k], m1[k]);   // (unconditional: keeps m0/m1 in registers)
            cp += dr; row += dq;
            if (cp >= npair) { cp -= npair; row++; }
        }
#pragma unroll
        for (int k = 0; k < BATCH; k++) {
            if (key[k] >= 0) {
                const int r = key[k] >> 16, c2 = key[k] & 0xffff;
                const double ua = lds64(us + 8u * (unsigned)r);
                double u0, u1;
                lds128(us + 16u * (unsigned)c2, u0, u1);
                double* q = S + (unsigned)(r * ld + 2 * c2);
                // (column nw of an odd block is not part of the matrix: the caller may be writing it)
                if (2 * c2 + 1 < nw) stg128(q, m0[k] + (ua * u0) * c, m1[k] + (ua * u1) * c);
                else stg64(q, m0[k] + (ua * u0) * c);
            }
        }
    }
#else
    for (int a = 0; a < nw; a++)
        for (int b = 0; b < nw; b++) S[(size_t)a * ld + b] += (u[a] * u[b]) * c;
#endif
}

// ---- explicit inverse of T[W,W] + delta I (packed), maintained by bordering ------------------------
// Append inequality row j to the working set (position nw).  A row that is numerically a combination of
// the rows already active -- Schur pivot kappa at the level of the regularisation, kappa <= 10 delta
// (1 + u'u) -- is NOT appended (the working set is kept linearly independent, as the active-set theory
// requires); returns 1 in that case, 0 otherwise.  Scratch: dI, lI.
LCQ_DEVN int tinv_append(QP& s, int j)
{
    const Work& w = *s.w;
    const int nw = s.nw, m = s.d->m;
    double* Si = w.Tinv;
    const double* Tj = s.mt->T + (size_t)j * m;
    LCQ_LOOP for (int a = LCQ_TID; a < nw; a += LCQ_NT) w.dI[a] = Tj[w.idx[a]];
    LCQ_SYNC();
    sym_apply(Si, w.tld, nw, w.dI, 1.0, w.lI, nullptr, nullptr, nullptr);
    LCQ_SYNC();
    double p1 = 0, p2 = 0;
    LCQ_LOOP for (int a = LCQ_TID; a < nw; a += LCQ_NT) { const double v = w.lI[a]; p1 += w.dI[a] * v; p2 += v * v; }
    fast_sum2(p1, p2, w.sc);
    const double kappa = Tj[j] + s.o->qp_delta - p1;
    if (!(kappa > 10.0 * s.o->qp_delta * (1.0 + p2))) return 1;
    const double ik = 1.0 / kappa;
    if (w.tld) {
        const int ld = w.tld;
        rank1_update_full(Si, ld, nw, w.lI, ik);   // S += ik * lI lI'
        double* row = Si + (size_t)nw * ld;
        LCQ_LOOP for (int a = LCQ_TID; a < nw; a += LCQ_NT) { const double v = -w.lI[a] * ik; row[a] = v; Si[(size_t)a * ld + nw] = v; }
        if (LCQ_TID == 0) { row[nw] = ik; w.idx[nw] = j; }
    } else {
        LCQ_LOOP for (int a = LCQ_WARP; a < nw; a += LCQ_NWARP) {
            const double ua = w.lI[a] * ik;
            double* row = Si + prow(a);
            LCQ_LOOP for (int b = LCQ_LANE; b <= a; b += LCQ_LANES) row[b] += ua * w.lI[b];
        }
        double* row = Si + prow(nw);
        LCQ_LOOP for (int a = LCQ_TID; a < nw; a += LCQ_NT) row[a] = -w.lI[a] * ik;
        if (LCQ_TID == 0) { row[nw] = ik; w.idx[nw] = j; }
    }
    if (LCQ_TID == 0) s.nw = nw + 1;
    LCQ_SYNC();
    return 0;
}

// remove position p from the working set (the last position is moved into p).  Scratch: dI, lI.
LCQ_DEVN void tinv_remove(QP& s, int p)
{
    const Work& w = *s.w;
    const int nw = s.nw, last = nw - 1;
    double* Si = w.Tinv;
    const int ld = w.tld;
    LCQ_LOOP for (int b = LCQ_TID; b < nw; b += LCQ_NT) w.dI[b] = ld ? Si[(size_t)b * ld + p] : Si[pidx(b, p)];
    LCQ_SYNC();
    const double ic = 1.0 / w.dI[p];
    if (ld) rank1_update_full(Si, ld, nw, w.dI, -ic);   // S -= ic * dI dI'
    else
        LCQ_LOOP for (int a = LCQ_WARP; a < nw; a += LCQ_NWARP) {
            const double ca = w.dI[a] * ic;
            double* row = Si + prow(a);
            LCQ_LOOP for (int b = LCQ_LANE; b <= a; b += LCQ_LANES) row[b] -= ca * w.dI[b];
        }
    LCQ_SYNC();
    if (p != last) {
        // move row/column `last` into position p
        LCQ_LOOP for (int b = LCQ_TID; b < nw; b += LCQ_NT) w.lI[b] = ld ? Si[(size_t)last * ld + b] : Si[pidx(last, b)];
        LCQ_SYNC();
        LCQ_LOOP for (int b = LCQ_TID; b < last; b += LCQ_NT) {
            const double v = (b == p) ? w.lI[last] : w.lI[b];
            if (ld) { Si[(size_t)p * ld + b] = v; Si[(size_t)b * ld + p] = v; }
            else Si[pidx(p, b)] = v;
        }
        if (LCQ_TID == 0) w.idx[p] = w.idx[last];
    }
    if (LCQ_TID == 0) s.nw = last;
    LCQ_SYNC();
}

// Rebuild idx/Tinv from W by successive bordering; inequality rows found linearly dependent on the rows
// before them (or beyond the capacity) are taken out of W.  Equality rows: W follows the static block.
LCQ_DEVN void tinv_build(QP& s, signed char* W)
{
    const int m = s.d->m;
    LCQ_SYNC();
    if (LCQ_TID == 0) { s.nw = 0; s.tinv_valid = 0; }
    LCQ_LOOP for (int i = LCQ_TID; i < m; i += LCQ_NT)
        if (s.w->ctype[i] >= 1) W[i] = (s.w->ctype[i] == 1) ? 1 : 0;
    LCQ_SYNC();
    LCQ_LOOP for (int i = 0; i < m; i++) {  // W is in shared memory: uniform branches
        if (!W[i] || s.w->ctype[i] != 0) continue;
        if (s.nw >= s.d->cap || tinv_append(s, i)) {
            LCQ_SYNC();
            if (LCQ_TID == 0) W[i] = 0;
            LCQ_SYNC();
        }
    }
    LCQ_SYNC();
    if (LCQ_TID == 0) s.tinv_valid = 1;
    LCQ_SYNC();
}

// Solve the regularised KKT system of the working set,
//     [P + dI, Aw'; Aw, -dI] [dx; dlam] = [r1; r2],
// r2 / dlam full-length (m) vectors of which only the rows in W count / are written.
// Block elimination: Hinv, then the static equality block (SEinv), then the inequality rows (Tinv).
// In: w.r1, w.r2.  Out: w.dx, w.dlam.  Scratch: u, t, cE, vE, dI, yf.
// Invariants kept by the callers: yf is zero on entry (and on exit); dlam is zero outside the working set.
LCQ_DEVN void kkt_solve(QP& s)
{
    const int n = s.d->n, nw = s.nw;
    const Work& w = *s.w;
    const Mats& mt = *s.mt;
    const int mE = mt.mE;
    const int ldE = s.d->ldE;
    const double* SE = mt.SEinvP ? mt.SEinvP : mt.SEinv;
    const int seld = mt.SEinvP ? 0 : ldE;
    // K^-1 [r1; r2_E]:  u = Hinv r1, vE = SEinv (A_E u - r2_E), t = u - (Hinv A_E') vE
    LCQ_PROF(w.sc, 1);
    op_mv_s(mt.oHinv, w.r1, nullptr, 1.0, (mE > 0 || nw > 0) ? w.u : w.dx);
    if (mE > 0) {
        op_mv_s(mt.oAHE, w.r1, w.r2, -1.0, w.cE, mt.eidx);   // cE = r2_E - AHE r1: the sign is undone below
        LCQ_SYNC();
        sym_apply(SE, seld, mE, w.cE, -1.0, w.vE, nw == 0 ? mt.eidx : nullptr, w.dlam, nullptr);
        LCQ_SYNC();
        op_mv_s(mt.oAHtE, w.vE, w.u, -1.0, nw == 0 ? w.dx : w.t);
    } else if (nw > 0) {
        LCQ_SYNC();
        LCQ_LOOP for (int j = LCQ_TID; j < n; j += LCQ_NT) w.t[j] = w.u[j];
    }
    LCQ_SYNC();
    if (nw == 0) return;
    // lI = Tinv (A_I t - r2_I)
    LCQ_PROF(w.sc, 2);
    op_mv_rows(mt.oA, w.idx, nw, w.t, w.r2, w.dI);
    LCQ_SYNC();
    sym_apply(w.Tinv, w.tld, nw, w.dI, 1.0, nullptr, w.idx, w.yf, w.dlam);
    LCQ_SYNC();
    // second K^-1 on [r1 - A_I' lI; r2_E]
    LCQ_PROF(w.sc, 3);
    op_mv_s(mt.oAt, w.yf, w.r1, -1.0, w.t);
    LCQ_SYNC();
    op_mv_s(mt.oHinv, w.t, nullptr, 1.0, mE > 0 ? w.u : w.dx);
    LCQ_LOOP for (int a = LCQ_TID; a < nw; a += LCQ_NT) w.yf[w.idx[a]] = 0.0;
    if (mE > 0) {
        op_mv_s(mt.oAHE, w.t, w.r2, -1.0, w.cE, mt.eidx);
        LCQ_SYNC();
        sym_apply(SE, seld, mE, w.cE, -1.0, w.vE, mt.eidx, w.dlam, nullptr);
        LCQ_SYNC();
        op_mv_s(mt.oAHtE, w.vE, w.u, -1.0, w.dx);
    }
    LCQ_SYNC();
}

// Residual of the UNREGULARISED KKT system of working set W at (x, lam):
//   r1 = -q - P x - A' lam,  r2[i] = b_i - (A x)_i for rows in W (0 elsewhere);  also leaves zx = A x.
// Returns the infinity norm of (r1, r2).
LCQ_DEVN double kkt_residual(QP& s, const signed char* W, const double* x, const double* lam)
{
    const int n = s.d->n, m = s.d->m;
    const Work& w = *s.w;
    const Mats& mt = *s.mt;
    LCQ_PROF(w.sc, 14);
    op_mv_s(mt.oP, x, w.q, 1.0, w.px);            // q + P x
    LCQ_PROF(w.sc, 15);
    op_mv_s(mt.oAt, lam, nullptr, 1.0, w.u);      // A' lam
    LCQ_PROF(w.sc, 14);
    op_mv_s(mt.oA, x, nullptr, 1.0, w.zx);
    LCQ_SYNC();
    LCQ_PROF(w.sc, 0);
    double rn = 0;
    LCQ_LOOP for (int i = LCQ_TID; i < m; i += LCQ_NT) {
        double r = 0.0;
        if (W[i]) r = ((W[i] == 1) ? w.l[i] : w.ub[i]) - w.zx[i];
        w.r2[i] = r;
        rn = fmax(rn, fabs(r));
    }
    LCQ_LOOP for (int j = LCQ_TID; j < n; j += LCQ_NT) {
        const double r = -w.px[j] - w.u[j];
        w.r1[j] = r;
        rn = fmax(rn, fabs(r));
    }
    return fast_max(rn, w.sc);
}

// KKT conditions of the full QP at (x, lam) with zx, r1 as left by kkt_residual.  0 = satisfied;
// 2 stationarity, 3 active row off its bound, 4 inactive row violated, 5/6 wrong multiplier sign
// (*worst = position in idx to drop).
LCQ_DEVN int kkt_check(QP& s, const signed char* W, const double* lam, int* worst)
{
    const int n = s.d->n, m = s.d->m, nw = s.nw;
    const Work& w = *s.w;
    const double ftol = s.o->qp_feas_tol, dtol = s.o->qp_dual_tol;
    double ln = 0, rs = 0;
    LCQ_LOOP for (int i = LCQ_TID; i < m; i += LCQ_NT) ln = fmax(ln, fabs(lam[i]));
    LCQ_LOOP for (int j = LCQ_TID; j < n; j += LCQ_NT) rs = fmax(rs, fabs(w.r1[j]));
    ln = block_max(ln, w.sc);
    rs = block_max(rs, w.sc);
    if (!(rs <= kResTol * (1.0 + ln))) return 2;
    int bad = 0;
    LCQ_LOOP for (int i = LCQ_TID; i < m; i += LCQ_NT) {
        const double zi = w.zx[i];
        if (W[i]) {
            const double b = (W[i] == 1) ? w.l[i] : w.ub[i];
            if (!(fabs(b - zi) <= kResTol * (1.0 + fabs(b)))) bad |= 2;
        }
        const double tol = ftol * (1.0 + fabs(zi));
        if (zi < w.l[i] - tol || zi > w.ub[i] + tol) bad |= 1;
    }
    bad = block_or(bad, w.sc);
    if (bad & 2) return 3;
    if (bad & 1) return 4;
    const double thr = dtol * (1.0 + ln);
    double bv = -1.0;
    int bi = -1;
    LCQ_LOOP for (int a = LCQ_TID; a < nw; a += LCQ_NT) {
        const int i = w.idx[a];
        if (w.pin[i] & 1) continue;
        const double v = (W[i] == 1) ? lam[i] : -lam[i];  // OSQP sign: lower-active needs lam <= 0
        if (v > thr && (bi < 0 || v > bv)) { bv = v; bi = a; }
    }
    double vout;
    const int pos = block_argmax(bv, bi, &vout, w.sc);
    if (pos < 0) return 0;
    if (worst) *worst = pos;
    return (W[w.idx[pos]] == 1) ? 5 : 6;
}

// (xa, lam) is the exact optimum on working set W: make it the accepted solution of this QP.
LCQ_DEVN void accept_solution(QP& s, const signed char* W)
{
    const int n = s.d->n, m = s.d->m;
    const Work& w = *s.w;
    const Mats& mt = *s.mt;
    LCQ_LOOP for (int i = LCQ_TID; i < m; i += LCQ_NT) {
        w.W[i] = W[i];
        const double yi = W[i] ? w.lam[i] : 0.0;
        w.y[i] = yi;
        w.ys[i] = -(mt.E[i] * yi);   // un-scale (osqp auxil.c:524-562, c = 1), qpOASES sign
        w.z[i] = w.zx[i];
    }
    LCQ_LOOP for (int j = LCQ_TID; j < n; j += LCQ_NT) w.x[j] = w.xa[j];
    if (LCQ_TID == 0) s.have_W = 1;
    LCQ_SYNC();
}

// Active-set iteration from the feasible point xa whose active rows contain the working set W (Tinv valid
// for W), multiplier estimate lam (zero outside W).
//   ratio_test = false: plain iterative refinement of the EQP on W (used to probe an ADMM guess), stops at
//                       the converged point without looking at the inactive rows; returns the KKT verdict.
//   ratio_test = true : the full primal active-set method; returns 0 with the solution accepted, 1 if it gave up.
LCQ_DEVN int active_set(QP& s, signed char* W, bool ratio_test, int* worst_out)
{
    const int n = s.d->n, m = s.d->m;
    const Work& w = *s.w;
    const Mats& mt = *s.mt;
    const int cap_it = 20 * (n + m) + 100;
    int last_dropped = -1;
    double best = INFINITY;
    int passes = 0;      // corrections since the last working-set change
    LCQ_LOOP for (int i = LCQ_TID; i < m; i += LCQ_NT) { w.yf[i] = 0.0; w.dlam[i] = 0.0; }   // invariants of kkt_solve
    LCQ_SYNC();
    bool dirty = ratio_test;   // (xa, lam) is not the from-zero solution on W
    bool clean = false;        // the from-zero recomputation is running: no ratio tests
    LCQ_LOOP for (int it = 0; it < cap_it; it++) {
        LCQ_PROF(w.sc, 0);
        const double rn = kkt_residual(s, W, w.xa, w.lam);
        LCQ_PROF(w.sc, 7);
        bool converged = false, slow = false;
        if (passes > 0 && !(rn < best)) {
            // the last correction did not help: undo it and take that point as the EQP solution
            LCQ_LOOP for (int j = LCQ_TID; j < n; j += LCQ_NT) w.xa[j] -= w.dx[j];
            LCQ_LOOP for (int i = LCQ_TID; i < m; i += LCQ_NT) w.lam[i] -= w.dlam[i];
            LCQ_SYNC();
            // At the round-off floor (both residuals tiny) zx and r1 of the point just tried differ from those of
            // the restored point by the size of the last correction, far below every tolerance of kkt_check:
            // they are not recomputed.
            if (!(best < kFloorTol && rn < kFloorTol)) kkt_residual(s, W, w.xa, w.lam);
            converged = true;
        } else {
            // A correction that barely helps means the remaining residual lies along directions whose
            // curvature is far below the regularisation delta (the refinement contracts by delta/(c+delta)):
            // the next direction is stretched to the exact minimiser along it (see below).
            slow = passes > 0 && rn > 0.25 * best && rn > 1e-13;
            best = rn;
            if (rn < 1e-15 || passes >= s.o->qp_refine_iter) converged = true;
        }
        if (converged) {
            int worst = -1;
            LCQ_PROF(w.sc, 8);
            const int reason = kkt_check(s, W, w.lam, &worst);
            LCQ_PROF(w.sc, 7);
            if (!ratio_test) { if (worst_out) *worst_out = worst; return reason; }
            if (reason == 0) {
                accept_solution(s, W);
                if (!dirty) return 0;
                // The point was reached along a path of partial steps.  Recompute it from zero on the final
                // working set so that the accepted solution is a function of (W, q) only (the path leaves
                // round-off that a symmetric problem would amplify, and the oracle computes it this way).
                // Where the EQP on W has no unique solution the recomputation may land elsewhere; then the
                // point accepted above stands.
                LCQ_LOOP for (int j = LCQ_TID; j < n; j += LCQ_NT) w.xa[j] = 0.0;
                LCQ_LOOP for (int i = LCQ_TID; i < m; i += LCQ_NT) w.lam[i] = 0.0;
                LCQ_SYNC();
                dirty = false;
                best = INFINITY;
                passes = 0;
                clean = true;
                continue;
            }
            if (clean) return 0;   // the recomputation did not verify: keep the solution accepted before it
            if ((reason == 5 || reason == 6) && worst >= 0) {
                const int row = w.idx[worst];
                LCQ_SYNC();
                if (LCQ_TID == 0) { W[row] = 0; w.lam[row] = 0.0; w.dlam[row] = 0.0; }
                LCQ_LOOP for (int i = LCQ_TID; i < m; i += LCQ_NT) w.pin[i] &= 1;
                LCQ_PROF(w.sc, 9);
                tinv_remove(s, worst);
                LCQ_PROF(w.sc, 7);
                last_dropped = row;
                if (LCQ_TID == 0) s.n_changes++;
                best = INFINITY;
                passes = 0;
                dirty = true;
                clean = false;
                continue;
            }
            return 1;  // EQP not solvable to tolerance on this set
        }
        if (LCQ_TID == 0) s.n_pass++;
        kkt_solve(s);   // r1, r2 -> dx, dlam
        LCQ_PROF(w.sc, 4);
        if (slow && ratio_test && !clean) {
            // exact line search along dx (the working-set rows are satisfied, so dx moves inside them):
            // tau = r1'dx / dx'P dx >= 1; the ratio test below then cuts the stretched step at the first bound
            op_mv_s(mt.oP, w.dx, nullptr, 1.0, w.px);
            LCQ_SYNC();
            double pd = 0, cd = 0;
            LCQ_LOOP for (int j = LCQ_TID; j < n; j += LCQ_NT) { pd += w.r1[j] * w.dx[j]; cd += w.px[j] * w.dx[j]; }
            fast_sum2(pd, cd, w.sc);
            if (pd > 0.0) {
                const double tau = (cd > 0.0 && pd < 1e12 * cd) ? pd / cd : 1e12;
                if (tau > 1.0) LCQ_LOOP for (int j = LCQ_TID; j < n; j += LCQ_NT) w.dx[j] *= tau;
                LCQ_SYNC();
            }
        }
        double amin = 1.0;
        int block = -1;
        double apn = 0;
        if (ratio_test && !clean) {
            // ratio test against the inactive rows (Nocedal & Wright alg. 16.3)
            LCQ_PROF(w.sc, 5);
            op_mv_s(mt.oA, w.dx, nullptr, 1.0, w.zp);
            LCQ_SYNC();
            LCQ_LOOP for (int i = LCQ_TID; i < m; i += LCQ_NT) apn = fmax(apn, fabs(w.zp[i]));
            apn = fast_max(apn, w.sc);
            const double seps = 1e-13 * (1.0 + apn);
            LCQ_LOOP for (;;) {
                // the smallest step length; among the rows attaining it the largest |s|, then the smallest row
                double ba = 2.0, bs = 0.0;
                int bi = -1;
                LCQ_LOOP for (int i = LCQ_TID; i < m; i += LCQ_NT) {
                    if (W[i] || w.ctype[i] != 0 || (w.pin[i] & 2)) continue;
                    const double sv = w.zp[i];
                    double a;
                    if (sv < -seps && w.l[i] > -INFINITY) a = fmax(w.zx[i] - w.l[i], 0.0) / (-sv);
                    else if (sv > seps && w.ub[i] < INFINITY) a = fmax(w.ub[i] - w.zx[i], 0.0) / sv;
                    else continue;
                    if (a < 1.0 && lex_better(ba, bs, bi, a, fabs(sv), i)) { ba = a; bs = fabs(sv); bi = i; }
                }
                double am;
                block = fast_argmin_lex(ba, bs, bi, &am, w.sc);
                amin = block >= 0 ? am : 1.0;
                if (block < 0) break;
                if (s.nw >= s.d->cap) return 1;
                LCQ_PROF(w.sc, 6);
                const int dep_ = tinv_append(s, block);
                LCQ_PROF(w.sc, 5);
                if (dep_ == 0) break;
                // The blocking row is a combination of the active rows (e.g. a box bound duplicating an active
                // selection row): along the working set it cannot move; the apparent motion is the leak of the
                // regularisation.  It is left out of the ratio tests until a row leaves the working set.
                if (LCQ_TID == 0) w.pin[block] |= 2;
                LCQ_SYNC();
                amin = 1.0;
            }
        }
        LCQ_PROF(w.sc, 7);
        if (block >= 0) {
            const signed char side = (w.zp[block] < 0) ? 1 : 2;
            LCQ_LOOP for (int j = LCQ_TID; j < n; j += LCQ_NT) w.xa[j] += amin * w.dx[j];
            LCQ_LOOP for (int i = LCQ_TID; i < m; i += LCQ_NT) w.lam[i] += amin * w.dlam[i];
            // a row that comes straight back after a zero-length step was dropped on multiplier noise:
            // it is weakly active; exempt it from the sign test for the rest of this QP (anti-cycling)
            if (LCQ_TID == 0) { W[block] = side; if (block == last_dropped && amin * (1.0 + apn) <= 1e-12) w.pin[block] |= 1; }
            last_dropped = -1;
            LCQ_SYNC();
            if (LCQ_TID == 0) s.n_changes++;
            best = INFINITY;
            passes = 0;
            continue;
        }
        LCQ_LOOP for (int j = LCQ_TID; j < n; j += LCQ_NT) w.xa[j] += w.dx[j];
        LCQ_LOOP for (int i = LCQ_TID; i < m; i += LCQ_NT) w.lam[i] += w.dlam[i];
        LCQ_SYNC();
        passes++;
    }
    return 1;
}

// H-metric projection of xin onto the rows of the working set + feasibility test of every row -> xa.
LCQ_DEVN int project_feasible(QP& s, const signed char* W, const double* xin)
{
    const int n = s.d->n, m = s.d->m;
    const Work& w = *s.w;
    const Mats& mt = *s.mt;
    LCQ_LOOP for (int j = LCQ_TID; j < n; j += LCQ_NT) w.xa[j] = xin[j];
    LCQ_LOOP for (int i = LCQ_TID; i < m; i += LCQ_NT) { w.yf[i] = 0.0; w.dlam[i] = 0.0; }   // invariants of kkt_solve
    LCQ_SYNC();
    LCQ_LOOP for (int pass = 0; pass < 4; pass++) {
        op_mv_s(mt.oA, w.xa, nullptr, 1.0, w.zx);
        LCQ_SYNC();
        double rn = 0;
        LCQ_LOOP for (int i = LCQ_TID; i < m; i += LCQ_NT) {
            double r = 0.0;
            if (W[i]) r = ((W[i] == 1) ? w.l[i] : w.ub[i]) - w.zx[i];
            w.r2[i] = r;
            rn = fmax(rn, fabs(r));
        }
        LCQ_LOOP for (int j = LCQ_TID; j < n; j += LCQ_NT) w.r1[j] = 0.0;
        rn = block_max(rn, w.sc);
        if (rn < 1e-15) break;
        kkt_solve(s);
        LCQ_LOOP for (int j = LCQ_TID; j < n; j += LCQ_NT) w.xa[j] += w.dx[j];
        LCQ_SYNC();
    }
    op_mv_s(mt.oA, w.xa, nullptr, 1.0, w.zx);
    LCQ_SYNC();
    const double ftol = s.o->qp_feas_tol;
    int bad = 0;
    LCQ_LOOP for (int i = LCQ_TID; i < m; i += LCQ_NT) {
        const double tol = ftol * (1.0 + fabs(w.zx[i]));
        if (w.zx[i] < w.l[i] - tol || w.zx[i] > w.ub[i] + tol) bad = 1;
        if (W[i]) {
            const double b = (W[i] == 1) ? w.l[i] : w.ub[i];
            if (fabs(w.zx[i] - b) > tol) bad = 1;
        }
    }
    return block_or(bad, w.sc) == 0;
}

// SubsolverBase::solve contract (/root/reference/include/SubsolverBase.hpp:37-56).  g unscaled,
// x0 / y0 (m entries, qpOASES sign) may be null.  Returns 0 or a non-zero flag;
// *iterations = ADMM iterations + working-set changes.
LCQ_DEVN int qp_solve(QP& s, bool initial, const double* g, const double* x0, const double* y0A, const double* y0box, int* iterations, bool infeasible)
{
    const int n = s.d->n, m = s.d->m, mA = s.d->mA;
    const Work& w = *s.w;
    const Mats& mt = *s.mt;
    const lcqp_cuda_options& o = *s.o;
    *iterations = 0;
    if (infeasible) return 37;
    const long long ch0 = s.n_changes;
    LCQ_LOOP for (int j = LCQ_TID; j < n; j += LCQ_NT) w.q[j] = mt.D[j] * g[j];  // osqp.c:752-779, c = 1
    LCQ_LOOP for (int i = LCQ_TID; i < m; i += LCQ_NT) w.pin[i] = 0;
    LCQ_SYNC();
    if (initial) {
        // osqp_warm_start_x/_y (osqp.c:700-745)
        LCQ_LOOP for (int j = LCQ_TID; j < n; j += LCQ_NT) w.x[j] = x0 ? x0[j] / mt.D[j] : 0.0;
        LCQ_LOOP for (int i = LCQ_TID; i < m; i += LCQ_NT) {
            double yv = 0.0;
            if (i < mA) { if (y0A) yv = y0A[i]; }
            else if (y0box) yv = y0box[i - mA];
            w.y[i] = -yv / mt.E[i];
        }
        LCQ_SYNC();
        op_mv_s(mt.oA, w.x, nullptr, 1.0, w.z);
        if (LCQ_TID == 0) { s.have_W = 0; s.tinv_valid = 0; }
        LCQ_SYNC();
    } else if (s.have_W && s.tinv_valid) {
        // hot start: the previous working set is tried first (EQP from zero: the usual case late in the
        // homotopy, where the active set no longer changes) ...
        LCQ_LOOP for (int j = LCQ_TID; j < n; j += LCQ_NT) w.xa[j] = 0.0;
        LCQ_LOOP for (int i = LCQ_TID; i < m; i += LCQ_NT) { w.Wtry[i] = w.W[i]; w.lam[i] = 0.0; }
        LCQ_SYNC();
        const int reason = active_set(s, w.Wtry, false, nullptr);
        if (reason == 0) { accept_solution(s, w.Wtry); return 0; }
        if (reason != 5 && reason != 6) {
            // ... else the active-set iteration continues from the previous optimum (feasible: the bounds did
            // not change) with its multipliers; a feasible EQP point with a wrong-signed multiplier is kept
            LCQ_LOOP for (int j = LCQ_TID; j < n; j += LCQ_NT) w.xa[j] = w.x[j];
            LCQ_LOOP for (int i = LCQ_TID; i < m; i += LCQ_NT) w.lam[i] = w.y[i];
            LCQ_SYNC();
        }
        if (active_set(s, w.Wtry, true, nullptr) == 0) { *iterations = (int)(s.n_changes - ch0); return 0; }
        // fall through to ADMM from the previous solution
        LCQ_SYNC();
        if (LCQ_TID == 0) s.tinv_valid = 0;
        op_mv_s(mt.oA, w.x, nullptr, 1.0, w.z);
        LCQ_SYNC();
    }
    int have_fail = 0;
    int it = 0;
    while (it < o.qp_max_iter) {
        LCQ_PROF(w.sc, 10);
        LCQ_LOOP for (int k = 0; k < o.qp_check_interval && it < o.qp_max_iter; k++, it++) admm_iter(s);
        LCQ_PROF(w.sc, 12);
        if (LCQ_TID == 0) s.n_admm += o.qp_check_interval;
        guess_working_set(s, w.Wtry);
        if (have_fail) {
            int diff = 0;
            LCQ_LOOP for (int i = LCQ_TID; i < m; i += LCQ_NT) diff |= (w.Wtry[i] != w.Wfail[i]);
            if (!block_or(diff, w.sc)) continue;
        }
        LCQ_LOOP for (int i = LCQ_TID; i < m; i += LCQ_NT) w.Wfail[i] = w.Wtry[i];
        have_fail = 1;
        LCQ_SYNC();
        tinv_build(s, w.Wtry);
        // EQP on the guessed set, from zero
        LCQ_LOOP for (int j = LCQ_TID; j < n; j += LCQ_NT) w.xa[j] = 0.0;
        LCQ_LOOP for (int i = LCQ_TID; i < m; i += LCQ_NT) w.lam[i] = 0.0;
        LCQ_SYNC();
        const int reason = active_set(s, w.Wtry, false, nullptr);
        if (reason == 0) { accept_solution(s, w.Wtry); *iterations = it + (int)(s.n_changes - ch0); return 0; }
        if (reason == 5 || reason == 6) {
            // primal feasible EQP point with a wrong-signed multiplier: a valid active-set start
            if (active_set(s, w.Wtry, true, nullptr) == 0) { *iterations = it + (int)(s.n_changes - ch0); return 0; }
            LCQ_SYNC();
            if (LCQ_TID == 0) s.tinv_valid = 0;
            LCQ_SYNC();
        } else if (project_feasible(s, w.Wtry, w.x)) {
            LCQ_LOOP for (int i = LCQ_TID; i < m; i += LCQ_NT) w.lam[i] = 0.0;
            LCQ_SYNC();
            if (active_set(s, w.Wtry, true, nullptr) == 0) { *iterations = it + (int)(s.n_changes - ch0); return 0; }
            LCQ_SYNC();
            if (LCQ_TID == 0) s.tinv_valid = 0;
            LCQ_SYNC();
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

// out = Q v + rho * (L' (R v) + R' (L v)) + add   (i.e. Qk v + add); leaves Lx = L v, Rx = R v
LCQ_DEVN void qk_apply(const RawOps& ro, double rho, const double* v, const double* add, double* out, const Work& w)
{
    op_mv(ro.Q, v, add, 1.0, w.tn);
    op_mv(ro.L, v, nullptr, 1.0, w.Lx);
    op_mv(ro.R, v, nullptr, 1.0, w.Rx);
    LCQ_SYNC();
    op_mv(ro.Lt, w.Rx, w.tn, rho, out);
    LCQ_SYNC();
    op_mv(ro.Rt, w.Lx, out, rho, out);
    LCQ_SYNC();
}

LCQ_DEVN void lcqp_loop(QP& s, const Inst& in, const RawOps& ro, unsigned long long instance,
                       bool infeasible, double* xout, double* yout, LoopOut& out)
{
    const Dims& d = *s.d;
    const int n = d.n, nC = d.nC, nComp = d.nComp, mA = d.mA;
    const Work& w = *s.w;
    const Mats& mt = *s.mt;
    const lcqp_cuda_options& o = *s.o;
    const bool osqp_flavour = (o.qpSolver == 2);
    const int boxOff = osqp_flavour ? 0 : n;

    double hist[kMaxLeyffer];
    int nh = 0;
    double alphak = 1.0, rho = o.initialPenaltyParameter, phi_const = 0.0;
    int outerIter = 0, totalIter = 0, subIter = 0, qpIter = 0, exitFlag = 0, status = 0, ret = RET_OK;
    out.rhoOpt = 0.0;
    const bool have_gphi = (in.lbL != nullptr) || (in.lbR != nullptr);

    LCQ_LOOP for (int j = LCQ_TID; j < n; j += LCQ_NT) {
        w.xk[j] = in.x0 ? in.x0[j] : 0.0;   // LCQProblem.ipp:138-142
        w.gt[j] = in.g[j];                  // g_tilde = g (LCQProblem.cpp:966-967)
        w.pk[j] = 0.0;
        w.gphi[j] = 0.0;
    }
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
        const double* y0A = nullptr;
        const double* y0box = nullptr;
        if (initial && in.y0) { y0A = in.y0 + n; y0box = d.has_box ? in.y0 : nullptr; }
        LCQ_PROF(w.sc, 12);
        const int fl = qp_solve(s, initial, w.gk, w.xk, y0A, y0box, &qpIter, infeasible);
        LCQ_PROF(w.sc, 11);
        subIter += qpIter;
        exitFlag = osqp_flavour ? (fl == 0 ? 1 : fl) : fl;
        if (fl != 0) { ret = (osqp_flavour && infeasible) ? RET_OSQP_GUESS : RET_SUBPROBLEM; return false; }
        LCQ_LOOP for (int j = LCQ_TID; j < n; j += LCQ_NT) w.pk[j] = mt.D[j] * w.x[j] - w.xk[j];  // xnew = D xbar (auxil.c:524-562)
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
        qk_apply(ro, rho, w.xk, w.gt, w.stat, w);
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
        // (the reference linearises here, :508, and again at :545 before the QP; only the second one is used)

        double sm = 0;
        LCQ_LOOP for (int j = LCQ_TID; j < n; j += LCQ_NT) sm = fmax(sm, fabs(w.stat[j]));
        sm = block_max(sm, w.sc);
        if (sm < o.stationarityTolerance) {  // :511
            if (phi() < o.complementarityTolerance) {
                // determineStationarityType :1412-1453 on yk_A (the PENALISED duals, :1420),
                // weak set :1456-1482 (phi() left Lx, Rx in shared memory)
                const double tc = o.complementarityTolerance;
                int fl = 0;  // bit0: s fails, bit1: m fails, bit2: weakly stationary only
                LCQ_LOOP for (int i = LCQ_TID; i < nComp; i += LCQ_NT) {
                    if (!(w.Lx[i] <= tc && w.Rx[i] <= tc)) continue;
                    const double yl = w.ys[nC + i], yr = w.ys[nC + nComp + i];
                    const double prod = yl * yr, mn = fmin(yl, yr);
                    if (mn < 0) fl |= 1;
                    if (fabs(prod) >= tc && mn <= 0) { if (prod <= tc) fl |= 4; else fl |= 2; }
                }
                // the reference returns W at the FIRST weak index whose test fails c-stationarity; any
                // such index gives W, so an OR over the weak set is the same decision
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
            qk_apply(ro, rho, w.pk, nullptr, w.stat, w);  // Qk pk
            double p1 = 0;
            LCQ_LOOP for (int j = LCQ_TID; j < n; j += LCQ_NT) p1 += w.stat[j] * w.pk[j];
            const double qk = block_sum(p1, w.sc);
            qk_apply(ro, rho, w.xk, w.gt, w.stat, w);     // Qk xk + g_tilde
            p1 = 0;
            LCQ_LOOP for (int j = LCQ_TID; j < n; j += LCQ_NT) p1 += w.stat[j] * w.pk[j];
            const double lk = block_sum(p1, w.sc);
            alphak = 1.0;
            if (qk > 0 && lk < 0) alphak = fmin(-lk / qk, 1.0);
        }
    }

    // outputs: x = xk ; y = [box duals ; yk_A] (transformDuals :1381-1409 applied on success)
    LCQ_LOOP for (int j = LCQ_TID; j < n; j += LCQ_NT) xout[j] = w.xk[j];
    const bool have_y = !infeasible && s.have_W;
    if (!osqp_flavour)
        LCQ_LOOP for (int j = LCQ_TID; j < n; j += LCQ_NT) yout[j] = (d.has_box && have_y) ? w.ys[mA + j] : 0.0;
    if (success) {
        // Lx, Rx from the last phi() call hold L xk, R xk
        LCQ_LOOP for (int i = LCQ_TID; i < mA; i += LCQ_NT) {
            double v = w.ys[i];
            if (i >= nC && i < nC + nComp) v -= rho * w.Rx[i - nC];
            else if (i >= nC + nComp) v -= rho * w.Lx[i - nC - nComp];
            yout[boxOff + i] = v;
        }
    } else {
        LCQ_LOOP for (int i = LCQ_TID; i < mA; i += LCQ_NT) yout[boxOff + i] = have_y ? w.ys[i] : 0.0;
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
// memory plan of one CTA
// ------------------------------------------------------------------------------------------------
// Doubles that are always in shared memory (QP iterates), doubles of the outer loop (shared memory when
// they fit, else the CTA's global scratch), packed Tinv (likewise).
struct SmemPlan {
    size_t bytes;        // dynamic shared memory of the CTA
    int outer_in_smem;
    int tinv_in_smem;
    size_t gl_doubles;   // doubles of global scratch per CTA for what did not fit
    int ys_global;       // the accepted duals (read by the outer loop only) live in the global scratch
    int bounds_shared;   // l / ub of all groups alias one copy behind the groups (bounds shared by the batch)
};

// every vector starts on a 16-byte boundary (vector loads): lengths are rounded up to even
inline LCQ_HD size_t ev(size_t k) { return (k + 1) & ~(size_t)1; }
inline LCQ_HD size_t qp_doubles(const Dims& d) { return 7ull * ev(d.n) + 10ull * ev(d.m) + 2ull * ev(d.cap) + 2ull * ev(d.capE) + ev(d.m) /*ys*/; }
inline LCQ_HD size_t outer_doubles(const Dims& d) { return 7ull * ev(d.n) + 2ull * ev(d.nComp); }
inline LCQ_HD size_t tinv_doubles(const Dims& d) { return ev(packed_doubles(d.cap)); }
// working-set inverse in global memory: full storage, even leading dimension (16-byte loads of column pairs)
inline LCQ_HD int tinv_ld(const Dims& d) { return (d.cap + 1) & ~1; }
inline LCQ_HD size_t tinv_gl_doubles(const Dims& d) { return (size_t)d.cap * tinv_ld(d); }
inline LCQ_HD size_t misc_bytes(const Dims& d)
{
    return (size_t)d.cap * sizeof(int) + 5ull * ((d.m + 15) / 16) * 16 + sizeof(Scalars) + 64;
}

// tinv_global: keep the working-set inverse in global memory (full storage) even if it would fit
inline LCQ_HD SmemPlan make_plan(const Dims& d, size_t budget, bool tinv_global = false)
{
    SmemPlan p;
    size_t b = qp_doubles(d) * sizeof(double) + misc_bytes(d);
    p.gl_doubles = 0;
    p.ys_global = 0;
    p.bounds_shared = 0;
    p.tinv_in_smem = !tinv_global && (b + tinv_doubles(d) * sizeof(double) <= budget);
    if (p.tinv_in_smem) b += tinv_doubles(d) * sizeof(double); else p.gl_doubles += tinv_gl_doubles(d);
    p.outer_in_smem = (b + outer_doubles(d) * sizeof(double) <= budget);
    if (p.outer_in_smem) b += outer_doubles(d) * sizeof(double); else p.gl_doubles += outer_doubles(d);
    p.bytes = b;
    return p;
}

// shared_bounds: 2 ev(m) doubles shared by the groups of the CTA (plan.bounds_shared), else null
LCQ_DEV void carve(Work& w, const Dims& d, const SmemPlan& p, unsigned char* base, double* gl, double* shared_bounds = nullptr)
{
    double* q = reinterpret_cast<double*>(base);
    auto take = [&](size_t k) { double* r = q; q += ev(k); return r; };
    auto takeg = [&](size_t k) { double* r = gl; gl += ev(k); return r; };
    const int n = d.n, m = d.m;
    w.q = take(n); w.x = take(n); w.xa = take(n); w.r1 = take(n); w.u = take(n); w.t = take(n); w.dx = take(n);
    w.px = w.t;   // q + P x lives only inside kkt_residual, t only inside kkt_solve / admm_iter
    w.z = take(m); w.y = take(m); w.lam = take(m); w.dlam = take(m); w.r2 = take(m);
    if (p.bounds_shared && shared_bounds) { w.l = shared_bounds; w.ub = shared_bounds + ev(m); }
    else { w.l = take(m); w.ub = take(m); }
    w.zx = take(m); w.zp = take(m); w.yf = take(m);
    w.w = w.yf;   // ADMM scratch; yf is scratch of the active-set passes (zeroed before they start)
    w.dI = take(d.cap); w.lI = take(d.cap);
    w.cE = take(d.capE); w.vE = take(d.capE);
    w.ys = p.ys_global ? takeg(m) : take(m);
    w.Tinv = p.tinv_in_smem ? take(tinv_doubles(d)) : takeg(tinv_gl_doubles(d));
    w.tld = p.tinv_in_smem ? 0 : tinv_ld(d);
    auto tk = [&](size_t k) { return p.outer_in_smem ? take(k) : takeg(k); };
    w.xk = tk(n); w.pk = tk(n); w.gk = tk(n); w.gt = tk(n); w.gphi = tk(n); w.stat = tk(n); w.tn = tk(n);
    w.Lx = tk(d.nComp); w.Rx = tk(d.nComp);
    w.idx = reinterpret_cast<int*>(q);
    signed char* c = reinterpret_cast<signed char*>(w.idx + d.cap);
    const int mpad = ((m + 15) / 16) * 16;
    w.W = c; w.Wtry = c + mpad; w.Wfail = c + 2 * mpad; w.ctype = c + 3 * mpad; w.pin = c + 4 * mpad;
    uintptr_t sp = reinterpret_cast<uintptr_t>(c + 5 * mpad);
    sp = (sp + 15) & ~(uintptr_t)15;
    w.sc = reinterpret_cast<Scalars*>(sp);
}

// Layout of a Mats block in a flat array of doubles (+ ints at the end).
inline LCQ_HD size_t mats_doubles(const Dims& d)
{
    const size_t n = d.n, m = d.m;
    return 3 * ev(n * n) + 3 * ev(m * n) + ev(n) + ev(m) + ev(m * m) + ev((size_t)d.ldE * d.ldE) + ev((size_t)d.ldE * n) + ev((m + 1) / 2) /*eidx*/ + ev((m + 7) / 8) /*ctype*/ + 2;
}

// every block starts on a 16-byte boundary (the store itself comes from cudaMalloc / an even offset)
LCQ_DEV void carve_mats(Mats& mt, double* base, const Dims& d)
{
    const int mEcap = d.ldE;
    const size_t n = d.n, m = d.m;
    mt.P = base; base += ev(n * n);
    mt.A = base; base += ev(m * n);
    mt.At = base; base += ev(m * n);
    mt.D = base; base += ev(n);
    mt.E = base; base += ev(m);
    mt.Hinv = base; base += ev(n * n);
    mt.AH = base; base += ev(m * n);
    mt.T = base; base += ev(m * m);
    mt.Minv = base; base += ev(n * n);
    mt.SEinv = base; base += ev((size_t)mEcap * mEcap);
    mt.AHE = base; base += ev((size_t)mEcap * n);
    mt.eidx = reinterpret_cast<int*>(base); base += ev((m + 1) / 2);
    mt.ctype = reinterpret_cast<signed char*>(base);
    mt.SEinvP = nullptr;
    mt.mE = 0; mt.status = 0; mt.cache_bytes_se = 0; mt.cache_bytes_hot = 0; mt.cache_bytes_raw = 0;
}

// ------------------------------------------------------------------------------------------------
// One instance, start to finish (used by the persistent kernel and by the host emulation).
// `mats_shared`: mt was prepared by the batch-level prepare step; otherwise this CTA prepares it here.
// ------------------------------------------------------------------------------------------------
LCQ_DEVN void run_instance(QP& s, Mats& mt, bool mats_shared, const Inst& in, const RawOps& ro, unsigned long long instance,
                           double* xo, double* yo, LoopOut& out, bool bounds_ready = false)
{
    const Dims& d = *s.d;
    const Work& w = *s.w;
    const lcqp_cuda_options& o = *s.o;
    const int nD = d.n + d.mA;
    LCQ_SYNC();
    if (LCQ_TID == 0) { s.mt = &mt; s.nw = 0; s.have_W = 0; s.tinv_valid = 0; s.n_admm = 0; s.n_pass = 0; s.n_changes = 0; }
    LCQ_LOOP for (int k = LCQ_TID; k < 8; k += LCQ_NT) w.sc->wph[k] = 0;
    LCQ_SYNC();
    out.ret = 0; out.status = 0; out.iterTotal = 0; out.iterOuter = 0; out.subIter = 0; out.exitFlag = 0; out.rhoOpt = 0;
    bool skip = false;
    // initializeSolver checks (LCQProblem.cpp:930-957): the OSQP-style layout has no box constraints
    if (o.qpSolver == 2 && (in.lb || in.ub)) { out.ret = RET_INVALID_OSQP_BOX; skip = true; }
    if (!skip) {
        int prep_rc = 0;
        if (!mats_shared) {
            prepare_scale(d, in, mt, w.u, w.zx);
            if (LCQ_TID == 0) mats_dense_ops(d, mt);
            LCQ_SYNC();
        }
        const int bflags = set_bounds(d, in, mt.E, w.l, w.ub, w.ctype, w.sc, !bounds_ready);
        if (bflags & 2) { out.ret = RET_INVALID_LOWER_COMP; skip = true; }  // loadLCQP fails (:747,:767)
        const bool infeasible = (bflags & 1) != 0;
        if (!skip && !infeasible) {
            if (!mats_shared) {
                prep_rc = prepare_factor(d, mt, w.ctype, o, w.u, w.t, w.zx, w.zp, w.w, w.sc);
                if (LCQ_TID == 0 && !prep_rc) mats_dense_ops_post(d, mt);
                LCQ_SYNC();
            } else {
                int diff = (mt.status != 0);
                LCQ_LOOP for (int i = LCQ_TID; i < d.m; i += LCQ_NT) diff |= ((w.ctype[i] == 1) != (mt.ctype[i] >= 1)) || ((w.ctype[i] < 0) != (mt.ctype[i] < 0));
                prep_rc = block_or(diff, w.sc);
            }
            if (!prep_rc) {
                // rows found dependent at prepare time keep their type 2
                LCQ_LOOP for (int i = LCQ_TID; i < d.m; i += LCQ_NT) w.ctype[i] = mt.ctype[i];
                LCQ_SYNC();
            }
        }
        if (!skip) {
            if (prep_rc) { out.ret = RET_SUBPROBLEM; out.exitFlag = 38; skip = true; }
            else lcqp_loop(s, in, ro, instance, infeasible, xo, yo, out);
        }
    }
    if (skip) {
        LCQ_LOOP for (int j = LCQ_TID; j < d.n; j += LCQ_NT) xo[j] = in.x0 ? in.x0[j] : 0.0;
        LCQ_LOOP for (int j = LCQ_TID; j < nD; j += LCQ_NT) yo[j] = 0.0;
    }
    LCQ_SYNC();
}

}  // namespace lcqp

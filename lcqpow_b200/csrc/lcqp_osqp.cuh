// lcqp_osqp.cuh -- the OSQP flavour of the B200-native LCQP path (sm_100a device code; compiles as plain C++ under
// LCQP_HOST_EMU for the CPU test-suite).
//
// What it restates.  With QPSolver::OSQP_SPARSE the reference hands every inner QP to OSQP 0.6.2
// (/root/reference/src/SubsolverOSQP.cpp:124-200): Ruiz equilibration with cost scaling (external/osqp/src/scaling.c:44-156),
// the ADMM iteration on the quasi-definite KKT system (osqp.c:354-519, auxil.c:161-225, proj.c:4-14), a sparse LDL'
// factorisation of K = [P + sigma I, A'; A, -diag(1/rho)] (kkt.c:6-177, qdldl.c:72-281), residuals, termination and
// infeasibility certificates (auxil.c:240-520, :681-786), rho adaptation with refactorisation (auxil.c:13-74,
// osqp.c:1268-1318), polish with iterative refinement and its accept rule (polish.c:19-350), warm start from the
// previous QP (osqp.c:752-832, :954-994), and the adapter's conventions (exit flag, negated duals).  Each function of
// lcqp_osqp_impl.inc cites the lines it follows; the arithmetic is fp64 and every sum runs in the reference's order of
// accumulation, so that a run reproduces the reference's run down to the number of ADMM iterations.
//
// How it is built for the GPU.  A batch shares its sparsity pattern, so everything that depends on the pattern only
// -- ordering, elimination tree, pattern of L, the row programs of the factorisation, the level sets of the triangular
// solves, the row lists of the products (lcqp_sparse_host.hpp) -- is computed once on the host, and the device executes
// flat index lists on per-instance values: no integer workspace, no pattern logic on the device.  The solver source
// (lcqp_osqp_impl.inc) is compiled twice:
//   namespace osqt  ONE THREAD PER INSTANCE.  The arrays of the 32 instances of a warp are interleaved (element i of
//       lane l at [i * 32 + l]): every load and store is one coalesced 256-byte transaction per warp and the index
//       arrays are warp-uniform.  Data-dependent loops run a warp-uniform number of trips with finished lanes masked
//       out, so the lanes stay in lockstep.  For very large batches of small or dense-ish problems (C5-like).
//   namespace osqw  ONE WARP PER INSTANCE.  Vectors are contiguous per instance, elementwise work and the rows of the
//       products are dealt round-robin to the 32 lanes, the triangular solves run level by level (rows of a level are
//       independent; `__syncwarp` between levels), the factorisation updates the touched rows of a pivot step in
//       parallel; scalars are computed redundantly by all lanes, sums that the reference does sequentially stay
//       sequential (bit-identical results in both modes).  For large sparse problems and small batches (C4), where one
//       thread per instance leaves the GPU empty and a single slow instance holds a kernel for minutes.
// The scratch vector of the triangular solves and of the factorisation lives in shared memory when it fits.
// Bound: HBM/L2 streaming of L and of the iterates (16 nnz(L) + 8 N + 8 (3n + 5m) bytes per
// instance-iteration).
//
// Deviations from the reference, stated: (1) the fill-reducing ordering is minimum degree, not AMD -- L differs in
// pattern, the computed iterates differ by round-off only; (2) `adaptive_rho_interval = 0` means "every
// 4 x check_termination iterations" (OSQP's own rule when it is built without its wall-clock profiler) -- the
// reference derives the interval from timings, which makes its own iteration counts machine dependent
// (osqp.c:459-485); (3) a user-supplied y0 warm-starts the duals of [A; L; R] (the reference copies nDuals BYTES of
// them, /root/reference/src/LCQProblem.cpp:947).
#pragma once

#include "lcqp_device.cuh"
#include "lcqp_sparse_host.hpp"   // kStreamVals, kStreamIdx (and the host-side analysis itself)

namespace lcqp {
namespace osq {

// constants.h:59-114
// (kRhoMin, kRhoTol, kRhoEqOverIneq, kMinScaling, kMaxScaling: lcqp_device.cuh)
constexpr double kRhoMax = 1e6;
constexpr double kOsqpInfty = 1e30;
constexpr double kDivTol = 1.0 / 1e30;
enum { OSQP_SOLVED = 1, OSQP_SOLVED_INACCURATE = 2, OSQP_PRIMAL_INFEASIBLE_INACCURATE = 3, OSQP_DUAL_INFEASIBLE_INACCURATE = 4,
       OSQP_MAX_ITER_REACHED = -2, OSQP_PRIMAL_INFEASIBLE = -3, OSQP_DUAL_INFEASIBLE = -4, OSQP_NON_CVX = -7, OSQP_UNSOLVED = -10 };

// device view of osq::Symbolic (lcqp_sparse_host.hpp)
struct SymDev {
    int n, m, N, nC, nComp, nnzP, nnzA, nnzQ, nnzK, nnzL, nflev, nblev;
    int fsChunks, bsChunks, stream;   // streamed triangular solves (lcqp_sparse_host.hpp: Symbolic::fsI ...), warp mode only
    const int *Pp, *Pi, *Psrc, *Ap, *Ai, *Asrc, *Qp, *Qi, *Qsrc, *perm, *Kp, *Ki, *Ksrc, *Lp, *Li, *rp, *rcol, *rpos, *Pcol, *Acol, *Qcol,
        *ArP, *ArE, *PrP, *PrE, *QrP, *QrE, *LrP, *LrC, *rposr, *flP, *flR, *blP, *blC, *fsI, *bsI, *fsSrc, *bsSrc,
        *sOff, *fpIdx, *fpLi, *fpStep, *rowPair;
};
constexpr int kSymArrays = 43;
// shared-memory ring of the streamed sweeps: kStreamStages x (value chunk + index chunk) + one mbarrier per stage
constexpr int kStreamStages = 2;
LCQ_HD inline size_t stream_stage_bytes() { return (size_t)kStreamVals * sizeof(double) + (size_t)kStreamIdx * sizeof(unsigned short); }
LCQ_HD inline size_t stream_ring_bytes() { return kStreamStages * stream_stage_bytes() + 64; }

// one instance's inputs (value arrays as the caller laid them out; NULL = absent)
struct View {
    const double *Q, *A, *L, *R, *g, *lbL, *ubL, *lbR, *ubR, *lbA, *ubA, *x0, *y0;
};

struct State {
    double rho, c, cinv;
    double pri_res, dua_res;
    int status_val, iter, interval;
    int factor_bad;          // a zero pivot, or not exactly n positive pivots (qdldl_interface.c:93-99)
    long long admm_total, factor_count;
};

LCQ_HD inline size_t sm_len(const SymDev& S) { return (size_t)S.N + 8; }   // solve / factor scratch (+ a few scalars)
LCQ_HD inline size_t ws_doubles(const SymDev& S)
{
    return (size_t)S.nnzP + S.nnzA + 4 * (size_t)S.nnzL + 2 * (size_t)S.N + 17 * (size_t)S.n + 17 * (size_t)S.m + 2 * (size_t)S.N + sm_len(S)
           + (S.stream ? 2 * ((size_t)S.fsChunks + S.bsChunks) * kStreamVals + 2 : 0);
}

LCQ_DEV double osq_limit(double d) { d = d < kMinScaling ? 1.0 : d; return d > kMaxScaling ? kMaxScaling : d; }   // scaling.c:7-14
LCQ_DEV bool has_solution(int s)
{
    return s != OSQP_PRIMAL_INFEASIBLE && s != OSQP_PRIMAL_INFEASIBLE_INACCURATE && s != OSQP_DUAL_INFEASIBLE && s != OSQP_DUAL_INFEASIBLE_INACCURATE && s != OSQP_NON_CVX;
}

}  // namespace osq

// ---- mode T: one thread per instance --------------------------------------------------------------------------
#define OSQ_NS osqt
#define OSQ_WARP 0
#include "lcqp_osqp_impl.inc"
#undef OSQ_NS
#undef OSQ_WARP
// ---- mode W: one warp per instance ----------------------------------------------------------------------------
#define OSQ_NS osqw
#define OSQ_WARP 1
#include "lcqp_osqp_impl.inc"
#undef OSQ_NS
#undef OSQ_WARP

}  // namespace lcqp

/*
 * lcqp_cuda.h -- C ABI of the B200-native batched LCQP solver (liblcqp_cuda.so).
 *
 * This is the drop-in boundary for ONE path of nosnoc/LCQPow: LCQProblem::runSolver() and the QP
 * subsolver under it (/root/reference/src/LCQProblem.cpp:444-560).  Plain pointers and sizes only;
 * no C++/torch types cross it; nothing throws across it.  Every entry point returns an int that is
 * either 0, a LCQPow::ReturnValue (/root/reference/include/Utilities.hpp:37-87) or a CUDA-side code
 * >= 500 (below).  All matrices are dense, row-major, fp64, exactly as LCQProblem::loadLCQP takes
 * them (/root/reference/include/LCQProblem.hpp:87-103).  Callers keep ownership of their buffers;
 * the library copies (the reference deep-copies too, LCQProblem.ipp:32-33).
 *
 * Two doors:
 *   (1) the batched front door  lcqp_cuda_create / _load / _run / _get_*   -- B independent LCQPs,
 *       the whole penalty-homotopy loop runs on the device (replaces LCQProblem::runSolver for a
 *       batch; LCQPow::LCQProblemBatch and LCQPow::LCQProblem in lcqpow_b200/host wrap it);
 *   (2) the plugin door  lcqp_cuda_qp_*  -- one convex QP with hot start, the exact shape of
 *       SubsolverBase::solve / getSolution (/root/reference/include/SubsolverBase.hpp:37-56); this is
 *       what LCQPow::SubsolverCUDA binds.
 *
 * There is no CPU fallback: every entry point that computes fails with LCQP_CUDA_NO_DEVICE when no
 * sm_100 device is usable.
 */
#ifndef LCQP_CUDA_H
#define LCQP_CUDA_H

#ifdef __cplusplus
extern "C" {
#endif

#define LCQP_CUDA_ABI_VERSION 2

/* CUDA-side return codes (LCQPow::ReturnValue stops at 402) */
#define LCQP_CUDA_OK 0
#define LCQP_CUDA_NO_DEVICE 500        /* no CUDA device / not sm_100 / driver error at create   */
#define LCQP_CUDA_BAD_HANDLE 501
#define LCQP_CUDA_BAD_ARGUMENT 502
#define LCQP_CUDA_OUT_OF_MEMORY 503
#define LCQP_CUDA_LAUNCH_FAILED 504    /* kernel launch / execution error (see _last_error)      */
#define LCQP_CUDA_NOT_LOADED 505       /* _run before _load                                     */
#define LCQP_CUDA_NOT_RUN 506          /* _get_* before _run                                    */
#define LCQP_CUDA_TOO_LARGE 507        /* problem does not fit the per-CTA shared-memory budget  */

/* bit k of a mask refers to argument k of loadLCQP:
 * 0 Q, 1 g, 2 L, 3 R, 4 lbL, 5 ubL, 6 lbR, 7 ubR, 8 A, 9 lbA, 10 ubA, 11 lb, 12 ub, 13 x0, 14 y0 */
enum {
    LCQP_Q = 0, LCQP_G, LCQP_L, LCQP_R, LCQP_LBL, LCQP_UBL, LCQP_LBR, LCQP_UBR,
    LCQP_A, LCQP_LBA, LCQP_UBA, LCQP_LB, LCQP_UB, LCQP_X0, LCQP_Y0, LCQP_NUM_ARRAYS
};

/* Mirrors LCQPow::Options (/root/reference/include/Options.hpp:192-213, defaults
 * /root/reference/src/Options.cpp:296-333) plus the knobs of the device QP solver. */
typedef struct {
    double stationarityTolerance;     /* 1e6 * EPS                                   */
    double complementarityTolerance;  /* 1e3 * EPS                                   */
    double initialPenaltyParameter;   /* 0.01                                        */
    double penaltyUpdateFactor;       /* 2                                           */
    double maxPenaltyParameter;       /* 1e8                                         */
    double etaDynamicPenalty;         /* 0.9                                         */
    int solveZeroPenaltyFirst;        /* 1                                           */
    int perturbStep;                  /* 1                                           */
    int maxIterations;                /* 1000                                        */
    int nDynamicPenalty;              /* 3  (<= 16)                                  */
    int qpSolver;                     /* dual layout: 0/1 qpOASES-style (nV box duals first), 2 OSQP-style */
    int osqp_admm;                    /* qpSolver == 2 only.  0: exact-vertex solver behind the OSQP dual layout;
                                         1: the OSQP restatement (ADMM + polish, lcqp_osqp.cuh) -- what QPSolver::OSQP_SPARSE means in the reference */
    double qp_rho;                    /* ADMM step of the phase-1 iteration, 0.1     */
    double qp_sigma;                  /* 1e-6                                        */
    double qp_alpha;                  /* 1.6                                         */
    double qp_delta;                  /* KKT regularisation of the EQP solves, 1e-6  */
    double qp_feas_tol;               /* KKT verification: bound violation / multiplier sign, 1e-12 (relative) */
    double qp_dual_tol;
    int qp_max_iter;                  /* ADMM iteration cap per QP, 4000             */
    int qp_check_interval;            /* ADMM iterations between active-set probes, 10 */
    int qp_refine_iter;               /* refinement passes per EQP solve, 10         */
    int qp_adaptive_rho;              /* reserved (0)                                */
    unsigned long long perturb_seed;  /* counter-based RNG key for perturbStep       */
    /* OSQPSettings of the OSQP restatement (/root/reference/external/osqp/include/types.h:158-195, defaults
     * constants.h:59-114 with LCQPow's overrides eps_prim_inf = EPS, polish = 1, /root/reference/src/Options.cpp:326-331) */
    double osqp_rho;                  /* 0.1    */
    double osqp_sigma;                /* 1e-6   */
    double osqp_alpha;                /* 1.6    */
    double osqp_delta;                /* 1e-6   */
    double osqp_eps_abs;              /* 1e-3   */
    double osqp_eps_rel;              /* 1e-3   */
    double osqp_eps_prim_inf;         /* 2.221e-16 (LCQPow) */
    double osqp_eps_dual_inf;         /* 1e-4   */
    double osqp_adaptive_rho_tolerance; /* 5    */
    int osqp_max_iter;                /* 4000   */
    int osqp_check_termination;       /* 25     */
    int osqp_scaling;                 /* 10 Ruiz passes */
    int osqp_adaptive_rho;            /* 1      */
    int osqp_adaptive_rho_interval;   /* 0: every 4 x check_termination iterations (the reference: wall-clock driven) */
    int osqp_polish;                  /* 1      */
    int osqp_polish_refine_iter;      /* 3      */
    int osqp_reserved;
    /* qpOASES::Options that the parametric active-set solver honours (Options::setqpOASESOptions,
     * /root/reference/src/Options.cpp:262-266; defaults /root/reference/external/qpOASES/src/Options.cpp:115,116) */
    double qpoases_terminationTolerance;   /* 5e6 * EPS: remaining relative homotopy length at which a QP ends */
    double qpoases_boundTolerance;         /* 1e6 * EPS: distance at which a bound counts as active (initial working set, far bounds) */
} lcqp_cuda_options;

/* Mirrors LCQPow::OutputStatistics counters (/root/reference/include/OutputStatistics.hpp:209-226),
 * one record per instance. */
typedef struct {
    int ret;             /* LCQPow::ReturnValue of runSolver for this instance */
    int status;          /* LCQPow::AlgorithmStatus                            */
    int iterTotal;
    int iterOuter;
    int subproblemIter;
    int qpExitFlag;
    int nDuals;
    int kktSolves;       /* regularised KKT solves (active-set passes) spent on this instance */
    double rhoOpt;
    double admmIters;    /* ADMM iterations spent on this instance (phase 1 of the first QP, fall-backs) */
} lcqp_cuda_stats;

typedef struct lcqp_cuda_handle_s* lcqp_cuda_handle;
typedef struct lcqp_cuda_qp_s* lcqp_cuda_qp;

int lcqp_cuda_abi_version(void);
void lcqp_cuda_default_options(lcqp_cuda_options* opts);          /* Options::setToDefault, Options.cpp:296 */

/* ---- (1) batched front door ------------------------------------------------------------------ */
/* LCQProblem::LCQProblem(nV,nC,nComp) (LCQProblem.cpp:43-79) for up to batch_capacity instances on
 * CUDA device `device`. */
int lcqp_cuda_create(int nV, int nC, int nComp, int batch_capacity, int device, lcqp_cuda_handle* out);
int lcqp_cuda_destroy(lcqp_cuda_handle h);
/* LCQProblem::setOptions (LCQProblem.ipp:160) */
int lcqp_cuda_set_options(lcqp_cuda_handle h, const lcqp_cuda_options* opts);
/* LCQProblem::loadLCQP (LCQProblem.cpp:87-144) for `batch` instances; HOST pointers; arrays whose
 * bit is set in shared_mask are read once, the others hold `batch` consecutive copies.  NULL = absent. */
int lcqp_cuda_load(lcqp_cuda_handle h, int batch, unsigned shared_mask,
                   const double* Q, const double* g, const double* L, const double* R,
                   const double* lbL, const double* ubL, const double* lbR, const double* ubR,
                   const double* A, const double* lbA, const double* ubA,
                   const double* lb, const double* ub, const double* x0, const double* y0);
/* Same, DEVICE pointers on the handle's device (no copy is made; they must stay valid until the
 * run completes). */
int lcqp_cuda_load_device(lcqp_cuda_handle h, int batch, unsigned shared_mask,
                          const double* Q, const double* g, const double* L, const double* R,
                          const double* lbL, const double* ubL, const double* lbR, const double* ubR,
                          const double* A, const double* lbA, const double* ubA,
                          const double* lb, const double* ub, const double* x0, const double* y0);
/* LCQProblem::loadLCQP(const csc* Q, g, const csc* L, const csc* R, ..., const csc* A, ...)
 * (/root/reference/src/LCQProblem.cpp:312-387; csc = /root/reference/external/osqp/include/types.h:21-29: column
 * pointers p[ncol+1], row indices i[nnz], values x[nnz]) for `batch` instances that share the sparsity PATTERNS:
 * Q is nV x nV (full symmetric), L and R are nComp x nV, A is nC x nV (NULL pointers when nC == 0).  The index
 * arrays are read once; a value array whose bit is set in shared_mask is read once, the others hold `batch`
 * consecutive copies of nnz values.  HOST pointers.  Serves qpSolver == 2 with osqp_admm == 1 (the OSQP flavour:
 * the reference's sparse door only exists for its sparse subsolvers); other settings return LCQP_CUDA_BAD_ARGUMENT. */
int lcqp_cuda_load_csc(lcqp_cuda_handle h, int batch, unsigned shared_mask,
                       const int* Q_p, const int* Q_i, const double* Q_x, const double* g,
                       const int* L_p, const int* L_i, const double* L_x,
                       const int* R_p, const int* R_i, const double* R_x,
                       const double* lbL, const double* ubL, const double* lbR, const double* ubR,
                       const int* A_p, const int* A_i, const double* A_x, const double* lbA, const double* ubA,
                       const double* x0, const double* y0);
/* Global index of instance 0 of this handle (keys the perturbStep RNG so that a batch sharded over
 * several GPUs draws the same perturbations as the unsharded batch).  Default 0. */
int lcqp_cuda_set_instance_offset(lcqp_cuda_handle h, unsigned long long offset);
/* LCQProblem::runSolver (LCQProblem.cpp:444-560) for every loaded instance.  `stream` is a
 * cudaStream_t (NULL = default stream); the call is asynchronous. */
int lcqp_cuda_run(lcqp_cuda_handle h, void* stream);
int lcqp_cuda_synchronize(lcqp_cuda_handle h);
/* LCQProblem::getPrimalSolution / getDualSolution / getOutputStatistics (LCQProblem.cpp:1485-1522):
 * x is batch x nV, y is batch x (nV+nC+2nComp) (first nDuals of each row valid), HOST buffers;
 * these synchronise. */
int lcqp_cuda_get_primal(lcqp_cuda_handle h, double* x);
int lcqp_cuda_get_dual(lcqp_cuda_handle h, double* y);
int lcqp_cuda_get_stats(lcqp_cuda_handle h, lcqp_cuda_stats* stats);
/* device-resident results (valid until the next _run / _destroy) */
int lcqp_cuda_get_device_results(lcqp_cuda_handle h, const double** x, const double** y, const lcqp_cuda_stats** stats);
int lcqp_cuda_num_duals(lcqp_cuda_handle h);                      /* LCQProblem::getNumberOfDuals */
/* kernels launched by this handle so far (bench.py's gpu_launches) and device time of the last run */
long long lcqp_cuda_launch_count(lcqp_cuda_handle h);
int lcqp_cuda_last_run_ms(lcqp_cuda_handle h, float* solve_kernel_ms, float* total_ms);
/* launch geometry of the last run: CTAs, dynamic shared memory per CTA, order of the static equality block */
int lcqp_cuda_last_launch_info(lcqp_cuda_handle h, int* grid, int* smem_bytes, int* equality_rows);
const char* lcqp_cuda_last_error(lcqp_cuda_handle h);
/* OSQP flavour: order N of the KKT system, non-zeros of its factor L and levels of a triangular solve (forward +
 * backward) of the loaded pattern; mode = 1 when the last run used one warp per instance, 0 for one thread per instance */
int lcqp_cuda_osqp_info(lcqp_cuda_handle h, int* N, int* nnzL, int* levels, int* mode);
/* fp64 FMA throughput of the device measured by a short probe kernel (TFLOP/s): the denominator bench.py uses for
 * the fp64 SIMT work of the solver */
int lcqp_cuda_measure_fp64_tflops(int device, double* tflops);
/* read bandwidth of the L2 (GB/s, 16-byte loads of a 48 MB buffer by every SM): the denominator for the solver's
 * L2-resident streams (the per-instance inverse of the working-set system and the shared operand Tt) */
int lcqp_cuda_measure_l2_gbs(int device, double* gbs);
/* work of the last run of the parametric active-set kernel, summed over the batch: fp64 multiply-adds of its dense
 * products (per-instance inverse, Tt columns, prepared operators) and the bytes those products read and write --
 * the algorithmic minimum, counted by the kernel itself.  Waits for the run.  Only the counting build of the library
 * (liblcqp_cuda_work.so, -DLCQP_COUNT_WORK) keeps these books -- they cost the kernel 3-5 % -- the shipped build returns
 * LCQP_CUDA_BAD_ARGUMENT. */
int lcqp_cuda_last_work(lcqp_cuda_handle h, double* fp64_macs, double* bytes);

/* ---- (2) plugin door: one convex QP, SubsolverBase semantics ---------------------------------- */
/* SubsolverQPOASES(int nV, int nC, double* Q, double* A) (SubsolverQPOASES.hpp:44-47): nCtot rows of
 * A_full = [A; L; R]. */
int lcqp_cuda_qp_create(int nV, int nCtot, const double* Q, const double* A, int device, lcqp_cuda_qp* out);
int lcqp_cuda_qp_destroy(lcqp_cuda_qp qp);
int lcqp_cuda_qp_set_options(lcqp_cuda_qp qp, const lcqp_cuda_options* opts);
/* SubsolverBase::solve (SubsolverBase.hpp:49-56).  Returns SUCCESSFUL_RETURN (0) or
 * SUBPROBLEM_SOLVER_ERROR (203). */
int lcqp_cuda_qp_solve(lcqp_cuda_qp qp, int initialSolve, int* iterations, int* exit_flag,
                       const double* g, const double* lbA, const double* ubA,
                       const double* x0, const double* y0, const double* lb, const double* ub);
/* SubsolverBase::getSolution (SubsolverBase.hpp:37): x[nV], y[nV + nCtot] (box duals first). */
int lcqp_cuda_qp_get_solution(lcqp_cuda_qp qp, double* x, double* y);

#ifdef __cplusplus
}
#endif
#endif /* LCQP_CUDA_H */

/*
 * oracle/lcqp_oracle.h -- TEST INFRASTRUCTURE ONLY (CPU restatement used as a checker).
 * See lcqp_oracle.c for what is restated from where.
 */
#ifndef LCQP_ORACLE_H
#define LCQP_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* First 12 fields are layout-compatible with lcqp_ref_options in ref_shim.cpp. */
typedef struct {
    double stationarityTolerance;
    double complementarityTolerance;
    double initialPenaltyParameter;
    double penaltyUpdateFactor;
    double maxPenaltyParameter;
    double etaDynamicPenalty;
    int solveZeroPenaltyFirst;
    int perturbStep;
    int maxIterations;
    int nDynamicPenalty;
    int qpSolver;      /* 0/1: qpOASES-flavour dual layout (box duals first); 2: OSQP-flavour (no box) */
    int reserved0;
    /* inner exact-QP solver (ADMM + KKT-verified polish) */
    double qp_rho;            /* ADMM step (OSQP rho, constants.h:45)            default 0.1  */
    double qp_sigma;          /* ADMM sigma (constants.h:46)                     default 1e-6 */
    double qp_alpha;          /* relaxation (constants.h:58)                     default 1.6  */
    double qp_delta;          /* polish regularisation (constants.h:76)          default 1e-6 */
    double qp_feas_tol;       /* KKT verification: primal feasibility (scaled)   default 1e-12 */
    double qp_dual_tol;       /* KKT verification: dual sign (scaled)            default 1e-14 */
    int qp_max_iter;          /* ADMM iteration cap per QP (constants.h:61)      default 4000 */
    int qp_check_interval;    /* ADMM iterations between active-set probes       default 10   */
    int qp_refine_iter;       /* max refinement passes per polish                default 10   */
    int qp_adaptive_rho;      /* 1: move along the rho ladder (x5 steps)         default 0    */
    unsigned long long perturb_seed; /* counter-based RNG key for perturbStep     default 1    */
} lcqp_oracle_options;

typedef struct {
    int ret;
    int status;
    int iterTotal;
    int iterOuter;
    int subproblemIter;
    int qpExitFlag;
    int nDuals;
    int pad;
    double rhoOpt;
    double seconds;
} lcqp_oracle_result;

void lcqp_oracle_default_options(lcqp_oracle_options* o);

int lcqp_oracle_solve_batch(int batch, int nV, int nC, int nComp, unsigned shared_mask,
                            const double* Q, const double* g, const double* L, const double* R,
                            const double* lbL, const double* ubL, const double* lbR, const double* ubR,
                            const double* A, const double* lbA, const double* ubA,
                            const double* lb, const double* ub,
                            const double* x0, const double* y0,
                            const lcqp_oracle_options* o, double* x, double* y, lcqp_oracle_result* res);

/* Restated Utilities kernels (/root/reference/src/Utilities.cpp:38-265), exported for the KATs of
 * /root/reference/test/RunUnitTests.cpp:33-246. */
void lcqp_oracle_MatrixMultiplication(const double* A, const double* B, double* C, int m, int n, int p);
void lcqp_oracle_TransponsedMatrixMultiplication(const double* A, const double* B, double* C, int m, int n, int p);
void lcqp_oracle_AddTransponsedMatrixMultiplication(const double* A, const double* B, double* C, int m, int n, int p);
void lcqp_oracle_MatrixSymmetrizationProduct(const double* A, const double* B, double* C, int m, int n);
void lcqp_oracle_AffineLinearTransformation(double alpha, const double* A, const double* b, const double* c, double* d, int m, int n);
void lcqp_oracle_WeightedMatrixAdd(double alpha, const double* A, double beta, const double* B, double* C, int m, int n);
void lcqp_oracle_WeightedVectorAdd(double alpha, const double* a, double beta, const double* b, double* c, int m);
double lcqp_oracle_QuadraticFormProduct(const double* Q, const double* p, int m);
double lcqp_oracle_DotProduct(const double* a, const double* b, int m);
double lcqp_oracle_MaxAbs(const double* a, int m);
int lcqp_oracle_perturb_draw(unsigned long long seed, unsigned long long instance, unsigned iter, unsigned i);

#ifdef __cplusplus
}
#endif
#endif

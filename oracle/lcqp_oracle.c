/*
 * oracle/lcqp_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C CPU restatement of the hot path of nosnoc/LCQPow, used by tests/, by
 * __graft_entry__.smoke() and by bench.py's cpu_baseline leg as the CHECKER of the CUDA path.
 * It is never linked into, imported by or called from the product (lcqpow_b200/).
 *
 * What is restated from where (all paths under /root/reference):
 *   - the penalty-homotopy loop        src/LCQProblem.cpp:444-560   -> lcqp_run()
 *   - initializeSolver                 src/LCQProblem.cpp:885-1034  -> lcqp_run() prologue
 *   - setConstraints/ComplementarityB. src/LCQProblem.cpp:563-626, 726-785 -> build_constraints()
 *   - updateLinearization .. determineStationarityType  src/LCQProblem.cpp:1105-1482
 *   - the dense Utilities kernels      src/Utilities.cpp:38-265     -> lcqp_oracle_<name>()
 *   - defaults                         src/Options.cpp:296-333      -> lcqp_oracle_default_options()
 *
 * The convex QP under the loop.  The reference delegates it to qpOASES 3.2 (exact active-set
 * optimum, external/qpOASES/src/QProblem.cpp:316-645) or OSQP 0.6.2 (ADMM + polish,
 * external/osqp/src/osqp.c:288-641).  The contract the loop relies on is "the exact
 * optimum x*, the multipliers y* in qpOASES' sign convention, warm/hot start from the previous
 * solution".  The restatement meets that contract with ONE solver used by both dual layouts:
 *
 *     OSQP's ADMM iteration (auxil.c:161-225, in condensed form), OSQP's Ruiz equilibration
 *     (scaling.c:44-156, without the q-dependent cost term so that batches sharing Q/A share it),
 *     OSQP's active-set guess and regularised-KKT polish with iterative refinement
 *     (polish.c:33-49, :134-181, :232-300) -- but the polish result is only accepted when it
 *     satisfies the KKT conditions of the QP to 1e-9 (primal feasibility of the inactive rows,
 *     sign of the active multipliers), i.e. when it IS the exact optimum; otherwise ADMM goes on.
 *     On later calls the previous working set is tried first (hot start, zero ADMM iterations when
 *     the active set did not change -- the analogue of qpOASES' nWSR = 0 hotstart).
 *
 * Because the accepted QP solution is the exact optimum, the outer trajectory (iterOuter,
 * iterTotal, rhoOpt, stationarity type, x) is that of the reference's qpOASES runs; this is pinned
 * against the real reference (oracle/_ref) in tests/test_oracle_vs_reference.py and against the
 * committed golden vectors in tests/golden/.
 */
#define _POSIX_C_SOURCE 200809L
#include "lcqp_oracle.h"

#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define LCQ_EPS 2.221e-16 /* Utilities.hpp:350 */
#define QP_INF 1e20       /* Utilities.hpp:362 (INFTY); anything beyond is "no bound" */
#define RHO_MIN 1e-6      /* osqp constants.h:48 */
#define RHO_MAX 1e6
#define RHO_TOL 1e-4            /* constants.h:50 */
#define RHO_EQ_OVER_INEQ 1e3    /* constants.h:51 */
#define MIN_SCALING 1e-4        /* constants.h:94 */
#define MAX_SCALING 1e4
#define SCALING_ITERS 10        /* constants.h:56 */
#define RES_TOL 1e-9            /* residual of an EQP solve that still counts as solved */

enum { RET_OK = 0, RET_INVALID_OSQP_BOX = 110, RET_INVALID_LOWER_COMP = 120, RET_MAX_ITER = 200, RET_MAX_PEN = 201,
       RET_SUBPROBLEM = 203, RET_OSQP_GUESS = 208 };

/* ================================================================================================
 * Utilities.cpp restated
 * ==============================================================================================*/
/* Utilities.cpp:38-47 */
void lcqp_oracle_MatrixMultiplication(const double* A, const double* B, double* C, int m, int n, int p)
{
    for (int i = 0; i < m; i++)
        for (int j = 0; j < p; j++) {
            double s = 0;
            for (int k = 0; k < n; k++) s += A[i * n + k] * B[k * p + j];
            C[i * p + j] = s;
        }
}

/* Utilities.cpp:62-72 */
void lcqp_oracle_TransponsedMatrixMultiplication(const double* A, const double* B, double* C, int m, int n, int p)
{
    for (int i = 0; i < n; i++)
        for (int j = 0; j < p; j++) {
            double s = 0;
            for (int k = 0; k < m; k++) s += A[k * n + i] * B[k * p + j];
            C[i * p + j] = s;
        }
}

/* Utilities.cpp:85-93 */
void lcqp_oracle_AddTransponsedMatrixMultiplication(const double* A, const double* B, double* C, int m, int n, int p)
{
    for (int i = 0; i < n; i++)
        for (int j = 0; j < p; j++)
            for (int k = 0; k < m; k++) C[i * p + j] += A[k * n + i] * B[k * p + j];
}

/* Utilities.cpp:104-116: C = A'B + B'A, lower triangle computed, mirrored */
void lcqp_oracle_MatrixSymmetrizationProduct(const double* A, const double* B, double* C, int m, int n)
{
    for (int i = 0; i < n; i++)
        for (int j = 0; j <= i; j++) {
            double s = 0;
            for (int k = 0; k < m; k++) s += A[k * n + i] * B[k * n + j] + B[k * n + i] * A[k * n + j];
            C[i * n + j] = s;
            C[j * n + i] = s;
        }
}

/* Utilities.cpp:176-186: d = alpha*A*b + c */
void lcqp_oracle_AffineLinearTransformation(double alpha, const double* A, const double* b, const double* c, double* d, int m, int n)
{
    for (int i = 0; i < m; i++) {
        double s = 0;
        for (int k = 0; k < n; k++) s += A[i * n + k] * b[k];
        d[i] = alpha * s + c[i];
    }
}

/* Utilities.cpp:202-206 */
void lcqp_oracle_WeightedMatrixAdd(double alpha, const double* A, double beta, const double* B, double* C, int m, int n)
{
    for (int i = 0; i < m * n; i++) C[i] = alpha * A[i] + beta * B[i];
}

/* Utilities.cpp:209-211 */
void lcqp_oracle_WeightedVectorAdd(double alpha, const double* a, double beta, const double* b, double* c, int m)
{
    lcqp_oracle_WeightedMatrixAdd(alpha, a, beta, b, c, m, 1);
}

/* Utilities.cpp:214-225 */
double lcqp_oracle_QuadraticFormProduct(const double* Q, const double* p, int m)
{
    double ret = 0;
    for (int i = 0; i < m; i++) {
        double s = 0;
        for (int j = 0; j < m; j++) s += Q[i * m + j] * p[j];
        ret += s * p[i];
    }
    return ret;
}

/* Utilities.cpp:244-250 */
double lcqp_oracle_DotProduct(const double* a, const double* b, int m)
{
    double r = 0;
    for (int i = 0; i < m; i++) r += a[i] * b[i];
    return r;
}

/* Utilities.cpp:253-265: despite the doc string this is the infinity norm */
double lcqp_oracle_MaxAbs(const double* a, int m)
{
    double mx = 0, mn = 0;
    for (int i = 0; i < m; i++) {
        if (a[i] > mx) mx = a[i];
        else if (a[i] < mn) mn = a[i];
    }
    return mx > -mn ? mx : -mn;
}

/* perturbStep (LCQProblem.cpp:1353-1362) draws rand()%3-1 from a time-seeded libc stream; the
 * restatement (and the CUDA path) use a counter-based generator keyed by (seed, instance, iterate,
 * coordinate) so that a run is reproducible.  splitmix64 finaliser. */
int lcqp_oracle_perturb_draw(unsigned long long seed, unsigned long long instance, unsigned iter, unsigned i)
{
    uint64_t z = seed * 0x9E3779B97F4A7C15ull + instance * 0xBF58476D1CE4E5B9ull + ((uint64_t)iter << 32 | i);
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    return (int)(z % 3ull) - 1;
}

/* ================================================================================================
 * small dense linear algebra
 * ==============================================================================================*/
static void* xcalloc(size_t n, size_t s)
{
    void* p = calloc(n ? n : 1, s);
    if (!p) { fprintf(stderr, "lcqp_oracle: out of memory\n"); abort(); }
    return p;
}

/* in-place lower Cholesky of the n x n row-major SPD matrix M (leading dimension ld);
 * returns 0 on success, 1 if a pivot is not positive */
static int chol_lower(double* M, int n, int ld)
{
    for (int j = 0; j < n; j++) {
        double d = M[j * ld + j];
        for (int k = 0; k < j; k++) d -= M[j * ld + k] * M[j * ld + k];
        if (!(d > 0)) return 1;
        d = sqrt(d);
        M[j * ld + j] = d;
        for (int i = j + 1; i < n; i++) {
            double s = M[i * ld + j];
            for (int k = 0; k < j; k++) s -= M[i * ld + k] * M[j * ld + k];
            M[i * ld + j] = s / d;
        }
    }
    return 0;
}

/* solve L L' x = b in place (L lower, row-major, leading dimension ld) */
static void chol_solve(const double* Lm, int n, int ld, double* b)
{
    for (int i = 0; i < n; i++) {
        double s = b[i];
        for (int k = 0; k < i; k++) s -= Lm[i * ld + k] * b[k];
        b[i] = s / Lm[i * ld + i];
    }
    for (int i = n - 1; i >= 0; i--) {
        double s = b[i];
        for (int k = i + 1; k < n; k++) s -= Lm[k * ld + i] * b[k];
        b[i] = s / Lm[i * ld + i];
    }
}

/* Minv = M^{-1} for SPD M (n x n); M is overwritten by its Cholesky factor. */
static int spd_inverse(double* M, double* Minv, int n)
{
    if (chol_lower(M, n, n)) return 1;
    double* col = (double*)xcalloc((size_t)n, sizeof(double));
    for (int j = 0; j < n; j++) {
        memset(col, 0, (size_t)n * sizeof(double));
        col[j] = 1.0;
        chol_solve(M, n, n, col);
        for (int i = 0; i < n; i++) Minv[i * n + j] = col[i];
    }
    /* symmetrise (rounding) */
    for (int i = 0; i < n; i++)
        for (int j = 0; j < i; j++) {
            double v = 0.5 * (Minv[i * n + j] + Minv[j * n + i]);
            Minv[i * n + j] = v;
            Minv[j * n + i] = v;
        }
    free(col);
    return 0;
}

static void matvec(const double* M, const double* v, double* out, int rows, int cols)
{
    for (int i = 0; i < rows; i++) {
        double s = 0;
        const double* r = M + (size_t)i * cols;
        for (int k = 0; k < cols; k++) s += r[k] * v[k];
        out[i] = s;
    }
}

/* ================================================================================================
 * The exact convex QP solver (ADMM + KKT-verified polish)
 *     min 1/2 x'Px + q'x   s.t.  l <= Ahat x <= u,    Ahat = [A_full ; I (only when box bounds exist)]
 * ==============================================================================================*/
#define MAX_MINV_CACHE 24

typedef struct {
    double rho;
    int* ctype;
    double* Minv;
} minv_entry;

/* everything that depends on (Q, A_full, has_box) only: shareable across a batch */
typedef struct {
    int n, mA, m, has_box;
    double sigma, delta;
    double* P;    /* n*n   scaled  c D P D            */
    double* A;    /* m*n   scaled  E Ahat D           */
    int* Ap;      /* CSR of scaled A (row pointers)   */
    int* Aj;
    double* Ax;
    double* D;    /* n */
    double* E;    /* m */
    double c;
    double* Hinv; /* n*n  (P + delta I)^-1            */
    double* G;    /* m*m  A Hinv A'                   */
    minv_entry cache[MAX_MINV_CACHE];
    int ncache;
} qp_mats;

typedef struct {
    qp_mats* M;
    const lcqp_oracle_options* o;
    double rho;
    double *l, *u;       /* scaled bounds, m */
    int* ctype;          /* -1 free row, 0 inequality, 1 equality (auxil.c:76-98) */
    double* rho_vec;
    const double* Minv;  /* points into the cache */
    double *x, *z, *y;   /* ADMM iterate, scaled, OSQP sign */
    double* q;           /* scaled linear term */
    int* W;              /* working set of the accepted solution: 0 inactive, 1 at lower, 2 at upper */
    int* Wtry;
    int* Wfail;
    int* pin;            /* rows exempt from the multiplier sign test (anti-cycling) */
    int have_W, have_fail;
    double* xs;          /* accepted solution, unscaled (n) */
    double* ys;          /* accepted multipliers, unscaled, qpOASES sign (m) */
    /* scratch */
    double *w, *rhs, *xt, *zt, *zt2, *px, *lam, *t1, *t2, *S, *r1, *r2, *dl, *dx, *xe, *xa;
    double eqp_res;
    int* idx;
    int infeasible_bounds;
    long total_admm, total_polish;
} qp_inst;

static void csr_matvec(const qp_mats* M, const double* v, double* out)
{
    for (int i = 0; i < M->m; i++) {
        double s = 0;
        for (int k = M->Ap[i]; k < M->Ap[i + 1]; k++) s += M->Ax[k] * v[M->Aj[k]];
        out[i] = s;
    }
}

static void csr_tmatvec_add(const qp_mats* M, const double* w, double* out /* n, accumulated */)
{
    for (int i = 0; i < M->m; i++) {
        double wi = w[i];
        if (wi == 0.0) continue;
        for (int k = M->Ap[i]; k < M->Ap[i + 1]; k++) out[M->Aj[k]] += M->Ax[k] * wi;
    }
}

static void qp_mats_free(qp_mats* M)
{
    if (!M) return;
    free(M->P); free(M->A); free(M->Ap); free(M->Aj); free(M->Ax); free(M->D); free(M->E); free(M->Hinv); free(M->G);
    for (int i = 0; i < M->ncache; i++) { free(M->cache[i].ctype); free(M->cache[i].Minv); }
    free(M);
}

static double limit_scaling(double v)
{
    if (v < MIN_SCALING) return 1.0; /* scaling.c:22-30 */
    if (v > MAX_SCALING) return MAX_SCALING;
    return v;
}

/* Build the shareable part: scaled matrices (Ruiz, scaling.c:44-156 without the q term), Hinv, G. */
static qp_mats* qp_mats_create(int n, int mA, int has_box, const double* Q, const double* Afull, double sigma, double delta)
{
    qp_mats* M = (qp_mats*)xcalloc(1, sizeof(qp_mats));
    const int m = mA + (has_box ? n : 0);
    M->n = n; M->mA = mA; M->m = m; M->has_box = has_box; M->sigma = sigma; M->delta = delta;
    M->P = (double*)xcalloc((size_t)n * n, sizeof(double));
    M->A = (double*)xcalloc((size_t)m * n, sizeof(double));
    M->D = (double*)xcalloc((size_t)n, sizeof(double));
    M->E = (double*)xcalloc((size_t)m, sizeof(double));
    memcpy(M->P, Q, (size_t)n * n * sizeof(double));
    if (mA) memcpy(M->A, Afull, (size_t)mA * n * sizeof(double));
    if (has_box)
        for (int i = 0; i < n; i++) M->A[(size_t)(mA + i) * n + i] = 1.0;
    for (int i = 0; i < n; i++) M->D[i] = 1.0;
    for (int i = 0; i < m; i++) M->E[i] = 1.0;
    M->c = 1.0;

    double* Dt = (double*)xcalloc((size_t)n, sizeof(double));
    double* Et = (double*)xcalloc((size_t)m, sizeof(double));
    for (int it = 0; it < SCALING_ITERS; it++) {
        for (int j = 0; j < n; j++) {
            double v = 0;
            for (int i = 0; i < n; i++) { double a = fabs(M->P[(size_t)i * n + j]); if (a > v) v = a; }
            for (int i = 0; i < m; i++) { double a = fabs(M->A[(size_t)i * n + j]); if (a > v) v = a; }
            Dt[j] = 1.0 / sqrt(limit_scaling(v));
        }
        for (int i = 0; i < m; i++) {
            double v = 0;
            for (int j = 0; j < n; j++) { double a = fabs(M->A[(size_t)i * n + j]); if (a > v) v = a; }
            Et[i] = 1.0 / sqrt(limit_scaling(v));
        }
        for (int i = 0; i < n; i++)
            for (int j = 0; j < n; j++) M->P[(size_t)i * n + j] *= Dt[i] * Dt[j];
        for (int i = 0; i < m; i++)
            for (int j = 0; j < n; j++) M->A[(size_t)i * n + j] *= Et[i] * Dt[j];
        for (int j = 0; j < n; j++) M->D[j] *= Dt[j];
        for (int i = 0; i < m; i++) M->E[i] *= Et[i];
    }
    /* OSQP's cost normalisation c = 1/max(mean col-norm P, ||q||) (scaling.c:114-141) depends on q, which
     * would make the scaled KKT matrix instance-specific; the restatement keeps c = 1 and lets the rho
     * ladder absorb the cost scale (ADMM with (cP, cq, rho) is ADMM with (P, q, rho/c)). */
    free(Dt); free(Et);

    /* CSR copy of the scaled A for the ADMM mat-vecs */
    int nnz = 0;
    for (size_t k = 0; k < (size_t)m * n; k++) nnz += (M->A[k] != 0.0);
    M->Ap = (int*)xcalloc((size_t)m + 1, sizeof(int));
    M->Aj = (int*)xcalloc((size_t)nnz, sizeof(int));
    M->Ax = (double*)xcalloc((size_t)nnz, sizeof(double));
    int p = 0;
    for (int i = 0; i < m; i++) {
        M->Ap[i] = p;
        for (int j = 0; j < n; j++)
            if (M->A[(size_t)i * n + j] != 0.0) { M->Aj[p] = j; M->Ax[p] = M->A[(size_t)i * n + j]; p++; }
    }
    M->Ap[m] = p;

    /* Hinv = (P + delta I)^-1,  G = A Hinv A' */
    double* H = (double*)xcalloc((size_t)n * n, sizeof(double));
    memcpy(H, M->P, (size_t)n * n * sizeof(double));
    for (int i = 0; i < n; i++) H[(size_t)i * n + i] += delta;
    M->Hinv = (double*)xcalloc((size_t)n * n, sizeof(double));
    if (spd_inverse(H, M->Hinv, n)) { free(H); qp_mats_free(M); return NULL; }
    free(H);
    double* AH = (double*)xcalloc((size_t)m * n, sizeof(double)); /* A Hinv */
    for (int i = 0; i < m; i++)
        for (int k = M->Ap[i]; k < M->Ap[i + 1]; k++) {
            const double a = M->Ax[k];
            const double* hr = M->Hinv + (size_t)M->Aj[k] * n;
            double* out = AH + (size_t)i * n;
            for (int j = 0; j < n; j++) out[j] += a * hr[j];
        }
    M->G = (double*)xcalloc((size_t)m * m, sizeof(double));
    for (int i = 0; i < m; i++)
        for (int r = 0; r <= i; r++) {
            double s = 0;
            for (int k = M->Ap[r]; k < M->Ap[r + 1]; k++) s += AH[(size_t)i * n + M->Aj[k]] * M->Ax[k];
            M->G[(size_t)i * m + r] = s;
            M->G[(size_t)r * m + i] = s;
        }
    free(AH);
    return M;
}

/* M = P + sigma I + A' diag(rho_vec) A, inverted; cached per (rho, ctype) */
static const double* qp_get_minv(qp_mats* M, double rho, const int* ctype, const double* rho_vec)
{
    for (int e = 0; e < M->ncache; e++)
        if (M->cache[e].rho == rho && memcmp(M->cache[e].ctype, ctype, (size_t)M->m * sizeof(int)) == 0) return M->cache[e].Minv;
    const int n = M->n, m = M->m;
    double* K = (double*)xcalloc((size_t)n * n, sizeof(double));
    memcpy(K, M->P, (size_t)n * n * sizeof(double));
    for (int i = 0; i < n; i++) K[(size_t)i * n + i] += M->sigma;
    for (int r = 0; r < m; r++)
        for (int a = M->Ap[r]; a < M->Ap[r + 1]; a++)
            for (int b = M->Ap[r]; b < M->Ap[r + 1]; b++) K[(size_t)M->Aj[a] * n + M->Aj[b]] += rho_vec[r] * M->Ax[a] * M->Ax[b];
    double* Minv = (double*)xcalloc((size_t)n * n, sizeof(double));
    if (spd_inverse(K, Minv, n)) { free(K); free(Minv); return NULL; }
    free(K);
    int e = M->ncache < MAX_MINV_CACHE ? M->ncache++ : MAX_MINV_CACHE - 1;
    if (e == MAX_MINV_CACHE - 1 && M->cache[e].Minv) { free(M->cache[e].Minv); free(M->cache[e].ctype); }
    M->cache[e].rho = rho;
    M->cache[e].ctype = (int*)xcalloc((size_t)m, sizeof(int));
    memcpy(M->cache[e].ctype, ctype, (size_t)m * sizeof(int));
    M->cache[e].Minv = Minv;
    return Minv;
}

static void qp_set_rho(qp_inst* q, double rho)
{
    const int m = q->M->m;
    q->rho = rho;
    for (int i = 0; i < m; i++) /* auxil.c:76-98 */
        q->rho_vec[i] = q->ctype[i] < 0 ? RHO_MIN : (q->ctype[i] == 1 ? RHO_EQ_OVER_INEQ * rho : rho);
    q->Minv = qp_get_minv(q->M, rho, q->ctype, q->rho_vec);
}

static void qp_inst_free(qp_inst* q)
{
    if (!q) return;
    free(q->l); free(q->u); free(q->ctype); free(q->rho_vec); free(q->x); free(q->z); free(q->y); free(q->q);
    free(q->W); free(q->Wtry); free(q->Wfail); free(q->pin); free(q->xs); free(q->ys);
    free(q->w); free(q->rhs); free(q->xt); free(q->zt); free(q->px); free(q->lam); free(q->t1); free(q->t2);
    free(q->zt2); free(q->xe); free(q->xa); free(q->S); free(q->r1); free(q->r2); free(q->dl); free(q->dx); free(q->idx);
    free(q);
}

/* lbA/ubA are the mA bounds of A_full; lb/ub (n) only when has_box */
static qp_inst* qp_inst_create(qp_mats* M, const lcqp_oracle_options* o, const double* lbA, const double* ubA, const double* lb, const double* ub)
{
    const int n = M->n, m = M->m, mA = M->mA;
    qp_inst* q = (qp_inst*)xcalloc(1, sizeof(qp_inst));
    q->M = M; q->o = o;
    q->l = (double*)xcalloc((size_t)m, sizeof(double));
    q->u = (double*)xcalloc((size_t)m, sizeof(double));
    q->ctype = (int*)xcalloc((size_t)m, sizeof(int));
    q->rho_vec = (double*)xcalloc((size_t)m, sizeof(double));
    q->x = (double*)xcalloc((size_t)n, sizeof(double));
    q->z = (double*)xcalloc((size_t)m, sizeof(double));
    q->y = (double*)xcalloc((size_t)m, sizeof(double));
    q->q = (double*)xcalloc((size_t)n, sizeof(double));
    q->W = (int*)xcalloc((size_t)m, sizeof(int));
    q->Wtry = (int*)xcalloc((size_t)m, sizeof(int));
    q->Wfail = (int*)xcalloc((size_t)m, sizeof(int));
    q->pin = (int*)xcalloc((size_t)m, sizeof(int));
    q->xs = (double*)xcalloc((size_t)n, sizeof(double));
    q->ys = (double*)xcalloc((size_t)m, sizeof(double));
    q->w = (double*)xcalloc((size_t)m, sizeof(double));
    q->rhs = (double*)xcalloc((size_t)n, sizeof(double));
    q->xt = (double*)xcalloc((size_t)n, sizeof(double));
    q->zt = (double*)xcalloc((size_t)m, sizeof(double));
    q->zt2 = (double*)xcalloc((size_t)m, sizeof(double));
    q->xe = (double*)xcalloc((size_t)n, sizeof(double));
    q->xa = (double*)xcalloc((size_t)n, sizeof(double));
    q->px = (double*)xcalloc((size_t)n, sizeof(double));
    q->lam = (double*)xcalloc((size_t)m, sizeof(double));
    q->t1 = (double*)xcalloc((size_t)(n > m ? n : m), sizeof(double));
    q->t2 = (double*)xcalloc((size_t)(n > m ? n : m), sizeof(double));
    q->S = (double*)xcalloc((size_t)m * m, sizeof(double));
    q->r1 = (double*)xcalloc((size_t)n, sizeof(double));
    q->r2 = (double*)xcalloc((size_t)m, sizeof(double));
    q->dl = (double*)xcalloc((size_t)m, sizeof(double));
    q->dx = (double*)xcalloc((size_t)n, sizeof(double));
    q->idx = (int*)xcalloc((size_t)m, sizeof(int));
    for (int i = 0; i < m; i++) {
        double lo = i < mA ? lbA[i] : lb[i - mA];
        double up = i < mA ? ubA[i] : ub[i - mA];
        if (lo > up) q->infeasible_bounds = 1;
        int linf = !(lo > -QP_INF), uinf = !(up < QP_INF);
        q->l[i] = linf ? -INFINITY : lo * M->E[i];
        q->u[i] = uinf ? INFINITY : up * M->E[i];
        if (linf && uinf) q->ctype[i] = -1;
        else if (!linf && !uinf && q->u[i] - q->l[i] < RHO_TOL) q->ctype[i] = 1;
        else q->ctype[i] = 0;
    }
    qp_set_rho(q, o->qp_rho);
    return q;
}

/* One ADMM iteration, OSQP auxil.c:161-225 with the KKT solve condensed to the n x n system
 * (P + sigma I + A' R A) xt = sigma x - q + A'(R z - y),  zt = A xt. */
static void admm_iter(qp_inst* q)
{
    const qp_mats* M = q->M;
    const int n = M->n, m = M->m;
    const double alpha = q->o->qp_alpha, sigma = M->sigma;
    for (int i = 0; i < m; i++) q->w[i] = q->rho_vec[i] * q->z[i] - q->y[i];
    for (int j = 0; j < n; j++) q->rhs[j] = sigma * q->x[j] - q->q[j];
    csr_tmatvec_add(M, q->w, q->rhs);
    matvec(q->Minv, q->rhs, q->xt, n, n);
    csr_matvec(M, q->xt, q->zt);
    for (int j = 0; j < n; j++) q->x[j] = alpha * q->xt[j] + (1.0 - alpha) * q->x[j];
    for (int i = 0; i < m; i++) {
        double v = alpha * q->zt[i] + (1.0 - alpha) * q->z[i];
        double zn = v + q->y[i] / q->rho_vec[i];
        if (zn < q->l[i]) zn = q->l[i];
        if (zn > q->u[i]) zn = q->u[i];
        q->y[i] += q->rho_vec[i] * (v - zn);
        q->z[i] = zn;
    }
}

/* Active-set guess from the ADMM iterate, polish.c:33-49 (equality rows are always active). */
static void guess_working_set(const qp_inst* q, int* W)
{
    const int m = q->M->m;
    for (int i = 0; i < m; i++) {
        if (q->ctype[i] == 1) W[i] = 1;
        else if (q->ctype[i] < 0) W[i] = 0;
        else if (q->z[i] - q->l[i] < -q->y[i]) W[i] = 1;
        else if (q->u[i] - q->z[i] < q->y[i]) W[i] = 2;
        else W[i] = 0;
    }
}

/* Solve the equality-constrained QP on working set W with the regularised KKT system
 *   [P + dI, Aw'; Aw, -dI] [x; lam] = [-q; b]     (polish.c:232-300)
 * by block elimination (Hinv, S = G[W,W] + dI) and iterative refinement against the unregularised
 * system (polish.c:134-181).  Then verify the KKT conditions of the full QP.  On success the scaled
 * point is left in q->xt (x) / q->lam (multipliers, OSQP sign, zero on inactive rows) and 1 is
 * returned. */
static int g_dbg = -1;
static int dbg(void)
{
    if (g_dbg < 0) g_dbg = getenv("LCQP_ORACLE_DEBUG") ? atoi(getenv("LCQP_ORACLE_DEBUG")) : 0;
    return g_dbg;
}

/* Equality-constrained QP on working set W:  regularised KKT system
 *   [P + dI, Aw'; Aw, -dI] [x; lam] = [-q; b]            (OSQP polish.c:232-300)
 * solved by block elimination -- Hinv = (P+dI)^-1 and S = G[W,W] + dI with G = A Hinv A' are
 * working-set independent except for the gather -- followed by iterative refinement against the
 * UNREGULARISED system (polish.c:134-181); the iterate with the smallest residual is kept.
 * Output: q->xe (n), q->lam (compact, nw entries in the order of q->idx), returns nw or -1. */
static int eqp_solve(qp_inst* q, const int* W)
{
    const qp_mats* M = q->M;
    const int n = M->n, m = M->m;
    const double delta = M->delta;
    int nw = 0;
    for (int i = 0; i < m; i++)
        if (W[i]) q->idx[nw++] = i;
    q->total_polish++;

    double* S = q->S;
    for (int a = 0; a < nw; a++) {
        const double* gr = M->G + (size_t)q->idx[a] * m;
        for (int b = 0; b <= a; b++) S[(size_t)a * nw + b] = gr[q->idx[b]];
        S[(size_t)a * nw + a] += delta;
    }
    if (nw && chol_lower(S, nw, nw)) return -1;

    double* x = q->xe;
    double* lam = q->lam;
    memset(x, 0, (size_t)n * sizeof(double));
    memset(lam, 0, (size_t)m * sizeof(double));
    double best = INFINITY;
    for (int pass = 0;; pass++) {
        /* residual of the unregularised system at (x, lam) */
        matvec(M->P, x, q->px, n, n);
        for (int j = 0; j < n; j++) q->r1[j] = -q->q[j] - q->px[j];
        for (int a = 0; a < nw; a++) {
            const int i = q->idx[a];
            const double la = lam[a];
            double s = 0;
            for (int k = M->Ap[i]; k < M->Ap[i + 1]; k++) {
                q->r1[M->Aj[k]] -= M->Ax[k] * la;
                s += M->Ax[k] * x[M->Aj[k]];
            }
            q->r2[a] = (W[i] == 1 ? q->l[i] : q->u[i]) - s;
        }
        double rn = 0;
        for (int j = 0; j < n; j++) if (fabs(q->r1[j]) > rn) rn = fabs(q->r1[j]);
        for (int a = 0; a < nw; a++) if (fabs(q->r2[a]) > rn) rn = fabs(q->r2[a]);
        if (pass > 0 && !(rn < best)) { /* last correction did not help: undo it and stop */
            for (int j = 0; j < n; j++) x[j] -= q->dx[j];
            for (int a = 0; a < nw; a++) lam[a] -= q->dl[a];
            break;
        }
        /* a correction that barely helped: the residual lies along directions whose curvature is far below
         * delta (the refinement contracts by delta/(c+delta) there); stretch the next one (below) */
        const int slow = pass > 0 && rn > 0.25 * best && rn > 1e-13;
        best = rn;
        if (rn < 1e-15 || pass == q->o->qp_refine_iter) break;
        /* dlam = S^-1 (Aw Hinv r1 - r2);  dx = Hinv (r1 - Aw' dlam) */
        matvec(M->Hinv, q->r1, q->t1, n, n);
        for (int a = 0; a < nw; a++) {
            const int i = q->idx[a];
            double s = 0;
            for (int k = M->Ap[i]; k < M->Ap[i + 1]; k++) s += M->Ax[k] * q->t1[M->Aj[k]];
            q->dl[a] = s - q->r2[a];
        }
        if (nw) chol_solve(S, nw, nw, q->dl);
        memcpy(q->t2, q->r1, (size_t)n * sizeof(double));
        for (int a = 0; a < nw; a++) {
            const int i = q->idx[a];
            for (int k = M->Ap[i]; k < M->Ap[i + 1]; k++) q->t2[M->Aj[k]] -= M->Ax[k] * q->dl[a];
        }
        matvec(M->Hinv, q->t2, q->dx, n, n);
        if (slow) {
            /* exact line search along dx (the rows of W are satisfied, dx moves inside them):
             * tau = r1'dx / dx'P dx >= 1 */
            matvec(M->P, q->dx, q->px, n, n);
            double pd = 0, cd = 0;
            for (int j = 0; j < n; j++) { pd += q->r1[j] * q->dx[j]; cd += q->px[j] * q->dx[j]; }
            if (pd > 0.0) {
                const double tau = (cd > 0.0 && pd < 1e12 * cd) ? pd / cd : 1e12;
                if (tau > 1.0) for (int j = 0; j < n; j++) q->dx[j] *= tau;
            }
        }
        for (int j = 0; j < n; j++) x[j] += q->dx[j];
        for (int a = 0; a < nw; a++) lam[a] += q->dl[a];
    }
    q->eqp_res = best;
    return nw;
}

/* KKT conditions of the full QP at the EQP point (q->xe, q->lam compact, nw rows in q->idx).
 * Returns 0 when satisfied, else a reason code: 2 stationarity, 3 active row off its bound,
 * 4 an inactive row violated, 5/6 wrong multiplier sign.  *worst receives the row to drop (5/6). */
static int kkt_check(qp_inst* q, const int* W, int nw, int* worst)
{
    const qp_mats* M = q->M;
    const int n = M->n, m = M->m;
    const double ftol = q->o->qp_feas_tol, dtol = q->o->qp_dual_tol;
    const double* x = q->xe;
    const double* lam = q->lam;
    double ln = 0;
    matvec(M->P, x, q->px, n, n);
    for (int j = 0; j < n; j++) q->r1[j] = -q->q[j] - q->px[j];
    for (int a = 0; a < nw; a++) {
        const int i = q->idx[a];
        for (int k = M->Ap[i]; k < M->Ap[i + 1]; k++) q->r1[M->Aj[k]] -= M->Ax[k] * lam[a];
        if (fabs(lam[a]) > ln) ln = fabs(lam[a]);
    }
    double rs = 0;
    for (int j = 0; j < n; j++) if (fabs(q->r1[j]) > rs) rs = fabs(q->r1[j]);
    if (!(rs <= RES_TOL * (1.0 + ln))) return 2;
    csr_matvec(M, x, q->zt);
    for (int a = 0; a < nw; a++) {
        const int i = q->idx[a];
        const double b = (W[i] == 1 ? q->l[i] : q->u[i]);
        if (!(fabs(b - q->zt[i]) <= RES_TOL * (1.0 + fabs(b)))) return 3;
    }
    for (int i = 0; i < m; i++) {
        const double tol = ftol * (1.0 + fabs(q->zt[i]));
        if (q->zt[i] < q->l[i] - tol || q->zt[i] > q->u[i] + tol) return 4;
    }
    double wv = dtol * (1.0 + ln);
    int reason = 0;
    for (int a = 0; a < nw; a++) {
        const int i = q->idx[a];
        if (q->ctype[i] == 1 || q->pin[i]) continue;
        const double v = (W[i] == 1) ? lam[a] : -lam[a]; /* OSQP sign: lower-active needs lam <= 0 */
        if (v > wv) { wv = v; reason = (W[i] == 1) ? 5 : 6; if (worst) *worst = i; }
    }
    return reason;
}

static void accept_solution(qp_inst* q, const int* W, int nw)
{
    const qp_mats* M = q->M;
    const int n = M->n, m = M->m;
    if (W != q->W) memcpy(q->W, W, (size_t)m * sizeof(int));
    q->have_W = 1;
    /* expand the multipliers to full length */
    memset(q->y, 0, (size_t)m * sizeof(double));
    for (int a = 0; a < nw; a++) q->y[q->idx[a]] = q->lam[a];
    /* un-scale (auxil.c:524-562): x = D xbar, y = E ybar / c; qpOASES sign = -OSQP sign */
    for (int j = 0; j < n; j++) q->xs[j] = M->D[j] * q->xe[j];
    for (int i = 0; i < m; i++) q->ys[i] = -(M->E[i] * q->y[i]) / M->c;
    /* state for the next call: the exact solution (ADMM warm start and active-set start point) */
    memcpy(q->x, q->xe, (size_t)n * sizeof(double));
    csr_matvec(M, q->x, q->z);
}

/* Phase 1 of the crossover: move the ADMM iterate onto the rows of W (H-metric projection through the
 * Cholesky factor of S left in q->S by the last eqp_solve(W)) and test feasibility of every row.
 * Returns 1 with the feasible point in xout. */
static int project_feasible(qp_inst* q, const int* W, int nw, const double* xin, double* xout)
{
    const qp_mats* M = q->M;
    const int n = M->n, m = M->m;
    memcpy(xout, xin, (size_t)n * sizeof(double));
    for (int pass = 0; pass < 4; pass++) {
        double rn = 0;
        for (int a = 0; a < nw; a++) {
            const int i = q->idx[a];
            double s = 0;
            for (int k = M->Ap[i]; k < M->Ap[i + 1]; k++) s += M->Ax[k] * xout[M->Aj[k]];
            q->dl[a] = (W[i] == 1 ? q->l[i] : q->u[i]) - s;
            if (fabs(q->dl[a]) > rn) rn = fabs(q->dl[a]);
        }
        if (rn < 1e-15) break;
        if (nw) chol_solve(q->S, nw, nw, q->dl);
        memset(q->t2, 0, (size_t)n * sizeof(double));
        for (int a = 0; a < nw; a++) {
            const int i = q->idx[a];
            for (int k = M->Ap[i]; k < M->Ap[i + 1]; k++) q->t2[M->Aj[k]] += M->Ax[k] * q->dl[a];
        }
        matvec(M->Hinv, q->t2, q->dx, n, n);
        for (int j = 0; j < n; j++) xout[j] += q->dx[j];
    }
    csr_matvec(M, xout, q->zt);
    const double ftol = q->o->qp_feas_tol;
    for (int i = 0; i < m; i++) {
        const double tol = ftol * (1.0 + fabs(q->zt[i]));
        if (q->zt[i] < q->l[i] - tol || q->zt[i] > q->u[i] + tol) return 0;
        if (W[i]) { const double b = (W[i] == 1 ? q->l[i] : q->u[i]); if (fabs(q->zt[i] - b) > tol) return 0; }
    }
    return 1;
}

/* Primal active-set iteration from a feasible point x (scaled) whose active rows contain W.
 * Each pass solves EQP(W) (above); the step toward the EQP minimiser is cut at the first blocking
 * row, which joins W; at an EQP minimiser the row with the most wrong-signed multiplier leaves W.
 * This is the textbook method (the role qpOASES' working-set changes play in the reference,
 * external/qpOASES/src/QProblem.cpp:1555-1722), expressed through the shared Hinv/G so that every
 * linear solve is a gather + small Cholesky.  Returns 0 with the solution accepted, 1 if it gave up. */
static int active_set(qp_inst* q, double* x, int* W, int* changes)
{
    const qp_mats* M = q->M;
    const int n = M->n, m = M->m;
    const int cap = 20 * (n + m) + 100;
    double* Axv = q->w;   /* A x  */
    double* Apv = q->zt2; /* A p  */
    int last_dropped = -1;
    for (int it = 0; it < cap; it++) {
        int nw = eqp_solve(q, W);
        if (nw < 0) return 1;
        /* direction to the EQP minimiser and ratio test (Nocedal & Wright alg. 16.3) */
        double pn = 0, xn = 0;
        for (int j = 0; j < n; j++) {
            q->dx[j] = q->xe[j] - x[j];
            if (fabs(q->dx[j]) > pn) pn = fabs(q->dx[j]);
            if (fabs(x[j]) > xn) xn = fabs(x[j]);
        }
        csr_matvec(M, x, Axv);
        csr_matvec(M, q->dx, Apv);
        double apn = 0;
        for (int i = 0; i < m; i++) if (fabs(Apv[i]) > apn) apn = fabs(Apv[i]);
        double alpha = 1.0, sbest = 0;
        int block = -1, bside = 0;
        const double seps = 1e-13 * (1.0 + apn);
        for (int i = 0; i < m; i++) {
            if (W[i] || q->ctype[i] < 0) continue;
            const double s = Apv[i];
            double a = 2.0; int side = 0;
            if (s < -seps && q->l[i] > -INFINITY) { double sl = Axv[i] - q->l[i]; if (sl < 0) sl = 0; a = sl / (-s); side = 1; }
            else if (s > seps && q->u[i] < INFINITY) { double su = q->u[i] - Axv[i]; if (su < 0) su = 0; a = su / s; side = 2; }
            if (side && (a < alpha || (a == alpha && block >= 0 && fabs(s) > sbest))) { alpha = a; block = i; bside = side; sbest = fabs(s); }
        }
        if (block >= 0) {
            for (int j = 0; j < n; j++) x[j] += alpha * q->dx[j];
            W[block] = bside;
            /* a row that comes straight back after a zero-length step was dropped on multiplier noise: it
             * is weakly active; exempt it from the sign test for the rest of this QP (anti-cycling) */
            if (block == last_dropped && alpha * (1.0 + apn) <= 1e-12) q->pin[block] = 1;
            last_dropped = -1;
            (*changes)++;
            if (dbg() > 1) fprintf(stderr, "    as it=%d nw=%d add row %d side %d alpha=%.3e |p|=%.3e res=%.1e\n", it, nw, block, bside, alpha, pn, q->eqp_res);
            continue;
        }
        /* full step: x is the EQP minimiser; check its multipliers */
        memcpy(x, q->xe, (size_t)n * sizeof(double));
        int worst = -1;
        int reason = kkt_check(q, W, nw, &worst);
        if (reason == 0) { accept_solution(q, W, nw); return 0; }
        if ((reason == 5 || reason == 6) && worst >= 0) {
            if (dbg() > 1) fprintf(stderr, "    as it=%d nw=%d drop row %d (reason %d) res=%.1e\n", it, nw, worst, reason, q->eqp_res);
            W[worst] = 0;
            last_dropped = worst;
            (*changes)++;
            continue;
        }
        if (dbg()) fprintf(stderr, "    as it=%d nw=%d gives up: kkt reason %d res=%.1e\n", it, nw, reason, q->eqp_res);
        return 1; /* EQP not solvable to tolerance on this set (dependent/inconsistent rows) */
    }
    return 1;
}

/* rho estimate, auxil.c:13-45 */
static double rho_estimate(qp_inst* q)
{
    const qp_mats* M = q->M;
    const int n = M->n, m = M->m;
    csr_matvec(M, q->x, q->zt);
    double pri = 0, nz = 0, nax = 0;
    for (int i = 0; i < m; i++) {
        double d = fabs(q->zt[i] - q->z[i]);
        if (d > pri) pri = d;
        if (fabs(q->z[i]) > nz) nz = fabs(q->z[i]);
        if (fabs(q->zt[i]) > nax) nax = fabs(q->zt[i]);
    }
    matvec(M->P, q->x, q->px, n, n);
    memset(q->t1, 0, (size_t)n * sizeof(double));
    csr_tmatvec_add(M, q->y, q->t1);
    double dua = 0, npx = 0, naty = 0, nq = 0;
    for (int j = 0; j < n; j++) {
        double d = fabs(q->px[j] + q->q[j] + q->t1[j]);
        if (d > dua) dua = d;
        if (fabs(q->px[j]) > npx) npx = fabs(q->px[j]);
        if (fabs(q->t1[j]) > naty) naty = fabs(q->t1[j]);
        if (fabs(q->q[j]) > nq) nq = fabs(q->q[j]);
    }
    double pn = nz > nax ? nz : nax;
    double dn = npx > naty ? npx : naty;
    if (nq > dn) dn = nq;
    pri /= (pn + 1e-10);
    dua /= (dn + 1e-10);
    double r = q->rho * sqrt(pri / (dua + 1e-10));
    if (r < RHO_MIN) r = RHO_MIN;
    if (r > RHO_MAX) r = RHO_MAX;
    return r;
}

/* SubsolverBase::solve contract (SubsolverBase.hpp:37-56).  g is the unscaled linear term.
 * Returns 0 on success; *iterations = ADMM iterations + working-set changes of this call. */
static int qp_solve(qp_inst* q, int initial, const double* g, const double* x0, const double* y0_full /* m, qpOASES sign */, int* iterations)
{
    const qp_mats* M = q->M;
    const lcqp_oracle_options* o = q->o;
    const int n = M->n, m = M->m;
    *iterations = 0;
    if (q->infeasible_bounds) return 37; /* qpOASES RET_INIT_FAILED_INFEASIBILITY class */
    if (!q->Minv) return 38;
    for (int j = 0; j < n; j++) q->q[j] = M->c * M->D[j] * g[j]; /* osqp.c:752-779 */
    memset(q->pin, 0, (size_t)m * sizeof(int));

    int changes = 0;
    if (initial) {
        /* warm start like osqp_warm_start_x / _y (osqp.c:700-745): x = D^-1 x0, z = A x, y = c E^-1 y0 */
        for (int j = 0; j < n; j++) q->x[j] = x0 ? x0[j] / M->D[j] : 0.0;
        csr_matvec(M, q->x, q->z);
        for (int i = 0; i < m; i++) q->y[i] = y0_full ? -(M->c * y0_full[i]) / M->E[i] : 0.0;
        q->have_W = 0;
    } else if (q->have_W) {
        /* hot start (the analogue of qpOASES' hotstart, SubsolverQPOASES.cpp:154-160): the previous
         * optimum is feasible for the new gradient; continue the active-set iteration from it */
        memcpy(q->xa, q->x, (size_t)n * sizeof(double));
        memcpy(q->Wtry, q->W, (size_t)m * sizeof(int));
        if (active_set(q, q->xa, q->Wtry, &changes) == 0) { *iterations = changes; return 0; }
        /* fall through to ADMM from the previous solution */
        csr_matvec(M, q->x, q->z);
    }
    q->have_fail = 0;
    int it = 0;
    while (it < o->qp_max_iter) {
        for (int k = 0; k < o->qp_check_interval && it < o->qp_max_iter; k++, it++) admm_iter(q);
        guess_working_set(q, q->Wtry);
        if (!(q->have_fail && memcmp(q->Wtry, q->Wfail, (size_t)m * sizeof(int)) == 0)) {
            memcpy(q->Wfail, q->Wtry, (size_t)m * sizeof(int));
            q->have_fail = 1;
            int nw = eqp_solve(q, q->Wtry);
            int reason = nw < 0 ? 1 : kkt_check(q, q->Wtry, nw, NULL);
            if (dbg()) fprintf(stderr, "  admm it=%d nw=%d probe reason=%d res=%.1e rho=%g\n", it, nw, reason, q->eqp_res, q->rho);
            if (reason == 0) { accept_solution(q, q->Wtry, nw); *iterations = it + changes; return 0; }
            if (reason == 5 || reason == 6) {
                /* primal feasible EQP point with a wrong-signed multiplier: a valid active-set start */
                memcpy(q->xa, q->xe, (size_t)n * sizeof(double));
                if (active_set(q, q->xa, q->Wtry, &changes) == 0) { *iterations = it + changes; return 0; }
            } else if (nw >= 0 && project_feasible(q, q->Wtry, nw, q->x, q->xa)) {
                /* EQP point not usable, but the ADMM iterate projects to a feasible point on W */
                if (dbg()) fprintf(stderr, "  admm it=%d projected start feasible\n", it);
                if (active_set(q, q->xa, q->Wtry, &changes) == 0) { *iterations = it + changes; return 0; }
            }
        }
        if (o->qp_adaptive_rho && it % 50 == 0) {
            /* rho ladder: steps of x5 (osqp.c:459-516 applies an update when the estimate is off by > 5x) */
            double est = rho_estimate(q);
            double ratio = est / q->rho;
            if (ratio > 5.0 || ratio < 0.2) {
                int steps = (int)lround(log(ratio) / log(5.0));
                double nr = q->rho * pow(5.0, steps);
                if (nr < RHO_MIN) nr = RHO_MIN;
                if (nr > RHO_MAX) nr = RHO_MAX;
                qp_set_rho(q, nr); /* y, z kept: OSQP's update_rho_vec only refactors (auxil.c:100-142) */
                if (!q->Minv) return 38;
            }
        }
    }
    *iterations = it + changes;
    return -2; /* OSQP_MAX_ITER_REACHED class */
}

/* ================================================================================================
 * The LCQP penalty-homotopy loop  (LCQProblem.cpp:444-560 and the helpers it calls)
 * ==============================================================================================*/
typedef struct {
    int nV, nC, nComp;
    const double *Q, *g, *L, *R, *lbL, *ubL, *lbR, *ubR, *A, *lbA, *ubA, *lb, *ub, *x0, *y0;
} lcqp_data;

/* prepared, shareable when Q/L/R/A are shared across the batch */
typedef struct {
    double* Afull; /* (nC+2nComp) x nV  (LCQProblem.cpp:572-582) */
    double* C;     /* nV x nV           (LCQProblem.cpp:622-623) */
    qp_mats* M;
    int has_box;
} lcqp_shared;

static void lcqp_shared_free(lcqp_shared* s)
{
    if (!s) return;
    free(s->Afull); free(s->C); qp_mats_free(s->M); free(s);
}

static lcqp_shared* lcqp_shared_create(const lcqp_data* d, const lcqp_oracle_options* o, int has_box)
{
    const int n = d->nV, mA = d->nC + 2 * d->nComp;
    lcqp_shared* s = (lcqp_shared*)xcalloc(1, sizeof(lcqp_shared));
    s->has_box = has_box;
    s->Afull = (double*)xcalloc((size_t)mA * n, sizeof(double));
    if (d->nC) memcpy(s->Afull, d->A, (size_t)d->nC * n * sizeof(double));
    memcpy(s->Afull + (size_t)d->nC * n, d->L, (size_t)d->nComp * n * sizeof(double));
    memcpy(s->Afull + (size_t)(d->nC + d->nComp) * n, d->R, (size_t)d->nComp * n * sizeof(double));
    s->C = (double*)xcalloc((size_t)n * n, sizeof(double));
    lcqp_oracle_MatrixSymmetrizationProduct(d->L, d->R, s->C, d->nComp, n);
    s->M = qp_mats_create(n, mA, has_box, d->Q, s->Afull, o->qp_sigma, o->qp_delta);
    return s;
}

static int lcqp_run(const lcqp_data* d, const lcqp_oracle_options* o, lcqp_shared* sh, unsigned long long instance,
                    double* xout, double* yout, lcqp_oracle_result* res)
{
    const int n = d->nV, nC = d->nC, nComp = d->nComp, mA = nC + 2 * nComp;
    const int osqp_flavour = (o->qpSolver == 2);
    memset(res, 0, sizeof(*res));

    /* initializeSolver: dual layout (LCQProblem.cpp:888-890, 934-935) */
    const int has_box = sh->has_box;
    const int nDuals = osqp_flavour ? mA : n + mA;
    const int boxOff = osqp_flavour ? 0 : n;
    res->nDuals = nDuals;
    if (osqp_flavour && (d->lb || d->ub)) { res->ret = RET_INVALID_OSQP_BOX; return res->ret; } /* :955-957 */

    /* setConstraints + setComplementarityBounds (LCQProblem.cpp:584-608, 745-782) */
    double* lbAf = (double*)xcalloc((size_t)mA, sizeof(double));
    double* ubAf = (double*)xcalloc((size_t)mA, sizeof(double));
    for (int i = 0; i < nC; i++) {
        lbAf[i] = d->lbA ? d->lbA[i] : -INFINITY;
        ubAf[i] = d->ubA ? d->ubA[i] : INFINITY;
    }
    for (int i = 0; i < nComp; i++) {
        if (d->lbL && d->lbL[i] <= -INFINITY) { res->ret = RET_INVALID_LOWER_COMP; free(lbAf); free(ubAf); return res->ret; }
        if (d->lbR && d->lbR[i] <= -INFINITY) { res->ret = RET_INVALID_LOWER_COMP; free(lbAf); free(ubAf); return res->ret; }
        lbAf[nC + i] = d->lbL ? d->lbL[i] : 0.0;
        ubAf[nC + i] = d->ubL ? d->ubL[i] : INFINITY;
        lbAf[nC + nComp + i] = d->lbR ? d->lbR[i] : 0.0;
        ubAf[nC + nComp + i] = d->ubR ? d->ubR[i] : INFINITY;
    }
    /* setLB/setUB (LCQProblem.ipp:53-118): qpOASES flavour always has box arrays (+-inf default) */
    double* lbv = NULL; double* ubv = NULL;
    if (has_box) {
        lbv = (double*)xcalloc((size_t)n, sizeof(double));
        ubv = (double*)xcalloc((size_t)n, sizeof(double));
        for (int i = 0; i < n; i++) { lbv[i] = d->lb ? d->lb[i] : -INFINITY; ubv[i] = d->ub ? d->ub[i] : INFINITY; }
    }

    const double* Afull = sh->Afull;
    const double* Cm = sh->C;
    qp_inst* qp = qp_inst_create(sh->M, o, lbAf, ubAf, lbv, ubv);

    double* xk = (double*)xcalloc((size_t)n, sizeof(double));
    double* yk = (double*)xcalloc((size_t)(n + mA), sizeof(double));
    double* ykA = (double*)xcalloc((size_t)mA, sizeof(double));
    double* ybox = (double*)xcalloc((size_t)n, sizeof(double));
    double* xnew = (double*)xcalloc((size_t)n, sizeof(double));
    double* pk = (double*)xcalloc((size_t)n, sizeof(double));
    double* gk = (double*)xcalloc((size_t)n, sizeof(double));
    double* gt = (double*)xcalloc((size_t)n, sizeof(double));
    double* gphi = NULL;
    double* statk = (double*)xcalloc((size_t)n, sizeof(double));
    double* cst = (double*)xcalloc((size_t)n, sizeof(double));
    double* lkt = (double*)xcalloc((size_t)n, sizeof(double));
    double* Qk = (double*)xcalloc((size_t)n * n, sizeof(double));
    double* y0full = NULL;
    double hist[64]; int nh = 0;
    int ret = RET_OK;

    if (d->x0) memcpy(xk, d->x0, (size_t)n * sizeof(double)); /* LCQProblem.ipp:138-142 */
    if (d->y0) {
        /* user duals: box(n) + A + L + R (LCQProblem.ipp:144-151); the QP solver takes the m = mA(+n) rows */
        y0full = (double*)xcalloc((size_t)(mA + n), sizeof(double));
        for (int i = 0; i < mA; i++) y0full[i] = d->y0[n + i];
        if (has_box) for (int i = 0; i < n; i++) y0full[mA + i] = d->y0[i];
    }

    /* g_tilde, phi_const, g_phi (LCQProblem.cpp:966-996) */
    memcpy(gt, d->g, (size_t)n * sizeof(double));
    double phi_const = 0;
    if (d->lbL || d->lbR) {
        static const double zero64[1] = {0};
        (void)zero64;
        double* zl = NULL;
        const double* pl = d->lbL; const double* pr = d->lbR;
        if (!pl || !pr) { zl = (double*)xcalloc((size_t)nComp, sizeof(double)); if (!pl) pl = zl; if (!pr) pr = zl; }
        phi_const = lcqp_oracle_DotProduct(pl, pr, nComp);
        gphi = (double*)xcalloc((size_t)n, sizeof(double));
        if (d->lbL) lcqp_oracle_AddTransponsedMatrixMultiplication(d->R, d->lbL, gphi, nComp, n, 1);
        if (d->lbR) lcqp_oracle_AddTransponsedMatrixMultiplication(d->L, d->lbR, gphi, nComp, n, 1);
        for (int i = 0; i < n; i++) gphi[i] = -gphi[i];
        free(zl);
    }
    double alphak = 1.0, rho = o->initialPenaltyParameter;
    int outerIter = 0, totalIter = 0, subIter = 0, qpIter = 0, exitFlag = 0;
    int status = 0;

#define PHI() (phi_const + (gphi ? lcqp_oracle_DotProduct(gphi, xk, n) : 0.0) + lcqp_oracle_QuadraticFormProduct(Cm, xk, n) / 2.0) /* :1172-1185 */
#define UPDATE_PENALTY() do { nh = 0; rho *= o->penaltyUpdateFactor; res->rhoOpt = rho; \
        lcqp_oracle_WeightedMatrixAdd(1, d->Q, rho, Cm, Qk, n, n); \
        if (gphi) lcqp_oracle_WeightedVectorAdd(1.0, d->g, rho, gphi, gt, n); } while (0) /* :1199-1214 */
#define SOLVE_QP(initial) do { \
        int fl = qp_solve(qp, initial, gk, xk, y0full, &qpIter); \
        subIter += qpIter; exitFlag = osqp_flavour ? (fl == 0 ? 1 : fl) : fl; \
        if (fl != 0) { ret = (osqp_flavour && qp->infeasible_bounds) ? RET_OSQP_GUESS : RET_SUBPROBLEM; goto done; } \
        memcpy(xnew, qp->xs, (size_t)n * sizeof(double)); \
        memcpy(ykA, qp->ys, (size_t)mA * sizeof(double)); \
        if (has_box) memcpy(ybox, qp->ys + mA, (size_t)n * sizeof(double)); \
        for (int i_ = 0; i_ < n; i_++) pk[i_] = xnew[i_] - xk[i_]; } while (0) /* :1115-1148 */

    /* first QP (LCQProblem.cpp:452-467) */
    if (o->solveZeroPenaltyFirst) memcpy(gk, d->g, (size_t)n * sizeof(double));
    else lcqp_oracle_AffineLinearTransformation(rho, Cm, xk, gt, gk, n, n);
    SOLVE_QP(1);

    lcqp_oracle_WeightedMatrixAdd(1, d->Q, rho, Cm, Qk, n, n); /* setQk :880 */
    res->rhoOpt = rho;                                          /* :473 */

    for (;;) {
        /* updateStep :1240-1243 */
        for (int i = 0; i < n; i++) xk[i] = xk[i] + alphak * pk[i];
        /* updateStationarity :1246-1272 */
        lcqp_oracle_AffineLinearTransformation(1, Qk, xk, gt, statk, n, n);
        lcqp_oracle_TransponsedMatrixMultiplication(Afull, ykA, cst, mA, n, 1);
        for (int i = 0; i < n; i++) statk[i] = statk[i] - cst[i];
        if (has_box) for (int i = 0; i < n; i++) statk[i] = statk[i] - ybox[i];

        totalIter++; /* :493-496 */

        /* leyfferCheckPositive :1275-1313 */
        {
            int nd = o->nDynamicPenalty, fire = 0;
            if (nd > 0) {
                double cur = PHI();
                if (nh < nd) { hist[nh++] = cur; }
                else if (cur < o->complementarityTolerance) { memmove(hist, hist + 1, (size_t)(nd - 1) * sizeof(double)); hist[nd - 1] = cur; }
                else {
                    fire = 1;
                    for (int i = 0; i < nd; i++) if (cur < o->etaDynamicPenalty * hist[i]) { fire = 0; break; }
                    memmove(hist, hist + 1, (size_t)(nd - 1) * sizeof(double)); hist[nd - 1] = cur;
                }
            }
            if (fire) { UPDATE_PENALTY(); outerIter++; }
        }

        lcqp_oracle_AffineLinearTransformation(rho, Cm, xk, gt, gk, n, n); /* :508 */

        if (lcqp_oracle_MaxAbs(statk, n) < o->stationarityTolerance) { /* :511 */
            if (PHI() < o->complementarityTolerance) {
                /* transformDuals :1381-1409 (writes yk only; the classifier reads yk_A, :1420) */
                for (int i = 0; i < n; i++) yk[i] = has_box ? ybox[i] : 0.0;
                for (int i = 0; i < mA; i++) yk[boxOff + i] = ykA[i];
                double* tmp = (double*)xcalloc((size_t)nComp, sizeof(double));
                lcqp_oracle_MatrixMultiplication(d->R, xk, tmp, nComp, n, 1);
                for (int i = 0; i < nComp; i++) yk[boxOff + nC + i] -= rho * tmp[i];
                lcqp_oracle_MatrixMultiplication(d->L, xk, tmp, nComp, n, 1);
                for (int i = 0; i < nComp; i++) yk[boxOff + nC + nComp + i] -= rho * tmp[i];
                /* determineStationarityType :1412-1453, getWeakComplementarities :1456-1482 */
                double* Lx = (double*)xcalloc((size_t)nComp, sizeof(double));
                lcqp_oracle_MatrixMultiplication(d->L, xk, Lx, nComp, n, 1);
                lcqp_oracle_MatrixMultiplication(d->R, xk, tmp, nComp, n, 1);
                int s_stat = 1, m_stat = 1, w_only = 0;
                const double tc = o->complementarityTolerance;
                for (int i = 0; i < nComp && !w_only; i++) {
                    if (!(Lx[i] <= tc && tmp[i] <= tc)) continue;
                    double yl = ykA[nC + i], yr = ykA[nC + nComp + i];
                    double prod = yl * yr, mn = yl < yr ? yl : yr;
                    if (mn < 0) s_stat = 0;
                    if (fabs(prod) >= tc && mn <= 0) {
                        if (prod <= tc) { w_only = 1; break; }
                        m_stat = 0;
                    }
                }
                status = w_only ? 1 : (s_stat ? 4 : (m_stat ? 3 : 2));
                free(Lx); free(tmp);
                ret = RET_OK;
                goto done_success;
            } else {
                UPDATE_PENALTY(); outerIter++;
            }
        }
        if (totalIter > o->maxIterations) { ret = RET_MAX_ITER; goto done; } /* :537 */
        if (rho > o->maxPenaltyParameter) { ret = RET_MAX_PEN; goto done; }   /* :541 */

        lcqp_oracle_AffineLinearTransformation(rho, Cm, xk, gt, gk, n, n); /* :545 */
        SOLVE_QP(0);                                                        /* :548 */

        if (o->perturbStep) /* :553-555, :1353-1362 */
            for (int i = 0; i < n; i++) xk[i] += lcqp_oracle_perturb_draw(o->perturb_seed, instance, (unsigned)totalIter, (unsigned)i) * LCQ_EPS;

        /* getOptimalStepLength :1217-1237 */
        {
            double qk = lcqp_oracle_QuadraticFormProduct(Qk, pk, n);
            lcqp_oracle_AffineLinearTransformation(1, Qk, xk, gt, lkt, n, n);
            double lk = lcqp_oracle_DotProduct(pk, lkt, n);
            alphak = 1;
            if (qk > 0 && lk < 0) { double a = -lk / qk; alphak = a < 1.0 ? a : 1.0; }
        }
    }

done:
    /* failure exits leave yk as last returned by the subsolver (LCQProblem.cpp:1138) */
    for (int i = 0; i < n; i++) yk[i] = has_box ? ybox[i] : 0.0;
    for (int i = 0; i < mA; i++) yk[boxOff + i] = ykA[i];
done_success:
    res->ret = ret;
    res->status = status;
    res->iterTotal = totalIter;
    res->iterOuter = outerIter;
    res->subproblemIter = subIter;
    res->qpExitFlag = exitFlag;
    memcpy(xout, xk, (size_t)n * sizeof(double));
    memcpy(yout, yk, (size_t)nDuals * sizeof(double));
    free(lbAf); free(ubAf); free(lbv); free(ubv); qp_inst_free(qp);
    free(xk); free(yk); free(ykA); free(ybox); free(xnew); free(pk); free(gk); free(gt); free(gphi);
    free(statk); free(cst); free(lkt); free(Qk); free(y0full);
    return ret;
#undef PHI
#undef UPDATE_PENALTY
#undef SOLVE_QP
}

/* Options.cpp:296-333 */
void lcqp_oracle_default_options(lcqp_oracle_options* o)
{
    memset(o, 0, sizeof(*o));
    o->complementarityTolerance = 1.0e3 * LCQ_EPS;
    o->stationarityTolerance = 1.0e6 * LCQ_EPS;
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

int lcqp_oracle_solve_batch(int batch, int nV, int nC, int nComp, unsigned shared_mask,
                            const double* Q, const double* g, const double* L, const double* R,
                            const double* lbL, const double* ubL, const double* lbR, const double* ubR,
                            const double* A, const double* lbA, const double* ubA,
                            const double* lb, const double* ub,
                            const double* x0, const double* y0,
                            const lcqp_oracle_options* o, double* x, double* y, lcqp_oracle_result* res)
{
    const size_t nD = (size_t)nV + nC + 2 * nComp;
    const size_t len[15] = {(size_t)nV * nV, (size_t)nV, (size_t)nComp * nV, (size_t)nComp * nV,
                            (size_t)nComp, (size_t)nComp, (size_t)nComp, (size_t)nComp,
                            (size_t)nC * nV, (size_t)nC, (size_t)nC, (size_t)nV, (size_t)nV, (size_t)nV, nD};
    const double* base[15] = {Q, g, L, R, lbL, ubL, lbR, ubR, A, lbA, ubA, lb, ub, x0, y0};
    /* matrices shared by the whole batch -> one scaling / Hinv / G for all instances */
    const unsigned mat_bits = (1u << 0) | (1u << 2) | (1u << 3) | (A ? (1u << 8) : 0u);
    const int mats_shared = (shared_mask & mat_bits) == mat_bits;
    /* box rows only when the caller passed lb/ub: with the +-inf defaults of LCQProblem.ipp:53-118 the box
     * multipliers are identically zero, so omitting the rows changes nothing observable */
    const int has_box = (o->qpSolver != 2) && (lb || ub);
    lcqp_shared* sh = NULL;
    int nfail = 0;
    for (int b = 0; b < batch; b++) {
        const double* p[15];
        for (int k = 0; k < 15; k++) {
            if (!base[k]) p[k] = 0;
            else p[k] = (shared_mask >> k) & 1u ? base[k] : base[k] + (size_t)b * len[k];
        }
        lcqp_data d = {nV, nC, nComp, p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7], p[8], p[9], p[10], p[11], p[12], p[13], p[14]};
        struct timespec t0, t1;
        clock_gettime(CLOCK_MONOTONIC, &t0);
        if (!sh || !mats_shared) {
            lcqp_shared_free(sh);
            sh = lcqp_shared_create(&d, o, has_box);
        }
        int rv;
        if (!sh->M) { memset(res + b, 0, sizeof(*res)); res[b].ret = RET_SUBPROBLEM; rv = RET_SUBPROBLEM; }
        else rv = lcqp_run(&d, o, sh, (unsigned long long)b, x + (size_t)b * nV, y + (size_t)b * nD, res + b);
        clock_gettime(CLOCK_MONOTONIC, &t1);
        res[b].seconds = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
        if (rv != 0) nfail++;
    }
    lcqp_shared_free(sh);
    return nfail;
}

// Utilities.cpp -- host-side helpers of the LCQPow API (see ../include/Utilities.hpp).
// Semantics follow /root/reference/src/Utilities.cpp (lines cited per function); the code is new.
#include "Utilities.hpp"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace LCQPow {

// ---- dense, row-major ------------------------------------------------------------------------------
// C (m x p) = A (m x n) B (n x p)                                              [Utilities.cpp:38-47]
void Utilities::MatrixMultiplication(const double* A, const double* B, double* C, int m, int n, int p)
{
    for (int r = 0; r < m; ++r)
        for (int c = 0; c < p; ++c) {
            double acc = 0.0;
            for (int k = 0; k < n; ++k) acc += A[r * n + k] * B[k * p + c];
            C[r * p + c] = acc;
        }
}

// c (m) = A b for a csc matrix                                                 [:49-59]
void Utilities::MatrixMultiplication(const csc* A, const double* b, double* c)
{
    std::fill(c, c + A->m, 0.0);
    for (int col = 0; col < A->n; ++col)
        for (int k = A->p[col]; k < A->p[col + 1]; ++k) c[A->i[k]] += A->x[k] * b[col];
}

// C (n x p) = A' B with A m x n, B m x p                                       [:62-72]
void Utilities::TransponsedMatrixMultiplication(const double* A, const double* B, double* C, int m, int n, int p)
{
    for (int r = 0; r < n; ++r)
        for (int c = 0; c < p; ++c) {
            double acc = 0.0;
            for (int k = 0; k < m; ++k) acc += A[k * n + r] * B[k * p + c];
            C[r * p + c] = acc;
        }
}

// c (n) = A' b                                                                 [:75-82]
void Utilities::TransponsedMatrixMultiplication(const csc* A, const double* b, double* c)
{
    for (int col = 0; col < A->n; ++col) {
        double acc = 0.0;
        for (int k = A->p[col]; k < A->p[col + 1]; ++k) acc += b[A->i[k]] * A->x[k];
        c[col] = acc;
    }
}

// C += A' B                                                                    [:85-93]
void Utilities::AddTransponsedMatrixMultiplication(const double* A, const double* B, double* C, int m, int n, int p)
{
    for (int r = 0; r < n; ++r)
        for (int c = 0; c < p; ++c)
            for (int k = 0; k < m; ++k) C[r * p + c] += A[k * n + r] * B[k * p + c];
}

// c += A' b                                                                    [:96-102]
void Utilities::AddTransponsedMatrixMultiplication(const csc* A, const double* b, double* c)
{
    for (int col = 0; col < A->n; ++col)
        for (int k = A->p[col]; k < A->p[col + 1]; ++k) c[col] += b[A->i[k]] * A->x[k];
}

// C (n x n) = A'B + B'A with A, B m x n; lower triangle computed, mirrored     [:104-116]
void Utilities::MatrixSymmetrizationProduct(const double* A, const double* B, double* C, int m, int n)
{
    for (int r = 0; r < n; ++r)
        for (int c = 0; c <= r; ++c) {
            double acc = 0.0;
            for (int k = 0; k < m; ++k) acc += A[k * n + r] * B[k * n + c] + B[k * n + r] * A[k * n + c];
            C[r * n + c] = acc;
            C[c * n + r] = acc;
        }
}

// sparse C = L'R + R'L; entries with |c| <= ZERO are dropped; an all-zero C gives a null pointer  [:118-173]
csc* Utilities::MatrixSymmetrizationProduct(const csc* L, const csc* R)
{
    const int n = L->n;
    // column j of L'R + R'L: sum over the rows shared by column i of one factor and column j of the other
    std::vector<double> colL(L->m > R->m ? L->m : R->m, 0.0), colR(colL.size(), 0.0);
    std::vector<int> rows;
    std::vector<double> vals;
    int* cp = static_cast<int*>(malloc(sizeof(int) * (size_t)(n + 1)));
    cp[0] = 0;
    for (int j = 0; j < n; ++j) {
        for (int k = L->p[j]; k < L->p[j + 1]; ++k) colL[L->i[k]] = L->x[k];
        for (int k = R->p[j]; k < R->p[j + 1]; ++k) colR[R->i[k]] = R->x[k];
        cp[j + 1] = cp[j];
        for (int i = 0; i < n; ++i) {
            double acc = 0.0;
            for (int k = L->p[i]; k < L->p[i + 1]; ++k) acc += L->x[k] * colR[L->i[k]];
            for (int k = R->p[i]; k < R->p[i + 1]; ++k) acc += R->x[k] * colL[R->i[k]];
            if (!isZero(acc)) { rows.push_back(i); vals.push_back(acc); cp[j + 1]++; }
        }
        for (int k = L->p[j]; k < L->p[j + 1]; ++k) colL[L->i[k]] = 0.0;
        for (int k = R->p[j]; k < R->p[j + 1]; ++k) colR[R->i[k]] = 0.0;
    }
    if (cp[n] == 0) { free(cp); return nullptr; }
    int* ci = static_cast<int*>(malloc(sizeof(int) * rows.size()));
    double* cx = static_cast<double*>(malloc(sizeof(double) * vals.size()));
    std::memcpy(ci, rows.data(), sizeof(int) * rows.size());
    std::memcpy(cx, vals.data(), sizeof(double) * vals.size());
    return createCSC(n, n, cp[n], cx, ci, cp);
}

// d (m) = alpha A b + c                                                        [:176-186]
void Utilities::AffineLinearTransformation(double alpha, const double* A, const double* b, const double* c, double* d, int m, int n)
{
    for (int r = 0; r < m; ++r) {
        double acc = 0.0;
        for (int k = 0; k < n; ++k) acc += A[r * n + k] * b[k];
        d[r] = alpha * acc + c[r];
    }
}

// same for a symmetric csc matrix (column j used as row j)                     [:189-199]
void Utilities::AffineLinearTransformation(double alpha, const csc* S, const double* b, const double* c, double* d, int m)
{
    for (int j = 0; j < m; ++j) {
        double acc = 0.0;
        for (int k = S->p[j]; k < S->p[j + 1]; ++k) acc += S->x[k] * b[S->i[k]];
        d[j] = alpha * acc + c[j];
    }
}

void Utilities::WeightedMatrixAdd(double alpha, const double* A, double beta, const double* B, double* C, int m, int n)
{
    for (long e = 0; e < (long)m * n; ++e) C[e] = alpha * A[e] + beta * B[e];   // [:202-206]
}

void Utilities::WeightedVectorAdd(double alpha, const double* a, double beta, const double* b, double* c, int m)
{
    WeightedMatrixAdd(alpha, a, beta, b, c, m, 1);                               // [:209-211]
}

// p' Q p                                                                       [:214-225]
double Utilities::QuadraticFormProduct(const double* Q, const double* p, int m)
{
    double total = 0.0;
    for (int r = 0; r < m; ++r) {
        double acc = 0.0;
        for (int c = 0; c < m; ++c) acc += Q[r * m + c] * p[c];
        total += acc * p[r];
    }
    return total;
}

double Utilities::QuadraticFormProduct(const csc* S, const double* p, int m)   // [:228-241]
{
    double total = 0.0;
    for (int j = 0; j < m; ++j) {
        double acc = 0.0;
        for (int k = S->p[j]; k < S->p[j + 1]; ++k) acc += S->x[k] * p[S->i[k]];
        total += p[j] * acc;
    }
    return total;
}

double Utilities::DotProduct(const double* a, const double* b, int m)          // [:244-250]
{
    double acc = 0.0;
    for (int k = 0; k < m; ++k) acc += a[k] * b[k];
    return acc;
}

// infinity norm (documented as "1-norm" in the reference)  [:253-265]
double Utilities::MaxAbs(const double* a, int m)
{
    double best = 0.0;
    for (int k = 0; k < m; ++k) best = getMax(best, getAbs(a[k]));
    return best;
}

// ---- csc -------------------------------------------------------------------------------------------
csc* Utilities::createCSC(int m, int n, int nnz, double* x, int* i, int* p)   // [:469-484]
{
    csc* M = static_cast<csc*>(malloc(sizeof(csc)));
    if (!M) return nullptr;
    M->nzmax = nnz; M->m = m; M->n = n; M->p = p; M->i = i; M->x = x; M->nz = -1;
    return M;
}

csc* Utilities::copyCSC(int m, int n, int nnz, const double* x, const int* i, const int* p)   // [:487-513]
{
    int* ci = static_cast<int*>(malloc(sizeof(int) * (size_t)(nnz > 0 ? nnz : 1)));
    double* cx = static_cast<double*>(malloc(sizeof(double) * (size_t)(nnz > 0 ? nnz : 1)));
    int* cp = static_cast<int*>(malloc(sizeof(int) * (size_t)(n + 1)));
    if (nnz > 0) { std::memcpy(ci, i, sizeof(int) * (size_t)nnz); std::memcpy(cx, x, sizeof(double) * (size_t)nnz); }
    std::memcpy(cp, p, sizeof(int) * (size_t)(n + 1));
    return createCSC(m, n, nnz, cx, ci, cp);
}

csc* Utilities::copyCSC(const csc* M, bool toUpperTriangular)                  // [:516-583]
{
    if (!toUpperTriangular) return copyCSC(M->m, M->n, M->nzmax, M->x, M->i, M->p);
    std::vector<int> rows;
    std::vector<double> vals;
    int* cp = static_cast<int*>(malloc(sizeof(int) * (size_t)(M->n + 1)));
    cp[0] = 0;
    for (int j = 0; j < M->n; ++j) {
        for (int k = M->p[j]; k < M->p[j + 1]; ++k)
            if (M->i[k] <= j) { rows.push_back(M->i[k]); vals.push_back(M->x[k]); }
        cp[j + 1] = (int)rows.size();
    }
    csc* U = copyCSC(M->m, M->n, (int)rows.size(), vals.data(), rows.data(), cp);
    free(cp);
    return U;
}

void Utilities::ClearSparseMat(csc* M)
{
    if (!M) return;
    free(M->p); free(M->i); free(M->x); free(M);
}

void Utilities::ClearSparseMat(csc** M)
{
    if (M && *M) { ClearSparseMat(*M); *M = nullptr; }
}

double* Utilities::csc_to_dns(const csc* sparse)                               // [:593-617]
{
    const int m = sparse->m, n = sparse->n;
    double* full = new double[(size_t)m * n]();
    for (int j = 0; j < n; ++j)
        for (int k = sparse->p[j]; k < sparse->p[j + 1] && k < sparse->nzmax; ++k) {
            const int r = sparse->i[k];
            if (r < 0 || r >= m) { MessageHandler::PrintMessage(INDEX_OUT_OF_BOUNDS, ERROR); delete[] full; return nullptr; }
            full[(size_t)r * n + j] = sparse->x[k];
        }
    return full;
}

csc* Utilities::dns_to_csc(const double* full, int m, int n)                   // [:620-650]
{
    std::vector<int> rows;
    std::vector<double> vals;
    int* cp = static_cast<int*>(malloc(sizeof(int) * (size_t)(n + 1)));
    cp[0] = 0;
    for (int j = 0; j < n; ++j) {
        for (int r = 0; r < m; ++r) {
            const double v = full[(size_t)r * n + j];
            if (v > 0 || v < 0) { rows.push_back(r); vals.push_back(v); }
        }
        cp[j + 1] = (int)rows.size();
    }
    csc* S = copyCSC(m, n, (int)rows.size(), vals.data(), rows.data(), cp);
    free(cp);
    return S;
}

// ---- text files ------------------------------------------------------------------------------------
ReturnValue Utilities::readFromFile(int* data, int n, const char* datafilename)
{
    FILE* f = datafilename ? fopen(datafilename, "r") : nullptr;
    if (!f) return UNABLE_TO_READ_FILE;
    for (int k = 0; k < n; ++k)
        if (fscanf(f, "%d", &data[k]) != 1) { fclose(f); return UNABLE_TO_READ_FILE; }
    fclose(f);
    return SUCCESSFUL_RETURN;
}

// one value per line, "%lf" (so "Inf"/"-Inf"/"nan" are accepted, as in examples/example_data)   [:341-366]
ReturnValue Utilities::readFromFile(double* data, int n, const char* datafilename)
{
    FILE* f = datafilename ? fopen(datafilename, "r") : nullptr;
    if (!f) return UNABLE_TO_READ_FILE;
    for (int k = 0; k < n; ++k)
        if (fscanf(f, "%lf", &data[k]) != 1) { fclose(f); return UNABLE_TO_READ_FILE; }
    fclose(f);
    return SUCCESSFUL_RETURN;
}

ReturnValue Utilities::writeToFile(const double* data, int n, const char* datafilename)   // [:369-395]
{
    FILE* f = datafilename ? fopen(datafilename, "w") : nullptr;
    if (!f) return UNABLE_TO_READ_FILE;
    for (int k = 0; k < n; ++k)
        if (fprintf(f, "%.17g\n", data[k]) <= 0) { fclose(f); return UNABLE_TO_READ_FILE; }
    fclose(f);
    return SUCCESSFUL_RETURN;
}

void Utilities::printMatrix(const double* A, int m, int n, const char* name)
{
    printf("Printing matrix %s:\n", name);
    for (int r = 0; r < m; ++r) {
        for (int c = 0; c < n; ++c) printf("%.5f ", A[(size_t)r * n + c]);
        printf("\n");
    }
    printf("\n");
}

void Utilities::printMatrix(const csc* A, const char* name)
{
    double* full = csc_to_dns(A);
    if (!full) return;
    printMatrix(full, A->m, A->n, name);
    delete[] full;
}

// ---- messages --------------------------------------------------------------------------------------
const char* MessageHandler::ReturnValueText(ReturnValue ret)
{
    switch (ret) {
        case NOT_YET_IMPLEMENTED: return "This method has not yet been implemented.";
        case SUCCESSFUL_RETURN: return "Successful return.";
        case INVALID_ARGUMENT: return "Invalid argument.";
        case INVALID_PENALTY_UPDATE_VALUE: return "Invalid penalty update value (must be > 1).";
        case INVALID_COMPLEMENTARITY_TOLERANCE: return "Invalid complementarity tolerance (must exceed machine precision).";
        case INVALID_INITIAL_PENALTY_VALUE: return "Invalid initial penalty parameter (must be positive).";
        case INVALID_MAX_ITERATIONS_VALUE: return "Invalid maximum number of iterations (must be a positive integer).";
        case INVALID_STATIONARITY_TOLERANCE: return "Invalid stationarity tolerance (must exceed machine precision).";
        case INVALID_NUMBER_OF_OPTIM_VARS: return "Invalid number of optimization variables (must be positive).";
        case INVALID_NUMBER_OF_COMP_VARS: return "Invalid number of complementarity pairs (must be positive).";
        case INVALID_NUMBER_OF_CONSTRAINT_VARS: return "Invalid number of linear constraints (must be non-negative).";
        case INVALID_QPSOLVER: return "Invalid QP solver.";
        case INVALID_OSQP_BOX_CONSTRAINTS: return "Box constraints cannot be used with the OSQP-style layout; pass them as linear constraints.";
        case INVALID_TOTAL_ITER_COUNT: return "Invalid total iteration delta (must be non-negative).";
        case INVALID_TOTAL_OUTER_ITER: return "Invalid outer iteration delta (must be non-negative).";
        case IVALID_SUBPROBLEM_ITER: return "Invalid subproblem iteration delta (must be non-negative).";
        case INVALID_RHO_OPT: return "Invalid penalty value at the solution (must be positive).";
        case INVALID_PRINT_LEVEL_VALUE: return "Invalid print level.";
        case INVALID_OBJECTIVE_LINEAR_TERM: return "Invalid objective linear term (null pointer).";
        case INVALID_CONSTRAINT_MATRIX: return "Invalid constraint matrix (null pointer but nC > 0).";
        case INVALID_COMPLEMENTARITY_MATRIX: return "Invalid complementarity matrix (null pointer).";
        case INVALID_ETA_VALUE: return "Invalid etaDynamicPenalty (must lie in (0,1)).";
        case INVALID_LOWER_COMPLEMENTARITY_BOUND: return "Lower complementarity bounds must be finite.";
        case INVALID_MAX_RHO_VALUE: return "Invalid maximal penalty value (must be positive).";
        case MAX_ITERATIONS_REACHED: return "Maximum number of iterations reached.";
        case MAX_PENALTY_REACHED: return "Maximum penalty value reached.";
        case INITIAL_SUBPROBLEM_FAILED: return "Failed to solve the initial QP.";
        case SUBPROBLEM_SOLVER_ERROR: return "An error occurred in the subproblem solver.";
        case FAILED_SYM_COMPLEMENTARITY_MATRIX: return "Failed to compute the symmetric complementarity matrix.";
        case FAILED_SWITCH_TO_SPARSE: return "Failed to switch to sparse mode.";
        case FAILED_SWITCH_TO_DENSE: return "Failed to switch to dense mode.";
        case OSQP_WORKSPACE_NOT_SET_UP: return "OSQP workspace is not set up.";
        case OSQP_INITIAL_PRIMAL_GUESS_FAILED: return "The initial primal guess could not be used.";
        case OSQP_INITIAL_DUAL_GUESS_FAILED: return "The initial dual guess could not be used.";
        case LCQPOBJECT_NOT_SETUP: return "The LCQP object has not been set up.";
        case INDEX_OUT_OF_BOUNDS: return "Index out of bounds.";
        case UNABLE_TO_READ_FILE: return "Unable to read file.";
        case INVALID_INDEX_POINTER: return "Invalid csc index pointer.";
        case INVALID_INDEX_ARRAY: return "Invalid csc index array.";
        case DENSE_SPARSE_MISSMATCH: return "Dense / sparse mismatch between the loaded data and the chosen method.";
        case CUDA_NO_DEVICE: return "No usable sm_100 CUDA device (this build has no CPU fallback).";
        case CUDA_BAD_HANDLE: return "Invalid CUDA solver handle.";
        case CUDA_BAD_ARGUMENT: return "Invalid argument passed to the CUDA solver.";
        case CUDA_OUT_OF_MEMORY: return "Out of device memory.";
        case CUDA_LAUNCH_FAILED: return "A CUDA kernel launch or copy failed.";
        case CUDA_NOT_LOADED: return "runSolver called before loadLCQP.";
        case CUDA_NOT_RUN: return "Results requested before runSolver.";
        case CUDA_TOO_LARGE: return "The problem does not fit the shared-memory budget of the CUDA solver.";
    }
    return "Unknown return value.";
}

ReturnValue MessageHandler::PrintMessage(ReturnValue ret, MessageType type)
{
    if (ret == SUCCESSFUL_RETURN) return ret;
    const char* tag = type == MESSAGE ? "MESSAGE" : (type == WARNING ? "WARNING" : "ERROR");
    fprintf(type == MESSAGE ? stdout : stderr, "[LCQPow-B200 %s %d] %s\n", tag, (int)ret, ReturnValueText(ret));
    return ret;
}

AlgorithmStatus MessageHandler::PrintSolution(AlgorithmStatus s)
{
    static const char* const names[] = {"problem not solved", "W-stationary solution", "C-stationary solution",
                                        "M-stationary solution", "S-stationary solution"};
    printf("LCQPow-B200: %s\n", names[(int)s >= 0 && (int)s <= 4 ? (int)s : 0]);
    return s;
}

}  // namespace LCQPow

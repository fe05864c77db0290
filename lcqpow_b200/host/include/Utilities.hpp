// Utilities.hpp -- enums, constants, the csc type and the small dense/CSC helpers of the LCQPow API,
// restated for the B200 build (host side; the solver itself lives in liblcqp_cuda.so).
//
// Mirrors /root/reference/include/Utilities.hpp: ReturnValue :37-87, AlgorithmStatus :103-109, PrintLevel
// :115-119, QPSolver :125-129 (+ CUDA_DENSE), class Utilities :139-362.  The csc layout is OSQP's
// (/root/reference/external/osqp/include/types.h:21-29) so that user code building csc matrices keeps working.
#ifndef LCQPOW_B200_UTILITIES_HPP
#define LCQPOW_B200_UTILITIES_HPP

#include <cstddef>

namespace LCQPow {

enum ReturnValue {
    NOT_YET_IMPLEMENTED = -1,
    SUCCESSFUL_RETURN = 0,
    // invalid arguments
    INVALID_ARGUMENT = 100,
    INVALID_PENALTY_UPDATE_VALUE = 101,
    INVALID_COMPLEMENTARITY_TOLERANCE = 102,
    INVALID_INITIAL_PENALTY_VALUE = 103,
    INVALID_MAX_ITERATIONS_VALUE = 104,
    INVALID_STATIONARITY_TOLERANCE = 105,
    INVALID_NUMBER_OF_OPTIM_VARS = 106,
    INVALID_NUMBER_OF_COMP_VARS = 107,
    INVALID_NUMBER_OF_CONSTRAINT_VARS = 108,
    INVALID_QPSOLVER = 109,
    INVALID_OSQP_BOX_CONSTRAINTS = 110,
    INVALID_TOTAL_ITER_COUNT = 111,
    INVALID_TOTAL_OUTER_ITER = 112,
    IVALID_SUBPROBLEM_ITER = 113,   // (sic) the reference's spelling
    INVALID_RHO_OPT = 114,
    INVALID_PRINT_LEVEL_VALUE = 115,
    INVALID_OBJECTIVE_LINEAR_TERM = 116,
    INVALID_CONSTRAINT_MATRIX = 117,
    INVALID_COMPLEMENTARITY_MATRIX = 118,
    INVALID_ETA_VALUE = 119,
    INVALID_LOWER_COMPLEMENTARITY_BOUND = 120,
    INVALID_MAX_RHO_VALUE = 121,
    // algorithmic
    MAX_ITERATIONS_REACHED = 200,
    MAX_PENALTY_REACHED = 201,
    INITIAL_SUBPROBLEM_FAILED = 202,
    SUBPROBLEM_SOLVER_ERROR = 203,
    FAILED_SYM_COMPLEMENTARITY_MATRIX = 204,
    FAILED_SWITCH_TO_SPARSE = 205,
    FAILED_SWITCH_TO_DENSE = 206,
    OSQP_WORKSPACE_NOT_SET_UP = 207,
    OSQP_INITIAL_PRIMAL_GUESS_FAILED = 208,
    OSQP_INITIAL_DUAL_GUESS_FAILED = 209,
    // generic
    LCQPOBJECT_NOT_SETUP = 300,
    INDEX_OUT_OF_BOUNDS = 301,
    UNABLE_TO_READ_FILE = 302,
    // sparse matrices
    INVALID_INDEX_POINTER = 400,
    INVALID_INDEX_ARRAY = 401,
    DENSE_SPARSE_MISSMATCH = 402,
    // CUDA side (include/lcqp_cuda.h; the reference stops at 402)
    CUDA_NO_DEVICE = 500,
    CUDA_BAD_HANDLE = 501,
    CUDA_BAD_ARGUMENT = 502,
    CUDA_OUT_OF_MEMORY = 503,
    CUDA_LAUNCH_FAILED = 504,
    CUDA_NOT_LOADED = 505,
    CUDA_NOT_RUN = 506,
    CUDA_TOO_LARGE = 507
};

enum MessageType { MESSAGE = 0, WARNING = 1, ERROR = 2 };

enum AlgorithmStatus {
    PROBLEM_NOT_SOLVED = 0,
    W_STATIONARY_SOLUTION = 1,
    C_STATIONARY_SOLUTION = 2,
    M_STATIONARY_SOLUTION = 3,
    S_STATIONARY_SOLUTION = 4
};

enum PrintLevel { NONE = 0, OUTER_LOOP_ITERATES = 1, INNER_LOOP_ITERATES = 2 };

// The first three values keep their meaning as a DUAL LAYOUT (qpOASES-style: nV box duals first;
// OSQP-style: constraint duals only, box bounds rejected); every value is served by SubsolverCUDA -- this
// build contains no CPU QP solver.  CUDA_DENSE is the explicit name of the device path (layout of QPOASES_DENSE).
enum QPSolver { QPOASES_DENSE = 0, QPOASES_SPARSE = 1, OSQP_SPARSE = 2, CUDA_DENSE = 3 };

// Compressed-sparse-column matrix, field for field OSQP's `csc`.
struct csc {
    int nzmax;  // allocated entries
    int m;      // rows
    int n;      // columns
    int* p;     // column pointers (n + 1)
    int* i;     // row indices
    double* x;  // values
    int nz;     // -1 for compressed-column form
};

class Utilities {
public:
    // dense, row-major (reference: src/Utilities.cpp:38-265)
    static void MatrixMultiplication(const double* A, const double* B, double* C, int m, int n, int p);
    static void MatrixMultiplication(const csc* A, const double* b, double* c);
    static void TransponsedMatrixMultiplication(const double* A, const double* B, double* C, int m, int n, int p);
    static void TransponsedMatrixMultiplication(const csc* A, const double* b, double* c);
    static void AddTransponsedMatrixMultiplication(const double* A, const double* B, double* C, int m, int n, int p);
    static void AddTransponsedMatrixMultiplication(const csc* A, const double* b, double* c);
    static void MatrixSymmetrizationProduct(const double* A, const double* B, double* C, int m, int n);
    static csc* MatrixSymmetrizationProduct(const csc* L, const csc* R);
    static void AffineLinearTransformation(double alpha, const double* A, const double* b, const double* c, double* d, int m, int n);
    static void AffineLinearTransformation(double alpha, const csc* S, const double* b, const double* c, double* d, int m);
    static void WeightedMatrixAdd(double alpha, const double* A, double beta, const double* B, double* C, int m, int n);
    static void WeightedVectorAdd(double alpha, const double* a, double beta, const double* b, double* c, int m);
    static double QuadraticFormProduct(const double* Q, const double* p, int m);
    static double QuadraticFormProduct(const csc* S, const double* p, int m);
    static double DotProduct(const double* a, const double* b, int m);
    static double MaxAbs(const double* a, int m);

    // csc helpers (src/Utilities.cpp:469-650)
    static csc* createCSC(int m, int n, int nnz, double* x, int* i, int* p);   // takes ownership of x, i, p (malloc'ed)
    static csc* copyCSC(int m, int n, int nnz, const double* x, const int* i, const int* p);
    static csc* copyCSC(const csc* M, bool toUpperTriangular = false);
    static void ClearSparseMat(csc* M);
    static void ClearSparseMat(csc** M);
    static double* csc_to_dns(const csc* sparse);               // new[]-allocated m*n row-major
    static csc* dns_to_csc(const double* full, int m, int n);

    // text files, one value per line (src/Utilities.cpp:312-395)
    static ReturnValue readFromFile(int* data, int n, const char* datafilename);
    static ReturnValue readFromFile(double* data, int n, const char* datafilename);
    static ReturnValue writeToFile(const double* data, int n, const char* datafilename);

    static void printMatrix(const double* A, int m, int n, const char* name);
    static void printMatrix(const csc* A, const char* name);

    static double getAbs(double x) { return x >= 0 ? x : -x; }
    static bool isEqual(double x, double y, double tol = ZERO) { return getAbs(x - y) <= tol; }
    static bool isZero(double x, double tol = ZERO) { return getAbs(x) <= tol; }
    static double getSign(double x) { return x < 0 ? -1.0 : 1.0; }
    static int getMax(int x, int y) { return x < y ? y : x; }
    static int getMin(int x, int y) { return x < y ? x : y; }
    static double getMax(double x, double y) { return x < y ? y : x; }
    static double getMin(double x, double y) { return x < y ? x : y; }
    template <typename P> static bool isNullPtr(P ptr) { return ptr == nullptr; }
    template <typename P> static bool isNotNullPtr(P ptr) { return ptr != nullptr; }

    constexpr static double EPS = 2.221e-16;    // Utilities.hpp:350
    constexpr static double ZERO = 1.0e-25;     // :356
    constexpr static double INFTY = 1.0e20;     // :362
    constexpr static unsigned MAX_STRING_LENGTH = 160;
};

// enum -> text (the reference's MessageHandler, src/MessageHandler.cpp:28-245, reduced to what callers use)
class MessageHandler {
public:
    static ReturnValue PrintMessage(ReturnValue ret, MessageType type = ERROR);
    static AlgorithmStatus PrintSolution(AlgorithmStatus algoStat);
    static const char* ReturnValueText(ReturnValue ret);
};

}  // namespace LCQPow

#endif

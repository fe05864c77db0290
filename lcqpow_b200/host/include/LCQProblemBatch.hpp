// LCQProblemBatch.hpp -- the batched front door in C++: B independent LCQPs of one shape, sharded by
// instance over one or several GPUs of a box (contiguous instance blocks, shared operands
// replicated per GPU, no collective on the data path, host-side gather of x / y / statistics).
//
// The API is LCQProblem's with a leading batch dimension: every array of loadLCQP holds `batch` consecutive
// copies unless its bit is set in `shared_mask` (bit k = argument k of loadLCQP, see include/lcqp_cuda.h).
#ifndef LCQPOW_B200_LCQPROBLEMBATCH_HPP
#define LCQPOW_B200_LCQPROBLEMBATCH_HPP

#include <string>
#include <vector>

#include "Options.hpp"
#include "OutputStatistics.hpp"
#include "Utilities.hpp"

namespace LCQPow {

class LCQProblemBatch {
public:
    // devices: CUDA ordinals to shard over (empty = device 0)
    LCQProblemBatch(int nV, int nC, int nComp, int batch, const std::vector<int>& devices = std::vector<int>());
    ~LCQProblemBatch();
    LCQProblemBatch(const LCQProblemBatch&) = delete;
    LCQProblemBatch& operator=(const LCQProblemBatch&) = delete;

    bool isValid() const { return valid; }
    ReturnValue setOptions(const Options& options);

    // host pointers; the library copies (callers keep ownership, as in the reference)
    ReturnValue loadLCQP(unsigned shared_mask, const double* Q, const double* g, const double* L, const double* R,
                         const double* lbL = 0, const double* ubL = 0, const double* lbR = 0, const double* ubR = 0,
                         const double* A = 0, const double* lbA = 0, const double* ubA = 0, const double* lb = 0,
                         const double* ub = 0, const double* x0 = 0, const double* y0 = 0);

    // LCQProblem::runSolver for every instance; returns SUCCESSFUL_RETURN when every shard ran (per-instance
    // return values: getReturnValues()).  All shards are launched before any is waited for.
    ReturnValue runSolver();

    // x: batch x nV;  y: batch x (nV + nC + 2 nComp), the first getNumberOfDuals() entries of a row are valid
    ReturnValue getPrimalSolution(double* x) const;
    ReturnValue getDualSolution(double* y) const;
    ReturnValue getOutputStatistics(std::vector<OutputStatistics>& stats) const;
    ReturnValue getRawStatistics(std::vector<lcqp_cuda_stats>& stats) const;
    std::vector<int> getReturnValues() const;
    int getNumberOfPrimals() const { return nV; }
    int getNumberOfDuals() const;
    int getBatchSize() const { return batch; }
    int getNumberOfShards() const { return (int)shards.size(); }
    long long getLaunchCount() const;
    const std::string& getLastError() const { return lastError; }

private:
    struct Shard {
        int device = 0;
        int first = 0;   // global index of the shard's first instance
        int count = 0;
        lcqp_cuda_handle handle = nullptr;
    };
    int nV, nC, nComp, batch;
    bool valid = false, ran = false;
    std::vector<Shard> shards;
    Options options;
    mutable std::string lastError;
    ReturnValue fail(int code, const Shard& s) const;
};

}  // namespace LCQPow

#endif

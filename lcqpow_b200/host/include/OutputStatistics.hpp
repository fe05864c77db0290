// OutputStatistics.hpp -- LCQPow::OutputStatistics for the B200 build.
// Same counters, update rules and tracking vectors as /root/reference/include/OutputStatistics.hpp:35-226
// (src/OutputStatistics.cpp:81-164).  fromCuda() fills the counters from the per-instance record the device
// loop writes (lcqp_cuda_stats, include/lcqp_cuda.h).
#ifndef LCQPOW_B200_OUTPUTSTATISTICS_HPP
#define LCQPOW_B200_OUTPUTSTATISTICS_HPP

#include <vector>

#include "Utilities.hpp"
#include "../../../include/lcqp_cuda.h"

namespace LCQPow {

class OutputStatistics {
public:
    OutputStatistics();
    OutputStatistics& operator=(const OutputStatistics& rhs);

    void reset();

    ReturnValue updateIterTotal(int delta_iter);
    ReturnValue updateIterOuter(int delta_iter);
    ReturnValue updateSubproblemIter(int delta_iter);
    ReturnValue updateRhoOpt(double _rho);
    ReturnValue updateSolutionStatus(AlgorithmStatus _status);
    ReturnValue updateQPSolverExitFlag(int _flag);
    ReturnValue updateTrackingVectors(double* thisxSteps, int thisInnerIter, int thisSubproblemIter, double thisStepLength,
                                      double thisStepSize, double statVal, double objVal, double phiVal, double meritVal, int nV);

    int getIterTotal() const;
    int getIterOuter() const;
    int getSubproblemIter() const;
    double getRhoOpt() const;
    AlgorithmStatus getSolutionStatus() const;
    int getQPSolverExitFlag() const;

    int* getInnerIters() const;
    std::vector<int> getInnerItersStdVec() const;
    int* getSubproblemIters() const;
    std::vector<int> getSubproblemItersStdVec() const;
    int* getAccuSubproblemIters() const;
    std::vector<int> getAccuSubproblemItersStdVec() const;
    double* getStepLength() const;
    std::vector<double> getStepLengthStdVec() const;
    double* getStepSize() const;
    std::vector<double> getStepSizeStdVec() const;
    double* getStatVals() const;
    std::vector<double> getStatValsStdVec() const;
    double* getObjVals() const;
    std::vector<double> getObjValsStdVec() const;
    double* getPhiVals() const;
    std::vector<double> getPhiValsStdVec() const;
    double* getMeritVals() const;
    std::vector<double> getMeritValsStdVec() const;
    std::vector<std::vector<double>> getxStepsStdVec() const;

    // counters of one instance of a device run
    void fromCuda(const lcqp_cuda_stats& rec);

private:
    int iterTotal = 0;
    int iterOuter = 0;
    int subproblemIter = 0;
    double rhoOpt = 0.0;
    AlgorithmStatus status = PROBLEM_NOT_SOLVED;
    int qpSolver_exit_flag = 0;

    std::vector<std::vector<double>> xSteps;
    std::vector<int> innerIters;
    std::vector<int> subproblemIters;
    std::vector<int> accuSubproblemIters;
    std::vector<double> stepLength;
    std::vector<double> stepSize;
    std::vector<double> statVals;
    std::vector<double> objVals;
    std::vector<double> phiVals;
    std::vector<double> meritVals;
};

}  // namespace LCQPow

#endif

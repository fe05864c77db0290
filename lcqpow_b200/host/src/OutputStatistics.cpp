// OutputStatistics.cpp -- see ../include/OutputStatistics.hpp (reference: src/OutputStatistics.cpp:28-377).
#include "OutputStatistics.hpp"

namespace LCQPow {

OutputStatistics::OutputStatistics() {}

OutputStatistics& OutputStatistics::operator=(const OutputStatistics& rhs)
{
    if (this == &rhs) return *this;
    iterTotal = rhs.iterTotal;
    iterOuter = rhs.iterOuter;
    subproblemIter = rhs.subproblemIter;
    rhoOpt = rhs.rhoOpt;
    status = rhs.status;
    qpSolver_exit_flag = rhs.qpSolver_exit_flag;
    xSteps = rhs.xSteps;
    innerIters = rhs.innerIters;
    subproblemIters = rhs.subproblemIters;
    accuSubproblemIters = rhs.accuSubproblemIters;
    stepLength = rhs.stepLength;
    stepSize = rhs.stepSize;
    statVals = rhs.statVals;
    objVals = rhs.objVals;
    phiVals = rhs.phiVals;
    meritVals = rhs.meritVals;
    return *this;
}

void OutputStatistics::reset()
{
    iterTotal = 0;
    iterOuter = 0;
    subproblemIter = 0;
    rhoOpt = 0.0;
    status = PROBLEM_NOT_SOLVED;
    qpSolver_exit_flag = 0;
    xSteps.clear();
    innerIters.clear();
    subproblemIters.clear();
    accuSubproblemIters.clear();
    stepLength.clear();
    stepSize.clear();
    statVals.clear();
    objVals.clear();
    phiVals.clear();
    meritVals.clear();
}

// negative increments are rejected, the counter keeps its value (reference :81-105)
ReturnValue OutputStatistics::updateIterTotal(int delta_iter)
{
    if (delta_iter < 0) return INVALID_TOTAL_ITER_COUNT;
    iterTotal += delta_iter;
    return SUCCESSFUL_RETURN;
}

ReturnValue OutputStatistics::updateIterOuter(int delta_iter)
{
    if (delta_iter < 0) return INVALID_TOTAL_OUTER_ITER;
    iterOuter += delta_iter;
    return SUCCESSFUL_RETURN;
}

ReturnValue OutputStatistics::updateSubproblemIter(int delta_iter)
{
    if (delta_iter < 0) return IVALID_SUBPROBLEM_ITER;
    subproblemIter += delta_iter;
    return SUCCESSFUL_RETURN;
}

ReturnValue OutputStatistics::updateRhoOpt(double _rho)
{
    if (_rho <= 0) return INVALID_RHO_OPT;
    rhoOpt = _rho;
    return SUCCESSFUL_RETURN;
}

ReturnValue OutputStatistics::updateSolutionStatus(AlgorithmStatus _status)
{
    status = _status;
    return SUCCESSFUL_RETURN;
}

ReturnValue OutputStatistics::updateQPSolverExitFlag(int _flag)
{
    qpSolver_exit_flag = _flag;
    return SUCCESSFUL_RETURN;
}

ReturnValue OutputStatistics::updateTrackingVectors(double* thisxSteps, int thisInnerIter, int thisSubproblemIter,
                                                    double thisStepLength, double thisStepSize, double statVal, double objVal,
                                                    double phiVal, double meritVal, int nV)
{
    xSteps.emplace_back(thisxSteps, thisxSteps + nV);
    innerIters.push_back(thisInnerIter);
    subproblemIters.push_back(thisSubproblemIter);
    accuSubproblemIters.push_back((accuSubproblemIters.empty() ? 0 : accuSubproblemIters.back()) + thisSubproblemIter);
    stepLength.push_back(thisStepLength);
    stepSize.push_back(thisStepSize);
    statVals.push_back(statVal);
    objVals.push_back(objVal);
    phiVals.push_back(phiVal);
    meritVals.push_back(meritVal);
    return SUCCESSFUL_RETURN;
}

int OutputStatistics::getIterTotal() const { return iterTotal; }
int OutputStatistics::getIterOuter() const { return iterOuter; }
int OutputStatistics::getSubproblemIter() const { return subproblemIter; }
double OutputStatistics::getRhoOpt() const { return rhoOpt; }
AlgorithmStatus OutputStatistics::getSolutionStatus() const { return status; }
int OutputStatistics::getQPSolverExitFlag() const { return qpSolver_exit_flag; }

int* OutputStatistics::getInnerIters() const { return const_cast<int*>(innerIters.data()); }
std::vector<int> OutputStatistics::getInnerItersStdVec() const { return innerIters; }
int* OutputStatistics::getSubproblemIters() const { return const_cast<int*>(subproblemIters.data()); }
std::vector<int> OutputStatistics::getSubproblemItersStdVec() const { return subproblemIters; }
int* OutputStatistics::getAccuSubproblemIters() const { return const_cast<int*>(accuSubproblemIters.data()); }
std::vector<int> OutputStatistics::getAccuSubproblemItersStdVec() const { return accuSubproblemIters; }
double* OutputStatistics::getStepLength() const { return const_cast<double*>(stepLength.data()); }
std::vector<double> OutputStatistics::getStepLengthStdVec() const { return stepLength; }
double* OutputStatistics::getStepSize() const { return const_cast<double*>(stepSize.data()); }
std::vector<double> OutputStatistics::getStepSizeStdVec() const { return stepSize; }
double* OutputStatistics::getStatVals() const { return const_cast<double*>(statVals.data()); }
std::vector<double> OutputStatistics::getStatValsStdVec() const { return statVals; }
double* OutputStatistics::getObjVals() const { return const_cast<double*>(objVals.data()); }
std::vector<double> OutputStatistics::getObjValsStdVec() const { return objVals; }
double* OutputStatistics::getPhiVals() const { return const_cast<double*>(phiVals.data()); }
std::vector<double> OutputStatistics::getPhiValsStdVec() const { return phiVals; }
double* OutputStatistics::getMeritVals() const { return const_cast<double*>(meritVals.data()); }
std::vector<double> OutputStatistics::getMeritValsStdVec() const { return meritVals; }
std::vector<std::vector<double>> OutputStatistics::getxStepsStdVec() const { return xSteps; }

void OutputStatistics::fromCuda(const lcqp_cuda_stats& rec)
{
    reset();
    iterTotal = rec.iterTotal;
    iterOuter = rec.iterOuter;
    subproblemIter = rec.subproblemIter;
    rhoOpt = rec.rhoOpt;
    status = (AlgorithmStatus)rec.status;
    qpSolver_exit_flag = rec.qpExitFlag;
}

}  // namespace LCQPow

"""lcqpow_b200 -- B200-native batched LCQP solver behind LCQPow's API for LCQProblem::runSolver and the
QP subsolver plugin.  The only implementation is the sm_100a CUDA library
lcqpow_b200/lib/liblcqp_cuda.so (C ABI: include/lcqp_cuda.h); there is no CPU fallback."""
from .api import (LCQProblem, LCQProblemBatch, Options, OutputStatistics, SubsolverCUDA, LCQPError, load_library,  # noqa: F401
                  SUCCESSFUL_RETURN, MAX_ITERATIONS_REACHED, MAX_PENALTY_REACHED, SUBPROBLEM_SOLVER_ERROR,
                  PROBLEM_NOT_SOLVED, W_STATIONARY_SOLUTION, C_STATIONARY_SOLUTION, M_STATIONARY_SOLUTION,
                  S_STATIONARY_SOLUTION, QPOASES_DENSE, QPOASES_SPARSE, OSQP_SPARSE)
from . import problems  # noqa: F401

"""Section profile of the PAS kernel (-DLCQP_PROFILE build selected by LCQP_CUDA_LIB).  Development aid."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["LCQP_CUDA_VERBOSE"] = "1"
import numpy as np
import lcqpow_b200 as L
from lcqpow_b200 import problems as P
which = sys.argv[1] if len(sys.argv) > 1 else "circle"
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
pb = P.circle_batch_fast(nb) if which == "circle" else P.dense_random_batch(nb)
prob = L.LCQProblemBatch(pb.nV, pb.nC, pb.nComp, pb.batch)
o = L.Options()
if which == "circle": o.setStationarityTolerance(10e-3)
prob.setOptions(o)
prob.loadBatch(pb)
for _ in range(2):
    prob.runSolver()
st = prob.getOutputStatistics()
ms = prob.lastRunMs()
print(which, "batch", nb, "solved", float((st['ret'] == 0).mean()), "kernel ms", ms, "LCQP/s %.0f" % (nb / (ms[0] * 1e-3)), "solves/LCQP %.1f sub %.1f" % (st['kktSolves'].mean(), st['subproblemIter'].mean()))

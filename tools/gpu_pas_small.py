"""Small runs of the PAS path for compute-sanitizer (development aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import lcqpow_b200 as L
from lcqpow_b200 import problems as P
which = sys.argv[1]; nb = int(sys.argv[2])
pb = P.circle_batch_fast(nb, N=int(sys.argv[3]) if len(sys.argv) > 3 else 100) if which == "circle" else P.dense_random_batch(nb)
prob = L.LCQProblemBatch(pb.nV, pb.nC, pb.nComp, pb.batch)
o = L.Options()
if which == "circle": o.setStationarityTolerance(10e-3)
prob.setOptions(o)
prob.loadBatch(pb)
prob.runSolver()
st = prob.getOutputStatistics()
print(which, nb, "ret", np.unique(st['ret'], return_counts=True), "k", st['iterOuter'][:8], "i", st['iterTotal'][:8])

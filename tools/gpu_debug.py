"""GPU debugging aid: run the small fixtures and print the per-instance records (not a test)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import lcqpow_b200 as L
from lcqpow_b200 import problems as P

def run(pb, **over):
    prob = L.LCQProblemBatch(pb.nV, pb.nC, pb.nComp, pb.batch)
    o = L.Options()
    o.setPerturbStep(False)
    for k, v in over.items():
        getattr(o, "set" + k[0].upper() + k[1:])(v)
    prob.setOptions(o)
    assert prob.loadBatch(pb) == 0
    prob.runSolver()
    st = prob.getOutputStatistics()
    x = prob.getPrimalSolution()
    print(pb.name, "launch", prob.lastLaunchInfo(), "ms", prob.lastRunMs())
    for b in range(min(pb.batch, 4)):
        print("  ", {k: st[k][b].item() for k in st.dtype.names}, x[b][:4])

run(P.warm_up())
run(P.dense_random_batch(4))
run(P.circle_batch(2), stationarityTolerance=10e-3)

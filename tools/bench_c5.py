"""Throughput of config C5 (per-instance dense LCQPs, n=64, nComp=32, nC=16) -- a development
measurement, not the bench line (bench.py measures C2).  usage: python tools/bench_c5.py [batch]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lcqpow_b200 as L
from lcqpow_b200 import problems as P

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
pb = P.dense_random_batch(batch)
prob = L.LCQProblemBatch(pb.nV, pb.nC, pb.nComp, pb.batch)
o = L.Options()
assert prob.setOptions(o) == 0
assert prob.loadBatch(pb) == 0
for rep in range(3):
    t = time.time()
    prob.runSolver()
    x = prob.getPrimalSolution()
    dt = time.time() - t
    st = prob.getOutputStatistics()
    print(f"C5 batch {batch}: {batch / dt:.0f} LCQP/s wall (incl. D2H), solved {np.mean(st['ret'] == 0):.4f}, "
          f"mean iterTotal {st['iterTotal'].mean():.1f}, kkt solves/LCQP {st['kktSolves'].mean():.1f}")

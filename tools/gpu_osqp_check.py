"""Development aid: the OSQP flavour on the GPU -- timing at the configurations' shapes (C2 circle, C5 dense, C4 sparse)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lcqpow_b200 as L
from lcqpow_b200 import problems as P

def run(pb, over, tag, interval=0):
    prob = L.LCQProblemBatch(pb.nV, pb.nC, pb.nComp, pb.batch, device=0)
    o = L.Options()
    for k, v in over.items():
        getattr(o, "set" + k[0].upper() + k[1:])(v)
    o.setOSQPADMM(True, adaptive_rho_interval=interval)
    if os.environ.get("MAXIT"):
        o.setMaxIterations(int(os.environ["MAXIT"]))
    assert prob.setOptions(o) == 0
    t = time.time()
    if isinstance(pb, P.SparseLCQPBatch):
        rc = prob.loadCSC(pb.Q, pb.g, pb.L, pb.R, A=pb.A, lbA=pb.lbA, ubA=pb.ubA, batch=pb.batch, shared=tuple(pb.shared))
    else:
        rc = prob.loadBatch(pb)
    assert rc == 0, prob._err()
    tl = time.time() - t
    for rep in range(int(os.environ.get("REPS", 1))):
        prob.runSolver()
        ms, _ = prob.lastRunMs()
    st = prob.getOutputStatistics()
    ok = (st["ret"] == 0).mean()
    print(f"{tag}: batch {pb.batch} load {tl:.2f}s kernel {ms:.1f} ms -> {pb.batch / ms * 1e3:.0f} LCQP/s ({ok * pb.batch / ms * 1e3:.0f} solved/s), solved {ok:.3f}, "
          f"ret {dict(zip(*np.unique(st['ret'], return_counts=True)))}, mean k {st['iterOuter'].mean():.1f} i {st['iterTotal'].mean():.1f} "
          f"ADMM/LCQP {st['admmIters'].mean():.0f} factorisations/LCQP {st['kktSolves'].mean():.1f}", flush=True)
    return prob, st

if __name__ == "__main__":
    which = sys.argv[1:] or ["c5", "c2", "c4"]
    if "c5" in which:
        run(P.dense_random_batch(int(os.environ.get("C5_BATCH", 16384))), {}, "C5 dense n=64")
    if "c2" in which:
        run(P.circle_batch_fast(int(os.environ.get("C2_BATCH", 32768))), {"stationarityTolerance": 10e-3}, "C2 circle")
    if "c4" in which:
        run(P.sparse_banded_batch(int(os.environ.get("C4_BATCH", 4096))), {}, "C4 sparse n=1000")

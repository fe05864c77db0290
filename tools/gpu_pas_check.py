"""GPU check of the parametric active-set path: parity against the unmodified reference (oracle/_ref) on fresh
instances + a quick throughput number.  Development aid (run under gpurun)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import lcqpow_b200 as L
from lcqpow_b200 import problems as P
from oracle import pyref


def solve(pb, over, perturb=0):
    prob = L.LCQProblemBatch(pb.nV, pb.nC, pb.nComp, pb.batch)
    o = L.Options()
    o.setPerturbStep(bool(perturb))
    for k, v in over.items():
        assert getattr(o, "set" + k[0].upper() + k[1:])(v) == 0
    assert prob.setOptions(o) == 0
    t = time.time()
    assert prob.loadBatch(pb) == 0
    tl = time.time() - t
    t = time.time()
    prob.runSolver()
    tr = time.time() - t
    return prob.getPrimalSolution(), prob.getDualSolution(), prob.getOutputStatistics(), prob, tl, tr


def cmp(pb, over, refsolver):
    ref = pyref.RefLib()
    sr = ref.solve_batch(pb, ref.default_options(perturbStep=0, qpSolver=refsolver, **over))
    x, y, st, prob, tl, tr = solve(pb, over)
    bad = 0
    for b in range(pb.batch):
        r = sr.res[b]
        dx = np.abs(sr.x[b] - x[b]).max()
        ok = r['ret'] == st['ret'][b] and r['status'] == st['status'][b] and r['iterOuter'] == st['iterOuter'][b] and r['iterTotal'] == st['iterTotal'][b] and (dx < 1e-6 or r['ret'] != 0)
        if not ok:
            bad += 1
            if bad < 10:
                print("  MISMATCH", pb.name, b, "ref", r['ret'], r['status'], r['iterOuter'], r['iterTotal'], "cuda", st['ret'][b], st['status'][b], st['iterOuter'][b], st['iterTotal'][b], st['qpExitFlag'][b], "dx %.2e" % dx)
    print(pb.name, "batch", pb.batch, "bad", bad, "load %.3fs run %.3fs" % (tl, tr), "launch", prob.lastLaunchInfo(), "kernel ms", prob.lastRunMs(), flush=True)
    return bad


if __name__ == "__main__":
    os.environ.setdefault("LCQP_CUDA_VERBOSE", "1")
    nb = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    cmp(P.warm_up_binary(), {}, 0)
    cmp(P.circle_batch_fast(nb), {"stationarityTolerance": 10e-3}, 1)
    cmp(P.dense_random_batch(nb), {}, 0)
    for big in (4096, 32768):
        pb = P.circle_batch_fast(big)
        x, y, st, prob, tl, tr = solve(pb, {"stationarityTolerance": 10e-3}, perturb=1)
        ms = prob.lastRunMs()
        print("circle perf batch", big, "solved", float((st['ret'] == 0).mean()), "kernel ms", ms, "LCQP/s %.0f" % (big / (ms[0] * 1e-3)), "solves/LCQP %.1f" % st['kktSolves'].mean(), "sub %.1f" % st['subproblemIter'].mean(), flush=True)
    pb = P.dense_random_batch(4096)
    x, y, st, prob, tl, tr = solve(pb, {}, perturb=1)
    ms = prob.lastRunMs()
    print("dense perf batch 4096 solved", float((st['ret'] == 0).mean()), "kernel ms", ms, "LCQP/s %.0f" % (4096 / (ms[0] * 1e-3)), flush=True)

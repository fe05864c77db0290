"""Parity tests proper (run on the B200 box: `pytest -m gpu`).  Everything goes through the C ABI
(include/lcqp_cuda.h) via lcqpow_b200.api; the checker is the committed golden output of the real reference
(tests/golden/reference_outputs.npz) and the plain-C oracle."""
import numpy as np
import pytest

from conftest import check_against_golden, golden_cases

pytestmark = pytest.mark.gpu


def _solve_cuda(pb, over, perturb=0, qp_solver=0):
    import lcqpow_b200 as L
    prob = L.LCQProblemBatch(pb.nV, pb.nC, pb.nComp, pb.batch)
    o = L.Options()
    o.setPerturbStep(bool(perturb))
    o.setQPSolver(qp_solver)
    for k, v in over.items():
        assert getattr(o, "set" + k[0].upper() + k[1:])(v) == 0
    assert prob.setOptions(o) == 0
    assert prob.loadBatch(pb) == 0
    prob.runSolver()
    return prob.getPrimalSolution(), prob.getDualSolution(), prob.getOutputStatistics(), prob


@pytest.mark.parametrize("name", ["warm_up", "warm_up_noguess", "warm_up_w_A", "warm_up_binary", "warm_up_shifted",
                                  "infeasible_qp", "max_penalty", "circle", "dense", "example_data"])
def test_cuda_matches_reference_golden(name, golden, example_data):
    pb, over = golden_cases(example_data)[name]
    x, y, st, _ = _solve_cuda(pb, over)
    check_against_golden(name, x, y, st, golden)


@pytest.mark.parametrize("name", ["warm_up", "warm_up_binary", "circle", "dense", "example_data"])
def test_cuda_matches_oracle(name, oracle, example_data):
    pb, over = golden_cases(example_data)[name]
    x, y, st, _ = _solve_cuda(pb, over)
    so = oracle.solve_batch(pb, oracle.default_options(perturbStep=0, **over))
    for b in range(pb.batch):
        for f in ("ret", "status", "iterOuter", "iterTotal"):
            assert int(st[f][b]) == int(so.res[f][b]), (name, b, f)
        assert np.abs(x[b] - so.x[b]).max() <= 1e-8 * max(1.0, np.abs(so.x[b]).max()), (name, b)


def test_warm_up_with_perturbation_orbit():
    """test/RunUnitTests.cpp:505-551: with perturbStep (default) the solution is (1,0) or (0,1) and
    2x_i - 2 - y_i - y_{2+i} = 0; k = 15 in every reference run."""
    import lcqpow_b200 as L
    from lcqpow_b200 import problems as P
    pb = P.warm_up(False)
    for seed in range(1, 21):
        prob = L.LCQProblemBatch(2, 0, 1, 1)
        o = L.Options()
        o.setPerturbSeed(seed)
        prob.setOptions(o)
        assert prob.loadBatch(pb) == 0
        prob.runSolver()
        st = prob.getOutputStatistics()[0]
        x = prob.getPrimalSolution()[0]
        y = prob.getDualSolution()[0]
        tol = o.getStationarityTolerance()
        assert int(st["ret"]) == 0 and int(st["status"]) == 4
        assert (abs(x[0] - 1) <= tol and abs(x[1]) <= tol) or (abs(x[1] - 1) <= tol and abs(x[0]) <= tol)
        assert abs(2 * x[0] - 2 - y[0] - y[2]) <= tol and abs(2 * x[1] - 2 - y[1] - y[3]) <= tol
        assert int(st["iterOuter"]) in (15, 16)


def test_osqp_style_dual_layout(golden, example_data):
    """qpSolver = OSQP_SPARSE: nDuals = nC + 2 nComp, no box duals (LCQProblem.cpp:934-935); box bounds are
    rejected with INVALID_OSQP_BOX_CONSTRAINTS (:955-957)."""
    pb, over = golden_cases(example_data)["dense"]
    x, y, st, prob = _solve_cuda(pb, over, qp_solver=2)
    assert prob.getNumberOfDuals() == pb.nC + 2 * pb.nComp
    x0, y0, st0, _ = _solve_cuda(pb, over, qp_solver=0)
    assert np.array_equal(st["iterTotal"], st0["iterTotal"])
    assert np.abs(y[:, : pb.nC + 2 * pb.nComp] - y0[:, pb.nV:]).max() <= 1e-9
    pbx, overx = golden_cases(example_data)["example_data"]
    _, _, stx, _ = _solve_cuda(pbx, overx, qp_solver=2)
    assert int(stx["ret"][0]) == 110



def test_circle_family_trajectories_match_oracle(oracle):
    """128 seeded instances of the C2 family with perturbStep on (the shipped default): the CUDA path and the
    oracle must walk the same penalty homotopy -- identical ReturnValue, stationarity type, outer and total
    iteration counts for every instance -- and agree in x to 1e-8.  (The oracle is pinned against the real
    reference on this family by tests/test_oracle_parity.py.)  This is the sensitive detector for changes of
    summation order in the device kernels."""
    from lcqpow_b200 import problems as P
    pb = P.circle_batch(128)
    over = {"stationarityTolerance": 10e-3}
    x, y, st, _ = _solve_cuda(pb, over, perturb=1)
    so = oracle.solve_batch(pb, oracle.default_options(perturbStep=1, **over))
    for f in ("ret", "status", "iterOuter", "iterTotal"):
        bad = np.nonzero(np.asarray(st[f]) != np.asarray(so.res[f]))[0]
        assert bad.size == 0, (f, bad.tolist())
    assert np.abs(x - so.x).max() <= 1e-8 * max(1.0, np.abs(so.x).max())


def test_circle_large_batch_properties():
    """C2 at a bench-like size (4096 instances, the vectorised generator of bench.py) through size-independent
    properties: every instance ends SUCCESSFUL_RETURN and S-stationary (as every reference qpOASES run of this
    family does), satisfies the circle constraints A x = 1 and the complementarity of each pair, instance 0 is
    the shipped solution (examples/OptimizeOnCircle.cpp: x = (0.1811, -0.9835)), and duplicated instances give
    bit-identical results wherever they sit in the batch (groups pull instances dynamically)."""
    from lcqpow_b200 import problems as P
    pb = P.circle_batch_fast(4096).normalised()
    g = pb.g.copy(); x0 = pb.x0.copy()
    g[4095] = g[7]; x0[4095] = x0[7]          # a duplicate far away in the batch
    import dataclasses
    pb = dataclasses.replace(pb, g=g, x0=x0)
    x, y, st, _ = _solve_cuda(pb, {"stationarityTolerance": 10e-3}, perturb=0)
    assert (st["ret"] == 0).all() and (st["status"] == 4).all()
    assert abs(x[0, 0] - 0.181110968) < 1e-6 and abs(x[0, 1] + 0.983483383) < 1e-6
    A = pb.A.reshape(pb.nC, pb.nV)
    assert np.abs(x @ A.T - 1.0).max() <= 1e-8
    u, v = x[:, 2::2], x[:, 3::2]
    assert (u >= -1e-9).all() and (v >= -1e-9).all() and np.abs(u * v).sum(axis=1).max() < 1e-9
    assert np.array_equal(x[7], x[4095]) and int(st["iterTotal"][7]) == int(st["iterTotal"][4095])

def test_large_batch_properties():
    """Full-size behaviour through size-independent properties: every solved instance is complementary
    (phi < tol), feasible, and a re-run is bit-identical (deterministic reductions)."""
    from lcqpow_b200 import problems as P
    pb = P.dense_random_batch(512)
    x, y, st, _ = _solve_cuda(pb, {})
    x2, y2, st2, _ = _solve_cuda(pb, {})
    assert np.array_equal(x, x2) and np.array_equal(st["iterTotal"], st2["iterTotal"])
    ok = st["ret"] == 0
    assert ok.mean() > 0.95
    n, p = pb.nV, pb.nComp
    Lx = x[:, :p]
    Rx = x[:, p:2 * p]
    assert (np.abs(Lx * Rx).sum(axis=1)[ok] < 1e-9).all()
    assert (Lx[ok] > -1e-9).all() and (Rx[ok] > -1e-9).all()
    pbn = pb.normalised()
    Ax = np.einsum("bij,bj->bi", pbn.A.reshape(pb.batch, pb.nC, n), x)
    assert (Ax[ok] >= pbn.lbA[ok] - 1e-8).all() and (Ax[ok] <= pbn.ubA[ok] + 1e-8).all()


def test_plugin_door_matches_batched_qp():
    """SubsolverCUDA (SubsolverBase::solve semantics): initial solve then hot start with a new gradient."""
    import lcqpow_b200 as L
    rng = np.random.default_rng(7)
    n, m = 12, 9
    M = rng.standard_normal((n, n))
    Q = M.T @ M / n + 0.1 * np.eye(n)
    A = rng.standard_normal((m, n))
    xs = rng.standard_normal(n)
    lbA = A @ xs - rng.uniform(0.1, 1.0, m)
    ubA = A @ xs + rng.uniform(0.1, 1.0, m)
    sub = L.SubsolverCUDA(n, m, Q, A)
    for it, g in enumerate([rng.standard_normal(n), rng.standard_normal(n) * 3]):
        rc, iters, flag = sub.solve(it == 0, g, lbA, ubA, x0=np.zeros(n))
        assert rc == 0 and flag == 0
        x, y = sub.getSolution()
        yA = y[n:]
        # KKT of min 1/2 x'Qx + g'x s.t. lbA <= Ax <= ubA with qpOASES signs: Qx + g = A'y
        assert np.abs(Q @ x + g - A.T @ yA).max() <= 1e-8
        Ax = A @ x
        assert (Ax >= lbA - 1e-9).all() and (Ax <= ubA + 1e-9).all()
        assert (yA[Ax > lbA + 1e-7] <= 1e-9).all() and (yA[Ax < ubA - 1e-7] >= -1e-9).all()

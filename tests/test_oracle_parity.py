"""Pins the checkers.  CPU only.

* oracle/pas_oracle.py -- the numpy restatement of the path (runSolver + qpOASES' online active-set strategy):
  against the committed outputs of the UNMODIFIED reference (tests/golden/*.npz) and, where oracle/_ref is present,
  live against the reference on fresh seeds.
* oracle/lcqp_oracle.c -- the plain-C restatement of round 1 (the loop, the Utilities kernels, a primal active-set QP
  solver): its Utilities kernels are pinned by tests/test_oracle_utilities.py; its trajectories are pinned here on the
  fixtures whose QPs have unique multipliers (on degenerate QPs a primal active-set method may return another
  optimal vertex than qpOASES' homotopy -- pas_oracle.py is the trajectory checker for those).
"""
import numpy as np
import pytest

from conftest import check_against_golden, golden_cases


def _as_stats(so, n):
    st = {k: so[k] for k in ("ret", "status", "iterOuter", "iterTotal")}
    st["qpExitFlag"] = np.where(so["ret"] == 203, 1, 0)
    st["rhoOpt"] = None
    return st


@pytest.mark.parametrize("name", ["warm_up_binary", "infeasible_qp", "max_penalty", "circle", "dense"])
def test_pas_oracle_matches_reference_golden(name, golden, example_data):
    from oracle import pas_oracle
    pb, over = golden_cases(example_data)[name]
    so = pas_oracle.solve_batch(pb, **over)
    g = {k.split("/", 2)[2]: v for k, v in golden.items() if k.startswith(name + "/qpoases/")}
    from conftest import ROUNDOFF_DECIDED
    for b in range(pb.batch):
        if (name, b) in ROUNDOFF_DECIDED:
            assert so["ret"][b] == 0 and so["status"][b] == 4
            continue
        tag = (name, b)
        assert so["ret"][b] == g["ret"][b] and so["status"][b] == g["status"][b], tag
        if g["ret"][b] == 203:
            continue
        assert so["iterOuter"][b] == g["iterOuter"][b] and so["iterTotal"][b] == g["iterTotal"][b], tag
        assert np.abs(so["x"][b] - g["x"][b]).max() <= 1e-6 * max(1.0, np.abs(g["x"][b]).max()), tag


def test_pas_oracle_counts_the_references_working_set_changes(golden, example_data):
    """On the dense family (no equality rows, nothing eliminated) the restated homotopy takes exactly the reference's
    number of working-set changes: subproblemIter is identical instance by instance."""
    from oracle import pas_oracle
    pb, over = golden_cases(example_data)["dense"]
    so = pas_oracle.solve_batch(pb, **over)
    assert np.array_equal(so["subproblemIter"], golden["dense/qpoases/subproblemIter"])


@pytest.mark.parametrize("family,count", [("dense", 16), ("circle", 6)])
def test_pas_oracle_matches_live_reference(family, count, reflib):
    """Fresh seeds (not the golden ones) against the reference itself, where it is built."""
    from lcqpow_b200 import problems as P
    from oracle import pas_oracle
    if family == "dense":
        pb, over = P.dense_random_batch(count, seed0=61000), {}
    else:
        pb, over = P.circle_batch(count + 1, seed0=27000).slice(1, count + 1), {"stationarityTolerance": 10e-3}
    r = reflib.solve_batch(pb, reflib.default_options(perturbStep=0, qpSolver=0, **over))
    so = pas_oracle.solve_batch(pb, **over)
    for b in range(pb.batch):
        assert all(int(r.res[f][b]) == int(so[f][b]) for f in ("ret", "status", "iterOuter", "iterTotal")), (family, b)
        assert np.abs(r.x[b] - so["x"][b]).max() <= 1e-6 * max(1.0, np.abs(r.x[b]).max()), (family, b)


@pytest.mark.parametrize("name", ["warm_up_binary", "warm_up_shifted", "infeasible_qp", "max_penalty", "dense", "example_data"])
def test_c_oracle_matches_reference_golden(name, oracle, golden, example_data):
    pb, over = golden_cases(example_data)[name]
    s = oracle.solve_batch(pb, oracle.default_options(perturbStep=0, **over))
    check_against_golden(name, s.x, s.y, s.res, golden)


def test_golden_table_of_survey(golden):
    """SURVEY.md section 4 golden table (qpOASES runs, perturbStep off where it matters)."""
    g = golden
    assert g["circle/qpoases/iterOuter"][0] == 8 and g["circle/qpoases/iterTotal"][0] == 19
    assert g["circle/osqp/iterOuter"][0] == 10 and g["circle/osqp/iterTotal"][0] == 39
    np.testing.assert_allclose(g["circle/qpoases/x"][0][:2], [0.181110968, -0.983483383], atol=1e-8)
    assert g["example_data/qpoases/iterOuter"][0] == 8 and g["example_data/qpoases/iterTotal"][0] == 34
    np.testing.assert_allclose(g["example_data/qpoases/x"][0][:3], [-1.48, -1.36, 0.0], atol=1e-9)
    assert g["warm_up_binary/qpoases/iterOuter"][0] == 9 and g["warm_up_binary/qpoases/iterTotal"][0] == 21
    assert g["max_penalty/qpoases/ret"][0] == 201          # test/examples/test_max_penalty.cpp:75-79
    assert g["infeasible_qp/qpoases/ret"][0] == 203        # test/RunUnitTests.cpp:463-502
    assert g["infeasible_qp/qpoases/qpExitFlag"][0] != 0


def test_family_goldens_are_what_the_bench_solves(families):
    """The committed families: 256 bench-family circle instances and 256 dense instances, all S-stationary in the
    reference's qpOASES run; the perturbed example_data family ends MAX_PENALTY_REACHED (201) except instance 0."""
    assert len(families["circle_bench/ret"]) == 256 and (families["circle_bench/ret"] == 0).all() and (families["circle_bench/status"] == 4).all()
    assert len(families["dense_bench/ret"]) == 256 and (families["dense_bench/ret"] == 0).all()
    assert families["example_data_family/ret"][0] == 0 and (families["example_data_family/ret"][1:] == 201).all()
    np.testing.assert_allclose(families["circle_bench/x"][0][:2], [0.181110968, -0.983483383], atol=1e-8)


def test_c_oracle_osqp_layout(oracle, example_data):
    """OSQP-style dual layout: nDuals = nC + 2 nComp; box bounds are rejected (LCQProblem.cpp:934-957)."""
    pb, over = golden_cases(example_data)["dense"]
    a = oracle.solve_batch(pb, oracle.default_options(perturbStep=0, qpSolver=0))
    b = oracle.solve_batch(pb, oracle.default_options(perturbStep=0, qpSolver=2))
    assert (b.res["nDuals"] == pb.nC + 2 * pb.nComp).all()
    assert np.array_equal(a.res["iterTotal"], b.res["iterTotal"])
    np.testing.assert_allclose(b.y[:, : pb.nC + 2 * pb.nComp], a.y[:, pb.nV:], atol=1e-12)
    pbx, _ = golden_cases(example_data)["example_data"]
    c = oracle.solve_batch(pbx, oracle.default_options(perturbStep=0, qpSolver=2))
    assert int(c.res["ret"][0]) == 110

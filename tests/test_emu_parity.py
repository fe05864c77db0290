"""The product's DEVICE code, compiled as single-threaded C++ (tests/emu), against the committed outputs of the real
reference and against the numpy oracle -- CPU coverage of the kernel logic (the `-m gpu` tests repeat these through
the C ABI on the B200)."""
import numpy as np
import pytest

from conftest import check_against_golden, check_family, check_family_regularised, family_cases, golden_cases


@pytest.fixture(scope="module")
def emu():
    from emu import EmuLib
    return EmuLib()


@pytest.mark.parametrize("name", ["warm_up", "warm_up_noguess", "warm_up_w_A", "warm_up_binary", "warm_up_shifted",
                                  "infeasible_qp", "max_penalty", "circle", "dense", "example_data"])
def test_emu_matches_reference_golden(name, emu, golden, example_data):
    pb, over = golden_cases(example_data)[name]
    s = emu.solve_batch(pb, emu.default_options(perturbStep=0, **over))
    check_against_golden(name, s.x, s.y, s.res, golden)


@pytest.mark.parametrize("name", ["circle_bench", "dense_bench", "circle_N20", "dense_n32"])
def test_emu_matches_reference_families(name, emu, families, example_data):
    """All 256 + 256 + 16 + 32 instances walk the reference's trajectory (ReturnValue, stationarity type, outer and
    total iteration counts) and end at its x to 1e-6 -- no instance is exempt."""
    pb, over = family_cases(example_data)[name]
    s = emu.solve_batch(pb, emu.default_options(perturbStep=0, **over))
    check_family(name, s.x, s.res, families)


def test_emu_counts_the_references_working_set_changes(emu, families, example_data):
    """Dense family (nothing eliminated): the number of working-set changes per LCQP equals the reference's nWSR sum."""
    pb, over = family_cases(example_data)["dense_bench"]
    s = emu.solve_batch(pb, emu.default_options(perturbStep=0, **over))
    assert np.array_equal(s.res["subproblemIter"], families["dense_bench/subproblemIter"])


def test_emu_example_data_family(emu, families, example_data):
    """The shipped example_data instance (semidefinite Hessian: the regularised solver) matches the reference; the
    perturbed instances end in a terminal failure as in the reference (201 there; 203 here -- DESIGN.md, known gap)."""
    pb, over = family_cases(example_data)["example_data_family"]
    s = emu.solve_batch(pb.slice(0, 4), emu.default_options(perturbStep=0, **over))
    check_family("example_data_family", s.x[:1], {k: s.res[k][:1] for k in ("ret", "status", "iterOuter", "iterTotal")}, families, subset=[0])
    assert (s.res["ret"][1:] != 0).all() and (families["example_data_family/ret"][1:4] != 0).all()


def test_emu_example_data_bounds_family(emu, families, example_data):
    """Config C3 as benchmarked: example_data with perturbed g and lbA = ubA (box bounds as shipped).  The reference solves
    all 64 instances; so does the device code, with the reference's ReturnValue, stationarity type, penalty updates and x
    on every instance and its total iteration count on at least 9 of 10."""
    pb, over = family_cases(example_data)["example_data_bounds"]
    s = emu.solve_batch(pb.slice(0, 32), emu.default_options(perturbStep=0, **over))
    off = check_family_regularised("example_data_bounds", s.x, s.res, families, subset=range(32))
    assert off <= 3, off


@pytest.mark.parametrize("perturb", [0, 1])
def test_emu_matches_numpy_oracle_on_seeded_batches(perturb, emu):
    """Same trajectory and x to 1e-8 as the numpy oracle on fresh seeds, with and without perturbStep (the two share
    the counter-based perturbation generator)."""
    from lcqpow_b200 import problems as P
    from oracle import pas_oracle
    for pb, over in ((P.circle_batch(5, seed0=23000), {"stationarityTolerance": 10e-3}), (P.dense_random_batch(12, seed0=52000), {})):
        se = emu.solve_batch(pb, emu.default_options(perturbStep=perturb, **over))
        so = pas_oracle.solve_batch(pb, perturb=perturb, **over)
        for f in ("ret", "status", "iterOuter", "iterTotal"):
            assert np.array_equal(se.res[f], so[f]), (pb.name, f, se.res[f], so[f])
        assert np.abs(se.x - so["x"]).max() <= 1e-8 * max(1.0, np.abs(so["x"]).max()), pb.name


def test_emu_mixed_row_types_in_one_batch(emu):
    """A batch that shares Q/A/L/R but not the bounds: a row that is an equality in one instance and an inequality in
    another (ADVICE of round 1).  Every instance must give the result it gives alone, in either order."""
    from lcqpow_b200 import problems as P
    import dataclasses
    base = P.warm_up_w_A().normalised()
    n = base.nV
    lbA = np.array([[-0.5], [-0.5], [-np.inf]])
    ubA = np.array([[-0.5], [np.inf], [0.25]])
    def batch(order):
        return dataclasses.replace(base, batch=len(order), g=np.tile(base.g, (len(order), 1)), lbA=lbA[order], ubA=ubA[order],
                                   shared=frozenset(("Q", "L", "R", "A")))
    alone = [emu.solve_batch(batch([k]), emu.default_options()) for k in range(3)]
    for order in ([0, 1, 2], [2, 1, 0], [1, 0, 2]):
        s = emu.solve_batch(batch(order), emu.default_options())
        for pos, k in enumerate(order):
            assert int(s.res["ret"][pos]) == int(alone[k].res["ret"][0]) == 0, (order, pos)
            assert int(s.res["iterOuter"][pos]) == int(alone[k].res["iterOuter"][0])
            assert np.abs(s.x[pos] - alone[k].x[0]).max() <= 1e-9, (order, pos)

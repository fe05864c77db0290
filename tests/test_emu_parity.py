"""The product's DEVICE code, compiled as single-threaded C++ (tests/emu), against the golden outputs of the
real reference and against the oracle -- CPU coverage of the kernel logic (the `-m gpu` tests repeat these
through the C ABI on the B200)."""
import numpy as np
import pytest

from conftest import check_against_golden, golden_cases


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


@pytest.mark.parametrize("perturb", [0, 1])
def test_emu_matches_oracle_on_seeded_batches(perturb, emu, oracle):
    """Same trajectory (ReturnValue, stationarity type, outer and total iterations) and x to 1e-8 on seeded
    instances of C2 (shared matrices, CSR operators, static equality block) and C5 (per-instance dense)."""
    from lcqpow_b200 import problems as P
    for pb, over in ((P.circle_batch(6), {"stationarityTolerance": 10e-3}), (P.dense_random_batch(24), {})):
        se = emu.solve_batch(pb, emu.default_options(perturbStep=perturb, **over))
        so = oracle.solve_batch(pb, oracle.default_options(perturbStep=perturb, **over))
        for f in ("ret", "status", "iterOuter", "iterTotal"):
            assert np.array_equal(se.res[f], so.res[f]), (pb.name, f, se.res[f], so.res[f])
        assert np.abs(se.x - so.x).max() <= 1e-8 * max(1.0, np.abs(so.x).max()), pb.name

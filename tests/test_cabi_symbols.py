"""The C-ABI library loads and exports every symbol include/lcqp_cuda.h declares (no compute without a GPU),
and the host-side mirror validates options like the reference."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "lcqp_cuda.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lcqp_cuda_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_are_exported():
    import lcqpow_b200 as L
    from lcqpow_b200 import build
    build.build()
    lib = L.load_library()
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), n
    assert lib.lcqp_cuda_abi_version() == 2
    assert set(L.api.EXPORTS) == set(names)


def test_struct_layouts_match_header():
    import lcqpow_b200 as L
    assert C.sizeof(L.api.CudaOptions) == 6 * 8 + 6 * 4 + 6 * 8 + 4 * 4 + 8 + 9 * 8 + 8 * 4 + 2 * 8   # ABI 2: + the OSQPSettings block + two qpOASES tolerances
    assert L.api.STATS_DTYPE.itemsize == 8 * 4 + 2 * 8


def test_no_device_is_loud():
    """Without a usable sm_100 device create() must fail (no CPU fallback)."""
    import torch
    import lcqpow_b200 as L
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(L.LCQPError) as e:
        L.LCQProblemBatch(2, 0, 1, 1)
    assert e.value.code == 500


def test_options_defaults_and_validation():
    """Options.cpp:296-333 defaults and the setter validations of Options.cpp:85-259."""
    import lcqpow_b200 as L
    o = L.Options()
    assert o.getComplementarityTolerance() == pytest.approx(1e3 * 2.221e-16)
    assert o.getStationarityTolerance() == pytest.approx(1e6 * 2.221e-16)
    assert o.getInitialPenaltyParameter() == 0.01 and o.getPenaltyUpdateFactor() == 2.0
    assert o.getSolveZeroPenaltyFirst() and o.getPerturbStep()
    assert o.getMaxIterations() == 1000 and o.getMaxPenaltyParameter() == 1e8
    assert o.getNDynamicPenalty() == 3 and o.getEtaDynamicPenalty() == 0.9 and o.getQPSolver() == 0
    assert o.setStationarityTolerance(1e-17) == 105 and o.getStationarityTolerance() == pytest.approx(2.221e-10)
    assert o.setComplementarityTolerance(0.0) == 102
    assert o.setInitialPenaltyParameter(0.0) == 103
    assert o.setPenaltyUpdateFactor(1.0) == 101
    assert o.setMaxIterations(0) == 104
    assert o.setMaxPenaltyParameter(0.0) == 121
    assert o.setEtaDynamicPenalty(1.0) == 119
    assert o.setQPSolver(3) == 109 and o.setQPSolver(2) == 0
    p = L.Options(o)  # copy (RunUnitTests.cpp:249-262)
    assert p.getQPSolver() == 2
    p.setQPSolver(0)
    assert o.getQPSolver() == 2


def test_oracle_is_not_imported_by_the_product():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "lcqpow_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "pyref" not in text and "lcqp_oracle" not in text and "oracle/" not in text.replace("# oracle/", ""), f

"""Known-answer tests of /root/reference/test/RunUnitTests.cpp:33-246 applied to the oracle's restated
Utilities kernels (oracle/lcqp_oracle.c).  Integer-exact, as in the reference (ASSERT_EQ)."""
import ctypes as C

import numpy as np
import pytest


@pytest.fixture(scope="module")
def lib(oracle):
    return oracle.lib


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _d(*v):
    return np.array(v, dtype=np.float64)


def test_matrix_multiplication(lib):  # RunUnitTests.cpp:33-57
    A, B, Cm = _d(1, 0, 2, 3, 1, 1), _d(2, 0, 0, 2, 1, 0, 0, 1, 0, -1, -1, 0), np.zeros(8)
    lib.lcqp_oracle_MatrixMultiplication(_p(A), _p(B), _p(Cm), 2, 3, 4)
    assert Cm.tolist() == [2, -2, -2, 2, 7, -1, -1, 7]


def test_transposed_matrix_multiplication(lib):  # :60-78
    A, B, Cm = _d(1, 0, 2, 3, 1, 1), _d(98, -10), np.zeros(3)
    lib.lcqp_oracle_TransponsedMatrixMultiplication(_p(A), _p(B), _p(Cm), 2, 3, 1)
    assert Cm.tolist() == [68, -10, 186]


def test_matrix_symmetrization(lib):  # :81-104 (the asserts, not the header comment, are the golden values)
    A, B, Cm = _d(1, 0, 2, 3, 1, 1), _d(2, 0, 1, 0, 0, -1), np.zeros(9)
    lib.lcqp_oracle_MatrixSymmetrizationProduct(_p(A), _p(B), _p(Cm), 2, 3)
    assert Cm.tolist() == [4, 0, 2, 0, 0, -1, 2, -1, 2]


def test_affine_transformation(lib):  # :107-129
    A, b, c, d = _d(1, 0, 2, 3, 1, 1), _d(2, 0, 1), _d(-3, -3), np.zeros(2)
    lib.lcqp_oracle_AffineLinearTransformation.argtypes = [C.c_double] + [C.POINTER(C.c_double)] * 4 + [C.c_int, C.c_int]
    lib.lcqp_oracle_AffineLinearTransformation(2.0, _p(A), _p(b), _p(c), _p(d), 2, 3)
    assert d.tolist() == [5, 11]


def test_matrix_add(lib):  # :132-159
    A, B, Cm = _d(0, 1, 3, 1, 10, 1), _d(2, 0, 0, 4, 2, 2), np.zeros(6)
    lib.lcqp_oracle_WeightedMatrixAdd.argtypes = [C.c_double, C.POINTER(C.c_double), C.c_double, C.POINTER(C.c_double),
                                                  C.POINTER(C.c_double), C.c_int, C.c_int]
    lib.lcqp_oracle_WeightedMatrixAdd(-1.0, _p(A), 0.5, _p(B), _p(Cm), 3, 2)
    assert Cm.tolist() == [1, -1, -3, 1, -9, 0]


def test_vector_add(lib):  # :162-187
    a, b, d = _d(0, 1, 2, 3), _d(10, 2, 0, 3), np.zeros(4)
    lib.lcqp_oracle_WeightedVectorAdd.argtypes = [C.c_double, C.POINTER(C.c_double), C.c_double, C.POINTER(C.c_double),
                                                  C.POINTER(C.c_double), C.c_int]
    lib.lcqp_oracle_WeightedVectorAdd(2.0, _p(a), -1.0, _p(b), _p(d), 4)
    assert d.tolist() == [-10, 0, 4, 3]


def test_quadratic_form_and_dot_and_maxabs(lib):  # :190-246
    for f in ("QuadraticFormProduct", "DotProduct", "MaxAbs"):
        getattr(lib, "lcqp_oracle_" + f).restype = C.c_double
    p, Q = _d(1, 2, 3), _d(0, 1, 0, 1, 2, 1, 0, 1, 0)
    assert lib.lcqp_oracle_QuadraticFormProduct(_p(Q), _p(p), 3) == 24
    a, b = _d(0, 1, 2, 3), _d(10, 2, 0, 3)
    assert lib.lcqp_oracle_DotProduct(_p(a), _p(b), 4) == 11
    assert lib.lcqp_oracle_MaxAbs(_p(_d(0, 1, 2, 3)), 4) == 3
    assert lib.lcqp_oracle_MaxAbs(_p(_d(0, -1, 2, 0)), 4) == 2
    assert lib.lcqp_oracle_MaxAbs(_p(_d(0, -4, 2, 0)), 4) == 4


def test_perturb_draw_is_ternary_and_balanced(lib):
    lib.lcqp_oracle_perturb_draw.argtypes = [C.c_ulonglong, C.c_ulonglong, C.c_uint, C.c_uint]
    draws = np.array([lib.lcqp_oracle_perturb_draw(1, b, it, i) for b in range(8) for it in range(1, 20) for i in range(40)])
    assert set(draws.tolist()) == {-1, 0, 1}
    counts = np.bincount(draws + 1)
    assert counts.min() > 0.28 * len(draws)

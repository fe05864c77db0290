"""ctypes doors onto the two CPU checkers -- TEST INFRASTRUCTURE ONLY.

* ``RefLib``    : oracle/_ref/liblcqpow_ref.so -- the unmodified reference (LCQPow + qpOASES + OSQP),
                  built by oracle/Makefile from /root/reference.  ``kind = "reference"``.
* ``OracleLib`` : oracle/_build/liblcqp_oracle.so -- our plain-C restatement (oracle/lcqp_oracle.c).
                  ``kind = "port"``.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product package (lcqpow_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "liblcqpow_ref.so")
ORACLE_SO = os.path.join(HERE, "_build", "liblcqp_oracle.so")

FIELDS = ("Q", "g", "L", "R", "lbL", "ubL", "lbR", "ubR", "A", "lbA", "ubA", "lb", "ub", "x0", "y0")

QPOASES_DENSE, QPOASES_SPARSE, OSQP_SPARSE = 0, 1, 2


class RefOptions(C.Structure):
    _fields_ = [("stationarityTolerance", C.c_double), ("complementarityTolerance", C.c_double),
                ("initialPenaltyParameter", C.c_double), ("penaltyUpdateFactor", C.c_double),
                ("maxPenaltyParameter", C.c_double), ("etaDynamicPenalty", C.c_double),
                ("solveZeroPenaltyFirst", C.c_int), ("perturbStep", C.c_int), ("maxIterations", C.c_int),
                ("nDynamicPenalty", C.c_int), ("qpSolver", C.c_int), ("osqp_adaptive_rho_interval", C.c_int)]


class OracleOptions(C.Structure):
    """lcqp_oracle_options (oracle/lcqp_oracle.h): RefOptions fields + the inner exact-QP solver knobs."""
    _fields_ = RefOptions._fields_[:11] + [("reserved0", C.c_int),
                ("qp_rho", C.c_double), ("qp_sigma", C.c_double), ("qp_alpha", C.c_double), ("qp_delta", C.c_double),
                ("qp_feas_tol", C.c_double), ("qp_dual_tol", C.c_double),
                ("qp_max_iter", C.c_int), ("qp_check_interval", C.c_int), ("qp_refine_iter", C.c_int),
                ("qp_adaptive_rho", C.c_int), ("perturb_seed", C.c_ulonglong)]


class RefResult(C.Structure):
    _fields_ = [("ret", C.c_int), ("status", C.c_int), ("iterTotal", C.c_int), ("iterOuter", C.c_int),
                ("subproblemIter", C.c_int), ("qpExitFlag", C.c_int), ("nDuals", C.c_int), ("pad", C.c_int),
                ("rhoOpt", C.c_double), ("seconds", C.c_double)]


RESULT_DTYPE = np.dtype([("ret", "i4"), ("status", "i4"), ("iterTotal", "i4"), ("iterOuter", "i4"),
                         ("subproblemIter", "i4"), ("qpExitFlag", "i4"), ("nDuals", "i4"), ("pad", "i4"),
                         ("rhoOpt", "f8"), ("seconds", "f8")])


@dataclass
class BatchSolution:
    x: np.ndarray      # (batch, nV)
    y: np.ndarray      # (batch, nV+nC+2nComp)  (first nDuals entries valid)
    res: np.ndarray    # structured, RESULT_DTYPE


def build(target: str) -> None:
    subprocess.run(["make", "-s", "-C", HERE, target], check=True)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


class _Lib:
    so_path = ""
    prefix = ""
    options_cls = RefOptions

    def __init__(self):
        if not os.path.exists(self.so_path):
            raise FileNotFoundError(self.so_path)
        self.lib = C.CDLL(self.so_path)
        dp = C.POINTER(C.c_double)
        f = getattr(self.lib, self.prefix + "_default_options")
        f.argtypes = [C.POINTER(self.options_cls)]
        f.restype = None
        f = getattr(self.lib, self.prefix + "_solve_batch")
        f.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint] + [dp] * 15 + [C.POINTER(self.options_cls), dp, dp, C.c_void_p]
        f.restype = C.c_int

    def default_options(self, **over):
        o = self.options_cls()
        getattr(self.lib, self.prefix + "_default_options")(C.byref(o))
        for k, v in over.items():
            setattr(o, k, v)
        return o

    def solve_batch(self, pb, opts=None) -> BatchSolution:
        """``pb`` is an lcqpow_b200.problems.LCQPBatch (duck-typed: needs .normalised())."""
        pb = pb.normalised()
        o = opts if opts is not None else self.default_options()
        nD = pb.nV + pb.nC + 2 * pb.nComp
        x = np.zeros((pb.batch, pb.nV))
        y = np.zeros((pb.batch, nD))
        res = np.zeros(pb.batch, dtype=RESULT_DTYPE)
        ptrs = [_ptr(getattr(pb, f)) for f in FIELDS]
        getattr(self.lib, self.prefix + "_solve_batch")(
            pb.batch, pb.nV, pb.nC, pb.nComp, pb.shared_mask(), *ptrs, C.byref(o), _ptr(x), _ptr(y),
            res.ctypes.data_as(C.c_void_p))
        return BatchSolution(x=x, y=y, res=res)


class RefLib(_Lib):
    so_path = REF_SO
    prefix = "lcqpow_ref"
    kind = "reference"


class OracleLib(_Lib):
    so_path = ORACLE_SO
    prefix = "lcqp_oracle"
    kind = "port"
    options_cls = OracleOptions


def have_ref() -> bool:
    return os.path.exists(REF_SO)

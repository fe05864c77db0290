"""Python binding of the C ABI (include/lcqp_cuda.h) and a host-side mirror of the reference's API.

Names, argument meaning and error behaviour follow LCQPow (/root/reference/include/LCQProblem.hpp:47-242,
Options.hpp:192-213, OutputStatistics.hpp:110-130): ``LCQProblem(nV, nC, nComp)``, ``loadLCQP``,
``setOptions``, ``runSolver``, ``getPrimalSolution``, ``getDualSolution``, ``getOutputStatistics`` -- plus
the batched front door ``LCQProblemBatch``.  All compute happens in liblcqp_cuda.so on a B200; there is no
CPU path: importing works without a GPU (so that symbols can be checked), computing does not.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

from . import build as _build

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = _build.LIB

FIELDS = ("Q", "g", "L", "R", "lbL", "ubL", "lbR", "ubR", "A", "lbA", "ubA", "lb", "ub", "x0", "y0")

# LCQPow::ReturnValue (Utilities.hpp:37-87) -- the values used here
SUCCESSFUL_RETURN = 0
MAX_ITERATIONS_REACHED = 200
MAX_PENALTY_REACHED = 201
SUBPROBLEM_SOLVER_ERROR = 203
# LCQPow::AlgorithmStatus (Utilities.hpp:103-109)
PROBLEM_NOT_SOLVED, W_STATIONARY_SOLUTION, C_STATIONARY_SOLUTION, M_STATIONARY_SOLUTION, S_STATIONARY_SOLUTION = range(5)
# LCQPow::QPSolver (Utilities.hpp:125-129) + the CUDA plugin
QPOASES_DENSE, QPOASES_SPARSE, OSQP_SPARSE = 0, 1, 2

EXPORTS = (
    "lcqp_cuda_abi_version", "lcqp_cuda_default_options", "lcqp_cuda_create", "lcqp_cuda_destroy",
    "lcqp_cuda_set_options", "lcqp_cuda_load", "lcqp_cuda_load_device", "lcqp_cuda_load_csc", "lcqp_cuda_set_instance_offset",
    "lcqp_cuda_run", "lcqp_cuda_synchronize", "lcqp_cuda_get_primal", "lcqp_cuda_get_dual",
    "lcqp_cuda_get_stats", "lcqp_cuda_get_device_results", "lcqp_cuda_num_duals", "lcqp_cuda_launch_count",
    "lcqp_cuda_last_run_ms", "lcqp_cuda_last_launch_info", "lcqp_cuda_last_error", "lcqp_cuda_osqp_info", "lcqp_cuda_qp_create", "lcqp_cuda_qp_destroy",
    "lcqp_cuda_qp_set_options", "lcqp_cuda_qp_solve", "lcqp_cuda_qp_get_solution", "lcqp_cuda_measure_fp64_tflops",
    "lcqp_cuda_measure_l2_gbs", "lcqp_cuda_last_work",
)


class CudaOptions(C.Structure):
    """lcqp_cuda_options."""
    _fields_ = [("stationarityTolerance", C.c_double), ("complementarityTolerance", C.c_double),
                ("initialPenaltyParameter", C.c_double), ("penaltyUpdateFactor", C.c_double),
                ("maxPenaltyParameter", C.c_double), ("etaDynamicPenalty", C.c_double),
                ("solveZeroPenaltyFirst", C.c_int), ("perturbStep", C.c_int), ("maxIterations", C.c_int),
                ("nDynamicPenalty", C.c_int), ("qpSolver", C.c_int), ("osqp_admm", C.c_int),
                ("qp_rho", C.c_double), ("qp_sigma", C.c_double), ("qp_alpha", C.c_double), ("qp_delta", C.c_double),
                ("qp_feas_tol", C.c_double), ("qp_dual_tol", C.c_double),
                ("qp_max_iter", C.c_int), ("qp_check_interval", C.c_int), ("qp_refine_iter", C.c_int),
                ("qp_adaptive_rho", C.c_int), ("perturb_seed", C.c_ulonglong),
                # OSQPSettings of the OSQP restatement (ABI 2)
                ("osqp_rho", C.c_double), ("osqp_sigma", C.c_double), ("osqp_alpha", C.c_double), ("osqp_delta", C.c_double),
                ("osqp_eps_abs", C.c_double), ("osqp_eps_rel", C.c_double), ("osqp_eps_prim_inf", C.c_double),
                ("osqp_eps_dual_inf", C.c_double), ("osqp_adaptive_rho_tolerance", C.c_double),
                ("osqp_max_iter", C.c_int), ("osqp_check_termination", C.c_int), ("osqp_scaling", C.c_int),
                ("osqp_adaptive_rho", C.c_int), ("osqp_adaptive_rho_interval", C.c_int), ("osqp_polish", C.c_int),
                ("osqp_polish_refine_iter", C.c_int), ("osqp_reserved", C.c_int),
                ("qpoases_terminationTolerance", C.c_double), ("qpoases_boundTolerance", C.c_double)]


STATS_DTYPE = np.dtype([("ret", "i4"), ("status", "i4"), ("iterTotal", "i4"), ("iterOuter", "i4"),
                        ("subproblemIter", "i4"), ("qpExitFlag", "i4"), ("nDuals", "i4"), ("kktSolves", "i4"),
                        ("rhoOpt", "f8"), ("admmIters", "f8")])

_lib = None


def load_library(build_if_missing: bool = False) -> C.CDLL:
    """dlopen liblcqp_cuda.so (in-tree).  Raises if it is missing: there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if build_if_missing:
            _build.build()
        else:
            raise RuntimeError(f"{LIB_PATH} is missing: run `python -m lcqpow_b200.build` (the CUDA library is the "
                               "only implementation; there is no CPU fallback)")
    # LCQP_CUDA_LIB: development aid (e.g. the -DLCQP_PROFILE build of the same sources)
    lib = C.CDLL(os.environ.get("LCQP_CUDA_LIB", LIB_PATH))
    dp = C.POINTER(C.c_double)
    vp = C.c_void_p
    lib.lcqp_cuda_abi_version.restype = C.c_int
    lib.lcqp_cuda_default_options.argtypes = [C.POINTER(CudaOptions)]
    lib.lcqp_cuda_default_options.restype = None
    lib.lcqp_cuda_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
    lib.lcqp_cuda_destroy.argtypes = [vp]
    lib.lcqp_cuda_set_options.argtypes = [vp, C.POINTER(CudaOptions)]
    lib.lcqp_cuda_load.argtypes = [vp, C.c_int, C.c_uint] + [vp] * 15
    lib.lcqp_cuda_load_device.argtypes = [vp, C.c_int, C.c_uint] + [vp] * 15
    lib.lcqp_cuda_load_csc.argtypes = [vp, C.c_int, C.c_uint] + [vp] * 21
    lib.lcqp_cuda_set_instance_offset.argtypes = [vp, C.c_ulonglong]
    lib.lcqp_cuda_run.argtypes = [vp, vp]
    lib.lcqp_cuda_synchronize.argtypes = [vp]
    lib.lcqp_cuda_get_primal.argtypes = [vp, vp]
    lib.lcqp_cuda_get_dual.argtypes = [vp, vp]
    lib.lcqp_cuda_get_stats.argtypes = [vp, vp]
    lib.lcqp_cuda_get_device_results.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
    lib.lcqp_cuda_num_duals.argtypes = [vp]
    lib.lcqp_cuda_launch_count.argtypes = [vp]
    lib.lcqp_cuda_launch_count.restype = C.c_longlong
    lib.lcqp_cuda_last_run_ms.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    lib.lcqp_cuda_last_launch_info.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.lcqp_cuda_last_error.argtypes = [vp]
    lib.lcqp_cuda_osqp_info.argtypes = [vp] + [C.POINTER(C.c_int)] * 4
    lib.lcqp_cuda_last_error.restype = C.c_char_p
    lib.lcqp_cuda_qp_create.argtypes = [C.c_int, C.c_int, dp, dp, C.c_int, C.POINTER(vp)]
    lib.lcqp_cuda_qp_destroy.argtypes = [vp]
    lib.lcqp_cuda_qp_set_options.argtypes = [vp, C.POINTER(CudaOptions)]
    lib.lcqp_cuda_qp_solve.argtypes = [vp, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)] + [dp] * 7
    lib.lcqp_cuda_qp_get_solution.argtypes = [vp, dp, dp]
    lib.lcqp_cuda_measure_fp64_tflops.argtypes = [C.c_int, C.POINTER(C.c_double)]
    lib.lcqp_cuda_measure_l2_gbs.argtypes = [C.c_int, C.POINTER(C.c_double)]
    lib.lcqp_cuda_last_work.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    for name in EXPORTS:
        f = getattr(lib, name)
        if name not in ("lcqp_cuda_default_options", "lcqp_cuda_launch_count", "lcqp_cuda_last_error"):
            f.restype = C.c_int
    _lib = lib
    return lib


class LCQPError(RuntimeError):
    def __init__(self, code: int, what: str):
        super().__init__(f"{what} failed with code {code}")
        self.code = code


class Options:
    """Mirror of LCQPow::Options (defaults /root/reference/src/Options.cpp:296-333).  Setters return the
    reference's ReturnValue (0 = ok) and leave the value unchanged when it is invalid, as the reference does."""

    EPS = 2.221e-16

    def __init__(self, other: "Options | None" = None):
        self.c = CudaOptions()
        load_library().lcqp_cuda_default_options(C.byref(self.c))
        if other is not None:
            C.memmove(C.byref(self.c), C.byref(other.c), C.sizeof(CudaOptions))

    # getters / setters with the validations of Options.cpp:85-259
    def getStationarityTolerance(self): return self.c.stationarityTolerance
    def setStationarityTolerance(self, v):
        if v <= self.EPS: return 105
        self.c.stationarityTolerance = v; return 0
    def getComplementarityTolerance(self): return self.c.complementarityTolerance
    def setComplementarityTolerance(self, v):
        if v <= self.EPS: return 102
        self.c.complementarityTolerance = v; return 0
    def getInitialPenaltyParameter(self): return self.c.initialPenaltyParameter
    def setInitialPenaltyParameter(self, v):
        if v <= 1e-25: return 103
        self.c.initialPenaltyParameter = v; return 0
    def getPenaltyUpdateFactor(self): return self.c.penaltyUpdateFactor
    def setPenaltyUpdateFactor(self, v):
        if v <= 1: return 101
        self.c.penaltyUpdateFactor = v; return 0
    def getSolveZeroPenaltyFirst(self): return bool(self.c.solveZeroPenaltyFirst)
    def setSolveZeroPenaltyFirst(self, v): self.c.solveZeroPenaltyFirst = int(bool(v)); return 0
    def getPerturbStep(self): return bool(self.c.perturbStep)
    def setPerturbStep(self, v): self.c.perturbStep = int(bool(v)); return 0
    def getMaxIterations(self): return self.c.maxIterations
    def setMaxIterations(self, v):
        if v <= 0: return 104
        self.c.maxIterations = int(v); return 0
    def getMaxPenaltyParameter(self): return self.c.maxPenaltyParameter
    def setMaxPenaltyParameter(self, v):
        if v <= 1e-25: return 121
        self.c.maxPenaltyParameter = v; return 0
    def getNDynamicPenalty(self): return self.c.nDynamicPenalty
    def setNDynamicPenalty(self, v): self.c.nDynamicPenalty = int(v); return 0
    def getEtaDynamicPenalty(self): return self.c.etaDynamicPenalty
    def setEtaDynamicPenalty(self, v):
        if v <= 0 or v >= 1: return 119
        self.c.etaDynamicPenalty = v; return 0
    def getQPSolver(self): return self.c.qpSolver
    def setQPSolver(self, v):
        if v < QPOASES_DENSE or v > OSQP_SPARSE: return 109
        self.c.qpSolver = int(v); return 0
    def setPerturbSeed(self, v): self.c.perturb_seed = int(v); return 0

    # OSQP flavour: QPSolver::OSQP_SPARSE as the reference runs it (ADMM + polish) instead of the exact-vertex solver
    # behind the OSQP dual layout; `settings` are OSQPSettings fields (rho, sigma, alpha, delta, eps_abs, eps_rel,
    # eps_prim_inf, eps_dual_inf, adaptive_rho, adaptive_rho_interval, adaptive_rho_tolerance, max_iter,
    # check_termination, scaling, polish, polish_refine_iter) -- Options::setOSQPOptions, Options.cpp:275-287
    def setOSQPADMM(self, on: bool = True, **settings):
        self.c.osqp_admm = int(bool(on))
        if on:
            self.c.qpSolver = OSQP_SPARSE
        return self.setOSQPOptions(**settings)

    def setqpOASESOptions(self, **settings):
        """qpOASES::Options fields the device solver honours: terminationTolerance, boundTolerance
        (Options::setqpOASESOptions, Options.cpp:262-266)."""
        for k, v in settings.items():
            if not hasattr(self.c, "qpoases_" + k):
                return 100
            setattr(self.c, "qpoases_" + k, v)
        return 0

    def setOSQPOptions(self, **settings):
        for k, v in settings.items():
            if not hasattr(self.c, "osqp_" + k):
                return 100
            setattr(self.c, "osqp_" + k, v)
        return 0


class OutputStatistics:
    """Mirror of LCQPow::OutputStatistics' counters (OutputStatistics.hpp:110-130) for one instance."""

    def __init__(self, rec=None):
        self.rec = rec

    def getIterTotal(self): return int(self.rec["iterTotal"])
    def getIterOuter(self): return int(self.rec["iterOuter"])
    def getSubproblemIter(self): return int(self.rec["subproblemIter"])
    def getRhoOpt(self): return float(self.rec["rhoOpt"])
    def getSolutionStatus(self): return int(self.rec["status"])
    def getQPSolverExitFlag(self): return int(self.rec["qpExitFlag"])


def _as_ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class LCQProblemBatch:
    """``batch`` independent LCQPs of one shape solved on one GPU (the batched front door of the C ABI)."""

    def __init__(self, nV: int, nC: int, nComp: int, batch: int, device: int = 0):
        self.lib = load_library()
        self.nV, self.nC, self.nComp, self.capacity, self.device = nV, nC, nComp, batch, device
        self.h = C.c_void_p()
        rc = self.lib.lcqp_cuda_create(nV, nC, nComp, batch, device, C.byref(self.h))
        if rc != 0:
            self.h = None
            raise LCQPError(rc, "lcqp_cuda_create")
        self.batch = 0
        self.options = Options()
        self._keep = None

    def close(self):
        if getattr(self, "h", None):
            self.lib.lcqp_cuda_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _err(self) -> str:
        return self.lib.lcqp_cuda_last_error(self.h).decode()

    def setOptions(self, options: Options) -> int:
        rc = self.lib.lcqp_cuda_set_options(self.h, C.byref(options.c))
        if rc == 0:
            self.options = Options(options)
        return rc

    def loadLCQP(self, Q, g, L, R, lbL=None, ubL=None, lbR=None, ubR=None, A=None, lbA=None, ubA=None,
                 lb=None, ub=None, x0=None, y0=None, batch: Optional[int] = None, shared: Sequence[str] = ()) -> int:
        """LCQProblem::loadLCQP (LCQProblem.cpp:87-144) with a leading batch dimension on every array that is
        not named in ``shared``.  Host numpy arrays (row-major fp64)."""
        vals = dict(Q=Q, g=g, L=L, R=R, lbL=lbL, ubL=ubL, lbR=lbR, ubR=ubR, A=A, lbA=lbA, ubA=ubA, lb=lb, ub=ub, x0=x0, y0=y0)
        batch = self.capacity if batch is None else batch
        if batch <= 0 or batch > self.capacity:
            raise ValueError(f"loadLCQP: batch {batch} outside 1..{self.capacity}")
        mask = 0
        arrs = []
        for k, f in enumerate(FIELDS):
            a = vals[f]
            if a is None:
                arrs.append(None)
                continue
            a = np.ascontiguousarray(a, dtype=np.float64)
            if f in shared:
                mask |= 1 << k
            # the library copies field_len * (1 | batch) doubles from this buffer: a short array would be read
            # past its end
            want = self._field_len(f) * (1 if (f in shared or batch == 1) else batch)
            if a.size != want:
                raise ValueError(f"loadLCQP: {f} has {a.size} entries, expected {want} "
                                 f"({'shared' if f in shared else 'batch of ' + str(batch)})")
            arrs.append(a)
        self._keep = arrs
        rc = self.lib.lcqp_cuda_load(self.h, batch, mask, *[_as_ptr(a) for a in arrs])
        if rc == 0:
            self.batch = batch
        return rc

    def _field_len(self, f: str) -> int:
        n, c, p = self.nV, self.nC, self.nComp
        return {"Q": n * n, "g": n, "L": p * n, "R": p * n, "lbL": p, "ubL": p, "lbR": p, "ubR": p,
                "A": c * n, "lbA": c, "ubA": c, "lb": n, "ub": n, "x0": n, "y0": n + c + 2 * p}[f]

    def loadCSC(self, Q, g, L, R, lbL=None, ubL=None, lbR=None, ubR=None, A=None, lbA=None, ubA=None, x0=None, y0=None,
                batch: Optional[int] = None, shared: Sequence[str] = ()) -> int:
        """LCQProblem::loadLCQP(const csc* ...) (LCQProblem.cpp:312-387) for a batch that shares its sparsity patterns.
        Q, L, R, A are (colptr int32[nV+1], rowidx int32[nnz], values float64[nnz] or [batch, nnz]) triples; the vectors
        are as in loadLCQP.  Serves the OSQP flavour (Options.setOSQPADMM)."""
        batch = self.capacity if batch is None else batch
        mask = sum(1 << k for k, f in enumerate(FIELDS) if f in shared)
        keep = []

        def mat(name, t):
            if t is None:
                return [None, None, None]
            p = np.ascontiguousarray(t[0], dtype=np.int32)
            i = np.ascontiguousarray(t[1], dtype=np.int32)
            x = np.ascontiguousarray(t[2], dtype=np.float64)
            if p.size != self.nV + 1 or i.size != int(p[-1]):
                raise ValueError(f"loadCSC: {name} has inconsistent index arrays")
            want = int(p[-1]) * (1 if (name in shared or batch == 1) else batch)
            if x.size != want:
                raise ValueError(f"loadCSC: {name} has {x.size} values, expected {want}")
            keep.extend((p, i, x))
            return [_as_ptr(p), _as_ptr(i), _as_ptr(x)]

        def vec(name, a):
            if a is None:
                return None
            a = np.ascontiguousarray(a, dtype=np.float64)
            want = self._field_len(name) * (1 if (name in shared or batch == 1) else batch)
            if a.size != want:
                raise ValueError(f"loadCSC: {name} has {a.size} entries, expected {want}")
            keep.append(a)
            return _as_ptr(a)

        args = mat("Q", Q) + [vec("g", g)] + mat("L", L) + mat("R", R) + [vec("lbL", lbL), vec("ubL", ubL), vec("lbR", lbR), vec("ubR", ubR)] \
            + mat("A", A) + [vec("lbA", lbA), vec("ubA", ubA), vec("x0", x0), vec("y0", y0)]
        self._keep = keep
        rc = self.lib.lcqp_cuda_load_csc(self.h, batch, mask, *args)
        if rc == 0:
            self.batch = batch
        return rc

    def loadBatch(self, pb) -> int:
        """Load an lcqpow_b200.problems.LCQPBatch."""
        pb = pb.normalised()
        return self.loadLCQP(**{f: getattr(pb, f) for f in FIELDS}, batch=pb.batch, shared=tuple(pb.shared))

    def loadDevicePointers(self, ptrs: dict, batch: int, shared: Sequence[str] = ()) -> int:
        """Inputs already resident in HBM: ``ptrs`` maps field name -> device address (int) or None."""
        mask = sum(1 << k for k, f in enumerate(FIELDS) if f in shared)
        args = [C.c_void_p(ptrs.get(f)) if ptrs.get(f) else None for f in FIELDS]
        rc = self.lib.lcqp_cuda_load_device(self.h, batch, mask, *args)
        if rc == 0:
            self.batch = batch
        return rc

    def setInstanceOffset(self, off: int) -> int:
        return self.lib.lcqp_cuda_set_instance_offset(self.h, off)

    def runSolver(self, stream: int = 0, sync: bool = True) -> int:
        """LCQProblem::runSolver for every instance.  Returns 0 or a launch error; per-instance return values
        are in ``getOutputStatistics()['ret']``."""
        rc = self.lib.lcqp_cuda_run(self.h, C.c_void_p(stream))
        if rc != 0:
            raise LCQPError(rc, "lcqp_cuda_run: " + self._err())
        if sync:
            rc = self.lib.lcqp_cuda_synchronize(self.h)
            if rc != 0:
                raise LCQPError(rc, "lcqp_cuda_synchronize: " + self._err())
        return rc

    def getPrimalSolution(self) -> np.ndarray:
        x = np.empty((self.batch, self.nV))
        rc = self.lib.lcqp_cuda_get_primal(self.h, _as_ptr(x))
        if rc != 0:
            raise LCQPError(rc, "lcqp_cuda_get_primal: " + self._err())
        return x

    def getDualSolution(self) -> np.ndarray:
        y = np.empty((self.batch, self.nV + self.nC + 2 * self.nComp))
        rc = self.lib.lcqp_cuda_get_dual(self.h, _as_ptr(y))
        if rc != 0:
            raise LCQPError(rc, "lcqp_cuda_get_dual: " + self._err())
        return y

    def getOutputStatistics(self) -> np.ndarray:
        st = np.empty(self.batch, dtype=STATS_DTYPE)
        rc = self.lib.lcqp_cuda_get_stats(self.h, _as_ptr(st))
        if rc != 0:
            raise LCQPError(rc, "lcqp_cuda_get_stats: " + self._err())
        return st

    def getNumberOfPrimals(self) -> int:
        return self.nV

    def getNumberOfDuals(self) -> int:
        return self.lib.lcqp_cuda_num_duals(self.h)

    def launchCount(self) -> int:
        return int(self.lib.lcqp_cuda_launch_count(self.h))

    def lastLaunchInfo(self):
        """(CTAs, dynamic shared memory per CTA in bytes, order of the static equality block) of the last run."""
        g, sm, me = C.c_int(), C.c_int(), C.c_int()
        rc = self.lib.lcqp_cuda_last_launch_info(self.h, C.byref(g), C.byref(sm), C.byref(me))
        if rc != 0:
            raise LCQPError(rc, "lcqp_cuda_last_launch_info")
        return g.value, sm.value, me.value

    def osqpInfo(self):
        """(N, nnz(L), levels of a forward + backward solve, 1 if the last run used one warp per instance) of an OSQP-flavour load."""
        v = [C.c_int() for _ in range(4)]
        rc = self.lib.lcqp_cuda_osqp_info(self.h, *[C.byref(t) for t in v])
        if rc != 0:
            raise LCQPError(rc, "lcqp_cuda_osqp_info")
        return tuple(t.value for t in v)

    def lastRunMs(self):
        a, b = C.c_float(), C.c_float()
        rc = self.lib.lcqp_cuda_last_run_ms(self.h, C.byref(a), C.byref(b))
        if rc != 0:
            raise LCQPError(rc, "lcqp_cuda_last_run_ms")
        return a.value, b.value


class LCQProblem:
    """Single-instance mirror of LCQPow::LCQProblem (LCQProblem.hpp:47-242) on top of the batched door."""

    def __init__(self, nV: int, nC: int, nComp: int, device: int = 0):
        self._b = LCQProblemBatch(nV, nC, nComp, 1, device)
        self._stats = None
        self._x = None
        self._y = None

    def setOptions(self, options: Options) -> int:
        return self._b.setOptions(options)

    def loadLCQP(self, Q, g, L, R, lbL=None, ubL=None, lbR=None, ubR=None, A=None, lbA=None, ubA=None,
                 lb=None, ub=None, x0=None, y0=None) -> int:
        return self._b.loadLCQP(Q, g, L, R, lbL, ubL, lbR, ubR, A, lbA, ubA, lb, ub, x0, y0, batch=1,
                                shared=FIELDS)

    def runSolver(self) -> int:
        self._b.runSolver()
        self._stats = self._b.getOutputStatistics()[0]
        self._x = self._b.getPrimalSolution()[0]
        self._y = self._b.getDualSolution()[0][: self._b.getNumberOfDuals()]
        return int(self._stats["ret"])

    def getPrimalSolution(self):
        return self._x, int(self._stats["status"])

    def getDualSolution(self):
        return self._y, int(self._stats["status"])

    def getNumberOfPrimals(self): return self._b.nV
    def getNumberOfDuals(self): return self._b.getNumberOfDuals()
    def getOutputStatistics(self) -> OutputStatistics: return OutputStatistics(self._stats)


class SubsolverCUDA:
    """Python view of the plugin door (SubsolverBase::solve / getSolution, SubsolverBase.hpp:37-56)."""

    def __init__(self, nV: int, nCtot: int, Q, A, device: int = 0):
        self.lib = load_library()
        self.nV, self.nC = nV, nCtot
        Q = np.ascontiguousarray(Q, dtype=np.float64)
        A = np.ascontiguousarray(A, dtype=np.float64) if nCtot > 0 else np.zeros(1)
        self.h = C.c_void_p()
        dp = C.POINTER(C.c_double)
        rc = self.lib.lcqp_cuda_qp_create(nV, nCtot, Q.ctypes.data_as(dp), A.ctypes.data_as(dp), device, C.byref(self.h))
        if rc != 0:
            self.h = None
            raise LCQPError(rc, "lcqp_cuda_qp_create")

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.lib.lcqp_cuda_qp_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def setOptions(self, options: Options) -> int:
        return self.lib.lcqp_cuda_qp_set_options(self.h, C.byref(options.c))

    def solve(self, initialSolve: bool, g, lbA, ubA, x0=None, y0=None, lb=None, ub=None):
        """Returns (ReturnValue, iterations, exit_flag)."""
        dp = C.POINTER(C.c_double)
        keep = [None if a is None else np.ascontiguousarray(a, dtype=np.float64) for a in (g, lbA, ubA, x0, y0, lb, ub)]
        it, fl = C.c_int(0), C.c_int(0)
        rc = self.lib.lcqp_cuda_qp_solve(self.h, int(bool(initialSolve)), C.byref(it), C.byref(fl),
                                         *[None if a is None else a.ctypes.data_as(dp) for a in keep])
        return rc, it.value, fl.value

    def getSolution(self):
        dp = C.POINTER(C.c_double)
        x = np.empty(self.nV)
        y = np.empty(self.nV + self.nC)
        rc = self.lib.lcqp_cuda_qp_get_solution(self.h, x.ctypes.data_as(dp), y.ctypes.data_as(dp))
        if rc != 0:
            raise LCQPError(rc, "lcqp_cuda_qp_get_solution")
        return x, y

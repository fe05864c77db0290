"""Generate tests/golden/*.npz by running the UNMODIFIED reference (oracle/_ref/liblcqpow_ref.so, built by
oracle/Makefile from /root/reference) on the parity fixtures.  Run in the dev container:

    make -C oracle ref && python tests/golden/make_golden.py

The committed outputs are what the `-m gpu` tests and the oracle tests compare against on machines where
/root/reference does not exist.  perturbStep is OFF for the golden runs: with it on the reference seeds
libc rand() from time(NULL) (LCQProblem.cpp:1016) and is not reproducible.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from lcqpow_b200 import problems as P  # noqa: E402
from oracle import pyref  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DATA = "/root/reference/examples/example_data"
STEMS = ("Q", "g", "L", "R", "lbL", "ubL", "lbR", "ubR", "A", "lbA", "ubA", "lb", "ub", "x0")


def family_cases(data):
    """name -> (LCQPBatch, option overrides, reference subsolver).  Also imported by tests/conftest.py."""
    return {
        "circle_bench": (P.circle_batch_fast(256), {"stationarityTolerance": 10e-3}, pyref.QPOASES_SPARSE),
        "dense_bench": (P.dense_random_batch(256), {}, pyref.QPOASES_DENSE),
        "circle_N20": (P.circle_batch(16, N=20), {"stationarityTolerance": 10e-3}, pyref.QPOASES_DENSE),
        "dense_n32": (P.dense_random_batch(32, n=32, nComp=16, nC=8), {}, pyref.QPOASES_DENSE),
        "example_data_family": (P.example_data_batch(data, 17, perturb_ub=True), {}, pyref.QPOASES_DENSE),
        "example_data_bounds": (P.example_data_batch(data, 64), {}, pyref.QPOASES_SPARSE),
    }


def main():
    ref = pyref.RefLib()
    # 1) the shipped input fixture (input data only)
    data = {k: np.loadtxt(os.path.join(REF_DATA, k + ".txt")) for k in STEMS}
    np.savez_compressed(os.path.join(HERE, "example_data.npz"), **data)

    cases = {
        "warm_up": (P.warm_up(), {}),
        "warm_up_noguess": (P.warm_up(False), {}),
        "warm_up_w_A": (P.warm_up_w_A(), {}),
        "warm_up_binary": (P.warm_up_binary(), {}),
        "warm_up_shifted": (P.warm_up_shifted(), {}),
        "infeasible_qp": (P.infeasible_qp(), {}),
        "max_penalty": (P.warm_up(), {"maxPenaltyParameter": 1.0}),          # test/examples/test_max_penalty.cpp:49
        "circle": (P.circle_batch(8), {"stationarityTolerance": 10e-3}),      # examples/OptimizeOnCircle.cpp:45
        "dense": (P.dense_random_batch(16), {}),
        "example_data": (P.example_data_batch(data, 1), {}),
        "stationarity_S": (P.stationarity_fixture("S"), {}),
        "stationarity_M": (P.stationarity_fixture("M"), {}),
        "stationarity_C": (P.stationarity_fixture("C"), {}),
        "stationarity_W": (P.stationarity_fixture("W"), {}),
    }
    out = {}
    for name, (pb, over) in cases.items():
        for flavour, tag in ((pyref.QPOASES_DENSE, "qpoases"), (pyref.OSQP_SPARSE, "osqp")):
            if tag == "osqp" and (pb.lb is not None or pb.ub is not None):
                continue
            o = ref.default_options(qpSolver=flavour, perturbStep=0, **over)
            if tag == "osqp":
                o.osqp_adaptive_rho_interval = 25  # make the OSQP run wall-clock independent
            s = ref.solve_batch(pb, o)
            out[f"{name}/{tag}/x"] = s.x
            out[f"{name}/{tag}/y"] = s.y
            for f in ("ret", "status", "iterTotal", "iterOuter", "subproblemIter", "qpExitFlag", "nDuals"):
                out[f"{name}/{tag}/{f}"] = s.res[f].astype(np.int64)
            out[f"{name}/{tag}/rhoOpt"] = s.res["rhoOpt"]
            print(name, tag, "ret", s.res["ret"].tolist(), "k", s.res["iterOuter"].tolist(), "i", s.res["iterTotal"].tolist())
    np.savez_compressed(os.path.join(HERE, "reference_outputs.npz"), **out)

    # 2) the families the bench and the verdict of round 1 use: 256 instances each, qpOASES runs of the reference
    #    (QPOASES_SPARSE for the circle family -- three times faster than QPOASES_DENSE, same trajectories)
    fam = {}
    for name, (pb, over, flavour) in family_cases(data).items():
        s = ref.solve_batch(pb, ref.default_options(qpSolver=flavour, perturbStep=0, **over))
        fam[f"{name}/x"] = s.x
        fam[f"{name}/y"] = s.y.astype(np.float32)   # duals: signs / classification only
        for f in ("ret", "status", "iterTotal", "iterOuter", "subproblemIter", "qpExitFlag"):
            fam[f"{name}/{f}"] = s.res[f].astype(np.int32)
        fam[f"{name}/rhoOpt"] = s.res["rhoOpt"]
        print(name, "ret", np.unique(s.res["ret"], return_counts=True), "status", np.unique(s.res["status"], return_counts=True),
              "k mean", s.res["iterOuter"].mean(), "i mean", s.res["iterTotal"].mean(), "sub mean", s.res["subproblemIter"].mean())
    np.savez_compressed(os.path.join(HERE, "reference_families.npz"), **fam)


if __name__ == "__main__":
    main()

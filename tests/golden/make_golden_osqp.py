"""Generate tests/golden/reference_families_osqp.npz: the UNMODIFIED reference's OSQP_SPARSE runs (oracle/_ref) of the
families the OSQP flavour is checked against.  Run in the dev container after `make -C oracle ref`:

    python tests/golden/make_golden_osqp.py

perturbStep is off and adaptive_rho_interval is fixed to 25: with its default (0) OSQP derives the interval from
wall-clock timings (external/osqp/src/osqp.c:459-485) and the reference's own iteration counts change from run to run.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from lcqpow_b200 import problems as P  # noqa: E402
from oracle import pyref  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def osqp_family_cases():
    """name -> (batch, option overrides).  Dense LCQPBatch, or SparseLCQPBatch for the sparse configuration C4."""
    return {
        "dense_bench": (P.dense_random_batch(256), {}),
        "dense_n32": (P.dense_random_batch(32, n=32, nComp=16, nC=8), {}),
        "circle_bench": (P.circle_batch_fast(64), {"stationarityTolerance": 10e-3}),
        "sparse_c4": (P.sparse_banded_batch(16), {}),
    }


def main():
    ref = pyref.RefLib()
    fam = {}
    for name, (pb, over) in osqp_family_cases().items():
        o = ref.default_options(qpSolver=pyref.OSQP_SPARSE, perturbStep=0, **over)
        o.osqp_adaptive_rho_interval = 25
        if isinstance(pb, P.SparseLCQPBatch):
            parts = [ref.solve_batch(pb.to_dense(b, b + 1), o) for b in range(pb.batch)]
            x = np.concatenate([s.x for s in parts]); y = np.concatenate([s.y for s in parts]); res = np.concatenate([s.res for s in parts])
        else:
            s = ref.solve_batch(pb, o)
            x, y, res = s.x, s.y, s.res
        fam[f"{name}/x"] = x
        fam[f"{name}/y"] = y.astype(np.float32)
        for f in ("ret", "status", "iterTotal", "iterOuter", "subproblemIter", "qpExitFlag"):
            fam[f"{name}/{f}"] = res[f].astype(np.int32)
        fam[f"{name}/rhoOpt"] = res["rhoOpt"]
        print(name, "ret", np.unique(res["ret"], return_counts=True), "k mean", res["iterOuter"].mean(), "i mean", res["iterTotal"].mean(),
              "ADMM iterations mean", res["subproblemIter"].mean())
    # subsolver option pass-through (Options::setqpOASESOptions / setOSQPOptions): the 16 dense golden
    # instances with qpOASES terminationTolerance = 1e-2 (instance 1 then runs into MAX_ITERATIONS_REACHED) and with
    # OSQP eps_abs = eps_rel = 1e-5
    import ctypes as C
    ref.lib.lcqpow_ref_set_subsolver_options.argtypes = [C.c_double, C.c_double, C.c_double, C.c_int]
    pb = P.dense_random_batch(16)
    for tag, args, qs in (("opt_qpoases_termtol_1e-2", (1e-2, 0.0, 0.0, 0), pyref.QPOASES_DENSE), ("opt_osqp_eps_1e-5", (0.0, 0.0, 1e-5, 0), pyref.OSQP_SPARSE)):
        ref.lib.lcqpow_ref_set_subsolver_options(*args)
        o = ref.default_options(qpSolver=qs, perturbStep=0)
        o.osqp_adaptive_rho_interval = 25
        s = ref.solve_batch(pb, o)
        fam[f"{tag}/x"] = s.x
        for f in ("ret", "status", "iterTotal", "iterOuter", "subproblemIter", "qpExitFlag"):
            fam[f"{tag}/{f}"] = s.res[f].astype(np.int32)
        print(tag, "ret", s.res["ret"].tolist(), "sub", s.res["subproblemIter"].tolist())
    ref.lib.lcqpow_ref_set_subsolver_options(0.0, 0.0, 0.0, 0)
    np.savez_compressed(os.path.join(HERE, "reference_families_osqp.npz"), **fam)


if __name__ == "__main__":
    main()

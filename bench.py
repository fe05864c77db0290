#!/usr/bin/env python
"""bench.py -- throughput of the LCQP hot path (LCQProblem::runSolver + QP subsolver) on B200.

Metric (BASELINE.json): LCQPs solved / second, batched, on N GPUs of one node.
Workloads, chosen by --config:
  c2 (default, the configuration the metric is quoted on): OptimizeOnCircle-shaped LCQPs (nV=202, nC=101,
      nComp=100) sharing Q/A/L/R/lbA/ubA, per-instance g and x0 (examples/OptimizeOnCircle.cpp:62-99 with a random
      x_ref per instance), stationarityTolerance = 1e-2 as the example sets it; 2^20 instances per GPU per step.
  c5: random dense LCQPs (n=64, 32 pairs, 16 constraints), every matrix per instance; 131072 per GPU per step.
  c3: examples/example_data (nV=151, nC=50, nComp=100, box bounds) replicated with perturbed g and lbA=ubA; 10000 per step.
  c4: synthetic sparse LCQPs (n=1000, 500 pairs, 300 constraints, banded Q; one sparsity pattern, per-instance values)
      through the CSC door and the OSQP-ADMM flavour (QPSolver::OSQP_SPARSE semantics); 4096 per GPU per step.  A step
      is minutes long (one warp per instance, thousands of ADMM iterations each): --warmup is honoured as given (>= 1).
--flavour osqp runs c2 / c5 through the OSQP restatement as well (c4 always does).
perturbStep is on (the default of the reference's Options).

One "step" = one pass of the hot path over one batch of `--batch` instances per GPU (weak scaling).
  value : whole-job LCQPs/s, inputs already resident in HBM when the timed region starts (CUDA events).
  e2e   : the same metric through the C ABI with HOST (pinned) buffers: H2D of the per-instance inputs, run, D2H of
          x, the duals and the statistics inside the timed region.
  parity_subset : outside the timed region, `--parity` random instances of the batch are solved once more with
          perturbStep off (the reference's perturbation draws libc rand() seeded by time(NULL)) on the GPU and by the
          reference's qpOASES run (oracle/_ref); mismatches in ReturnValue / stationarity type / outer and total
          iteration counts / x (1e-6) are counted.
  --impl reference : the reference's own CPU implementation of the path (oracle/_ref = unmodified LCQPow +
          qpOASES + OSQP when it was built, else the plain-C oracle port) on all host cores, same config.

    python bench.py --gpus 1 --steps 3 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "lcqps_solved_per_sec"
UNIT = "LCQP/s"
STAT_TOL = 10e-3  # examples/OptimizeOnCircle.cpp:45


class Config:
    """A workload: generator, dimensions, option overrides, the reference subsolver that ships with it."""

    def __init__(self, name):
        self.name = name
        if name == "c2":
            self.nV, self.nC, self.nComp = 202, 101, 100
            self.over = {"stationarityTolerance": STAT_TOL}
            self.default_batch = 1 << 20
            self.workload = "C2 OptimizeOnCircle N=100 (nV=202,nC=101,nComp=100), shared Q/A/L/R, per-instance g/x0"
            self.ref_solver_shipped = 2   # OSQP_SPARSE (examples/OptimizeOnCircle.cpp:44)
            self.ref_solver_parity = 1    # QPOASES_SPARSE: the exact-QP flavour the CUDA path reproduces
            self.cpu_per_core = 48
        elif name == "c5":
            self.nV, self.nC, self.nComp = 64, 16, 32
            self.over = {}
            self.default_batch = 1 << 17
            self.workload = "C5 random dense LCQPs (n=64, 32 pairs, nC=16), every matrix per instance"
            self.ref_solver_shipped = 0   # QPOASES_DENSE
            self.ref_solver_parity = 0
            self.cpu_per_core = 256
        elif name == "c3":
            self.nV, self.nC, self.nComp = 151, 50, 100
            self.over = {}
            self.default_batch = 10000
            self.workload = "C3 examples/example_data (nV=151,nC=50,nComp=100, box) x batch with perturbed g and lbA=ubA, shared Q/A/L/R"
            self.ref_solver_shipped = 1   # QPOASES_SPARSE (examples/solve_lcqp_from_file.cpp:128)
            self.ref_solver_parity = 1
            self.cpu_per_core = 16
        elif name == "c4":
            self.nV, self.nC, self.nComp = 1000, 300, 500
            self.over = {}
            self.default_batch = 4096
            self.workload = "C4 sparse banded LCQPs (n=1000, 500 pairs, nC=300), one pattern, per-instance values, CSC door, OSQP-ADMM flavour"
            self.ref_solver_shipped = 2
            self.ref_solver_parity = 2
            self.cpu_per_core = 2
        else:
            raise SystemExit("unknown --config " + name)
        self.sparse = (name == "c4")

    def generate(self, n, lo=0, hi=None, dense=True):
        from lcqpow_b200 import problems as P
        hi = n if hi is None else hi
        if self.name == "c4":
            sb = P.sparse_banded_batch(hi - lo, lo=lo)
            return sb.to_dense(0, hi - lo) if dense else sb
        if self.name == "c2":
            return P.circle_batch_fast(n, lo=lo, hi=hi)   # (only the shard's g / x0 are materialised)
        if self.name == "c5":
            return P.dense_random_batch(hi, seed_lo=lo) if lo else P.dense_random_batch(hi)
        data = dict(np.load(os.path.join(ROOT, "tests", "golden", "example_data.npz")))
        return P.example_data_batch(data, n).slice(lo, hi)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops_sustained", 1400.0), "measured"
    return 6650.0, 1590.0, "fallback"


# ------------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md recipe)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                                         text=True)
        except Exception:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [t.strip() for t in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU arms (rank 0 only).  The only place where bench.py executes oracle/.
# ------------------------------------------------------------------------------------------------
_REFLIB = None   # loaded once in the parent (the driver's loaded-library hook sees it); fork()ed workers inherit it


def cpu_kind():
    from oracle import pyref
    global _REFLIB
    if pyref.have_ref():
        if _REFLIB is None:
            _REFLIB = pyref.RefLib()
        return "reference"
    if not os.path.exists(pyref.ORACLE_SO):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "oracle"], check=True)
    return "port"


def _cpu_worker(args):
    """Solve instances [lo, hi) of the config's family serially in this process.  perturb < 0: options default."""
    kind, cfg_name, lo, hi, total, solver, perturb, want_x = args
    from oracle import pyref
    cfg = Config(cfg_name)
    pb = cfg.generate(total, lo, hi)
    if kind == "reference":
        lib = _REFLIB if _REFLIB is not None else pyref.RefLib()
        o = lib.default_options(qpSolver=solver, **cfg.over)
    else:
        lib = pyref.OracleLib()
        o = lib.default_options(**cfg.over)
    if perturb >= 0:
        o.perturbStep = perturb
    t = time.perf_counter()
    s = lib.solve_batch(pb, o)
    dt = time.perf_counter() - t
    return dt, int((s.res["ret"] == 0).sum()), hi - lo, (s.x if want_x else None), (s.res if want_x else None)


def cpu_pass(kind, cfg, cores, per_core, total, pool):
    """One pass: `cores` processes, each solving `per_core` instances serially.  Returns (LCQP/s, solved, n, wall)."""
    jobs = [(kind, cfg.name, c * per_core, (c + 1) * per_core, total, cfg.ref_solver_shipped, -1, False) for c in range(cores)]
    t = time.perf_counter()
    out = pool.map(_cpu_worker, jobs)
    wall = time.perf_counter() - t
    return sum(o[2] for o in out) / wall, sum(o[1] for o in out), sum(o[2] for o in out), wall


def cpu_sample_desc(kind, cfg, cores, per_core):
    names = {0: "QPOASES_DENSE", 1: "QPOASES_SPARSE", 2: "OSQP_SPARSE"}
    what = (f"unmodified reference (oracle/_ref: LCQPow + qpOASES 3.2 + OSQP 0.6.2), QPSolver::{names[cfg.ref_solver_shipped]} "
            "as the reference's example ships it") if kind == "reference" else "plain-C oracle port (oracle/lcqp_oracle.c)"
    return (f"first {cores * per_core} instances of the same {cfg.name.upper()} batch, {per_core} per process, {cores} processes, {what}")


def run_reference_arm(args, cfg, rank, world):
    if rank != 0:
        return
    kind = cpu_kind()
    cores = len(os.sched_getaffinity(0))
    per_core = args.cpu_per_core or cfg.cpu_per_core
    total = cores * per_core
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        for _ in range(args.warmup):
            cpu_pass(kind, cfg, cores, max(1, per_core // 4), total, pool)
        t = time.perf_counter()
        n = solved = 0
        for _ in range(args.steps):
            _, s, k, _ = cpu_pass(kind, cfg, cores, per_core, total, pool)
            n += k
            solved += s
        wall = time.perf_counter() - t
    v = n / wall
    line = {"metric": METRIC, "value": v, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": cfg.workload, "instances_per_step": cores * per_core,
                       "sample": "bounded sample of the workload: the first instances_per_step instances of the same family",
                       **cfg.over},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": cpu_sample_desc(kind, cfg, cores, per_core)},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "solved_frac": solved / max(1, n)}
    print(json.dumps(line))


def parity_subset(cfg, prob_cls, L, pb, batch, lo_global, n_check, device, seed=12345):
    """`n_check` random instances of this rank's batch, perturbStep off: GPU against the reference's qpOASES run."""
    kind = cpu_kind()
    rng = np.random.default_rng(seed)
    idx = np.sort(rng.choice(batch, size=min(n_check, batch), replace=False))
    sub = pb.normalised()
    kw = {}
    for f in L.api.FIELDS:
        a = getattr(sub, f)
        kw[f] = None if a is None else (a if f in pb.shared else np.ascontiguousarray(a[idx]))
    import dataclasses
    sub = dataclasses.replace(sub, batch=len(idx), **kw)
    prob = prob_cls(cfg.nV, cfg.nC, cfg.nComp, len(idx), device=device)
    o = L.Options()
    o.setPerturbStep(False)
    for k, v in cfg.over.items():
        getattr(o, "set" + k[0].upper() + k[1:])(v)
    prob.setOptions(o)
    assert prob.loadBatch(sub) == 0
    prob.runSolver()
    x, st = prob.getPrimalSolution(), prob.getOutputStatistics()
    prob.close()
    # reference: the same instances, one slice per core
    from oracle import pyref
    cores = len(os.sched_getaffinity(0))
    lib = _REFLIB if kind == "reference" else pyref.OracleLib()
    chunks = np.array_split(np.arange(len(idx)), cores)
    def solve_chunk(ch):
        kw2 = {}
        for f in L.api.FIELDS:
            a = getattr(sub, f)
            kw2[f] = None if a is None else (a if f in pb.shared else np.ascontiguousarray(a[ch]))
        s2 = dataclasses.replace(sub, batch=len(ch), **kw2)
        oo = lib.default_options(perturbStep=0, **cfg.over) if kind != "reference" else lib.default_options(perturbStep=0, qpSolver=cfg.ref_solver_parity, **cfg.over)
        return lib.solve_batch(s2, oo)
    # threads would serialise on the reference's globals: fork one process per chunk
    ctx = mp.get_context("fork")
    global _PARITY_JOB
    _PARITY_JOB = solve_chunk
    with ctx.Pool(cores) as pool:
        outs = pool.map(_parity_call, [c for c in chunks if len(c)])
    rx = np.concatenate([o.x for o in outs]); rres = np.concatenate([o.res for o in outs])
    mism = mism_end = 0
    for b in range(len(idx)):
        end = (rres["ret"][b] == st["ret"][b] and rres["status"][b] == st["status"][b] and rres["iterOuter"][b] == st["iterOuter"][b])
        if end and rres["ret"][b] == 0:
            end = np.abs(rx[b] - x[b]).max() <= 1e-6 * max(1.0, np.abs(rx[b]).max())
        same = end and rres["iterTotal"][b] == st["iterTotal"][b]
        mism += (not same)
        mism_end += (not end)
    return {"n": int(len(idx)), "mismatches": int(mism), "mismatches_ignoring_iterTotal": int(mism_end), "checker": kind,
            "criterion": "ReturnValue, stationarity type, iterOuter, iterTotal identical; x within 1e-6 relative; perturbStep off"}


def parity_subset_osqp(cfg, L, batch, lo_global, n_check, device):
    """OSQP flavour: the first `n_check` instances of this rank's batch, perturbStep off and adaptive_rho_interval = 25 on
    both sides (the reference's default interval is wall-clock driven), against the reference's OSQP_SPARSE run.  Bar of
    tests/test_osqp_flavour.py: an instance the reference solves must be reproduced with identical ReturnValue,
    stationarity type, iteration counts (ADMM iterations included) and x to 1e-6; one it fails must fail here too."""
    kind = cpu_kind()
    if kind != "reference":
        return {"n": 0, "mismatches": None, "error": "the OSQP flavour is checked against oracle/_ref only"}
    n = min(n_check, batch)
    src = cfg.generate(lo_global + n, lo_global, lo_global + n, dense=False) if cfg.sparse else cfg.generate(lo_global + n, lo_global, lo_global + n).normalised()
    prob = L.LCQProblemBatch(cfg.nV, cfg.nC, cfg.nComp, n, device=device)
    o = L.Options()
    o.setPerturbStep(False)
    for k, v in cfg.over.items():
        getattr(o, "set" + k[0].upper() + k[1:])(v)
    o.setOSQPADMM(True, adaptive_rho_interval=25)
    prob.setOptions(o)
    if cfg.sparse:
        assert prob.loadCSC(src.Q, src.g, src.L, src.R, A=src.A, lbA=src.lbA, ubA=src.ubA, batch=n, shared=tuple(src.shared)) == 0, prob._err()
        dense = src.to_dense(0, n).normalised()
    else:
        assert prob.loadBatch(src) == 0, prob._err()
        dense = src
    prob.runSolver()
    x, st = prob.getPrimalSolution(), prob.getOutputStatistics()
    prob.close()
    import dataclasses
    cores = min(len(os.sched_getaffinity(0)), n)
    chunks = np.array_split(np.arange(n), cores)

    def solve_chunk(ch):
        kw2 = {}
        for f in L.api.FIELDS:
            a = getattr(dense, f)
            kw2[f] = None if a is None else (a if f in dense.shared else np.ascontiguousarray(a[ch]))
        s2 = dataclasses.replace(dense, batch=len(ch), **kw2)
        return _REFLIB.solve_batch(s2, _REFLIB.default_options(perturbStep=0, qpSolver=2, osqp_adaptive_rho_interval=25, **cfg.over))
    global _PARITY_JOB
    _PARITY_JOB = solve_chunk
    with mp.get_context("fork").Pool(cores) as pool:
        outs = pool.map(_parity_call, [c for c in chunks if len(c)])
    rx = np.concatenate([o.x for o in outs]); rres = np.concatenate([o.res for o in outs])
    mism = 0
    for b in range(n):
        if rres["ret"][b] == 0:
            same = all(rres[f][b] == st[f][b] for f in ("ret", "status", "iterOuter", "iterTotal", "subproblemIter"))
            same = same and np.abs(rx[b] - x[b]).max() <= 1e-6 * max(1.0, np.abs(rx[b]).max())
        else:
            same = st["ret"][b] != 0
        mism += (not same)
    return {"n": int(n), "mismatches": int(mism), "checker": kind, "reference_solved": int((rres["ret"] == 0).sum()),
            "criterion": "instances the reference's OSQP run solves: ReturnValue, stationarity type, iterOuter, iterTotal, ADMM iterations "
                         "identical and x within 1e-6 relative; instances it fails: a failure code here too; perturbStep off, adaptive_rho_interval 25",
            "note": "on the circle family the reference's own OSQP trajectory is round-off determined wherever polish fails (Hessian curvatures of 5e-12): about half of the "
                    "instances are reproduced bit for bit, the others end at another stationary point (DESIGN.md section 4); dense and C4 families are reproduced"}


def work_probe(args, cfg):
    """Child process of the bench (rank 0): the first `--work-probe` instances of the workload through the COUNTING build of
    the library; prints {"macs_per_lcqp", "bytes_per_lcqp", "l2_gbs", "n"} -- the active-set kernel's own count of its fp64
    multiply-adds and of the bytes its dense products read / write (lcqp_cuda_last_work), and the L2 read-bandwidth probe."""
    import ctypes as C
    import lcqpow_b200 as L
    n = args.work_probe
    pb = cfg.generate(n, 0, n).normalised()
    prob = L.LCQProblemBatch(cfg.nV, cfg.nC, cfg.nComp, n, device=int(os.environ.get("LOCAL_RANK", "0")))
    o = L.Options()
    for k, v in cfg.over.items():
        getattr(o, "set" + k[0].upper() + k[1:])(v)
    o.setPerturbStep(bool(args.perturb))
    prob.setOptions(o)
    assert prob.loadBatch(pb) == 0
    prob.runSolver()
    macs, wbytes, l2 = C.c_double(0.0), C.c_double(0.0), C.c_double(0.0)
    rc = prob.lib.lcqp_cuda_last_work(prob.h, C.byref(macs), C.byref(wbytes))
    prob.lib.lcqp_cuda_measure_l2_gbs(int(os.environ.get("LOCAL_RANK", "0")), C.byref(l2))
    print(json.dumps({"rc": rc, "macs_per_lcqp": macs.value / n, "bytes_per_lcqp": wbytes.value / n, "l2_gbs": l2.value, "n": n}))


_PARITY_JOB = None


def _parity_call(ch):
    return _PARITY_JOB(ch)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c2", choices=["c2", "c5", "c3", "c4"])
    ap.add_argument("--flavour", default="exact", choices=["exact", "osqp"], help="QP subsolver flavour: the exact-vertex solver "
                    "(qpOASES semantics) or the OSQP restatement (ADMM + polish); c4 always runs the latter")
    ap.add_argument("--batch", type=int, default=0, help="instances per GPU per step (0: the configuration's own size: "
                    "2^20 for c2, 2^17 for c5, 10000 for c3)")
    ap.add_argument("--cpu-per-core", type=int, default=0, help="instances per host process in the CPU baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--parity", type=int, default=1024, help="instances of the parity subset (0: skip)")
    ap.add_argument("--perturb", type=int, default=1)
    ap.add_argument("--work-probe", type=int, default=0, help="internal: solve this many instances with the counting build of the "
                    "library (LCQP_CUDA_LIB = liblcqp_cuda_work.so) and print its work counters per LCQP")
    args = ap.parse_args()
    cfg = Config(args.config)
    if args.work_probe > 0:
        work_probe(args, cfg)
        return
    osqp = cfg.sparse or args.flavour == "osqp"
    if osqp:
        cfg.ref_solver_parity = 2
        if not cfg.sparse:
            cfg.workload += ", OSQP-ADMM flavour"

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference_arm(args, cfg, rank, world)
        return

    import torch
    import torch.distributed as dist
    import lcqpow_b200 as L

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    hbm_gbs, bf16_tf, peak_src = load_peaks()
    NV, NC, NCOMP = cfg.nV, cfg.nC, cfg.nComp

    def make_options():
        o = L.Options()
        for k, v in cfg.over.items():
            getattr(o, "set" + k[0].upper() + k[1:])(v)
        o.setPerturbStep(bool(args.perturb))
        if osqp:
            o.setOSQPADMM(True)
        return o

    batch = args.batch if args.batch > 0 else cfg.default_batch

    # ---- inputs: the same family on every rank, instances [rank*batch, (rank+1)*batch) ------------
    from lcqpow_b200 import sharding
    import ctypes as C
    lo, hi = sharding.shard_range(batch * world, rank, world)
    prob = L.LCQProblemBatch(NV, NC, NCOMP, batch, device=local_rank)
    assert prob.setOptions(make_options()) == 0
    prob.setInstanceOffset(lo)   # perturbStep draws are keyed by the global instance index
    nD = NV + NC + 2 * NCOMP
    x_pin_t = torch.empty((batch, NV), dtype=torch.float64).pin_memory()
    y_pin_t = torch.empty((batch, nD), dtype=torch.float64).pin_memory()
    st_pin_t = torch.empty((batch, L.api.STATS_DTYPE.itemsize), dtype=torch.uint8).pin_memory()
    d2h = x_pin_t.numel() * 8 + y_pin_t.numel() * 8 + st_pin_t.numel()
    stream = torch.cuda.current_stream(dev)

    def fetch_results():
        for fn, buf in ((prob.lib.lcqp_cuda_get_primal, x_pin_t), (prob.lib.lcqp_cuda_get_dual, y_pin_t), (prob.lib.lcqp_cuda_get_stats, st_pin_t)):
            rc = fn(prob.h, C.c_void_p(buf.data_ptr()))
            assert rc == 0, rc

    if cfg.sparse:
        # the CSC door takes host arrays; a load converts and uploads them, the run works on the resident copy
        pb = cfg.generate(batch * world, lo, hi, dense=False)
        shared = tuple(pb.shared)

        def load_csc():
            rc = prob.loadCSC(pb.Q, pb.g, pb.L, pb.R, A=pb.A, lbA=pb.lbA, ubA=pb.ubA, batch=batch, shared=shared)
            assert rc == 0, prob._err()

        h2d = sum(t[2].nbytes for f, t in (("Q", pb.Q), ("L", pb.L), ("R", pb.R), ("A", pb.A)) if t is not None and f not in shared)
        h2d += sum(a.nbytes for f, a in (("g", pb.g), ("lbA", pb.lbA), ("ubA", pb.ubA)) if a is not None and f not in shared)
        load_csc()

        def step_resident():
            prob.runSolver(stream=stream.cuda_stream, sync=False)

        def step_e2e():
            load_csc()
            prob.runSolver(stream=0, sync=False)
            fetch_results()
    else:
        pb = cfg.generate(batch * world, lo, hi).normalised()
        shared = tuple(pb.shared)
        # device-resident copies (torch is only the allocator here)
        dev_t = {}
        for f in L.api.FIELDS:
            a = getattr(pb, f)
            if a is not None:
                dev_t[f] = torch.from_numpy(a).to(dev)
        ptrs = {f: t.data_ptr() for f, t in dev_t.items()}
        # pinned host copies of the per-instance inputs for the e2e leg
        pin = {}
        for f in L.api.FIELDS:
            a = getattr(pb, f)
            if a is None:
                pin[f] = None
            elif f in shared:
                pin[f] = a
            else:
                t = torch.from_numpy(a).pin_memory()
                pin[f] = t.numpy()
                pin["_keep_" + f] = t
        h2d = sum(pin[f].nbytes for f in L.api.FIELDS if pin[f] is not None and f not in shared)

        if osqp:
            # the OSQP flavour analyses the pattern at load time on the host: load once, the runs work on the resident copy
            rc = prob.loadLCQP(**{f: pin[f] for f in L.api.FIELDS}, batch=batch, shared=shared)
            assert rc == 0, prob._err()

        def step_resident():
            if not osqp:
                rc = prob.loadDevicePointers(ptrs, batch, shared)
                assert rc == 0, rc
            prob.runSolver(stream=stream.cuda_stream, sync=False)

        def step_e2e():
            rc = prob.loadLCQP(**{f: pin[f] for f in L.api.FIELDS}, batch=batch, shared=shared)
            assert rc == 0, rc
            prob.runSolver(stream=0, sync=False)
            fetch_results()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident leg ------------------------------------------------------------------------
    n_warm = max(args.warmup, 1) if cfg.sparse else max(args.warmup, 3)
    for _ in range(n_warm):
        step_resident()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = prob.launchCount()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    k_ms_sum = 0.0
    for _ in range(args.steps):
        step_resident()
        # duration of the dominant kernel of this step from the library's own CUDA events on the launching stream
        # (reading them waits for the launch; the next step would wait for it anyway: a load reuses the staging state)
        k_ms_sum += prob.lastRunMs()[0]
    e1.record(stream)
    barrier()
    elapsed_ms = e0.elapsed_time(e1)
    launches = prob.launchCount() - l0
    k_ms = k_ms_sum / args.steps
    clocks = sampler.stop()
    st = prob.getOutputStatistics()
    x = prob.getPrimalSolution()
    grid, smem_b, mE = prob.lastLaunchInfo()

    # ---- e2e leg ------------------------------------------------------------------------------------
    # (c4: a step is a minute long and the resident steps above have warmed the kernel; its load path ran once already)
    # The end-to-end rate is measured over at most five steps (a step of the full C2 batch is 11 s and the driver's
    # scaling run gives each N 870 s for warm-up + K timed steps + this leg): `e2e.steps` says how many.
    for _ in range(0 if cfg.sparse else 1):
        step_e2e()
    e2e_steps = min(args.steps, 5)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_e2e()
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0

    # max over ranks
    elapsed_ms, e2e_ms, k_ms = sharding.reduce_scalars([elapsed_ms, e2e_s * 1e3, float(k_ms)], "max", dev)
    n_solved, n_units, n_total_it, n_outer_it, n_sub = sharding.reduce_scalars(
        [float((st["ret"] == 0).sum()), float(st["kktSolves"].sum()) + float(st["admmIters"].sum()),
         float(st["iterTotal"].sum()), float(st["iterOuter"].sum()), float(st["subproblemIter"].sum())], "sum", dev)

    par = None
    if args.parity > 0 and rank == 0:
        try:
            if osqp:
                par = parity_subset_osqp(cfg, L, batch, lo, min(args.parity, 32 if cfg.sparse else 256), local_rank)
            else:
                par = parity_subset(cfg, L.LCQProblemBatch, L, pb, batch, lo, args.parity, local_rank)
        except Exception as ex:  # the checker is absent on this box: say so, do not fail the measurement
            par = {"n": 0, "mismatches": None, "error": repr(ex)}

    if rank == 0:
        nb = batch * world
        ret_hist = {int(k): int(v) for k, v in zip(*np.unique(st["ret"], return_counts=True))}
        # the metric counts SOLVED LCQPs (terminal ReturnValue SUCCESSFUL_RETURN); every step solves the same batch
        if n_solved < (0.25 if osqp else 0.5) * nb:
            raise SystemExit(f"bench.py: only {n_solved:.0f} of {nb} instances were solved -- the CUDA path is broken, "
                             "refusing to report a throughput")
        counted = n_solved
        value = counted * args.steps / (elapsed_ms * 1e-3)
        e2e_v = counted * e2e_steps / (e2e_ms * 1e-3)
        units_per_launch = n_units / world   # explicit-inverse solves (homotopy steps + polish passes) of one launch on one GPU
        fp64 = C.c_double(0.0)
        prob.lib.lcqp_cuda_measure_fp64_tflops(local_rank, C.byref(fp64))
        # the work the parametric active-set kernel actually did in its last launch (counted by the kernel: fp64
        # multiply-adds of its dense products and the bytes they read / write -- per-instance inverse, Tt columns and
        # prepared operators, all L2 resident), against the measured fp64 pipe and the measured L2 read bandwidth
        work = None
        work_lib = os.path.join(ROOT, "lcqpow_b200", "lib", "liblcqp_cuda_work.so")
        if not osqp and cfg.name != "c3" and os.path.exists(work_lib):
            # the counting build (3-5 % slower: not the one that is timed) on the first instances of the same workload, in a
            # child process, after the timed region; the counts are a property of the instances, scaled to this launch
            import subprocess
            n_probe = min(batch, 8192)
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), "--config", cfg.name, "--work-probe", str(n_probe), "--perturb", str(args.perturb)],
                                   env=dict(os.environ, LCQP_CUDA_LIB=work_lib, LOCAL_RANK=str(local_rank), WORLD_SIZE="1", RANK="0"),
                                   capture_output=True, text=True, timeout=600)
                wp = json.loads(r.stdout.strip().splitlines()[-1])
            except Exception as ex:
                wp = {"rc": -1, "error": repr(ex)}
            if wp.get("rc") == 0 and wp["macs_per_lcqp"] > 0:
                tf = 2.0 * wp["macs_per_lcqp"] * batch / (k_ms * 1e-3) / 1e12
                gbs = wp["bytes_per_lcqp"] * batch / (k_ms * 1e-3) / 1e9
                work = {"fp64_macs_per_lcqp": wp["macs_per_lcqp"], "l2_bytes_per_lcqp": wp["bytes_per_lcqp"],
                        "fp64_achieved_tflops": tf, "fp64_frac": tf / fp64.value if fp64.value > 0 else None,
                        "l2_achieved_gbs": gbs, "l2_peak_gbs_measured": wp["l2_gbs"], "l2_frac": gbs / wp["l2_gbs"] if wp["l2_gbs"] > 0 else None,
                        "note": f"dense products of the active-set kernel counted on the device by the counting build (lcqp_cuda_last_work) on the first {n_probe} "
                                "instances of this workload, per LCQP, times this launch's instances over its kernel time: algorithmic minimum"}
            else:
                work = {"error": wp.get("error", "work probe failed"), "rc": wp.get("rc")}
        if osqp:
            # SURVEY.md 8(d) row "OSQP-ADMM iteration": HBM bound, bytes per ADMM iteration of one instance =
            # 16 nnz(L) (values + indices of the factor, forward and backward sweep read it once each as 8 + 8)
            # + 8 N (the KKT right-hand side / solution) + 8 (3n + 5m) (x, x_prev, x_tilde; z, z_prev, z_tilde, y, rho)
            oN, onnzL, olev, omode = prob.osqpInfo()
            m = NC + 2 * NCOMP
            bytes_per_it = 16.0 * onnzL + 8.0 * oN + 8.0 * (3 * NV + 5 * m)
            n_admm = float(st["admmIters"].sum())
            achieved = bytes_per_it * n_admm / (k_ms * 1e-3) / 1e9
            roof = {"bound": "hbm", "achieved": achieved, "peak": hbm_gbs, "unit": "GB/s", "frac": achieved / hbm_gbs, "traffic": None,
                    "bytes_per_unit": bytes_per_it, "unit_def": "one ADMM iteration of one instance: 16 nnz(L) + 8 N + 8 (3n + 5m) bytes (SURVEY.md 8d)",
                    "kkt_order": oN, "nnz_L": onnzL, "levels_per_solve": olev, "admm_iters_per_lcqp": n_admm / batch,
                    "factorisations_per_lcqp": float(st["kktSolves"].sum()) / batch,
                    "launch_mode": "one warp per instance" if omode else "one thread per instance"}
        elif cfg.name == "c5":
            # SURVEY.md 8(d) row "Per-instance dense KKT in SMEM": HBM bound, 8 (n^2 + m n + 2n + 2m + n) in + 8 (n + m + 4) out
            m = NC + 2 * NCOMP
            bytes_per_lcqp = 8.0 * (NV * NV + m * NV + 2 * NV + 2 * m + NV) + 8.0 * (NV + m + 4)
            achieved = bytes_per_lcqp * batch / (k_ms * 1e-3) / 1e9
            roof = {"bound": "hbm", "achieved": achieved, "peak": hbm_gbs, "unit": "GB/s", "frac": achieved / hbm_gbs, "traffic": None,
                    "bytes_per_unit": bytes_per_lcqp, "unit_def": "one LCQP (SURVEY.md 8d: 77 728 B at nC=16)"}
        else:
            # SURVEY.md 8(d) row "Shared-factor multi-RHS (C2)": one unit = one KKT solve for one instance = 2 N^2 flop,
            # N = nV + nC + 2 nComp (+ nV box rows), against the tensor pipe
            N = NV + NC + 2 * NCOMP + (NV if cfg.name == "c3" else 0)
            achieved = units_per_launch * 2.0 * N * N / (k_ms * 1e-3) / 1e12
            roof = {"bound": "tensor", "achieved": achieved, "peak": bf16_tf, "unit": "TFLOP/s", "frac": achieved / bf16_tf, "traffic": None,
                    "flop_per_unit": 2.0 * N * N, "unit_def": "one KKT solve of one instance counted as 2 N^2 flop (SURVEY.md 8d)"}
        tpath = os.path.join(ROOT, "profiles", "dram_traffic.json")
        if os.path.exists(tpath):
            try:
                tj = json.load(open(tpath))
                if tj.get("config", "c2") == cfg.name:
                    roof["traffic"] = float(tj["dram_bytes_per_lcqp"]) * batch
            except Exception:
                pass
        roof.update({"kernel": ("lcqp_osqpw_kernel" if (osqp and roof.get("launch_mode", "").startswith("one warp")) else "lcqp_osqp_kernel") if osqp else "lcqp_pas_kernel", "kernel_ms": k_ms, "units_per_launch": units_per_launch,
                     "peak_source": f"{peak_src} (MEASURED_PEAKS.json); the arithmetic is fp64 SIMT -- see `fp64` for the pipe it runs on",
                     # the work actually done, against the pipes it runs on (model counts of DESIGN.md section 5)
                     "fp64": {"peak_tflops_measured": fp64.value,
                              "note": "fp64 FMA probe of this device (lcqp_cuda_measure_fp64_tflops); DESIGN.md 5 has the MAC model"},
                     "work": work})
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": n_warm,
                "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": cfg.workload, "instances_per_gpu_per_step": batch, "perturbStep": bool(args.perturb), "warmup_steps": n_warm,
                           "parallelism": f"instance-sharded x{world}, no collective on the data path",
                           "l2": "inputs+outputs per step exceed L2 (%.0f MB)" % ((h2d + d2h) / 1e6), **cfg.over},
                "clocks": clocks, "gpu_launches": int(launches),
                "e2e": {"value": e2e_v, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "steps": e2e_steps},
                "roofline": roof,
                "solved_frac": n_solved / nb, "return_values": ret_hist,
                "mean_iter_outer": n_outer_it / nb, "mean_iter_total": n_total_it / nb,
                "mean_subproblem_iters": n_sub / nb, "kkt_solves_per_lcqp": n_units / nb,
                "launch": {"groups": grid, "smem_bytes_per_cta": smem_b, "eliminated_equality_rows": mE},
                "parity_subset": par}
        if cfg.name == "c2" and not osqp:
            line["instance0_matches_shipped_solution"] = bool(abs(x[0, 0] - 0.181110968) < 1e-6 and abs(x[0, 1] + 0.983483383) < 1e-6 and st["status"][0] == 4)
        if not args.no_cpu_baseline and world == 1:
            kind = cpu_kind()
            cores = len(os.sched_getaffinity(0))
            per_core = args.cpu_per_core or cfg.cpu_per_core
            with mp.get_context("fork").Pool(cores) as pool:
                v, s, n, wall = cpu_pass(kind, cfg, cores, per_core, cores * per_core, pool)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": kind,
                                    "sample": cpu_sample_desc(kind, cfg, cores, per_core), "seconds": wall,
                                    "solved_frac": s / max(1, n)}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

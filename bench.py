#!/usr/bin/env python
"""bench.py -- throughput of the LCQP hot path (LCQProblem::runSolver + QP subsolver) on B200.

Metric (BASELINE.json): LCQPs solved / second, batched, on N GPUs of one node.
Workload: config C2 of SURVEY.md 8(d) -- OptimizeOnCircle-shaped LCQPs (nV=202, nC=101, nComp=100) that
share Q/A/L/R/lbA/ubA and differ in g and x0 (examples/OptimizeOnCircle.cpp:62-99 with a random x_ref per
instance), stationarityTolerance = 1e-2 as the example sets it, perturbStep on (the default).

One "step" = one pass of the hot path over one batch of `--batch` instances per GPU (weak scaling).
  value : whole-job LCQPs/s, inputs already resident in HBM when the timed region starts (CUDA events).
  e2e   : the same metric through the C ABI with HOST (pinned) buffers: H2D of g/x0, run, D2H of x and the
          statistics inside the timed region.
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
NV, NC, NCOMP = 202, 101, 100
STAT_TOL = 10e-3  # examples/OptimizeOnCircle.cpp:45


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
def _cpu_worker(args):
    kind, lo, hi, seed0 = args
    from lcqpow_b200 import problems as P
    from oracle import pyref
    pb = P.circle_batch_fast(hi, seed0=seed0).slice(lo, hi)
    if kind == "reference":
        lib = pyref.RefLib()
        o = lib.default_options(qpSolver=pyref.OSQP_SPARSE, stationarityTolerance=STAT_TOL)  # as shipped (:44-45)
    else:
        lib = pyref.OracleLib()
        o = lib.default_options(stationarityTolerance=STAT_TOL)
    t = time.perf_counter()
    s = lib.solve_batch(pb, o)
    return time.perf_counter() - t, int((s.res["ret"] == 0).sum()), hi - lo


def cpu_kind():
    from oracle import pyref
    if pyref.have_ref():
        return "reference"
    if not os.path.exists(pyref.ORACLE_SO):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "oracle"], check=True)
    return "port"


def cpu_pass(kind: str, cores: int, per_core: int, seed0: int, pool):
    """One pass: `cores` processes, each solving `per_core` instances serially.  Returns (LCQP/s, solved, n, wall)."""
    jobs = [(kind, c * per_core, (c + 1) * per_core, seed0) for c in range(cores)]
    t = time.perf_counter()
    out = pool.map(_cpu_worker, jobs)
    wall = time.perf_counter() - t
    n = sum(o[2] for o in out)
    solved = sum(o[1] for o in out)
    return n / wall, solved, n, wall


def cpu_sample_desc(kind, cores, per_core):
    what = ("unmodified reference (oracle/_ref: LCQPow + OSQP 0.6.2, QPSolver::OSQP_SPARSE as shipped in "
            "examples/OptimizeOnCircle.cpp:44)") if kind == "reference" else "plain-C oracle port (oracle/lcqp_oracle.c)"
    return f"first {cores * per_core} instances of the same C2 batch, {per_core} per process, {cores} processes, {what}"


# ------------------------------------------------------------------------------------------------
def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    kind = cpu_kind()
    cores = len(os.sched_getaffinity(0))
    per_core = args.cpu_per_core
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        for _ in range(args.warmup):
            cpu_pass(kind, cores, max(1, per_core // 4), 20000, pool)
        t = time.perf_counter()
        n = solved = 0
        for _ in range(args.steps):
            _, s, k, _ = cpu_pass(kind, cores, per_core, 20000, pool)
            n += k
            solved += s
        wall = time.perf_counter() - t
    v = n / wall
    line = {"metric": METRIC, "value": v, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "C2 OptimizeOnCircle N=100 (nV=202,nC=101,nComp=100), shared Q/A/L/R, per-instance g/x0",
                       "instances_per_step": cores * per_core, "stationarityTolerance": STAT_TOL},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": cpu_sample_desc(kind, cores, per_core)},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "solved_frac": solved / max(1, n)}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=0, help="instances per GPU per step (0: largest power of two <= 2^20 "
                    "whose step is estimated to take <= --step-seconds)")
    ap.add_argument("--step-seconds", type=float, default=8.0)
    ap.add_argument("--cpu-per-core", type=int, default=48, help="instances per host process in the CPU baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--perturb", type=int, default=1)
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import lcqpow_b200 as L
    from lcqpow_b200 import problems as P

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    hbm_gbs, bf16_tf, peak_src = load_peaks()

    def make_options():
        o = L.Options()
        o.setStationarityTolerance(STAT_TOL)
        o.setPerturbStep(bool(args.perturb))
        return o

    # ---- choose the per-GPU batch ---------------------------------------------------------------
    batch = args.batch
    if batch <= 0:
        probe = 2048
        pb = P.circle_batch_fast(probe)
        pr = L.LCQProblemBatch(NV, NC, NCOMP, probe, device=local_rank)
        pr.setOptions(make_options())
        pr.loadBatch(pb)
        pr.runSolver()
        t = time.perf_counter()
        pr.runSolver()
        rate = probe / (time.perf_counter() - t)
        pr.close()
        batch = 1 << 20
        while batch > 4096 and batch / rate > args.step_seconds:
            batch >>= 1
        if world > 1:
            tb = torch.tensor([batch], device=dev, dtype=torch.int64)
            dist.all_reduce(tb, op=dist.ReduceOp.MIN)
            batch = int(tb.item())

    # ---- inputs: the same family on every rank, instances [rank*batch, (rank+1)*batch) ------------
    from lcqpow_b200 import sharding
    lo, hi = sharding.shard_range(batch * world, rank, world)
    pb_all = P.circle_batch_fast(batch * world)
    pb = pb_all.slice(lo, hi).normalised()
    del pb_all
    shared = tuple(pb.shared)
    prob = L.LCQProblemBatch(NV, NC, NCOMP, batch, device=local_rank)
    assert prob.setOptions(make_options()) == 0
    prob.setInstanceOffset(lo)   # perturbStep draws are keyed by the global instance index

    # device-resident copies (torch is only the allocator here)
    dev_t = {}
    for f in L.api.FIELDS:
        a = getattr(pb, f)
        if a is not None:
            dev_t[f] = torch.from_numpy(a).to(dev)
    ptrs = {f: t.data_ptr() for f, t in dev_t.items()}
    # pinned host copies of the per-instance inputs and pinned result buffers for the e2e leg
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
    x_pin_t = torch.empty((batch, NV), dtype=torch.float64).pin_memory()
    st_pin_t = torch.empty((batch, L.api.STATS_DTYPE.itemsize), dtype=torch.uint8).pin_memory()
    h2d = sum(pin[f].nbytes for f in L.api.FIELDS if pin[f] is not None and f not in shared)
    d2h = x_pin_t.numel() * 8 + st_pin_t.numel()

    stream = torch.cuda.current_stream(dev)

    def step_resident():
        rc = prob.loadDevicePointers(ptrs, batch, shared)
        assert rc == 0, rc
        prob.runSolver(stream=stream.cuda_stream, sync=False)

    import ctypes as C

    def step_e2e():
        rc = prob.loadLCQP(**{f: pin[f] for f in L.api.FIELDS}, batch=batch, shared=shared)
        assert rc == 0, rc
        prob.runSolver(stream=0, sync=False)
        rc = prob.lib.lcqp_cuda_get_primal(prob.h, C.c_void_p(x_pin_t.data_ptr()))
        assert rc == 0, rc
        rc = prob.lib.lcqp_cuda_get_stats(prob.h, C.c_void_p(st_pin_t.data_ptr()))
        assert rc == 0, rc

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident leg ------------------------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        step_resident()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = prob.launchCount()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    solve_ms = []
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        step_resident()
        # per-launch duration of the dominant kernel, from the library's own events on the launching stream
        # (queried after the loop would serialise nothing: cudaEventElapsedTime only reads)
    e1.record(stream)
    barrier()
    elapsed_ms = e0.elapsed_time(e1)
    launches = prob.launchCount() - l0
    k_ms, tot_ms = prob.lastRunMs()
    solve_ms.append(k_ms)
    clocks = sampler.stop()
    st = prob.getOutputStatistics()
    x = prob.getPrimalSolution()

    # ---- e2e leg ------------------------------------------------------------------------------------
    for _ in range(2):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0

    # max over ranks
    elapsed_ms, e2e_ms, k_ms = sharding.reduce_scalars([elapsed_ms, e2e_s * 1e3, float(k_ms)], "max", dev)
    n_solved, n_units, n_outer, n_sub = sharding.reduce_scalars(
        [float((st["ret"] == 0).sum()), float(st["kktSolves"].sum()) + float(st["admmIters"].sum()),
         float(st["iterTotal"].sum()), float(st["subproblemIter"].sum())], "sum", dev)
    total = batch * world * args.steps

    if rank == 0:
        # the metric counts SOLVED LCQPs (terminal ReturnValue SUCCESSFUL_RETURN), every step solves the same batch
        if n_solved < 0.5 * batch * world:
            raise SystemExit(f"bench.py: only {n_solved:.0f} of {batch * world} instances were solved -- the CUDA path is broken, "
                             "refusing to report a throughput")
        value = n_solved * args.steps / (elapsed_ms * 1e-3)
        e2e_v = n_solved * args.steps / (e2e_ms * 1e-3)
        # roofline of the dominant kernel (lcqp_solve_kernel), SURVEY.md 8(d) row "Shared-factor multi-RHS (C2)":
        # one unit = one KKT solve for one instance = 2 N^2 flop with N = nV + nC + 2 nComp = 503.
        N = NV + NC + 2 * NCOMP
        units_per_launch = n_units / world  # KKT solves (ADMM iterations + refined EQP passes) of one launch on one GPU
        flops_per_launch = units_per_launch * 2.0 * N * N
        achieved = flops_per_launch / (k_ms * 1e-3) / 1e12
        # DRAM traffic of the kernel per launch: bytes per LCQP from the committed ncu --set full capture
        # (profiles/dram_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum / instances of that launch)
        # scaled to the instances of one launch here; null when the summary is absent
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "dram_traffic.json")
        if os.path.exists(tpath):
            try:
                traffic = float(json.load(open(tpath))["dram_bytes_per_lcqp"]) * batch
            except Exception:
                traffic = None
        roof = {"bound": "tensor", "achieved": achieved, "peak": bf16_tf, "unit": "TFLOP/s", "frac": achieved / bf16_tf,
                "traffic": traffic, "kernel": "lcqp_solve_kernel", "kernel_ms": k_ms, "units_per_launch": units_per_launch,
                "flop_per_unit": 2.0 * N * N, "peak_source": f"bf16_tflops_sustained of {peak_src}; arithmetic is fp64 (see DESIGN.md)"}
        # parity spot check inside the bench: instance 0 is the shipped x_ref=(0.5,-0.6)
        ok0 = bool(abs(x[0, 0] - 0.181110968) < 1e-6 and abs(x[0, 1] + 0.983483383) < 1e-6 and st["status"][0] == 4)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": "C2 OptimizeOnCircle N=100 (nV=202,nC=101,nComp=100), shared Q/A/L/R, per-instance g/x0",
                           "instances_per_gpu_per_step": batch, "stationarityTolerance": STAT_TOL, "perturbStep": bool(args.perturb),
                           "parallelism": f"instance-sharded x{world}, no collective on the data path",
                           "l2": "inputs+outputs per step exceed L2 (%.0f MB)" % ((h2d + d2h) / 1e6)},
                "clocks": clocks, "gpu_launches": int(launches),
                "e2e": {"value": e2e_v, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
                "roofline": roof,
                "solved_frac": n_solved / (batch * world), "mean_outer_iters": n_outer / (batch * world),
                "mean_subproblem_iters": n_sub / (batch * world), "kkt_solves_per_lcqp": n_units / (batch * world),
                "instance0_matches_shipped_solution": ok0}
        if not args.no_cpu_baseline and world == 1:
            kind = cpu_kind()
            cores = len(os.sched_getaffinity(0))
            with mp.get_context("fork").Pool(cores) as pool:
                v, s, n, wall = cpu_pass(kind, cores, args.cpu_per_core, 20000, pool)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": kind,
                                    "sample": cpu_sample_desc(kind, cores, args.cpu_per_core), "seconds": wall,
                                    "solved_frac": s / max(1, n)}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""Input instances: the reference's shipped LCQPs and the synthetic batches of SURVEY.md section 8(d).

Every generator returns an ``LCQPBatch`` -- dense, row-major, fp64, exactly the argument list of
``LCQProblem::loadLCQP`` (/root/reference/include/LCQProblem.hpp:87-103) with a leading batch
dimension on every array that is not shared by the whole batch.  ``None`` means "NULL pointer" in the
reference's sense (e.g. no lbL -> bounds default to 0, /root/reference/src/LCQProblem.cpp:745-782).

Nothing in here touches /root/reference at run time: the literal fixtures are restated from the
reference's examples/tests (file:line cited per function) and example_data is shipped as a golden
fixture under tests/golden/.
"""
from __future__ import annotations

import dataclasses
import math
from typing import Optional

import numpy as np

FIELDS = ("Q", "g", "L", "R", "lbL", "ubL", "lbR", "ubR", "A", "lbA", "ubA", "lb", "ub", "x0", "y0")


@dataclasses.dataclass
class LCQPBatch:
    """A batch of LCQPs.  Arrays in ``shared`` have no batch dimension."""

    nV: int
    nC: int
    nComp: int
    batch: int
    Q: np.ndarray
    g: np.ndarray
    L: np.ndarray
    R: np.ndarray
    lbL: Optional[np.ndarray] = None
    ubL: Optional[np.ndarray] = None
    lbR: Optional[np.ndarray] = None
    ubR: Optional[np.ndarray] = None
    A: Optional[np.ndarray] = None
    lbA: Optional[np.ndarray] = None
    ubA: Optional[np.ndarray] = None
    lb: Optional[np.ndarray] = None
    ub: Optional[np.ndarray] = None
    x0: Optional[np.ndarray] = None
    y0: Optional[np.ndarray] = None
    shared: frozenset = frozenset()
    name: str = ""

    def field_len(self, f: str) -> int:
        n, c, p = self.nV, self.nC, self.nComp
        return {"Q": n * n, "g": n, "L": p * n, "R": p * n, "lbL": p, "ubL": p, "lbR": p, "ubR": p,
                "A": c * n, "lbA": c, "ubA": c, "lb": n, "ub": n, "x0": n, "y0": n + c + 2 * p}[f]

    def shared_mask(self) -> int:
        return sum(1 << i for i, f in enumerate(FIELDS) if f in self.shared)

    def present_mask(self) -> int:
        return sum(1 << i for i, f in enumerate(FIELDS) if getattr(self, f) is not None)

    def normalised(self) -> "LCQPBatch":
        """C-contiguous float64 views with shape (len,) if shared else (batch, len)."""
        kw = {}
        for f in FIELDS:
            a = getattr(self, f)
            if a is None:
                kw[f] = None
                continue
            a = np.ascontiguousarray(a, dtype=np.float64)
            ln = self.field_len(f)
            a = a.reshape(ln) if f in self.shared else a.reshape(self.batch, ln)
            kw[f] = a
        return dataclasses.replace(self, **kw)

    def instance(self, b: int) -> "LCQPBatch":
        """Instance ``b`` as a batch of one with everything 'shared'."""
        s = self.normalised()
        kw = {}
        for f in FIELDS:
            a = getattr(s, f)
            kw[f] = None if a is None else (a if f in self.shared else a[b]).copy()
        return dataclasses.replace(s, batch=1, shared=frozenset(f for f in FIELDS if kw[f] is not None), **kw)

    def slice(self, lo: int, hi: int) -> "LCQPBatch":
        s = self.normalised()
        kw = {}
        for f in FIELDS:
            a = getattr(s, f)
            kw[f] = None if a is None else (a if f in self.shared else a[lo:hi])
        return dataclasses.replace(s, batch=hi - lo, **kw)


def _single(name, nV, nC, nComp, **kw) -> LCQPBatch:
    arrs = {k: (None if v is None else np.asarray(v, dtype=np.float64)) for k, v in kw.items()}
    shared = frozenset(k for k, v in arrs.items() if v is not None)
    return LCQPBatch(nV=nV, nC=nC, nComp=nComp, batch=1, shared=shared, name=name, **arrs)


# ------------------------------------------------------------------------------------------------
# The reference's literal fixtures
# ------------------------------------------------------------------------------------------------
def warm_up(with_guess: bool = True) -> LCQPBatch:
    """examples/warm_up.cpp:32-42 (and test/RunUnitTests.cpp:506-512 when ``with_guess`` is False)."""
    return _single("warm_up", 2, 0, 1, Q=[2, 0, 0, 2], g=[-2, -2], L=[1, 0], R=[0, 1],
                   x0=[1, 1] if with_guess else None, y0=[0, 0, 0, 0] if with_guess else None)


def warm_up_w_A() -> LCQPBatch:
    """test/examples/warm_up_w_A.cpp:32-38."""
    return _single("warm_up_w_A", 2, 1, 1, Q=[2, 0, 0, 2], g=[-2, -2], L=[1, 0], R=[0, 1],
                   A=[1, -1], lbA=[-0.5], ubA=[math.inf])


def warm_up_binary() -> LCQPBatch:
    """test/examples/warm_up_binary.cpp:32-42: 0<=x _|_ y>=0 and 0<=x _|_ 0.5-x>=0."""
    return _single("warm_up_binary", 2, 0, 2, Q=[2, 0, 0, 2], g=[-2, -2], L=[1, 0, 1, 0], R=[0, 1, -1, 0],
                   lbL=[0, 0], lbR=[0, -0.5], x0=[0, 0])


def warm_up_shifted() -> LCQPBatch:
    """test/warm_up_shifted.cpp:32-41 (lbL=lbR=1)."""
    return _single("warm_up_shifted", 2, 0, 1, Q=[2, 0, 0, 2], g=[-4, -4], L=[1, 0], R=[0, 1],
                   lbL=[1], lbR=[1], x0=[1, 1], y0=[0, 0, 0, 0])


def infeasible_qp() -> LCQPBatch:
    """test/RunUnitTests.cpp:464-472: lbA=0 > ubA=-1 -> SUBPROBLEM_SOLVER_ERROR."""
    return _single("infeasible_qp", 2, 1, 1, Q=[2, 0, 0, 2], g=[-2, -2], L=[1, 0], R=[0, 1],
                   A=[1, 0], lbA=[0], ubA=[-1])


def stationarity_fixture(kind: str) -> LCQPBatch:
    """Two variables pinned to the bi-active point x = (0, 0) by 0 <= x1 <= 0 _|_ 0 <= x2 <= 0, so that the multipliers
    of the pair are y = g and the classifier of /root/reference/src/LCQProblem.cpp:1412-1453 sees every sign pattern:
    S (y >= 0), M (one zero, one negative), C (both negative), W (opposite signs)."""
    g = {"S": [1.0, 1.0], "M": [0.0, -1.0], "C": [-1.0, -1.0], "W": [1.0, -1.0]}[kind]
    return _single("stationarity_" + kind, 2, 0, 1, Q=[2, 0, 0, 2], g=g, L=[1, 0], R=[0, 1],
                   lbL=[0], ubL=[0], lbR=[0], ubR=[0], x0=[0, 0])


def circle_shared(N: int = 100):
    """Shared operands of examples/OptimizeOnCircle.cpp:62-99."""
    nV, nC, nComp = 2 + 2 * N, N + 1, N
    Q = np.zeros((nV, nV))
    Q[0, 0] = Q[1, 1] = 17.0
    Q[0, 1] = Q[1, 0] = -15.0
    for i in range(2, nV):
        Q[i, i] = 5e-12
    L = np.zeros((nComp, nV))
    R = np.zeros((nComp, nV))
    A = np.zeros((nC, nV))
    for i in range(N):
        A[i, 0] = math.cos((2 * math.pi * i) / N)
        A[i, 1] = math.sin((2 * math.pi * i) / N)
        A[i, 2 + 2 * i] = 1.0
        A[N, 3 + 2 * i] = 1.0
        L[i, 2 + 2 * i] = 1.0
        R[i, 3 + 2 * i] = 1.0
    lbA = np.ones(nC)
    ubA = np.ones(nC)
    return nV, nC, nComp, Q, L, R, A, lbA, ubA


def circle_batch(batch: int = 1, N: int = 100, seed0: int = 20000) -> LCQPBatch:
    """Config C2: OptimizeOnCircle with shared Q/A/L/R/lbA/ubA and per-instance g, x0.

    Instance 0 is the shipped x_ref=(0.5,-0.6) (examples/OptimizeOnCircle.cpp:36); instance b>0 draws
    x_ref ~ U([-1,1]^2) from default_rng(seed0+b), rejected until ||x_ref||_2 <= 0.95.
    """
    nV, nC, nComp, Q, L, R, A, lbA, ubA = circle_shared(N)
    Qx = np.array([[17.0, -15.0], [-15.0, 17.0]])
    g = np.zeros((batch, nV))
    x0 = np.ones((batch, nV))
    for b in range(batch):
        if b == 0:
            xr = np.array([0.5, -0.6])
        else:
            rng = np.random.default_rng(seed0 + b)
            while True:
                xr = rng.uniform(-1.0, 1.0, size=2)
                if np.linalg.norm(xr) <= 0.95:
                    break
        # examples/OptimizeOnCircle.cpp:72-75: g = -(Qx * x_ref), plain row-times-vector sums
        g[b, 0] = -(Qx[0, 0] * xr[0] + Qx[0, 1] * xr[1])
        g[b, 1] = -(Qx[1, 0] * xr[0] + Qx[1, 1] * xr[1])
        x0[b, 0:2] = xr
    return LCQPBatch(nV=nV, nC=nC, nComp=nComp, batch=batch, Q=Q, g=g, L=L, R=R, A=A, lbA=lbA, ubA=ubA,
                     x0=x0, shared=frozenset(("Q", "L", "R", "A", "lbA", "ubA")), name=f"circle_N{N}")


def circle_batch_fast(batch: int, N: int = 100, seed0: int = 20000, lo: int = 0, hi: int | None = None) -> LCQPBatch:
    """Same family as ``circle_batch`` but vectorised (one generator, seed ``seed0``) for bench-sized
    batches; instance 0 is still the shipped one.  Used for throughput only; parity subsets use
    ``circle_batch`` (per-instance seeds).

    ``lo``/``hi``: build only instances [lo, hi) of the family of ``batch`` instances (a shard of a multi-GPU run: the
    reference points of the whole family are drawn -- 16 bytes per instance -- the 3.2 KB of g and x0 per instance only
    for the shard)."""
    nV, nC, nComp, Q, L, R, A, lbA, ubA = circle_shared(N)
    hi = batch if hi is None else hi
    rng = np.random.default_rng(seed0)
    xr = np.empty((batch, 2))
    filled = 0
    while filled < batch:
        cand = rng.uniform(-1.0, 1.0, size=(2 * (batch - filled) + 16, 2))
        cand = cand[np.linalg.norm(cand, axis=1) <= 0.95][: batch - filled]
        xr[filled:filled + len(cand)] = cand
        filled += len(cand)
    xr[0] = (0.5, -0.6)
    xr = xr[lo:hi]
    nb = hi - lo
    g = np.zeros((nb, nV))
    g[:, 0] = -(17.0 * xr[:, 0] + -15.0 * xr[:, 1])
    g[:, 1] = -(-15.0 * xr[:, 0] + 17.0 * xr[:, 1])
    x0 = np.ones((nb, nV))
    x0[:, 0:2] = xr
    return LCQPBatch(nV=nV, nC=nC, nComp=nComp, batch=nb, Q=Q, g=g, L=L, R=R, A=A, lbA=lbA, ubA=ubA,
                     x0=x0, shared=frozenset(("Q", "L", "R", "A", "lbA", "ubA")), name=f"circle_N{N}")


def dense_random_batch(batch: int, n: int = 64, nComp: int = 32, nC: int = 16, seed0: int = 50000, seed_lo: int = 0) -> LCQPBatch:
    """Config C5: per-instance dense LCQPs, instance b from default_rng(seed0+b).

    L=[I 0], R=[0 I]; Q = M'M/n + 0.1 I; A ~ N(0,1)/8; a feasible complementary x* (one of each pair 0,
    the other U(0,1)); lbA = A x* - 0.1 - U(0,1), ubA = A x* + 0.1 + U(0,1); g ~ N(0,1).
    """
    assert 2 * nComp <= n
    if seed_lo:   # instances [seed_lo, batch) of the family only (a shard)
        seed0, batch = seed0 + seed_lo, batch - seed_lo
    Q = np.empty((batch, n, n))
    A = np.empty((batch, nC, n))
    g = np.empty((batch, n))
    lbA = np.empty((batch, nC))
    ubA = np.empty((batch, nC))
    L = np.zeros((nComp, n))
    R = np.zeros((nComp, n))
    for i in range(nComp):
        L[i, i] = 1.0
        R[i, nComp + i] = 1.0
    for b in range(batch):
        rng = np.random.default_rng(seed0 + b)
        M = rng.standard_normal((n, n))
        Qb = M.T @ M / n + 0.1 * np.eye(n)
        Q[b] = 0.5 * (Qb + Qb.T)
        A[b] = rng.standard_normal((nC, n)) / 8.0
        xs = rng.uniform(0.0, 1.0, size=n)
        pick = rng.integers(0, 2, size=nComp)
        for i in range(nComp):
            if pick[i]:
                xs[i] = 0.0
            else:
                xs[nComp + i] = 0.0
        Ax = A[b] @ xs
        lbA[b] = Ax - 0.1 - rng.uniform(0.0, 1.0, size=nC)
        ubA[b] = Ax + 0.1 + rng.uniform(0.0, 1.0, size=nC)
        g[b] = rng.standard_normal(n)
    return LCQPBatch(nV=n, nC=nC, nComp=nComp, batch=batch, Q=Q, g=g, L=L, R=R, A=A, lbA=lbA, ubA=ubA,
                     shared=frozenset(("L", "R")), name=f"dense_n{n}")


def example_data_batch(data: dict, batch: int = 1, seed0: int = 30000, perturb_ub: bool = False) -> LCQPBatch:
    """Config C3: examples/example_data (nV=151, nC=50, nComp=100) replicated with
    perturbed g and constraint bounds lbA=ubA (each entry scaled by 1 + 0.05 N(0,1)); instance 0 unperturbed.
    ``data`` maps the file stems (Q,g,L,R,lbL,ubL,lbR,ubR,A,lbA,ubA,lb,ub,x0) to arrays (tests/golden/example_data.npz).

    ``perturb_ub`` also scales the finite box bounds ub: that breaks the problem -- the reference (qpOASES) runs every
    such instance to MAX_PENALTY_PARAMETER_REACHED -- and is kept as the "terminal failure" family of the tests; with
    the box bounds as shipped the reference solves every instance of the family."""
    nV, nC, nComp = 151, 50, 100
    g = np.tile(np.asarray(data["g"], dtype=np.float64), (batch, 1))
    lbA = np.tile(np.asarray(data["lbA"], dtype=np.float64), (batch, 1))
    ub = np.tile(np.asarray(data["ub"], dtype=np.float64), (batch, 1))
    for b in range(1, batch):
        rng = np.random.default_rng(seed0 + b)
        g[b] *= 1.0 + 0.05 * rng.standard_normal(nV)
        lbA[b] *= 1.0 + 0.05 * rng.standard_normal(nC)
        fin = np.isfinite(ub[b])
        du = 1.0 + 0.05 * rng.standard_normal(int(fin.sum()))   # (drawn either way: the same g / lbA in both families)
        if perturb_ub:
            ub[b, fin] *= du
    kw = {k: np.asarray(data[k], dtype=np.float64) for k in ("Q", "L", "R", "lbL", "ubL", "lbR", "ubR", "A", "lb", "x0")}
    return LCQPBatch(nV=nV, nC=nC, nComp=nComp, batch=batch, g=g, lbA=lbA, ubA=lbA.copy(), ub=ub,
                     shared=frozenset(kw.keys()), name="example_data", **kw)


# ------------------------------------------------------------------------------------------------
# Config C4: sparse LCQPs that share their sparsity pattern, in CSC form
# ------------------------------------------------------------------------------------------------
@dataclasses.dataclass
class SparseLCQPBatch:
    """A batch of LCQPs in the layout of LCQProblem::loadLCQP(const csc* ...) (/root/reference/src/LCQProblem.cpp:312-387):
    each matrix is (colptr int32[nV+1], rowidx int32[nnz], values) with values of shape (nnz,) when shared by the
    batch, else (batch, nnz).  The patterns are shared by the whole batch."""

    nV: int
    nC: int
    nComp: int
    batch: int
    Q: tuple
    g: np.ndarray
    L: tuple
    R: tuple
    A: Optional[tuple] = None
    lbA: Optional[np.ndarray] = None
    ubA: Optional[np.ndarray] = None
    lbL: Optional[np.ndarray] = None
    ubL: Optional[np.ndarray] = None
    lbR: Optional[np.ndarray] = None
    ubR: Optional[np.ndarray] = None
    x0: Optional[np.ndarray] = None
    y0: Optional[np.ndarray] = None
    shared: frozenset = frozenset()
    name: str = ""

    def _dense(self, t, rows, b):
        p, i, x = t
        xv = x if x.ndim == 1 else x[b]
        M = np.zeros((rows, self.nV))
        for c in range(self.nV):
            M[i[p[c]:p[c + 1]], c] = xv[p[c]:p[c + 1]]
        return M

    def to_dense(self, lo: int, hi: int) -> LCQPBatch:
        """Instances [lo, hi) as a dense LCQPBatch (for the reference's dense front door)."""
        nb = hi - lo
        def vec(a, f):
            if a is None:
                return None
            return a if f in self.shared else a[lo:hi]
        mats = {}
        for f, rows in (("Q", self.nV), ("L", self.nComp), ("R", self.nComp), ("A", self.nC)):
            t = getattr(self, f)
            if t is None:
                mats[f] = None
            elif f in self.shared:
                mats[f] = self._dense(t, rows, 0)
            else:
                mats[f] = np.stack([self._dense(t, rows, b) for b in range(lo, hi)])
        return LCQPBatch(nV=self.nV, nC=self.nC, nComp=self.nComp, batch=nb, Q=mats["Q"], g=vec(self.g, "g"), L=mats["L"], R=mats["R"],
                         A=mats["A"], lbA=vec(self.lbA, "lbA"), ubA=vec(self.ubA, "ubA"), lbL=vec(self.lbL, "lbL"), ubL=vec(self.ubL, "ubL"),
                         lbR=vec(self.lbR, "lbR"), ubR=vec(self.ubR, "ubR"), x0=vec(self.x0, "x0"), y0=vec(self.y0, "y0"),
                         shared=self.shared, name=self.name)

    def slice(self, lo: int, hi: int) -> "SparseLCQPBatch":
        kw = {}
        for f in ("Q", "L", "R", "A"):
            t = getattr(self, f)
            kw[f] = None if t is None else (t if (f in self.shared) else (t[0], t[1], t[2][lo:hi]))
        for f in ("g", "lbA", "ubA", "lbL", "ubL", "lbR", "ubR", "x0", "y0"):
            a = getattr(self, f)
            kw[f] = None if a is None else (a if f in self.shared else a[lo:hi])
        return dataclasses.replace(self, batch=hi - lo, **kw)


def sparse_banded_batch(batch: int, n: int = 1000, nComp: int = 500, nC: int = 300, seed0: int = 40000, lo: int = 0) -> SparseLCQPBatch:
    """Config C4: banded Q (off-diagonals (i, i+1..i+3) ~ 0.1 N(0,1), diagonal = sum of |off-diagonals| of the row
    + 0.5 + U(0,1): strictly diagonally dominant), A with 5 non-zeros per row ~ N(0,1) in a +-8 column window around
    r n / nC, L picks x_2i and R picks x_2i+1; a feasible complementary x* (one of each pair 0, the other U(0,1)),
    lbA = A x* - 0.1 - U(0,1), ubA = A x* + 0.1 + U(0,1), g ~ N(0,1).  The pattern comes from default_rng(seed0) and
    is shared by the batch, the values of instance b from default_rng(seed0 + 1 + b).  Instances [lo, lo + batch)."""
    assert 2 * nComp <= n
    prng = np.random.default_rng(seed0)
    # pattern of A: 5 distinct columns per row
    arows, acols = [], []
    for r in range(nC):
        c = (r * n) // nC
        window = np.arange(max(0, c - 8), min(n, c + 9))
        cols = np.sort(prng.choice(window, size=5, replace=False))
        arows += [r] * 5
        acols += cols.tolist()
    arows, acols = np.array(arows), np.array(acols)
    order = np.lexsort((arows, acols))          # CSC order: by column, then row
    Ap = np.zeros(n + 1, dtype=np.int32)
    np.add.at(Ap, acols + 1, 1)
    Ap = np.cumsum(Ap).astype(np.int32)
    Ai = arows[order].astype(np.int32)
    # pattern of Q: |i - j| <= 3
    qrows, qcols = [], []
    for j in range(n):
        for i in range(max(0, j - 3), min(n, j + 4)):
            qrows.append(i); qcols.append(j)
    qrows, qcols = np.array(qrows), np.array(qcols)
    Qp = np.zeros(n + 1, dtype=np.int32)
    np.add.at(Qp, qcols + 1, 1)
    Qp = np.cumsum(Qp).astype(np.int32)
    Qi = qrows.astype(np.int32)
    Lp = np.zeros(n + 1, dtype=np.int32); Rp = np.zeros(n + 1, dtype=np.int32)
    for i in range(nComp):
        Lp[2 * i + 1:] += 1
        Rp[2 * i + 2:] += 1
    Li = np.arange(nComp, dtype=np.int32); Ri = np.arange(nComp, dtype=np.int32)
    Lx = np.ones(nComp); Rx = np.ones(nComp)

    Qx = np.empty((batch, len(Qi))); Ax = np.empty((batch, len(Ai)))
    g = np.empty((batch, n)); lbA = np.empty((batch, nC)); ubA = np.empty((batch, nC))
    for k in range(batch):
        rng = np.random.default_rng(seed0 + 1 + lo + k)
        off = 0.1 * rng.standard_normal((n, 3))               # off[i, d-1] = Q[i, i+d]
        for d in range(1, 4):
            off[n - d:, d - 1] = 0.0
        rowabs = np.abs(off).sum(axis=1)
        for d in range(1, 4):
            rowabs[d:] += np.abs(off[:n - d, d - 1])
        diag = rowabs + 0.5 + rng.uniform(0.0, 1.0, size=n)
        dd = qrows - qcols                                     # row - col
        qv = np.where(dd == 0, diag[qcols], 0.0)
        for d in range(1, 4):
            up = dd == -d                                      # (i, i + d): row = col - d
            qv[up] = off[qrows[up], d - 1]
            lw = dd == d                                       # (i + d, i): mirrors (col, col + d)
            qv[lw] = off[qcols[lw], d - 1]
        Qx[k] = qv
        av = rng.standard_normal((nC, 5)).reshape(-1)          # row-major (r, slot)
        Ax[k] = av[order]
        xs = rng.uniform(0.0, 1.0, size=n)
        pick = rng.integers(0, 2, size=nComp)
        xs[2 * np.arange(nComp) + (1 - pick)] = 0.0            # pick = 1: x_2i = 0 ... one of each pair
        Axs = np.zeros(nC)
        np.add.at(Axs, arows, av * xs[acols])
        lbA[k] = Axs - 0.1 - rng.uniform(0.0, 1.0, size=nC)
        ubA[k] = Axs + 0.1 + rng.uniform(0.0, 1.0, size=nC)
        g[k] = rng.standard_normal(n)
    return SparseLCQPBatch(nV=n, nC=nC, nComp=nComp, batch=batch, Q=(Qp, Qi, Qx), g=g, L=(Lp, Li, Lx), R=(Rp, Ri, Rx),
                           A=(Ap, Ai, Ax), lbA=lbA, ubA=ubA, shared=frozenset(("L", "R")), name=f"sparse_n{n}")

"""oracle/pas_oracle.py -- TEST INFRASTRUCTURE ONLY (never imported by the product).

A numpy restatement of the path the product replaces: LCQProblem::runSolver (/root/reference/src/LCQProblem.cpp:444-560
and helpers :885-1034, :1105-1482) over qpOASES' online active-set strategy as LCQPow drives it
(/root/reference/src/SubsolverQPOASES.cpp:134-169):
    init      QProblem::solveInitialQP             external/qpOASES/src/QProblem.cpp:1301-1471 (auxiliary QP :2199-2344, :2668-2813)
    hotstart  QProblem::hotstart (far bounds)      QProblem.cpp:446-640, :5498-5552
    loop      QProblem::solveQP                    QProblem.cpp:1477-1747
    ratio     performStep / performRatioTest       QProblem.cpp:4981-5278, QProblemB.cpp:2065-2107
    LI        addConstraint_ensureLI               QProblem.cpp:3117-3300
    ramping   performRamping                       QProblem.cpp:5416-5492
    drift     performDriftCorrection               QProblem.cpp:5559-5652
    stop      getRelativeHomotopyLength            QProblem.cpp:5372-5410, QProblemB.cpp:2113-2157
The decisions (which row blocks, when the homotopy counts as finished, which row leaves when an addition is linearly
dependent, what a zero step does) are qpOASES'; the linear algebra is plain dense numpy in range space with the rows
that are equalities eliminated (they stay active with free multipliers -- qpOASES flips them between their two
sides with zero-length steps instead, which changes its iteration count but not the path of x).

Pinned: tests/test_oracle_parity.py checks it against the committed outputs of the UNMODIFIED reference
(tests/golden/*.npz: ReturnValue, stationarity type, outer and total iteration counts, x to 1e-6) and, where
oracle/_ref is present, live against the reference on fresh seeds.  Slow (dense numpy, a second per LCQP at n = 202):
for small cases only.
"""
import numpy as np

EPS = 2.221e-16
INFTY = 1e20
TERM_TOL = 5.0e6 * EPS
BOUND_TOL = 1.0e6 * EPS
BOUND_RELAX = 1.0e4
EPS_NUM = -1.0e3 * EPS
EPS_DEN = 1.0e3 * EPS
MAX_DUAL_JUMP = 1.0e8
RAMP0, RAMP1 = 0.5, 1.0
FAR0, FAR_GROW = 1.0e6, 1.0e3
INACTIVE, LOWER, UPPER = 0, 1, -1


class QP:
    def __init__(self, H, Afull, nV_q, nC_q, rib_full, eq_mask, verbose=0, li_tol=1e-10):
        self.n = n = H.shape[0]
        self.H = H
        self.Hinv = np.linalg.inv(H)
        E = self.Eidx = np.nonzero(eq_mask)[0]
        I = self.Iidx = np.nonzero(~eq_mask)[0]
        self.AE = Afull[E]; self.A = Afull[I]
        self.mE = len(E); self.m = len(I)
        Gf = Afull @ self.Hinv
        Tf = Gf @ Afull.T; Tf = 0.5 * (Tf + Tf.T)
        self.GE = Gf[E]; self.GI = Gf[I]
        if self.mE:
            self.TEEinv = np.linalg.inv(Tf[np.ix_(E, E)])
            self.K = Tf[np.ix_(I, E)] @ self.TEEinv
            self.T = Tf[np.ix_(I, I)] - self.K @ Tf[np.ix_(E, I)]
        else:
            self.TEEinv = np.zeros((0, 0)); self.K = np.zeros((self.m, 0)); self.T = Tf[np.ix_(I, I)]
        self.T = 0.5 * (self.T + self.T.T)
        self.TEI = Tf[np.ix_(E, I)]
        self.nVq, self.nCq = nV_q, nC_q
        self.rib = [rib_full[i] for i in I]
        self.verbose = verbose; self.li_tol = li_tol
        self.rampOffset = 0
        self.W = []; self.status = np.zeros(self.m, dtype=int); self.Sinv = np.zeros((0, 0))
        self.stamp = np.arange(self.m); self.stamp_next = self.m
        self.nwsr = 0; self.nsolve = 0; self._dc_dirty = False
        self.cap = max(1, min(self.m, self.n - self.mE))   # a linearly independent working set has at most n rows

    def Gt(self, v, dbE):   # c-space image of a gradient change v and an equality-bound change dbE
        return self.GI @ v - self.K @ (self.GE @ v + dbE)

    # ---- working set ----------------------------------------------------------------------------
    def pivot(self, k):
        if not self.W: return self.T[k, k], np.zeros(0)
        u = self.Sinv @ self.T[self.W, k]
        return self.T[k, k] - self.T[self.W, k] @ u, u

    def add_row(self, k, st, u, p):
        nw = len(self.W)
        S = np.zeros((nw + 1, nw + 1))
        S[:nw, :nw] = self.Sinv + np.outer(u, u) / p
        S[:nw, nw] = -u / p; S[nw, :nw] = -u / p; S[nw, nw] = 1.0 / p
        self.Sinv = S; self.W.append(k); self.status[k] = st
        self.stamp[k] = self.stamp_next; self.stamp_next += 1

    def remove_row(self, k):
        a = self.W.index(k)
        s = self.Sinv[:, a].copy()
        S = self.Sinv - np.outer(s, s) / s[a]
        keep = [i for i in range(len(self.W)) if i != a]
        self.Sinv = S[np.ix_(keep, keep)]
        self.W.pop(a); self.status[k] = INACTIVE
        self.stamp[k] = self.stamp_next; self.stamp_next += 1

    # ---- ramp values (QProblem.cpp:5416-5492, :5498-5552) ----------------------------------------
    def ramp_vals(self, i):
        nV, nC = self.nVq, self.nCq
        nRamp = nV + nC + nC + nV
        isb, li = self.rib[i]; off = self.rampOffset
        if isb: tP = ((li + off) % nRamp) / (nRamp - 1); tD = ((nV + nC + nC + li + off) % nRamp) / (nRamp - 1)
        else: tP = ((nV + li + off) % nRamp) / (nRamp - 1); tD = ((nV + nC + li + off) % nRamp) / (nRamp - 1)
        return (1 - tP) * RAMP0 + tP * RAMP1, (1 - tD) * RAMP0 + tD * RAMP1

    def far_bounds(self, far, l_new, u_new):
        nV, nC = self.nVq, self.nCq; nRamp = nV + nC
        lf = np.empty(self.m); uf = np.empty(self.m)
        for i in range(self.m):
            isb, li = self.rib[i]
            t = (((li if isb else nV + li) + self.rampOffset) % nRamp) / (nRamp - 1)
            rv = far * (1.0 + (1.0 - t) * RAMP0 + t * RAMP1)
            lf[i] = max(-rv, l_new[i]); uf[i] = min(rv, u_new[i])
        return lf, uf

    # ---- the primal: xq is the QP's own iterate; rebase() brings it (and gq, y0s) to the current homotopy point
    def yE_now(self, g):
        W = self.W
        return self.TEEinv @ (self.bE + self.GE @ g - (self.TEI[:, W] @ self.y[W] if W else 0.0))

    def g_now(self):
        return self.g_new - self.phi * self.dg0

    def rebase(self):
        g = self.g_now()
        yE = self.yE_now(g)
        self.xq = self.xq + self.Hinv @ (self.A.T @ (self.y - self.y_base) + self.AE.T @ (yE - self.yE_base) - (g - self.gq))
        self.gq = g; self.y_base = self.y.copy(); self.yE_base = yE
        return g

    def ramping(self):
        self.rebase()
        for i in range(self.m):
            rP, rD = self.ramp_vals(i)
            sca = max(abs(self.z[i]), 1.0); st = self.status[i]
            if st != LOWER: self.l[i] = self.z[i] - sca * rP
            if st != UPPER: self.u[i] = self.z[i] + sca * rP
            if st == LOWER: self.l[i] = self.z[i]; self.y[i] = +rD
            if st == UPPER: self.u[i] = self.z[i]; self.y[i] = -rD
            if st == INACTIVE: self.y[i] = 0.0
        self.gq = -self.H @ self.xq + self.A.T @ self.y + self.AE.T @ self.yE_base
        self.y_base = self.y.copy()
        self.set_target(self.g_new)
        W = self.W
        self.c = (self.T[:, W] @ self.y[W] if W else np.zeros(self.m)) - self.z
        self.rampOffset += 1

    def set_target(self, g_new):
        self.g_new = g_new
        self.dg0 = g_new - self.gq
        self.phi = 1.0
        self.len0 = np.max(np.abs(self.dg0) / np.maximum(np.abs(g_new), 1.0))

    # ---- init (QProblem.cpp:1301-1471) ------------------------------------------------------------
    def init(self, g, lfull, ufull, x0, y0full):
        n, m = self.n, self.m
        self.xq = np.zeros(n) if x0 is None else x0.copy()
        self.z = self.A @ self.xq
        l = lfull[self.Iidx]; u = ufull[self.Iidx]
        y0 = None if y0full is None else y0full[self.Iidx]
        self.bE = self.AE @ self.xq
        self.yE_base = np.zeros(self.mE) if y0full is None else y0full[self.Eidx].copy()
        self.y = np.zeros(m) if y0 is None else y0.copy()
        aux = np.zeros(m, dtype=int)
        for i in range(m):
            if y0 is not None: aux[i] = LOWER if y0[i] > EPS else (UPPER if y0[i] < -EPS else INACTIVE)
            elif x0 is not None:
                if self.z[i] - l[i] <= BOUND_TOL: aux[i] = LOWER
                elif u[i] - self.z[i] <= BOUND_TOL: aux[i] = UPPER
            else: aux[i] = LOWER if self.rib[i][0] else INACTIVE
        order = [i for i in range(m) if self.rib[i][0]] + [i for i in range(m) if not self.rib[i][0]]
        for i in order:
            if aux[i] != INACTIVE:
                p, uu = self.pivot(i)
                if p > self.li_tol * self.T[i, i] and len(self.W) < self.cap: self.add_row(i, aux[i], uu, p)
        self.l = np.empty(m); self.u = np.empty(m)
        for i in range(m):
            st = self.status[i]
            if st == INACTIVE:
                self.l[i] = self.z[i] if aux[i] == LOWER else self.z[i] - BOUND_RELAX
                self.u[i] = self.z[i] if aux[i] == UPPER else self.z[i] + BOUND_RELAX
            elif st == LOWER: self.l[i] = self.z[i]; self.u[i] = self.z[i] + BOUND_RELAX
            else: self.u[i] = self.z[i]; self.l[i] = self.z[i] - BOUND_RELAX
        self.gq = -self.H @ self.xq + self.A.T @ self.y + self.AE.T @ self.yE_base
        self.y_base = self.y.copy()
        self.g_new = g
        self.set_target(g)
        self.ramping()
        return self.hotstart(g, lfull, ufull)

    # ---- hotstart with far bounds (QProblem.cpp:446-640) -----------------------------------------
    def hotstart(self, g_new, lfull, ufull):
        if np.any(lfull > ufull + EPS): return -1
        l_new = lfull[self.Iidx]; u_new = ufull[self.Iidx]
        self.bE_new = lfull[self.Eidx]
        far = FAR0
        allb = np.concatenate([lfull, ufull]); fin = np.abs(allb[np.abs(allb) < INFTY])
        if fin.size: far = max(far, fin.max())
        lf, uf = self.far_bounds(far, l_new, u_new)
        self.nwsr = 0
        self.set_target(g_new)
        rc = 0
        while True:
            rc = self.solve(lf, uf)
            far *= FAR_GROW
            if rc == -2:
                if far >= INFTY: break
                lf, uf = self.far_bounds(far, l_new, u_new)
            elif rc == 0:
                tol = far / FAR_GROW * BOUND_TOL
                nact = int(np.sum((lf > l_new) & (np.abs(lf - self.z) < tol)) + np.sum((uf < u_new) & (np.abs(uf - self.z) < tol)))
                if nact == 0: break
                if far >= INFTY: rc = -3; break
                lf, uf = self.far_bounds(far, l_new, u_new)
            else: break
            self.rampOffset += 1
        if rc == 0: self.finish()
        return rc

    # ---- the homotopy loop (QProblem.cpp:1477-1747) -- only vectors over the rows ----------------
    def solve(self, l_new, u_new, max_iter=5000):
        m = self.m
        for it in range(max_iter):
            dbE = self.bE_new - self.bE
            # c_new - c :  the remaining gradient change phi*dg0 and the remaining equality-bound change
            dc = self.Gt(self.phi * self.dg0, dbE) if it == 0 or self._dc_dirty else dc * (1.0 - tau)
            self._dc_dirty = False
            dl = l_new - self.l; du = u_new - self.u
            W = self.W
            if W:
                db = np.where(self.status[W] == LOWER, dl[W], du[W])
                dyW = self.Sinv @ (db + dc[W]); self.nsolve += 1
                dz = self.T[:, W] @ dyW - dc
                dz[W] = db
            else:
                dyW = np.zeros(0); dz = -dc
            tau = 1.0; bc = -1; bcst = None
            posW = {k: a for a, k in enumerate(W)}
            act = sorted(W, key=lambda i: self.stamp[i])
            inact = sorted([i for i in range(m) if self.status[i] == INACTIVE], key=lambda i: self.stamp[i])
            isb = lambda i: self.rib[i][0]
            def test(num, den, i, st_new):
                nonlocal tau, bc, bcst
                if den >= EPS_DEN and num >= EPS_NUM and num < tau * den:
                    tau = num / den; bc = i; bcst = st_new
            for grp in (False, True):
                for i in act:
                    if isb(i) != grp: continue
                    num, den = self.y[i], -dyW[posW[i]]
                    if self.status[i] == UPPER: num, den = -num, -den
                    test(num, den, i, INACTIVE)
            for grp in (False, True):
                for i in inact:
                    if isb(i) == grp: test(max(self.z[i] - self.l[i], 0.0), dl[i] - dz[i], i, LOWER)
                for i in inact:
                    if isb(i) == grp: test(max(self.u[i] - self.z[i], 0.0), dz[i] - du[i], i, UPPER)
            if tau > 1e-25:
                if W: self.y[W] += tau * dyW
                self.z += tau * dz; self.c += tau * dc
                self.l = self.l + tau * dl; self.u = self.u + tau * du; self.bE = self.bE + tau * dbE
                self.phi *= (1.0 - tau)
            else:
                tau = 0.0
            def rel(new, cur):
                return np.max(np.abs(new - cur) / np.maximum(np.abs(new), 1.0)) if new.size else 0.0
            hl = max(self.phi * self.len0, rel(l_new, self.l), rel(u_new, self.u), rel(self.bE_new, self.bE))
            if self.verbose: print("   it %d tau %.3e bc %s st %s hl %.3e nW %d" % (self.nwsr, tau, bc, bcst, hl, len(W)))
            if hl <= TERM_TOL: return 0
            self.nwsr += 1
            if bc >= 0:
                if bcst == INACTIVE:
                    self.remove_row(bc); self.y[bc] = 0.0
                else:
                    rc = self.add_with_li(bc, bcst)
                    if rc != 0: return rc
                if tau <= EPS:
                    self.ramping(); self._dc_dirty = True
                else:
                    self.drift()
            else:
                self.drift()
        return -4

    def drift(self):
        for i in range(self.m):
            st = self.status[i]; yo = self.y[i]
            if st == LOWER: self.l[i] = self.z[i]; self.u[i] = max(self.u[i], self.z[i]); self.y[i] = max(yo, 0.0)
            elif st == UPPER: self.u[i] = self.z[i]; self.l[i] = min(self.l[i], self.z[i]); self.y[i] = min(yo, 0.0)
            else: self.l[i] = min(self.l[i], self.z[i]); self.u[i] = max(self.u[i], self.z[i]); self.y[i] = 0.0
            if self.y[i] != yo:
                d = self.y[i] - yo
                self.c = self.c + self.T[:, i] * d     # x stays, the gradient absorbs the change
                self.gq = self.gq  # (the n-dimensional gradient is re-derived at the next rebase; d is round-off)
                self._dc_dirty = True

    def add_with_li(self, k, st):
        p, u = self.pivot(k)
        if p > self.li_tol * self.T[k, k] and len(self.W) < self.cap:
            self.add_row(k, st, u, p); return 0
        sgn = 1.0 if st == LOWER else -1.0
        xi = sgn * u
        ymin = MAX_DUAL_JUMP; jmin = -1
        order = sorted(range(len(self.W)), key=lambda a: self.stamp[self.W[a]])
        for grp in (False, True):
            for a in order:
                i = self.W[a]
                if self.rib[i][0] != grp: continue
                num, den = self.y[i], xi[a]
                if self.status[i] == UPPER: num, den = -num, -den
                if den >= EPS_DEN and num >= EPS_NUM and num < ymin * den: ymin = num / den; jmin = a
        if jmin < 0: return -2
        for a, i in enumerate(self.W): self.y[i] -= ymin * xi[a]
        self.y[k] = ymin if st == LOWER else -ymin
        rem = self.W[jmin]
        self.remove_row(rem); self.y[rem] = 0.0
        p, u = self.pivot(k)
        self.add_row(k, st, u, p)
        return 0

    def finish(self):
        """End of a QP: bring xq to the end point, polish primal feasibility of the active rows in range space,
        recompute the row activities (drift)."""
        self.rebase()
        W = self.W
        self.npolish = getattr(self, 'npolish', 0)
        for _ in range(8):
            r2E = self.bE - self.AE @ self.xq
            rn = np.abs(r2E).max() if self.mE else 0.0
            if W:
                b = np.where(self.status[W] == LOWER, self.l[W], self.u[W])
                r2W = b - self.A[W] @ self.xq
                rn = max(rn, np.abs(r2W).max())
            if rn <= 1e-15: break
            self.npolish += 1
            if W:
                dyW = self.Sinv @ (r2W - self.K[W] @ r2E)
                dyE = self.TEEinv @ (r2E - self.TEI[:, W] @ dyW)
                self.xq = self.xq + self.Hinv @ (self.A[W].T @ dyW + self.AE.T @ dyE)
                self.y[W] += dyW
            else:
                dyE = self.TEEinv @ r2E
                self.xq = self.xq + self.Hinv @ (self.AE.T @ dyE)
            self.yE_base = self.yE_base + dyE
        for i in range(self.m):   # the polish moves multipliers by round-off: signs as after a drift correction
            st = self.status[i]
            self.y[i] = max(self.y[i], 0.0) if st == LOWER else (min(self.y[i], 0.0) if st == UPPER else 0.0)
        self.y_base = self.y.copy()
        self.z = self.A @ self.xq
        if W: self.c = self.T[:, W] @ self.y[W] - self.z
        else: self.c = -self.z

    def solution(self):
        yfull = np.zeros(self.mE + self.m)
        yfull[self.Eidx] = self.yE_base; yfull[self.Iidx] = self.y
        return self.xq.copy(), yfull


def splitmix_draw(seed, inst, it, i):
    M = (1 << 64) - 1
    z = (seed * 0x9E3779B97F4A7C15 + inst * 0xBF58476D1CE4E5B9 + ((it << 32) | i)) & M
    z = (z + 0x9E3779B97F4A7C15) & M
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M
    z = z ^ (z >> 31)
    return int(z % 3) - 1


def lcqp_solve(pb, statTol=None, perturb=False, verbose=0, qp_verbose=0, maxIter=1000, maxRho=1e8, li_tol=1e-10, elim=True, seed=1, inst=0):
    n, nC, nComp = pb.nV, pb.nC, pb.nComp
    Q = pb.Q.reshape(n, n); g = pb.g.reshape(n)
    L = pb.L.reshape(nComp, n); R = pb.R.reshape(nComp, n)
    A = pb.A.reshape(nC, n) if pb.A is not None else np.zeros((0, n))
    inf = np.inf
    lbA = pb.lbA.reshape(nC) if pb.lbA is not None else np.full(nC, -inf)
    ubA = pb.ubA.reshape(nC) if pb.ubA is not None else np.full(nC, inf)
    lbL = pb.lbL.reshape(nComp) if pb.lbL is not None else np.zeros(nComp)
    ubL = pb.ubL.reshape(nComp) if pb.ubL is not None else np.full(nComp, inf)
    lbR = pb.lbR.reshape(nComp) if pb.lbR is not None else np.zeros(nComp)
    ubR = pb.ubR.reshape(nComp) if pb.ubR is not None else np.full(nComp, inf)
    has_box = pb.lb is not None or pb.ub is not None
    rows = [A, L, R]; lo = [lbA, lbL, lbR]; up = [ubA, ubL, ubR]
    mA = nC + 2 * nComp
    rib = [(False, i) for i in range(mA)]
    if has_box:
        rows.append(np.eye(n))
        lo.append(pb.lb.reshape(n) if pb.lb is not None else np.full(n, -inf))
        up.append(pb.ub.reshape(n) if pb.ub is not None else np.full(n, inf))
        rib += [(True, i) for i in range(n)]
    Ahat = np.vstack(rows); l = np.concatenate(lo); u = np.concatenate(up)
    eq = (l == u) & np.isfinite(l) if elim else np.zeros(len(l), dtype=bool)
    qp = QP(Q, Ahat, n, mA, rib, eq, verbose=qp_verbose, li_tol=li_tol)
    C = L.T @ R + R.T @ L
    statTol = 1e6 * EPS if statTol is None else statTol
    compTol = 1e3 * EPS
    rho = 0.01; alpha = 1.0
    xk = pb.x0.reshape(n).copy() if pb.x0 is not None else np.zeros(n)
    y0 = None
    if pb.y0 is not None:
        yy = pb.y0.reshape(-1)
        y0 = np.concatenate([yy[n:n + mA], yy[:n] if has_box else np.zeros(0)])
    gphi = -(R.T @ lbL + L.T @ lbR); phic = lbL @ lbR
    gt = g.copy()
    phi = lambda x: phic + gphi @ x + 0.5 * x @ (C @ x)
    outer = 0; total = 0; sub = 0; hist = []
    rc = qp.init(g.copy(), l, u, xk, y0)   # LCQProblem.ipp:138-142: xk is zero-filled without a guess, never NULL
    sub += qp.nwsr
    if rc != 0: return dict(ret=203, status=0, k=outer, i=total, sub=sub, x=xk, exit=rc, nsolve=qp.nsolve)
    xnew, y = qp.solution()
    pk = xnew - xk
    nd = 3; eta = 0.9
    while True:
        xk = xk + alpha * pk
        Qk = Q + rho * C
        stat = Qk @ xk + gt - Ahat.T @ y
        total += 1
        cur = phi(xk); fire = False
        if len(hist) < nd: hist.append(cur)
        else:
            if not (cur < compTol): fire = all(not (cur < eta * h) for h in hist)
            hist = hist[1:] + [cur]
        if fire: hist = []; rho *= 2; gt = g + rho * gphi; outer += 1
        if verbose: print("i=%d k=%d rho=%g stat=%.3e phi=%.3e sub=%d x12=%s" % (total, outer, rho, np.abs(stat).max(), cur, qp.nwsr, xk[:2]))
        if np.abs(stat).max() < statTol:
            if phi(xk) < compTol:
                # determineStationarityType on the penalised duals (LCQProblem.cpp:1412-1453, weak set :1456-1482)
                Lx, Rx = L @ xk, R @ xk
                s_ok = m_ok = True; status = None
                for i in range(nComp):
                    if Lx[i] <= compTol and Rx[i] <= compTol:
                        yl, yr = y[nC + i], y[nC + nComp + i]
                        prod, mn = yl * yr, min(yl, yr)
                        if mn < 0: s_ok = False
                        if abs(prod) >= compTol and mn <= 0:
                            if prod <= compTol: status = 1; break
                            m_ok = False
                if status is None: status = 4 if s_ok else (3 if m_ok else 2)
                yt = y.copy()   # transformDuals :1381-1409
                yt[nC:nC + nComp] -= rho * Rx; yt[nC + nComp:mA] -= rho * Lx
                return dict(ret=0, status=status, k=outer, i=total, sub=sub, x=xk, y=yt, rho=rho, nsolve=qp.nsolve)
            hist = []; rho *= 2; gt = g + rho * gphi; outer += 1
        if total > maxIter: return dict(ret=200, status=0, k=outer, i=total, sub=sub, x=xk, nsolve=qp.nsolve)
        if rho > maxRho: return dict(ret=201, status=0, k=outer, i=total, sub=sub, x=xk, nsolve=qp.nsolve)
        gk = rho * (C @ xk) + gt
        rc = qp.hotstart(gk, l, u)
        sub += qp.nwsr
        if rc != 0: return dict(ret=203, status=0, k=outer, i=total, sub=sub, x=xk, exit=rc, nsolve=qp.nsolve)
        xnew, y = qp.solution()
        pk = xnew - xk
        if perturb:
            for j in range(n): xk[j] += splitmix_draw(seed, inst, total, j) * EPS
        Qk = Q + rho * C
        qq = pk @ (Qk @ pk); lk = pk @ (Qk @ xk + gt)
        alpha = min(-lk / qq, 1.0) if (qq > 0 and lk < 0) else 1.0


def solve_batch(pb, perturb=0, **over):
    """Solve every instance of an lcqpow_b200.problems.LCQPBatch; returns dict of arrays (ret, status, iterOuter,
    iterTotal, subproblemIter, x)."""
    out = dict(ret=[], status=[], iterOuter=[], iterTotal=[], subproblemIter=[], x=[])
    for b in range(pb.batch):
        inst = pb.instance(b) if pb.batch > 1 else pb.normalised()
        r = lcqp_solve(inst, statTol=over.get("stationarityTolerance"), maxRho=over.get("maxPenaltyParameter", 1e8),
                       maxIter=over.get("maxIterations", 1000), perturb=bool(perturb), inst=b)
        out["ret"].append(r["ret"]); out["status"].append(r["status"]); out["iterOuter"].append(r["k"])
        out["iterTotal"].append(r["i"]); out["subproblemIter"].append(r["sub"]); out["x"].append(r["x"])
    return {k: np.array(v) for k, v in out.items()}

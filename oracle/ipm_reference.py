"""CPU oracle, part 2: an IPOPT-style primal-dual interior-point solver (numpy/scipy).

TEST INFRASTRUCTURE ONLY (see ascent_nlp.py header).  PARITY UNPINNED: IPOPT itself is
absent from this container and from /root/reference; this is a restatement of its
published algorithm (Waechter & Biegler, Math. Prog. 106 (2006), "IPOPT paper" below),
which the reference selects with ``m.options.SOLVER = 3`` (LO:26) and reaches through
``m.solve`` (LO:177).

It deliberately shares NO linear algebra with the CUDA product: the KKT system is
assembled as one general sparse matrix and factorised by SuperLU with no knowledge of
the stage structure (like IPOPT+MUMPS/MA27 in the reference), and the inertia
condition is replaced by the curvature test of Chiang & Zavala (2016) because a sparse
LU exposes no inertia.

Algorithm (section numbers of the IPOPT paper):
  * barrier problem, primal-dual Newton system, Sigma = Z/X ............... (3), (11), (13)
  * fraction-to-boundary rule, tau = max(tau_min, 1-mu) ..................... (15), (8)
  * filter line search with switching condition / Armijo .................. 2.3, (18)-(20)
  * second-order correction ................................................ 2.4
  * monotone mu update, kappa_mu = 0.2, theta_mu = 1.5, kappa_eps = 10 ...... (7)
  * scaled optimality error E_mu ........................................... (5), (6)
  * multiplier safeguard kappa_Sigma = 1e10 ................................ (16)
  * Hessian regularisation schedule delta_w ................................ 3.1 (IC)
  * feasibility restoration: a Gauss-Newton/Levenberg step on theta only, kept simple.
"""
from __future__ import annotations

import dataclasses
from typing import Optional

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla


@dataclasses.dataclass
class IPMOptions:
    tol: float = 1e-8
    max_iter: int = 3000
    mu_init: float = 0.1
    kappa_mu: float = 0.2
    theta_mu: float = 1.5
    kappa_eps: float = 10.0
    tau_min: float = 0.99
    kappa_sigma: float = 1e10
    s_max: float = 100.0
    gamma_theta: float = 1e-5
    gamma_phi: float = 1e-8
    delta: float = 1.0
    s_theta: float = 1.1
    s_phi: float = 2.3
    eta_phi: float = 1e-8
    max_soc: int = 4
    kappa_soc: float = 0.99
    bound_push: float = 1e-2
    bound_frac: float = 1e-2
    delta_w_first: float = 1e-4
    delta_w_min: float = 1e-20
    delta_w_max: float = 1e40
    kappa_w_minus: float = 1.0 / 3.0
    kappa_w_plus: float = 8.0
    kappa_w_plus_first: float = 100.0
    delta_c: float = 1e-8
    kappa_c: float = 0.25
    verbose: bool = False


@dataclasses.dataclass
class IPMResult:
    x: np.ndarray
    lam: np.ndarray
    zL: np.ndarray
    zU: np.ndarray
    status: int            # 0 converged, 1 max_iter, 2 line-search/restoration failure
    iterations: int
    kkt_error: float
    mu: float
    obj: float


def _push_interior(x, lb, ub, k1, k2):
    x = x.copy()
    hasL, hasU = np.isfinite(lb), np.isfinite(ub)
    both = hasL & hasU
    pL = np.where(hasL, k1 * np.maximum(1.0, np.abs(np.where(hasL, lb, 0.0))), 0.0)
    pU = np.where(hasU, k1 * np.maximum(1.0, np.abs(np.where(hasU, ub, 0.0))), 0.0)
    span = np.where(both, ub - lb, np.inf)
    pL = np.where(both, np.minimum(pL, k2 * span), pL)
    pU = np.where(both, np.minimum(pU, k2 * span), pU)
    x = np.where(hasL, np.maximum(x, lb + pL), x)
    x = np.where(hasU, np.minimum(x, ub - pU), x)
    return x


def solve_ipm(nlp, x0: np.ndarray, opts: Optional[IPMOptions] = None) -> IPMResult:
    o = opts or IPMOptions()
    lb, ub = nlp.lb, nlp.ub
    hasL, hasU = np.isfinite(lb), np.isfinite(ub)
    n, m = nlp.n, nlp.m
    x = _push_interior(x0, lb, ub, o.bound_push, o.bound_frac)
    zL = np.where(hasL, 1.0, 0.0)
    zU = np.where(hasU, 1.0, 0.0)
    mu = o.mu_init
    tau = max(o.tau_min, 1.0 - mu)

    def dL(x):
        return np.where(hasL, x - lb, 1.0)

    def dU(x):
        return np.where(hasU, ub - x, 1.0)

    def barrier(x):
        return nlp.f(x) - mu * (np.log(dL(x))[hasL].sum() + np.log(dU(x))[hasU].sum())

    def theta_of(c):
        return np.abs(c).sum()

    # least-squares multipliers at the start (IPOPT paper 3.6)
    g = nlp.grad(x)
    J = nlp.jac(x)
    K0 = sp.bmat([[sp.identity(n), J.T], [J, None]], format="csc")
    try:
        sol = spla.splu(K0).solve(np.concatenate([-(g - zL + zU), np.zeros(m)]))
        lam = sol[n:]
        if np.abs(lam).max() > 1e3:
            lam = np.zeros(m)
    except Exception:
        lam = np.zeros(m)

    filt = []                       # list of (theta, phi)
    c = nlp.c(x)
    theta_max = 1e4 * max(1.0, theta_of(c))
    theta_min = 1e-4 * max(1.0, theta_of(c))
    delta_w_last = 0.0
    status = 1
    it = 0
    err0 = np.inf

    def kkt_error(x, lam, zL, zU, g, J, c, mu_):
        sd = max(o.s_max, (np.abs(lam).sum() + np.abs(zL).sum() + np.abs(zU).sum()) / (m + hasL.sum() + hasU.sum())) / o.s_max
        sc = max(o.s_max, (np.abs(zL).sum() + np.abs(zU).sum()) / max(1, hasL.sum() + hasU.sum())) / o.s_max
        dual = np.abs(g + J.T @ lam - zL + zU).max() / sd
        prim = np.abs(c).max()
        compL = np.abs(dL(x) * zL - mu_)[hasL].max() if hasL.any() else 0.0
        compU = np.abs(dU(x) * zU - mu_)[hasU].max() if hasU.any() else 0.0
        return max(dual, prim, max(compL, compU) / sc)

    while it < o.max_iter:
        g = nlp.grad(x)
        J = nlp.jac(x)
        c = nlp.c(x)
        err0 = kkt_error(x, lam, zL, zU, g, J, c, 0.0)
        if err0 <= o.tol:
            status = 0
            break
        # barrier update (possibly several times)
        while True:
            errmu = kkt_error(x, lam, zL, zU, g, J, c, mu)
            if errmu <= o.kappa_eps * mu and mu > o.tol / 10 * (1 + 1e-12):
                mu = max(o.tol / 10.0, min(o.kappa_mu * mu, mu ** o.theta_mu))
                tau = max(o.tau_min, 1.0 - mu)
                filt = []
            else:
                break
        # --- search direction --------------------------------------------------------
        W = nlp.hess(x, lam)
        sigL = np.where(hasL, zL / dL(x), 0.0)
        sigU = np.where(hasU, zU / dU(x), 0.0)
        Sig = sigL + sigU
        gphi = g - np.where(hasL, mu / dL(x), 0.0) + np.where(hasU, mu / dU(x), 0.0)
        r1 = gphi + J.T @ lam
        rhs = -np.concatenate([r1, c])
        delta_w = 0.0
        delta_c = 0.0
        attempt = 0
        while True:
            Kmat = sp.bmat([[W + sp.diags(Sig + delta_w), J.T],
                            [J, -delta_c * sp.identity(m) if delta_c > 0 else None]], format="csc")
            ok = True
            try:
                lu = spla.splu(Kmat)
                sol = lu.solve(rhs)
                # one step of iterative refinement
                sol = sol + lu.solve(rhs - Kmat @ sol)
                if not np.all(np.isfinite(sol)):
                    ok = False
            except RuntimeError:
                ok = False
                delta_c = o.delta_c * mu ** o.kappa_c
            if ok:
                dx, dlam = sol[:n], sol[n:]
                curv = dx @ (W @ dx) + (Sig + delta_w) @ (dx * dx)
                if curv >= 1e-10 * (dx @ dx) or (dx @ dx) == 0.0:
                    break
            # regularise (IPOPT paper, Algorithm IC)
            if delta_w == 0.0:
                delta_w = o.delta_w_first if delta_w_last == 0.0 else max(o.delta_w_min, o.kappa_w_minus * delta_w_last)
            else:
                delta_w *= o.kappa_w_plus_first if delta_w_last == 0.0 else o.kappa_w_plus
            attempt += 1
            if delta_w > o.delta_w_max or attempt > 60:
                return IPMResult(x, lam, zL, zU, 2, it, err0, mu, nlp.f(x))
        if delta_w > 0:
            delta_w_last = delta_w
        dzL = np.where(hasL, mu / dL(x) - zL - sigL * dx, 0.0)
        dzU = np.where(hasU, mu / dU(x) - zU + sigU * dx, 0.0)

        # --- fraction to boundary ----------------------------------------------------
        def max_step(v, dv, mask):
            neg = mask & (dv < 0)
            if not neg.any():
                return 1.0
            return min(1.0, float((-tau * v[neg] / dv[neg]).min()))

        a_max = min(max_step(dL(x), dx, hasL), max_step(dU(x), -dx, hasU))
        a_z = min(max_step(zL, dzL, hasL), max_step(zU, dzU, hasU))

        # --- filter line search ------------------------------------------------------
        theta = theta_of(c)
        phi = barrier(x)
        dphi = gphi @ dx
        alpha = a_max
        accepted = False
        alpha_min_fac = 0.05
        if dphi < 0 and theta <= theta_min:
            amin = alpha_min_fac * min(o.gamma_theta, o.gamma_phi * theta / (-dphi) if theta > 0 else np.inf,
                                       o.delta * theta ** o.s_theta / (-dphi) ** o.s_phi if theta > 0 else np.inf)
        elif dphi < 0:
            amin = alpha_min_fac * min(o.gamma_theta, o.gamma_phi * theta / (-dphi))
        else:
            amin = alpha_min_fac * o.gamma_theta
        amin = max(amin, 1e-14)
        first = True
        dx_ls = dx
        while alpha >= amin:
            xt = x + alpha * dx_ls
            ct = nlp.c(xt)
            tht = theta_of(ct)
            pht = barrier(xt)

            def acceptable(tht, pht, alpha):
                if not np.isfinite(pht) or tht > theta_max:
                    return False, False
                for (tj, pj) in filt:
                    if tht >= tj and pht >= pj:
                        return False, False
                ftype = (theta <= theta_min and dphi < 0 and
                         alpha * (-dphi) ** o.s_phi > o.delta * theta ** o.s_theta)
                if ftype:
                    return (pht <= phi + o.eta_phi * alpha * dphi + 10 * np.finfo(float).eps * abs(phi)), True
                ok_ = (tht <= (1 - o.gamma_theta) * theta) or (pht <= phi - o.gamma_phi * theta)
                return ok_, False

            ok_, ftype = acceptable(tht, pht, alpha)
            if ok_:
                accepted = True
                break
            # second-order correction on the first trial (IPOPT paper 2.4)
            if first and tht >= theta and o.max_soc > 0:
                csoc = alpha * c + ct
                th_old = theta
                soc_ok = False
                for p_ in range(o.max_soc):
                    sol = lu.solve(-np.concatenate([r1, csoc]))
                    dxs = sol[:n]
                    a_s = min(max_step(dL(x), dxs, hasL), max_step(dU(x), -dxs, hasU))
                    xs = x + a_s * dxs
                    cs = nlp.c(xs)
                    ths = theta_of(cs)
                    phs = barrier(xs)
                    ok2, ft2 = acceptable(ths, phs, alpha)
                    if ok2:
                        xt, ct, tht, pht, ftype = xs, cs, ths, phs, ft2
                        dlam = sol[n:]
                        dzL = np.where(hasL, mu / dL(x) - zL - sigL * dxs, 0.0)
                        dzU = np.where(hasU, mu / dU(x) - zU + sigU * dxs, 0.0)
                        a_z = min(max_step(zL, dzL, hasL), max_step(zU, dzU, hasU))
                        alpha = a_s
                        soc_ok = True
                        break
                    if ths > o.kappa_soc * th_old:
                        break
                    th_old = ths
                    csoc = a_s * csoc + cs
                if soc_ok:
                    accepted = True
                    break
            first = False
            alpha *= 0.5
        if not accepted:
            # restoration: damped Gauss-Newton on the constraint violation until the
            # filter accepts the point (kept simple; IPOPT paper 3.3 is the full version)
            filt.append(((1 - o.gamma_theta) * theta, phi - o.gamma_phi * theta))
            xr = x.copy()
            restored = False
            for _ in range(50):
                cr = nlp.c(xr)
                Jr = nlp.jac(xr)
                Dr = 1.0 / np.maximum(1.0, np.abs(xr))
                Kr = sp.bmat([[sp.diags(1e-4 * Dr + 1e-8), Jr.T], [Jr, -1e-10 * sp.identity(m)]], format="csc")
                sr = spla.splu(Kr).solve(-np.concatenate([np.zeros(n), cr]))
                dxr = sr[:n]
                ar = min(max_step(dL(xr), dxr, hasL), max_step(dU(xr), -dxr, hasU))
                thr = theta_of(cr)
                while ar > 1e-12:
                    xn = xr + ar * dxr
                    if theta_of(nlp.c(xn)) < (1 - 1e-4 * ar) * thr:
                        break
                    ar *= 0.5
                else:
                    break
                xr = xn
                thn = theta_of(nlp.c(xr))
                phn = barrier(xr)
                if thn <= 0.9 * theta and all(not (thn >= tj and phn >= pj) for (tj, pj) in filt):
                    restored = True
                    break
            if not restored:
                status = 2
                break
            x = xr
            zL = np.where(hasL, np.clip(zL, mu / (o.kappa_sigma * dL(x)), o.kappa_sigma * mu / dL(x)), 0.0)
            zU = np.where(hasU, np.clip(zU, mu / (o.kappa_sigma * dU(x)), o.kappa_sigma * mu / dU(x)), 0.0)
            it += 1
            if o.verbose:
                print(f"{it:4d} restoration  theta {theta:.3e} -> {theta_of(nlp.c(x)):.3e}")
            continue
        if not ftype:
            filt.append(((1 - o.gamma_theta) * theta, phi - o.gamma_phi * theta))
        x = xt
        lam = lam + alpha * dlam
        zL = zL + a_z * dzL
        zU = zU + a_z * dzU
        zL = np.where(hasL, np.clip(zL, mu / (o.kappa_sigma * dL(x)), o.kappa_sigma * mu / dL(x)), 0.0)
        zU = np.where(hasU, np.clip(zU, mu / (o.kappa_sigma * dU(x)), o.kappa_sigma * mu / dU(x)), 0.0)
        it += 1
        if o.verbose:
            print(f"{it:4d} f {nlp.f(x):.10f} theta {tht:.3e} err {err0:.3e} mu {mu:.1e} "
                  f"a {alpha:.3e} az {a_z:.3e} dw {delta_w:.1e} |dx| {np.abs(dx).max():.2e}")
    return IPMResult(x, lam, zL, zU, status, it, err0, mu, nlp.f(x))

"""A GPU-RESIDENT interior-point iteration loop over the C ABI (test scaffolding for SURVEY §8 a13: the solver hand-off).

The reference hands `backend.model` to MadNLP / Ipopt (ext/InfiniteExaModelsMadNLP.jl:49-50, ext/InfiniteExaModelsIpopt.jl:48-49),
which call the NLPModels callbacks at every iterate and factorise the KKT system.  Neither solver (nor Julia) exists here, so this
file is the smallest loop with the same STRUCTURE: a primal-dual log-barrier method whose iterate (x, slacks, multipliers) never
leaves the device —

  * per iteration ONE `iexa_eval3` (cons! + jac_coord! + hess_coord! at the current (x, y), device buffers, one fused launch),
    one `iexa_grad`, and `iexa_obj_device` + `iexa_cons` per line-search trial;
  * the COO values are scattered into the KKT matrix ON the device (index_put_ with accumulation: duplicates summed, the job
    `iexa_csr_apply` does for cuDSS) and the KKT solve — the reported NON-TARGET cost of the north star — is a dense
    `torch.linalg.solve` (the test models are tiny; cuDSS is not installed);
  * host synchronisation happens only where a solver has to branch on a scalar (merit value, convergence test).

It is NOT MadNLP: no filter line search, no inertia-correcting LDLᵀ — iterate traces cannot be compared with the reference's
(unmeasurable here), objective values can (tests/test_gpu_resident_solve.py: the reference's solve-level goldens).
"""
import ctypes as C

import numpy as np


class DeviceCallbacks:
    """the C ABI with DEVICE buffers (MadNLP-style): every method enqueues kernels on the current stream, nothing is copied"""

    def __init__(self, m, ex):
        import torch
        self.m, self.ex, self.torch = m, ex, torch
        self.dev = torch.device("cuda", m.device)
        self.nvar, self.ncon, self.nnzj, self.nnzh = m.meta.nvar, m.meta.ncon, m.meta.nnzj, m.meta.nnzh
        self.minimize = m.meta.minimize
        self.meta = m.meta
        z = lambda n: torch.zeros(max(n, 1), dtype=torch.float64, device=self.dev)
        self.c, self.jv, self.hv, self.g, self.f, self.ct = z(self.ncon), z(self.nnzj), z(self.nnzh), z(self.nvar), z(1), z(self.ncon)
        self.counts = dict(eval3=0, grad=0, obj=0, cons=0)

    def structures(self):
        torch, ex, m = self.torch, self.ex, self.m
        jr = torch.zeros(max(self.nnzj, 1), dtype=torch.int64, device=self.dev); jc = torch.zeros_like(jr)
        hr = torch.zeros(max(self.nnzh, 1), dtype=torch.int64, device=self.dev); hc = torch.zeros_like(hr)
        ex.jac_structure_(m, jr, jc); ex.hess_structure_(m, hr, hc)
        return jr[: self.nnzj] - 1, jc[: self.nnzj] - 1, hr[: self.nnzh] - 1, hc[: self.nnzh] - 1

    def eval3(self, x, y, sigma):   # ONE fused launch: cons! + jac_coord! + hess_coord!
        self.ex.eval3_(self.m, x, y, self.c, self.jv, self.hv, sigma); self.counts["eval3"] += 1
        return self.c[: self.ncon], self.jv[: self.nnzj], self.hv[: self.nnzh]

    def grad(self, x):
        self.ex.grad_(self.m, x, self.g); self.counts["grad"] += 1
        return self.g

    def obj(self, x):               # device scalar, no host synchronisation
        m = self.m
        st = self.torch.cuda.current_stream(self.dev).cuda_stream
        self.ex.lib.check(m.L, m.L.iexa_obj_device(m.h, C.c_void_p(x.data_ptr()), C.c_void_p(self.f.data_ptr()), C.c_void_p(st)))
        self.counts["obj"] += 1
        return self.f[0].clone()

    def cons(self, x):
        self.ex.cons_(self.m, x, self.ct); self.counts["cons"] += 1
        return self.ct[: self.ncon].clone()


class OracleCallbacks:
    """the same interface over the CPU oracle (host tensors): lets the ALGORITHM be tested without a GPU"""

    def __init__(self, om):
        import torch
        self.om, self.torch = om, torch
        self.dev = torch.device("cpu")
        self.nvar, self.ncon, self.nnzj, self.nnzh = om.nvar, om.ncon, om.nnzj, om.nnzh
        self.minimize = om.minimize
        self.meta = type("M", (), dict(lvar=om.lvar, uvar=om.uvar, lcon=om.lcon, ucon=om.ucon, x0=om.x0))
        self.counts = dict(eval3=0, grad=0, obj=0, cons=0)

    def structures(self):
        t = self.torch
        jr, jc = self.om.jac_structure(); hr, hc = self.om.hess_structure()
        return tuple(t.from_numpy(a.astype(np.int64)) - 1 for a in (jr, jc, hr, hc))

    def eval3(self, x, y, sigma):
        t, xn, yn = self.torch, x.numpy(), y.numpy()
        self.counts["eval3"] += 1
        return t.from_numpy(self.om.cons(xn)), t.from_numpy(self.om.jac_coord(xn)), t.from_numpy(self.om.hess_coord(xn, yn, sigma))

    def grad(self, x): self.counts["grad"] += 1; return self.torch.from_numpy(self.om.grad(x.numpy()))
    def obj(self, x): self.counts["obj"] += 1; return self.torch.tensor(self.om.obj(x.numpy()), dtype=self.torch.float64)
    def cons(self, x): self.counts["cons"] += 1; return self.torch.from_numpy(self.om.cons(x.numpy()))


def solve_on_device(cb, tol=1e-9, max_iter=300, mu0=0.1, x0=None, verbose=False):
    """primal-dual log-barrier Newton iteration; every vector lives on ``cb.dev``"""
    import torch
    dev, dt = cb.dev, torch.float64
    nv, nc = cb.nvar, cb.ncon
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev)
    lx, ux, lc, uc = T(cb.meta.lvar), T(cb.meta.uvar), T(cb.meta.lcon), T(cb.meta.ucon)
    eq = lc == uc
    iq = torch.nonzero(~eq).flatten()
    ni = int(iq.numel())
    jr, jc, hr, hc = cb.structures()
    off = hr != hc
    lw = torch.cat([lx, lc[iq]]); uw = torch.cat([ux, uc[iq]])
    has_l, has_u = torch.isfinite(lw), torch.isfinite(uw)

    def interior(v, lo, hi, k1=1e-2, k2=1e-2):   # Ipopt's bound_push / bound_frac
        span = torch.where(torch.isfinite(hi - lo), hi - lo, torch.full_like(v, float("inf")))
        pl = torch.minimum(k1 * torch.clamp(lo.abs(), min=1.0), k2 * span)
        pu = torch.minimum(k1 * torch.clamp(hi.abs(), min=1.0), k2 * span)
        v = torch.where(torch.isfinite(lo), torch.maximum(v, lo + pl), v)
        v = torch.where(torch.isfinite(hi), torch.minimum(v, hi - pu), v)
        return v

    x = interior(T(cb.meta.x0 if x0 is None else x0).clone(), lx, ux)
    sgn = 1.0 if cb.minimize else -1.0
    y = torch.zeros(nc, dtype=dt, device=dev)
    s = interior(cb.cons(x)[iq].clone(), lc[iq], uc[iq]) if ni else torch.zeros(0, dtype=dt, device=dev)
    w = torch.cat([x, s])
    n = nv + ni
    mu = mu0
    one, zero = torch.ones(n, dtype=dt, device=dev), torch.zeros(n, dtype=dt, device=dev)
    zl = torch.where(has_l, mu / torch.clamp(w - lw, min=1e-12), zero)
    zu = torch.where(has_u, mu / torch.clamp(uw - w, min=1e-12), zero)
    slack_cols = nv + torch.arange(ni, device=dev)

    def resid(cc, ss):
        r = cc.clone()
        r[eq] = r[eq] - lc[eq]
        if ni:
            r[iq] = r[iq] - ss
        return r

    def barrier(ww):
        b = torch.zeros((), dtype=dt, device=dev)
        if has_l.any():
            b = b - torch.log(ww[has_l] - lw[has_l]).sum()
        if has_u.any():
            b = b - torch.log(uw[has_u] - ww[has_u]).sum()
        return b

    err0, iters = float("inf"), 0
    for it in range(max_iter):
        iters = it + 1
        xcur = w[:nv].contiguous()
        c, jv, hv = cb.eval3(xcur, y if nc else torch.zeros(1, dtype=dt, device=dev), sgn)
        gr = sgn * cb.grad(xcur)[:nv]
        J = torch.zeros(nc, nv, dtype=dt, device=dev)
        J.index_put_((jr, jc), jv, accumulate=True)          # duplicates summed on the device (what iexa_csr_apply does for cuDSS)
        H = torch.zeros(nv, nv, dtype=dt, device=dev)
        H.index_put_((hr, hc), hv, accumulate=True)
        H.index_put_((hc[off], hr[off]), hv[off], accumulate=True)
        A = torch.zeros(nc, n, dtype=dt, device=dev)
        A[:, :nv] = J
        if ni:
            A[iq, slack_cols] = -1.0
        r = resid(c, w[nv:])
        gw = torch.cat([gr, torch.zeros(ni, dtype=dt, device=dev)])
        dl, du = torch.where(has_l, w - lw, one), torch.where(has_u, uw - w, one)
        dual = gw + A.T @ y - zl + zu
        comp = torch.cat([(zl * dl)[has_l], (zu * du)[has_u], torch.zeros(1, dtype=dt, device=dev)])
        rmax = r.abs().max().item() if nc else 0.0
        err0 = max(dual.abs().max().item(), rmax, comp.abs().max().item())
        if verbose:
            print(f"it {it:3d} f {cb.obj(xcur).item():+.10e} mu {mu:.1e} err {err0:.2e}")
        if err0 <= tol:
            break
        cm = torch.cat([(zl * dl - mu)[has_l], (zu * du - mu)[has_u], torch.zeros(1, dtype=dt, device=dev)])
        errmu = max(dual.abs().max().item(), rmax, cm.abs().max().item())
        while errmu <= 10.0 * mu and mu > tol / 10:
            mu = max(tol / 10, min(0.2 * mu, mu ** 1.5))
            cm = torch.cat([(zl * dl - mu)[has_l], (zu * du - mu)[has_u], torch.zeros(1, dtype=dt, device=dev)])
            errmu = max(dual.abs().max().item(), rmax, cm.abs().max().item())
        Sig = torch.where(has_l, zl / dl, zero) + torch.where(has_u, zu / du, zero)
        bgrad = gw - torch.where(has_l, mu / dl, zero) + torch.where(has_u, mu / du, zero)
        rhs_w = bgrad + A.T @ y
        Hw = torch.zeros(n, n, dtype=dt, device=dev)
        Hw[:nv, :nv] = H
        delta, dw, dy = 0.0, None, None
        for _try in range(14):   # primal regularisation until the step is a descent direction of the barrier problem
            K = torch.zeros(n + nc, n + nc, dtype=dt, device=dev)
            K[:n, :n] = Hw + torch.diag(Sig + delta)
            K[:n, n:] = A.T
            K[n:, :n] = A
            K[n:, n:] = -1e-11 * torch.eye(nc, dtype=dt, device=dev)
            try:
                sol = torch.linalg.solve(K, -torch.cat([rhs_w, r]))
            except Exception:
                delta = max(1e-8, 10 * delta); continue
            dw, dy = sol[:n], sol[n:]
            curv = dw @ ((Hw + torch.diag(Sig + delta)) @ dw)
            if torch.isfinite(sol).all() and curv.item() > 1e-14 * (dw @ dw).item():
                break
            delta = max(1e-8, 10 * delta)
        dzl = torch.where(has_l, mu / dl - zl - (zl / dl) * dw, zero)
        dzu = torch.where(has_u, mu / du - zu + (zu / du) * dw, zero)
        tau = max(0.99, 1.0 - mu)

        def max_step(v, dv, mask):   # largest a <= 1 with v + a dv >= (1 - tau) v
            neg = mask & (dv < 0)
            return min(1.0, float((-tau * v[neg] / dv[neg]).min().item())) if neg.any() else 1.0
        ap = min(max_step(dl, dw, has_l), max_step(du, -dw, has_u))
        ad = min(max_step(zl, dzl, has_l), max_step(zu, dzu, has_u))
        ynew_max = float((y + dy).abs().max().item()) if nc else 0.0
        nu = max(1.0, 2.0 * ynew_max)
        rn = r.abs().sum()
        phi0 = sgn * cb.obj(xcur) + mu * barrier(w) + nu * rn
        dphi = bgrad @ dw - nu * rn
        a = ap
        for _ls in range(40):
            wt = w + a * dw
            xt = wt[:nv].contiguous()
            rt = resid(cb.cons(xt), wt[nv:]) if nc else r
            phit = sgn * cb.obj(xt) + mu * barrier(wt) + nu * rt.abs().sum()
            if bool(torch.isfinite(phit)) and (phit.item() <= phi0.item() + 1e-8 * a * min(dphi.item(), 0.0) + 1e-14 * abs(phi0.item()) or a < 1e-12):
                break
            a *= 0.5
        w = w + a * dw
        y = y + a * dy
        zl = torch.where(has_l, zl + ad * dzl, zl); zu = torch.where(has_u, zu + ad * dzu, zu)
        dl, du = torch.where(has_l, w - lw, one), torch.where(has_u, uw - w, one)
        zl = torch.where(has_l, torch.minimum(torch.maximum(zl, mu / (1e10 * dl)), 1e10 * mu / dl), zl)
        zu = torch.where(has_u, torch.minimum(torch.maximum(zu, mu / (1e10 * du)), 1e10 * mu / du), zu)
    xfin = w[:nv].contiguous()
    return dict(x=xfin, y=y, f=float(cb.obj(xfin).item()), err=err0, iters=iters, mu=mu, counts=dict(cb.counts))

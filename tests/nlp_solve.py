"""Tiny NLP driver used by the golden-solve tests: scipy's trust-constr fed ONLY through the
NLPModels-style callbacks (obj, grad, cons, jac COO, Hessian-of-Lagrangian COO), i.e. the same
information MadNLP / Ipopt receive from the model (ext/InfiniteExaModelsMadNLP.jl:49-50,
ext/InfiniteExaModelsIpopt.jl:48-49)."""
import numpy as np
import scipy.sparse as sp
from scipy.optimize import Bounds, NonlinearConstraint, minimize


class Callbacks:
    """adapter: anything exposing obj/grad/cons/jac_coord/hess_coord + structures"""

    def __init__(self, nvar, ncon, lvar, uvar, lcon, ucon, x0, obj, grad, cons, jac_coord, hess_coord,
                 jac_structure, hess_structure):
        self.nvar, self.ncon = nvar, ncon
        self.lvar, self.uvar, self.lcon, self.ucon, self.x0 = lvar, uvar, lcon, ucon, x0
        self.obj, self.grad, self.cons, self.jac_coord, self.hess_coord = obj, grad, cons, jac_coord, hess_coord
        self.jr, self.jc = jac_structure
        self.hr, self.hc = hess_structure

    def jac(self, x):
        return sp.coo_matrix((self.jac_coord(x), (self.jr - 1, self.jc - 1)), shape=(self.ncon, self.nvar)).tocsr()

    def hess(self, x, y, sigma):
        v = self.hess_coord(x, y, sigma)
        L = sp.coo_matrix((v, (self.hr - 1, self.hc - 1)), shape=(self.nvar, self.nvar)).tocsr()
        return L + sp.tril(L, -1).T


def from_oracle(om):
    return Callbacks(om.nvar, om.ncon, om.lvar, om.uvar, om.lcon, om.ucon, om.x0, om.obj, om.grad, om.cons,
                     om.jac_coord, lambda x, y, s: om.hess_coord(x, y, s), om.jac_structure(), om.hess_structure())


def from_examodel(m):
    """GPU engine through the C ABI with host buffers (the Ipopt-style path)."""
    import iexa_b200 as ex
    nv, nc, nj, nh = m.meta.nvar, m.meta.ncon, m.meta.nnzj, m.meta.nnzh
    jr, jc = np.zeros(nj, dtype=np.int64), np.zeros(nj, dtype=np.int64)
    hr, hc = np.zeros(nh, dtype=np.int64), np.zeros(nh, dtype=np.int64)
    ex.jac_structure_(m, jr, jc)
    ex.hess_structure_(m, hr, hc)
    c = lambda x: np.ascontiguousarray(x, dtype=np.float64)
    return Callbacks(
        nv, nc, m.meta.lvar, m.meta.uvar, m.meta.lcon, m.meta.ucon, m.meta.x0,
        lambda x: ex.obj(m, c(x)), lambda x: ex.grad_(m, c(x), np.zeros(nv)), lambda x: ex.cons_(m, c(x), np.zeros(nc)),
        lambda x: ex.jac_coord_(m, c(x), np.zeros(nj)),
        lambda x, y, s: ex.hess_coord_(m, c(x), None if y is None else c(y), np.zeros(nh), s), (jr, jc), (hr, hc))


def solve(cb: Callbacks, x0=None, tol=1e-13, maxiter=3000):
    x0 = cb.x0.copy() if x0 is None else np.asarray(x0, dtype=np.float64)
    x0 = np.minimum(np.maximum(x0, np.where(np.isfinite(cb.lvar), cb.lvar, -1e20)), np.where(np.isfinite(cb.uvar), cb.uvar, 1e20))
    zeros = np.zeros(cb.ncon)
    con = NonlinearConstraint(cb.cons, cb.lcon, cb.ucon, jac=cb.jac, hess=lambda x, v: cb.hess(x, v, 0.0))
    res = minimize(cb.obj, x0, jac=cb.grad, hess=lambda x: cb.hess(x, zeros, 1.0), method="trust-constr",
                   constraints=[con] if cb.ncon else [], bounds=Bounds(cb.lvar, cb.uvar),
                   options=dict(gtol=tol, xtol=1e-14, barrier_tol=1e-12, maxiter=maxiter, initial_barrier_parameter=0.1,
                                initial_barrier_tolerance=0.1))
    return res

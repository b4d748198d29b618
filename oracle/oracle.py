"""ctypes wrapper of oracle/liboracle.so — TEST INFRASTRUCTURE ONLY (see oracle.c header).

PARITY UNPINNED at the callback level: the reference's evaluator (ExaModels.jl 0.11.2) is an
absent third-party dependency; this oracle is pinned only through the reference's solve-level
goldens (tests/test_golden_solves.py).

Importers allowed: tests/, __graft_entry__.smoke(), bench.py (cpu_baseline / --impl reference).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        vp, i64, i32, dbl = C.c_void_p, C.c_int64, C.c_int32, C.c_double
        L.orc_create.restype = vp
        L.orc_set_threads.argtypes = [i32]
        L.orc_max_threads.restype = i32
        L.orc_free.argtypes = [vp]
        L.orc_set_dims.argtypes = [vp, i64, i64, vp]
        L.orc_set_theta.argtypes = [vp, i64, i64, vp]
        L.orc_add_gen.argtypes = [vp, i32, vp, i32, vp, i32, i64, i32, vp, i32, vp, dbl, dbl]
        L.orc_add_gen.restype = i32
        L.orc_add_gen_borrow.argtypes = L.orc_add_gen.argtypes
        L.orc_add_gen_borrow.restype = i32
        L.orc_set_slot_order.argtypes = [vp, i32]
        L.orc_jac_is_csr.argtypes = [vp]; L.orc_jac_is_csr.restype = i32
        L.orc_finalize.argtypes = [vp]
        for f in ("orc_ncon", "orc_nnzj", "orc_nnzh"):
            getattr(L, f).argtypes = [vp]
            getattr(L, f).restype = i64
        L.orc_gen_info.argtypes = [vp, i32, vp]
        L.orc_obj.argtypes = [vp, vp]
        L.orc_obj.restype = dbl
        L.orc_cons.argtypes = [vp, vp, vp]
        L.orc_grad.argtypes = [vp, vp, vp]
        L.orc_jac_structure.argtypes = [vp, vp, vp]
        L.orc_jac_coord.argtypes = [vp, vp, vp]
        L.orc_jprod.argtypes = [vp, vp, vp, vp]
        L.orc_jtprod.argtypes = [vp, vp, vp, vp]
        L.orc_hess_structure.argtypes = [vp, vp, vp]
        L.orc_hess_coord.argtypes = [vp, vp, vp, dbl, vp]
        L.orc_hprod.argtypes = [vp, vp, vp, vp, dbl, vp]
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def set_threads(n: int) -> None:
    """OpenMP threads of the oracle's loops over supports (omp_set_num_threads: works even when another
    library initialised libgomp first)"""
    lib().orc_set_threads(int(n))


def max_threads() -> int:
    return int(lib().orc_max_threads())


class OracleModel:
    """CPU evaluation of an ``ExaCore`` description with the restated ExaModels algorithm."""

    def __init__(self, core, slot_order: int = 0):
        L = lib()
        self.L = L
        self.h = L.orc_create()
        L.orc_set_slot_order(self.h, int(slot_order))   # the slot-ORDER policy is data (0: inner1 then inner2, 1: reversed)
        self.nvar = core.nvar
        theta = np.ascontiguousarray(core.theta_vec, dtype=np.float64)
        L.orc_set_dims(self.h, core.nvar, core.npar, _p(theta))
        self.x0 = core.x0_vec.copy()
        self.lvar, self.uvar = core.lvar_vec.copy(), core.uvar_vec.copy()
        lcon, ucon = [], []
        self._keep = []
        cols = {}   # iterator -> materialised columns, shared by every generator over it (kept alive by self._keep)
        for g in core.gens:
            if id(g.itr) not in cols:
                ic, fc = g.itr.materialise()
                ic = [np.ascontiguousarray(c, dtype=np.int64) for c in ic]
                fc = [np.ascontiguousarray(c, dtype=np.float64) for c in fc]
                icp = (C.c_void_p * max(len(ic), 1))(*[c.ctypes.data for c in ic])
                fcp = (C.c_void_p * max(len(fc), 1))(*[c.ctypes.data for c in fc])
                cols[id(g.itr)] = (ic, fc, icp, fcp)
                self._keep.append((g.itr, ic, fc, icp, fcp))
            ic, fc, icp, fcp = cols[id(g.itr)]
            nodes = np.ascontiguousarray(g.tape.nodes)
            index = np.ascontiguousarray(g.tape.index)
            L.orc_add_gen_borrow(self.h, int(g.is_obj), _p(nodes), len(nodes), _p(index), len(index),
                                 g.itr.K, len(ic), icp, len(fc), fcp, g.lcon, g.ucon)
            if not g.is_obj:
                lcon.append(np.full(g.itr.K, g.lcon)); ucon.append(np.full(g.itr.K, g.ucon))
        L.orc_finalize(self.h)
        self.ngen = len(core.gens)
        self.ncon = L.orc_ncon(self.h)
        self.nnzj = L.orc_nnzj(self.h)
        self.nnzh = L.orc_nnzh(self.h)
        self.lcon = np.concatenate(lcon) if lcon else np.zeros(0)
        self.ucon = np.concatenate(ucon) if ucon else np.zeros(0)
        self.minimize = core.minimize

    def __del__(self):
        try:
            self.L.orc_free(self.h)
        except Exception:
            pass

    def gen_info(self, i):
        out = np.zeros(8, dtype=np.int64)
        self.L.orc_gen_info(self.h, i, _p(out))
        return dict(o0=out[0], o1=out[1], o2=out[2], o1step=out[3], o2step=out[4], K=out[5],
                    nocc1=out[6], nocc2=out[7])

    def set_parameter(self, offset0, vals):
        v = np.ascontiguousarray(vals, dtype=np.float64)
        self.L.orc_set_theta(self.h, int(offset0), len(v), _p(v))

    @staticmethod
    def _x(x):
        return np.ascontiguousarray(x, dtype=np.float64)

    def obj(self, x):
        return float(self.L.orc_obj(self.h, _p(self._x(x))))

    def cons(self, x):
        c = np.zeros(self.ncon)
        self.L.orc_cons(self.h, _p(self._x(x)), _p(c))
        return c

    def grad(self, x):
        g = np.zeros(self.nvar)
        self.L.orc_grad(self.h, _p(self._x(x)), _p(g))
        return g

    def jac_structure(self):
        r = np.zeros(self.nnzj, dtype=np.int64); c = np.zeros(self.nnzj, dtype=np.int64)
        self.L.orc_jac_structure(self.h, _p(r), _p(c))
        return r, c

    def jac_coord(self, x):
        v = np.zeros(self.nnzj)
        self.L.orc_jac_coord(self.h, _p(self._x(x)), _p(v))
        return v

    def hess_structure(self):
        r = np.zeros(self.nnzh, dtype=np.int64); c = np.zeros(self.nnzh, dtype=np.int64)
        self.L.orc_hess_structure(self.h, _p(r), _p(c))
        return r, c

    def hess_coord(self, x, y=None, obj_weight=1.0):
        v = np.zeros(self.nnzh)
        yy = None if y is None else np.ascontiguousarray(y, dtype=np.float64)
        self.L.orc_hess_coord(self.h, _p(self._x(x)), _p(yy), float(obj_weight), _p(v))
        return v

    def jprod(self, x, v):
        out = np.zeros(self.ncon)
        self.L.orc_jprod(self.h, _p(self._x(x)), _p(self._x(v)), _p(out))
        return out

    def jtprod(self, x, v):
        out = np.zeros(self.nvar)
        self.L.orc_jtprod(self.h, _p(self._x(x)), _p(self._x(v)), _p(out))
        return out

    def hprod(self, x, y, v, obj_weight=1.0):
        out = np.zeros(self.nvar)
        yy = None if y is None else np.ascontiguousarray(y, dtype=np.float64)
        self.L.orc_hprod(self.h, _p(self._x(x)), _p(yy), _p(self._x(v)), float(obj_weight), _p(out))
        return out

"""``ExaModel`` — the NLPModel the solvers consume — and the NLPModels callback API.

Mirrors what the reference hands to ``MadNLPSolver(model; ...)`` / ``IpoptSolver(model)``
(ext/InfiniteExaModelsMadNLP.jl:49, ext/InfiniteExaModelsIpopt.jl:48): an object with ``meta``
(nvar, ncon, nnzj, nnzh, x0, lvar, uvar, y0, lcon, ucon, minimize), a mutable parameter vector
``θ`` (infiniteopt_backend.jl:479,522,546) and the callbacks ``obj, grad!, cons!,
jac_structure!, jac_coord!, hess_structure!, hess_coord!, jprod!, jtprod!, hprod!``.
Python has no ``!``; the in-place callbacks carry a trailing underscore.

Every callback is a single call into the C ABI (libiexa_b200.so).  Buffers may be numpy arrays
(host memory: the Ipopt-style path, copies happen inside the call) or CUDA torch tensors (device
memory: the MadNLP-style path, asynchronous on the current torch stream).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import lib as _lib
from .core import ExaCore, Itr


class NLPModelMeta:
    """``NLPModels.NLPModelMeta``: nvar, ncon, nnzj, nnzh, minimize and the vectors x0, lvar, uvar, y0, lcon, ucon.
    The vectors are fetched from the plan on first access (and then kept: they are ordinary mutable arrays, like
    ``NLPModels.get_x0(model)`` — infiniteopt_backend.jl:600-601); a 10^6-support model never pays for the six
    44-million-entry copies unless somebody asks for them."""
    _VECTORS = {"x0": 0, "lvar": 1, "uvar": 2, "lcon": 3, "ucon": 4, "y0": 5}

    def __init__(self, nvar, ncon, nnzj, nnzh, minimize, fetch):
        self.nvar, self.ncon, self.nnzj, self.nnzh, self.minimize = nvar, ncon, nnzj, nnzh, minimize
        self._fetch, self._cache = fetch, {}

    def __getattr__(self, name):
        if name in NLPModelMeta._VECTORS:
            if name not in self._cache:
                which = NLPModelMeta._VECTORS[name]
                self._cache[name] = self._fetch(which, self.nvar if which <= 2 else self.ncon)
            return self._cache[name]
        raise AttributeError(name)


def _is_torch(a):
    return type(a).__module__.startswith("torch")


class ExaModel:
    """``ExaModels.ExaModel(core)`` (infiniteopt_backend.jl:156) backed by the CUDA engine."""

    def __init__(self, core: ExaCore, device: int = 0, rank: int = 0, world: int = 1,
                 flags: int = _lib.IEXA_F_DEFAULT, library=None, slot_order: int = 0, strict_ieee: Optional[bool] = None):
        self.L = L = library or _lib.load()
        self.core = core
        h = C.c_void_p()
        _lib.check(L, L.iexa_plan_create(C.byref(h), int(core.minimize)))
        self.h = h
        if slot_order:
            _lib.check(L, L.iexa_set_option(h, _lib.IEXA_OPT_SLOT_ORDER, int(slot_order)))
        if strict_ieee is not None:   # None: the engine's default (strict)
            _lib.check(L, L.iexa_set_option(h, _lib.IEXA_OPT_STRICT_IEEE, int(bool(strict_ieee))))
        self._keep = []
        off = C.c_int64()
        if core._x0 is None and len(core.var_defaults) == len(core.x0) and core.nvar:
            # block by block, as the reference calls add_var (transform.jl:113,154): no concatenated 8*nvar-byte host vectors, and a
            # block whose start / bounds are the defaults (0, -inf, +inf) passes NULL — the engine fills them itself
            for (n, d0, dl, du), a0, al, au in zip(core.var_defaults, core.x0, core.lvar, core.uvar):
                _lib.check(L, L.iexa_add_var(h, n, None if d0 else a0.ctypes.data, None if dl else al.ctypes.data,
                                             None if du else au.ctypes.data, C.byref(off)))
        elif core.nvar:
            x0, lv, uv = (np.ascontiguousarray(v, dtype=np.float64) for v in (core.x0_vec, core.lvar_vec, core.uvar_vec))
            _lib.check(L, L.iexa_add_var(h, core.nvar, x0.ctypes.data, lv.ctypes.data, uv.ctypes.data, C.byref(off)))
        th = np.ascontiguousarray(core.theta_vec, dtype=np.float64)
        itr_ids = {}

        def itr_id(it: Itr) -> int:
            if id(it) in itr_ids:
                return itr_ids[id(it)]
            out = C.c_int32()
            if it.factors is None:
                if it.K == 1 and not it.ints and not it.fps:
                    itr_ids[id(it)] = 0  # the empty iterator [(;)]
                    return 0
                ic = [it.ints[n] for n in it.ints]
                fc = [it.fps[n] for n in it.fps]
                if getattr(it, "gens", None) or getattr(it, "iota", None):
                    # device-side transcription: generated fp columns / iota int columns are DESCRIBED, not passed as data
                    icp = (C.c_void_p * max(len(ic), 1))(*[None if n in it.iota else it.ints[n].ctypes.data for n in it.ints])
                    fcp = (C.c_void_p * max(len(fc), 1))(*[None if n in it.gens else it.fps[n].ctypes.data for n in it.fps])
                    gens = (_lib.ColGen * max(len(fc), 1))()
                    names = list(it.fps)
                    for j, n in enumerate(names):
                        g = it.gens.get(n)
                        if g is not None:
                            gens[j].kind, gens[j].n, gens[j].a, gens[j].b = g.kind, g.n, g.a, g.b
                            gens[j].src = names.index(g.src) if g.src is not None else -1
                    _lib.check(L, L.iexa_itr_generated(h, it.K, len(ic), icp, len(fc), gens, fcp, C.byref(out)))
                    itr_ids[id(it)] = out.value
                    return out.value
                icp = (C.c_void_p * max(len(ic), 1))(*[c.ctypes.data for c in ic])
                fcp = (C.c_void_p * max(len(fc), 1))(*[c.ctypes.data for c in fc])
                _lib.check(L, L.iexa_itr_base(h, it.K, len(ic), icp, len(fc), fcp, C.byref(out)))
            else:
                ids = (C.c_int32 * len(it.factors))(*[itr_id(f) for f in it.factors])
                _lib.check(L, L.iexa_itr_product(h, len(it.factors), ids, C.byref(out)))
            itr_ids[id(it)] = out.value
            return out.value

        # theta: data blocks through iexa_add_par, parameter FUNCTIONS as tapes the engine evaluates on the device
        pos = 0
        for par, tape, pit in sorted(getattr(core, "par_functions", []), key=lambda t: t[0].offset):
            if par.offset > pos:
                _lib.check(L, L.iexa_add_par(h, par.offset - pos, th[pos:].ctypes.data, C.byref(off)))
            nodes = np.ascontiguousarray(tape.nodes); index = np.ascontiguousarray(tape.index)
            _lib.check(L, L.iexa_add_par_function(h, nodes.ctypes.data, len(nodes), index.ctypes.data if len(index) else None,
                                                  len(index), itr_id(pit), C.byref(off)))
            assert off.value == par.offset
            pos = par.offset + par.length
        if core.npar > pos:
            _lib.check(L, L.iexa_add_par(h, core.npar - pos, th[pos:].ctypes.data, C.byref(off)))
        for g in core.gens:
            nodes = np.ascontiguousarray(g.tape.nodes)
            index = np.ascontiguousarray(g.tape.index)
            iid = itr_id(g.itr)
            ip = index.ctypes.data if len(index) else None
            if g.is_obj:
                _lib.check(L, L.iexa_add_obj(h, nodes.ctypes.data, len(nodes), ip, len(index), iid))
            else:
                _lib.check(L, L.iexa_add_con(h, nodes.ctypes.data, len(nodes), ip, len(index), iid,
                                             g.lcon, g.ucon, C.byref(off)))
        _lib.check(L, L.iexa_finalize(h, device, rank, world, flags))
        m = _lib.Meta()
        _lib.check(L, L.iexa_get_meta(h, C.byref(m)))
        self.cmeta = m
        self.device, self.rank, self.world = device, rank, world

        def vec(which, n):
            a = np.zeros(n)
            if n:
                _lib.check(L, L.iexa_get_vector(h, which, a.ctypes.data))
            return a

        self.meta = NLPModelMeta(m.nvar, m.ncon, m.nnzj, m.nnzh, bool(m.minimize), vec)
        # local (this rank's) sizes; equal to the global ones when world == 1
        self.loc_ncon, self.loc_nnzj, self.loc_nnzh = m.loc_ncon, m.loc_nnzj, m.loc_nnzh

    def __del__(self):
        try:
            if self.h:
                self.L.iexa_plan_destroy(self.h)
                self.h = None
        except Exception:
            pass

    # ---- θ (model.θ, set_parameter!) -----------------------------------------------------------
    @property
    def θ(self) -> np.ndarray:
        n = self.cmeta.npar
        a = np.zeros(n)
        if n:
            _lib.check(self.L, self.L.iexa_get_par(self.h, 0, n, a.ctypes.data))
        return a

    theta = θ

    def set_parameter(self, par, vals) -> None:
        """``ExaModels.set_parameter!(core, param, vals)`` — infiniteopt_backend.jl:522,546."""
        v = np.ascontiguousarray(np.asarray(vals, dtype=np.float64).reshape(-1, order="F"))
        assert v.size == par.length, "parameter block length mismatch"
        _lib.check(self.L, self.L.iexa_set_par(self.h, par.offset, v.size, v.ctypes.data))

    # ---- buffers ---------------------------------------------------------------------------------
    def _buf(self, a, n=None, write=False):
        if a is None:
            return None, None, None
        if _is_torch(a):
            import torch
            assert a.dtype in (torch.float64, torch.int32, torch.int64) and a.is_contiguous()
            if n is not None:
                assert a.numel() >= n, f"buffer too small: {a.numel()} < {n}"
            if a.is_cuda:
                return a.data_ptr(), _lib.IEXA_MEM_DEVICE, torch.cuda.current_stream(a.device).cuda_stream
            return a.data_ptr(), _lib.IEXA_MEM_HOST, None
        assert isinstance(a, np.ndarray) and a.flags["C_CONTIGUOUS"]
        if n is not None:
            assert a.size >= n, f"buffer too small: {a.size} < {n}"
        return a.ctypes.data, _lib.IEXA_MEM_HOST, None

    def _pair(self, *bufs):
        ms = {b[1] for b in bufs if b[0] is not None}
        assert len(ms) == 1, "all buffers of one call must live in the same memory space"
        st = next((b[2] for b in bufs if b[2] is not None), None)
        return ms.pop(), st


# ---- NLPModels API ------------------------------------------------------------------------------------
def get_x0(m: ExaModel): return m.meta.x0
def get_y0(m: ExaModel): return m.meta.y0


def obj(m: ExaModel, x) -> float:
    bx = m._buf(x, m.meta.nvar)
    f = C.c_double()
    _lib.check(m.L, m.L.iexa_obj(m.h, bx[0], C.byref(f), bx[1], bx[2]))
    return f.value


def grad_(m: ExaModel, x, g):
    bx, bg = m._buf(x, m.meta.nvar), m._buf(g, m.meta.nvar)
    ms, st = m._pair(bx, bg)
    _lib.check(m.L, m.L.iexa_grad(m.h, bx[0], bg[0], ms, st))
    return g


def cons_(m: ExaModel, x, c):
    bx, bc = m._buf(x, m.meta.nvar), m._buf(c, m.loc_ncon)
    ms, st = m._pair(bx, bc)
    _lib.check(m.L, m.L.iexa_cons(m.h, bx[0], bc[0], ms, st))
    return c


def _idx_bytes(a):
    if _is_torch(a):
        import torch
        return 4 if a.dtype == torch.int32 else 8
    return a.dtype.itemsize


def jac_structure_(m: ExaModel, rows, cols):
    br, bc = m._buf(rows, m.loc_nnzj), m._buf(cols, m.loc_nnzj)
    ms, st = m._pair(br, bc)
    _lib.check(m.L, m.L.iexa_jac_structure(m.h, br[0], bc[0], _idx_bytes(rows), ms, st))
    return rows, cols


def device_bytes(m: ExaModel) -> dict:
    """device memory the engine holds for this plan on this rank (iexa.h iexa_device_bytes)"""
    out = (C.c_int64 * 6)()
    _lib.check(m.L, m.L.iexa_device_bytes(m.h, out))
    return dict(zip(("columns", "columns_unsharded", "theta", "theta_unsharded", "programs_tables", "host_path_staging"), [int(v) for v in out]))


def jac_is_csr(m: ExaModel) -> bool:
    """The jac_coord! array is already a CSR value array (slot_order=2, iexa.h IEXA_SLOT_ORDER_JAC_ROW_SORTED)."""
    out = C.c_int32(0)
    _lib.check(m.L, m.L.iexa_jac_is_csr(m.h, C.byref(out)))
    return bool(out.value)


def jac_csr_rowptr_(m: ExaModel, rowptr):
    """0-based CSR row pointers (loc_ncon + 1 entries) of the row-sorted Jacobian layout; the column indices are jac_structure's cols."""
    b = m._buf(rowptr, m.loc_ncon + 1)
    ms, st = m._pair(b)
    _lib.check(m.L, m.L.iexa_jac_csr_rowptr(m.h, b[0], _idx_bytes(rowptr), ms, st))
    return rowptr


def hess_structure_(m: ExaModel, rows, cols):
    br, bc = m._buf(rows, m.loc_nnzh), m._buf(cols, m.loc_nnzh)
    ms, st = m._pair(br, bc)
    _lib.check(m.L, m.L.iexa_hess_structure(m.h, br[0], bc[0], _idx_bytes(rows), ms, st))
    return rows, cols


def jac_coord_(m: ExaModel, x, vals):
    bx, bv = m._buf(x, m.meta.nvar), m._buf(vals, m.loc_nnzj)
    ms, st = m._pair(bx, bv)
    _lib.check(m.L, m.L.iexa_jac_coord(m.h, bx[0], bv[0], ms, st))
    return vals


def hess_coord_(m: ExaModel, x, y, vals, obj_weight: float = 1.0):
    bx, by, bv = m._buf(x, m.meta.nvar), m._buf(y, m.loc_ncon), m._buf(vals, m.loc_nnzh)
    ms, st = m._pair(bx, by, bv)
    _lib.check(m.L, m.L.iexa_hess_coord(m.h, bx[0], by[0], float(obj_weight), bv[0], ms, st))
    return vals


def eval3_(m: ExaModel, x, y, c, jvals, hvals, obj_weight: float = 1.0):
    """cons! + jac_coord! + hess_coord! at the same (x, y) in one call (one fused kernel for device buffers)"""
    bx, by = m._buf(x, m.meta.nvar), m._buf(y, m.loc_ncon)
    bc, bj, bh = m._buf(c, m.loc_ncon), m._buf(jvals, m.loc_nnzj), m._buf(hvals, m.loc_nnzh)
    ms, st = m._pair(bx, by, bc, bj, bh)
    _lib.check(m.L, m.L.iexa_eval3(m.h, bx[0], by[0], float(obj_weight), bc[0], bj[0], bh[0], ms, st))
    return c, jvals, hvals


def jprod_(m: ExaModel, x, v, Jv):
    bx, bv, bo = m._buf(x, m.meta.nvar), m._buf(v, m.meta.nvar), m._buf(Jv, m.loc_ncon)
    ms, st = m._pair(bx, bv, bo)
    _lib.check(m.L, m.L.iexa_jprod(m.h, bx[0], bv[0], bo[0], ms, st))
    return Jv


def jtprod_(m: ExaModel, x, v, Jtv):
    bx, bv, bo = m._buf(x, m.meta.nvar), m._buf(v, m.loc_ncon), m._buf(Jtv, m.meta.nvar)
    ms, st = m._pair(bx, bv, bo)
    _lib.check(m.L, m.L.iexa_jtprod(m.h, bx[0], bv[0], bo[0], ms, st))
    return Jtv


def hprod_(m: ExaModel, x, y, v, Hv, obj_weight: float = 1.0):
    bx, by, bv, bo = m._buf(x, m.meta.nvar), m._buf(y, m.loc_ncon), m._buf(v, m.meta.nvar), m._buf(Hv, m.meta.nvar)
    ms, st = m._pair(bx, by, bv, bo)
    _lib.check(m.L, m.L.iexa_hprod(m.h, bx[0], by[0], bv[0], float(obj_weight), bo[0], ms, st))
    return Hv


def algorithmic_bytes(m: ExaModel, which: int) -> int:
    return int(m.L.iexa_algorithmic_bytes(m.h, which))


def launches_per_call(m: ExaModel, which: int) -> int:
    return int(m.L.iexa_launches_per_call(m.h, which))


# ---- pre-bound callbacks ------------------------------------------------------------------------
class BoundCall:
    """A callback with its buffers resolved once (raw pointers, memory space, stream), so that each
    call is ONE foreign call — what a Julia ``ccall`` on ``CuArray`` pointers costs.  Use when the
    solver reuses the same vectors every iteration (MadNLP and Ipopt both do)."""

    __slots__ = ("_f", "_args", "_m", "_keep")

    def __init__(self, m: ExaModel, fn, args, keep):
        self._m, self._f, self._args, self._keep = m, fn, args, keep

    def __call__(self):
        rc = self._f(*self._args)
        if rc:
            _lib.check(self._m.L, rc)


def bind(m: ExaModel, name: str, x, out, y=None, obj_weight: float = 1.0, new_x: bool = True, v=None) -> BoundCall:
    """``bind(m, "cons", x, c)``, ``bind(m, "jac_coord", x, vals)``, ``bind(m, "hess_coord", x, vals, y, σ)``,
    ``bind(m, "grad", x, g)``, ``bind(m, "jprod", x, Jv, v=v)``, ``bind(m, "jtprod", x, Jtv, v=w)``,
    ``bind(m, "hprod", x, Hv, y, σ, v=v)``.  ``new_x=False`` (host buffers only) is Ipopt's ``new_x`` flag: x is the x
    of the previous host call on this model, so the engine reuses its device copy (IEXA_MEM_HOST_SAME_X)."""
    bx, by, bv = m._buf(x, m.meta.nvar), m._buf(y), m._buf(v)
    bo = m._buf(out[0] if isinstance(out, tuple) else out)
    ms, st = m._pair(*[b for b in (bx, bo, by, bv) if b[0] is not None])
    if not new_x and ms == _lib.IEXA_MEM_HOST:
        ms = _lib.IEXA_MEM_HOST_SAME_X
    h = m.h
    L = m.L
    vp = C.c_void_p
    if name == "cons":
        return BoundCall(m, L.iexa_cons, (h, vp(bx[0]), vp(bo[0]), ms, vp(st)), (x, out))
    if name == "jac_coord":
        return BoundCall(m, L.iexa_jac_coord, (h, vp(bx[0]), vp(bo[0]), ms, vp(st)), (x, out))
    if name == "grad":
        return BoundCall(m, L.iexa_grad, (h, vp(bx[0]), vp(bo[0]), ms, vp(st)), (x, out))
    if name == "hess_coord":
        return BoundCall(m, L.iexa_hess_coord, (h, vp(bx[0]), vp(by[0]), C.c_double(obj_weight), vp(bo[0]), ms, vp(st)), (x, out, y))
    if name == "eval3":   # out = (c, jac_vals, hess_vals)
        bc, bj, bh = (m._buf(o) for o in out)
        return BoundCall(m, L.iexa_eval3, (h, vp(bx[0]), vp(by[0]), C.c_double(obj_weight), vp(bc[0]), vp(bj[0]), vp(bh[0]), ms, vp(st)), (x, out, y))
    if name == "jprod":
        return BoundCall(m, L.iexa_jprod, (h, vp(bx[0]), vp(bv[0]), vp(bo[0]), ms, vp(st)), (x, out, v))
    if name == "jtprod":
        return BoundCall(m, L.iexa_jtprod, (h, vp(bx[0]), vp(bv[0]), vp(bo[0]), ms, vp(st)), (x, out, v))
    if name == "hprod":
        return BoundCall(m, L.iexa_hprod, (h, vp(bx[0]), vp(by[0]), vp(bv[0]), C.c_double(obj_weight), vp(bo[0]), ms, vp(st)), (x, out, y, v))
    raise ValueError(name)

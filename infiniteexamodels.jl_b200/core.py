"""Host-side mirror of the ExaModels builder API that ``src/transform.jl`` drives.

``ExaCore`` records, in call order, exactly what the reference emits into an
``ExaModels.ExaCore`` (``add_var`` transform.jl:113,154; ``add_par`` :127,179; ``add_con``
:458,559,597; ``add_obj`` :614,700,741).  It is a neutral description: ``ExaModel`` hands it to
the CUDA engine through the C ABI (include/iexa.h); the test oracle reads the same object.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple, Union

import numpy as np

from .expr import (Const, DataField, IndexExpr, Node, Par, Tape, Var, as_node, lower)


# --- generated columns (device-side transcription) ----------------------------------------------------
@dataclass
class ColGen:
    """An fp iterator column described by a closed form instead of data (include/iexa.h: iexa_colgen).  The engine
    generates it ON the device; ``values(K, cols)`` is the numpy statement of the same arithmetic — what the host path
    would have uploaded, and what the oracle reads."""
    kind: int          # 1 LINSPACE, 2 LINSPACE_MID, 3 TRAPEZOID, 4 CONST
    a: float = 0.0
    b: float = 0.0
    n: int = 0
    src: Optional[str] = None   # TRAPEZOID: name of an earlier fp column of the same iterator

    def values(self, K: int, cols: Dict[str, np.ndarray]) -> np.ndarray:
        if self.kind == 1:
            return np.linspace(self.a, self.b, self.n)
        if self.kind == 2:
            pub = np.linspace(self.a, self.b, self.n)
            out = np.empty(2 * self.n - 1)
            out[0::2] = pub
            out[1::2] = 0.5 * (pub[:-1] + pub[1:])
            return out
        if self.kind == 3:
            s = cols[self.src]
            c = np.zeros_like(s)
            d = np.diff(s)
            c[:-1] += d / 2
            c[1:] += d / 2
            return c
        if self.kind == 4:
            return np.full(K, float(self.a))
        raise ValueError(self.kind)


def linspace_col(a, b, n): return ColGen(1, float(a), float(b), int(n))
def linspace_mid_col(a, b, n): return ColGen(2, float(a), float(b), int(n))
def trapezoid_col(src): return ColGen(3, src=src)
def const_col(v): return ColGen(4, float(v))


# --- iterators --------------------------------------------------------------------------------
class Itr:
    """SoA form of the reference's ``Vector{NamedTuple}`` iterators (transform.jl:31).

    A *base* iterator owns named integer columns (group_idx / i1,i2 / support indices) and named
    float columns (support values, ``d_arg`` coefficients, quadrature weight ``c``).  A *product*
    iterator (transform.jl:445,541,591,670; first factor fastest) only references its factors.
    """

    def __init__(self, K: int, ints: Optional[Dict[str, np.ndarray]] = None,
                 fps: Optional[Dict[str, np.ndarray]] = None, factors: Optional[List["Itr"]] = None):
        self.K = int(K)
        self.factors = factors
        self.ints: Dict[str, np.ndarray] = {}
        self.fps: Dict[str, np.ndarray] = {}
        self.gens: Dict[str, ColGen] = {}     # fp columns generated on the device (their numpy values are in ``fps`` too)
        self.iota: set = set()                # int columns given as None: 1..K, never materialised for the engine
        if factors is None:
            for k, v in (ints or {}).items():
                if v is None:
                    self.iota.add(k)
                    v = np.arange(1, self.K + 1)
                a = np.ascontiguousarray(v, dtype=np.int64)
                assert a.shape == (self.K,), f"int column {k}: shape {a.shape} != ({self.K},)"
                self.ints[k] = a
            for k, v in (fps or {}).items():
                if isinstance(v, ColGen):
                    self.gens[k] = v
                    v = v.values(self.K, self.fps)
                a = np.ascontiguousarray(v, dtype=np.float64)
                assert a.shape == (self.K,), f"fp column {k}: shape {a.shape} != ({self.K},)"
                self.fps[k] = a
        self._handle = {}  # per-backend iterator ids

    @staticmethod
    def empty() -> "Itr":
        """``[(;)]`` — the iterator of finite constraints / objective terms (transform.jl:440,614)."""
        return Itr(1)

    @staticmethod
    def product(factors: Sequence["Itr"]) -> "Itr":
        K = 1
        for f in factors:
            K *= f.K
        return Itr(K, factors=list(factors))

    # column names in C-ABI order: factors' columns concatenated (later names shadow earlier ones
    # like Julia's merge(); the lowering resolves a name to its LAST occurrence)
    def int_names(self) -> List[str]:
        if self.factors is None:
            return list(self.ints)
        return [n for f in self.factors for n in f.int_names()]

    def fp_names(self) -> List[str]:
        if self.factors is None:
            return list(self.fps)
        return [n for f in self.factors for n in f.fp_names()]

    def materialise(self) -> Tuple[List[np.ndarray], List[np.ndarray]]:
        """Full-length columns (what ExaModels stores per element); used by the oracle only."""
        if self.factors is None:
            return [self.ints[n] for n in self.ints], [self.fps[n] for n in self.fps]
        ic, fc = [], []
        stride = 1
        for f in self.factors:
            fi, ff = f.materialise()
            reps_inner, reps_outer = stride, self.K // (stride * f.K) if f.K else 0
            for col in fi:
                ic.append(np.tile(np.repeat(col, reps_inner), reps_outer))
            for col in ff:
                fc.append(np.tile(np.repeat(col, reps_inner), reps_outer))
            stride *= f.K
        return ic, fc

    def filtered(self, mask: np.ndarray) -> "Itr":
        """Domain restriction: ``filter(itr)`` (transform.jl:448-451) -> a materialised base iterator."""
        ic, fc = self.materialise()
        inames, fnames = self.int_names(), self.fp_names()
        mask = np.asarray(mask, dtype=bool)
        return Itr(int(mask.sum()), {n: c[mask] for n, c in zip(inames, ic)},
                   {n: c[mask] for n, c in zip(fnames, fc)})


def _resolve_names(names: List[str]) -> Tuple[List[str], Dict[str, int]]:
    """position of the LAST column carrying each name (merge() semantics)."""
    pos = {}
    for i, n in enumerate(names):
        pos[n] = i
    return names, pos


# --- variables / parameters ---------------------------------------------------------------------
class _Indexed:
    def __init__(self, offset: int, size: Tuple[int, ...]):
        self.offset = int(offset)      # 0-based offset of the block (ExaModels `offset`)
        self.size = tuple(int(s) for s in size)
        self.length = int(np.prod(self.size)) if self.size else 1

    def _index(self, key) -> IndexExpr:
        if not isinstance(key, tuple):
            key = (key,)
        if len(key) != len(self.size):
            raise IndexError(f"expected {len(self.size)} indices, got {len(key)}")
        ix = IndexExpr(self.offset + 1)
        stride = 1
        for k, n in zip(key, self.size):  # column-major, 1-based
            ix = ix + (IndexExpr.of(k) - 1) * stride
            stride *= n
        return ix


class Variable(_Indexed):
    """``ExaModels.Variable``: ``v[i, j]`` -> ``Var`` (transform.jl:250,269,310,318)."""

    def __getitem__(self, key) -> Var:
        return Var(self._index(key))

    def index(self, *key) -> int:
        """1-based x index for integer subscripts."""
        ix = self._index(tuple(key))
        assert not ix.terms
        return ix.const


class Parameter(_Indexed):
    """``ExaModels.Parameter``: ``p[i, j]`` -> theta lookup (transform.jl:324,329)."""

    def __getitem__(self, key) -> Par:
        return Par(self._index(key))


@dataclass
class GenSpec:
    is_obj: bool
    expr: Node
    itr: Itr
    tape: Tape
    lcon: float = 0.0
    ucon: float = 0.0
    row_offset: int = 0  # 0-based first row (constraints)


@dataclass
class Constraint:
    """``ExaModels.Constraint`` handle: ``offset``/``size`` as used by ``multipliers(results, con)``
    (infiniteopt_backend.jl:500-505)."""
    offset: int
    size: int
    itr: Itr


class ExaCore:
    """``ExaModels.ExaCore(; backend, minimize, concrete)`` — transform.jl:815."""

    def __init__(self, minimize: bool = True):
        self.minimize = bool(minimize)
        self.x0: List[np.ndarray] = []
        self.lvar: List[np.ndarray] = []
        self.uvar: List[np.ndarray] = []
        self.theta: List[np.ndarray] = []
        self.nvar = 0
        self.npar = 0
        self.ncon = 0
        self.gens: List[GenSpec] = []
        self.par_functions = []           # (Parameter, Tape, Itr): blocks the engine evaluates on the device
        self.var_defaults = []            # per add_var block: (n, start is 0, lvar is -inf, uvar is +inf)
        self._x0 = self._lvar = self._uvar = self._theta = None

    # -- add_var / add_par ------------------------------------------------------------------------
    def add_var(self, *dims: int, start=0.0, lvar=-np.inf, uvar=np.inf) -> Variable:
        dims = tuple(int(d) for d in dims) or (1,)
        n = int(np.prod(dims))

        def col(v):
            a = np.asarray(v, dtype=np.float64)
            if a.ndim == 0:
                return np.full(n, float(a))
            assert a.shape == dims, f"bound/start array shape {a.shape} != {dims}"
            return np.ascontiguousarray(a.reshape(-1, order="F"))  # column-major like Julia

        v = Variable(self.nvar, dims)
        self.x0.append(col(start)); self.lvar.append(col(lvar)); self.uvar.append(col(uvar))
        # blocks whose start / bounds are the engine's defaults (0, -inf, +inf) need no host array at all (model.py)
        sc = lambda a, d: np.ndim(a) == 0 and float(a) == d
        self.var_defaults.append((n, sc(start, 0.0), sc(lvar, -np.inf), sc(uvar, np.inf)))
        self.nvar += n
        self._x0 = None
        return v

    def add_par_function(self, expr, itr: "Itr") -> Parameter:
        """a parameter function (transform.jl:161-183) as an EXPRESSION over the iterator's fields: the engine evaluates the
        theta block on the device at finalize (iexa_add_par_function).  The host copy kept here (for the oracle) is the numpy
        evaluation of the same expression."""
        from .expr import eval_numpy
        tape = self._lower(expr, itr)
        _, fc = itr.materialise()
        vals = eval_numpy(tape, fc, self.theta_vec if self.npar else np.zeros(0), itr.K)
        p = self.add_par(vals)
        self.par_functions.append((p, tape, itr))
        return p

    def add_par(self, vals) -> Parameter:
        a = np.asarray(vals, dtype=np.float64)
        dims = a.shape if a.ndim else (1,)
        p = Parameter(self.npar, dims)
        self.theta.append(np.ascontiguousarray(a.reshape(-1, order="F")))
        self.npar += p.length
        self._theta = None
        return p

    # flat vectors (mutable until the model is built: transform.jl:216-231 patches them)
    def _flat(self, what="x"):
        if what == "x" and self._x0 is None:
            cat = lambda l: np.concatenate(l) if l else np.zeros(0)
            self._x0, self._lvar, self._uvar = cat(self.x0), cat(self.lvar), cat(self.uvar)
            self.x0, self.lvar, self.uvar = [self._x0], [self._lvar], [self._uvar]
        if what == "theta" and self._theta is None:
            self._theta = np.concatenate(self.theta) if self.theta else np.zeros(0)
            self.theta = [self._theta]

    @property
    def x0_vec(self): self._flat(); return self._x0
    @property
    def lvar_vec(self): self._flat(); return self._lvar
    @property
    def uvar_vec(self): self._flat(); return self._uvar
    @property
    def theta_vec(self): self._flat("theta"); return self._theta

    # -- add_con / add_obj ------------------------------------------------------------------------
    def _lower(self, expr, itr: Itr) -> Tape:
        inames, ipos = _resolve_names(itr.int_names())
        fnames, fpos = _resolve_names(itr.fp_names())
        # lower() maps names to positions; give it lists where shadowed names are unreachable
        il = [n if ipos[n] == i else f"\0shadow{i}" for i, n in enumerate(inames)]
        fl = [n if fpos[n] == i else f"\0shadow{i}" for i, n in enumerate(fnames)]
        return lower(expr, il, fl)

    def add_con(self, expr, itr: Optional[Itr] = None, lcon: float = 0.0, ucon: float = 0.0) -> Constraint:
        itr = itr or Itr.empty()
        tape = self._lower(expr, itr)
        g = GenSpec(False, as_node(expr), itr, tape, float(lcon), float(ucon), self.ncon)
        self.gens.append(g)
        self.ncon += itr.K
        return Constraint(g.row_offset, itr.K, itr)

    def add_obj(self, expr, itr: Optional[Itr] = None) -> None:
        itr = itr or Itr.empty()
        tape = self._lower(expr, itr)
        self.gens.append(GenSpec(True, as_node(expr), itr, tape))

    @property
    def cons(self) -> List[GenSpec]:
        return [g for g in self.gens if not g.is_obj]

    @property
    def objs(self) -> List[GenSpec]:
        return [g for g in self.gens if g.is_obj]
